"""B200-native GCN hot path behind the walexi/gnn.cpp C++ surface."""
