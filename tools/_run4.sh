W="12 16 24 32 48 64 100 128 256"
export PROBE_SORTED=0
PROBE_TAG=" lib=r1" GNN_LIB=$PWD/tools/_variants/libgnn_b200_r1.so python tools/spmm_width_probe.py products $W
PROBE_TAG=" lib=occ4" GNN_LIB=$PWD/tools/_variants/libgnn_b200_occ4.so python tools/spmm_width_probe.py products $W
PROBE_TAG=" lib=occ5" GNN_LIB=$PWD/tools/_variants/libgnn_b200_occ5.so python tools/spmm_width_probe.py products $W
PROBE_TAG=" lib=occ6" python tools/spmm_width_probe.py products $W
PROBE_TAG=" lib=occ6+alt" GNN_SPMM_ALT=1 python tools/spmm_width_probe.py products 12 24 48 100
PROBE_TAG=" lib=occ4+alt" GNN_SPMM_ALT=1 GNN_LIB=$PWD/tools/_variants/libgnn_b200_occ4.so python tools/spmm_width_probe.py products 12 24 48 100
PROBE_TAG=" lib=r1 again" GNN_LIB=$PWD/tools/_variants/libgnn_b200_r1.so python tools/spmm_width_probe.py products 48 100 256
