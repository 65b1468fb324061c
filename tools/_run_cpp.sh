python -m pytest tests/test_host_cpp.py -q -m gpu > gpurun_out/r2_pytest_cpp3.log 2>&1; echo "cpp tests rc=$?"; tail -3 gpurun_out/r2_pytest_cpp3.log
gnn.cpp_b200/host/gcn_main --config products --epochs 8 --lr 0.01 > gpurun_out/r2_gcn_main_products_1gpu.log 2>&1; echo "gcn_main 1 rc=$?"; cat gpurun_out/r2_gcn_main_products_1gpu.log
GNN_GEMM_PRECISION=0 gnn.cpp_b200/host/gcn_main --config products --epochs 5 --lr 0.01 > gpurun_out/r2_gcn_main_products_1gpu_fma.log 2>&1; cat gpurun_out/r2_gcn_main_products_1gpu_fma.log
gnn.cpp_b200/host/gcn_main --config arxiv --epochs 6 --lr 0.01 | tail -3
