python -m pytest tests/test_host_cpp.py -q -m gpu > gpurun_out/r2_pytest_cpp.log 2>&1; echo "cpp tests rc=$?"; tail -3 gpurun_out/r2_pytest_cpp.log
CUDA_VISIBLE_DEVICES=0 gnn.cpp_b200/host/gcn_main --config products --epochs 6 --lr 0.01 > gpurun_out/r2_gcn_main_products_1gpu.log 2>&1; echo "gcn_main 1 rc=$?"; cat gpurun_out/r2_gcn_main_products_1gpu.log
gnn.cpp_b200/host/gcn_main --gpus 2 --config products --epochs 6 --lr 0.01 > gpurun_out/r2_gcn_main_products_2gpu.log 2>&1; echo "gcn_main 2 rc=$?"; cat gpurun_out/r2_gcn_main_products_2gpu.log
bash tools/_run_scale.sh 2
