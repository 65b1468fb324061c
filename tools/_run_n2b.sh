python tools/nvlink_debug.py
python -m pytest tests/test_host_cpp.py -q -m gpu > gpurun_out/r2_pytest_cpp2.log 2>&1; echo "cpp tests rc=$?"; tail -3 gpurun_out/r2_pytest_cpp2.log
CUDA_VISIBLE_DEVICES=0 gnn.cpp_b200/host/gcn_main --config products --epochs 8 --lr 0.01 > gpurun_out/r2_gcn_main_products_1gpu.log 2>&1; echo "gcn_main 1 rc=$?"; cat gpurun_out/r2_gcn_main_products_1gpu.log
gnn.cpp_b200/host/gcn_main --gpus 2 --config products --epochs 8 --lr 0.01 > gpurun_out/r2_gcn_main_products_2gpu.log 2>&1; echo "gcn_main 2 rc=$?"; cat gpurun_out/r2_gcn_main_products_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_final_products_2gpu.json 2> gpurun_out/r2_final_products_2gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_final_products_2gpu.json").read().strip().splitlines()[-1])
print(round(d["value"],2), d["nvlink"], d["parity"]["ok"])
PY
