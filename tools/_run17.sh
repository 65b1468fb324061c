# round 2, session 2: 2-GPU check of the final library (CTA-pair GEMMs + truncating split inside the partitioned trainers)
set +e
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for g in 1x2 row; do
  GNN_GRID=$g timeout 500 $TR --master-port 29521 tests/dist_check.py > gpurun_out/r2b_dist2_$g.log 2>&1; echo "dist_check $g rc=$?"; grep -E "dist_check|rror|Traceback" gpurun_out/r2b_dist2_$g.log | tail -8
done
timeout 600 $TR --master-port 29522 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu > gpurun_out/r2b_bench_products_2gpu.json 2> gpurun_out/r2b_bench_products_2gpu.err; echo "bench 2 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2b_bench_products_2gpu.json").read().strip().splitlines()[-1])
print("products 2 GPUs", round(d["value"],2), {k:round(v,2) for k,v in d["breakdown_ms"].items()}, "parity", d["parity"]["ok"], d["parity"]["max_rel_err"], d["config"]["parallelism"][:60])
PY
