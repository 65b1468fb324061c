TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
GNN_GRID=2x4 timeout 600 $TR --master-port 29521 tests/dist_check.py > gpurun_out/r2_dist8_2x4.log 2>&1; echo "dist_check 2x4 rc=$?"; grep dist_check gpurun_out/r2_dist8_2x4.log | tail -12
for g in 2x4 1x8 4x2; do
  GNN_GRID=$g timeout 600 $TR --master-port 29522 bench.py --gpus 8 --steps 8 --warmup 3 --no-cpu > gpurun_out/r2_bench8_$g.json 2> gpurun_out/r2_bench8_$g.err; echo "bench $g rc=$?"; tail -2 gpurun_out/r2_bench8_$g.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench8_$g.json").read().strip().splitlines()[-1])
print("$g", round(d["value"],2), {k:round(v,2) for k,v in d["breakdown_ms"].items()}, "parity", d["parity"]["ok"], d["parity"]["max_rel_err"])
PY
done
GNN_GRID=2x4 timeout 600 $TR --master-port 29523 bench.py --gpus 8 --steps 8 --warmup 3 --no-cpu --config reddit > gpurun_out/r2_bench8_reddit_2x4.json 2> gpurun_out/r2_bench8_reddit_2x4.err; echo "reddit rc=$?"
