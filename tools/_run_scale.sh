# usage: bash tools/_run_scale.sh N   — default-partition bench lines of the three large configs at N GPUs (+ dist_check at 8)
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "8" ]; then
  GNN_GRID=2x4 timeout 900 $TR --master-port 29521 tests/dist_check.py > gpurun_out/r2_dist${N}_2x4.log 2>&1; echo "dist_check 2x4 rc=$?"; grep "\[dist_check\]" gpurun_out/r2_dist${N}_2x4.log | tail -16
fi
for c in products reddit products_local; do
  timeout 900 $TR --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --config $c > gpurun_out/r2_final_${c}_${N}gpu.json 2> gpurun_out/r2_final_${c}_${N}gpu.err; echo "bench $c rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2_final_${c}_${N}gpu.json").read().strip().splitlines()[-1])
print("$c N=$N", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), {k:round(v,2) for k,v in d["breakdown_ms"].items()}, "parity", (d.get("parity") or {}).get("ok"), (d.get("parity") or {}).get("max_rel_err"), d["config"]["parallelism"][:60], d["config"].get("exchange"))
PY
done
