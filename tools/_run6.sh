TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for c in 64 148 296 592; do
  GNN_PEER_CTAS=$c GNN_GRID=1x2 timeout 600 $TR --master-port 29522 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu > gpurun_out/r2_bench2_1x2_c$c.json 2> gpurun_out/r2_bench2_1x2_c$c.err; echo "bench ctas=$c rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench2_1x2_c$c.json").read().strip().splitlines()[-1])
print("ctas=$c", round(d["value"],2), {k:round(v,2) for k,v in d["breakdown_ms"].items()})
PY
done
