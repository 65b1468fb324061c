# round 2, session 2: CTA pairs in both tensor-core kernels + bias gradient fused into the NN epilogue
set +e
export PYTHONUNBUFFERED=1
dbg() { # name, env...
  name=$1; shift
  env "$@" DEBUG_NO_TIMING=1 timeout 240 python tools/gemm_tc_debug.py > gpurun_out/r2b_gemm_debug_$name.log 2>&1
  rc=$?; echo "gemm_tc_debug $name rc=$rc"; grep -E "rel_err|WORST|rror|Traceback" gpurun_out/r2b_gemm_debug_$name.log | awk '{print "   " $0}' | tail -48
  return $rc
}
probe() { name=$1; shift; env "$@" timeout 200 python tools/gemm_probe.py > gpurun_out/r2b_gemm_probe_$name.log 2>&1; echo "probe $name rc=$?"; cat gpurun_out/r2b_gemm_probe_$name.log | tail -11; }
dbg pair3 GNN_GEMM_PAIR=3; RC3=$?
if [ $RC3 -ne 0 ]; then dbg pair1 GNN_GEMM_PAIR=1; RC1=$?; else RC1=0; fi
probe single GNN_GEMM_PAIR=0
if [ $RC3 -eq 0 ]; then
  export GNN_GEMM_PAIR=3
  probe pair3 GNN_GEMM_PAIR=3
  probe pair3_bk16 GNN_GEMM_PAIR=3 GNN_GEMM_BK=16
  probe pair3_light GNN_GEMM_PAIR=3 GNN_GEMM_DEBUG=8
elif [ $RC1 -eq 0 ]; then
  export GNN_GEMM_PAIR=1
  probe pair1 GNN_GEMM_PAIR=1
  probe pair1_bk16 GNN_GEMM_PAIR=1 GNN_GEMM_BK=16
  probe pair1_light GNN_GEMM_PAIR=1 GNN_GEMM_DEBUG=8
else
  export GNN_GEMM_PAIR=0
fi
export GNN_FUSED_BIAS_GRAD=1
echo "suite runs with GNN_GEMM_PAIR=${GNN_GEMM_PAIR:-default} GNN_FUSED_BIAS_GRAD=1"
python -m pytest tests -x -q -m gpu > gpurun_out/r2b_pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2b_pytest_gpu2.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2b_bench2_products_1gpu.json 2> gpurun_out/r2b_bench2_products_1gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2b_bench2_products_1gpu.json").read().strip().splitlines()[-1])
print("products", round(d["value"],3), "e2e", round(d["e2e"]["value"],3), {k:round(v,3) for k,v in d["breakdown_ms"].items()}, "parity", (d.get("parity") or {}).get("ok"), (d.get("parity") or {}).get("max_rel_err"), "gemm_roofline", d.get("gemm_roofline",{}).get("frac"), d["clocks"])
PY
