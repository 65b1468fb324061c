# round 2, session 2: CTA-pair (cta_group::2) dense transforms — correctness, A/B timing, then the whole GPU suite
set +e
export PYTHONUNBUFFERED=1
dbg() { # name, env...
  name=$1; shift
  env "$@" DEBUG_NO_TIMING=1 timeout 240 python tools/gemm_tc_debug.py > gpurun_out/r2b_gemm_debug_$name.log 2>&1
  rc=$?; echo "gemm_tc_debug $name rc=$rc"; grep -E "rel_err|WORST|rror|Traceback" gpurun_out/r2b_gemm_debug_$name.log | awk '{print "   " $0}' | tail -45
  return $rc
}
dbg pair GNN_GEMM_DEBUG_PRINT=1; PAIR_RC=$?
if [ $PAIR_RC -ne 0 ]; then dbg pair_alloc2 GNN_GEMM_DEBUG=16; ALLOC2_RC=$?; else ALLOC2_RC=1; fi
probe() { name=$1; shift; env "$@" timeout 200 python tools/gemm_probe.py > gpurun_out/r2b_gemm_probe_$name.log 2>&1; echo "probe $name rc=$?"; cat gpurun_out/r2b_gemm_probe_$name.log | tail -11; }
probe single GNN_GEMM_PAIR=0
if [ $PAIR_RC -eq 0 ]; then
  probe pair GNN_GEMM_PAIR=1
  probe pair_bk16 GNN_GEMM_PAIR=1 GNN_GEMM_BK=16
  probe pair_light GNN_GEMM_DEBUG=8
elif [ $ALLOC2_RC -eq 0 ]; then
  probe pair_alloc2 GNN_GEMM_DEBUG=16
  export GNN_GEMM_DEBUG=16
else
  export GNN_GEMM_PAIR=0
fi
echo "suite runs with GNN_GEMM_PAIR=${GNN_GEMM_PAIR:-default} GNN_GEMM_DEBUG=${GNN_GEMM_DEBUG:-}"
python -m pytest tests -x -q -m gpu > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2b_pytest_gpu.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2b_bench_products_1gpu.json 2> gpurun_out/r2b_bench_products_1gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2b_bench_products_1gpu.json").read().strip().splitlines()[-1])
print("products", round(d["value"],3), "e2e", round(d["e2e"]["value"],3), {k:round(v,3) for k,v in d["breakdown_ms"].items()}, "parity", (d.get("parity") or {}).get("ok"), (d.get("parity") or {}).get("max_rel_err"), "gemm_roofline", d.get("gemm_roofline",{}).get("frac"))
PY
timeout 300 gnn.cpp_b200/host/gcn_main --config products --epochs 8 --lr 0.01 > gpurun_out/r2b_gcn_main_products_1gpu.log 2>&1; echo "gcn_main rc=$?"; cat gpurun_out/r2b_gcn_main_products_1gpu.log
