# round 2, session 2: truncating hi/lo split (GNN_GEMM_SPLIT=1) — does the tensor core ignore the low 13 bits? accuracy, speed, suite
set +e
export PYTHONUNBUFFERED=1
GNN_GEMM_SPLIT=1 DEBUG_NO_TIMING=1 timeout 240 python tools/gemm_tc_debug.py > gpurun_out/r2b_gemm_debug_trunc.log 2>&1
RC=$?; echo "gemm_tc_debug trunc rc=$RC"; grep -E "rel_err|WORST|rror|Traceback" gpurun_out/r2b_gemm_debug_trunc.log | grep -E "M=150000|M=70000|M=4099|WORST|rror|Traceback" | tail -24
probe() { name=$1; shift; env "$@" timeout 200 python tools/gemm_probe.py > gpurun_out/r2b_gemm_probe_$name.log 2>&1; echo "probe $name rc=$?"; tail -11 gpurun_out/r2b_gemm_probe_$name.log; }
probe rna X=1
probe trunc GNN_GEMM_SPLIT=1
if [ $RC -eq 0 ]; then
  export GNN_GEMM_SPLIT=1
  python -m pytest tests -x -q -m gpu > gpurun_out/r2b_pytest_gpu_trunc.log 2>&1; echo "pytest (trunc) rc=$?"; tail -6 gpurun_out/r2b_pytest_gpu_trunc.log
  cp gpurun_out/parity_report.jsonl gpurun_out/r2b_parity_report_trunc.jsonl 2>/dev/null
  python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2b_bench_products_1gpu_trunc.json 2> gpurun_out/r2b_bench_products_1gpu_trunc.err; echo "bench rc=$?"
  python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2b_bench_products_1gpu_trunc.json").read().strip().splitlines()[-1])
print("products trunc", round(d["value"],3), "e2e", round(d["e2e"]["value"],3), {k:round(v,3) for k,v in d["breakdown_ms"].items()}, "parity", d["parity"]["ok"], d["parity"]["max_rel_err"], d["parity"]["rel_err"], "gemm", d["gemm_roofline"]["frac"], d["clocks"])
PY
fi
