# round 2, session 2: lo of the truncating split rounded to nearest (bit-pattern + 0x1000) — accuracy and parity check
set +e
export PYTHONUNBUFFERED=1
DEBUG_NO_TIMING=1 timeout 240 python tools/gemm_tc_debug.py > gpurun_out/r2b_gemm_debug_trunc_rn.log 2>&1
echo "gemm_tc_debug rc=$?"; grep -E "M=150000|M=70000|WORST|rror|Traceback" gpurun_out/r2b_gemm_debug_trunc_rn.log | tail -12
python -m pytest tests/test_gpu_parity_fullsize.py tests/test_gpu_parity.py -x -q -m gpu -k "fullsize or gemm or train_step or fused or tensor_core or medium" > gpurun_out/r2b_pytest_trunc_rn.log 2>&1; echo "pytest subset rc=$?"; tail -4 gpurun_out/r2b_pytest_trunc_rn.log
cp gpurun_out/parity_report.jsonl gpurun_out/r2b_parity_report_trunc_rn.jsonl 2>/dev/null
for c in arxiv products; do python bench.py --config $c --steps 10 --warmup 3 --no-cpu > gpurun_out/r2b_bench_${c}_trunc_rn.json 2> gpurun_out/r2b_bench_${c}_trunc_rn.err; echo "bench $c rc=$?"; done
python - <<'PY'
import json
for c in ("arxiv","products"):
    d=json.loads(open("gpurun_out/r2b_bench_%s_trunc_rn.json"%c).read().strip().splitlines()[-1])
    print(c, round(d["value"],3), {k:round(v,3) for k,v in d["breakdown_ms"].items()}, "parity", d["parity"]["ok"], d["parity"]["max_rel_err"], {k:float("%.2e"%v) for k,v in d["parity"]["rel_err"].items()})
PY
