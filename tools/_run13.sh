# round 2, session 2: pair kernels with the light stage publish (default), fused bias gradient (default), staged SpMM A/B
set +e
export PYTHONUNBUFFERED=1
DEBUG_NO_TIMING=1 timeout 240 python tools/gemm_tc_debug.py > gpurun_out/r2b_gemm_debug_default.log 2>&1
echo "gemm_tc_debug default rc=$?"; grep -E "WORST|rror|Traceback|bad=[1-9]" gpurun_out/r2b_gemm_debug_default.log | tail -8
probe() { name=$1; shift; env "$@" timeout 200 python tools/gemm_probe.py > gpurun_out/r2b_gemm_probe_$name.log 2>&1; echo "probe $name rc=$?"; cat gpurun_out/r2b_gemm_probe_$name.log | tail -11; }
probe default2 X=1
probe single2 GNN_GEMM_PAIR=0
timeout 400 python tools/spmm_async_probe.py products 16 32 47 64 100 128 > gpurun_out/r2b_spmm_async_probe.log 2>&1; echo "spmm_async_probe rc=$?"; cat gpurun_out/r2b_spmm_async_probe.log | tail -30
GNN_SPMM_ASYNC=8 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "spmm or train_step or out_of_bounds" > gpurun_out/r2b_pytest_async.log 2>&1; echo "pytest async subset rc=$?"; tail -4 gpurun_out/r2b_pytest_async.log
python -m pytest tests -x -q -m gpu > gpurun_out/r2b_pytest_gpu3.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2b_pytest_gpu3.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2b_bench3_products_1gpu.json 2> gpurun_out/r2b_bench3_products_1gpu.err; echo "bench rc=$?"
GNN_SPMM_ASYNC=8 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2b_bench3_products_1gpu_async.json 2> gpurun_out/r2b_bench3_products_1gpu_async.err; echo "bench async rc=$?"
python - <<'PY'
import json
for n in ("r2b_bench3_products_1gpu", "r2b_bench3_products_1gpu_async"):
    try:
        d=json.loads(open("gpurun_out/%s.json" % n).read().strip().splitlines()[-1])
        print(n, round(d["value"],3), "e2e", round(d["e2e"]["value"],3), {k:round(v,3) for k,v in d["breakdown_ms"].items()}, "parity", (d.get("parity") or {}).get("ok"), (d.get("parity") or {}).get("max_rel_err"), "gemm_roofline", d.get("gemm_roofline",{}).get("frac"), {k: round(v["frac"],3) for k,v in d["roofline"]["all_aggregations_of_a_step"]["by_width"].items()}, d["clocks"])
    except Exception as e:
        print(n, "unreadable", e)
PY
