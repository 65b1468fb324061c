# round 2, session 2: validation of the final library (pairs + fused bias gradient + converter/stage-counter changes)
set +e
export PYTHONUNBUFFERED=1
DEBUG_NO_TIMING=1 timeout 240 python tools/gemm_tc_debug.py > gpurun_out/r2b_gemm_debug_final.log 2>&1
echo "gemm_tc_debug rc=$?"; grep -E "WORST|rror|Traceback|bad=[1-9]" gpurun_out/r2b_gemm_debug_final.log | tail -8
timeout 200 python tools/gemm_probe.py > gpurun_out/r2b_gemm_probe_final.log 2>&1; echo "probe rc=$?"; tail -11 gpurun_out/r2b_gemm_probe_final.log
python -m pytest tests -x -q -m gpu > gpurun_out/r2b_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2b_pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2b_smoke.log
python bench.py > gpurun_out/r2b_bench_products_1gpu.json 2> gpurun_out/r2b_bench_products_1gpu.err; echo "bench rc=$?"
for c in cora pubmed arxiv; do python bench.py --config $c --no-cpu > gpurun_out/r2b_bench_${c}_1gpu.json 2> gpurun_out/r2b_bench_${c}_1gpu.err; echo "bench $c rc=$?"; done
python - <<'PY'
import json
for c in ["products","cora","pubmed","arxiv"]:
    try:
        d=json.loads(open("gpurun_out/r2b_bench_%s_1gpu.json"%c).read().strip().splitlines()[-1])
        print(c, round(d["value"],3), "e2e", round(d["e2e"]["value"],3), {k:round(v,3) for k,v in d["breakdown_ms"].items()}, "build", round(d["config"]["structure_build_ms"],1), round(d["config"]["structure_build_warm_ms"],1),
          "parity", (d.get("parity") or {}).get("ok"), (d.get("parity") or {}).get("max_rel_err"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "roof", round(d["roofline"]["frac"],3), "gemm", (d.get("gemm_roofline") or {}).get("frac"), "launches", d.get("gpu_launches"), d["clocks"])
    except Exception as e:
        print(c, "unreadable", e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2b_launches_products.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2b_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_rows_kernel|tc_tn_kernel" -s 3 -c 3 -f -o gpurun_out/r2b_prof_gemm_pair_final python tools/gemm_ncu_target.py > gpurun_out/r2b_ncu_gemm_pair_final.log 2>&1; echo "ncu full rc=$?"
