# round 2, session 2: final 1-GPU bench line (full defaults, CPU baseline) + launch list + ncu of the final GEMM kernels
set +e
export PYTHONUNBUFFERED=1
python bench.py > gpurun_out/r2b_bench_products_1gpu.json 2> gpurun_out/r2b_bench_products_1gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2b_bench_products_1gpu.json").read().strip().splitlines()[-1])
print("products", round(d["value"],3), "e2e", round(d["e2e"]["value"],3), {k:round(v,3) for k,v in d["breakdown_ms"].items()}, "build", round(d["config"]["structure_build_ms"],1), round(d["config"]["structure_build_warm_ms"],1),
  "parity", d["parity"]["ok"], d["parity"]["max_rel_err"], "cpu", (d.get("cpu_baseline") or {}).get("value"), "roof", round(d["roofline"]["frac"],3), {k: round(v["frac"],3) for k,v in d["roofline"]["all_aggregations_of_a_step"]["by_width"].items()}, "gemm", d["gemm_roofline"]["frac"], "launches", d.get("gpu_launches"), d["clocks"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2b_launches_products.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2b_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"tc_rows_kernel|tc_tn_kernel" -s 3 -c 3 -f -o gpurun_out/r2b_prof_gemm_pair_final python tools/gemm_ncu_target.py > gpurun_out/r2b_ncu_gemm_pair_final.log 2>&1; echo "ncu full rc=$?"
python bench.py --config arxiv --no-cpu > gpurun_out/r2b_bench_arxiv_1gpu.json 2> gpurun_out/r2b_bench_arxiv_1gpu.err; echo "bench arxiv rc=$?"
