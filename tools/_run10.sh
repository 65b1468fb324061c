python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu3.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest_gpu3.log
python bench.py > gpurun_out/r2_bench_products_1gpu.json 2> gpurun_out/r2_bench_products_1gpu.err; echo "bench rc=$?"
for c in cora pubmed arxiv; do python bench.py --config $c > gpurun_out/r2_bench_${c}_1gpu.json 2> gpurun_out/r2_bench_${c}_1gpu.err; echo "bench $c rc=$?"; done
python bench.py --config reddit --no-cpu > gpurun_out/r2_bench_reddit_1gpu.json 2> gpurun_out/r2_bench_reddit_1gpu.err; echo "bench reddit rc=$?"
python - <<'PY'
import json
for c in ["products","cora","pubmed","arxiv","reddit"]:
    d=json.loads(open("gpurun_out/r2_bench_%s_1gpu.json"%c).read().strip().splitlines()[-1])
    print(c, round(d["value"],3), "e2e", round(d["e2e"]["value"],3), {k:round(v,3) for k,v in d["breakdown_ms"].items()}, "build", round(d["config"]["structure_build_ms"],1), round(d["config"]["structure_build_warm_ms"],1),
          "parity", (d.get("parity") or {}).get("ok"), (d.get("parity") or {}).get("max_rel_err"), "cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("kind"), "roof", round(d["roofline"]["frac"],3), d["roofline"]["kernel"][:40])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_products.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"spmm_merge_kernel" -s 15 -c 5 -o gpurun_out/r2_prof_spmm python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_ncu_spmm.log 2>&1; echo "ncu spmm rc=$?"
