set +e
export PYTHONUNBUFFERED=1
python bench.py > gpurun_out/r2b_bench_products_1gpu.json 2> gpurun_out/r2b_bench_products_1gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2b_bench_products_1gpu.json").read().strip().splitlines()[-1])
print("products", round(d["value"],3), "e2e", round(d["e2e"]["value"],3), {k:round(v,3) for k,v in d["breakdown_ms"].items()}, "build", round(d["config"]["structure_build_ms"],1), round(d["config"]["structure_build_warm_ms"],1),
  "parity", d["parity"]["ok"], d["parity"]["max_rel_err"], "cpu", (d.get("cpu_baseline") or {}).get("value"), "roof", round(d["roofline"]["frac"],3), "gemm", d["gemm_roofline"]["frac"], "launches", d.get("gpu_launches"), d["clocks"])
PY
timeout 100 python tools/gemm_probe.py > gpurun_out/r2b_gemm_probe_final2.log 2>&1; tail -11 gpurun_out/r2b_gemm_probe_final2.log
