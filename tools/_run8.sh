python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_gpu2.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench1d.json 2> gpurun_out/r2_bench1d.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench1d.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench1d.json").read().strip().splitlines()[-1])
print(round(d["value"],2), {k:round(v,2) for k,v in d["breakdown_ms"].items()}, "build", d["config"]["structure_build_ms"], d["config"]["structure_build_warm_ms"])
print("gemm_roofline", d["gemm_roofline"]); print("parity", d["parity"]["ok"], d["parity"]["max_rel_err"]); print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PY
for c in arxiv pubmed cora reddit; do python tools/spmm_probe.py $c 1 2 > gpurun_out/r2_variant_probe_$c.log 2>&1; cat gpurun_out/r2_variant_probe_$c.log; done
