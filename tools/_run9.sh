TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
for g in 1x4 4x1 2x2; do
  GNN_GRID=$g GNN_DIST_BIG=0 timeout 600 $TR --master-port 29521 tests/dist_check.py > gpurun_out/r2_dist4b_$g.log 2>&1; echo "dist_check $g rc=$?"; grep dist_check gpurun_out/r2_dist4b_$g.log | tail -14
done
python -m pytest tests/test_host_cpp.py -q -m gpu -k "multi_gpu" > gpurun_out/r2_pytest_cpp_multi.log 2>&1; echo "cpp multi rc=$?"; tail -4 gpurun_out/r2_pytest_cpp_multi.log
run_bench() { # name, env..., config
  name=$1; shift
  env "$@" timeout 600 $TR --master-port 29522 bench.py --gpus 4 --steps 8 --warmup 3 --no-cpu --config ${BENCH_CFG:-products} > gpurun_out/r2_bench4_$name.json 2> gpurun_out/r2_bench4_$name.err; echo "bench $name rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench4_$name.json").read().strip().splitlines()[-1])
print("$name", round(d["value"],2), {k:round(v,2) for k,v in d["breakdown_ms"].items()}, "parity", (d.get("parity") or {}).get("ok"), d["config"].get("exchange"))
PY
}
run_bench prod_2x2 GNN_GRID=2x2 BENCH_CFG=products
run_bench prod_1x4 GNN_GRID=1x4 BENCH_CFG=products
export BENCH_CFG=products_local
run_bench local_4x1 GNN_GRID=4x1
run_bench local_4x1_nohalo GNN_GRID=4x1 GNN_HALO=0 GNN_SPLIT=0
run_bench local_2x2 GNN_GRID=2x2
run_bench local_row GNN_GRID=row
CUDA_VISIBLE_DEVICES=0 python bench.py --steps 8 --warmup 3 --no-cpu --config products_local > gpurun_out/r2_bench1_local.json 2> gpurun_out/r2_bench1_local.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench1_local.json").read().strip().splitlines()[-1])
print("local 1 GPU", round(d["value"],2), {k:round(v,2) for k,v in d["breakdown_ms"].items()})
PY
export PROBE_SORTED=0
for v in 1 2; do CUDA_VISIBLE_DEVICES=0 PROBE_VARIANT=$v PROBE_TAG=" variant=$v" python tools/spmm_width_probe.py products 8 16 32 48 64 100 128; done
