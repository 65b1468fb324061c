TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
CUDA_VISIBLE_DEVICES=0 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "partition_arrays or spmm_fwd_bwd" > gpurun_out/r2_pytest_part.log 2>&1; tail -3 gpurun_out/r2_pytest_part.log
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench1c.json 2> gpurun_out/r2_bench1c.err; echo "bench1 rc=$?"
GNN_GRID=2x2 GNN_DIST_BIG=1 timeout 600 $TR --master-port 29521 tests/dist_check.py > gpurun_out/r2_dist4_2x2.log 2>&1; echo "dist_check 2x2 rc=$?"; grep dist_check gpurun_out/r2_dist4_2x2.log | tail -12
for g in row 2x2 1x4; do
  GNN_GRID=$g timeout 600 $TR --master-port 29522 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench4_$g.json 2> gpurun_out/r2_bench4_$g.err; echo "bench $g rc=$?"; tail -2 gpurun_out/r2_bench4_$g.err
done
