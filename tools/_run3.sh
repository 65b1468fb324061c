TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for g in row 1x2 2x1; do
  GNN_GRID=$g timeout 600 $TR --master-port 29521 tests/dist_check.py > gpurun_out/r2_dist2_$g.log 2>&1; echo "dist_check $g rc=$?"; grep dist_check gpurun_out/r2_dist2_$g.log | tail -12
done
for g in row 1x2; do
  GNN_GRID=$g timeout 600 $TR --master-port 29522 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench2_$g.json 2> gpurun_out/r2_bench2_$g.err; echo "bench $g rc=$?"; tail -2 gpurun_out/r2_bench2_$g.err
done
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench1b.json 2> gpurun_out/r2_bench1b.err; echo "bench1 rc=$?"
