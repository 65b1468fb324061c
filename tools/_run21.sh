# the GPU tests the last subset run (tools/_run19.sh) did not cover, against the final library
set +e
export PYTHONUNBUFFERED=1
timeout 120 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu -k "not (gemm or train_step or fused or tensor_core or medium)" > gpurun_out/r2b_pytest_rest.log 2>&1; echo "pytest rest rc=$?"; tail -5 gpurun_out/r2b_pytest_rest.log
