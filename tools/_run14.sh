set +e
export PYTHONUNBUFFERED=1
timeout 120 python tools/gemm_ncu_target.py > gpurun_out/r2b_ncu_target_plain.log 2>&1; echo "plain rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_rows_kernel|tc_tn_kernel" -s 3 -c 3 -f -o gpurun_out/r2b_prof_gemm_pair python tools/gemm_ncu_target.py > gpurun_out/r2b_ncu_gemm_pair.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2b_ncu_gemm_pair.log
ls -la gpurun_out/*.ncu-rep
