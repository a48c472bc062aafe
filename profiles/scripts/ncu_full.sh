cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export PATH=$PATH:/usr/local/cuda/bin
timeout 300 ncu --set full --clock-control none --import-source on -k regex:price_tma -s 6 -c 1 -f -o gpurun_out/prof_price_v6 python tests/ncu_target.py c2 12 > gpurun_out/ncu18a.log 2>&1; tail -1 gpurun_out/ncu18a.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:row_pass -s 6 -c 1 -f -o gpurun_out/prof_rowpass_v6 python tests/ncu_target.py c2 12 > gpurun_out/ncu18b.log 2>&1; tail -1 gpurun_out/ncu18b.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemv_rows -s 12 -c 2 -f -o gpurun_out/prof_gemv_v6 python tests/ncu_target.py c2 12 > gpurun_out/ncu18c.log 2>&1; tail -1 gpurun_out/ncu18c.log
ls -la gpurun_out/*.ncu-rep
