cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export PATH=$PATH:/usr/local/cuda/bin
timeout 600 ncu --set full --clock-control none --import-source on -k regex:price_tma -s 6 -c 1 -o gpurun_out/prof_price -f python tests/ncu_target.py c2 12 > gpurun_out/ncu2.log 2>&1; tail -1 gpurun_out/ncu2.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(lu_|trsm|gemm_sub|set_perm|transpose|zero_pad|gather_nucleus_matrix).*' -c 4000 --csv --log-file gpurun_out/launches_refactor.csv python tests/ncu_target.py c2 2 > gpurun_out/ncu4.log 2>&1; tail -1 gpurun_out/ncu4.log
