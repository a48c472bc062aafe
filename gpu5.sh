cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export PATH=$PATH:/usr/local/cuda/bin
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3000 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
# launch list of one mid-solve window (iteration kernels only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(?!lu_|trsm|gemm_sub|set_perm|transpose|zero_pad|gather_nucleus_matrix).*' -s 60 -c 300 --csv --log-file gpurun_out/launches_c2.csv python tests/ncu_target.py c2 24 > gpurun_out/ncu1.log 2>&1; tail -2 gpurun_out/ncu1.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:price_kernel -s 6 -c 2 -o gpurun_out/prof_price -f python tests/ncu_target.py c2 12 > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemv_rows -s 12 -c 2 -o gpurun_out/prof_gemv -f python tests/ncu_target.py c2 12 > gpurun_out/ncu3.log 2>&1; tail -2 gpurun_out/ncu3.log
ls -la gpurun_out
