cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export PATH=$PATH:/usr/local/cuda/bin
( time timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > gpurun_out/r2_tests.log 2>&1
tail -8 gpurun_out/r2_tests.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(?!lu_|trsm|gemm_sub|set_perm|transpose|zero_pad|gather_nucleus_matrix|cutlass).*' -s 10300 -c 50 --csv --log-file gpurun_out/r2_launches_c2.csv python tests/ncu_target.py c2 1012 > gpurun_out/ncuA.log 2>&1; tail -1 gpurun_out/ncuA.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(lu_|trsm|gemm_sub|set_perm|transpose|zero_pad|gather_nucleus_matrix|cutlass).*' -c 4000 --csv --log-file gpurun_out/r2_launches_refactor.csv python tests/ncu_target.py c2 2 > gpurun_out/ncuB.log 2>&1; tail -1 gpurun_out/ncuB.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemv_rows -s 12 -c 2 -f -o gpurun_out/r2_prof_gemv python tests/ncu_target.py c2 12 > gpurun_out/ncuC.log 2>&1; tail -1 gpurun_out/ncuC.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:price_ldg -s 6 -c 1 -f -o gpurun_out/r2_prof_price python tests/ncu_target.py c2 12 > gpurun_out/ncuD.log 2>&1; tail -1 gpurun_out/ncuD.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:row_pass -s 6 -c 1 -f -o gpurun_out/r2_prof_rowpass python tests/ncu_target.py c2 12 > gpurun_out/ncuE.log 2>&1; tail -1 gpurun_out/ncuE.log
for f in gemv price rowpass; do ncu -i gpurun_out/r2_prof_$f.ncu-rep --page raw --csv > gpurun_out/r2_ncu_$f.csv 2>/dev/null; done
CLPB_REFACTOR_TRACE=1 timeout 100 python tests/refactor_probe.py 50 200 2>&1 | tail -3
CLPB_NO_CLUSTER_PANEL=1 CLPB_REFACTOR_TRACE=1 timeout 100 python tests/refactor_probe.py 50 200 2>&1 | tail -2
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; python - <<'PY'
import json
try:
    r=json.load(open('gpurun_out/r2_bench_n1.json'))
    print({k:r[k] for k in ('value','ms_per_step','timed_iterations','nucleus_size')})
    print(r['roofline']['all']); print(r['roofline']['phase_us_per_iteration'], r['roofline']['refactor_ms_total'])
    print(r['e2e']['value'], r['cpu_baseline']['value'], r['cpu_baseline'].get('others'))
    print(r.get('wall_to_optimal_s'), r.get('optimal')); print(r.get('objective_after_window'))
except Exception as e: print('bench parse fail', e)
PY
tail -2 gpurun_out/r2_bench_n1.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
