cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export PATH=$PATH:/usr/local/cuda/bin
( time timeout 400 python -m pytest tests -m gpu -q --timeout 120 -x 2>&1 | tail -6 ) > gpurun_out/tF.log 2>&1
tail -6 gpurun_out/tF.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(?!lu_|trsm|gemm_sub|set_perm|transpose|zero_pad|gather_nucleus_matrix).*' -s 10300 -c 50 --csv --log-file gpurun_out/launchesF.csv python tests/ncu_target.py c2 1012 > gpurun_out/ncuF.log 2>&1; tail -1 gpurun_out/ncuF.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:price_ldg -s 6 -c 1 -f -o gpurun_out/prof_price_v9 python tests/ncu_target.py c2 12 > gpurun_out/ncuFa.log 2>&1; tail -1 gpurun_out/ncuFa.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemv_rows -s 12 -c 2 -f -o gpurun_out/prof_gemv_v9 python tests/ncu_target.py c2 12 > gpurun_out/ncuFb.log 2>&1; tail -1 gpurun_out/ncuFb.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:row_pass -s 6 -c 1 -f -o gpurun_out/prof_rowpass_v9 python tests/ncu_target.py c2 12 > gpurun_out/ncuFc.log 2>&1; tail -1 gpurun_out/ncuFc.log
timeout 400 python bench.py --steps 4 --warmup 3 > gpurun_out/benchF.json 2> gpurun_out/benchF.err; python - <<'PY'
import json
try:
    r=json.load(open('gpurun_out/benchF.json'))
    print({k:r[k] for k in ('value','ms_per_step')}, r['config']['timed_iterations'], r['config']['nucleus_size'])
    print(r['roofline']['all']); print(r['roofline']['phase_us_per_iteration'], r['roofline']['refactor_ms_total'])
    print(r['e2e']['value'], r['cpu_baseline']['value'])
except Exception as e: print('bench parse fail', e)
PY
tail -2 gpurun_out/benchF.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
