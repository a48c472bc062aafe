cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > gpurun_out/r2_tests_final.log 2>&1
tail -8 gpurun_out/r2_tests_final.log
timeout 400 python tests/fullsize_probe.py c5 c4 > gpurun_out/r2_fullsize_final.out 2> gpurun_out/r2_fullsize_final.err; cut -c1-420 gpurun_out/r2_fullsize_final.out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err; python - <<'PY'
import json
try:
    r=json.load(open('gpurun_out/r2_bench_n1_final.json'))
    print({k:r[k] for k in ('value','ms_per_step','timed_iterations','nucleus_size')})
    print({q:(round(v['GBps']),round(v['frac'],3)) for q,v in r['roofline']['all'].items()}); print(r['roofline']['phase_us_per_iteration'], r['roofline']['refactor_ms_total'])
    print('e2e', r['e2e']['value'], 'cpu', r['cpu_baseline']['value'], [ (o['kind'], o['value']) for o in r['cpu_baseline'].get('others',[])])
    print('wall_to_optimal', r.get('wall_to_optimal_s'), r.get('optimal',{}).get('parity_ok'), r.get('objective_after_window'))
except Exception as e: print('bench parse fail', e)
PY
tail -2 gpurun_out/r2_bench_n1_final.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
