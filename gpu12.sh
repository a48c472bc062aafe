cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x --durations=8 2>&1 | tail -20 ) > gpurun_out/t13.log 2>&1
tail -25 gpurun_out/t13.log
timeout 600 python bench.py --steps 4 --warmup 3 > gpurun_out/bench13.json 2> gpurun_out/bench13.err; python - <<'PY'
import json
try:
    r=json.load(open('gpurun_out/bench13.json'))
    print({k:r[k] for k in ('value','ms_per_step')}, r['config']['timed_iterations'], r['config']['nucleus_size'])
    print(r['roofline']['all']); print(r['roofline']['phase_us_per_iteration'], r['roofline']['refactor_ms_total'])
    print(r['e2e'], r['cpu_baseline'])
except Exception as e: print('bench parse fail', e)
PY
tail -3 gpurun_out/bench13.err
