cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 -x -k "dense_invert" 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -5
timeout 600 python bench.py --steps 4 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; python - <<'PY'
import json
try:
    r=json.load(open('gpurun_out/bench_n1.json'))
    print({k:r[k] for k in ('value','ms_per_step')}, r['config']['timed_iterations'], r['config']['nucleus_size'])
    print(r['roofline']['all']); print(r['roofline']['phase_us_per_iteration'], r['roofline']['refactor_ms_total'])
except Exception as e: print('bench parse fail', e)
PY
tail -3 gpurun_out/bench_n1.err
