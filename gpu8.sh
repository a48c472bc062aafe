cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export PATH=$PATH:/usr/local/cuda/bin
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; python - <<'PY'
import json
try:
    r=json.load(open('gpurun_out/bench_n1.json'))
    print({k:r[k] for k in ('value','ms_per_step','gpu_launches')}, r['config']['timed_iterations'], r['config']['nucleus_size'])
    print(r['roofline']['all']); print(r['roofline']['phase_us_per_iteration'], r['roofline']['refactor_ms_total'])
    print(r['e2e']['value'], r['cpu_baseline']['value'])
except Exception as e: print('bench parse fail', e)
PY
tail -3 gpurun_out/bench_n1.err
