cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export PATH=$PATH:/usr/local/cuda/bin
( time timeout 400 python -m pytest tests -m gpu -q --timeout 120 -x 2>&1 | tail -8 ) > gpurun_out/t23.log 2>&1
tail -10 gpurun_out/t23.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(?!lu_|trsm|gemm_sub|set_perm|transpose|zero_pad|gather_nucleus_matrix).*' -s 10300 -c 50 --csv --log-file gpurun_out/launches23.csv python tests/ncu_target.py c2 1012 > gpurun_out/ncu23.log 2>&1; tail -2 gpurun_out/ncu23.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(lu_|trsm|gemm_sub|set_perm|transpose|zero_pad|gather_nucleus_matrix).*' -c 3000 --csv --log-file gpurun_out/launches23_refactor.csv python tests/ncu_target.py c2 2 > gpurun_out/ncu23b.log 2>&1; tail -1 gpurun_out/ncu23b.log
timeout 600 python bench.py --steps 4 --warmup 3 > gpurun_out/bench23.json 2> gpurun_out/bench23.err; python - <<'PY'
import json
try:
    r=json.load(open('gpurun_out/bench23.json'))
    print({k:r[k] for k in ('value','ms_per_step')}, r['config']['timed_iterations'], r['config']['nucleus_size'], 'launches/iter', r['gpu_launches']/max(1,r['config']['timed_iterations']+3*2000))
    print(r['roofline']['all']); print(r['roofline']['phase_us_per_iteration'], r['roofline']['refactor_ms_total'])
    print(r['e2e']['value'], r['cpu_baseline']['value'])
except Exception as e: print('bench parse fail', e)
PY
tail -3 gpurun_out/bench23.err
