cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export PATH=$PATH:/usr/local/cuda/bin
nvidia-smi -L | head -3
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 4 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; python - <<'PY'
import json
try:
    r=json.load(open('gpurun_out/bench_n1.json'))
    print({k:r[k] for k in ('value','ms_per_step')}, r['config']['timed_iterations'], r['config']['nucleus_size'])
    print(r['roofline']['all']); print(r['roofline']['phase_us_per_iteration'], r['roofline']['refactor_ms_total'])
except Exception as e: print('bench parse fail', e)
PY
tail -3 gpurun_out/bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 600 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
