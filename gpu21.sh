cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export PATH=$PATH:/usr/local/cuda/bin
nvidia-smi -L
( time timeout 500 python -m pytest tests -m gpu -q --timeout 200 -x 2>&1 | tail -8 ) > gpurun_out/t21.log 2>&1
tail -6 gpurun_out/t21.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench21_n2.json 2> gpurun_out/bench21_n2.err; tail -c 600 gpurun_out/bench21_n2.json; tail -3 gpurun_out/bench21_n2.err
timeout 400 python bench.py --steps 4 --warmup 3 > gpurun_out/bench21.json 2> gpurun_out/bench21.err; python - <<'PY'
import json
try:
    r=json.load(open('gpurun_out/bench21.json'))
    print({k:r[k] for k in ('value','ms_per_step')}, r['config']['timed_iterations'], r['config']['nucleus_size'])
    print(r['roofline']['all']); print(r['roofline']['phase_us_per_iteration'], r['roofline']['refactor_ms_total'])
    print(r['e2e']['value'], r['cpu_baseline']['value'])
except Exception as e: print('bench parse fail', e)
PY
tail -3 gpurun_out/bench21.err
