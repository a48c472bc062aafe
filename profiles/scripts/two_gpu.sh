cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi -L
( time timeout 500 python -m pytest tests -m gpu -q --timeout 200 -x 2>&1 | tail -8 ) > gpurun_out/t22.log 2>&1
tail -6 gpurun_out/t22.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench22_n2.json 2> gpurun_out/bench22_n2.err; grep -o '"value": [0-9.]*' gpurun_out/bench22_n2.json | head -1; head -c 100 gpurun_out/bench22_n2.json; echo; tail -2 gpurun_out/bench22_n2.err
timeout 300 python tests/ab_probe.py pfiApplyVariant 0 1 2>&1 | tail -6
