cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv,noheader
nproc
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -70 > gpurun_out/pytest_gpu.log
tail -45 gpurun_out/pytest_gpu.log
timeout 300 python tests/perf_probe.py 1000 10000 100000 1 16 0 > gpurun_out/probe_1k.log 2>&1; tail -5 gpurun_out/probe_1k.log
