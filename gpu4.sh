cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python tests/perf_probe.py 1000 10000 100000 0 32 0 > gpurun_out/probe_1k.log 2>&1; tail -2 gpurun_out/probe_1k.log
timeout 600 python tests/perf_probe.py 10000 100000 1100 1 16 0 > gpurun_out/probe_c2_t.log 2>&1; tail -3 gpurun_out/probe_c2_t.log
timeout 600 python tests/perf_probe.py 10000 100000 8000 0 32 0 > gpurun_out/probe_c2_g.log 2>&1; tail -2 gpurun_out/probe_c2_g.log
