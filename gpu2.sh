cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python tests/perf_probe.py 10000 100000 3000 1 16 0 > gpurun_out/probe_c2_t.log 2>&1; tail -3 gpurun_out/probe_c2_t.log
timeout 1200 python tests/perf_probe.py 10000 100000 1000000 0 32 0 1 > gpurun_out/probe_c2_full.log 2>&1; tail -4 gpurun_out/probe_c2_full.log
