cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python tests/trace_compare.py 1000 10000 0.01 400 > gpurun_out/trace_1k.log 2>&1; tail -12 gpurun_out/trace_1k.log
timeout 900 python tests/trace_compare.py 4000 40000 0.01 600 > gpurun_out/trace_4k.log 2>&1; tail -14 gpurun_out/trace_4k.log
