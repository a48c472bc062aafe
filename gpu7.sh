cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 120 ./tests/microbench/sv > gpurun_out/microbench.log 2>&1; cat gpurun_out/microbench.log
