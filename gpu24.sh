cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -m gpu -q --timeout 120 -x 2>&1 | tail -6 ) > gpurun_out/t24.log 2>&1
tail -5 gpurun_out/t24.log
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/bench24.json 2> gpurun_out/bench24.err; python - <<'PY'
import json
r=json.load(open('gpurun_out/bench24.json'))
print(r['value'], r['roofline']['frac'], r['roofline']['traffic'], r['roofline']['kernel'], r['e2e']['value'])
PY
tail -2 gpurun_out/bench24.err
