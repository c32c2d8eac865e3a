set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -2 gpurun_out/bench_full.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','roofline','gpu_launches')})
print({k:(v.get('iters_per_s') if isinstance(v,dict) else v) for k,v in d.get('secondary',{}).items() if 'vmc' in k})
PY
