set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python scripts/tfm_bwd_prof.py 10000 3 2>&1 | tail -3
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -2 gpurun_out/bench_full.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')})
s=d.get('secondary',{})
print({k:(v.get('iters_per_s') if isinstance(v,dict) else v) for k,v in s.items() if 'vmc' in k})
for k,v in s.items():
    if 'made' in k or 'amplitude' in k or 'backward' in k: print(k, v)
PY
