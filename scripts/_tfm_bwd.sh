set -x
timeout 300 python -m pytest tests/test_gpu_transformer.py -x -q 2>&1 | tail -5
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tfm.json 2> gpurun_out/bench_tfm.err; tail -2 gpurun_out/bench_tfm.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_tfm.json').read().strip().splitlines()[-1])
print({k:(v.get('iters_per_s') if isinstance(v,dict) else v) for k,v in d.get('secondary',{}).items() if 'vmc' in k})
PY
