"""One-line summary of a bench.py log (value, e2e, ms/step, kernel share, roofline fraction)."""
import json, sys
for line in open(sys.argv[1]):
    if line.startswith('{'):
        d = json.loads(line)
        k = d.get('kernel', {})
        print(f"n={d['config'].get('n_unq_per_gpu')} value={d['value']:.4g} e2e={d['e2e']['value']:.4g} ms/step={d['ms_per_step']:.3f} "
              f"kernel_ms={k.get('ms', 0):.3f} share={k.get('share_of_step', 0):.2f} frac={d.get('roofline', {}).get('frac', 0):.3f}")
    elif 'Error' in line or 'error' in line:
        print(line.rstrip()[:300])
