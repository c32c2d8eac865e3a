"""Times SampleTable builds (k2_hash.cu) at the key counts of the 1..8-GPU bench."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from anqs_quantum_chemistry_b200 import SampleTable, synthetic, dist as adist
dev = torch.device('cuda:0')
def tm(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
SIZES = [int(v) for v in os.environ.get("SIZES", "%d,%d,%d,%d" % (1 << 20, 1 << 21, 1 << 22, 1 << 23)).split(",")]
for n in SIZES:
    s = synthetic.random_physical_samples(56, 7, 7, n, seed=1)
    a = synthetic.random_amplitudes(s.shape[0], seed=2)
    ds, da = torch.from_numpy(s.view(np.int64)).to(dev), torch.from_numpy(a).to(dev)
    t = SampleTable(ds, da)
    reps = int(os.environ.get('REPS', 5))
    print(n, 'keys: rebuild in place', round(tm(lambda: t.rebuild(ds, da), reps=reps, warm=1), 3), 'ms')
    del t
    e = torch.randn(n // 8 if n > (1 << 20) else n, dtype=torch.complex128, device=dev)
    print('   stats', round(tm(lambda: adist.local_energy_stats(e, da[:e.shape[0]])), 3), 'ms')
