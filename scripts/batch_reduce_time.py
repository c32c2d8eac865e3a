"""Time of the batch reductions of the MADE backward at the C5 shape against torch (cuBLAS) on the same operands."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from anqs_quantum_chemistry_b200 import _lib
dev = torch.device('cuda:0')
m = int(sys.argv[1]) if len(sys.argv) > 1 else 87000
QD, width, n, depth = 640, 64, 56, 2
dY = torch.randn(2, m, QD, dtype=torch.float64, device=dev); da = torch.randn(2, depth, m, width, dtype=torch.float64, device=dev)
h = torch.randn(2, depth, m, width, dtype=torch.float64, device=dev); x = torch.randn(m, n, dtype=torch.float64, device=dev)
f64 = dict(dtype=torch.float64, device=dev)
gW_out, gb_out, gW0, gb_h, gWm = torch.empty(2, QD, width, **f64), torch.empty(2, QD, **f64), torch.empty(2, width, n, **f64), torch.empty(2, depth, width, **f64), torch.empty(2, depth - 1, width, width, **f64)
problems = []
for net in range(2):
    problems.append((dY[net].data_ptr(), QD, QD, h[net, depth - 1].data_ptr(), width, width, gW_out[net].data_ptr(), width, gb_out[net].data_ptr()))
    problems.append((da[net, 0].data_ptr(), width, width, x.data_ptr(), n, n, gW0[net].data_ptr(), n, gb_h[net, 0].data_ptr()))
    for l in range(1, depth):
        problems.append((da[net, l].data_ptr(), width, width, h[net, l - 1].data_ptr(), width, width, gWm[net, l - 1].data_ptr(), width, gb_h[net, l].data_ptr()))
def ours(): _lib.batch_reduce(problems, m, False, dev)
def lib():
    return (torch.bmm(dY.transpose(1, 2), h[:, depth - 1]), dY.sum(1), torch.matmul(da[:, 0].transpose(1, 2), x),
            torch.matmul(da[:, 1:].transpose(2, 3), h[:, :depth - 1]), da.sum(2))
def t(fn):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10
a, b = t(ours), t(lib)
flop = 2.0 * m * 2 * (QD * width + width * n + (depth - 1) * width * width)
r = lib(); ours()
print(f'rows {m}: batch_reduce_gemm {a:.3f} ms ({flop / a / 1e9:.1f} TFLOP/s fp64), torch/cuBLAS {b:.3f} ms ({flop / b / 1e9:.1f}); '
      f'max diff W_out {float((gW_out - r[0]).abs().max()):.2e} b_out {float((gb_out - r[1]).abs().max()):.2e} W0 {float((gW0 - r[2]).abs().max()):.2e}')
