"""A/B of the three forms of qsx_readout_gemm (QSX_GEMM=old|simt|dmma) at the shapes of the C4 response tree
(CUDA events, L2 flushed between launches; results checked against torch's ZGEMM)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qsextra_b200.engine.runtime import get_engine
eng = get_engine()
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
peak = eng.fp64_peak()
only = os.environ.get('QSX_GEMM_ONLY')
shapes = [(8192, 128, 6144), (8192, 128, 1024), (2048, 128, 6144), (2048, 128, 1024), (256, 128, 6144), (1024, 2048, 64),
          (1000, 77, 333)]
print('measured DFMA peak %.1f TFLOP/s' % peak)
for m, n, k in shapes:
    a = torch.randn(m, k, dtype=torch.complex128, device='cuda')
    b = torch.randn(n, k, dtype=torch.complex128, device='cuda')
    want = a @ b.T
    for form in ([only] if only else ['old', 'simt', 'dmma']):
        os.environ['QSX_GEMM'] = form
        got = eng.readout_gemm(a, b)
        err = float((got - want).abs().max() / want.abs().max())
        times = []
        for _ in range(int(os.environ.get('QSX_REPS', '10'))):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); eng.readout_gemm(a, b); e1.record(); torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = sorted(times)[len(times) // 2]
        tf = 8.0 * m * n * k / (ms * 1e-3) / 1e12
        print('%5d x %4d x %4d  %-4s: %.3f ms, %5.1f TFLOP/s = %3.0f %% of the DFMA peak, rel err %.1e' % (m, n, k, form, ms, tf, 100 * tf / peak, err))
    # batched (ensemble) form of the first shape class
a = torch.randn(8, 1024, 6144, dtype=torch.complex128, device='cuda')
b = torch.randn(8, 128, 6144, dtype=torch.complex128, device='cuda')
want = torch.matmul(a, b.transpose(1, 2))
for form in ([only] if only else ['old', 'simt', 'dmma']):
    os.environ['QSX_GEMM'] = form
    got = eng.readout_gemm(a, b)
    print('batched 8 x (1024 x 128 x 6144) %-4s rel err %.1e' % (form, float((got - want).abs().max() / want.abs().max())))
