"""Time qsx_readout_gemm at the C4 shapes (CUDA events, L2 flushed between launches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qsextra_b200.engine.runtime import get_engine
eng = get_engine()
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
peak = eng.fp64_peak()
for m, n, k in [(2048, 128, 6144), (2048, 128, 1024), (256, 128, 6144)]:
    a = torch.randn(m, k, dtype=torch.complex128, device='cuda')
    b = torch.randn(n, k, dtype=torch.complex128, device='cuda')
    want = a @ b.T
    got = eng.readout_gemm(a, b)
    err = float((got - want).abs().max() / want.abs().max())
    times = []
    for _ in range(10):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.readout_gemm(a, b); e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = sorted(times)[len(times) // 2]
    tf = 8.0 * m * n * k / (ms * 1e-3) / 1e12
    print('%d x %d x %d: %.3f ms, %.1f TFLOP/s = %.0f %% of the measured DFMA peak (%.1f), rel err %.1e' % (m, n, k, ms, tf, 100 * tf / peak, peak, err))
