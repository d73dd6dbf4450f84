"""Latency regime of the CTA-wide kernels: a few instances of the (1,1) block of the C4 aggregate, each alone on an SM,
walked through 1500 steps by the first, second and third form (QSX_RRC_V2 / QSX_RRC_V3)."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter('ignore')
import numpy as np, torch
import bench
from qsextra_b200.engine.runtime import get_engine
from qsextra_b200.spectroscopy.simulation.qspectroscopy import ensemble_program
eng = get_engine()
program, mu, sequences, counts = ensemble_program(bench.c4_members(1), bench.c4_specs(), dt=bench.DT, coll_rates=[0.2])
for ka, kb in ((1, 1), (0, 1), (1, 0)):
    block = program.block(ka, kb)
    dev = eng.upload_program(program.bind(block, fused=False), program)
    rng = np.random.default_rng(0)
    for n_inst in (8, 64):
        sigma = eng.to_device(rng.normal(size=(n_inst, block.size)) + 1j * rng.normal(size=(n_inst, block.size)))
        emit = eng.to_device(np.arange(100, 1501, 100), torch.int32)
        ref = None
        for label, v2, v3 in (('first form', '0', '0'), ('second form', '1', '0'), ('third form', '1', '1')):
            os.environ['QSX_RRC_V2'], os.environ['QSX_RRC_V3'] = v2, v3
            run = lambda: eng.evolve(dev, sigma, n_inst, 1500, emit, snapshots=True)[1]
            out = run(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                out = run()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            if ref is None:
                ref = out
            print('(%d,%d) %3d instances  %-12s %.3f ms = %.0f SM-cycles per step   max |diff to first form| %.2e' % (
                ka, kb, n_inst, label, ms, ms * 1e-3 * 1.965e9 / 1500, float((out - ref).abs().max())))

# transposed steps (the read-out propagations of the Heisenberg picture): one instance per member, an emission every 15
# steps; first form against the damping-diagonal form for transposed programs (qsx_rrc3t.cuh)
for ka, kb in ((1, 0), (0, 1), (2, 1)):
    block = program.block(ka, kb)
    dev = eng.upload_program(program.bind(block, fused=False).transposed(), program)
    rng = np.random.default_rng(1)
    for n_inst in (1, 8):
        sigma = eng.to_device(rng.normal(size=(n_inst, block.size)) + 1j * rng.normal(size=(n_inst, block.size)))
        emit = eng.to_device(np.arange(0, 1906, 15), torch.int32)
        ref = None
        for label, v3 in (('first form', '0'), ('damping-diagonal', '1')):
            os.environ['QSX_RRC_V2'], os.environ['QSX_RRC_V3'] = '0', v3
            run = lambda: eng.evolve(dev, sigma, n_inst, 1905, emit, snapshots=True)[1]
            out = run(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                out = run()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            if ref is None:
                ref = out
            print('(%d,%d) transposed %d instance(s)  %-17s %.3f ms = %.0f SM-cycles per step   max |diff| %.2e of %.2e' % (
                ka, kb, n_inst, label, ms, ms * 1e-3 * 1.965e9 / 1905, float((out - ref).abs().max()), float(ref.abs().max())))
