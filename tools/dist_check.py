"""Multi-GPU parity: run under torchrun; every rank must return the same signal / populations as a single GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import warnings; warnings.simplefilter('ignore')
os.environ['QSX_SHARD_MIN'] = '1'          # small grids: shard them anyway, the sharded path is what is checked
import numpy as np, torch, torch.distributed as dist
from qsextra_b200 import ChromophoreSystem
from qsextra_b200.engine import dist as qdist
from qsextra_b200.spectroscopy import FeynmanDiagram, qspectroscopy
from qsextra_b200.qcomo import qevolve_batch
from qsextra_b200.engine.program import SystemBatch

local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
rank, size = qdist.world()
DT = 0.1 / 0.658212
J = np.zeros((4, 4))
for i in range(3):
    J[i, i + 1] = J[i + 1, i] = -0.01
s = ChromophoreSystem(energies=[1.55, 1.50, 1.46, 1.52], couplings=J, dipole_moments=[1., 0.9, 1.1, 1.],
                      frequencies_pseudomode=0.02, levels_pseudomode=2, couplings_ep=0.054)
kw = dict(dt=DT, coll_rates=[0.2])
delays = [(15 * DT * np.arange(5)).tolist(), (100 * DT * np.arange(3)).tolist(), (15 * DT * np.arange(7)).tolist()]
worst = 0.0
for lab in ('gsb', 'se', 'esa'):
    sharded = qspectroscopy(s, FeynmanDiagram(lab, delays), shots=0, verbose=False, **kw)
    # single-rank reference: temporarily pretend there is no process group
    w = qdist.world
    qdist.world = lambda: (0, 1)
    alone = qspectroscopy(s, FeynmanDiagram(lab, delays), shots=0, verbose=False, **kw)
    qdist.world = w
    worst = max(worst, float(np.abs(sharded - alone).max()))
rng = np.random.default_rng(0)
B = 37
batch = SystemBatch(rng.uniform(1.4, 1.6, (B, 2)), np.stack([[[0, j], [j, 0]] for j in rng.uniform(-.05, .05, B)]), np.ones((B, 2)),
                    rng.uniform(.01, .2, (B, 1)), rng.uniform(.02, .08, (B, 1)), (2,))
rates = rng.uniform(0.1, 0.4, (B, 1))
t = np.arange(0, 51) * DT
sh = qevolve_batch(batch, t, dt=DT, coll_rates=rates)
al = qevolve_batch(batch, t, dt=DT, coll_rates=rates, shard=False)
worst = max(worst, float(np.abs(sh - al).max()))
# fewer parameter sets than ranks: some ranks hold an empty shard
one = SystemBatch(batch.energies[:1], batch.couplings[:1], batch.dipole_moments[:1], batch.omega[:1], batch.coupl_ep[:1], (2,))
sh1 = qevolve_batch(one, t, dt=DT, coll_rates=rates[:1])
worst = max(worst, float(np.abs(sh1 - al[:1]).max()))
# an ensemble (parameter sets of one program): members over the ranks, one gather; 3 members also leave ranks empty at N = 4, 8
from qsextra_b200.spectroscopy import qspectroscopy_ensemble
members = [ChromophoreSystem(energies=(np.array([1.55, 1.50, 1.46, 1.52]) + 0.01 * k).tolist(), couplings=J,
                             dipole_moments=[1., 0.9, 1.1, 1.], frequencies_pseudomode=0.02, levels_pseudomode=2,
                             couplings_ep=0.054) for k in range(3)]
specs = [FeynmanDiagram(lab, delays) for lab in ('gsb', 'esa')]
lo3, hi3 = qdist.shard_bounds(3, rank, size)
mine = [torch.zeros((hi3 - lo3, 5, 3, 7), dtype=torch.complex128).pin_memory() for _ in specs]
together = qspectroscopy_ensemble(members, specs, verbose=False, host_out=mine, **kw)
for d in range(len(specs)):                       # every rank's own members were copied to the host on the way
    worst = max(worst, float(np.abs(mine[d].numpy() - together[d][lo3:hi3]).max()) if hi3 > lo3 else 0.0)
w = qdist.world
qdist.world = lambda: (0, 1)
for k, member in enumerate(members):
    for d, spec in enumerate(specs):
        alone = qspectroscopy(member, spec, shots=0, verbose=False, **kw)
        worst = max(worst, float(np.abs(together[d][k] - alone).max()))
qdist.world = w
flag = torch.tensor([worst], device='cuda')
dist.all_reduce(flag, op=dist.ReduceOp.MAX)
if rank == 0:
    print('world', size, 'max |sharded - single| over ranks =', float(flag.item()))
assert float(flag.item()) < 1e-12
dist.destroy_process_group()
