"""One C4 system through qspectroscopy_set under torchrun (every rank evaluates it whole): timing and traceback probe."""
import os, sys, time, traceback, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter('ignore')
import numpy as np, torch, torch.distributed as dist
import bench
from qsextra_b200.spectroscopy import qspectroscopy_set
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
if int(os.environ.get('WORLD_SIZE', '1')) > 1:
    import datetime
    dist.init_process_group('nccl', device_id=torch.device('cuda', local), timeout=datetime.timedelta(seconds=120))
chain, specs = bench.c4_member(), bench.c4_specs()
try:
    for rep in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sigs = qspectroscopy_set(chain, specs, verbose=False, dt=bench.DT, coll_rates=[0.2])
        torch.cuda.synchronize()
        if local == 0:
            print('run %d: %.2f ms checksum %.9f' % (rep, (time.perf_counter() - t0) * 1e3, sum(float(np.abs(s).sum()) for s in sigs)), flush=True)
except Exception:
    traceback.print_exc()
if dist.is_initialized():
    dist.destroy_process_group()
