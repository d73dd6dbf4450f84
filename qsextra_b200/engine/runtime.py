"""Device-side runtime: PyTorch owns memory and streams, ``libqsx.so`` does the arithmetic.

``Engine`` keeps the small per-block tables on the device and turns host-side programs
(``program.BoundProgram``) into ``qsx_evolve`` / ``qsx_apply_dipole`` calls on the current CUDA stream.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os

import numpy as np
import torch

from . import cabi
from .program import BoundProgram
from .sectors import Block, sector_rank


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class DeviceBlock:
    """Sector tables of one block resident on the device."""

    def __init__(self, block: Block, device):
        self.block = block
        self.cfg_a = torch.tensor(block.cfg_a, dtype=torch.int32, device=device)
        self.cfg_b = torch.tensor(block.cfg_b, dtype=torch.int32, device=device)
        self.rank_a = torch.from_numpy(sector_rank(block.n_sites, block.ka).copy()).to(device)
        self.rank_b = torch.from_numpy(sector_rank(block.n_sites, block.kb).copy()).to(device)
        self.c = cabi.QsxBlock(block.n_sites, block.n_bits, block.n_a, block.n_b,
                               self.cfg_a.data_ptr(), self.cfg_b.data_ptr(),
                               self.rank_a.data_ptr(), self.rank_b.data_ptr())


class Engine:
    """One engine per process/GPU.  Raises when CUDA or the extension is unavailable (no CPU path)."""

    def __init__(self, device=None):
        self.lib = cabi.load()
        if not torch.cuda.is_available():
            raise cabi.QsxError('no CUDA device: the qsextra_b200 engine has no CPU fallback')
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self._blocks: dict[Block, DeviceBlock] = {}
        self._recipes: dict = {}              # device copies of parameter recipes, by content
        self._side_stream = None
        self._staging = None                  # pinned read-back buffer (to_host)
        self.trace = None                     # list of (name, meta, start, stop) while bench.py profiles launches
        self._memo = None                     # while a response plan is built: uploads by content (see to_device)
        self._capturing = False
        self.plans = {}                       # response plans (CUDA graphs) by content key (engine/response.py)
        self.plan_seen = {}
        self.aux_launches = 0                 # table-assembly kernels (qsx_bind_params, qsx_rr_tables)
        self.launches = 0                     # kernels enqueued through this engine (bench bookkeeping)
        self.h2d_bytes = 0                    # bytes copied host -> device through to_device()
        with torch.cuda.device(self.device):
            sm, smem, cc = C.c_int32(), C.c_int32(), C.c_int32()
            cabi.check(self.lib.qsx_device_info(C.byref(sm), C.byref(smem), C.byref(cc)))
        self.sm_count, self.smem_optin, self.cc = sm.value, smem.value, cc.value
        self.n_shipped_rrc = 20                 # programs of csrc/qsx_rrc_programs.inc (indices below it: shipped menu)

    # ---- helpers --------------------------------------------------------------------------
    @contextlib.contextmanager
    def span(self, name: str, **meta):
        """When ``self.trace`` is a list: CUDA events around one launch on the stream it is enqueued on (bench.py's
        per-kernel device times; never active inside a captured graph)."""
        if self.trace is None:
            yield
            return
        stream = torch.cuda.current_stream(self.device)
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record(stream)
        yield
        stop.record(stream)
        self.trace.append((name, meta, start, stop))

    @property
    def side_stream(self):
        """One extra stream per engine for work that is independent of the main sequence (kept, so that the caching
        allocator reuses its blocks from call to call)."""
        return self.side_stream_of(0)

    def side_stream_of(self, index: int):
        """Side stream number ``index`` (a small fixed set, created once)."""
        if self._side_stream is None:
            self._side_stream = []
        index %= 12
        while len(self._side_stream) <= index:
            # streams 0-3 carry the read-out propagations: a few CTAs that walk every step alone and are needed at the
            # very end -- they must not queue behind a level that fills every SM, hence the higher priority (inside a
            # captured response plan the kernel nodes inherit it)
            priority = -1 if len(self._side_stream) < 4 and not os.environ.get('QSX_FLAT_PRIORITIES') else 0
            self._side_stream.append(torch.cuda.Stream(device=self.device, priority=priority))
        return self._side_stream[index]

    def device_block(self, block: Block) -> DeviceBlock:
        db = self._blocks.get(block)
        if db is None:
            db = self._blocks[block] = DeviceBlock(block, self.device)
        return db

    def to_host(self, tensor: torch.Tensor) -> np.ndarray:
        """Device tensor -> NumPy array through a pinned staging buffer kept by the engine (a pageable read-back of a
        few MB costs milliseconds; the staged copy runs at PCIe speed and the final host copy at memory speed)."""
        flat = tensor.contiguous().view(-1)
        n_bytes = flat.numel() * flat.element_size()
        if self._staging is None or self._staging.numel() < n_bytes:
            self._staging = torch.empty(max(n_bytes, 1 << 20), dtype=torch.uint8).pin_memory()
        stage = self._staging[:n_bytes].view(flat.dtype)
        stage.copy_(flat, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return stage.numpy().reshape(tuple(tensor.shape)).copy()

    def to_device(self, array, dtype=None):
        """Host array -> device tensor.  While a response plan is being built (``_memo`` is a dict) uploads are kept by
        content: the capture pass of the CUDA graph then finds every table of the warm pass already resident -- a
        host-to-device copy cannot be captured -- and the plan keeps them alive."""
        host = np.ascontiguousarray(array)
        key = None
        if self._memo is not None:
            import hashlib
            key = (host.dtype.str, host.shape, str(dtype), hashlib.blake2b(host.tobytes(), digest_size=16).digest())
            hit = self._memo.get(key)
            if hit is not None:
                return hit
            if self._capturing:
                raise cabi.QsxError('an upload the warm pass did not make was requested during graph capture')
        t = torch.as_tensor(host)
        if dtype is not None:
            t = t.to(dtype)
        self.h2d_bytes += t.numel() * t.element_size()
        out = t.to(self.device, non_blocking=True)
        if key is not None:
            self._memo[key] = out
        return out

    def upload_program(self, bound: BoundProgram, program=None, use_rr: bool = True, reg_bits=None):
        """-> dict of device tensors for a bound program (ops + parameter sets).  With ``program`` (the StepProgram
        it was bound from) small power-of-two blocks are additionally mapped onto the register-resident kernel."""
        rr_entries = {}
        transposed = bool(getattr(bound, 'is_transposed', False))      # only the CTA-wide kernel and the op list run it
        if program is not None and use_rr and not transposed and not os.environ.get('QSX_DISABLE_RR'):
            from . import rr as rr_mod
            rrp = rr_mod.build(program, bound, reg_bits=reg_bits)
            if rrp is not None:
                shared = rrp._shared if rrp._shared is not None else {}
                lkey = ('dev_logical', self.device.index)
                if lkey not in shared:
                    shared[lkey] = self.to_device(rrp.logical_index, torch.int32)
                rr_entries = dict(rr=rrp, rr_params=rrp.device_params(self), logical=shared[lkey])
        dev = dict(block=self.device_block(bound.block),
                   ops=None if rr_entries else self.to_device(bound.ops, torch.int32),
                   params=None if rr_entries else self.to_device(bound.params, torch.float64),
                   bound=bound, n_ops=int(bound.ops.shape[0]), stride=int(bound.stride),
                   n_sets=int(bound.n_sets), n_passes=0, team=0)
        if bound.passes is not None and len(bound.passes):
            dev.update(n_passes=int(bound.passes.shape[0]), team=int(bound.team),
                       passes=self.to_device(bound.passes, torch.int32),
                       tiles=self.to_device(bound.tiles if len(bound.tiles) else np.zeros(2, np.int32), torch.int32),
                       hops=self.to_device(bound.hops if len(bound.hops) else np.zeros(4, np.int32), torch.int32),
                       n_tile_ints=int(len(bound.tiles)), n_hop_ints=int(len(bound.hops)))
        dev.update(rr_entries)
        # The CTA-wide kernel serves the blocks the register-resident small-block kernel cannot -- and the ground-state
        # block (no exciton on either side) next to it: its step is the diagonal multiply and the dampings only, which
        # the damping-diagonal form reduces to ONE complex multiply per element and step (see ``evolve``).
        no_exciton = bound.block.ka == 0 and bound.block.kb == 0
        if program is not None and use_rr and (not rr_entries or no_exciton) and not os.environ.get('QSX_DISABLE_RRC'):
            from . import rrc as rrc_mod
            big = rrc_mod.build(program, bound)
            if big is not None:
                dev.update(rrc=big, rrc_params=self.to_device(big.params, torch.float64),
                           rrc_signature=np.ascontiguousarray(big.signature()),
                           rrc_diagonal=all(op[0] in (rrc_mod.OP_DIAG, rrc_mod.OP_AD) for op in big.ops))
        return dev

    def upload_observables(self, ptr, idx, w, kind: str | None = None, mu=None):
        """CSR read-outs -> device tables.  ``kind`` ('populations' | 'emission' with ``mu``) tells the engine what the
        read-outs are, so that compile-time programs can use compile-time read-out masks."""
        return dict(n=len(ptr) - 1, ptr=self.to_device(ptr, torch.int32), idx=self.to_device(idx, torch.int32),
                    w=self.to_device(np.asarray(w, dtype=np.complex128)), host=(np.asarray(ptr), np.asarray(idx), np.asarray(w)),
                    kind={'populations': 1, 'emission': 2}.get(kind, 0), mu=None if mu is None else np.asarray(mu, dtype=float))

    def has_static_rrc(self, prog: dict) -> bool:
        """Can the CTA-wide register-resident kernel run this uploaded program?  Either a compile-time program of the
        shipped menu has its signature, or ``engine/jit.py`` specialises the kernel for it now (NVRTC; cached on disk)."""
        if prog.get('rrc') is None:
            return False
        sig = prog['rrc_signature']
        found = self.lib.qsx_rrc_find_program(C.c_void_p(sig.ctypes.data), int(sig.size))
        if found < 0:
            from . import jit
            if jit.ensure_registered(self.lib, prog['rrc'], sig):
                found = self.lib.qsx_rrc_find_program(C.c_void_p(sig.ctypes.data), int(sig.size))
        if found < 0:
            prog['rrc'] = None
        else:
            prog['rrc_family'] = 'shipped' if found < self.n_shipped_rrc else 'specialised at run time'
        return found >= 0

    def _evolve_rrc(self, prog, sigma_in, n_inst, n_steps, emit, obs, snapshots, inst_set, inst_src):
        """CTA-wide register-resident kernel (compile-time programs only).  Returns None when no compile-time program
        matches, so that the caller falls back to the shared-memory pass kernel."""
        big = prog['rrc']
        E = big.block.size
        n_emit = int(emit.numel())
        args = cabi.QsxRRCArgs()
        sig = prog['rrc_signature']
        args.n_signature, args.signature = int(sig.size), sig.ctypes.data
        args.param_stride, args.params = int(big.params.shape[1]), prog['rrc_params'].data_ptr()
        args.n_inst = n_inst
        args.inst_set = None if inst_set is None else inst_set.data_ptr()
        args.inst_src = None if inst_src is None else inst_src.data_ptr()
        args.sigma_in = sigma_in.data_ptr()
        args.n_steps, args.n_emit, args.emit_steps = int(n_steps), n_emit, emit.data_ptr()
        out_obs = out_snap = None
        if obs is not None and obs.get('kind') == 1:            # populations of a (KA, KA) block
            args.emit_trace = 2
            out_obs = torch.empty((n_inst, n_emit, big.block.n_sites), dtype=torch.float64, device=self.device)
            args.out_obs = out_obs.data_ptr()
        elif obs is not None:
            args.emit_trace = 1
            for i, m in enumerate(obs['mu'][:16]):
                args.mu[i] = float(m)
            out_obs = torch.empty((n_inst, n_emit, 1), dtype=torch.complex128, device=self.device)
            args.out_obs = out_obs.data_ptr()
        progress = None
        if snapshots:
            out_snap = torch.empty((n_inst, n_emit, E), dtype=torch.complex128, device=self.device)
            args.out_snap = out_snap.data_ptr()
            if obs is None and n_inst > 2 * self.sm_count and not os.environ.get('QSX_DISABLE_SEGMENTS'):
                # more instances than the GPU keeps resident: let the kernel schedule (instance, segment) pairs
                progress = torch.zeros(n_inst + 1, dtype=torch.int32, device=self.device)
                args.progress = progress.data_ptr()
        ex_flops, ex_instr = big.executed_flops_per_step()
        with torch.cuda.device(self.device), self.span(
                'evolve_rrc ({},{}) {}x{}{}'.format(big.block.ka, big.block.kb, big.block.dim_a, big.block.dim_b,
                                                   ' transposed' if any(op[0] == 6 for op in big.ops) else ''),
                instances=n_inst, steps=int(n_steps), elements=E, executed_flops_per_step=ex_flops,
                fp64_instructions_per_step=ex_instr, threads=big.threads):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            rc = self.lib.qsx_evolve_rrc(C.byref(args), C.c_void_p(stream))
        if rc == cabi.ERR_NO_PROGRAM:
            return None
        cabi.check(rc)
        if n_inst > 0 and n_emit > 0:
            self.launches += 1
        return out_obs, out_snap

    def _evolve_rr(self, prog, sigma_in, n_inst, n_steps, emit, obs, obs_real, snapshots, inst_set, inst_src):
        rrp = prog['rr']
        E = rrp.block.size
        n_emit = int(emit.numel())
        out_obs = out_snap = None
        args = cabi.QsxRRArgs()
        args.reg_bits, args.lane_bits, args.n_diag, args.n_ops = rrp.reg_bits, rrp.lane_bits, rrp.n_diag, len(rrp.ops)
        for i, row in enumerate(rrp.ops):
            o = args.ops[i]
            o.kind, o.reg_xor, o.lane_xor = int(row[0]), int(row[1]), int(row[2])
            o.reg_mask = int(row[3]) & 0xffffffff
            o.p, o.sign, o.aux0, o.aux1 = int(row[4]), int(row[5]), int(row[6]), int(row[7])
        args.logical_index = prog['logical'].data_ptr()
        args.param_stride, args.params = int(rrp.stride), prog['rr_params'].data_ptr()
        args.n_inst = n_inst
        args.inst_set = None if inst_set is None else inst_set.data_ptr()
        args.inst_src = None if inst_src is None else inst_src.data_ptr()
        args.sigma_in = sigma_in.data_ptr()
        args.n_steps, args.n_emit, args.emit_steps = int(n_steps), n_emit, emit.data_ptr()
        if obs is not None:
            key = ('rr', id(rrp))
            if key not in obs:
                dense, masks, entries = rrp.observables(*obs['host'])
                obs[key] = (self.to_device(dense), masks, entries)
            dense_dev, masks, entries = obs[key]
            args.n_obs, args.obs_weights = obs['n'], dense_dev.data_ptr()
            for i, m in enumerate(masks):
                args.obs_reg_mask[i] = int(m)
            args.readout_kind = obs.get('kind', 0)
            if args.readout_kind:
                pkey = ('rr_sites', id(rrp))
                if pkey not in obs:
                    obs[pkey] = rrp.site_pieces(args.readout_kind, rrp.block.n_sites)
                if obs[pkey] is None or any(len(p) > cabi.RR_MAX_ENTRIES for p in obs[pkey]):
                    args.readout_kind = 0
                else:
                    for s_, pieces in enumerate(obs[pkey]):
                        for e_, (rm, lm) in enumerate(pieces):
                            args.site_reg_mask[s_][e_], args.site_lane_mask[s_][e_] = rm, lm
            if obs.get('mu') is not None:
                for i, m in enumerate(obs['mu'][:16]):
                    args.mu[i] = float(m)
            if entries is not None:
                for o, ents in enumerate(entries):
                    args.n_entries[o] = len(ents)
                    for e, (rm, lm, wgt) in enumerate(ents):
                        args.entry_reg_mask[o][e], args.entry_lane_mask[o][e], args.entry_weight[o][e] = rm, lm, wgt
            out_obs = torch.empty((n_inst, n_emit, obs['n']), dtype=torch.float64 if obs_real else torch.complex128,
                                  device=self.device)
            args.out_obs = out_obs.data_ptr()
        args.obs_real = int(obs_real)
        if snapshots:
            out_snap = torch.empty((n_inst, n_emit, E), dtype=torch.complex128, device=self.device)
            args.out_snap = out_snap.data_ptr()
        ex_flops, ex_instr = rrp.executed_flops_per_step()
        with torch.cuda.device(self.device), self.span(
                'evolve_rr ({},{}) {}x{}'.format(rrp.block.ka, rrp.block.kb, rrp.block.dim_a, rrp.block.dim_b),
                instances=n_inst, steps=int(n_steps), elements=E, executed_flops_per_step=ex_flops,
                fp64_instructions_per_step=ex_instr, threads=rrp.team):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            cabi.check(self.lib.qsx_evolve_rr(C.byref(args), C.c_void_p(stream)))
        if n_inst > 0 and n_emit > 0:
            self.launches += 1
        return out_obs, out_snap

    # ---- kernels ------------------------------------------------------------------------------
    def evolve(self, prog: dict, sigma_in: torch.Tensor, n_inst: int, n_steps: int, emit_steps, obs: dict | None = None,
               obs_real: bool = False, snapshots: bool = False, inst_set=None, inst_src=None, threads: int = 0):
        """Run ``n_inst`` instances for ``n_steps`` steps.  Returns (out_obs | None, out_snap | None):
        out_obs [n_inst, n_emit, n_obs] (float64 if obs_real else complex128), out_snap [n_inst, n_emit, E] complex128."""
        blk: DeviceBlock = prog['block']
        E = blk.block.size
        if isinstance(emit_steps, torch.Tensor):         # (a device tensor is the caller's promise of the precondition)
            emit = emit_steps
        else:
            host_emit = np.asarray(emit_steps, dtype=np.int64).reshape(-1)
            if host_emit.size and (np.any(np.diff(host_emit) <= 0) or host_emit[0] < 0 or host_emit[-1] > n_steps):
                raise cabi.QsxError('emit_steps must be strictly increasing step counts in [0, n_steps] (include/qsx.h)')
            emit = self.to_device(host_emit, torch.int32)
        n_emit = int(emit.numel())
        assert sigma_in.dtype == torch.complex128 and sigma_in.is_contiguous() and sigma_in.numel() % E == 0
        if 'rr' in prog and (obs is None or obs['n'] <= cabi.RR_MAX_OBS):
            # a step of diagonal multiply + dampings (ground-state block), snapshots at a few steps only: the
            # damping-diagonal form of the CTA-wide kernel walks it with one complex multiply per element and step
            diagonal = (prog.get('rrc') is not None and prog.get('rrc_diagonal') and obs is None and snapshots
                        and n_emit * 8 <= int(n_steps) + 8 and not os.environ.get('QSX_DISABLE_DIAGONAL_RRC'))
            done = None
            if diagonal:
                if 'rrc_shipped' not in prog:           # (shipped programs only: that form is not specialised at run time)
                    sig = prog['rrc_signature']
                    found = self.lib.qsx_rrc_find_program(C.c_void_p(sig.ctypes.data), int(sig.size))
                    prog['rrc_shipped'] = 0 <= found < self.n_shipped_rrc
                if prog['rrc_shipped']:
                    done = self._evolve_rrc(prog, sigma_in, n_inst, n_steps, emit, obs, snapshots, inst_set, inst_src)
                    if done is None:
                        prog['rrc_shipped'] = False
            if done is not None:
                return done
            return self._evolve_rr(prog, sigma_in, n_inst, n_steps, emit, obs, obs_real, snapshots, inst_set, inst_src)
        blk_ = blk.block
        if prog.get('rrc') is not None and 'rrc_family' not in prog:
            self.has_static_rrc(prog)                   # (specialises the kernel for a structure outside the menu)
        rrc_serves = ((not obs_real and (obs is None or (obs.get('kind') == 2 and obs['n'] == 1))) or
                      (obs_real and obs is not None and obs.get('kind') == 1 and obs['n'] == blk_.n_sites and blk_.ka == blk_.kb))
        if prog.get('rrc') is not None and rrc_serves:
            done = self._evolve_rrc(prog, sigma_in, n_inst, n_steps, emit, obs, snapshots, inst_set, inst_src)
            if done is not None:
                return done
            prog['rrc'] = None                        # no compile-time program: remember and use the pass kernel
        out_obs = out_snap = None
        if obs is not None:
            out_obs = torch.empty((n_inst, n_emit, obs['n']), dtype=torch.float64 if obs_real else torch.complex128,
                                  device=self.device)
        if snapshots:
            out_snap = torch.empty((n_inst, n_emit, E), dtype=torch.complex128, device=self.device)
        if prog['params'] is None:                    # the register-resident path was planned but cannot serve this call
            prog['params'] = self.to_device(prog['bound'].params, torch.float64)
            prog['ops'] = self.to_device(prog['bound'].ops, torch.int32)
        args = cabi.QsxEvolveArgs()
        args.block = blk.c
        args.n_ops, args.ops = prog['n_ops'], prog['ops'].data_ptr()
        args.param_stride, args.params = prog['stride'], prog['params'].data_ptr()
        args.n_inst = n_inst
        args.inst_set = None if inst_set is None else inst_set.data_ptr()
        args.inst_src = None if inst_src is None else inst_src.data_ptr()
        args.sigma_in = sigma_in.data_ptr()
        args.n_steps, args.n_emit, args.emit_steps = int(n_steps), n_emit, emit.data_ptr()
        if obs is not None:
            args.n_obs, args.obs_ptr, args.obs_idx, args.obs_w = obs['n'], obs['ptr'].data_ptr(), obs['idx'].data_ptr(), obs['w'].data_ptr()
            args.out_obs = out_obs.data_ptr()
        args.obs_real = int(obs_real)
        args.out_snap = None if out_snap is None else out_snap.data_ptr()
        args.threads = threads or prog.get('team', 0)
        if prog.get('n_passes', 0) > 0:
            args.n_passes, args.passes = prog['n_passes'], prog['passes'].data_ptr()
            args.n_tile_ints, args.tiles = prog['n_tile_ints'], prog['tiles'].data_ptr()
            args.n_hop_ints, args.hops = prog['n_hop_ints'], prog['hops'].data_ptr()
        with torch.cuda.device(self.device):
            ws_bytes = self.lib.qsx_evolve_workspace_bytes(C.byref(args))
            cabi.check(ws_bytes)
            workspace = None
            if ws_bytes > 0:
                workspace = torch.empty(ws_bytes // 8, dtype=torch.float64, device=self.device)
                args.workspace = workspace.data_ptr()
            stream = torch.cuda.current_stream(self.device).cuda_stream
            with self.span('evolve ({},{}) {}x{}'.format(blk_.ka, blk_.kb, blk_.dim_a, blk_.dim_b), instances=n_inst,
                           steps=int(n_steps), elements=E):
                cabi.check(self.lib.qsx_evolve(C.byref(args), C.c_void_p(stream)))
        if n_inst > 0 and n_emit > 0:
            self.launches += 1
        return out_obs, out_snap

    def apply_dipole(self, in_block: Block, out_block: Block, side: int, mu: torch.Tensor, sigma_in: torch.Tensor):
        """Collective dipole operator on a batch of blocks: [n, E_in] -> [n, E_out]."""
        n_inst = sigma_in.numel() // in_block.size
        out = torch.empty((n_inst, out_block.size), dtype=torch.complex128, device=self.device)
        bi, bo = self.device_block(in_block), self.device_block(out_block)
        with torch.cuda.device(self.device), self.span('apply_dipole', instances=n_inst, elements=out_block.size):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            cabi.check(self.lib.qsx_apply_dipole(C.byref(bi.c), C.byref(bo.c), side, C.c_void_p(mu.data_ptr()), n_inst,
                                                 C.c_void_p(sigma_in.data_ptr()), C.c_void_p(out.data_ptr()),
                                                 C.c_void_p(stream)))
        if n_inst > 0:
            self.launches += 1
        return out

    def bind_params(self, bound):
        """Parameter rows [B, stride] of a bound program on the device: evaluated there from the raw angles of the
        step (``qsx_bind_params``) when the program carries its recipe, else uploaded."""
        if bound.recipe is None or bound._params is not None or not bound.recipe.width:
            return self.to_device(bound.params, torch.float64)
        recipe = bound.recipe
        kind, ptr, idx, coef = recipe.tables()
        key = (kind.tobytes(), ptr.tobytes(), idx.tobytes(), coef.tobytes())
        tabs = self._recipes.get(key)
        if tabs is None:
            if len(self._recipes) > 256:
                self._recipes.clear()
            tabs = self._recipes[key] = (self.to_device(kind, torch.int32), self.to_device(ptr, torch.int32),
                                         self.to_device(idx if idx.size else np.zeros(1, np.int32), torch.int32),
                                         self.to_device(coef if coef.size else np.zeros(1)))
        angles = self.to_device(recipe.angles())
        out = torch.empty((recipe.n_sets, recipe.width), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            cabi.check(self.lib.qsx_bind_params(recipe.n_sets, recipe.width, int(angles.shape[1]),
                                                C.c_void_p(tabs[0].data_ptr()), C.c_void_p(tabs[1].data_ptr()),
                                                C.c_void_p(tabs[2].data_ptr()), C.c_void_p(tabs[3].data_ptr()),
                                                C.c_void_p(angles.data_ptr()), C.c_void_p(out.data_ptr()), C.c_void_p(stream)))
        self.aux_launches += 1
        return out

    def rr_tables(self, rrp, op_params: torch.Tensor, a_of: torch.Tensor, b_of: torch.Tensor, counts):
        """Parameter sets [B, stride] of the register-resident kernel from the op parameters (``qsx_rr_tables``)."""
        blk = rrp.block
        B, S = op_params.shape
        args = cabi.QsxRRTablesArgs()
        args.n_sets, args.n_elem, args.n_bits, args.n_cfg_b = int(B), blk.size, blk.n_bits, blk.n_b
        args.n_diag = len(rrp.diag_specs)
        for d, (sa, sb, st) in enumerate(rrp.diag_specs):
            args.slot_ga[d], args.slot_gb[d], args.slot_t[d] = int(sa), int(sb), int(st)
        args.n_classes = len(rrp.fold_classes)
        for c, (slot, _) in enumerate(rrp.fold_classes):
            args.class_slot[c] = int(slot)
        args.class_count = None if counts is None else counts.data_ptr()
        args.n_tangent = len(rrp.folded_slots)
        for k, slot in enumerate(rrp.folded_slots):
            args.tangent_slot[k] = int(slot)
        args.a_of, args.b_of = a_of.data_ptr(), b_of.data_ptr()
        args.op_stride, args.op_params = int(S), op_params.data_ptr()
        out = torch.empty((B, rrp.stride), dtype=torch.float64, device=self.device)
        args.out = out.data_ptr()
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            cabi.check(self.lib.qsx_rr_tables(C.byref(args), C.c_void_p(stream)))
        self.aux_launches += 1
        return out

    # ---- Lindblad limit ----------------------------------------------------------------------------
    def upload_lindblad(self, bound):
        """Device tables of a ``lindblad.BoundLindblad`` (ELL generator of one block)."""
        return dict(lindblad=bound, E=int(bound.ell_idx.shape[1]), width=int(bound.width),
                    idx=self.to_device(bound.ell_idx, torch.int32), val=self.to_device(bound.ell_val),
                    norm1=float(bound.norm1))

    def evolve_lindblad(self, prog: dict, sigma_in: torch.Tensor, n_inst: int, seg_dt, obs: dict | None = None,
                        obs_real: bool = False, snapshots: bool = False, inst_src=None, terms: int = 0, theta: float = 0.0):
        """Propagate ``n_inst`` blocks under the master equation; emission k follows segment k of duration
        ``seg_dt[k]``.  Returns (out_obs | None, out_snap | None) shaped like ``evolve``'s."""
        E = prog['E']
        seg = np.ascontiguousarray(seg_dt, dtype=np.float64)
        n_seg = int(seg.size)
        assert sigma_in.dtype == torch.complex128 and sigma_in.is_contiguous() and sigma_in.numel() % E == 0
        args = cabi.QsxLindbladArgs()
        args.E, args.width = E, prog['width']
        args.ell_idx, args.ell_val, args.norm1 = prog['idx'].data_ptr(), prog['val'].data_ptr(), prog['norm1']
        args.terms, args.theta = int(terms), float(theta)
        args.n_inst = int(n_inst)
        args.inst_src = None if inst_src is None else inst_src.data_ptr()
        args.sigma_in = sigma_in.data_ptr()
        args.n_seg, args.seg_dt = n_seg, seg.ctypes.data_as(C.POINTER(C.c_double))
        out_obs = out_snap = None
        if obs is not None:
            args.n_obs, args.obs_ptr, args.obs_idx, args.obs_w = obs['n'], obs['ptr'].data_ptr(), obs['idx'].data_ptr(), obs['w'].data_ptr()
            out_obs = torch.empty((n_inst, n_seg, obs['n']), dtype=torch.float64 if obs_real else torch.complex128,
                                  device=self.device)
            args.out_obs, args.obs_real = out_obs.data_ptr(), int(obs_real)
        if snapshots:
            out_snap = torch.empty((n_inst, n_seg, E), dtype=torch.complex128, device=self.device)
            args.out_snap = out_snap.data_ptr()
        with torch.cuda.device(self.device):
            ws_bytes = self.lib.qsx_lindblad_workspace_bytes(C.byref(args))
            cabi.check(ws_bytes)
            workspace = torch.empty(max(ws_bytes // 8, 1), dtype=torch.float64, device=self.device)
            args.workspace = workspace.data_ptr()
            stream = torch.cuda.current_stream(self.device).cuda_stream
            cabi.check(self.lib.qsx_lindblad_evolve(C.byref(args), C.c_void_p(stream)))
        if n_inst > 0 and n_seg > 0:
            self.launches += 1
        return out_obs, out_snap

    def rotating_frame(self, signal: torch.Tensor, t1: torch.Tensor, t3, t2_index: int, pad: int, rf_freq: float,
                       damping: float, out: torch.Tensor | None = None):
        """Rotating frame x damping x zero padding (postprocessing.py:9-37, :57, :77).  ``signal`` complex128 [n1]
        (linear, ``t3`` None) or [n1, n2, n3]; returns [n1*pad] or [n1*pad, n3*pad]."""
        assert signal.dtype == torch.complex128 and signal.is_contiguous() and signal.dim() in (1, 3)
        linear = signal.dim() == 1
        args = cabi.QsxFrameArgs()
        args.n1 = signal.shape[0]
        args.n2, args.n3 = (1, 1) if linear else (signal.shape[1], signal.shape[2])
        args.t2_index, args.pad, args.linear = int(t2_index), int(pad), int(linear)
        args.rf_freq, args.damping = float(rf_freq), float(damping)
        args.t1 = t1.data_ptr()
        args.t3 = None if linear else t3.data_ptr()
        shape = (args.n1 * pad,) if linear else (args.n1 * pad, args.n3 * pad)
        if out is None:
            out = torch.empty(shape, dtype=torch.complex128, device=self.device)
        assert tuple(out.shape) == shape and out.dtype == torch.complex128 and out.is_contiguous()
        args.signal, args.out = signal.data_ptr(), out.data_ptr()
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            cabi.check(self.lib.qsx_rotating_frame(C.byref(args), C.c_void_p(stream)))
        if out.numel():
            self.launches += 1
        return out

    def spectrum_shift(self, data: torch.Tensor, s1: int, s3: int, scale: float, out: torch.Tensor | None = None):
        """out[i, j] = scale * data[(i + s1) % m1, (j + s3) % m3]  (fftshift / ifftshift + prefactor in one pass)."""
        assert data.dtype == torch.complex128 and data.is_contiguous() and data.dim() in (1, 2)
        m1, m3 = (data.shape[0], 1) if data.dim() == 1 else data.shape
        if out is None:
            out = torch.empty_like(data)
        assert out.shape == data.shape and out.dtype == data.dtype and out.is_contiguous() and out.data_ptr() != data.data_ptr()
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            cabi.check(self.lib.qsx_spectrum_shift(C.c_void_p(data.data_ptr()), C.c_void_p(out.data_ptr()), m1, m3,
                                                   s1 % max(m1, 1), s3 % max(m3, 1), float(scale), C.c_void_p(stream)))
        if out.numel():
            self.launches += 1
        return out

    def readout_gemm(self, sigma: torch.Tensor, readouts: torch.Tensor) -> torch.Tensor:
        """out[i, k] = sum_e sigma[i, e] readouts[k, e] (complex128, no conjugation; ``qsx_readout_gemm``).  With a
        leading batch axis on both operands (``sigma`` [S, n, E], ``readouts`` [S, m, E]: the parameter sets of an
        ensemble, each with its own propagated read-outs) the S products run in one launch -> [S, n, m]."""
        assert sigma.dtype == readouts.dtype == torch.complex128 and sigma.is_contiguous() and readouts.is_contiguous()
        batched = sigma.dim() == 3
        assert readouts.dim() == sigma.dim() and (not batched or readouts.shape[0] == sigma.shape[0])
        n_batch = sigma.shape[0] if batched else 1
        n_inst, E = sigma.shape[-2:]
        n_emit = readouts.shape[-2]
        assert readouts.shape[-1] == E
        out = torch.empty(((n_batch,) if batched else ()) + (n_inst, n_emit), dtype=torch.complex128, device=self.device)
        with torch.cuda.device(self.device):
            ws_bytes = self.lib.qsx_readout_gemm_batched_workspace_bytes(n_batch, n_inst, n_emit, E)
            cabi.check(ws_bytes)
            workspace = torch.empty(ws_bytes // 8, dtype=torch.float64, device=self.device) if ws_bytes else None
            stream = torch.cuda.current_stream(self.device).cuda_stream
            with self.span('readout_gemm {}x{}x{}'.format(n_batch * n_inst, n_emit, E),
                           flops=8.0 * n_batch * n_inst * n_emit * E):
                cabi.check(self.lib.qsx_readout_gemm_batched(
                    n_batch, n_inst, n_emit, E, C.c_void_p(sigma.data_ptr()), C.c_void_p(readouts.data_ptr()),
                    C.c_void_p(out.data_ptr()), C.c_void_p(workspace.data_ptr() if workspace is not None else None),
                    C.c_void_p(stream)))
        if n_batch > 0 and n_inst > 0 and n_emit > 0:
            self.launches += 1
        return out

    def fp64_peak(self, repeats: int = 5) -> float:
        """Measured DFMA throughput of the CUDA cores (TFLOP/s) -- the FP64 roofline denominator."""
        out = C.c_double()
        with torch.cuda.device(self.device):
            cabi.check(self.lib.qsx_fp64_peak(repeats, C.byref(out)))
        return out.value


_ENGINE = None


def get_engine() -> Engine:
    """Process-wide engine on the current CUDA device."""
    global _ENGINE
    if _ENGINE is None or _ENGINE.device.index != torch.cuda.current_device():
        _ENGINE = Engine()
    return _ENGINE
