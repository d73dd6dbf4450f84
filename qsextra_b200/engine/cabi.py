"""ctypes binding of ``libqsx.so`` (C ABI declared in ``include/qsx.h``).

The shared library is built in-tree by ``qsextra_b200/csrc/build.py`` (``__graft_entry__.build()``).  There is no
CPU fallback: ``load()`` raises if the library is missing, and every compute call raises ``QsxError`` with the
library's message when it fails.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), 'csrc', 'libqsx.so')
ABI_VERSION = 20
ERR_NO_PROGRAM = -4

EXPORTS = ('qsx_abi_version', 'qsx_last_error', 'qsx_device_info', 'qsx_evolve_workspace_bytes', 'qsx_evolve',
           'qsx_evolve_rr', 'qsx_bind_params', 'qsx_rr_tables', 'qsx_evolve_rrc', 'qsx_rrc_find_program', 'qsx_rrc_register', 'qsx_step_program_size', 'qsx_step_program_build', 'qsx_apply_dipole', 'qsx_readout_gemm_workspace_bytes', 'qsx_readout_gemm',
           'qsx_readout_gemm_batched_workspace_bytes', 'qsx_readout_gemm_batched', 'qsx_rotating_frame', 'qsx_spectrum_shift',
           'qsx_lindblad_workspace_bytes', 'qsx_lindblad_evolve',
           'qsx_fp64_peak')
RR_MAX_OPS, RR_MAX_OBS, RR_MAX_ENTRIES = 24, 8, 4


class QsxError(RuntimeError):
    pass


class QsxSystemDesc(C.Structure):
    _fields_ = [('n_sites', C.c_int32), ('n_modes', C.c_int32), ('energies', C.c_void_p), ('couplings', C.c_void_p),
                ('omega', C.c_void_p), ('coupl_ep', C.c_void_p), ('coll_rates', C.c_void_p), ('dt', C.c_double),
                ('trotter_number', C.c_int32)]


class QsxOp(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ('type', 'i0', 'i1', 'i2', 'i3', 'i4', 'p', 'flags')]


class QsxPass(C.Structure):
    _fields_ = [('type', C.c_int32), ('f', C.c_int32 * 11)]


class QsxBlock(C.Structure):
    _fields_ = [('n_sites', C.c_int32), ('n_bits', C.c_int32), ('nA', C.c_int32), ('nB', C.c_int32),
                ('cfgA', C.c_void_p), ('cfgB', C.c_void_p), ('rankA', C.c_void_p), ('rankB', C.c_void_p)]


class QsxEvolveArgs(C.Structure):
    _fields_ = [('block', QsxBlock),
                ('n_ops', C.c_int32), ('ops', C.c_void_p),
                ('param_stride', C.c_int32), ('params', C.c_void_p),
                ('n_inst', C.c_int32), ('inst_set', C.c_void_p), ('inst_src', C.c_void_p),
                ('sigma_in', C.c_void_p),
                ('n_steps', C.c_int32), ('n_emit', C.c_int32), ('emit_steps', C.c_void_p),
                ('n_obs', C.c_int32), ('obs_ptr', C.c_void_p), ('obs_idx', C.c_void_p), ('obs_w', C.c_void_p),
                ('out_obs', C.c_void_p), ('obs_real', C.c_int32),
                ('out_snap', C.c_void_p), ('workspace', C.c_void_p), ('threads', C.c_int32),
                ('n_passes', C.c_int32), ('passes', C.c_void_p),
                ('n_tile_ints', C.c_int32), ('tiles', C.c_void_p),
                ('n_hop_ints', C.c_int32), ('hops', C.c_void_p)]


class QsxRROp(C.Structure):
    _fields_ = [('kind', C.c_int32), ('reg_xor', C.c_int32), ('lane_xor', C.c_int32), ('reg_mask', C.c_uint32),
                ('p', C.c_int32), ('sign', C.c_int32), ('aux0', C.c_int32), ('aux1', C.c_int32)]


class QsxRRArgs(C.Structure):
    _fields_ = [('reg_bits', C.c_int32), ('lane_bits', C.c_int32), ('n_diag', C.c_int32), ('n_ops', C.c_int32),
                ('ops', QsxRROp * RR_MAX_OPS),
                ('logical_index', C.c_void_p),
                ('param_stride', C.c_int32), ('params', C.c_void_p),
                ('n_inst', C.c_int32), ('inst_set', C.c_void_p), ('inst_src', C.c_void_p), ('sigma_in', C.c_void_p),
                ('n_steps', C.c_int32), ('n_emit', C.c_int32), ('emit_steps', C.c_void_p),
                ('n_obs', C.c_int32), ('obs_weights', C.c_void_p), ('obs_reg_mask', C.c_uint32 * RR_MAX_OBS),
                ('n_entries', C.c_int32 * RR_MAX_OBS),
                ('entry_reg_mask', (C.c_uint32 * RR_MAX_ENTRIES) * RR_MAX_OBS),
                ('entry_lane_mask', (C.c_uint32 * RR_MAX_ENTRIES) * RR_MAX_OBS),
                ('entry_weight', (C.c_double * RR_MAX_ENTRIES) * RR_MAX_OBS),
                ('readout_kind', C.c_int32), ('mu', C.c_double * 16),
                ('site_reg_mask', (C.c_uint32 * RR_MAX_ENTRIES) * 16), ('site_lane_mask', (C.c_uint32 * RR_MAX_ENTRIES) * 16),
                ('out_obs', C.c_void_p), ('obs_real', C.c_int32), ('out_snap', C.c_void_p)]


class QsxRRCArgs(C.Structure):
    _fields_ = [('n_signature', C.c_int32), ('signature', C.c_void_p),
                ('param_stride', C.c_int32), ('params', C.c_void_p),
                ('n_inst', C.c_int32), ('inst_set', C.c_void_p), ('inst_src', C.c_void_p), ('sigma_in', C.c_void_p),
                ('n_steps', C.c_int32), ('n_emit', C.c_int32), ('emit_steps', C.c_void_p),
                ('emit_trace', C.c_int32), ('mu', C.c_double * 16),
                ('out_obs', C.c_void_p), ('out_snap', C.c_void_p), ('progress', C.c_void_p)]


_lib = None


class QsxRRTablesArgs(C.Structure):
    _fields_ = [('n_sets', C.c_int32), ('n_elem', C.c_int32), ('n_bits', C.c_int32), ('n_cfg_b', C.c_int32),
                ('n_diag', C.c_int32), ('slot_ga', C.c_int32 * 2), ('slot_gb', C.c_int32 * 2), ('slot_t', C.c_int32 * 2),
                ('n_classes', C.c_int32), ('class_slot', C.c_int32 * 8), ('class_count', C.c_void_p),
                ('n_tangent', C.c_int32), ('tangent_slot', C.c_int32 * 24), ('a_of', C.c_void_p), ('b_of', C.c_void_p),
                ('op_stride', C.c_int32), ('op_params', C.c_void_p), ('out', C.c_void_p)]


class QsxLindbladArgs(C.Structure):
    _fields_ = [('E', C.c_int32), ('width', C.c_int32), ('ell_idx', C.c_void_p), ('ell_val', C.c_void_p),
                ('norm1', C.c_double), ('terms', C.c_int32), ('theta', C.c_double), ('n_inst', C.c_int32),
                ('inst_src', C.c_void_p), ('sigma_in', C.c_void_p), ('n_seg', C.c_int32),
                ('seg_dt', C.POINTER(C.c_double)), ('n_obs', C.c_int32), ('obs_ptr', C.c_void_p),
                ('obs_idx', C.c_void_p), ('obs_w', C.c_void_p), ('obs_real', C.c_int32), ('out_obs', C.c_void_p),
                ('out_snap', C.c_void_p), ('workspace', C.c_void_p)]


class QsxFrameArgs(C.Structure):
    _fields_ = [('n1', C.c_int32), ('n2', C.c_int32), ('n3', C.c_int32), ('t2_index', C.c_int32), ('pad', C.c_int32),
                ('linear', C.c_int32), ('rf_freq', C.c_double), ('damping', C.c_double), ('t1', C.c_void_p),
                ('t3', C.c_void_p), ('signal', C.c_void_p), ('out', C.c_void_p)]


def load():
    """Load libqsx.so once; raise (never fall back) when it is absent or has the wrong ABI."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise QsxError('CUDA extension {} is missing: run `python -c "import __graft_entry__ as g; g.build()"` '
                       '(there is no CPU fallback)'.format(LIB_PATH))
    lib = C.CDLL(LIB_PATH)
    lib.qsx_abi_version.restype = C.c_int32
    lib.qsx_last_error.restype = C.c_char_p
    lib.qsx_device_info.argtypes = [C.POINTER(C.c_int32)] * 3
    lib.qsx_device_info.restype = C.c_int32
    lib.qsx_evolve_workspace_bytes.argtypes = [C.POINTER(QsxEvolveArgs)]
    lib.qsx_evolve_workspace_bytes.restype = C.c_int64
    lib.qsx_evolve.argtypes = [C.POINTER(QsxEvolveArgs), C.c_void_p]
    lib.qsx_evolve.restype = C.c_int32
    lib.qsx_evolve_rr.argtypes = [C.POINTER(QsxRRArgs), C.c_void_p]
    lib.qsx_evolve_rr.restype = C.c_int32
    lib.qsx_evolve_rrc.argtypes = [C.POINTER(QsxRRCArgs), C.c_void_p]
    lib.qsx_evolve_rrc.restype = C.c_int32
    lib.qsx_rrc_find_program.argtypes = [C.c_void_p, C.c_int32]
    lib.qsx_rrc_find_program.restype = C.c_int32
    lib.qsx_rrc_register.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_char_p, C.c_void_p, C.c_int32]
    lib.qsx_rrc_register.restype = C.c_int32
    lib.qsx_step_program_size.argtypes = [C.POINTER(QsxSystemDesc), C.c_int32, C.c_int32] + [C.POINTER(C.c_int32)] * 5
    lib.qsx_step_program_size.restype = C.c_int32
    lib.qsx_step_program_build.argtypes = [C.POINTER(QsxSystemDesc), C.c_int32, C.c_int32] + [C.c_void_p] * 6
    lib.qsx_step_program_build.restype = C.c_int32
    lib.qsx_apply_dipole.argtypes = [C.POINTER(QsxBlock), C.POINTER(QsxBlock), C.c_int32, C.c_void_p, C.c_int32,
                                     C.c_void_p, C.c_void_p, C.c_void_p]
    lib.qsx_apply_dipole.restype = C.c_int32
    lib.qsx_readout_gemm_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    lib.qsx_readout_gemm_workspace_bytes.restype = C.c_int64
    lib.qsx_readout_gemm.argtypes = [C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 5
    lib.qsx_readout_gemm.restype = C.c_int32
    lib.qsx_readout_gemm_batched_workspace_bytes.argtypes = [C.c_int32] * 4
    lib.qsx_readout_gemm_batched_workspace_bytes.restype = C.c_int64
    lib.qsx_readout_gemm_batched.argtypes = [C.c_int32] * 4 + [C.c_void_p] * 5
    lib.qsx_readout_gemm_batched.restype = C.c_int32
    lib.qsx_bind_params.argtypes = [C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 7
    lib.qsx_bind_params.restype = C.c_int32
    lib.qsx_rr_tables.argtypes = [C.POINTER(QsxRRTablesArgs), C.c_void_p]
    lib.qsx_rr_tables.restype = C.c_int32
    lib.qsx_lindblad_workspace_bytes.argtypes = [C.POINTER(QsxLindbladArgs)]
    lib.qsx_lindblad_workspace_bytes.restype = C.c_int64
    lib.qsx_lindblad_evolve.argtypes = [C.POINTER(QsxLindbladArgs), C.c_void_p]
    lib.qsx_lindblad_evolve.restype = C.c_int32
    lib.qsx_rotating_frame.argtypes = [C.POINTER(QsxFrameArgs), C.c_void_p]
    lib.qsx_rotating_frame.restype = C.c_int32
    lib.qsx_spectrum_shift.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double,
                                       C.c_void_p]
    lib.qsx_spectrum_shift.restype = C.c_int32
    lib.qsx_fp64_peak.argtypes = [C.c_int32, C.POINTER(C.c_double)]
    lib.qsx_fp64_peak.restype = C.c_int32
    if lib.qsx_abi_version() != ABI_VERSION:
        raise QsxError('libqsx.so ABI {} != expected {}'.format(lib.qsx_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(rc: int):
    if rc < 0:
        raise QsxError('libqsx: {} (code {})'.format(load().qsx_last_error().decode(), rc))
    return rc
