"""The C-ABI shared library loads and exports every symbol include/qsx.h declares (no compute calls: no GPU here),
and the ctypes mirrors have the layout the header declares."""
import ctypes
import os
import re
import subprocess

import pytest

from qsextra_b200.engine import cabi

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, 'include', 'qsx.h')


def declared_functions():
    text = open(HEADER).read()
    return sorted(set(re.findall(r'^\s*(?:const\s+char\*|int32_t|int64_t)\s+(qsx_\w+)\s*\(', text, flags=re.M)))


def test_library_exports_every_declared_symbol():
    lib = cabi.load()
    names = declared_functions()
    assert set(names) == set(cabi.EXPORTS)
    for name in names:
        assert hasattr(lib, name), name
    assert lib.qsx_abi_version() == cabi.ABI_VERSION
    assert lib.qsx_last_error() == b''


def test_struct_layout_matches_header(tmp_path):
    """Compile a tiny C program against include/qsx.h and compare sizeof() and the offset of the last field of every
    struct with the ctypes mirrors."""
    pairs = [('qsx_op', cabi.QsxOp, 'flags'), ('qsx_pass', cabi.QsxPass, 'f'), ('qsx_block', cabi.QsxBlock, 'rankB'),
             ('qsx_evolve_args', cabi.QsxEvolveArgs, 'hops'), ('qsx_rr_op', cabi.QsxRROp, 'aux1'),
             ('qsx_rr_args', cabi.QsxRRArgs, 'out_snap'), ('qsx_rrc_args', cabi.QsxRRCArgs, 'progress'),
             ('qsx_rr_tables_args', cabi.QsxRRTablesArgs, 'out'), ('qsx_lindblad_args', cabi.QsxLindbladArgs, 'workspace'),
             ('qsx_frame_args', cabi.QsxFrameArgs, 'out'), ('qsx_system_desc', cabi.QsxSystemDesc, 'trotter_number')]
    body = ''.join('printf("%zu %zu\\n", sizeof({0}), offsetof({0}, {1}));'.format(c_name, last) for c_name, _, last in pairs)
    src = tmp_path / 'sizes.c'
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "qsx.h"\nint main(){' + body + ' return 0;}\n')
    exe = tmp_path / 'sizes'
    subprocess.check_call(['gcc', '-I', os.path.join(REPO, 'include'), str(src), '-o', str(exe)])
    rows = [[int(v) for v in line.split()] for line in subprocess.check_output([str(exe)]).decode().splitlines()]
    for (c_name, mirror, last), (size, offset) in zip(pairs, rows):
        assert ctypes.sizeof(mirror) == size, c_name
        assert getattr(mirror, last).offset == offset, (c_name, last)


def test_invalid_arguments_fail_loudly_without_a_device():
    lib = cabi.load()
    assert lib.qsx_evolve(None, None) == -1 and b'null' in lib.qsx_last_error()
    out = ctypes.c_double()
    assert lib.qsx_fp64_peak(1, None) == -1
    with pytest.raises(cabi.QsxError):
        cabi.check(-1)


def test_engine_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from qsextra_b200.engine.runtime import Engine
    with pytest.raises(cabi.QsxError):
        Engine()


# ---- host-side program builder of the C ABI (qsx_step_program_size / _build): no device needed ---------------------
def c_built_program(system, ka, kb, dt, coll_rates=None, trotter_number=1):
    """-> (ops [n, 8] int32, params [n_params], cfg_a, cfg_b, n_bits) from the C builder for a qsextra system object."""
    import numpy as np
    lib = cabi.load()
    n = system.system_size
    chromophore = getattr(system, 'mode_dict', None) is not None
    w = len(system.mode_dict['omega_mode']) if chromophore else 0
    eps = np.ascontiguousarray(system.e_el, dtype=float)
    J = np.ascontiguousarray(system.coupl_el, dtype=float)
    omega = np.ascontiguousarray(system.mode_dict['omega_mode'], dtype=float) if chromophore else None
    g = np.ascontiguousarray(system.mode_dict['coupl_ep'], dtype=float) if chromophore else None
    rates = None if coll_rates is None else np.ascontiguousarray(np.atleast_1d(coll_rates), dtype=float)
    desc = cabi.QsxSystemDesc(n, w, eps.ctypes.data, J.ctypes.data, None if omega is None else omega.ctypes.data,
                              None if g is None else g.ctypes.data, None if rates is None else rates.ctypes.data, dt,
                              trotter_number)
    sizes = [ctypes.c_int32() for _ in range(5)]
    cabi.check(lib.qsx_step_program_size(ctypes.byref(desc), ka, kb, *[ctypes.byref(v) for v in sizes]))
    n_ops, n_params, n_a, n_b, n_bits = [v.value for v in sizes]
    ops = np.zeros((n_ops, 8), dtype=np.int32)
    params = np.zeros(n_params)
    cfg_a, cfg_b = np.zeros(n_a, dtype=np.int32), np.zeros(n_b, dtype=np.int32)
    rank_a, rank_b = np.zeros(1 << n, dtype=np.int32), np.zeros(1 << n, dtype=np.int32)
    cabi.check(lib.qsx_step_program_build(ctypes.byref(desc), ka, kb, ops.ctypes.data, params.ctypes.data, cfg_a.ctypes.data,
                                          cfg_b.ctypes.data, rank_a.ctypes.data, rank_b.ctypes.data))
    return ops, params, cfg_a, cfg_b, n_bits


def _builder_cases():
    from helpers import build_system
    dimer = dict(energies=[1.55, 1.46], couplings=[[0, -0.3], [-0.3, 0]], dipole_moments=[1., 1.])
    chain = dict(energies=[1.5, 1.57, 1.64, 1.71], couplings=[[0, -.3, 0, .1], [-.3, 0, -.33, 0], [0, -.33, 0, -.36], [.1, 0, -.36, 0]],
                 dipole_moments=[1.] * 4)
    modes1 = dict(frequencies_pseudomode=[0.3], levels_pseudomode=[2], couplings_ep=[0.7])
    modes2 = dict(frequencies_pseudomode=[0.3, 0.0], levels_pseudomode=[2, 2], couplings_ep=[0.7, 0.4])
    return [('dimer, one mode, damping', build_system({**dimer, **modes1}), [0.5], 1, [(1, 1), (1, 0), (2, 1)]),
            ('dimer, two modes, one undamped', build_system({**dimer, **modes2}), [0.5, 0.0], 1, [(1, 1), (0, 1)]),
            ('chain of four, one mode, two Trotter steps', build_system({**chain, **modes1}), [0.4], 2, [(1, 1), (2, 1)]),
            ('chain of four, dephasing', build_system(chain), 0.2, 1, [(1, 1), (2, 2), (1, 0)]),
            ('dimer without collisions', build_system(dimer), None, 3, [(1, 1)])]


@pytest.mark.parametrize('case', _builder_cases(), ids=lambda c: c[0])
def test_c_program_builder_equals_the_python_lowering(case):
    """One step of the program the C builder writes equals one step of the Python lowering (both through the emulated
    operation semantics of tests/emulator.py, which the CPU suite ties to the reference goldens) on random blocks."""
    import types
    import numpy as np
    import emulator
    from qsextra_b200.qcomo.evolution.qevolve import _program_for, _validate_rates
    name, system, rates, trotter_number, blocks = case
    dt = 0.15
    prog = _program_for(system, dt, dt / trotter_number, trotter_number, 1, _validate_rates(system, rates))
    rng = np.random.default_rng(5)
    for ka, kb in blocks:
        blk = prog.block(ka, kb)
        ops, params, cfg_a, cfg_b, n_bits = c_built_program(system, ka, kb, dt, rates, trotter_number)
        assert list(cfg_a) == list(blk.cfg_a) and list(cfg_b) == list(blk.cfg_b) and n_bits == blk.n_bits
        from_c = types.SimpleNamespace(block=blk, ops=ops, params=params[None, :])
        sigma = rng.normal(size=(blk.dim_a, blk.dim_b)) + 1j * rng.normal(size=(blk.dim_a, blk.dim_b))
        want = emulator.apply_step(sigma, prog.bind(blk, fused=False))
        got = emulator.apply_step(sigma, from_c)
        assert np.abs(got - want).max() < 1e-13 * np.abs(want).max(), (name, ka, kb)


def test_c_program_builder_rejects_bad_descriptions():
    lib = cabi.load()
    assert lib.qsx_step_program_size(None, 1, 1, None, None, None, None, None) == -1
    import numpy as np
    eps, J = np.zeros(2), np.zeros(4)
    desc = cabi.QsxSystemDesc(2, 0, eps.ctypes.data, J.ctypes.data, None, None, None, 0.0, 1)
    assert lib.qsx_step_program_size(ctypes.byref(desc), 1, 1, None, None, None, None, None) == -1      # dt = 0
    desc.dt = 0.1
    assert lib.qsx_step_program_size(ctypes.byref(desc), 3, 1, None, None, None, None, None) == -1      # ka > N
