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
             ('qsx_rr_args', cabi.QsxRRArgs, 'out_snap'), ('qsx_rrc_args', cabi.QsxRRCArgs, 'out_snap'),
             ('qsx_rr_tables_args', cabi.QsxRRTablesArgs, 'out'), ('qsx_lindblad_args', cabi.QsxLindbladArgs, 'workspace'),
             ('qsx_frame_args', cabi.QsxFrameArgs, 'out')]
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
