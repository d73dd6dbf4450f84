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
    """Compile a tiny C program against include/qsx.h and compare sizeof() with the ctypes mirrors."""
    src = tmp_path / 'sizes.c'
    src.write_text('#include <stdio.h>\n#include "qsx.h"\nint main(){printf("%zu %zu %zu %zu\\n", sizeof(qsx_op), '
                   'sizeof(qsx_pass), sizeof(qsx_block), sizeof(qsx_evolve_args)); return 0;}\n')
    exe = tmp_path / 'sizes'
    subprocess.check_call(['gcc', '-I', os.path.join(REPO, 'include'), str(src), '-o', str(exe)])
    sizes = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert sizes == [ctypes.sizeof(cabi.QsxOp), ctypes.sizeof(cabi.QsxPass), ctypes.sizeof(cabi.QsxBlock),
                     ctypes.sizeof(cabi.QsxEvolveArgs)]


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
