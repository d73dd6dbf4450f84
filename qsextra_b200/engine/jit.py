"""Run-time specialisation of the CTA-wide register-resident kernel (csrc/qsx_rrc.cuh, csrc/qsx_rrc2.cuh).

libqsx.so ships compile-time programs for a menu of block structures (csrc/qsx_rrc_programs.inc: the 3- and 4-site
chains of the BASELINE configurations).  Any other eligible structure -- rings, chains with a missing bond, long-range
couplings -- gets the SAME kernels, specialised for its own operation list the first time it is used:

  1. ``RRCProgram.source()`` writes the ``BigProg<0>`` table of the program (the emitter of the shipped menu);
  2. NVRTC (libnvrtc through ctypes, no host compiler involved) compiles csrc/qsx_rrc.cuh + csrc/qsx_rrc2.cuh around
     it for sm_100a, with -lineinfo, into a cubin; cubins are cached on disk by the hash of everything that went in;
  3. ``qsx_rrc_register`` (C ABI) loads the cubin and files the kernels under the program's signature, after which
     ``qsx_evolve_rrc`` / ``qsx_rrc_find_program`` treat the signature like a shipped one.

The first use of a structure costs a few seconds (logged with QSX_JIT_VERBOSE=1), later processes hit the disk cache.
``QSX_DISABLE_JIT`` switches the mechanism off (the shared-memory pass kernel then serves such blocks);
``QSX_JIT_CACHE`` moves the cache (default ~/.cache/qsextra_b200).
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import sys
import time

import numpy as np

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'csrc')
INCLUDE = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'include')
ARCH = 'sm_100a'
_NVRTC = None
_REGISTERED: dict = {}                    # signature bytes -> True (registered) / False (cannot be specialised)

UNIT = '''#define QSX_RRC_EXTERNAL_PROGRAMS "qsx_jit_program.inc"
#include "qsx.h"
#include "qsx_complex.cuh"
#include "qsx_rrc.cuh"
#include "qsx_rrc2.cuh"
// what the host needs to size a launch: threads, exchange-buffer elements of the two forms, operations
extern "C" __constant__ int qsx_jit_info[6] = {rrc::Geo<0>::NT, 2 * rrc::Geo<0>::NGROUP * rrc::Geo<0>::BUF,
                                               %(eligible)d, rrc2::Geo<0>::BUF, rrc::BigProg<0>::NOPS,
                                               rrc2::Geo<0>::MIN_CTAS};
template __global__ void rrc::evolve_rrc_kernel<0>(const rrc::Params);
%(second)s
'''
SECOND = 'template __global__ void rrc2::evolve_rrc2_kernel<0>(const rrc::Params);'
NAME1 = 'rrc::evolve_rrc_kernel<0>'
NAME2 = 'rrc2::evolve_rrc2_kernel<0>'


class JitError(RuntimeError):
    pass


def nvrtc():
    """libnvrtc of the CUDA toolkit, loaded on first use."""
    global _NVRTC
    if _NVRTC is not None:
        return _NVRTC
    last = None
    for name in ('libnvrtc.so.12', '/usr/local/cuda/lib64/libnvrtc.so.12', 'libnvrtc.so', '/usr/local/cuda/lib64/libnvrtc.so'):
        try:
            _NVRTC = C.CDLL(name)
            break
        except OSError as exc:
            last = exc
    if _NVRTC is None:
        raise JitError('libnvrtc not found: {}'.format(last))
    return _NVRTC


def _check(lib, code, what):
    if code != 0:
        lib.nvrtcGetErrorString.restype = C.c_char_p
        raise JitError('{}: {}'.format(what, lib.nvrtcGetErrorString(code).decode()))


def cuda_include_dir():
    for root in (os.environ.get('CUDA_HOME'), '/usr/local/cuda'):
        if root and os.path.exists(os.path.join(root, 'include', 'cuda', 'std', 'utility')):
            return os.path.join(root, 'include')
    return '/usr/local/cuda/include'


def compile_unit(program_source: str, second_form: bool):
    """-> (cubin bytes, lowered name of the first form, lowered name of the throughput form or None, log)."""
    lib = nvrtc()
    unit = UNIT % dict(eligible=int(second_form), second=SECOND if second_form else '')
    header = ('namespace rrc_jit_guard {}\n' + program_source + 'constexpr int kNumBigProgs = 1;\n').encode()
    prog = C.c_void_p()
    names = (C.c_char_p * 1)(b'qsx_jit_program.inc')
    texts = (C.c_char_p * 1)(header)
    _check(lib, lib.nvrtcCreateProgram(C.byref(prog), unit.encode(), b'qsx_jit_unit.cu', 1, texts, names), 'nvrtcCreateProgram')
    try:
        wanted = [NAME1] + ([NAME2] if second_form else [])
        for expr in wanted:
            _check(lib, lib.nvrtcAddNameExpression(prog, expr.encode()), 'nvrtcAddNameExpression')
        options = ['--gpu-architecture=' + ARCH, '-std=c++17', '-lineinfo', '-default-device', '-I' + CSRC, '-I' + INCLUDE,
                   '-I' + cuda_include_dir()]
        opts = (C.c_char_p * len(options))(*[o.encode() for o in options])
        rc = lib.nvrtcCompileProgram(prog, len(options), opts)
        size = C.c_size_t()
        lib.nvrtcGetProgramLogSize(prog, C.byref(size))
        log = C.create_string_buffer(max(size.value, 1))
        lib.nvrtcGetProgramLog(prog, log)
        if rc != 0:
            raise JitError('NVRTC could not compile the specialisation:\n' + log.value.decode(errors='replace')[-4000:])
        lowered = []
        for expr in wanted:
            out = C.c_char_p()
            _check(lib, lib.nvrtcGetLoweredName(prog, expr.encode(), C.byref(out)), 'nvrtcGetLoweredName')
            lowered.append(out.value.decode())
        _check(lib, lib.nvrtcGetCUBINSize(prog, C.byref(size)), 'nvrtcGetCUBINSize')
        cubin = C.create_string_buffer(size.value)
        _check(lib, lib.nvrtcGetCUBIN(prog, cubin), 'nvrtcGetCUBIN')
        return cubin.raw, lowered[0], lowered[1] if second_form else None, log.value.decode(errors='replace')
    finally:
        lib.nvrtcDestroyProgram(C.byref(prog))


def _inputs_digest(program_source: str, second_form: bool) -> str:
    h = hashlib.sha256()
    h.update(program_source.encode())
    h.update(bytes([second_form]))
    h.update(ARCH.encode())
    for path in (os.path.join(CSRC, 'qsx_rrc.cuh'), os.path.join(CSRC, 'qsx_rrc2.cuh'), os.path.join(CSRC, 'qsx_complex.cuh'),
                 os.path.join(INCLUDE, 'qsx.h')):
        h.update(open(path, 'rb').read())
    major, minor = C.c_int(), C.c_int()
    nvrtc().nvrtcVersion(C.byref(major), C.byref(minor))
    h.update(('%d.%d' % (major.value, minor.value)).encode())
    h.update(UNIT.encode())
    return h.hexdigest()[:32]


def cache_dir() -> str:
    return os.environ.get('QSX_JIT_CACHE') or os.path.join(os.path.expanduser('~'), '.cache', 'qsextra_b200')


def specialise(big):
    """Cubin and kernel names for an ``RRCProgram`` (disk cache first).  -> (cubin, name1, name2 | None, from_cache)."""
    source = big.source(0, 'run-time specialisation')
    second = big.throughput_form_eligible()
    key = _inputs_digest(source, second)
    folder = cache_dir()
    path = os.path.join(folder, key + '.cubin')
    meta = os.path.join(folder, key + '.names')
    if os.path.exists(path) and os.path.exists(meta):
        try:
            names = open(meta).read().split('\n')
            return open(path, 'rb').read(), names[0], (names[1] or None), True
        except OSError:
            pass
    t0 = time.perf_counter()
    cubin, name1, name2, _ = compile_unit(source, second)
    if os.environ.get('QSX_JIT_VERBOSE'):
        print('[qsextra_b200.jit] block ({}, {}) {} x {}: {} operations specialised in {:.1f} s ({} bytes)'.format(
            big.block.ka, big.block.kb, big.block.dim_a, big.block.dim_b, len(big.ops), time.perf_counter() - t0, len(cubin)),
            file=sys.stderr)
    try:
        os.makedirs(folder, exist_ok=True)
        tmp = path + '.%d.tmp' % os.getpid()
        open(tmp, 'wb').write(cubin)
        os.replace(tmp, path)
        open(meta, 'w').write(name1 + '\n' + (name2 or ''))
    except OSError:
        pass                                            # a read-only home: compile again next time
    return cubin, name1, name2, False


def ensure_registered(lib, big, signature: np.ndarray) -> bool:
    """Make ``qsx_evolve_rrc`` able to run ``big`` (an ``RRCProgram`` without a shipped compile-time program).
    -> False when the structure cannot be specialised (layout outside the kernel's limits): use the pass kernel."""
    if os.environ.get('QSX_DISABLE_JIT'):
        return False
    key = signature.tobytes()
    known = _REGISTERED.get(key)
    if known is not None:
        return known
    from . import cabi
    try:
        cubin, name1, name2, _ = specialise(big)
    except JitError as exc:
        if os.environ.get('QSX_JIT_VERBOSE'):
            print('[qsextra_b200.jit] not specialised: {}'.format(str(exc)[:600]), file=sys.stderr)
        _REGISTERED[key] = False
        return False
    sig = np.ascontiguousarray(signature, dtype=np.int32)
    rc = lib.qsx_rrc_register(cubin, len(cubin), name1.encode(), name2.encode() if name2 else None,
                              C.c_void_p(sig.ctypes.data), int(sig.size))
    cabi.check(rc)
    _REGISTERED[key] = True
    return True
