"""Build libqsx.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repository snapshot)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
SOURCES = ["qsx_engine.cu"]
OUT = os.path.join(HERE, 'libqsx.so')


def build(force: bool = False, verbose: bool = True) -> str:
    srcs = [os.path.join(HERE, s) for s in SOURCES]
    deps = srcs + [os.path.join(REPO, 'include', 'qsx.h')] + [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(('.cuh', '.inc'))]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
           '-shared', '-Xcompiler', '-fPIC', '-I' + os.path.join(REPO, 'include'), '-o', OUT] + srcs
    if verbose:
        print(' '.join(cmd))
    subprocess.check_call(cmd)
    return OUT


if __name__ == '__main__':
    build(force='--force' in sys.argv)
