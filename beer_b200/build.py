"""Build libbeer_b200.so (sm_100a only) in-tree with nvcc.

    python -m beer_b200.build [--force]

The shared library is the product: a C-ABI (include/beer_b200.h) over hand-written
CUDA kernels.  It is built into beer_b200/lib/ so that it travels with the repo
snapshot to the GPU box; nothing is JIT-compiled at import time.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
LIBPATH = os.path.join(LIBDIR, 'libbeer_b200.so')
SOURCES = ['dists.cu', 'emission.cu', 'emission_tc.cu', 'scan.cu', 'accumulate.cu', 'accumulate_tc.cu', 'features.cu', 'transitions.cu', 'chains.cu', 'probe.cu', 'mix16.cu', 'emission_bwd.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden']


def _nvcc():
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found: libbeer_b200.so cannot be built on this machine')
    return nvcc


def _newest_source_mtime():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    files.append(os.path.join(os.path.dirname(HERE), 'include', 'beer_b200.h'))
    return max(os.path.getmtime(f) for f in files)


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link the shared library."""
    if (not force and os.path.exists(LIBPATH)
            and os.path.getmtime(LIBPATH) >= _newest_source_mtime()):
        return LIBPATH
    nvcc = _nvcc()
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace('.cu', '.o'))
        cmd = [nvcc] + NVCC_FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj]
        if verbose:
            cmd.insert(1, '-Xptxas=-v')
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{res.stdout}\n{res.stderr}')
        if verbose:
            print(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    cmd = [nvcc, '-shared', '-o', LIBPATH] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a',
                                                     '-cudart', 'static']
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f'link failed:\n{res.stdout}\n{res.stderr}')
    return LIBPATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
