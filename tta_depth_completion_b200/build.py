"""Builds libptta_b200.so in-tree with nvcc for sm_100a (no torch headers: the library is plain C ABI)."""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, 'csrc')
LIB_DIR = os.path.join(PKG, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libptta_b200.so')
SOURCES = ['engine.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared']


def _newest_source_mtime():
    t = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(PKG), 'include')):
        for f in os.listdir(root):
            if f.endswith(('.cu', '.cuh', '.h')):
                t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


def needs_build():
    return not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < _newest_source_mtime()


def build_library(force=False, verbose=False):
    """Compile the CUDA sources into lib/libptta_b200.so.  Raises if nvcc is missing or fails."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found: cannot build libptta_b200.so')
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + ['-o', LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        print(' '.join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + res.stdout + res.stderr)
    return LIB_PATH


if __name__ == '__main__':
    print(build_library(force=True, verbose=True))
