"""Builds libptta_b200.so in-tree with nvcc for sm_100a (no torch headers: the library is plain C ABI).
Every .cu under csrc/ is one translation unit, compiled to an object only when it or a header is newer, then linked."""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, 'csrc')
LIB_DIR = os.path.join(PKG, 'lib')
LIB_PATH = os.environ.get('PTTA_B200_LIB') or os.path.join(LIB_DIR, 'libptta_b200.so')     # override: A/B experiments with another build
SOURCES = ['engine.cu', 'nlspn_net.cu', 'png_host.cu']      # png_host.cu: host-only code (PNG container + zlib inflate), links libz
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC']


def _header_mtime():
    t = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(PKG), 'include')):
        for f in os.listdir(root):
            if f.endswith(('.cuh', '.h')):
                t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


def _obj(src):
    return os.path.join(LIB_DIR, src[:-3] + '.o')


def _stale(src, hdr_t):
    o = _obj(src)
    return not os.path.exists(o) or os.path.getmtime(o) < max(hdr_t, os.path.getmtime(os.path.join(CSRC, src)))


def needs_build():
    hdr_t = _header_mtime()
    if not os.path.exists(LIB_PATH):
        return True
    lib_t = os.path.getmtime(LIB_PATH)
    return any(lib_t < max(hdr_t, os.path.getmtime(os.path.join(CSRC, s))) for s in SOURCES)


def build_library(force=False, verbose=False):
    """Compile the CUDA sources into lib/libptta_b200.so.  Raises if nvcc is missing or fails."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found: cannot build libptta_b200.so')
    os.makedirs(LIB_DIR, exist_ok=True)
    hdr_t = _header_mtime()
    procs = []
    for s in SOURCES:
        if force or _stale(s, hdr_t):
            cmd = [nvcc] + NVCC_FLAGS + ['-c', '-o', _obj(s), os.path.join(CSRC, s)]
            if verbose:
                print(' '.join(cmd))
            procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError('nvcc failed:\n' + out)
        if verbose and out.strip():
            print(out)
    cmd = [nvcc, '-shared', '-o', LIB_PATH] + [_obj(s) for s in SOURCES] + ['-lz']
    if verbose:
        print(' '.join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('link failed:\n' + res.stdout + res.stderr)
    return LIB_PATH


def build_stamps_library(verbose=False):
    """Profiling build (-DPTTA_STAMPS): every kernel records its in-situ start time (tools/graph_stamps.py).  Not the product."""
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    out = os.path.join(LIB_DIR, 'libptta_b200_stamps.so')
    cmd = [nvcc] + NVCC_FLAGS + ['-DPTTA_STAMPS', '-shared', '-o', out] + [os.path.join(CSRC, s) for s in SOURCES] + ['-lz']
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + res.stdout + res.stderr)
    return out


def build_experiments_library(verbose=False):
    """Experiments build (-DPTTA_EXPERIMENTS): convg_kernel with the work-skipping timing switches and cycle stamps of
    tools/convg_experiment.py / tools/convg_trace.py (PTTA_B200_LIB=.../libptta_b200_experiments.so).  Not the product."""
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    out = os.path.join(LIB_DIR, 'libptta_b200_experiments.so')
    cmd = [nvcc] + NVCC_FLAGS + ['-DPTTA_EXPERIMENTS', '-shared', '-o', out] + [os.path.join(CSRC, s) for s in SOURCES] + ['-lz']
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + res.stdout + res.stderr)
    return out


if __name__ == '__main__':
    import sys
    if 'stamps' in sys.argv:
        print(build_stamps_library(verbose=True))
    elif 'experiments' in sys.argv:
        print(build_experiments_library(verbose=True))
    else:
        print(build_library(force=True, verbose=True))
