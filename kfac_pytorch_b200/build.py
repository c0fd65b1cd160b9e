"""Builds ``kfac_pytorch_b200/lib/libkfacb.so`` (the C-ABI CUDA library) in-tree with
nvcc for sm_100a.  No torch headers are involved: the library is plain CUDA/C++
behind ``include/kfacb.h`` and is bound with ctypes (``native.py``).

    python -m kfac_pytorch_b200.build [--force]
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
OBJDIR = os.path.join(HERE, 'build')
LIB = os.path.join(LIBDIR, 'libkfacb.so')
INCLUDE = os.path.join(ROOT, 'include')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-I' + INCLUDE]


def _nvcc() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found: cannot build libkfacb.so')


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _fingerprint() -> str:
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + ['../../include/kfacb.h']:
        with open(os.path.join(CSRC, f), 'rb') as fh:
            h.update(f.encode())
            h.update(fh.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def built_fingerprint() -> str:
    """Fingerprint compiled INTO the library (kfacb_build_fingerprint): travels with the .so wherever it goes."""
    if not os.path.exists(LIB):
        return ''
    try:
        import ctypes
        lib = ctypes.CDLL(LIB)
        lib.kfacb_build_fingerprint.restype = ctypes.c_char_p
        return lib.kfacb_build_fingerprint().decode()
    except Exception:
        return ''


def needs_build() -> bool:
    return built_fingerprint() != _fingerprint()


def build(force: bool = False, verbose: bool = True) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    os.makedirs(OBJDIR, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)

    fp = _fingerprint()

    def compile_one(src):
        obj = os.path.join(OBJDIR, src[:-3] + '.o')
        cmd = [nvcc] + NVCC_FLAGS + ['-DKFACB_SRC_FINGERPRINT="%s"' % fp, '-c', os.path.join(CSRC, src), '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    if verbose:
        print('built', LIB)
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv)
