#!/usr/bin/env python
"""Recipe that materialises the UNMODIFIED reference package under ``oracle/_ref/`` so that it can be timed
on the GPU box's host cores (``bench.py --impl reference``, ``cpu_baseline.kind = "reference"``).

    python oracle/make_ref.py            # needs /root/reference (authoring container only)

What it does: copies ``/root/reference/kfac`` (1 380 lines of pure Python) byte for byte to
``oracle/_ref/kfac`` and writes a 6-line ``horovod.torch`` stand-in next to it (the reference imports Horovod
unconditionally, kfac/backend.py:2, and this image has no Horovod; the Torch backend only ever calls
``hvd.synchronize`` on it).  ``oracle/_ref/`` is git-ignored (no reference source enters the history) but
travels to the GPU box with the snapshot, like the built ``.so`` files.  Test infrastructure / baseline only:
nothing under ``kfac_pytorch_b200/`` imports it.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('KFAC_REF', '/root/reference')
DST = os.path.join(HERE, '_ref')

HVD_STUB = '''"""Stand-in for horovod.torch (absent from this image): the reference's Torch backend
(kfac/backend.py) only needs the names below to import and run on torch.distributed."""
def synchronize(handle):
    return None
def size():
    return 1
def rank():
    return 0
'''


def main():
    src = os.path.join(REF, 'kfac')
    if not os.path.isdir(src):
        print('make_ref: %s not present (GPU box / CI): keeping whatever oracle/_ref already holds' % src)
        return 0
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    shutil.copytree(src, os.path.join(DST, 'kfac'), ignore=shutil.ignore_patterns('__pycache__'))
    os.makedirs(os.path.join(DST, 'horovod'))
    with open(os.path.join(DST, 'horovod', '__init__.py'), 'w') as fh:
        fh.write('from . import torch\n')
    with open(os.path.join(DST, 'horovod', 'torch.py'), 'w') as fh:
        fh.write(HVD_STUB)
    with open(os.path.join(DST, 'PROVENANCE.txt'), 'w') as fh:
        fh.write('kfac/: verbatim copy of %s/kfac made by oracle/make_ref.py\nhorovod/: stand-in written by the same script\n' % REF)
    n = sum(len(files) for _, _, files in os.walk(os.path.join(DST, 'kfac')))
    print('make_ref: copied %d files to %s' % (n, DST))
    return 0


if __name__ == '__main__':
    sys.exit(main())
