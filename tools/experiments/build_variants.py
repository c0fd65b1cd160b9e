#!/usr/bin/env python
"""Build one variant of libkfacb.so per experimental patch (tools/experiments/*.patch) WITHOUT touching the tree:
the sources are copied to a scratch directory, patched there and compiled with the library's own nvcc flags into
tools/_bin/libkfacb_<patch name>.so (git-ignored, but it travels with `gpurun`).  Evaluate them on the GPU box with
tools/experiments/eval_variants.sh (tools/eigh_probe.py loads a variant through KFACB_PROBE_LIB).

    python tools/experiments/build_variants.py [patch name ...]        # default: every *.patch
"""
import glob
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from kfac_pytorch_b200 import build as kbuild  # noqa: E402


def main():
    names = sys.argv[1:] or sorted(os.path.basename(p)[:-6] for p in glob.glob(os.path.join(HERE, '*.patch')))
    out_dir = os.path.join(ROOT, 'tools', '_bin')
    os.makedirs(out_dir, exist_ok=True)
    nvcc = kbuild._nvcc()
    for name in names:
        patch = os.path.join(HERE, name + '.patch')
        with tempfile.TemporaryDirectory() as tmp:
            shutil.copytree(os.path.join(ROOT, 'kfac_pytorch_b200', 'csrc'), os.path.join(tmp, 'kfac_pytorch_b200', 'csrc'))
            shutil.copytree(os.path.join(ROOT, 'include'), os.path.join(tmp, 'include'))
            r = subprocess.run(['patch', '-p1', '-i', patch], cwd=tmp, capture_output=True, text=True)
            if r.returncode != 0:
                print('%s: patch does not apply\n%s%s' % (name, r.stdout, r.stderr))
                continue
            csrc = os.path.join(tmp, 'kfac_pytorch_b200', 'csrc')
            flags = [f.replace(os.path.join(ROOT, 'include'), os.path.join(tmp, 'include')) for f in kbuild.NVCC_FLAGS]
            objs = []
            for src in sorted(f for f in os.listdir(csrc) if f.endswith('.cu')):
                obj = os.path.join(tmp, src[:-3] + '.o')
                cmd = [nvcc] + flags + ['-DKFACB_SRC_FINGERPRINT="variant:%s"' % name, '-c', os.path.join(csrc, src), '-o', obj]
                r = subprocess.run(cmd, capture_output=True, text=True)
                if r.returncode != 0:
                    print('%s: nvcc failed for %s\n%s' % (name, src, r.stderr))
                    break
                objs.append(obj)
            else:
                lib = os.path.join(out_dir, 'libkfacb_%s.so' % name)
                r = subprocess.run([nvcc, '-shared', '-o', lib] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a'],
                                   capture_output=True, text=True)
                print('%s: %s' % (name, 'built ' + lib if r.returncode == 0 else 'link failed\n' + r.stderr))


if __name__ == '__main__':
    main()
