"""Builds stylegan_v_b200/libsgv_b200.so (in-tree, sm_100a only) with nvcc.

    python -m stylegan_v_b200.build [--force]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the snapshot.
"""
import os
import subprocess
import sys
import hashlib

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ_DIR = os.path.join(HERE, 'csrc', '_obj')
LIB = os.path.join(HERE, 'libsgv_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
CFLAGS = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', '-Xptxas', '-v']


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _deps_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), 'include')):
        for f in os.listdir(root):
            if f.endswith(('.cuh', '.h')):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def build(force=False, verbose=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_m = _deps_mtime()
    objs = []
    rebuilt = False
    for src in sources():
        sp = os.path.join(CSRC, src)
        op = os.path.join(OBJ_DIR, src[:-3] + '.o')
        objs.append(op)
        if force or not os.path.exists(op) or os.path.getmtime(op) < max(os.path.getmtime(sp), hdr_m):
            cmd = [NVCC] + ARCH + CFLAGS + ['-c', sp, '-o', op]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f'nvcc failed for {src}')
            with open(op + '.log', 'w') as f:
                f.write(r.stdout + r.stderr)
            rebuilt = True
    if rebuilt or force or not os.path.exists(LIB):
        cmd = [NVCC] + ARCH + ['-shared', '-o', LIB + '.tmp'] + objs + ['-lcudart']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError('link failed')
        os.replace(LIB + '.tmp', LIB)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
