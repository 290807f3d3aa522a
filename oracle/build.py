"""ORACLE (test infrastructure): builds oracle/_build/liboracle.so from oracle/c/sgv_oracle.c with gcc.

Called by __graft_entry__.build() and lazily by oracle.ops_ref.  `-ffp-contract=off` keeps gcc
from fusing anything we did not write as fmaf()/fma() explicitly, so float results do not depend
on the host's ISA.
"""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, 'c', 'sgv_oracle.c')
OUT_DIR = os.path.join(_HERE, '_build')
OUT = os.path.join(OUT_DIR, 'liboracle.so')


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    cmd = ['gcc', '-O2', '-std=c11', '-fPIC', '-shared', '-ffp-contract=off', '-fno-fast-math',
           SRC, '-o', OUT + '.tmp', '-lm']
    subprocess.run(cmd, check=True)
    os.replace(OUT + '.tmp', OUT)
    return OUT


if __name__ == '__main__':
    print(build(force=True))
