"""The C-ABI library loads and exports every symbol include/*.h declares; ctypes structs match the C layout.
No compute calls (CPU box)."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import pytest

from conftest import ROOT
from stylegan_v_b200 import _lib

INCLUDE = os.path.join(ROOT, 'include')


def _declared_functions():
    names = []
    for fn in sorted(os.listdir(INCLUDE)):
        if not fn.endswith('.h'):
            continue
        src = open(os.path.join(INCLUDE, fn)).read()
        src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
        names += re.findall(r'\b(sgv_[a-z0-9_]+)\s*\(', src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), 'build with python -m stylegan_v_b200.build'
    L = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_functions()
    assert 'sgv_upfirdn2d' in declared and 'sgv_bias_act' in declared
    for name in declared:
        assert hasattr(L, name), f'{name} declared in include/ but not exported'
    bound = {s[0] for s in _lib.SYMBOLS}
    assert set(declared) <= bound, f'binding misses {set(declared) - bound}'


def test_loader_and_version():
    L = _lib.lib()
    assert L.sgv_abi_version() == _lib.ABI_VERSION
    assert L.sgv_upfirdn2d_out_size(257, 1, 1, 1, 4, 1) == 256
    assert L.sgv_upfirdn2d_out_size(128, 2, 2, 1, 4, 1) == 256


def test_struct_layouts_match_c():
    structs = dict(_lib.STRUCTS)
    prog = ['#include <stdio.h>', '#include <stddef.h>']
    for fn in sorted(os.listdir(INCLUDE)):
        if fn.endswith('.h'):
            prog.append(f'#include "{fn}"')
    prog.append('int main(void){')
    for cname, st in structs.items():
        prog.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in st._fields_:
            prog.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    prog.append('return 0;}')
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, 't.c')
        open(c, 'w').write('\n'.join(prog))
        exe = os.path.join(d, 't')
        subprocess.run(['gcc', '-I', INCLUDE, c, '-o', exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    got = dict(line.split() for line in out.strip().splitlines())
    for cname, st in structs.items():
        assert int(got[cname]) == ctypes.sizeof(st), cname
        for fname, _ in st._fields_:
            assert int(got[f'{cname}.{fname}']) == getattr(st, fname).offset, (cname, fname)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(_lib.SgvError):
        _lib.lib()


def test_no_cpu_path_in_plugin():
    import torch
    from stylegan_v_b200 import plugin
    x = torch.zeros(1, 1, 4, 4)
    with pytest.raises(RuntimeError):
        plugin.upfirdn2d(x, torch.ones(1, 1), 1, 1, 1, 1, 0, 0, 0, 0, False, 1.0)
    with pytest.raises(RuntimeError):
        e = torch.empty(0)
        plugin.bias_act(x, e, e, e, e, 0, 1, 1, 0.0, 1.0, -1.0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'stylegan_v_b200')
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f
                assert 'oracle/' not in src or f.endswith('.md'), f


def test_aux_entry_points_have_no_cpu_path():
    """Time-encoder tail, fused optimiser step and the fused layer nodes refuse CPU tensors / a missing device instead of computing elsewhere."""
    import torch
    from stylegan_v_b200.optim import FlatModuleState, FusedAdamEMA
    from stylegan_v_b200.time_encoder import _TimeEncoderTail
    from stylegan_v_b200 import dconv
    with pytest.raises(Exception):
        _TimeEncoderTail.apply(torch.zeros(2, 16), torch.zeros(2, 8), torch.zeros(2), torch.ones(4), torch.ones(4), 16.0)
    with pytest.raises(AssertionError):
        FusedAdamEMA(FlatModuleState([torch.nn.Parameter(torch.zeros(8))]))
    assert not dconv.supported(torch.zeros(1, 64, 8, 8), torch.zeros(64, 64, 3, 3), 1, 1)          # CPU tensors never take the fused node
    L = _lib.lib()
    if not torch.cuda.is_available():
        buf = (ctypes.c_float * 64)()
        rc = L.sgv_time_encoder_fwd(buf, buf, buf, buf, buf, buf, 1, 4, ctypes.c_float(16.0), None)
        assert rc == 4 and b'no CPU path' in L.sgv_last_error()                                    # SGV_ERR_NO_DEVICE
