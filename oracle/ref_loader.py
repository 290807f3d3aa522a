"""ORACLE (test infrastructure only).  Imports the UNMODIFIED Python reference from /root/reference.

Imports /root/reference in place in the build container; on the GPU box (no /root/reference) the byte-for-byte copy that
oracle/stage_ref.py stages under the git-ignored oracle/_ref/pyref (hash-verified on load).  Used by oracle/make_goldens.py to mint
tests/golden/*.npz, by `bench.py --impl reference`, and by tests that are skipped when no reference tree is present.

The reference needs `omegaconf` (absent here) only for `OmegaConf.create/to_container` and the
`DictConfig` name (src/training/networks.py:12,384; layers.py:8; motion.py:6) — a minimal in-memory
stand-in is injected into sys.modules.
"""
import os
import sys
import types

def _find_root():
    """The reference tree: $SGV_REFERENCE_ROOT, else /root/reference (build container), else the byte-for-byte copy staged by
    oracle/stage_ref.py under oracle/_ref/pyref (what the GPU box has)."""
    env = os.environ.get('SGV_REFERENCE_ROOT')
    if env:
        return env
    if os.path.isdir('/root/reference/src/torch_utils/ops'):
        return '/root/reference'
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref', 'pyref')


REF_ROOT = _find_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, 'src', 'torch_utils', 'ops'))


def is_staged_copy() -> bool:
    return os.path.basename(os.path.normpath(REF_ROOT)) == 'pyref'


class _Cfg(dict):
    """Attribute-access dict (what the reference nets need from a DictConfig)."""
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)
    __setattr__ = dict.__setitem__


def to_cfg(d):
    if isinstance(d, dict):
        return _Cfg({k: to_cfg(v) for k, v in d.items()})
    return d


def _install_omegaconf_stub():
    if 'omegaconf' in sys.modules:
        return
    m = types.ModuleType('omegaconf')

    class OmegaConf:
        @staticmethod
        def create(x):
            return to_cfg(dict(x))

        @staticmethod
        def to_container(x, **_):
            def plain(v):
                return {k: plain(u) for k, u in v.items()} if isinstance(v, dict) else v
            return plain(x)
    m.OmegaConf = OmegaConf
    m.DictConfig = _Cfg
    sys.modules['omegaconf'] = m


_loaded = None


def load():
    """Returns a namespace with the reference modules (ops + networks)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f'reference tree not found at {REF_ROOT}')
    if is_staged_copy():
        from . import stage_ref
        assert stage_ref.verify(), 'oracle/_ref/pyref does not match its manifest: not the unmodified reference'
    _install_omegaconf_stub()
    for p in (os.path.join(REF_ROOT, 'src'), REF_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import importlib
    ns = types.SimpleNamespace()
    ns.upfirdn2d = importlib.import_module('src.torch_utils.ops.upfirdn2d')
    ns.bias_act = importlib.import_module('src.torch_utils.ops.bias_act')
    ns.conv2d_resample = importlib.import_module('src.torch_utils.ops.conv2d_resample')
    ns.conv2d_gradfix = importlib.import_module('src.torch_utils.ops.conv2d_gradfix')
    ns.fma = importlib.import_module('src.torch_utils.ops.fma')
    ns.networks = importlib.import_module('training.networks')
    ns.layers = importlib.import_module('training.layers')
    ns.motion = importlib.import_module('training.motion')
    ns.loss = importlib.import_module('training.loss')
    ns.augment = importlib.import_module('training.augment')
    _loaded = ns
    return ns
