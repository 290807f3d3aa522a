"""Makes the reference code base run on libsgv_b200 without editing it.

    import stylegan_v_b200.install as sgv; sgv.install_ops()
    from training.networks import Generator      # unchanged reference; its `torch_utils.ops` imports now resolve here

`install_ops()` registers stylegan_v_b200.ops.* under the module names the reference imports
(`src.torch_utils.ops.<name>` — networks.py:16, layers.py:11, loss.py:15, augment.py:14-16, training_loop.py:25-26 —
and `torch_utils.ops.<name>` for scripts that put src/ on sys.path).  It must run before the reference modules are
imported.  See INTEGRATION.md.
"""
import importlib
import sys
import types

OP_MODULES = ('upfirdn2d', 'bias_act', 'conv2d_resample', 'conv2d_gradfix', 'fma', 'grid_sample_gradfix')


def install_ops(prefixes=('src.torch_utils.ops', 'torch_utils.ops')):
    mine = importlib.import_module('stylegan_v_b200.ops')
    for prefix in prefixes:
        parent_name = prefix.rsplit('.', 1)[0]
        pkg = sys.modules.get(prefix)
        if pkg is None or getattr(pkg, '__sgv__', False) is False:
            pkg = types.ModuleType(prefix)
            pkg.__path__ = []
            pkg.__sgv__ = True
            sys.modules[prefix] = pkg
        for name in OP_MODULES:
            mod = getattr(mine, name)
            sys.modules[f'{prefix}.{name}'] = mod
            setattr(pkg, name, mod)
        parent = sys.modules.get(parent_name)
        if parent is not None:
            setattr(parent, 'ops', pkg)
    return mine
