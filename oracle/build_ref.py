"""ORACLE (test infrastructure only): builds the REFERENCE'S OWN CUDA plugins for sm_100a into oracle/_ref/.

    python -m oracle.build_ref            # needs /root/reference (build container); nvcc cross-compiles without a GPU

The sources are compiled where they lie — /root/reference/src/torch_utils/ops/{upfirdn2d,bias_act}.{cpp,cu} (+ .h), unmodified,
with the reference's own flag (`--use_fast_math`, upfirdn2d.py:29 / bias_act.py:45) — through torch.utils.cpp_extension (ninja + nvcc +
the torch / pybind11 headers the plugins include).  Nothing is copied into the repository; the outputs (`oracle/_ref/<name>/<name>.so`)
are git-ignored but travel to the GPU box with the snapshot, where /root/reference does not exist.  They are the north_star's oracle
"the reference's own JIT-compiled ops": tests/test_zz_reference_cuda_gpu.py compares libsgv_b200 against them on identical inputs.

(The reference's loader cannot produce these modules itself on current torch: custom_ops.get_plugin ends with
importlib.import_module(module_name), which no longer finds a cpp_extension.load()-ed module — SURVEY.md §8c — so the ops would
silently take their slow fallback.  `load_plugin` below imports the built file directly instead.)
"""
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(_HERE, '_ref')
PLUGINS = {'upfirdn2d_plugin': ['upfirdn2d.cpp', 'upfirdn2d.cu'], 'bias_act_plugin': ['bias_act.cpp', 'bias_act.cu']}


def _ops_dir():
    from . import ref_loader
    return os.path.join(ref_loader.REF_ROOT, 'src', 'torch_utils', 'ops')


def plugin_path(name):
    return os.path.join(OUT_DIR, name, name + '.so')


def build(force=False, verbose=False):
    """Compiles both plugins (skips those already built unless force).  Returns {name: path}.  No-op ({}) without the reference tree."""
    from . import ref_loader
    if not ref_loader.available():
        return {}
    import torch.utils.cpp_extension as ext
    os.environ.setdefault('TORCH_CUDA_ARCH_LIST', '10.0')          # no GPU here to detect; the explicit -gencode below is what is used
    out = {}
    for name, files in PLUGINS.items():
        path = plugin_path(name)
        srcs = [os.path.join(_ops_dir(), f) for f in files]
        if force or not os.path.exists(path) or os.path.getmtime(path) < max(os.path.getmtime(s) for s in srcs):
            bdir = os.path.join(OUT_DIR, name)
            os.makedirs(bdir, exist_ok=True)
            ext.load(name=name, sources=srcs, build_directory=bdir, verbose=verbose, with_cuda=True, is_python_module=False,
                     extra_cuda_cflags=['--use_fast_math', '-gencode', 'arch=compute_100a,code=sm_100a'])
        assert os.path.exists(path), path
        for junk in os.listdir(os.path.join(OUT_DIR, name)):          # keep only the plugin: the snapshot that travels to the GPU box stays small
            if junk.endswith(('.o', '.d')):
                os.remove(os.path.join(OUT_DIR, name, junk))
        out[name] = path
    return out


def load_plugin(name):
    """Imports oracle/_ref/<name>/<name>.so (built earlier; does not need /root/reference) or returns None."""
    path = plugin_path(name)
    if not os.path.exists(path):
        return None
    if name in sys.modules:
        return sys.modules[name]
    import torch  # noqa: F401  (libtorch symbols must be loaded first)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[name] = mod
    return mod


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
