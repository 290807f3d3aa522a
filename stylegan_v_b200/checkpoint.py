"""Loads the reference's `network-snapshot-*.pkl` files into the native modules — WITHOUT the reference source tree.

The reference pickles networks through `torch_utils.persistence` (src/torch_utils/persistence.py:118-126,179-203): every persistent module
is stored as `_reconstruct_persistent_obj(meta)` with meta = {type, version, module_src, class_name, state}, where `state` is the
module's `__dict__` (`_parameters`, `_buffers`, `_modules`, `_init_args`, `_init_kwargs`, plain attributes) and `module_src` the text of
the defining Python file, which the reference exec()s on load.  Here nothing is executed: a restricted unpickler maps that constructor
(under any of the module paths the reference was run with) to a passive `PersistedModule` record, `dnnlib.EasyDict` to a dict, and refuses
every other global that is not a torch / numpy / collections data type.  From the record tree
    state_dict(rec)                 rebuilds the flat parameter / buffer dict (the keys equal the native modules' — state-dict compatible)
    build_generator(rec) / build_discriminator(rec)
                                    read the constructor arguments the reference recorded (`_init_kwargs`: cfg node, channel_base, ...),
                                    build stylegan_v_b200.networks.Generator / Discriminator and load the weights
    load_snapshot(path)             -> {'G': Generator, 'G_ema': Generator, 'D': Discriminator, ...} like legacy.load_network_pkl (legacy.py:20-60).
The other direction needs no code: `native.state_dict()` loads into the reference's modules with `load_state_dict` (same keys and shapes).
"""
import collections
import io
import pickle

import torch

from .networks import Discriminator, Generator

_PERSISTENCE_MODULES = ('src.torch_utils.persistence', 'torch_utils.persistence')
_EASYDICT_MODULES = ('src.dnnlib.util', 'dnnlib.util', 'src.dnnlib', 'dnnlib')
_SAFE_PREFIXES = ('torch', 'numpy', 'collections', '_codecs', 'builtins')
_SAFE_BUILTINS = {'set', 'frozenset', 'list', 'dict', 'tuple', 'slice', 'complex', 'bytearray', 'range', 'getattr', 'int', 'float', 'bool', 'str', 'bytes'}


class EasyDict(dict):
    """Attribute-access dict (the reference's dnnlib.EasyDict)."""
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)
    __setattr__ = dict.__setitem__


class PersistedModule:
    """Passive record of one pickled persistent object: class name + its `__dict__`; the pickled source text is kept but never executed."""

    def __init__(self, meta):
        self.class_name = meta['class_name']
        self.version = meta.get('version')
        self.module_src = meta.get('module_src')
        self.state = dict(meta['state'])

    @property
    def init_kwargs(self):
        return self.state.get('_init_kwargs', {})

    @property
    def init_args(self):
        return self.state.get('_init_args', ())

    def children(self):
        return self.state.get('_modules', {}) or {}

    def __repr__(self):
        return f'PersistedModule({self.class_name}, {len(self.children())} children)'


def _reconstruct(meta):
    return PersistedModule(meta)


class _Opaque:
    """Passive stand-in for a pickled object whose class is not imported (omegaconf's DictConfig / ListConfig / value nodes, which real
    snapshots hold in `cfg`): keeps constructor arguments and state; `plain()` turns config containers into dicts / lists / values."""
    _opaque_name = '?'

    def __init__(self, *args, **kwargs):
        self._args = args

    def __setstate__(self, state):
        self.__dict__['_state'] = state

    def __repr__(self):
        return f'<opaque {self._opaque_name}>'


_opaque_classes = {}


def _opaque_class(module, name):
    key = f'{module}.{name}'
    if key not in _opaque_classes:
        _opaque_classes[key] = type(name, (_Opaque,), {'_opaque_name': key})
    return _opaque_classes[key]


def plain(o, _depth=0):
    """Config containers (EasyDict, omegaconf records read opaquely) -> EasyDict / list / python values."""
    assert _depth < 64
    if isinstance(o, _Opaque):
        st = o.__dict__.get('_state')
        st = st if isinstance(st, dict) else o.__dict__
        if '_content' in st:
            return plain(st['_content'], _depth + 1)
        if '_val' in st:
            return plain(st['_val'], _depth + 1)
        return o._args[0] if getattr(o, '_args', ()) else None           # enum-like value objects
    if isinstance(o, dict):
        return EasyDict({k: plain(v, _depth + 1) for k, v in o.items()})
    if isinstance(o, (list, tuple)):
        return [plain(v, _depth + 1) for v in o]
    return o


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module in _PERSISTENCE_MODULES and name == '_reconstruct_persistent_obj':
            return _reconstruct
        if module in _EASYDICT_MODULES and name == 'EasyDict':
            return EasyDict
        if module == 'builtins':
            if name in _SAFE_BUILTINS:
                return super().find_class(module, name)
            raise pickle.UnpicklingError(f'refusing builtins.{name}')
        if module.split('.')[0] in _SAFE_PREFIXES or module in ('typing', 'enum'):
            return super().find_class(module, name)
        if module.split('.')[0] == 'omegaconf' or module.startswith(('training.', 'src.training.')):
            # config containers, and the reference's non-persistent sub-modules (e.g. layers.TemporalDifferenceEncoder): read as passive records
            return _opaque_class(module, name)
        raise pickle.UnpicklingError(f'refusing to import {module}.{name} while reading a network pickle')


def load_records(f):
    """File object or bytes -> the unpickled container (usually a dict with 'G', 'D', 'G_ema' PersistedModule records)."""
    if isinstance(f, (bytes, bytearray)):
        f = io.BytesIO(f)
    return _Unpickler(f).load()


def state_dict(rec, prefix=''):
    """Flat {name: tensor} of a record tree in torch's state_dict order (parameters, persistent buffers, then children)."""
    out = collections.OrderedDict()
    if isinstance(rec, PersistedModule):
        st = rec.state
    else:
        st = rec.__dict__.get('_state') if isinstance(rec.__dict__.get('_state'), dict) else rec.__dict__
    non_persistent = st.get('_non_persistent_buffers_set', set())
    for k, v in (st.get('_parameters') or {}).items():
        if v is not None:
            out[prefix + k] = v.data if isinstance(v, torch.nn.Parameter) else v
    for k, v in (st.get('_buffers') or {}).items():
        if v is not None and k not in non_persistent:
            out[prefix + k] = v
    for k, child in (st.get('_modules') or {}).items():
        if child is not None:
            out.update(state_dict(child, prefix + k + '.'))
    return out


def _get(o, k, d=None):
    if o is None:
        return d
    if isinstance(o, dict):
        return o.get(k, d)
    return getattr(o, k, d)


def build_generator(rec, **overrides):
    """Native Generator from a pickled reference Generator record (constructor: networks.py:371-381; kwargs as train.py:163-175 sets them)."""
    assert rec.class_name == 'Generator', rec.class_name
    kw = rec.init_kwargs
    syn = _get(kw, 'synthesis_kwargs', {}) or {}
    mp = _get(kw, 'mapping_kwargs', {}) or {}
    assert _get(kw, 'c_dim', 0) in (0, None), 'conditional generators (c_dim > 0) are outside the StyleGAN-V model this package covers'
    G = Generator.from_reference_cfg(plain(_get(kw, 'cfg')), img_resolution=_get(kw, 'img_resolution'), img_channels=_get(kw, 'img_channels', 3),
                                     channel_base=_get(syn, 'channel_base', 32768), channel_max=_get(syn, 'channel_max', 512),
                                     mapping_layers=_get(mp, 'num_layers', 8), num_fp16_res=_get(syn, 'num_fp16_res', 0),
                                     conv_clamp=_get(syn, 'conv_clamp', None), **overrides)
    G.load_state_dict(state_dict(rec), strict=True)
    return G


def build_discriminator(rec, **overrides):
    """Native Discriminator from a pickled reference Discriminator record (constructor: networks.py:580-594)."""
    assert rec.class_name == 'Discriminator', rec.class_name
    kw = rec.init_kwargs
    epi = _get(kw, 'epilogue_kwargs', {}) or {}
    mp = _get(kw, 'mapping_kwargs', {}) or {}
    blk = _get(kw, 'block_kwargs', {}) or {}
    D = Discriminator.from_reference_cfg(plain(_get(kw, 'cfg')), img_resolution=_get(kw, 'img_resolution'), img_channels=_get(kw, 'img_channels', 3),
                                         c_dim=_get(kw, 'c_dim', 0) or 0, channel_base=_get(kw, 'channel_base', 32768),
                                         channel_max=_get(kw, 'channel_max', 512), mbstd_group_size=_get(epi, 'mbstd_group_size', 4),
                                         mapping_layers=_get(mp, 'num_layers', 8), architecture=_get(kw, 'architecture', 'resnet'),
                                         num_fp16_res=_get(kw, 'num_fp16_res', 0), conv_clamp=_get(kw, 'conv_clamp', None),
                                         freeze_layers=_get(blk, 'freeze_layers', 0), **overrides)
    D.load_state_dict(state_dict(rec), strict=True)
    return D


def load_snapshot(path_or_file, **overrides):
    """A reference `network-snapshot-*.pkl` (dict with G / D / G_ema, training_loop.py:447-456) -> the same dict with native modules in
    place of the pickled networks; other entries (training_set_kwargs, augment_pipe record, ...) are passed through as read."""
    if isinstance(path_or_file, str):
        with open(path_or_file, 'rb') as f:
            data = load_records(f)
    else:
        data = load_records(path_or_file)
    out = dict(data) if isinstance(data, dict) else {'G_ema': data}
    for k, v in list(out.items()):
        if isinstance(v, PersistedModule):
            if v.class_name == 'Generator':
                out[k] = build_generator(v, **overrides)
            elif v.class_name == 'Discriminator':
                out[k] = build_discriminator(v, **overrides)
    return out
