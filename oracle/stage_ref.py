"""ORACLE (test infrastructure only): stages the UNMODIFIED reference Python tree for the GPU box.

    python -m oracle.stage_ref            # needs /root/reference (build container)

/root/reference does not exist on the GPU box, so what the GPU-side checks need of it — `src/` (torch_utils, dnnlib, training: the
networks, layers, motion, loss, augment modules) and `configs/` — is copied byte for byte into oracle/_ref/pyref/, which is git-ignored
(no reference source enters the history) but travels with the snapshot, exactly like the reference's CUDA plugins built by
oracle/build_ref.py.  `__graft_entry__.build()` runs this where /root/reference exists.  Consumers (oracle/ref_loader.py):
  * `bench.py --impl reference`: the reference's own SynthesisNetwork on the host cores, its `impl='ref'` ops (custom CUDA disabled);
  * tests/test_zz_reference_on_dropin_gpu.py: the unmodified reference Generator / Discriminator running on the drop-in ops on the GPU.
A manifest with a SHA-256 per file is written next to the copy; the loader refuses a tree whose files do not match it.
"""
import hashlib
import json
import os
import shutil

_HERE = os.path.dirname(os.path.abspath(__file__))
STAGED_ROOT = os.path.join(_HERE, '_ref', 'pyref')
SOURCE_ROOT = '/root/reference'
SUBTREES = ('src', 'configs')
MANIFEST = 'MANIFEST.json'


def _files(root):
    for sub in SUBTREES:
        for d, _, fs in os.walk(os.path.join(root, sub)):
            if '__pycache__' in d:
                continue
            for f in fs:
                if f.endswith(('.pyc', '.so', '.o')):
                    continue
                p = os.path.join(d, f)
                yield os.path.relpath(p, root), p


def _sha(path):
    h = hashlib.sha256()
    with open(path, 'rb') as f:
        h.update(f.read())
    return h.hexdigest()


def stage(force=False):
    """Copies the reference subtrees (skips if the manifest already matches the source).  Returns the staged root, or None without a source tree."""
    if not os.path.isdir(os.path.join(SOURCE_ROOT, 'src', 'torch_utils', 'ops')):
        return None
    want = {rel: _sha(p) for rel, p in _files(SOURCE_ROOT)}
    mpath = os.path.join(STAGED_ROOT, MANIFEST)
    if not force and os.path.exists(mpath) and json.load(open(mpath)) == want and verify():
        return STAGED_ROOT
    shutil.rmtree(STAGED_ROOT, ignore_errors=True)
    for rel, p in _files(SOURCE_ROOT):
        dst = os.path.join(STAGED_ROOT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(p, dst)
    json.dump(want, open(mpath, 'w'), indent=0, sort_keys=True)
    assert verify()
    return STAGED_ROOT


def verify():
    """True when every staged file has the hash recorded at staging time (i.e. the tree is the unmodified reference)."""
    mpath = os.path.join(STAGED_ROOT, MANIFEST)
    if not os.path.exists(mpath):
        return False
    want = json.load(open(mpath))
    return all(os.path.exists(os.path.join(STAGED_ROOT, rel)) and _sha(os.path.join(STAGED_ROOT, rel)) == h for rel, h in want.items())


if __name__ == '__main__':
    print(stage(force=True))
