"""Overlay of the reference's `datasets` package: everything the reference defines stays the reference's, and the
dataset registry (`--dataset <name>` -> module `datasets.<name>_dataset`, datasets/__init__.py:9-31) additionally finds
the plugins that live next to this file (warp_b200_dataset.py: the warp dataset with the per-channel augmentation moved
to the device).

How: `python -m swapnet_b200.run train.py` puts `<repo>/dropin` ahead of the checkout, so `import datasets` lands here.
This module then (1) appends the checkout's own `datasets/` directory to `__path__` — submodules (`datasets.data_utils`,
`datasets.warp_dataset`, ...) keep resolving to the reference's files — and (2) executes the reference's own
`datasets/__init__.py` in this namespace, so `create_dataset`, `CappedDataLoader`, `get_transforms`, `BaseDataset`, ... are
the reference's code, unmodified and uncopied.
"""
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
_ref_init = None
for _p in _sys.path:
    _cand = _os.path.join(_os.path.abspath(_p or _os.getcwd()), "datasets", "__init__.py")
    # the reference's package, not e.g. the HuggingFace `datasets` library in site-packages: it has base_dataset.py
    if (_os.path.isfile(_cand) and _os.path.dirname(_cand) != _here
            and _os.path.isfile(_os.path.join(_os.path.dirname(_cand), "base_dataset.py"))):
        _ref_init = _cand
        break
if _ref_init is None:
    raise ImportError("dropin/datasets overlays the reference's `datasets` package, which is not on sys.path "
                      "(start the reference's scripts with `python -m swapnet_b200.run <script>` from its checkout)")
__path__.append(_os.path.dirname(_ref_init))
REFERENCE_INIT = _ref_init
with open(_ref_init) as _f:
    exec(compile(_f.read(), _ref_init, "exec"), globals())
