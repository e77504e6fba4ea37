"""Launcher that runs an UNMODIFIED script of the reference checkout (train.py, inference.py) on the B200 plugins:

    cd /path/to/SwapNet
    python -m swapnet_b200.run train.py --name warp_stage --model warp --dataroot data/deep_fashion
    torchrun --nproc-per-node 8 -m swapnet_b200.run train.py --model warp --batch_size 16 ...      # data parallel

`python train.py` puts the script's directory first on sys.path, so the reference's own `models` package would win
over any PYTHONPATH entry.  This launcher reproduces Python's script start-up (sys.argv, sys.path[0] = the script's
directory, `__main__` namespace) with ONE difference: `<repo>/dropin` — whose `models` package re-exports
swapnet_b200.models under the top-level name the reference imports (models/__init__.py:5-44, train.py:25,38) — is placed
AHEAD of the script directory.  `dropin/datasets` is an overlay of the reference's `datasets` package: it executes the
reference's own `datasets/__init__.py` and keeps every reference submodule, and only adds the dataset plugins
`--dataset warp_b200` / `texture_b200` (device-side input pipeline).  Everything else (options/, util/, optimizers/,
modules/) resolves to the reference checkout.  Under torchrun the process group is created by the plugin itself (BaseModel.__init__), so
train.py needs no distributed code.
"""
from __future__ import annotations

import os
import runpy
import sys


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ("-h", "--help"):
        print(__doc__)
        raise SystemExit(0 if argv else 2)
    script = os.path.abspath(argv[0])
    if not os.path.isfile(script):
        raise SystemExit(f"swapnet_b200.run: no such script: {argv[0]}")
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dropin = os.path.join(repo, "dropin")
    # Python's own layout for `python script.py` is [script_dir, PYTHONPATH..., site]; ours: [dropin, script_dir, repo, ...]
    for p in (repo, os.path.dirname(script), dropin):
        while p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    stale = sys.modules.get("models")
    if stale is not None and not os.path.abspath(getattr(stale, "__file__", "")).startswith(dropin):
        raise SystemExit("swapnet_b200.run: a foreign `models` package is already imported: " + repr(stale))
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
