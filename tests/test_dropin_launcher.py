"""The drop-in boundary, end to end on the CPU: the UNMODIFIED reference `train.py` started through
`python -m swapnet_b200.run` resolves `models` to this repo's plugins, parses its options through them
(options/base_options.py:171-186), builds the reference's own dataset from files on disk and reaches the plugin's
constructor — which, on a box without a GPU, refuses to run (there is no CPU fallback).  On a GPU box the same command
trains (tests/test_engine_gpu.py::test_train_loop_protocol restates train.py:31-116 there, where /root/reference does
not exist).  Needs /root/reference (build container); skipped elsewhere."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
STUBS = os.path.join(ROOT, "tests", "tools", "ref_stubs")

needs_ref = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "train.py")), reason="/root/reference not mounted")


def make_dataset(root, n=2, size=64):
    """A tiny dataset in the reference's on-disk format: cloth/*.npz (scipy CSC label maps, data_utils.py:311-327),
    body/*.jpg, texture/*.jpg, normalization_stats.json (json lines indexed by `path`, data_utils.py:30-38)."""
    from PIL import Image
    from scipy import sparse

    rng = np.random.RandomState(0)
    for d in ("cloth", "body", "texture"):
        os.makedirs(os.path.join(root, d), exist_ok=True)
    for i in range(n):
        lab = np.kron(rng.randint(0, 19, (size // 16, size // 16)), np.ones((16, 16), dtype=np.int64))
        sparse.save_npz(os.path.join(root, "cloth", f"{i}.npz"), sparse.csc_matrix(lab))
        for d in ("body", "texture"):
            Image.fromarray(rng.randint(0, 255, (size, size, 3), dtype=np.uint8)).save(os.path.join(root, d, f"{i}.jpg"))
    with open(os.path.join(root, "normalization_stats.json"), "w") as f:
        for k in ("body", "texture", "cloth"):
            f.write(json.dumps({"path": k, "means": [0.5, 0.5, 0.5], "stds": [0.25, 0.25, 0.25]}) + "\n")


def run(cmd, cwd, extra_path=()):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([*extra_path, ROOT, STUBS])
    env["CUDA_VISIBLE_DEVICES"] = ""
    return subprocess.run(cmd, cwd=cwd, env=env, capture_output=True, text=True, timeout=600)


@needs_ref
def test_plain_python_train_py_resolves_the_reference_models_package(tmp_path):
    """Why a launcher is needed: `python script.py` puts the script's directory first, so PYTHONPATH cannot shadow
    the reference's `models` package."""
    probe = "import models, sys; print(models.__file__)"
    r = run([sys.executable, "-c", f"import sys; sys.path.insert(0, {REF!r}); {probe}"], cwd=str(tmp_path),
            extra_path=[os.path.join(ROOT, "dropin")])
    assert r.returncode == 0 and r.stdout.strip().startswith(REF), (r.stdout, r.stderr[-2000:])


@needs_ref
def test_launcher_runs_unmodified_train_py_through_the_plugin(tmp_path):
    data = tmp_path / "data"
    make_dataset(str(data))
    ck = tmp_path / "ck"
    r = run([sys.executable, "-m", "swapnet_b200.run", "train.py", "--name", "t", "--model", "warp", "--dataroot",
             str(data), "--checkpoints_dir", str(ck), "--display_id", "0", "--batch_size", "1", "--load_size", "64",
             "--crop_size", "64", "--num_workers", "0", "--no_confirm", "--gpu_id", "0", "--n_epochs", "1"],
            cwd=REF)
    out = r.stdout + r.stderr
    # options parsed through the plugin: our extra flag is in the printed / stored option table (train_options.py)
    assert "b200_precision" in out, out[-3000:]
    assert "The number of training images = 2" in out, out[-3000:]
    # ... and the model constructor that refused is OURS (no GPU here, no CPU fallback)
    assert r.returncode != 0
    assert "swapnet_b200 models run on a CUDA device only" in out, out[-3000:]
    assert os.path.join("swapnet_b200", "models", "base_model.py") in out


@needs_ref
def test_launcher_option_defaults_match_the_reference_parser(tmp_path):
    """TrainOptions().parse() through the plugin yields the reference's defaults for every flag that reaches the hot
    path, for both models."""
    probe = tmp_path / "probe.py"
    probe.write_text(
        "import sys, json\n"
        "import models\n"
        "from options.train_options import TrainOptions\n"
        "opt = TrainOptions().parse(print_options=False) if 'print_options' in TrainOptions.parse.__code__.co_varnames"
        " else TrainOptions().parse()\n"
        "keys = ['lr','d_lr','weight_decay','d_weight_decay','b1','b2','gan_mode','gan_label_mode','norm','init_type',"
        "'lambda_gan','lambda_discriminator','discriminator','optimizer_G','optimizer_D','b200_precision']\n"
        "extra = {'warp': ['warp_mode','lambda_ce'], 'texture': ['lambda_l1','lambda_content','lambda_style','netG']}\n"
        "print('PROBE', json.dumps({'file': models.__file__, **{k: getattr(opt, k) for k in keys + extra[opt.model]}}))\n")
    want = dict(lr=1e-4, d_lr=4e-4, weight_decay=0, d_weight_decay=0.01, b1=0.9, b2=0.999, gan_mode="vanilla",
                gan_label_mode="smooth", norm="instance", init_type="kaiming", lambda_gan=1.0, lambda_discriminator=1.0,
                discriminator="basic", optimizer_G="AdamW", optimizer_D="AdamW", b200_precision="fp32x3")
    extra = dict(warp=dict(warp_mode="gan", lambda_ce=100), texture=dict(lambda_l1=10, lambda_content=20,
                                                                         lambda_style=1e-8, netG="swapnet"))
    data = tmp_path / "data"
    make_dataset(str(data))
    # the probe must live in the checkout directory for Python's script start-up rule to bite: emulate by cwd + path
    for model in ("warp", "texture"):
        r = run([sys.executable, "-m", "swapnet_b200.run", str(probe), "--name", "p", "--model", model, "--dataroot",
                 str(data), "--checkpoints_dir", str(tmp_path / "ck"), "--no_confirm"], cwd=REF, extra_path=[REF])
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("PROBE ")]
        assert r.returncode == 0 and line, (r.stdout[-2000:], r.stderr[-3000:])
        got = json.loads(line[-1][6:])
        assert got.pop("file").startswith(os.path.join(ROOT, "dropin")), got
        for k, v in {**want, **extra[model]}.items():
            assert got[k] == v, (model, k, got[k], v)
