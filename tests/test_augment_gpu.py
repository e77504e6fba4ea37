"""GPU parity of the device-side cloth augmentation (csrc/augment.cu through the C ABI).

Bit-exact against (a) the numpy oracle (oracle/augment.py), (b) the golden fixture generated from the reference's
`per_channel_transform`, (c) Pillow / torchvision run live on the same draws — at the benchmark's 512x512 too — and
the plugin consumes the result like the dense host tensor the reference's dataset would have produced.
"""
import random

import numpy as np
import pytest
import torch

from oracle import augment as A
from swapnet_b200 import data as D
from swapnet_b200 import ops
from test_augment_cpu import GOLDEN, label_map, pil_per_channel, reference_transform

pytestmark = pytest.mark.gpu
ALL = ("hflip", "vflip", "affine", "perspective")


def dev():
    return torch.device("cuda:0")


def run_device(labels_np, sample_ops, channels=19):
    aug = D.ClothAugmenter(None, channels)
    out = aug.apply(torch.from_numpy(labels_np).to(dev()), D.OpTable(sample_ops) if labels_np.shape[0] > 1 else sample_ops)
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("size", [(64, 64), (40, 56), (33, 130)])
@pytest.mark.parametrize("names", [ALL, ("affine",), ("hflip", "vflip"), ("perspective",)])
def test_device_equals_oracle(size, names):
    h, w = size
    tf = reference_transform(names)
    B = 3
    labs = np.stack([label_map(h, w, 20 + b) for b in range(B)])
    random.seed(1); torch.manual_seed(1)
    sample_ops = [D.draw_channel_ops(tf, 19, w, h) for _ in range(B)]
    got = run_device(labs, sample_ops)
    for b in range(B):
        ref = A.per_channel_transform(A.onehot(labs[b], 19), sample_ops[b])
        assert np.array_equal(got[b], ref), (b, int((got[b] != ref).sum()))


def test_device_dense_source_and_every_op_count():
    """dense fp32 source (real-valued planes: the float32 tap difference of the bilinear filter matters) and planes
    with 0, 1, 2, 3 and 4 ops in one launch sequence (ping-pong parity: the last pass must land in `out`)."""
    h, w, C = 48, 72, 5
    g = np.random.default_rng(3)
    dense = g.random((2, C, h, w)).astype(np.float32)
    torch.manual_seed(4)
    from torchvision import transforms as T
    from torchvision.transforms import functional as TF

    persp = lambda: (D.AUG_PERSPECTIVE_BILINEAR, tuple(TF._get_perspective_coeffs(*T.RandomPerspective.get_params(w, h, 0.5))))
    aff = lambda ang: D._affine_op(TF._get_inverse_affine_matrix([w * .5, h * .5], ang, [3, -2], 1.1, [7.0, 0.0]), w, h)
    menu = [[], [(D.AUG_HFLIP, ())], [aff(5.0), (D.AUG_VFLIP, ())], [persp(), aff(-8.0), (D.AUG_HFLIP, ())],
            [(D.AUG_VFLIP, ()), persp(), (D.AUG_HFLIP, ()), aff(2.5)]]
    sample_ops = [[menu[(b + c) % 5] for c in range(C)] for b in range(2)]
    table, max_ops = D.encode_ops([o for s in sample_ops for o in s])
    assert max_ops == 4
    src = torch.from_numpy(dense).to(dev())
    tab = torch.from_numpy(table.view(np.uint8).reshape(-1)).to(dev())
    out, tmp = torch.full_like(src, -7.0), torch.full_like(src, -9.0)
    ops.augment_channels(src, C, tab, table.shape[1], max_ops, out, tmp)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    for b in range(2):
        assert np.array_equal(got[b], A.per_channel_transform(dense[b], sample_ops[b])), b
    with pytest.raises(Exception):                      # >= 2 ops without the tmp buffer is refused, not guessed
        ops.augment_channels(src, C, tab, table.shape[1], max_ops, out, None)


def test_device_equals_the_golden_fixture_from_the_reference():
    z = np.load(GOLDEN)
    for name in ("all_64", "all_40x56", "affine_64", "flips_64"):
        lab, ref, seed = z[name + "_labels"], z[name + "_out"], int(z[name + "_seed"])
        tf = reference_transform(tuple(str(z[name + "_transforms"]).split(",")))
        random.seed(seed); torch.manual_seed(seed)
        sample_ops = [D.draw_channel_ops(tf, 19, lab.shape[1], lab.shape[0])]
        assert np.array_equal(run_device(lab[None], sample_ops)[0], ref), name


def test_device_equals_pillow_at_512():
    """the benchmark's plane size: 2 samples x 19 channels through Pillow/torchvision themselves on the same seeds."""
    h = w = 512
    tf = reference_transform(ALL)
    labs = np.stack([label_map(h, w, 40 + b) for b in range(2)])
    refs, sample_ops = [], []
    for b in range(2):
        random.seed(100 + b); torch.manual_seed(100 + b)
        refs.append(pil_per_channel(A.onehot(labs[b], 19), tf))
        random.seed(100 + b); torch.manual_seed(100 + b)
        sample_ops.append(D.draw_channel_ops(tf, 19, w, h))
    got = run_device(labs, sample_ops)
    for b in range(2):
        assert np.array_equal(got[b], refs[b]), (b, int((got[b] != refs[b]).sum()))


def test_plugin_takes_the_device_augmented_batch():
    """`set_input` with input_cloths = the device tensor of the augmenter and target_cloths = the uint8 label map gives
    the step the reference's host pipeline (dense fp32 tensors made by Pillow) gives."""
    from swapnet_b200.models import create_model
    from swapnet_b200.ops import SegMap
    from test_engine_gpu import _opt, _run_phases, relmax, synth_warp_batch

    B, S = 2, 64
    torch.manual_seed(0)
    model = create_model(_opt(B, S))
    model.setup(model.opt)
    model.is_train = True
    body, _, _ = synth_warp_batch(B, S)
    labs = np.stack([label_map(S, S, 60 + b) for b in range(B)])
    tf = reference_transform(ALL)
    aug = D.ClothAugmenter(tf, 19)
    host_in, sample_ops = [], []
    for b in range(B):
        random.seed(7 + b); torch.manual_seed(7 + b)
        host_in.append(pil_per_channel(A.onehot(labs[b], 19), tf))
        random.seed(7 + b); torch.manual_seed(7 + b)
        sample_ops.append(aug.draw(S, S))
    tgt = torch.from_numpy(np.stack([A.onehot(l, 19) for l in labs]))
    dense = dict(bodys=body, input_cloths=torch.from_numpy(np.stack(host_in)), target_cloths=tgt,
                 cloth_paths=["c"] * B, body_paths=["b"] * B)
    lab_dev = torch.from_numpy(labs).to(dev())
    on_dev = dict(dense, input_cloths=aug.apply(lab_dev, sample_ops), target_cloths=SegMap(lab_dev, 19))
    l0, gD0, gG0 = _run_phases(model, dense, 5)
    f0 = model.fakes.clone()
    l1, gD1, gG1 = _run_phases(model, on_dev, 5)
    assert torch.equal(f0, model.fakes)
    assert relmax(gD1, gD0) < 1e-5 and relmax(gG1, gG0) < 1e-5
    assert all(abs(l0[k] - l1[k]) <= 1e-6 * abs(l0[k]) for k in l0), (l0, l1)
    # ... and the batch format of `--dataset warp_b200` (dropin/datasets/warp_b200_dataset.py after the DataLoader's
    # default collate): label maps + encoded op tables on the HOST; set_input augments on the device
    ds_batch = dict(bodys=body, input_labels=torch.from_numpy(labs), target_labels=torch.from_numpy(labs.copy()),
                    input_ops=torch.stack([D.encode_sample(o, 4) for o in sample_ops]),
                    cloth_paths=["c"] * B, body_paths=["b"] * B)
    l2, gD2, gG2 = _run_phases(model, ds_batch, 5)
    assert torch.equal(model.inputs.cpu(), dense["input_cloths"]), "set_input's device augmentation differs from Pillow"
    assert torch.equal(model.dense(model.targets).cpu(), tgt)
    assert torch.equal(f0, model.fakes)
    assert relmax(gD2, gD0) < 1e-5 and relmax(gG2, gG0) < 1e-5
    assert all(abs(l0[k] - l2[k]) <= 1e-6 * abs(l0[k]) for k in l0), (l0, l2)
