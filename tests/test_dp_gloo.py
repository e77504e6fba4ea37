"""Data-parallel equivalence on CPU (gloo, world_size 2): sharding the batch over ranks and averaging
the flat gradient buffer reproduces the single-process gradient of the concatenated batch — the
property the B200 DP path relies on (per-sample InstanceNorm, batch-mean losses) — and
every rank draws the same smooth GAN label."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import nets as ON
from swapnet_b200 import modules as M
from swapnet_b200 import parallel
from test_engine_gpu import synth_warp_batch


def _d_grads(sdD, body, cloth_fake, cloth_real, t):
    lf = ON.gan_loss(ON.patchgan_forward(sdD, torch.cat((body, cloth_fake), 1)), t[0])
    lr = ON.gan_loss(ON.patchgan_forward(sdD, torch.cat((body, cloth_real), 1)), t[1])
    g = torch.autograd.grad(0.5 * (lf + lr), list(sdD.values()))
    return torch.cat([x.reshape(-1) for x in g])


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    D = M.NLayerDiscriminator(22, 64, 3, "instance")
    M.init_weights(D, "kaiming")
    if rank != 0:  # broadcast must repair divergent replicas
        for p in D.parameters():
            p.data.add_(1.0)
    parallel.broadcast_parameters(D.parameters())
    sdD = {k: v.detach().clone().requires_grad_() for k, v in D.state_dict().items()}
    B = 4
    body, inp, tgt = synth_warp_batch(B, 64)
    full = dict(bodys=body, input_cloths=inp, target_cloths=tgt, cloth_paths=["c"] * B, body_paths=["b"] * B)
    mine = parallel.shard_batch(full, rank, world)
    assert mine["bodys"].shape[0] == B // world and len(mine["cloth_paths"]) == B // world
    labels = parallel.LabelDraws(1234)
    t = [torch.tensor([labels.draw()]), torch.tensor([labels.draw()])]
    flat = _d_grads(sdD, mine["bodys"], mine["input_cloths"], mine["target_cloths"], t)
    flat2 = flat.clone()
    parallel.average_gradients(flat)
    # the bucketed (overlapped) averager must give the same result as the single all-reduce
    half = flat2.numel() // 2
    avg = parallel.BucketedAverager(flat2, [(half, flat2.numel()), (0, half)])
    avg.ready(0)
    avg.ready(1)
    avg.finish()
    assert torch.allclose(flat, flat2, rtol=0, atol=0)
    if rank == 0:
        ref = _d_grads(sdD, body, inp, tgt, t)
        q.put((float((flat - ref).abs().max() / ref.abs().max()), [float(x) for x in t]))
    else:
        q.put((None, [float(x) for x in t]))
    dist.destroy_process_group()


def test_gradient_average_equals_full_batch_gradient():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    errs = [o[0] for o in outs if o[0] is not None]
    assert errs and errs[0] < 1e-5, errs
    assert outs[0][1] == outs[1][1], "ranks drew different smooth labels"


def test_single_process_helpers_are_noops():
    x = torch.arange(4.0)
    parallel.average_gradients(x)
    assert torch.equal(x, torch.arange(4.0)) and parallel.world_size() == 1 and parallel.rank() == 0
