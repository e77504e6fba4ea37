"""2-rank CUDA data-parallel equivalence (run under torchrun, one rank per GPU; tests/test_dp_cuda.py launches it).

Each rank runs the D and G phases of one WarpModel training step on ITS half of a batch through the plugin —
NCCL process group created by BaseModel.__init__ from torchrun's environment, gradients averaged by the bucketed
all-reduce that overlaps G-backward (parallel.BucketedAverager) — and compares the averaged flat gradient buffers of
D and G with the gradients of the whole batch computed by a single-process model (world forced to 1) on the same
weights, label draws and (training mode) dropout masks: masks follow the GLOBAL sample index (SURVEY §8e ii).
Prints `DP_EQUIV OK ...` on rank 0 when every rank agrees to 5e-4 (measured: 2e-5 .. 2e-4 — the two runs accumulate the
same bf16-split products in different orders: per-rank partial sums + all-reduce vs one pass of atomics).
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from test_engine_gpu import _opt, synth_warp_batch  # noqa: E402

from swapnet_b200 import parallel  # noqa: E402
from swapnet_b200.models import create_model  # noqa: E402


def relmax(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def phases(model, batch):
    model.set_input(batch)
    model._acc.zero_()
    model.forward()
    model._eng_Dd.zero_grad()
    model.backward_D()
    gD = model._eng_Dd.flat_grad.detach().clone()
    model._eng_G.zero_grad()
    model.backward_G()
    torch.cuda.synchronize()
    return gD, model._eng_G.flat_grad.detach().clone(), dict(model.get_current_losses())


def main():
    S = int(os.environ.get("SN_DP_SIZE", "256"))
    per = int(os.environ.get("SN_DP_PER_RANK", "2"))
    mode = os.environ.get("SN_DP_MODE", "train")
    world = int(os.environ["WORLD_SIZE"])
    torch.manual_seed(0)
    dp = create_model(_opt(per, S))                # BaseModel.__init__ creates the NCCL group; rank r -> cuda:r
    dp.setup(dp.opt)
    rank = dist.get_rank()
    assert dist.get_world_size() == world and dp.device.index == int(os.environ["LOCAL_RANK"])
    if mode == "eval":
        dp.eval()
    dp.is_train = True
    B = per * world
    body, inp, tgt = synth_warp_batch(B, S)
    full = dict(bodys=body, input_cloths=inp, target_cloths=tgt, cloth_paths=["c"] * B, body_paths=["b"] * B)
    gD, gG, losses = phases(dp, parallel.shard_batch(full, rank, world))
    assert dp._eng_G.sample_base == rank * per
    # the all-reduce leaves the SUM over ranks in the flat buffers; the 1/world factor is applied by the fused AdamW
    # kernel as it reads them (BaseGAN.grad_scale)
    assert dp.grad_scale() == 1.0 / world
    gD, gG = gD * dp.grad_scale(), gG * dp.grad_scale()

    # single-process reference on the same (broadcast) weights: world forced to 1, no all-reduce
    torch.manual_seed(0)
    ref = create_model(_opt(B, S, name="warp_ref"))
    ref.setup(ref.opt)
    for a, b in zip(list(dp.net_generator.parameters()) + list(dp.net_discriminator.parameters()),
                    list(ref.net_generator.parameters()) + list(ref.net_discriminator.parameters())):
        assert torch.equal(a, b), "replicas diverged from the single-process initialisation"
    ref._world = 1
    ref.allreduce_grads = lambda eng: None
    ref._labels = parallel.LabelDraws(1234)        # the same draws as the DP run
    ref.ensure_engines(B, S)
    ref._eng_G.sample_base = 0
    if mode == "eval":
        ref.eval()
    ref.is_train = True
    rD, rG, rlosses = phases(ref, full)
    eD, eG = relmax(gD, rD), relmax(gG, rG)
    # each rank's loss values are means over ITS shard: their mean over ranks is the full-batch loss
    lt = torch.tensor([losses[k] for k in sorted(losses)], dtype=torch.float64, device=dp.device)
    dist.all_reduce(lt)
    el = max(abs(v / world - rlosses[k]) / abs(rlosses[k]) for v, k in zip(lt.tolist(), sorted(losses)))
    worst = torch.tensor([eD, eG, el], dtype=torch.float64, device=dp.device)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    if rank == 0:
        ok = bool((worst[:2] < 5e-4).all() and worst[2] < 1e-5)
        print(f"DP_EQUIV {'OK' if ok else 'FAIL'} world={world} size={S} per_rank={per} mode={mode} "
              f"flat_grad_D={worst[0].item():.3e} flat_grad_G={worst[1].item():.3e} losses={worst[2].item():.3e}",
              flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if not (worst[:2] < 5e-4).all():
        sys.exit(1)


if __name__ == "__main__":
    main()
