#!/usr/bin/env python
"""SwapNet warp-stage training throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference ...                     # the reference's CPU path (baseline arm)
    torchrun --nproc-per-node N bench.py --gpus N ...        # data parallel, one rank per GPU

One "step" = one WarpModel.optimize_parameters() (G fwd, D step, G step incl. both AdamW updates —
the full reference training step, models/warp_model.py:169-183) on a synthetic batch of
`--batch` 512x512 images per GPU (BASELINE.json configs[1]: warp_model 512x512, batch 16, 1xB200).
Prints ONE JSON line (rank 0).  Field notes:
  value     images/s, whole job, inputs already resident in HBM, K steps timed with CUDA events;
  e2e       same metric through the plugin API with HOST (pinned) input tensors: the timed region
            has, every step, the H2D copy of the batch (set_input) and the D2H read of the six
            losses (get_current_losses, as train.py:62-74 does);
  roofline  dominant kernel class = the tcgen05 tap-GEMM (`tap_gemm_kernel<3>`: forward + dgrad
            launches): algorithmic conv FLOPs of those launches / their summed CUDA-event time,
            against the measured dense bf16 peak.  The kernel issues 3 MMAs per algorithmic MAC
            (fp16/bf16-split fp32-faithful product), so frac <= 1/3 by construction; `pipe_frac`
            is the tensor-pipe view (3x).
  cpu_baseline  the CPU oracle port (oracle/nets.py, pinned bit-exactly to the reference modules)
            running the same training step at 512x512, batch 1, on the host cores the cgroup CPU quota
            allows (host_cores()), a few steps; `--impl reference` times the same port as the reference arm
            (K <= 10, W <= 2 so that the run stays within a few minutes).
  --model texture [--perceptual]   informational: the texture stage (BASELINE configs[2]), not the headline.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FULL_STEP_GFLOP_PER_IMG_512 = 1029.0   # full reference warp training step, nominal


def synth_batch(B, S, seed):
    """normalised-RGB-like body, 16x16-block one-hot cloth (label 0 = all-zero),
    input cloth = target rolled by (8, 8)."""
    g = torch.Generator().manual_seed(seed)
    body = torch.rand(B, 3, S, S, generator=g) * 4.8 - 0.31
    lab = torch.randint(0, 19, (B, S // 16, S // 16), generator=g).repeat_interleave(16, 1).repeat_interleave(16, 2)
    tgt = torch.zeros(B, 19, S, S)
    for c in range(1, 19):
        tgt[:, c] = (lab == c).float()
    inp = torch.roll(tgt, (8, 8), (2, 3))
    return dict(bodys=body, input_cloths=inp, target_cloths=tgt, cloth_paths=["synthetic"] * B,
                body_paths=["synthetic"] * B)


def synth_texture_batch(B, S, seed):
    """normalised-RGB-like textures, one-hot cloth, rois = notebook fixture (256 px
    space, incl. degenerate rows) scaled to S and rotated per sample."""
    g = torch.Generator().manual_seed(seed)
    tex = torch.rand(B, 3, S, S, generator=g) * 4.5 - 2.0
    tgt = torch.rand(B, 3, S, S, generator=g) * 4.5 - 2.0
    lab = torch.randint(0, 19, (B, S // 16, S // 16), generator=g).repeat_interleave(16, 1).repeat_interleave(16, 2)
    cloth = torch.zeros(B, 19, S, S)
    for c in range(1, 19):
        cloth[:, c] = (lab == c).float()
    base = torch.tensor([[159, 0, 193, 14], [144, 15, 206, 89], [255, 0, 255, 0], [196, 20, 215, 94],
                         [144, 151, 180, 229], [179, 151, 216, 226], [156, 1, 188, 24], [141, 83, 215, 155],
                         [128, 20, 160, 82], [206, 92, 226, 158], [145, 220, 168, 255], [174, 217, 203, 255]],
                        dtype=torch.float32) * (S / 256.0)
    rois = torch.stack([torch.roll(base, b, 0) for b in range(B)])
    return dict(input_textures=tex, rois=rois, cloths=cloth, target_textures=tgt, cloth_paths=["synthetic"] * B,
                texture_paths=["synthetic"] * B)


def warp_opt(B, S, precision):
    return argparse.Namespace(
        model="warp", gpu_id=int(os.environ.get("LOCAL_RANK", 0)), is_train=True,
        checkpoints_dir=tempfile.mkdtemp(prefix="sn_bench_"), name="warp", no_confirm=True,
        body_representation="rgb", body_channels=12, cloth_representation="labels", cloth_channels=19,
        texture_channels=3, init_type="kaiming", init_gain=0.02, discriminator="basic", n_layers_D=3,
        norm="instance", gan_mode="vanilla", gan_label_mode="smooth", lambda_gan=1.0, lambda_discriminator=1.0,
        lambda_gp=10, optimizer_G="AdamW", optimizer_D="AdamW", lr=1e-4, d_lr=4e-4, weight_decay=0,
        d_weight_decay=0.01, b1=0.9, b2=0.999, warp_mode="gan", lambda_ce=100, continue_train=False,
        load_epoch="latest", verbose=False, batch_size=B, crop_size=S, load_size=S, b200_precision=precision)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.rows.append(f)
            except Exception:
                pass
            self._halt.wait(0.05)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def ncu_traffic():
    """dram__bytes_read + dram__bytes_write per launch of the dominant kernel, from the committed
    `ncu --set full` capture (profiles/r01_ncu_traffic.json); null if absent."""
    p = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
    if os.path.exists(p):
        return json.load(open(p))
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", d.get("bf16_tflops")), d.get("hbm_gbs"), "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference training step restated in oracle/nets.py (bit-identical to the reference
# modules), with torch.optim.AdamW as optimizers/__init__.py builds it
# ------------------------------------------------------------------------------------------------
def host_cores() -> int:
    """CPUs this process may actually use: the scheduler affinity capped by the cgroup CPU quota (the GPU boxes
    expose 128 logical CPUs under a 16-CPU quota — 128 torch threads there are 8x oversubscribed)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    quota = None
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()[:2]          # cgroup v2
        if q != "max":
            quota = int(q) / int(p)
    except (OSError, ValueError):
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())      # cgroup v1
            p = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                quota = q / p
        except (OSError, ValueError):
            pass
    if quota is not None:
        n = min(n, max(1, int(quota + 0.999)))
    return max(1, n)


def cpu_reference_run(S, B, steps, warmup):
    from oracle import nets as ON
    from swapnet_b200 import modules as M

    import contextlib

    torch.set_num_threads(host_cores())
    torch.manual_seed(0)
    with contextlib.redirect_stdout(sys.stderr):
        G = M.WarpModule()
        M.init_weights(G, "kaiming")
        D = M.NLayerDiscriminator(22, 64, 3, "instance")
        M.init_weights(D, "kaiming")
    sdG = {k: v.detach().clone().requires_grad_() for k, v in G.state_dict().items()}
    sdD = {k: v.detach().clone().requires_grad_() for k, v in D.state_dict().items()}
    optG = torch.optim.AdamW(list(sdG.values()), lr=1e-4, weight_decay=0, betas=(0.9, 0.999))
    optD = torch.optim.AdamW(list(sdD.values()), lr=4e-4, weight_decay=0.01, betas=(0.9, 0.999))
    b = synth_batch(B, S, 1234)
    body, inp, tgt = b["bodys"], b["input_cloths"], b["target_cloths"]
    g = torch.Generator().manual_seed(7)

    def drop(name, x):  # training-mode dropout(0.5) as in the reference (cost parity; masks irrelevant)
        return torch.nn.functional.dropout(x, 0.5, True)

    def step():
        fakes = ON.warp_forward(sdG, body, inp, drop)
        optD.zero_grad()
        t = [ON.smooth_label(torch.rand(1, generator=g)) for _ in range(3)]
        lf = ON.gan_loss(ON.patchgan_forward(sdD, torch.cat((body, fakes), 1).detach()), t[0])
        lr = ON.gan_loss(ON.patchgan_forward(sdD, torch.cat((body, tgt), 1)), t[1])
        (0.5 * (lf + lr)).backward()
        optD.step()
        optG.zero_grad()
        ce = torch.nn.functional.cross_entropy(fakes, torch.argmax(tgt, 1)) * 100
        gan = ON.gan_loss(ON.patchgan_forward(sdD, torch.cat((body, fakes), 1)), t[2])
        (ce + gan).backward()
        optG.step()
        return float((ce + gan).detach())

    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    med = ts[len(ts) // 2]
    return B / med, med


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=("b200", "reference"))
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--batch", type=int, default=16, help="images per GPU")
    ap.add_argument("--precision", default="fp32x3", choices=("fp32x3", "bf16"))
    ap.add_argument("--cpu-steps", type=int, default=3,
                    help="timed CPU-oracle steps of the cpu_baseline leg (one step = ~5 s on the usable host cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--perceptual", action="store_true",
                    help="--model texture: add the VGG16 content + Gram style terms (lambda 20 / 1e-8)")
    ap.add_argument("--model", default="warp", choices=("warp", "texture"),
                    help="warp = the BASELINE.json metric (default); texture = configs[2] (informational)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    S, B = args.size, args.batch
    cores = host_cores()
    workload = f"warp_model {S}x{S} synthetic, batch {B}/GPU, full GAN step (G fwd, D step, G step, AdamW x2)"

    if args.impl == "reference":
        if rank != 0:
            return
        # one CPU step at 512x512, batch 1, takes ~5 s on the 16 usable cores of a GPU box: K and W are clamped so
        # that the run ends within a few minutes (the line reports the steps actually timed)
        steps = max(1, min(args.steps, 10))
        v, med = cpu_reference_run(S, 1, steps, min(args.warmup, 2))
        print(json.dumps({
            "impl": "reference", "metric": "images/sec (G+D fwd+bwd) warp-stage 512x512", "value": v,
            "unit": "images/s", "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 2),
            "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "note": "CPU arm runs batch 1 per step (bounded sample of the same workload)",
                       "steps_requested": args.steps, "warmup_requested": args.warmup},
            "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": f"{steps} timed full training steps at {S}x{S}, batch 1, torch CPU ({cores} threads)"},
            "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback of the hot path)")
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    from swapnet_b200 import ops
    from swapnet_b200.models import create_model

    torch.manual_seed(0)
    import contextlib

    with contextlib.redirect_stdout(sys.stderr):   # stdout carries exactly one JSON line
        o = warp_opt(B, S, args.precision)
        if args.model == "texture":
            o.model, o.name, o.netG, o.lambda_l1, o.lambda_content, o.lambda_style = "texture", "texture", "swapnet", 10, 0, 0
            if args.perceptual:   # the reference's default texture losses; seeded-random VGG16 (no weight file offline)
                o.lambda_content, o.lambda_style, o.b200_vgg = 20.0, 1e-8, "random"
        model = create_model(o)
        model.setup(model.opt)
    if args.model == "texture":
        host = synth_texture_batch(B, S, 1234 + rank)
        tkeys = ("input_textures", "rois", "cloths", "target_textures")
    else:
        host = synth_batch(B, S, 1234 + rank)
        tkeys = ("bodys", "input_cloths", "target_cloths")
    for k in tkeys:
        host[k] = host[k].pin_memory()
    dev_batch = dict(host)
    for k in tkeys:
        dev_batch[k] = host[k].cuda(non_blocking=True)
    h2d = sum(host[k].numel() * 4 for k in tkeys)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(n, batch, read_losses):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            model.set_input(batch)
            model.optimize_parameters()
            if read_losses:
                model.get_current_losses()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return ms.item()

    for _ in range(max(args.warmup, 3)):
        model.set_input(dev_batch)
        model.optimize_parameters()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ops.launch_count()
    ms = timed(args.steps, dev_batch, False)
    launches = ops.launch_count() - l0
    clocks = sampler.stop()
    ms_e2e = timed(args.steps, host, True)

    # ---- roofline pass (untimed): per-launch CUDA events on the tap-GEMM kernel ----
    roof = None
    # every rank runs the traced step (it contains the gradient all-reduces); rank 0 evaluates it
    ops.Plan.trace = []
    model.set_input(dev_batch)
    model.optimize_parameters()
    torch.cuda.synchronize()
    trace, ops.Plan.trace = ops.Plan.trace, None
    if rank == 0:
        info = {}
        engs = [model._eng_G, model._eng_Dd, model._eng_Dg]
        pe = getattr(model, "_eng_P", None)
        if pe is not None and pe.out is not None:
            engs += [pe.out, pe.tgt]
        for eng in engs:
            for st in eng.stages:
                fl, ly = 2.0 * st.nominal_macs(), st.layer
                for p in ly.fwd_plans:
                    info[id(p)] = ("fwd", fl / len(ly.fwd_plans))
                for p in ly.dgrad_plans:
                    info[id(p)] = ("dgrad", fl / len(ly.dgrad_plans))
                if ly.wgrad_plan is not None:
                    info[id(ly.wgrad_plan)] = ("wgrad", fl)
        peak, hbm, how = measured_peaks()
        tot = {"fwd": [0.0, 0.0, 0], "dgrad": [0.0, 0.0, 0], "wgrad": [0.0, 0.0, 0]}  # flops, ms, launches
        for plan, a, b_ in trace:
            kind, fl = info[id(plan)]
            tot[kind][0] += fl
            tot[kind][1] += a.elapsed_time(b_)
            tot[kind][2] += 1
        gemm_fl = tot["fwd"][0] + tot["dgrad"][0]
        gemm_ms = tot["fwd"][1] + tot["dgrad"][1]
        gemm_n = tot["fwd"][2] + tot["dgrad"][2]
        step_ms = ms / args.steps
        ach = gemm_fl / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        roof = {"bound": "tensor", "kernel": "tap_gemm_kernel<3> (tcgen05, fwd+dgrad launches)", "achieved": ach,
                "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "pipe_frac": 3 * ach / peak, "peak_source": how,
                "launches_per_step": gemm_n, "avg_launch_ms": gemm_ms / max(gemm_n, 1),
                "algorithmic_gflop_per_launch": gemm_fl / max(gemm_n, 1) / 1e9,
                "share_of_step": gemm_ms / step_ms,
                # DRAM bytes of one launch of the dominant shape from the committed `ncu --set full` capture
                "traffic": (ncu_traffic() or {}).get("dram_bytes_per_launch"), "traffic_detail": ncu_traffic(),
                "wgrad_kernel": {"achieved": (tot["wgrad"][0] / (tot["wgrad"][1] * 1e-3) / 1e12) if tot["wgrad"][1] else 0.0,
                                 "share_of_step": tot["wgrad"][1] / step_ms, "launches_per_step": tot["wgrad"][2]}}

    if rank != 0:
        torch.distributed.destroy_process_group()
        return
    cpu = None
    if args.model == "texture":
        workload = (f"texture_model {S}x{S} synthetic, 12-ROI, batch {B}/GPU, full GAN step "
                    + ("(L1 + GAN + VGG16 content + Gram style, seeded-random VGG weights)" if args.perceptual
                       else "(L1 + GAN; perceptual terms off)"))
    if not args.no_cpu_baseline and args.gpus == 1 and args.model == "warp":
        v, med = cpu_reference_run(S, 1, args.cpu_steps, 1)
        cpu = {"value": v, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": f"{args.cpu_steps} timed full training step(s) at {S}x{S}, batch 1, after 1 warm-up step, "
                         f"torch CPU ({cores} threads = usable cores under the cgroup quota)"}
    step_ms = ms / args.steps
    total_imgs = B * world
    out = {
        "metric": "images/sec (G+D fwd+bwd) warp-stage 512x512" if args.model == "warp"
        else "images/sec (G+D fwd+bwd) texture-stage 512x512",
        "value": total_imgs / (step_ms * 1e-3),
        "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp16/bf16-split x3 tensor-core products, fp32 accumulate (fp32-faithful)" if args.precision == "fp32x3"
        else "fp16 single-pass tensor-core products, fp32 accumulate",
        "data": "synthetic",
        "config": {"workload": workload, "global_batch": total_imgs, "parallelism": f"dp{world}",
                   "l2": "inputs+activations per step (>2 GB) exceed the 126 MB L2; no explicit flush",
                   "algorithmic_tflop_per_step": (FULL_STEP_GFLOP_PER_IMG_512 if args.model == "warp" else (896.0 if args.perceptual else 415.0))
                   * (S / 512) ** 2 * total_imgs / 1e3},
        "e2e": {"value": total_imgs / (ms_e2e / args.steps * 1e-3), "unit": "images/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 6 * 8},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
    }
    print(json.dumps(out))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
