"""Stage-by-stage comparison of the engines against the CPU oracle (fp64): conv outputs y,
dL/dy, dL/d(input) — prints relmax per stage.  Diagnostic; run on the GPU box."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn.functional as F
from oracle import nets as ON
from swapnet_b200 import engine as E, ops
from test_engine_gpu import make_nets, synth_warp_batch, relmax, stage_gates

dev = torch.device("cuda:0")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
B = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 2
G, D = make_nets()
sdG = {k: v.clone().double().requires_grad_() for k, v in G.state_dict().items()}
sdD = {k: v.clone().double().requires_grad_() for k, v in D.state_dict().items()}
body, inp, tgt = synth_warp_batch(B, S)
G.to(dev); D.to(dev)
nhwc = lambda t: t.permute(0, 2, 3, 1)

# ---------------- generator ----------------
eng = E.WarpEngine(G, B, S, dev); eng.alloc_grads(); eng.bind_backward(); eng.pack()
fakes = eng.forward(body.to(dev), inp.to(dev), training=False)
rec = {}
ON.record_into(rec)
torch.cuda.synchronize()
gatesG = stage_gates(eng)          # the oracle evaluates the activations at the gates the device used (see tests)
ON.gate_with(lambda name, x: gatesG.get(name))
ref = ON.warp_forward(sdG, body.double(), inp.double())
print("gate flips G:", {k: v for k, v in ON.GATE_STATS.items() if v and k != "__total__"}, "of", ON.GATE_STATS.get("__total__"))
ON.gate_with(None)
ON.record_into(None)
gout = torch.randn(ref.shape, generator=torch.Generator().manual_seed(5)).double() * 1e-3
ref.backward(gout)
rec_G = rec
eng.zero_grad()
eng.backward([ops.GradSrc(nhwc(gout).float().contiguous().to(dev))])
torch.cuda.synchronize()
print(f"== WarpEngine S={S}: fakes relmax {relmax(fakes.cpu(), nhwc(ref.detach())):.2e}")
for st in eng.stages:
    key = st.name + ".y"
    r = rec[key]
    line = f"{st.name:28s} y {relmax(st.y.cpu(), nhwc(r.detach())):.2e}"
    if st.plain:
        pass
    else:
        line += f"  dy {relmax(st.dy.dense().cpu()[..., :st.cout], nhwc(r.grad)):.2e}"
    wkey = [k for k in sdG if k.startswith(st.name.replace('.conv1', '.conv_block.1').replace('.conv2', '.conv_block.6')) and k.endswith("weight")]
    if wkey:
        line += f"  wgrad {relmax(st.conv.weight.grad.cpu(), sdG[wkey[0]].grad):.2e}"
    print(line)

print(f"head.dx[:, :64] (gradient w.r.t. dual_up3's activation) relmax "
      f"{relmax(eng.head.dx[..., :64].cpu(), nhwc(rec_G['dual_up3.a'].grad)):.2e}")
print(f"dual_up3.dx[:, :128] (w.r.t. dual_up2's activation) relmax "
      f"{relmax(eng.d3.dx[..., :128].cpu(), nhwc(rec_G['dual_up2.a'].grad)):.2e}")
print(f"dual_up3 stats: mean relmax {relmax(eng.d3.stats[..., 0].cpu(), rec_G['dual_up3.y'].detach().mean((2, 3))):.2e} "
      f"rstd relmax {relmax(eng.d3.stats[..., 1].cpu(), (rec_G['dual_up3.y'].detach().var((2, 3), unbiased=False) + 1e-5).rsqrt()):.2e}")

# the IN + ReLU backward of dual_up3 recomputed by torch ON THE DEVICE BUFFERS the kernel read (y, upstream gradient)
st = eng.d3
yd = st.y.permute(0, 3, 1, 2).double().requires_grad_()
up = eng.head.dx[..., :64].permute(0, 3, 1, 2).double()
(gt,) = torch.autograd.grad(F.relu(F.instance_norm(yd, eps=1e-5)), yd, up)
gt = gt.permute(0, 2, 3, 1)
print(f"dual_up3: kernel dy vs torch-on-device-buffers {relmax(st.dy.dense()[..., :64], gt):.2e} | torch-on-device-buffers vs oracle "
      f"{relmax(gt.cpu(), nhwc(rec_G['dual_up3.y'].grad)):.2e}")
dy2 = ops.Planes(st.n, st.oh, st.ow, 64, dev, fmt=ops.FMT_BF16)
gs2 = torch.zeros(st.n, st.cout, 2, dtype=torch.float64, device=dev)
ops.norm_act_bwd([ops.GradSrc(eng.head.dx, 0)], st.y, st.cout, st.stats, st.act, dy2, gs2, st.slope, 0.0, 0)
torch.cuda.synchronize()
print(f"dual_up3: a second kernel call on the same buffers vs torch {relmax(dy2.dense()[..., :64], gt):.2e}; vs the engine's dy "
      f"{relmax(dy2.dense()[..., :64], st.dy.dense()[..., :64]):.2e}")

# ---------------- discriminator ----------------
x = torch.cat((body, ref.detach().float()), 1)
Dd = E.PatchGANEngine(D, B, S, dev, input_grad=True); Dd.alloc_grads(); Dd.bind_backward(); Dd.pack()
ops.pack_planes(x.to(dev), Dd.din.slice(0, 22))
pred = Dd.forward()
rec = {}
ON.record_into(rec)
xr = x.double().requires_grad_()
torch.cuda.synchronize()
gatesD = stage_gates(Dd)
ON.gate_with(lambda name, x: gatesD.get(name))
pr = ON.patchgan_forward(sdD, xr)
print("gate flips D:", {k: v for k, v in ON.GATE_STATS.items() if v and k != "__total__"}, "of", ON.GATE_STATS.get("__total__"))
ON.gate_with(None)
ON.record_into(None)
gp = torch.randn(pr.shape, generator=torch.Generator().manual_seed(6)).double() * 1e-3
pr.backward(gp)
Dd.zero_grad()
Dd.backward(nhwc(gp).float().contiguous().to(dev))
torch.cuda.synchronize()
print(f"== PatchGAN: pred relmax {relmax(pred.cpu(), nhwc(pr.detach())):.2e}  dx_in {relmax(Dd.dx_in.cpu()[..., :22], nhwc(xr.grad)):.2e}")
for st in Dd.stages:
    r = rec[st.name + ".y"]
    line = f"{st.name:28s} y {relmax(st.y.cpu(), nhwc(r.detach())):.2e}"
    line += f"  dy {relmax(st.dy.dense().cpu()[..., :st.cout], nhwc(r.grad)):.2e}"
    line += f"  wgrad {relmax(st.conv.weight.grad.cpu(), sdD[st.name + '.weight'].grad):.2e}"
    if st.conv.bias is not None:
        bref = sdD[st.name + '.bias'].grad
        line += f"  bgrad abs {float((st.conv.bias.grad.cpu().double() - bref).abs().max()):.2e} (ref max {float(bref.abs().max()):.2e})"
    if st.dx is not None and not st.plain:
        pass
    print(line)

print(f"last.dx (gradient w.r.t. model.8's activation) relmax {relmax(Dd.last.dx.cpu(), nhwc(rec['model.8.a'].grad)):.2e}")
st8 = Dd.chain[-1]
print(f"model.8 stats: mean relmax {relmax(st8.stats[..., 0].cpu(), rec['model.8.y'].detach().mean((2, 3))):.2e} "
      f"rstd relmax {relmax(st8.stats[..., 1].cpu(), (rec['model.8.y'].detach().var((2, 3), unbiased=False) + 1e-5).rsqrt()):.2e}")

# ---------------- focus: dual_up2 backward ----------------
if "--focus" in sys.argv:
    st = eng.d2
    g_in = eng.d3.dx[..., :128].cpu()
    ra = rec_G["dual_up2.a"].grad
    print(f"upstream grad into dual_up2 (d3.dx[:128]) relmax {relmax(g_in, nhwc(ra)):.2e}")
    ry = rec_G["dual_up2.y"]
    ref_dy = nhwc(ry.grad)
    got = st.dy.dense().cpu()[..., :st.cout].double()
    err_c = (got - ref_dy).abs().amax((0, 1, 2))
    top = torch.topk(err_c, 5).indices.tolist()
    mean_ref = ry.detach().mean((2, 3)); var_ref = ry.detach().var((2, 3), unbiased=False)
    print("worst channels", top, "err", [f"{err_c[c]:.2e}" for c in top], "ref max|dy|", f"{ref_dy.abs().max():.2e}")
    for c in top:
        print(f" ch {c}: ref mean {mean_ref[:, c].tolist()} var {var_ref[:, c].tolist()} | eng mean {st.stats[:, c, 0].tolist()} rstd {st.stats[:, c, 1].tolist()} (ref rstd {(var_ref[:, c] + 1e-5).rsqrt().tolist()})")
        print(f"        ref dy absmax per n {ref_dy[..., c].abs().amax((1,2)).tolist()}  got {got[..., c].abs().amax((1,2)).tolist()}")
