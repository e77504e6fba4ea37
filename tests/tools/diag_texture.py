"""Stage-by-stage comparison of TextureEngine against the CPU oracle (fp64)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import nets as ON
from swapnet_b200 import engine as E, ops
from test_engine_gpu import make_texture_net, synth_texture_batch, relmax, stage_gates

dev = torch.device("cuda:0")
B, S = 2, int(sys.argv[1]) if len(sys.argv) > 1 else 128
T = make_texture_net(S)
sd = {k: v.clone().double().requires_grad_() for k, v in T.state_dict().items()}
tex, rois, cloth, _ = synth_texture_batch(B, S)
T.to(dev)
nhwc = lambda t: t.permute(0, 2, 3, 1)
eng = E.TextureEngine(T, B, S, dev); eng.alloc_grads(); eng.bind_backward(); eng.pack()
fakes = eng.forward(tex.to(dev), rois.to(dev), cloth.to(dev), training=False)
torch.cuda.synchronize()
gates = stage_gates(eng)
ON.gate_with(lambda name, x: gates.get(name))
rec = {}
ON.record_into(rec)
ref = ON.texture_forward(sd, tex.double(), rois.double(), cloth.double())
ON.record_into(None); ON.gate_with(None)
gout = torch.randn(ref.shape, generator=torch.Generator().manual_seed(5)).double() * 1e-3
ref.backward(gout)
eng.zero_grad()
eng.backward([ops.GradSrc(nhwc(gout).float().contiguous().to(dev))])
torch.cuda.synchronize()
print(f"== TextureEngine S={S}: fakes relmax {relmax(fakes.cpu(), nhwc(ref.detach())):.2e}")
names = {k: v for k, v in T.named_parameters()}
pmap = {id(p): k for k, p in T.named_parameters()}
for st in eng.stages:
    r = rec[st.name + ".y"]
    line = f"{st.name:12s} y {relmax(st.y.cpu(), nhwc(r.detach())):.2e}"
    if not st.plain:
        line += f"  dy {relmax(st.dy.dense().cpu()[..., :st.cout], nhwc(r.grad)):.2e}"
    k = pmap[id(st.conv.weight)]
    line += f"  wgrad {relmax(st.conv.weight.grad.cpu(), sd[k].grad):.2e}"
    if st.conv.bias is not None:
        kb = pmap[id(st.conv.bias)]
        line += f"  bgrad {float((st.conv.bias.grad.cpu().double() - sd[kb].grad).abs().max()):.2e}/{float(sd[kb].grad.abs().max()):.2e}"
    print(line)
