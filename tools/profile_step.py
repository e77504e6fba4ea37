"""One training step between cudaProfilerStart/Stop (for `ncu --profile-from-start off`), plus — with
SN_TRACE=1 — a per-plan table (CUDA events around every tap-GEMM / wgrad-GEMM launch) of one more step.

env: SN_MODEL=warp|texture  SN_PERCEPTUAL=1  SN_B  SN_S  SN_PREC  SN_TRACE
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_batch, synth_texture_batch, warp_opt
from swapnet_b200 import ops
from swapnet_b200.models import create_model

B = int(os.environ.get("SN_B", 16)); S = int(os.environ.get("SN_S", 512))
kind = os.environ.get("SN_MODEL", "warp")
torch.manual_seed(0)
o = warp_opt(B, S, os.environ.get("SN_PREC", "fp32x3"))
if kind == "texture":
    o.model, o.name, o.netG, o.lambda_l1, o.lambda_content, o.lambda_style = "texture", "texture", "swapnet", 10, 0, 0
    if os.environ.get("SN_PERCEPTUAL"):
        o.lambda_content, o.lambda_style, o.b200_vgg = 20.0, 1e-8, "random"
    batch = synth_texture_batch(B, S, 1234)
else:
    batch = synth_batch(B, S, 1234)
model = create_model(o)
model.setup(model.opt)
for k, v in batch.items():
    if torch.is_tensor(v):
        batch[k] = v.cuda()
for _ in range(2):
    model.set_input(batch); model.optimize_parameters()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
model.set_input(batch); model.optimize_parameters()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("losses", model.get_current_losses())
if os.environ.get("SN_TRACE"):
    engs = [("G", model._eng_G), ("Dd", model._eng_Dd), ("Dg", model._eng_Dg)]
    pe = getattr(model, "_eng_P", None)
    if pe is not None and pe.out is not None:
        engs += [("Po", pe.out), ("Pt", pe.tgt)]
    info = {}
    for en, eng in engs:
        for st in eng.stages:
            fl, ly = 2.0 * st.nominal_macs(), st.layer
            for p in ly.fwd_plans:
                info[id(p)] = (f"{en}.{st.name}.fwd", fl / len(ly.fwd_plans))
            for p in ly.dgrad_plans:
                info[id(p)] = (f"{en}.{st.name}.dgrad", fl / len(ly.dgrad_plans))
            if ly.wgrad_plan is not None:
                info[id(ly.wgrad_plan)] = (f"{en}.{st.name}.wgrad", fl)
    ops.Plan.trace = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    model.set_input(batch); model.optimize_parameters()
    e1.record()
    torch.cuda.synchronize()
    trace, ops.Plan.trace = ops.Plan.trace, None
    rows = {}
    for plan, a, b in trace:
        name, fl = info[id(plan)]
        r = rows.setdefault(name, [0.0, 0.0, 0])
        r[0] += a.elapsed_time(b); r[1] += fl; r[2] += 1
    tot = sum(r[0] for r in rows.values())
    print(f"traced step {e0.elapsed_time(e1):.2f} ms; GEMM plans {tot:.2f} ms")
    for name, (ms, fl, n) in sorted(rows.items(), key=lambda kv: -kv[1][0]):
        print(f"{name:42s} {ms:8.3f} ms  x{n}  {fl / ms / 1e9:8.1f} TFLOP/s")
