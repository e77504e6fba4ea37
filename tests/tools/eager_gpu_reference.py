"""Context measurement (not a product path, not bench.py): the reference-equivalent EAGER PyTorch training
step (oracle/nets.py == the reference modules bit for bit, torch.optim.AdamW) on the same GPU, with the
cuDNN TF32 switch on (torch default, what `python train.py` uses on a GPU) and off (true fp32).
    python tests/tools/eager_gpu_reference.py [batch] [size]
"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from bench import synth_batch
from oracle import nets as ON
from swapnet_b200 import modules as M

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
S = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda:0")
out = {}
for tf32 in (True, False):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    G = M.WarpModule(); M.init_weights(G, "kaiming")
    D = M.NLayerDiscriminator(22, 64, 3, "instance"); M.init_weights(D, "kaiming")
    sdG = {k: v.detach().to(dev).requires_grad_() for k, v in G.state_dict().items()}
    sdD = {k: v.detach().to(dev).requires_grad_() for k, v in D.state_dict().items()}
    optG = torch.optim.AdamW(list(sdG.values()), lr=1e-4, weight_decay=0, betas=(0.9, 0.999))
    optD = torch.optim.AdamW(list(sdD.values()), lr=4e-4, weight_decay=0.01, betas=(0.9, 0.999))
    b = synth_batch(B, S, 1234)
    body, inp, tgt = (b[k].to(dev) for k in ("bodys", "input_cloths", "target_cloths"))
    drop = lambda name, x: torch.nn.functional.dropout(x, 0.5, True)

    def step():
        fakes = ON.warp_forward(sdG, body, inp, drop)
        optD.zero_grad()
        t = [ON.smooth_label(torch.rand(1)).to(dev) for _ in range(3)]
        lf = ON.gan_loss(ON.patchgan_forward(sdD, torch.cat((body, fakes), 1).detach()), t[0])
        lr = ON.gan_loss(ON.patchgan_forward(sdD, torch.cat((body, tgt), 1)), t[1])
        (0.5 * (lf + lr)).backward()
        optD.step()
        optG.zero_grad()
        ce = torch.nn.functional.cross_entropy(fakes, torch.argmax(tgt, 1)) * 100
        gan = ON.gan_loss(ON.patchgan_forward(sdD, torch.cat((body, fakes), 1)), t[2])
        (ce + gan).backward()
        optG.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 6
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    out["tf32_convs" if tf32 else "fp32_convs"] = {"ms_per_step": ms, "images_per_s": B / ms * 1e3}
    del sdG, sdD, optG, optD
    torch.cuda.empty_cache()
print(json.dumps({"what": "eager PyTorch (cuDNN) reference-equivalent warp step", "batch": B, "size": S, **out}))
