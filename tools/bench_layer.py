"""Times the forward / input-gradient / weight-gradient plans of ONE conv layer in isolation (back-to-back
launches, CUDA events): python tools/bench_layer.py kind n cin cout h w [reps] [ksplit]."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from swapnet_b200 import lowering as L, ops
from swapnet_b200.layers import ConvLayer


def run(kind, n, cin, cout, h, w, reps=20):
    dev = torch.device("cuda:0")
    k = 3 if kind in ("conv3r", "conv3z") else 4
    wshape = (cin, cout, k, k) if kind == "convT4s2" else (cout, cin, k, k)
    wt = (torch.randn(*wshape) * 0.02).to(dev)
    ih, iw = (h + 2, w + 2) if kind == "conv3r" else (h, w)
    x = ops.Planes(n, ih, iw, L.padc(cin), dev, dual=True)
    x.hi.view(torch.float16).normal_(); x.twin.hi.normal_()
    layer = ConvLayer(kind, wt, None, x, nsplit=3, name="bench")
    oh, ow = L.out_hw(kind, h, w)
    y = torch.zeros(n, oh, ow, cout, device=dev)
    layer.bind_forward(y)
    dy = ops.Planes(n, oh, ow, L.padc(cout), dev, fmt=ops.FMT_BF16)
    dy.hi.normal_()
    dx = torch.zeros(n, ih, iw, (cin + 3) // 4 * 4, device=dev)[..., :cin]
    wg = torch.zeros_like(wt)
    layer.bind_backward(dy, dx, wg, None)
    layer.pack()
    macs = n * oh * ow * cin * cout * k * k if kind != "convT4s2" else n * h * w * cin * cout * k * k
    out = {}
    for name, plans in (("fwd", layer.fwd_plans), ("dgrad", layer.dgrad_plans), ("wgrad", [layer.wgrad_plan])):
        for _ in range(3):
            for p in plans:
                p.run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            for p in plans:
                p.run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out[name] = ms
        print(f"{kind} n={n} {cin}->{cout} {h}x{w} {name:6s} {ms * 1e3:9.1f} us  {2 * macs / ms / 1e9:7.1f} TFLOP/s", flush=True)
    return out


if __name__ == "__main__":
    a = sys.argv[1:]
    run(a[0], *[int(v) for v in a[1:6]], reps=int(a[6]) if len(a) > 6 else 20)
