"""One warp training step between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_batch, warp_opt
from swapnet_b200.models import create_model

B = int(os.environ.get("SN_B", 16)); S = int(os.environ.get("SN_S", 512))
torch.manual_seed(0)
model = create_model(warp_opt(B, S, os.environ.get("SN_PREC", "fp32x3")))
model.setup(model.opt)
batch = synth_batch(B, S, 1234)
for k in ("bodys", "input_cloths", "target_cloths"):
    batch[k] = batch[k].cuda()
for _ in range(2):
    model.set_input(batch); model.optimize_parameters()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
model.set_input(batch); model.optimize_parameters()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("losses", model.get_current_losses())
