"""One warm-up + N forward steps of the lifting path (for ncu launch lists / captures; never a bench value)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import egotap_b200  # noqa: E402
from egotap_b200.options import make_opt  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--precision", default="bf16x3")
ap.add_argument("--preset", default="UnrealEgo")
a = ap.parse_args()
torch.manual_seed(0)
net = egotap_b200.EgoTAPAutoEncoder(make_opt(a.preset, b200_precision=a.precision, b200_max_batch=a.batch), input_channel_scale=2)
net.init_weights("kaiming")
net = net.cuda().eval()
x = egotap_b200.synthetic_heatmaps(a.preset, a.batch, seed=1234).cuda()
net.predict_pose(x)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("timed")
for _ in range(a.steps):
    p = net.predict_pose(x)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("ok", tuple(p.shape), float(p.abs().max()))
