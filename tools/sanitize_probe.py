"""Small forward (both precisions, both presets) and one training step for compute-sanitizer runs (memcheck / initcheck / synccheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch
import egotap_b200, weights
from egotap_b200.options import make_opt
what = sys.argv[1] if len(sys.argv) > 1 else "infer"
if what == "infer":
    for preset in ("UnrealEgo", "EgoCap"):
        sd = weights.make_state_dict(preset, 5)
        for prec in ("bf16x3", "bf16"):
            net = egotap_b200.EgoTAPAutoEncoder(make_opt(preset, b200_precision=prec), 2); net.load_state_dict(sd)
            net = net.cuda().eval()
            for B in (3, 2):
                p = net.predict_pose(egotap_b200.synthetic_heatmaps(preset, B, seed=B).cuda())
                torch.cuda.synchronize()
                print(preset, prec, B, bool(torch.isfinite(p).all()))
else:
    from egotap_b200.training import TrainEngine
    preset = "UnrealEgo"
    sd = weights.make_state_dict(preset, 5)
    eng = TrainEngine(preset, {k: v.clone().cuda().contiguous() for k, v in sd.items()}, precision="bf16")
    x = egotap_b200.synthetic_heatmaps(preset, 2, seed=1).cuda()
    gt = torch.randn(2, 16, 3, device="cuda")
    print("loss", float(eng.train_step(x, gt)))
