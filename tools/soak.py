"""Soak: the whole forward (both precisions, several batch sizes, UnrealEgo and EgoCap) repeated many times -- every repetition must give
the same bits (no atomics anywhere), so a rare protocol race in a warp-specialised kernel shows as a differing repetition."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch
import egotap_b200, weights
from egotap_b200.options import make_opt
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
bad_total = 0
for preset in ("UnrealEgo", "EgoCap"):
    sd = weights.make_state_dict(preset, 5)
    for prec in ("bf16x3", "bf16"):
        net = egotap_b200.EgoTAPAutoEncoder(make_opt(preset, b200_precision=prec), 2); net.load_state_dict(sd)
        net = net.cuda().eval()
        for B in (3, 16, 96, 256):
            x = egotap_b200.synthetic_heatmaps(preset, B, seed=B).cuda()
            first = net.predict_pose(x).clone()
            bad = 0
            for _ in range(reps):
                if not torch.equal(net.predict_pose(x), first):
                    bad += 1
            torch.cuda.synchronize()
            bad_total += bad
            print("%s %s batch %d: %d repetitions, %d differ, finite %s" % (preset, prec, B, reps, bad, bool(torch.isfinite(first).all())))
print("SOAK", "FAILED" if bad_total else "OK")
