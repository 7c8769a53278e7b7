"""CPU study of the training step's operand precision (no GPU needed): the SAME host orchestration
(egotap_b200/training.py) driven through the op oracle in three operand-storage modes --
  exact   : fp32 operands (what torch.autograd on the reference computes, up to summation order)
  bf16x3  : bf16 hi/lo pairs, 3 MMAs per k-step (the fp32-parity mode)
  bf16    : plain bf16 operands, fp32 accumulation, fp32 master weights (BASELINE config 5's mode)
-- for a number of AdamW steps on a fixed synthetic batch, recording the loss trajectory and the distance of the weights
from the exact run.  Writes profiles/r01d_train_precision_cpu.json.

    python tests/studies/train_precision_study.py [--steps 12] [--batch 4]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import torch  # noqa: E402

import op_oracle  # noqa: E402
import weights  # noqa: E402
from egotap_b200 import training  # noqa: E402
from egotap_b200.synthetic import synthetic_heatmaps  # noqa: E402


def run(mode, preset, batch, steps, lr):
    sd = weights.make_state_dict(preset, seed=5)
    params = {k: v.clone().contiguous() for k, v in sd.items()}
    eng = training.TrainEngine(preset, params, precision="bf16" if mode == "bf16" else "bf16x3",
                               backend=op_oracle.OracleBackend(exact=(mode == "exact")))
    x = synthetic_heatmaps(preset, batch, seed=17, kind="gauss")
    nj = 16 if preset == "UnrealEgo" else 17
    gt = torch.randn(batch, nj, 3, generator=torch.Generator().manual_seed(19)) * 20
    losses = []
    for s in range(steps):
        losses.append(float(eng.train_step(x, gt, lr=lr)[0]))
    return losses, params


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--preset", default="UnrealEgo")
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r01d_train_precision_cpu.json"))
    args = ap.parse_args()
    out = dict(what="loss trajectory of TrainEngine.train_step on the CPU op oracle, fixed batch", preset=args.preset,
               batch=args.batch, steps=args.steps, lr=args.lr, modes={})
    ref_params = None
    for mode in ("exact", "bf16x3", "bf16"):
        t0 = time.time()
        losses, params = run(mode, args.preset, args.batch, args.steps, args.lr)
        rec = dict(loss=losses, seconds=round(time.time() - t0, 1))
        if mode == "exact":
            ref_params = params
        else:
            num = sum(float((params[k] - ref_params[k]).double().pow(2).sum()) for k in params if params[k].is_floating_point())
            den = sum(float((ref_params[k] - weights.make_state_dict(args.preset, seed=5)[k]).double().pow(2).sum())
                      for k in params if params[k].is_floating_point())
            rec["weight_drift_vs_exact_rel_to_total_update"] = (num / max(den, 1e-300)) ** 0.5
        out["modes"][mode] = rec
        print(mode, ["%.4f" % l for l in losses], rec.get("weight_drift_vs_exact_rel_to_total_update"))
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
