"""Generate tests/golden/*.npz from the UNMODIFIED reference module (run in the build container).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

For each preset and input distribution: weights = oracle/weights.make_state_dict(preset, seed=5)
loaded (strict) into the reference's EgoTAPAutoEncoder, inputs =
egotap_b200.synthetic.synthetic_heatmaps(preset, 2, seed=1234, kind).  Stored: the reference's
pose output and the propagation-network taps the reference itself exposes as attributes
(``skel_inputs``, ``skel_embed``; reference model/net_architecture.py:725,727).
Weights and inputs are regenerated from their seeds at test time (388 MB would not fit in git).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)

import ref_shim  # noqa: E402
import weights  # noqa: E402
from egotap_b200.synthetic import synthetic_heatmaps  # noqa: E402

WEIGHT_SEED, INPUT_SEED, BATCH = 5, 1234, 2


def main():
    for preset in ("UnrealEgo", "EgoCap"):
        net = ref_shim.build_reference_net(preset)
        net.load_state_dict(weights.make_state_dict(preset, seed=WEIGHT_SEED), strict=True)
        for kind in ("gauss", "uniform"):
            x = synthetic_heatmaps(preset, BATCH, seed=INPUT_SEED, kind=kind)
            with torch.no_grad():
                pose, rot, indep, hm = net(x)
            out = dict(pose=pose.detach().numpy(),
                       skel_inputs=net.skel_inputs.detach().numpy(),      # (J, B, 512) = [pos|limb] embeds
                       skel_embed=net.skel_embed.detach().numpy(),        # (J, B, 512)
                       aux_shapes=np.array([rot.shape[1], indep.shape[1], hm.shape[1]]),
                       aux_absmax=np.array([rot.abs().max().item(), indep.abs().max().item(), hm.abs().max().item()]),
                       input_checksum=np.array([x.double().sum().item(), x.double().pow(2).sum().item()]),
                       meta=np.array([WEIGHT_SEED, INPUT_SEED, BATCH]))
            path = os.path.join(HERE, "ref_%s_%s.npz" % (preset, kind))
            np.savez_compressed(path, **out)
            print("wrote", path, pose.shape)


def metrics_golden():
    """Evaluation-metric golden: the reference's own Procrustes + MPJPE on seeded poses."""
    ref_shim.import_reference()
    from utils.loss import LossFuncMPJPE
    from utils.util import batch_compute_similarity_transform_torch
    g = torch.Generator().manual_seed(3)
    gt = torch.randn(6, 16, 3, generator=g) * 20
    pred = gt + torch.randn(6, 16, 3, generator=g) * 3
    pred[1] = gt[1] @ torch.tensor([[0., 1, 0], [1, 0, 0], [0, 0, 1]])      # a mirrored pose (det < 0 branch)
    S = batch_compute_similarity_transform_torch(pred, gt)
    lf = LossFuncMPJPE()
    m = torch.stack([lf(pred[i], gt[i]) * 10 for i in range(6)])
    pa = torch.stack([lf(S[i], gt[i]) * 10 for i in range(6)])
    np.savez_compressed(os.path.join(HERE, "ref_metrics.npz"), pred=pred.numpy(), gt=gt.numpy(), mpjpe_mm=m.numpy(),
                        pa_mpjpe_mm=pa.numpy())
    print("wrote ref_metrics.npz")


def train_step_golden():
    """Training-step golden: loss and parameter updates of ONE reference optimisation step
    (reference module in train mode + its loss classes + torch.optim.AdamW), probes of six tensors."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_train_oracle import PROBE, reference_train_step
    ref_loss, ref_sd = reference_train_step("UnrealEgo")
    sd = weights.make_state_dict("UnrealEgo", seed=WEIGHT_SEED)
    out = dict(loss=np.float64(ref_loss))
    for i, k in enumerate(PROBE):
        out["upd_%d" % i] = (ref_sd[k] - sd[k]).flatten()[:4096].numpy()
    np.savez_compressed(os.path.join(HERE, "ref_train_step.npz"), **out)
    print("wrote ref_train_step.npz")


if __name__ == "__main__":
    main()
    metrics_golden()
    train_step_golden()
