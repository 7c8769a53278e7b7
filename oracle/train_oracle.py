"""TEST INFRASTRUCTURE ONLY -- CPU oracle of the pose-estimator TRAINING STEP (SURVEY.md section 8(f) row f2,
BASELINE config 5): the checker of the CUDA training path (egotap_b200/training.py, csrc/train_*.cu).

Restates, in fp32/fp64 torch on the CPU:
  * the train-mode forward: identical to the eval forward except that the six FC blocks use BatchNorm1d batch
    statistics over the B*2J rows and update running_mean / running_var (momentum 0.1, unbiased variance) and
    num_batches_tracked (reference model/network_utils.py:123-142; torch.nn.BatchNorm1d semantics)
  * the loss  lambda_mpjpe * MPJPE + lambda_cos_sim * lambda_mpjpe * CosSim(bones)
    (reference model/egotap_autoencoder_model.py:284-296, utils/loss.py:44-85; weights from
    scripts/train/PoseEstimator/*.sh: lambda_mpjpe 0.1, lambda_cos_sim -0.01)
  * one AdamW step (lr 1e-3, betas (0.9, 0.999), eps 1e-4, weight decay 0; reference model/network.py:72-78,
    options/train_options.py:31-37) over every parameter of the lifting net
  * the cosine schedule with linear warm-up (transformers.get_cosine_schedule_with_warmup, model/network.py:49-52)
Gradients come from torch.autograd on the restated forward.  Pinned against the unmodified reference module +
its own loss classes + torch.optim.AdamW in tests/test_train_oracle.py (live) and tests/golden/ref_train_step.npz.
"""
import math

import torch
import torch.nn.functional as F

import egotap_oracle as orc

KINEMATIC_PARENTS = {   # reference utils/util.py:51-52
    "UnrealEgo": [0, 0, 1, 1, 2, 3, 4, 5, 2, 3, 8, 9, 10, 11, 12, 13],
    "EgoCap": [0, 0, 1, 2, 3, 4, 1, 6, 7, 8, 2, 10, 11, 12, 6, 14, 15, 16],
}
BN_MOMENTUM = 0.1


def fc_block_train(sd, prefix, x, new_stats):
    """Linear -> BatchNorm1d(train: batch statistics) -> LeakyReLU(0.2); records the running-stat update."""
    y = F.linear(x, sd[prefix + ".fc.weight"], sd[prefix + ".fc.bias"])
    mu = y.mean(0)
    var_b = y.var(0, unbiased=False)
    n = y.shape[0]
    with torch.no_grad():
        new_stats[prefix + ".bn.running_mean"] = (1 - BN_MOMENTUM) * sd[prefix + ".bn.running_mean"] + BN_MOMENTUM * mu
        new_stats[prefix + ".bn.running_var"] = ((1 - BN_MOMENTUM) * sd[prefix + ".bn.running_var"]
                                                 + BN_MOMENTUM * var_b * n / max(n - 1, 1))
        new_stats[prefix + ".bn.num_batches_tracked"] = sd[prefix + ".bn.num_batches_tracked"] + 1
    y = (y - mu) / torch.sqrt(var_b + orc.BN_EPS) * sd[prefix + ".bn.weight"] + sd[prefix + ".bn.bias"]
    return F.leaky_relu(y, orc.LEAKY)


def forward_train(sd, x, preset):
    """Train-mode forward.  Returns (pose, new_stats) where new_stats holds the updated BatchNorm buffers."""
    new_stats = {}
    saved = orc.fc_block
    orc.fc_block = lambda sd_, prefix, x_: fc_block_train(sd_, prefix, x_, new_stats)
    try:
        pose = orc.forward(sd, x, preset)
    finally:
        orc.fc_block = saved
    return pose, new_stats


def loss_mpjpe(pred, gt):
    """reference utils/loss.py:79-85"""
    return torch.linalg.norm(gt - pred, dim=-1).mean()


def loss_cos_sim(pred, gt, preset):
    """reference utils/loss.py:44-77 (pred_rot=False).  EgoCap (estimate_head False) prepends a zero root joint and
    drops the first bone."""
    parents = KINEMATIC_PARENTS[preset]
    estimate_head = preset == "UnrealEgo"
    if not estimate_head:
        z = pred.new_zeros(pred.shape[0], 1, 3)
        pred, gt = torch.cat([z, pred], 1), torch.cat([z, gt], 1)
    pb = (pred - pred[:, parents])[:, 1:]
    gb = (gt - gt[:, parents])[:, 1:]
    cos = F.cosine_similarity(pb, gb, dim=2)
    if not estimate_head:
        cos = cos[:, 1:]
    return cos.sum(1).mean(0)


def total_loss(pred, gt, preset, lambda_mpjpe=0.1, lambda_cos_sim=-0.01):
    """reference model/egotap_autoencoder_model.py:284-296"""
    return loss_mpjpe(pred, gt) * lambda_mpjpe + loss_cos_sim(pred, gt, preset) * lambda_cos_sim * lambda_mpjpe


def cosine_warmup_lr(step, base_lr, warmup_steps, total_steps):
    """transformers.get_cosine_schedule_with_warmup (num_cycles 0.5), as used by model/network.py:49-52"""
    if step < warmup_steps:
        return base_lr * step / max(1, warmup_steps)
    prog = (step - warmup_steps) / max(1, total_steps - warmup_steps)
    return base_lr * max(0.0, 0.5 * (1.0 + math.cos(math.pi * prog)))


def adamw_update(p, g, m, v, step, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-4, weight_decay=0.0):
    """torch.optim.AdamW, single tensor, step counted from 1.  Returns (p, m, v)."""
    p = p * (1 - lr * weight_decay)
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    return p - (lr / bc1) * m / denom, m, v


def train_step(sd, x, gt, preset, opt_state=None, lr=1e-3, eps=1e-4, weight_decay=0.0):
    """One optimisation step.  sd: state_dict (fp32 or fp64); returns (loss, new_sd, opt_state, grads)."""
    keys = [k for k, v in sd.items() if v.is_floating_point() and "running_" not in k]
    work = dict(sd)
    leaves = {k: sd[k].detach().clone().requires_grad_(True) for k in keys}
    work.update(leaves)
    pose, new_stats = forward_train(work, x, preset)
    loss = total_loss(pose, gt.to(pose.dtype), preset)
    used = [k for k in keys if "cls_token" not in k and "pooler" not in k]          # dead parameters get no gradient
    grads = dict(zip(used, torch.autograd.grad(loss, [leaves[k] for k in used], allow_unused=True)))
    if opt_state is None:
        opt_state = dict(step=0, m={}, v={})
    step = opt_state["step"] + 1
    new_sd = dict(sd)
    new_sd.update(new_stats)
    for k in used:
        g = grads[k]
        if g is None:
            continue
        m = opt_state["m"].get(k, torch.zeros_like(sd[k]))
        v = opt_state["v"].get(k, torch.zeros_like(sd[k]))
        new_sd[k], opt_state["m"][k], opt_state["v"][k] = adamw_update(sd[k], g, m, v, step, lr, eps=eps,
                                                                       weight_decay=weight_decay)
    opt_state["step"] = step
    return loss.detach(), new_sd, opt_state, grads
