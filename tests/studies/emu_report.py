"""Writes profiles/r01d_emulation_report.json: what the CPU emulation of the kernel source (tests/cuda_emu, DESIGN.md
section 11) reproduces -- the inference entry points against the oracle, the training step against autograd -- with every
launch executed from the product's own .cu files.  No GPU involved; numbers are parity figures, not timings.

    python tests/studies/emu_report.py
"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "cuda_emu")):
    sys.path.insert(0, p)

import torch  # noqa: E402

import build_emu  # noqa: E402
import egotap_oracle as orc  # noqa: E402
import train_oracle as tro  # noqa: E402
import weights  # noqa: E402
from egotap_b200 import training  # noqa: E402
from egotap_b200.synthetic import synthetic_heatmaps  # noqa: E402


def inference(preset, precision):
    lib = C.CDLL(build_emu.build())
    lib.egotap_b200_last_error.restype = C.c_char_p
    lib.egotap_b200_param_name.restype = C.c_char_p
    lib.egotap_b200_launch_count.restype = C.c_longlong
    lib.emu_set_num_sms(32)
    pid = 0 if preset == "UnrealEgo" else 1
    prec = 0 if precision == "bf16x3" else 1
    pb, wb = C.c_size_t(), C.c_size_t()
    lib.egotap_b200_plan_sizes.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    lib.egotap_b200_plan_sizes(pid, prec, 1, C.byref(pb), C.byref(wb))
    packed = torch.zeros(pb.value + 1024, dtype=torch.uint8)
    work = torch.full((wb.value // 4 + 256,), float("nan"))
    al = lambda t: (t.data_ptr() + 1023) // 1024 * 1024
    plan = C.c_void_p()
    lib.egotap_b200_plan_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    assert lib.egotap_b200_plan_create(pid, prec, 1, al(packed), al(work), C.byref(plan)) == 0
    sd = weights.make_state_dict(preset, seed=5)
    names = [lib.egotap_b200_param_name(pid, i).decode() for i in range(lib.egotap_b200_num_params(pid))]
    tens = [sd[n].float().contiguous() for n in names]
    arr = (C.c_void_p * len(tens))(*[t.data_ptr() for t in tens])
    lib.egotap_b200_pack_weights.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_void_p]
    assert lib.egotap_b200_pack_weights(plan, arr, len(tens), None) == 0
    x = synthetic_heatmaps(preset, 1, seed=1234, kind="gauss").contiguous()
    pose = torch.full((1, 16 if preset == "UnrealEgo" else 17, 3), float("nan"))
    lib.egotap_b200_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    n0 = lib.egotap_b200_launch_count()
    t0 = time.time()
    assert lib.egotap_b200_forward(plan, x.data_ptr(), 1, pose.data_ptr(), -1, None) == 0, lib.egotap_b200_last_error()
    dt = time.time() - t0
    with torch.no_grad():
        ref = orc.forward(sd, x, preset)
    rep = orc.parity_report(pose, ref)
    rep.update(launches=int(lib.egotap_b200_launch_count() - n0), emulation_seconds=round(dt, 1))
    lib.emu_set_num_sms(6)
    return rep


def train(preset, precision):
    emu, _ = build_emu.make_backend(real_tensor_core=True)
    sd = weights.make_state_dict(preset, seed=5)
    params = {k: v.clone().contiguous() for k, v in sd.items()}
    eng = training.TrainEngine(preset, params, precision=precision, backend=emu)
    eng.use_tape = False
    x = synthetic_heatmaps(preset, 1, seed=17, kind="gauss")
    nj = 16 if preset == "UnrealEgo" else 17
    gt = torch.randn(1, nj, 3, generator=torch.Generator().manual_seed(19)) * 20
    ref_loss, _, _, ref_grads = tro.train_step(sd, x, gt, preset)
    n0 = emu.launches
    t0 = time.time()
    eng.forward(x.clone())
    loss = eng.loss_and_grad(gt.clone())
    grads = eng.backward()
    dt = time.time() - t0
    worst_cos, worst_key = 1.0, None
    for k, g in ref_grads.items():
        if g is None or g.abs().max() < 1e-7:
            continue
        a, b = grads[k].flatten().double(), g.flatten().double()
        c = float((a @ b) / (a.norm() * b.norm()))
        if c < worst_cos:
            worst_cos, worst_key = c, k
    return dict(loss=float(loss[0]), ref_loss=float(ref_loss), loss_rel_diff=abs(float(loss[0]) - float(ref_loss)) / abs(float(ref_loss)),
                worst_gradient_cosine=worst_cos, worst_gradient=worst_key, launches=int(emu.launches - n0),
                emulation_seconds=round(dt, 1))


def main():
    out = dict(what="product kernel source executed on the CPU emulation (tests/cuda_emu), batch 1; parity only, no timings",
               inference={}, training={})
    for preset in ("UnrealEgo", "EgoCap"):
        for precision in ("bf16x3", "bf16"):
            out["inference"]["%s/%s" % (preset, precision)] = inference(preset, precision)
            print("inference", preset, precision, out["inference"]["%s/%s" % (preset, precision)], flush=True)
    for preset, precision in (("UnrealEgo", "bf16x3"), ("EgoCap", "bf16x3"), ("UnrealEgo", "bf16")):
        out["training"]["%s/%s" % (preset, precision)] = train(preset, precision)
        print("training", preset, precision, out["training"]["%s/%s" % (preset, precision)], flush=True)
    with open(os.path.join(ROOT, "profiles", "r01d_emulation_report.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
