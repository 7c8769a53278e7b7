"""Runs the fused attention op alone (for ncu captures / timing)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from egotap_b200 import capi
Bf = int(sys.argv[1]) if len(sys.argv) > 1 else 64
prec = capi.PREC_BF16 if (len(sys.argv) > 2 and sys.argv[2] == "bf16") else capi.PREC_BF16X3
torch.manual_seed(0)
qk = torch.randn(Bf * 576, 2048, device="cuda"); vt = torch.randn(Bf * 8 * 128, 576, device="cuda")
qh, ql = capi.split_bf16(qk); vh, vl = capi.split_bf16(vt)
x3 = prec == capi.PREC_BF16X3
for _ in range(3):
    capi.attention(qh, ql if x3 else None, vh, vl if x3 else None, Bf, prec)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    capi.attention(qh, ql if x3 else None, vh, vl if x3 else None, Bf, prec)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
fl = 4.0 * Bf * 8 * 576 * 576 * 128
print("attention B=%d %s: %.3f ms  %.0f TFLOP/s algorithmic" % (Bf, "x3" if x3 else "bf16", ms, fl / ms / 1e9))
