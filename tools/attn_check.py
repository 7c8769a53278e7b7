"""Runs the fused attention op N times on fixed inputs and reports (i) whether every launch gives the same bits and (ii) a digest
of the context rows, so that two builds of the library (EGOTAP_B200_LIB) can be compared bit for bit (tools/gpu_job_r2*.sh)."""
import hashlib, os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from egotap_b200 import capi
Bf = int(sys.argv[1]) if len(sys.argv) > 1 else 64
prec = capi.PREC_BF16 if (len(sys.argv) > 2 and sys.argv[2] == "bf16") else capi.PREC_BF16X3
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
x3 = prec == capi.PREC_BF16X3
torch.manual_seed(0)
qk = torch.randn(Bf * 576, 2048, device="cuda"); vt = torch.randn(Bf * 8 * 128, 576, device="cuda")
qh, ql = capi.split_bf16(qk); vh, vl = capi.split_bf16(vt)
try:
    first, bad = None, 0
    for i in range(reps):
        out = capi.attention(qh, ql if x3 else None, vh, vl if x3 else None, Bf, prec)
        torch.cuda.synchronize()
        cur = [o.clone() for o in (out if isinstance(out, (tuple, list)) else [out]) if o is not None]
        if first is None:
            first = cur
        elif any(not torch.equal(a, b) for a, b in zip(first, cur)):
            bad += 1
    dig = hashlib.sha1(b"".join(t.cpu().view(torch.int16).numpy().tobytes() for t in first)).hexdigest()[:16]
    nan = any(bool(torch.isnan(t.float()).any()) for t in first)
    print("attn_check B=%d %s: %d launches, %d differ from the first, digest %s, nan %s" % (Bf, "x3" if x3 else "bf16", reps, bad, dig, nan))
except Exception:
    traceback.print_exc()
    print("attn_check B=%d %s: FAILED (%s)" % (Bf, "x3" if x3 else "bf16", str(sys.exc_info()[1]).replace("\n", " | ")[:300]))
