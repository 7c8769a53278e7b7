"""Debug: wait-cycle accounting of the attention kernel's warp roles (library built with -DEB_ATTN_TRACE)."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lib = os.path.join(ROOT, "tools", "libtrace.so")
import glob
_csrc = glob.glob(os.path.join(ROOT, "egotap_b200", "csrc", "*"))
if not os.path.isfile(lib) or "--build" in sys.argv or any(os.path.getmtime(f) > os.path.getmtime(lib) for f in _csrc):
    os.makedirs(os.path.dirname(lib), exist_ok=True)
    src = [os.path.join(ROOT, "egotap_b200", "csrc", f) for f in ("api.cu", "kernels.cu", "gemm_launch.cu", "attention.cu", "attention_bwd.cu", "pu_chain.cu", "pu_chain_bwd.cu",
                                                                       "metrics.cu", "plan.cu", "train_ops.cu", "train_model.cu", "gt_heatmap.cu")]
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--shared", "-Xcompiler", "-fPIC",
                           "-DEB_ATTN_TRACE", "-o", lib] + src)
if "--build" in sys.argv:
    sys.exit(0)
import torch
from egotap_b200 import capi
capi.LIB_PATH = lib
L = capi.lib()
trace = L.egotap_b200_attn_trace
Bf = int(os.environ.get("ATTN_TRACE_FRAMES", "256"))
for prec, name in ((capi.PREC_BF16X3, "x3"), (capi.PREC_BF16, "bf16")):
    x3 = prec == capi.PREC_BF16X3
    qk = torch.randn(Bf * 576, 2048, device="cuda"); vt = torch.randn(Bf * 8 * 128, 576, device="cuda")
    qh, ql = capi.split_bf16(qk); vh, vl = capi.split_bf16(vt)
    capi.attention(qh, ql if x3 else None, vh, vl if x3 else None, Bf, prec); torch.cuda.synchronize()
    out = (C.c_ulonglong * 32)()
    trace(out, 1)
    capi.attention(qh, ql if x3 else None, vh, vl if x3 else None, Bf, prec); torch.cuda.synchronize()
    trace(out, 1)
    v = [x / 148.0 for x in out]
    items = Bf * 8 * 5 / 148.0
    print("== attention_kernel %s (cycles per CTA, %.1f items = %.1f key tiles per CTA)" % (name, items, 5 * items))
    # warp 1 issues the even tiles (half of them): its busy time per issued tile = (total - waits) / (tiles / 2)
    waits = v[1] + v[2] + v[3] + v[4] + v[5]
    print(" MMA warp 1 : total %.0f | wait q_full %.0f  kv_full %.0f  pv_done %.0f  p_half/p_full %.0f  o_empty %.0f | period per tile %.0f, busy per own tile %.0f"
          % (v[0], v[1], v[2], v[3], v[4], v[5], v[0] / items / 5, (v[0] - waits) / (items * 2.5)))
    print(" softmax (warp 2) : total %.0f | wait s_full %.0f  pair_sync %.0f  pv_done(rescale) %.0f" % (v[8], v[9], v[10], v[11]))
    print(" producer : total %.0f | wait q_empty %.0f  kv_empty %.0f" % (v[16], v[17], v[18]))
    print(" epilogue : total %.0f | wait o_full %.0f" % (v[24], v[25]))
