"""Board power and SM clock while ONE kernel family runs in a loop for a couple of seconds (nvidia-smi, 100 ms samples):
the energy side of the step.  python tools/power_probe.py [seconds]"""
import os, subprocess, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from egotap_b200 import capi

SECS = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0


class Sampler:
    def __enter__(self):
        self.rows = []
        self.p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=power.draw,clocks.sm", "--format=csv,noheader,nounits", "-lms", "100"],
                                  stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        self.t = threading.Thread(target=lambda: [self.rows.append(l.split(",")) for l in self.p.stdout], daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.p.terminate(); self.t.join(timeout=2)

    def med(self, i):
        v = []
        for r in self.rows[3:]:                 # skip the ramp
            try:
                v.append(float(r[i]))
            except (IndexError, ValueError):
                pass
        v.sort()
        return v[len(v) // 2] if v else float("nan")


def probe(name, fn, work=None, unit=""):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    with Sampler() as s:
        t0 = time.time()
        e0.record()
        while time.time() - t0 < SECS:
            for _ in range(10):
                fn()
            n += 10
            torch.cuda.synchronize()
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    pw, clk = s.med(0), s.med(1)
    extra = "" if work is None else "  %.0f %s" % (work / ms / 1e9, unit)
    print("%-44s %8.3f ms/launch  %7.1f W  %5.0f MHz  %7.2f J/launch%s" % (name, ms, pw, clk, pw * ms * 1e-3, extra))


torch.manual_seed(0)
M = 147456
for prec, pname in ((capi.PREC_BF16X3, "bf16x3"), (capi.PREC_BF16, "bf16")):
    x3 = prec == capi.PREC_BF16X3
    for (N, K, tag) in ((3072, 1024, "QKV"), (1024, 4096, "MLP-down")):
        A, B = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda") * 0.05
        ah, al = capi.split_bf16(A); bh, bl = capi.split_bf16(B)
        if not x3:
            al = bl = None
        oh = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        ol = torch.empty_like(oh) if x3 else None
        probe("gemm %s %s (%dx%dx%d)" % (tag, pname, M, N, K),
              lambda: capi.gemm(ah, al, bh, bl, M, N, K, precision=prec, out_hi=oh, out_lo=ol), 2.0 * M * N * K, "TFLOP/s")
        del A, B, ah, al, bh, bl, oh, ol
    qk = torch.randn(256 * 576, 2048, device="cuda"); vt = torch.randn(256 * 8 * 128, 576, device="cuda")
    qh, ql = capi.split_bf16(qk); vh, vl = capi.split_bf16(vt)
    probe("attention %s (256 frames)" % pname, lambda: capi.attention(qh, ql if x3 else None, vh, vl if x3 else None, 256, prec),
          4.0 * 256 * 8 * 576 * 576 * 128, "TFLOP/s")
    del qk, vt, qh, ql, vh, vl
h = torch.randn(M, 1024, device="cuda"); w = torch.ones(1024, device="cuda"); b = torch.zeros(1024, device="cuda")
probe("layernorm (147456 rows, hi + lo + fp32 out)", lambda: capi.layernorm(h, w, b, 256, 576, 576), M * 1024 * 16.0, "GB/s")
