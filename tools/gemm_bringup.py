"""GPU bring-up harness for the tcgen05 GEMM: every tile configuration in its own subprocess
(a protocol bug then costs one timeout, not the whole call).  Writes gpurun_out/gemm_bringup.json.

    python tools/gemm_bringup.py            # all variants
    python tools/gemm_bringup.py --one V    # (internal) run variant V in this process
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_one(v):
    import torch
    from egotap_b200 import capi
    L = capi.lib()
    name = L.egotap_b200_gemm_variant_name(v).decode()
    x3 = "x3" in name
    prec = capi.PREC_BF16X3 if x3 else capi.PREC_BF16
    out = {"variant": v, "name": name, "cases": []}
    torch.manual_seed(0)
    dev = "cuda"

    def case(M, N, K, groups=1, tag=""):
        A = torch.randn(groups, M, K, device=dev)
        B = torch.randn(groups, N, K, device=dev)
        ah, al = capi.split_bf16(A)
        bh, bl = capi.split_bf16(B)
        D = torch.full((groups, M, N), float("nan"), device=dev)
        capi.gemm(ah, al if x3 else None, bh, bl if x3 else None, M, N, K, groups=groups,
                  a_group=(groups, M * K, 1, 0), b_group=(groups, N * K, 1, 0),
                  precision=prec, variant=v, out_f32=D, ldo=N, group_rows=M)
        torch.cuda.synchronize()
        if x3:
            ref = torch.matmul(A.double(), B.double().transpose(1, 2))
        else:
            ref = torch.matmul(ah.double(), bh.double().transpose(1, 2))
        err = (D.double() - ref).abs().max().item()
        scale = ref.abs().max().item()
        nan = int(torch.isnan(D).sum().item())
        out["cases"].append(dict(tag=tag, M=M, N=N, K=K, groups=groups, max_err=err, ref_max=scale, rel=err / scale, nan=nan))

    case(128, 256, 64, tag="one tile one kblock")
    case(128, 256, 256, tag="one tile 4 kblocks")
    case(512, 512, 512, tag="multi tile")
    case(300, 768, 128, tag="M tail")
    case(576, 576, 128, groups=5, tag="grouped, N tail")
    case(1000, 1024, 1024, tag="ring wrap")
    # timing on a layer-sized problem (QKV projection at batch 64)
    M, N, K = 64 * 576, 3072, 1024
    A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev)
    ah, al = capi.split_bf16(A); bh, bl = capi.split_bf16(B)
    oh = torch.empty(M, N, device=dev, dtype=torch.bfloat16); ol = torch.empty_like(oh)
    bias = torch.randn(N, device=dev)

    def go():
        capi.gemm(ah, al if x3 else None, bh, bl if x3 else None, M, N, K, precision=prec, variant=v,
                  bias=bias, out_hi=oh, out_lo=ol if x3 else None, ldo=N)
    for _ in range(3):
        go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        go()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    out["timing"] = dict(M=M, N=N, K=K, ms=ms, tflops_algorithmic=2.0 * M * N * K / ms / 1e9,
                         mma_tflops=2.0 * M * N * K * (3 if x3 else 1) / ms / 1e9)
    ref = torch.matmul(A, B.t()) + bias if x3 else torch.matmul(ah.float(), bh.float().t()) + bias
    got = oh.float() + (ol.float() if x3 else 0)
    out["timing"]["rel_err"] = ((got - ref).abs().max() / ref.abs().max()).item()
    print(json.dumps(out))


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        run_one(int(sys.argv[2]))
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    sys.path.insert(0, ROOT)
    from egotap_b200 import capi
    n = capi.lib().egotap_b200_gemm_num_variants()
    results = []
    for v in range(n):
        t = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", str(v)], capture_output=True,
                               text=True, timeout=120)
            line = [l for l in p.stdout.splitlines() if l.startswith("{")]
            r = json.loads(line[-1]) if line else {"variant": v, "error": (p.stdout + p.stderr)[-1500:], "rc": p.returncode}
        except subprocess.TimeoutExpired as e:
            r = {"variant": v, "error": "timeout", "tail": ((e.stdout or b"")[-500:]).decode(errors="replace") if isinstance(e.stdout, bytes) else str(e.stdout)[-500:]}
        r["wall_s"] = time.time() - t
        results.append(r)
        print(json.dumps(r)[:1200], flush=True)
    with open(os.path.join(ROOT, "gpurun_out", "gemm_bringup.json"), "w") as f:
        json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
