"""ctypes binding of ``libegotap_b200.so`` (C ABI declared in ``include/egotap_b200.h``).

This is the only place Python touches the native library.  There is no fallback: if the shared
library is missing the import raises, and every compute entry needs a CUDA device.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libegotap_b200.so")

PREC_BF16X3, PREC_BF16 = 0, 1
PRESET_ID = {"UnrealEgo": 0, "EgoCap": 1}
ACT_NONE, ACT_GELU, ACT_LRELU = 0, 1, 2
STORE_ROWMAJOR, STORE_QKV, STORE_JOINT_REGROUP, STORE_HEAD_MERGE = 0, 1, 2, 3

# every symbol include/egotap_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "egotap_b200_abi_version", "egotap_b200_last_error", "egotap_b200_launch_count",
    "egotap_b200_gemm", "egotap_b200_gemm_num_variants", "egotap_b200_gemm_variant_name",
    "egotap_b200_split_bf16", "egotap_b200_attention", "egotap_b200_ingest", "egotap_b200_layernorm",
    "egotap_b200_pu_permute_split", "egotap_b200_pu_chain", "egotap_b200_head", "egotap_b200_pose_metrics", "egotap_b200_profile_begin", "egotap_b200_profile_end", "egotap_b200_profile_record",
    "egotap_b200_num_params", "egotap_b200_param_name", "egotap_b200_plan_sizes", "egotap_b200_plan_create",
    "egotap_b200_plan_destroy", "egotap_b200_pack_weights", "egotap_b200_forward", "egotap_b200_plan_buffer",
]


class Operand(C.Structure):
    _fields_ = [("hi", C.c_void_p), ("lo", C.c_void_p), ("ld", C.c_longlong), ("rows", C.c_longlong),
                ("g0_count", C.c_longlong), ("g0_stride", C.c_longlong),
                ("g1_count", C.c_longlong), ("g1_stride", C.c_longlong)]


class Epilogue(C.Structure):
    _fields_ = [("alpha", C.c_float), ("scale", C.c_void_p), ("bias", C.c_void_p), ("act", C.c_int),
                ("resid", C.c_void_p), ("resid_ld", C.c_longlong), ("resid_mod", C.c_int),
                ("rows_in", C.c_int), ("rows_out", C.c_int), ("group_rows", C.c_longlong),
                ("out_f32", C.c_void_p), ("out_hi", C.c_void_p), ("out_lo", C.c_void_p),
                ("ldo", C.c_longlong), ("col_off", C.c_int), ("store", C.c_int),
                ("qk_cols", C.c_int), ("tokens", C.c_int), ("vt_hi", C.c_void_p), ("vt_lo", C.c_void_p),
                ("J", C.c_int), ("heads", C.c_int)]


class Gemm(C.Structure):
    _fields_ = [("a", Operand), ("b", Operand), ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
                ("groups", C.c_int), ("precision", C.c_int), ("variant", C.c_int), ("epi", Epilogue)]


_lib = None


def lib():
    """Load (once) and return the native library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError("egotap_b200: native library %s is missing -- run "
                               "`python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback exists)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.egotap_b200_last_error.restype = C.c_char_p
        L.egotap_b200_launch_count.restype = C.c_longlong
        L.egotap_b200_gemm_variant_name.restype = C.c_char_p
        L.egotap_b200_gemm.argtypes = [C.POINTER(Gemm), C.c_void_p]
        L.egotap_b200_split_bf16.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]
        L.egotap_b200_attention.argtypes = [C.c_void_p] * 6 + [C.c_int, C.c_int, C.c_void_p]
        L.egotap_b200_ingest.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 5
        L.egotap_b200_layernorm.argtypes = [C.c_void_p] * 3 + [C.c_longlong, C.c_int, C.c_int, C.c_float] + [C.c_void_p] * 4
        L.egotap_b200_pu_permute_split.argtypes = [C.c_void_p] * 4
        L.egotap_b200_pu_chain.argtypes = ([C.c_void_p] * 3 + [C.c_longlong] * 2 + [C.c_void_p] + [C.c_longlong] * 2 +
                                           [C.c_void_p] * 6 + [C.c_int] * 3 + [C.c_void_p])
        L.egotap_b200_head.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_longlong, C.c_int, C.c_void_p, C.c_void_p]
        L.egotap_b200_pose_metrics.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_float, C.c_void_p,
                                               C.c_void_p, C.c_void_p]
        L.egotap_b200_param_name.restype = C.c_char_p
        L.egotap_b200_plan_sizes.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.egotap_b200_plan_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        L.egotap_b200_plan_destroy.argtypes = [C.c_void_p]
        L.egotap_b200_pack_weights.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_void_p]
        L.egotap_b200_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.egotap_b200_plan_buffer.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        raise RuntimeError("egotap_b200: %s failed (%d): %s" % (what, rc, lib().egotap_b200_last_error().decode()))


def _ptr(t):
    return None if t is None else t.data_ptr()


def current_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("egotap_b200 has no CPU path: got a %s tensor" % t.device)


def split_bf16(x, want_lo=True):
    """fp32 CUDA tensor -> (hi, lo) bf16 tensors with x ~= hi + lo."""
    import torch
    require_cuda(x)
    x = x.contiguous()
    hi = torch.empty_like(x, dtype=torch.bfloat16)
    lo = torch.empty_like(x, dtype=torch.bfloat16) if want_lo else None
    check(lib().egotap_b200_split_bf16(x.data_ptr(), hi.data_ptr(), _ptr(lo), x.numel(), current_stream()), "split_bf16")
    return hi, lo


def attention(qk_hi, qk_lo, vt_hi, vt_lo, frames, precision=PREC_BF16X3):
    """Fused attention op: returns (ctx_hi, ctx_lo) bf16 tensors of shape (frames*576, 1024)."""
    import torch
    require_cuda(qk_hi, qk_lo, vt_hi, vt_lo)
    ctx_hi = torch.empty((frames * 576, 1024), dtype=torch.bfloat16, device=qk_hi.device)
    ctx_lo = torch.empty_like(ctx_hi) if precision == PREC_BF16X3 else None
    check(lib().egotap_b200_attention(_ptr(qk_hi), _ptr(qk_lo), _ptr(vt_hi), _ptr(vt_lo), _ptr(ctx_hi), _ptr(ctx_lo),
                                      frames, precision, current_stream()), "attention")
    return ctx_hi, ctx_lo


def ingest(x, preset):
    """(B, 6J, 64, 64) fp32 -> (patch_hi, patch_lo, limb_hi, limb_lo) bf16 operand matrices."""
    import torch
    require_cuda(x)
    B, J = x.shape[0], x.shape[1] // 6
    mk = lambda r, c: torch.empty((r, c), dtype=torch.bfloat16, device=x.device)
    ph, pl, lh, ll = mk(B * 2 * J * 16, 256), mk(B * 2 * J * 16, 256), mk(B * 2 * J, 8192), mk(B * 2 * J, 8192)
    check(lib().egotap_b200_ingest(x.data_ptr(), B, PRESET_ID[preset], ph.data_ptr(), pl.data_ptr(), lh.data_ptr(),
                                   ll.data_ptr(), current_stream()), "ingest")
    return ph, pl, lh, ll


def layernorm(x, w, b, frames, rows_in, rows_out, eps=1e-12):
    """fp32 (frames*rows_in, 1024) -> (hi, lo, f32) each (frames*rows_out, 1024)."""
    import torch
    require_cuda(x, w, b)
    hi = torch.empty((frames * rows_out, 1024), dtype=torch.bfloat16, device=x.device)
    lo = torch.empty_like(hi)
    f32 = torch.empty((frames * rows_out, 1024), dtype=torch.float32, device=x.device)
    check(lib().egotap_b200_layernorm(x.data_ptr(), w.data_ptr(), b.data_ptr(), frames, rows_in, rows_out, eps,
                                      hi.data_ptr(), lo.data_ptr(), f32.data_ptr(), current_stream()), "layernorm")
    return hi, lo, f32


def pu_chain(w_hh, G, F, frames, J, precision=PREC_BF16X3):
    """One propagation-unit layer: w_hh (2048, 512) fp32, G (frames*J, 2048) fp32, F (frames*J, 512) fp32
    -> h (frames*J, 512) fp32."""
    import torch
    require_cuda(w_hh, G, F)
    dev = w_hh.device
    wh = torch.empty((2048, 512), dtype=torch.bfloat16, device=dev); wl = torch.empty_like(wh)
    check(lib().egotap_b200_pu_permute_split(w_hh.contiguous().data_ptr(), wh.data_ptr(), wl.data_ptr(), current_stream()),
          "pu_permute_split")
    out = torch.empty((frames * J, 512), dtype=torch.float32, device=dev)
    hg_h = torch.zeros((2 * frames, 512), dtype=torch.bfloat16, device=dev); hg_l = torch.zeros_like(hg_h)
    cnt = torch.zeros(64, dtype=torch.int32, device=dev)
    x3 = precision == PREC_BF16X3
    check(lib().egotap_b200_pu_chain(wh.data_ptr(), wl.data_ptr() if x3 else None, G.data_ptr(), J * G.shape[1], G.shape[1],
                                     F.data_ptr(), J * F.shape[1], F.shape[1], out.data_ptr(), None, None, hg_h.data_ptr(),
                                     hg_l.data_ptr() if x3 else None, cnt.data_ptr(), frames, J, precision,
                                     current_stream()), "pu_chain")
    return out


def head(e, skel, Wp, bp, Wg, bg, frames, J):
    import torch
    require_cuda(e, skel, Wp, bp, Wg, bg)
    nj = J + 1 if Wg is not None else J
    pose = torch.empty((frames, nj, 3), dtype=torch.float32, device=e.device)
    check(lib().egotap_b200_head(e.data_ptr(), e.shape[1], skel.data_ptr(), Wp.data_ptr(), bp.data_ptr(), _ptr(Wg), _ptr(bg),
                                 frames, J, pose.data_ptr(), current_stream()), "head")
    return pose


def gemm(a_hi, a_lo, b_hi, b_lo, M, N, K, *, groups=1, a_group=(1, 0, 1, 0), b_group=(1, 0, 1, 0), a_rows=None,
         b_rows=None, lda=None, ldb=None, precision=PREC_BF16X3, variant=-1, **epi):
    """Op-level entry used by the tests: D = epi(A @ B^T).  ``epi`` keys mirror ``egotap_epilogue``;
    tensor-valued keys take CUDA tensors."""
    d = Gemm()
    d.a = Operand(_ptr(a_hi), _ptr(a_lo), lda or K, a_rows or M, *a_group)
    d.b = Operand(_ptr(b_hi), _ptr(b_lo), ldb or K, b_rows or N, *b_group)
    d.M, d.N, d.K, d.groups, d.precision, d.variant = M, N, K, groups, precision, variant
    e = d.epi
    e.alpha = float(epi.pop("alpha", 1.0))
    for name in ("scale", "bias", "resid", "out_f32", "out_hi", "out_lo", "vt_hi", "vt_lo"):
        t = epi.pop(name, None)
        require_cuda(t)
        setattr(e, name, _ptr(t))
    for name in ("act", "resid_ld", "resid_mod", "rows_in", "rows_out", "group_rows", "ldo", "col_off", "store",
                 "qk_cols", "tokens", "J", "heads"):
        setattr(e, name, int(epi.pop(name, 0)))
    if epi:
        raise TypeError("unknown epilogue fields: %s" % sorted(epi))
    if e.ldo == 0:
        e.ldo = N
    check(lib().egotap_b200_gemm(C.byref(d), current_stream()), "gemm")


def profile_kernels(fn):
    """Run ``fn()`` with per-launch CUDA-event timing of every kernel the library launches; returns a list of
    dicts (name, ms; GEMMs also M, N, K, groups, variant and flops = algorithmic 2*M*N*K*groups)."""
    L = lib()
    check(L.egotap_b200_profile_begin(), "profile_begin")
    try:
        fn()
    finally:
        n = C.c_int()
        check(L.egotap_b200_profile_end(C.byref(n)), "profile_end")
    out = []
    name = C.c_char_p()
    M, N, K, G, V, ms = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_float()
    for i in range(n.value):
        check(L.egotap_b200_profile_record(i, C.byref(name), C.byref(M), C.byref(N), C.byref(K), C.byref(G), C.byref(V),
                                           C.byref(ms)), "profile_record")
        r = dict(name=name.value.decode(), ms=ms.value)
        if M.value:
            r.update(M=M.value, N=N.value, K=K.value, groups=G.value,
                     variant=L.egotap_b200_gemm_variant_name(V.value).decode(),
                     flops=2.0 * M.value * N.value * K.value * G.value)
        out.append(r)
    return out


def profile_gemms(fn):
    return [r for r in profile_kernels(fn) if "flops" in r]
