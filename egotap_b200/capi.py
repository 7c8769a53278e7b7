"""ctypes binding of ``libegotap_b200.so`` (C ABI declared in ``include/egotap_b200.h``).

This is the only place Python touches the native library.  There is no fallback: if the shared
library is missing the import raises, and every compute entry needs a CUDA device.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# EGOTAP_B200_LIB: A/B runs of two builds of the same library (tools/gpu_job_*.sh); never a non-CUDA path
LIB_PATH = os.environ.get("EGOTAP_B200_LIB") or os.path.join(_HERE, "libegotap_b200.so")

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
    # training step (csrc/train_ops.cu, csrc/train_model.cu)
    "egotap_b200_zero", "egotap_b200_copy", "egotap_b200_add3", "egotap_b200_split2d", "egotap_b200_fill_dummy",
    "egotap_b200_pos_permute", "egotap_b200_pu_bridge_gate", "egotap_b200_transpose_split", "egotap_b200_transpose_bf16",
    "egotap_b200_colsum", "egotap_b200_reduce_partials", "egotap_b200_gelu_fwd", "egotap_b200_gelu_bwd",
    "egotap_b200_layernorm_bwd", "egotap_b200_softmax_bwd", "egotap_b200_bn_stats", "egotap_b200_bn_apply",
    "egotap_b200_bn_bwd", "egotap_b200_regroup_gather", "egotap_b200_pu_cell_fwd", "egotap_b200_pu_cell_bwd",
    "egotap_b200_pu_bridge_gate_bwd", "egotap_b200_pu_chain_bwd", "egotap_b200_head_bwd", "egotap_b200_embed_grads", "egotap_b200_pose_loss",
    "egotap_b200_adamw", "egotap_b200_gt_heatmaps",
    # fused attention backward (csrc/attention_bwd.cu)
    "egotap_b200_attention_lse", "egotap_b200_attn_dsum", "egotap_b200_attention_bwd",
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
                ("groups", C.c_int), ("precision", C.c_int), ("variant", C.c_int), ("layout", C.c_int), ("epi", Epilogue)]


_lib = None


def _TRAIN_ARGTYPES(P, LL, I, F):
    """argument types of the training entries, in the order include/egotap_b200.h declares them"""
    return {
        "egotap_b200_zero": [P, C.c_size_t, P],
        "egotap_b200_copy": [P, P, C.c_size_t, P],
        "egotap_b200_add3": [P, P, P, P, I, P],
        "egotap_b200_split2d": [P, LL, LL, LL, P, P, LL, P],
        "egotap_b200_fill_dummy": [P, P, I, I, I, P],
        "egotap_b200_pos_permute": [P, P, I, I, P, P, P],
        "egotap_b200_pu_bridge_gate": [P, I, I, P, I, I, LL, P, P, P],
        "egotap_b200_transpose_split": [P, LL, I, LL, I, I, P, P, LL, P, P, LL, LL, P, P, P, LL, P],
        "egotap_b200_transpose_bf16": [P, P, LL, I, LL, I, LL, I, LL, P, P, LL, LL, LL, LL, P],
        "egotap_b200_colsum": [P, LL, I, LL, I, I, P, P, LL, P],
        "egotap_b200_reduce_partials": [P, I, LL, P, P],
        "egotap_b200_gelu_fwd": [P, LL, P, P, P],
        "egotap_b200_gelu_bwd": [P, P, LL, P],
        "egotap_b200_layernorm_bwd": [P, P, P, LL, I, I, F, P, I, P, P, P, LL, P],
        "egotap_b200_softmax_bwd": [P, P, LL, I, F, P, P, P, P, P],
        "egotap_b200_attention_lse": [P, P, P, P, P, P, P, I, I, P],
        "egotap_b200_attn_dsum": [P, P, P, P, LL, P, P],
        "egotap_b200_attention_bwd": [P, P, P, P, P, P, I, P],
        "egotap_b200_bn_stats": [P, LL, I, P, P, P, P, P, F, F, P, P, P, P, P, LL, P],
        "egotap_b200_bn_apply": [P, LL, I, P, P, P, P, LL, P, LL, I, I, P],
        "egotap_b200_bn_bwd": [P, P, LL, I, P, P, P, P, P, P, P, LL, P],
        "egotap_b200_regroup_gather": [P, LL, I, LL, I, I, P, P],
        "egotap_b200_pu_cell_fwd": [P, LL, LL, P, LL, LL, P, P, P, P, P, P, I, I, LL, P],
        "egotap_b200_pu_cell_bwd": [P, LL, LL, P, LL, LL, P, P, P, P, P, P, LL, LL, P, LL, LL, P, P, I, I, LL, P],
        "egotap_b200_pu_chain_bwd": [P, P, P, LL, LL, P, LL, LL, P, P, P, P, LL, LL, P, LL, LL, P, P, P, I, I, I, P],
        "egotap_b200_pu_bridge_gate_bwd": [P, LL, P, LL, I, P, I, LL, P, LL, P],
        "egotap_b200_head_bwd": [P, P, LL, P, P, P, LL, I, P, LL, P, P, P, P, P, P, LL, P],
        "egotap_b200_embed_grads": [P, I, I, P, P, P],
        "egotap_b200_pose_loss": [P, P, LL, I, C.POINTER(C.c_int), I, I, F, F, P, P, P, LL, P],
        "egotap_b200_gt_heatmaps": [P, P, LL, I, P, P],
        "egotap_b200_adamw": [C.POINTER(P), C.POINTER(P), C.POINTER(P), C.POINTER(P), C.POINTER(LL), I, I] + [C.c_double] * 6 + [P],
    }


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
        P, LL, I, F = C.c_void_p, C.c_longlong, C.c_int, C.c_float
        for name, args in _TRAIN_ARGTYPES(P, LL, I, F).items():
            getattr(L, name).argtypes = args
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
    d = gemm_desc(a_hi, a_lo, b_hi, b_lo, M, N, K, groups=groups, a_group=a_group, b_group=b_group, a_rows=a_rows,
                  b_rows=b_rows, lda=lda, ldb=ldb, precision=precision, variant=variant, **epi)
    check(lib().egotap_b200_gemm(C.byref(d), current_stream()), "gemm")


def gemm_desc(a_hi, a_lo, b_hi, b_lo, M, N, K, *, groups=1, a_group=(1, 0, 1, 0), b_group=(1, 0, 1, 0), a_rows=None,
              b_rows=None, lda=None, ldb=None, precision=PREC_BF16X3, variant=-1, tn=False, **epi):
    """the ``egotap_gemm`` descriptor for D = epi(A @ B^T) (no launch); ``tn``: row-major operands with the contraction along
    the rows, D[g] = A[gK:(g+1)K]^T @ B[gK:(g+1)K] (a_rows = b_rows = total rows; include/egotap_b200.h EGOTAP_GEMM_TN)"""
    d = Gemm()
    d.layout = 1 if tn else 0
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
    return d


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


class CudaBackend:
    """The product backend of ``training.TrainEngine``: every method is ONE call into libegotap_b200.so on the current
    CUDA stream (same method names and argument meaning as the op oracle used by the tests, oracle/op_oracle.py).
    Tensors are passed as device pointers (``data_ptr`` of the given view); nothing here computes anything."""

    name = "cuda"

    def __init__(self, device=None):
        import torch
        self.L = lib()                                   # raises if the library has not been built
        if not torch.cuda.is_available():
            raise RuntimeError("egotap_b200 has no CPU path: the training engine needs a CUDA device")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._preset_of_J = {15: PRESET_ID["UnrealEgo"], 17: PRESET_ID["EgoCap"]}

    @property
    def launches(self):
        return self.L.egotap_b200_launch_count()

    @staticmethod
    def _st():
        return current_stream()

    # ---- one choke point for every library call, so a whole training step can be recorded once and replayed without
    # re-running the Python marshalling (the step issues ~700 launches; see TrainEngine.train_step)
    _tape = None

    def _fn(self, name):
        return getattr(self.L, name)

    def _raise(self, name, rc):
        check(rc, name)

    def _c(self, name, *args):
        fn = self._fn(name)
        rc = fn(*args)
        if rc:
            self._raise(name, rc)
        if self._tape is not None:
            self._tape.append((fn, args, name))

    def callback(self, fn, *args):
        """a host-side callback inside a step (e.g. launching the all-reduce of a finished gradient slice)"""
        fn(*args)
        if self._tape is not None:
            self._tape.append((fn, args, None))

    def begin_record(self):
        self._tape = []
        self._tape_stream = self._st()

    def end_record(self):
        tape, self._tape = self._tape, None
        return (tape, self._tape_stream)

    def can_replay(self, recorded):
        return recorded is not None and recorded[1] == self._st()

    def replay(self, recorded):
        for fn, args, name in recorded[0]:
            rc = fn(*args)
            if name is not None and rc:
                self._raise(name, rc)

    # ---- memory
    def empty(self, shape, dtype=None):
        import torch
        return torch.empty(shape, dtype=dtype or torch.float32, device=self.device)

    def zero(self, t):
        self._c("egotap_b200_zero", t.data_ptr(), t.numel() * t.element_size(), self._st())

    def copy(self, dst, src):
        assert dst.numel() == src.numel() and dst.element_size() == src.element_size()
        self._c("egotap_b200_copy", dst.data_ptr(), src.data_ptr(), dst.numel() * dst.element_size(), self._st())

    def add3(self, a, b, c, out, n):
        self._c("egotap_b200_add3", _ptr(a), _ptr(b), _ptr(c), _ptr(out), n, self._st())

    # ---- round-1 ops
    def gemm(self, *a, **k):
        d = gemm_desc(*a, **k)
        self._c("egotap_b200_gemm", C.byref(d), self._st())

    def split2d(self, src, rows, cols, src_ld, hi, lo, dst_ld):
        self._c("egotap_b200_split2d", _ptr(src), rows, cols, src_ld, _ptr(hi), _ptr(lo), dst_ld, self._st())

    def ingest(self, x, J, p_hi, p_lo, l_hi, l_lo):
        self._c("egotap_b200_ingest", x.data_ptr(), x.shape[0], self._preset_of_J[J], _ptr(p_hi), _ptr(p_lo), _ptr(l_hi),
                                        _ptr(l_lo), self._st())

    def fill_dummy(self, hidden, dummy, B, tokens, live):
        if tokens > live:
            self._c("egotap_b200_fill_dummy", _ptr(hidden), _ptr(dummy), B, tokens, live, self._st())

    def pos_permute(self, pos, mask_token, grid, n_hm, pos_perm, dummy):
        self._c("egotap_b200_pos_permute", _ptr(pos), _ptr(mask_token), grid, n_hm, _ptr(pos_perm), _ptr(dummy),
                                             self._st())

    def layernorm(self, x, w, b, frames, rows_in, rows_out, eps, out_hi, out_lo, out_f32):
        self._c("egotap_b200_layernorm", _ptr(x), _ptr(w), _ptr(b), frames, rows_in, rows_out, eps, _ptr(out_hi),
                                           _ptr(out_lo), _ptr(out_f32), self._st())

    def attention(self, qk_hi, qk_lo, vt_hi, vt_lo, ctx_hi, ctx_lo, frames, precision):
        self._c("egotap_b200_attention", _ptr(qk_hi), _ptr(qk_lo), _ptr(vt_hi), _ptr(vt_lo), _ptr(ctx_hi), _ptr(ctx_lo),
                                           frames, precision, self._st())

    def attention_lse(self, qk_hi, qk_lo, vt_hi, vt_lo, ctx_hi, ctx_lo, lse, frames, precision):
        self._c("egotap_b200_attention_lse", _ptr(qk_hi), _ptr(qk_lo), _ptr(vt_hi), _ptr(vt_lo), _ptr(ctx_hi), _ptr(ctx_lo),
                                               _ptr(lse), frames, precision, self._st())

    def attn_dsum(self, ctx_hi, ctx_lo, dctx_hi, dctx_lo, rows, dsum):
        self._c("egotap_b200_attn_dsum", _ptr(ctx_hi), _ptr(ctx_lo), _ptr(dctx_hi), _ptr(dctx_lo), rows, _ptr(dsum), self._st())

    def attention_bwd(self, qk_hi, vt_hi, dctx_hi, lse, dsum, dqkv, frames):
        self._c("egotap_b200_attention_bwd", _ptr(qk_hi), _ptr(vt_hi), _ptr(dctx_hi), _ptr(lse), _ptr(dsum), _ptr(dqkv), frames,
                                               self._st())

    def pu_bridge_gate(self, f, f_ld, f_col, e, e_ld, X, rows, hi, lo):
        self._c("egotap_b200_pu_bridge_gate", _ptr(f), f_ld, f_col, _ptr(e), e_ld, X, rows, _ptr(hi), _ptr(lo), self._st())

    def head(self, e, e_ld, skel, Wp, bp, Wg, bg, frames, J, pose):
        self._c("egotap_b200_head", _ptr(e), e_ld, _ptr(skel), _ptr(Wp), _ptr(bp), _ptr(Wg), _ptr(bg), frames, J, _ptr(pose),
                                      self._st())

    # ---- training ops
    def transpose_split(self, src, rows, cols, src_ld, rows_in, rows_out, rm_hi, rm_lo, rm_ld, t_hi, t_lo, t_ld, pad_rows,
                        gelu_u=None, colsum_out=None, scratch=None):
        self._c("egotap_b200_transpose_split", _ptr(src), rows, cols, src_ld, rows_in, rows_out, _ptr(rm_hi), _ptr(rm_lo),
                rm_ld, _ptr(t_hi), _ptr(t_lo), t_ld, pad_rows, _ptr(gelu_u), _ptr(colsum_out), _ptr(scratch),
                0 if scratch is None else scratch.numel(), self._st())

    def transpose_bf16(self, s_hi, s_lo, rows, cols, s_ld, g0c, s_g0s, g1c, s_g1s, d_hi, d_lo, d_ld, d_g0s, d_g1s, pad_rows):
        self._c("egotap_b200_transpose_bf16", _ptr(s_hi), _ptr(s_lo), rows, cols, s_ld, g0c, s_g0s, g1c, s_g1s, _ptr(d_hi),
                                                _ptr(d_lo), d_ld, d_g0s, d_g1s, pad_rows, self._st())

    def colsum(self, src, rows, cols, ld, rows_in, rows_out, out, scratch):
        self._c("egotap_b200_colsum", _ptr(src), rows, cols, ld, rows_in, rows_out, _ptr(out), _ptr(scratch),
                                        scratch.numel(), self._st())

    def reduce_partials(self, partials, G, n, out):
        self._c("egotap_b200_reduce_partials", _ptr(partials), G, n, _ptr(out), self._st())

    def gelu_fwd(self, u, n, out_hi, out_lo):
        self._c("egotap_b200_gelu_fwd", _ptr(u), n, _ptr(out_hi), _ptr(out_lo), self._st())

    def gelu_bwd(self, dg, u, n):
        self._c("egotap_b200_gelu_bwd", _ptr(dg), _ptr(u), n, self._st())

    def layernorm_bwd(self, dy, x, w, frames, rows_in, rows_out, eps, dx, accumulate, dw, db, scratch):
        self._c("egotap_b200_layernorm_bwd", _ptr(dy), _ptr(x), _ptr(w), frames, rows_in, rows_out, eps, _ptr(dx),
                                               int(accumulate), _ptr(dw), _ptr(db), _ptr(scratch), scratch.numel(),
                                               self._st())

    def softmax_bwd(self, S, dP, rows, cols, scale, p_hi, p_lo, ds_hi, ds_lo):
        self._c("egotap_b200_softmax_bwd", _ptr(S), _ptr(dP), rows, cols, scale, _ptr(p_hi), _ptr(p_lo), _ptr(ds_hi),
                                             _ptr(ds_lo), self._st())

    def bn_stats(self, y, rows, cols, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps, mean, rstd,
                 scale, shift, scratch):
        self._c("egotap_b200_bn_stats", _ptr(y), rows, cols, _ptr(gamma), _ptr(beta), _ptr(running_mean), _ptr(running_var),
                                          _ptr(num_batches_tracked), momentum, eps, _ptr(mean), _ptr(rstd), _ptr(scale),
                                          _ptr(shift), _ptr(scratch), scratch.numel(), self._st())

    def bn_apply(self, y, rows, cols, scale, shift, out_hi, out_lo, out_ld, out_f32, f32_ld, J, col_off):
        self._c("egotap_b200_bn_apply", _ptr(y), rows, cols, _ptr(scale), _ptr(shift), _ptr(out_hi), _ptr(out_lo), out_ld,
                                          _ptr(out_f32), f32_ld, J, col_off, self._st())

    def bn_bwd(self, da, y, rows, cols, scale, shift, mean, rstd, dgamma, dbeta, scratch):
        self._c("egotap_b200_bn_bwd", _ptr(da), _ptr(y), rows, cols, _ptr(scale), _ptr(shift), _ptr(mean), _ptr(rstd),
                                        _ptr(dgamma), _ptr(dbeta), _ptr(scratch), scratch.numel(), self._st())

    def regroup_gather(self, dE, e_ld, col_off, frames, J, cols, out):
        self._c("egotap_b200_regroup_gather", _ptr(dE), e_ld, col_off, frames, J, cols, _ptr(out), self._st())

    def pu_cell_fwd(self, G, g_rs, g_ts, F_, f_rs, f_ts, C_all, H, h_hi, h_lo, hg_hi, hg_lo, t, J, B):
        self._c("egotap_b200_pu_cell_fwd", _ptr(G), g_rs, g_ts, _ptr(F_), f_rs, f_ts, _ptr(C_all), _ptr(H), _ptr(h_hi),
                                             _ptr(h_lo), _ptr(hg_hi), _ptr(hg_lo), t, J, B, self._st())

    def pu_cell_bwd(self, G, g_rs, g_ts, F_, f_rs, f_ts, C_all, H, dOut, dhg, dc, dG, dg_rs, dg_ts, dF, df_rs, df_ts, dgp_hi,
                    dgp_lo, t, J, B):
        self._c("egotap_b200_pu_cell_bwd", _ptr(G), g_rs, g_ts, _ptr(F_), f_rs, f_ts, _ptr(C_all), _ptr(H), _ptr(dOut),
                                             _ptr(dhg), _ptr(dc), _ptr(dG), dg_rs, dg_ts, _ptr(dF), df_rs, df_ts, _ptr(dgp_hi),
                                             _ptr(dgp_lo), t, J, B, self._st())

    def pu_chain_bwd(self, wT_hi, wT_lo, G, g_rs, g_ts, F_, f_rs, f_ts, C_all, H, dOut, dG, dg_rs, dg_ts, dF, df_rs, df_ts, x_hi,
                     x_lo, counters, B, J, precision):
        self._c("egotap_b200_pu_chain_bwd", _ptr(wT_hi), _ptr(wT_lo), _ptr(G), g_rs, g_ts, _ptr(F_), f_rs, f_ts, _ptr(C_all),
                _ptr(H), _ptr(dOut), _ptr(dG), dg_rs, dg_ts, _ptr(dF), df_rs, df_ts, _ptr(x_hi), _ptr(x_lo), _ptr(counters), B, J,
                precision, self._st())

    def pu_bridge_gate_bwd(self, dE, e_ld, F0, f_ld, f_col, E, X, rows, dF, df_ld):
        self._c("egotap_b200_pu_bridge_gate_bwd", _ptr(dE), e_ld, _ptr(F0), f_ld, f_col, _ptr(E), X, rows, _ptr(dF), df_ld,
                                                    self._st())

    def head_bwd(self, dpose, e, e_ld, skel, Wp, Wg, frames, J, dE, de_ld, dSkel, dWp, dbp, dWg, dbg, scratch):
        self._c("egotap_b200_head_bwd", _ptr(dpose), _ptr(e), e_ld, _ptr(skel), _ptr(Wp), _ptr(Wg), frames, J, _ptr(dE), de_ld,
                                          _ptr(dSkel), _ptr(dWp), _ptr(dbp), _ptr(dWg), _ptr(dbg), _ptr(scratch),
                                          scratch.numel(), self._st())

    def embed_grads(self, dpos_perm, grid, n_hm, dpos, dmask):
        self._c("egotap_b200_embed_grads", _ptr(dpos_perm), grid, n_hm, _ptr(dpos), _ptr(dmask), self._st())

    def pose_loss(self, pred, gt, frames, nj, parents, drop_first, lambda_mpjpe, lambda_cos, loss, dpose, scratch=None):
        arr = (C.c_int * len(parents))(*parents)
        if scratch is None:
            scratch = self.empty((2 * frames + 64,))
        self._c("egotap_b200_pose_loss", _ptr(pred), _ptr(gt), frames, nj, arr, len(parents), int(drop_first), lambda_mpjpe,
                                           lambda_cos, _ptr(loss), _ptr(dpose), _ptr(scratch), scratch.numel(), self._st())

    def gt_heatmaps(self, pts2d, pts3d_left, frames, preset, out):
        self._c("egotap_b200_gt_heatmaps", _ptr(pts2d), _ptr(pts3d_left), frames, PRESET_ID[preset], _ptr(out), self._st())

    def adamw(self, params, grads, m, v, step, lr, beta1, beta2, eps, weight_decay, grad_scale=1.0):
        n = len(params)
        mk = lambda ts: (C.c_void_p * n)(*[t.data_ptr() for t in ts])
        numel = (C.c_longlong * n)(*[p.numel() for p in params])
        self._c("egotap_b200_adamw", mk(params), mk(grads), mk(m), mk(v), numel, n, step, lr, beta1, beta2, eps, weight_decay,
                                       grad_scale, self._st())
