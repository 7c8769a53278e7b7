"""Training step of the lifting network (BASELINE config 5).

Replaces, for ``net_AutoEncoder``, what the reference's ``optimize_parameters`` does through torch.autograd
(reference model/egotap_autoencoder_model.py:284-323): train-mode forward (BatchNorm1d batch statistics,
model/network_utils.py:123-142), MPJPE + bone-cosine loss (utils/loss.py:44-85), backward through the head, the
2-layer propagation chain (BPTT over the joints), both FC encoders and the 3-layer ViT, and AdamW
(model/network.py:72-78).

``TrainEngine`` is host-side orchestration only: every piece of arithmetic is one call into the C ABI
(include/egotap_b200.h) -- the round-1 tcgen05 GEMM / fused attention / LayerNorm / ingest kernels for the forward and
for every dX / dW contraction of the backward, plus the bandwidth-bound training kernels of csrc/train.cu.  The backend
object is the ctypes binding (``capi.CudaBackend``); there is no CPU or PyTorch fallback in the product (the
``backend=`` argument exists so tests can execute this same orchestration against the op oracle, oracle/op_oracle.py).

Dataflow conventions
  * activations that feed a GEMM are bf16 "pairs" (hi, lo = bf16(x - hi)); lo is absent in the 'bf16' precision mode
  * gradients flow in fp32; ``transpose_split`` turns a gradient into its row-major bf16 pair (and, in the same pass, applies
    gelu' and forms the bias-gradient column sums)
  * dW[n][k] = sum_r dY[r][n] X[r][k] runs as a split-K GEMM straight from the ROW-major pairs dY (r x n) and X (r x k) with the
    contraction along the rows (EGOTAP_GEMM_TN: MN-major shared-memory descriptors, no transposed copies): r is cut into G
    chunks (GEMM groups), partial products are summed by ``reduce_partials``
  * parameter gradients live in ONE flat fp32 buffer ordered by backward completion (``stages``), so data-parallel
    training all-reduces contiguous slices while earlier layers are still being differentiated
"""
import math
import os

import torch

HID, TOK, MLPD, HEADS, HD, EMB, PUH, PUX = 1024, 576, 4096, 8, 128, 128, 512, 256
LN_EPS, BN_EPS, BN_MOMENTUM = 1e-12, 1e-5, 0.1
ACT_NONE, ACT_GELU, ACT_LRELU = 0, 1, 2
STORE_QKV, STORE_HEAD_MERGE = 1, 3
KINEMATIC_PARENTS = {   # reference utils/util.py:51-52
    "UnrealEgo": [0, 0, 1, 1, 2, 3, 4, 5, 2, 3, 8, 9, 10, 11, 12, 13],
    "EgoCap": [0, 0, 1, 2, 3, 4, 1, 6, 7, 8, 2, 10, 11, 12, 6, 14, 15, 16],
}
SPLITK_MAX = 32


class Pair:
    """bf16 hi/lo operand pair; ``lo`` is None in the plain-bf16 mode.  Slicing returns pointer-offset views."""
    __slots__ = ("hi", "lo")

    def __init__(self, hi, lo=None):
        self.hi, self.lo = hi, lo

    @classmethod
    def alloc(cls, be, rows, cols, x3, zero=False):
        hi = be.empty((rows, cols), torch.bfloat16)
        lo = be.empty((rows, cols), torch.bfloat16) if x3 else None
        if zero:
            be.zero(hi)
            if lo is not None:
                be.zero(lo)
        return cls(hi, lo)

    def at(self, row=0, col=0):
        return Pair(self.hi[row:, col:], None if self.lo is None else self.lo[row:, col:])


def pad_ld(rows):
    """leading dimension of a transposed (feature x row) copy: rows rounded up to 64 plus slack so that any split-K
    chunking (chunks are multiples of 64) stays inside zero-filled memory"""
    r64 = (rows + 63) // 64 * 64
    return r64 + 64 * min(SPLITK_MAX, r64 // 64)


def splitk(n_out, k_out, rows):
    """(G, chunk): number of reduction chunks and their length for a dW GEMM with an (n_out x k_out) result"""
    tiles = ((n_out + 255) // 256) * ((k_out + 255) // 256)
    r64 = (rows + 63) // 64 * 64
    G = max(1, min(SPLITK_MAX, -(-148 // tiles), max(1, r64 // 256)))
    chunk = -(-(r64 // 64) // G) * 64
    G = -(-r64 // chunk)
    return G, chunk


def param_order(preset):
    """parameter names grouped by the backward stage that completes their gradient (flat-buffer order)"""
    v = "pos_heatmap_encoder.vit."
    pu = "skel_sequential_layer.lstm_custom.layers."
    stages = []
    head = ["pose_mlp.pose_fcs.0.weight", "pose_mlp.pose_fcs.0.bias"]
    if preset == "UnrealEgo":
        head += ["global_mlp.pose_fcs.0.weight", "global_mlp.pose_fcs.0.bias"]
    stages.append(("head", head))
    stages.append(("chain", [pu + n + s for n in ("1.x2f", "1.x2h", "1.h2h", "0.x2f", "0.x2h", "0.b2h", "0.h2h")
                             for s in (".weight", ".bias")]))
    for enc in ("rot_heatmap_encoder.", "pos_heatmap_encoder."):
        stages.append((enc + "fc", [enc + fc + s for fc in ("fc3.", "fc2.", "fc1.")
                                    for s in ("fc.weight", "fc.bias", "bn.weight", "bn.bias")]))
    stages.append(("final_ln", [v + "layernorm.weight", v + "layernorm.bias"]))
    for l in (2, 1, 0):
        p = v + "encoder.layer.%d." % l
        names = []
        for s in ("output.dense", "intermediate.dense", "layernorm_after", "attention.output.dense",
                  "attention.attention.query", "attention.attention.key", "attention.attention.value", "layernorm_before"):
            names += [p + s + ".weight", p + s + ".bias"]
        stages.append(("layer%d" % l, names))
    stages.append(("embed", [v + "embeddings.patch_embeddings.projection.weight",
                             v + "embeddings.patch_embeddings.projection.bias",
                             v + "embeddings.position_embeddings", v + "embeddings.mask_token"]))
    return stages


class TrainEngine:
    def __init__(self, preset, params, precision="bf16", backend=None, attn_chunk=64,
                 lambda_mpjpe=0.1, lambda_cos_sim=-0.01):
        """params: dict name -> fp32 contiguous tensor with the reference's state_dict keys (parameters AND BatchNorm
        buffers); tensors are referenced, not copied: AdamW updates them in place."""
        if backend is None:
            from . import capi
            backend = capi.CudaBackend()          # raises if the native library is missing: no fallback
        self.be = backend
        if preset not in KINEMATIC_PARENTS:
            raise ValueError("joint_preset is {} which is undefined".format(preset))
        self.preset = preset
        self.J = 15 if preset == "UnrealEgo" else 17
        self.global_head = preset == "UnrealEgo"
        self.nj = self.J + 1 if self.global_head else self.J
        self.n_hm = 2 * self.J
        self.live = self.n_hm * 16
        self.grid = 6
        self.x3 = precision in ("bf16x3", "fp32")
        self.precision = 0 if self.x3 else 1
        self.attn_chunk = attn_chunk
        self.lambda_mpjpe, self.lambda_cos_sim = lambda_mpjpe, lambda_cos_sim
        self.P = params
        for k, t in params.items():
            if t.is_floating_point() and (t.dtype != torch.float32 or not t.is_contiguous()):
                raise RuntimeError("egotap_b200 expects contiguous fp32 parameters (%s is %s)" % (k, t.dtype))
        # ---- flat gradient / optimiser-state buffers
        self.stage_names, self.order, self.stages = [], [], []
        off = 0
        self.offsets = {}
        for name, keys in param_order(preset):
            start = off
            for k in keys:
                self.offsets[k] = off
                off += (params[k].numel() + 63) // 64 * 64          # 256-byte aligned slots
                self.order.append(k)
            self.stage_names.append(name)
            self.stages.append((start, off))
        self.flat_grad = self.be.empty((off,), torch.float32)
        self.be.zero(self.flat_grad)
        self.grad = {k: self.flat_grad[o:o + params[k].numel()].view(params[k].shape) for k, o in self.offsets.items()}
        self.opt_m = self.opt_v = None
        self.opt_step = 0
        self.mutation = 0          # bumped whenever the engine writes parameters or BatchNorm buffers through raw pointers
        self.fwd_gen = 0           # bumped by every forward: a backward belongs to exactly one forward's activations
        self.batch = 0
        # Replay of a recorded step: the ~700 library calls of forward + loss + backward are recorded once per batch size
        # (backend.begin_record) and re-issued from the record afterwards, skipping this file's Python entirely -- at
        # the reference's batch size (32) the host would otherwise be the bottleneck.  Needs stable device pointers:
        # train_step copies its inputs into engine-owned buffers.
        self.use_tape = hasattr(self.be, "begin_record")
        self._tape, self._tape_key = None, None
        # Opt-in: the same step captured in a CUDA graph (single-GPU; the staged all-reduce stays on the replay path).
        # Step 1 runs eagerly (allocations, per-device kernel attributes), step 2 is captured, later steps replay.
        self.use_cuda_graph = False
        # Opt-in: BPTT of each propagation layer as ONE persistent launch (csrc/pu_chain_bwd.cu) instead of J x
        # (cell-backward kernel + dgates . W_hh GEMM); same arithmetic
        self.persistent_bptt = False
        # bf16-operand mode: attention backward as two fused tcgen05 kernels that keep scores / probabilities on chip
        # (csrc/attention_bwd.cu); the fp32-parity mode (gradient checks) keeps the GEMM-based path below.  Set before the
        # first step; False selects the GEMM-based path in bf16 mode too (A/B).
        self.fused_attention_bwd = (not self.x3) and hasattr(self.be, "attention_bwd") and os.environ.get("EGOTAP_ATTN_BWD") != "gemm"
        self._graph, self._graph_key, self._graph_warm = None, None, None
        self._alloc_weights()
        self.scr = self.be.empty((4 * 1024 * 1024,), torch.float32)     # reduction scratch shared by the small ops
        self.partials = None
        self.loss = self.be.empty((4,), torch.float32)

    # ------------------------------------------------------------------------------------------ helpers
    def _pair(self, rows, cols, zero=False):
        return Pair.alloc(self.be, rows, cols, self.x3, zero)

    def _f32(self, *shape):
        return self.be.empty(tuple(shape), torch.float32)

    def _gemm(self, A, lda, M, K, W, ldb, N, **epi):
        out = epi.pop("out", None)
        if out is not None:
            epi["out_hi"], epi["out_lo"] = out.hi, out.lo
        self.be.gemm(A.hi, A.lo, W.hi, W.lo, M, N, K, lda=lda, ldb=ldb, precision=self.precision, **epi)

    def _dw(self, dY, ldY, X, ldX, n_out, k_out, rows, grad):
        """grad[n_out][k_out] = sum_r dY[r][n] X[r][k] from the row-major pairs dY (rows x n_out, row stride ldY) and
        X (rows x k_out, row stride ldX); rows beyond ``rows`` are zero-filled by the TMA loads"""
        G, chunk = splitk(n_out, k_out, rows)
        kw = dict(a_rows=rows, b_rows=rows, lda=ldY, ldb=ldX, precision=self.precision, tn=True, ldo=k_out)
        if G == 1:
            self.be.gemm(dY.hi, dY.lo, X.hi, X.lo, n_out, k_out, chunk, out_f32=grad, **kw)
            return
        need = G * n_out * k_out
        # sized once per batch size in _alloc: a recorded step (tape / CUDA graph) keeps raw device pointers, so no buffer
        # of the engine may be reallocated between _alloc calls
        assert self.partials is not None and self.partials.numel() >= need, "split-K scratch undersized (see _dw_shapes)"
        self.be.gemm(dY.hi, dY.lo, X.hi, X.lo, n_out, k_out, chunk, groups=G, out_f32=self.partials, group_rows=n_out, **kw)
        self.be.reduce_partials(self.partials, G, n_out * k_out, grad)

    def _dw_shapes(self, B):
        """(n_out, k_out, rows) of every weight-gradient contraction of one step (the calls of _dw in backward)"""
        M, R, RJ, Ml = B * TOK, B * self.n_hm, B * self.J, B * self.live
        shapes = [(PUH, PUH, RJ), (4 * PUH, PUH, RJ), (4 * PUH, PUX, RJ), (PUH + PUX, PUX, RJ)]
        shapes += [(2048, 16 * HID, R), (2048, 8192, R), (512, 2048, R), (EMB, 512, R)]
        shapes += [(HID, MLPD, M), (MLPD, HID, M), (HID, HID, M), (HID, 256, Ml)]
        return shapes

    def _tsplit(self, src, rows, cols, src_ld, rm, rows_in=0, rows_out=0, bias=None, gelu_u=None):
        """fp32 gradient -> row-major bf16 pair ``rm`` (rows x cols, row stride cols); in the same pass: multiply by
        gelu'(gelu_u) first, and the column sums (bias gradient) into ``bias``"""
        self.be.transpose_split(src, rows, cols, src_ld, rows_in, rows_out, rm.hi, rm.lo, cols, None, None, 0, 0,
                                gelu_u, bias, None if bias is None else self.S["colpart"])

    def _tbf16(self, src, rows, cols, src_ld, dst, dst_ld, groups=(1, 0, 1, 0), dst_groups=(0, 0), pad=None):
        g0c, s_g0s, g1c, s_g1s = groups
        self.be.transpose_bf16(src.hi, src.lo, rows, cols, src_ld, g0c, s_g0s, g1c, s_g1s, dst.hi, dst.lo, dst_ld,
                               dst_groups[0], dst_groups[1], dst_ld if pad is None else pad)

    # ------------------------------------------------------------------------------------------ weights
    def _alloc_weights(self):
        mk = self._pair
        self.W, self.WT = {}, {}

        def both(name, n, k, t=True):
            self.W[name] = mk(n, k)
            if t:
                self.WT[name] = mk(k, n)
        both("patch", HID, 256, t=False)
        for l in range(3):
            both("qkv%d" % l, 3 * HID, HID)
            both("o%d" % l, HID, HID)
            both("up%d" % l, MLPD, HID)
            both("down%d" % l, HID, MLPD)
        for e, k1 in (("p", 16 * HID), ("r", 2 * 64 * 64)):
            both(e + "fc1", 2048, k1, t=(e == "p"))
            both(e + "fc2", 512, 2048)
            both(e + "fc3", EMB, 512)
        both("x2f0", PUH + PUX, PUX)
        both("xb0", 4 * PUH, 2 * PUX)
        both("hh0", 4 * PUH, PUH)
        both("cat1", 5 * PUH, PUH)
        both("hh1", 4 * PUH, PUH)
        self.b_qkv = [self._f32(3 * HID) for _ in range(3)]
        self.b_g0 = self._f32(4 * PUH)
        self.b_cat1 = self._f32(5 * PUH)
        self.pos_perm = self._f32(TOK, HID)
        self.dummy = self._f32(max(TOK - self.live, 1), HID)
        self.packed = False

    def pack(self):
        """fp32 parameters -> bf16 operand pairs W (n x k) and W^T (k x n); run after every optimiser step"""
        be, P = self.be, self.P
        v = "pos_heatmap_encoder.vit."

        def put(name, key, n, k, row0=0, col0=0):
            """parameter `key` (n x k) into W[name] at (row0, col0) and W^T[name] at (col0, row0)"""
            W = self.W[name]
            WT = self.WT.get(name)
            wd = W.hi.shape[1]
            rm = W.at(row0, col0)
            if WT is None:
                be.split2d(P[key], n, k, k, rm.hi, rm.lo, wd)
            else:
                t = WT.at(col0, row0)
                td = WT.hi.shape[1]
                be.transpose_split(P[key], n, k, k, 0, 0, rm.hi, rm.lo, wd, t.hi, t.lo, td, n)
        put("patch", v + "embeddings.patch_embeddings.projection.weight", HID, 256)
        be.pos_permute(P[v + "embeddings.position_embeddings"], P[v + "embeddings.mask_token"], self.grid, self.n_hm,
                       self.pos_perm, self.dummy)
        for l in range(3):
            p = v + "encoder.layer.%d." % l
            for q, n in enumerate(("query", "key", "value")):
                put("qkv%d" % l, p + "attention.attention.%s.weight" % n, HID, HID, row0=q * HID)
                be.copy(self.b_qkv[l][q * HID:(q + 1) * HID], P[p + "attention.attention.%s.bias" % n])
            put("o%d" % l, p + "attention.output.dense.weight", HID, HID)
            put("up%d" % l, p + "intermediate.dense.weight", MLPD, HID)
            put("down%d" % l, p + "output.dense.weight", HID, MLPD)
        for e, enc in (("p", "pos_heatmap_encoder."), ("r", "rot_heatmap_encoder.")):
            for i, (n, k) in enumerate(((2048, self.W[e + "fc1"].hi.shape[1]), (512, 2048), (EMB, 512))):
                put("%sfc%d" % (e, i + 1), "%sfc%d.fc.weight" % (enc, i + 1), n, k)
        pu = "skel_sequential_layer.lstm_custom.layers."
        put("x2f0", pu + "0.x2f.weight", PUH + PUX, PUX)
        put("xb0", pu + "0.x2h.weight", 4 * PUH, PUX)                       # K-concatenated [x2h | b2h]
        put("xb0", pu + "0.b2h.weight", 4 * PUH, PUX, col0=PUX)
        put("hh0", pu + "0.h2h.weight", 4 * PUH, PUH)
        put("cat1", pu + "1.x2f.weight", PUH, PUH)                          # N-concatenated [x2f ; x2h]
        put("cat1", pu + "1.x2h.weight", 4 * PUH, PUH, row0=PUH)
        put("hh1", pu + "1.h2h.weight", 4 * PUH, PUH)
        be.add3(P[pu + "0.x2h.bias"], P[pu + "0.b2h.bias"], P[pu + "0.h2h.bias"], self.b_g0, 4 * PUH)
        be.copy(self.b_cat1[:PUH], P[pu + "1.x2f.bias"])
        be.add3(P[pu + "1.x2h.bias"], P[pu + "1.h2h.bias"], None, self.b_cat1[PUH:], 4 * PUH)
        self.packed = True

    # ------------------------------------------------------------------------------------------ activations
    def _alloc(self, B):
        if self.batch == B and getattr(self, "_alloc_fused", None) == bool(self.fused_attention_bwd):
            return
        J, n_hm, live = self.J, self.n_hm, self.live
        M, R, RJ = B * TOK, B * n_hm, B * J
        f32, pair = self._f32, self._pair
        A = self.A = {}
        A["a_patch"], A["a_limb"] = pair(B * live, 256), pair(R, 8192)
        A["h_in"] = [f32(M, HID) for _ in range(4)]
        A["h_mid"] = [f32(M, HID) for _ in range(3)]
        for n, c in (("ln1", HID), ("qk", 2 * HID), ("ctx", HID), ("ln2", HID), ("g", MLPD)):
            A[n] = [pair(M, c) for _ in range(3)]
        A["vt"] = [pair(B * HEADS * HD, TOK) for _ in range(3)]        # V^T per (frame, head), as STORE_QKV writes it
        A["u"] = [f32(M, MLPD) for _ in range(3)]
        fused = self._alloc_fused = bool(self.fused_attention_bwd)
        if fused:
            A["lse"] = [f32(B * HEADS * TOK) for _ in range(3)]     # per-row log-sum-exp of the forward (exp2 domain)
        A["fin"] = pair(R, 16 * HID)
        for e in ("p", "r"):
            for i, n in enumerate((2048, 512, EMB)):
                A["%sy%d" % (e, i + 1)] = f32(R, n)
                for s in ("mean", "rstd", "scale", "shift"):
                    A["%s%s%d" % (e, s, i + 1)] = f32(n)
            A[e + "a1"], A[e + "a2"] = pair(R, 2048), pair(R, 512)
        A["E"], A["xb"] = f32(RJ, 2 * PUX), pair(RJ, 2 * PUX)
        A["F0"], A["G0"], A["FG1"] = f32(RJ, PUH + PUX), f32(RJ, 4 * PUH), f32(RJ, 5 * PUH)
        A["C0"], A["C1"], A["H0"], A["skel"] = f32(RJ, PUH), f32(RJ, PUH), f32(RJ, PUH), f32(RJ, PUH)
        A["h0b"] = pair(RJ, PUH)
        A["HG0"], A["HG1"] = pair(RJ, PUH, zero=True), pair(RJ, PUH, zero=True)   # row b*J + 0 stays zero (h_-1 = 0)
        A["pose"] = f32(B, self.nj, 3)
        # ---- backward scratch
        S = self.S = {}
        S["dH"], S["dA"] = f32(M, HID), f32(M, MLPD)
        S["rm"] = pair(M, MLPD)                               # row-major bf16 copy of the current gradient
        S["dctx"] = pair(M, HID)
        if fused:
            S["dsum"] = f32(B * HEADS * TOK)
        else:       # score / probability scratch of the GEMM-based attention backward, per chunk of frames
            Bc = min(B, self.attn_chunk)
            Gc = Bc * HEADS
            S["Sc"], S["dP"] = f32(Gc * TOK, TOK), f32(Gc * TOK, TOK)
            for n in ("P", "dS", "PT", "dST"):
                S[n] = pair(Gc * TOK, TOK)
            for n in ("QT", "KT", "dctxT"):
                S[n] = pair(Gc * HD, TOK)
            S["V"] = pair(Gc * TOK, HD)
        S["da"] = f32(R, 16 * HID)                            # FC-encoder gradient ping
        S["db"] = f32(R, 2048)                                # ... and pong
        S["dE"], S["dSkel"], S["dH0"] = f32(RJ, 2 * PUX), f32(RJ, PUH), f32(RJ, PUH)
        S["dFG1"], S["dG0"], S["dF0"] = f32(RJ, 5 * PUH), f32(RJ, 4 * PUH), f32(RJ, PUH + PUX)
        S["dhg"], S["dc"], S["dgp"] = f32(B, PUH), f32(B, PUH), pair(B, 4 * PUH)
        S["dgx"] = pair(2 * B, 4 * PUH)                       # exchange buffer of the persistent BPTT kernel
        S["counters"] = self.be.empty((64,), torch.int32)
        S["dpos"] = f32(TOK, HID)
        S["colpart"] = f32(((M + 63) // 64) * MLPD)           # per-tile column sums of transpose_split (bias gradients)
        S["bias_tmp"] = f32(5 * PUH)
        S["dpose"] = f32(B, self.nj, 3)
        self.x_in, self.gt_in = f32(B, 6 * J, 64, 64), f32(B, self.nj, 3)
        self.partials = f32(max(splitk(n, k, r)[0] * n * k for n, k, r in self._dw_shapes(B)))
        self.batch = B
        self._tape = None
        self._graph = self._graph_key = self._graph_warm = None      # recorded steps hold pointers into the old buffers

    # ------------------------------------------------------------------------------------------ forward
    def forward(self, x):
        """train-mode forward; x: (B, 6J, 64, 64) fp32 contiguous.  Returns the engine-owned pose buffer (B, nj, 3)
        and leaves every activation the backward needs in self.A.  Updates the BatchNorm running buffers."""
        be, P, J = self.be, self.P, self.J
        B = x.shape[0]
        self._alloc(B)
        if not self.packed:
            self.pack()
        self.fwd_gen += 1
        self.mutation += 1          # BatchNorm running buffers are updated below
        A, W = self.A, self.W
        M, R, RJ, live = B * TOK, B * self.n_hm, B * J, self.live
        v = "pos_heatmap_encoder.vit."
        be.ingest(x, J, A["a_patch"].hi, A["a_patch"].lo, A["a_limb"].hi, A["a_limb"].lo)
        h = A["h_in"][0]
        self._gemm(A["a_patch"], 256, B * live, 256, W["patch"], 256, HID,
                   bias=P[v + "embeddings.patch_embeddings.projection.bias"], resid=self.pos_perm, resid_ld=HID,
                   resid_mod=live, rows_in=live, rows_out=TOK, out_f32=h, ldo=HID)
        be.fill_dummy(h, self.dummy, B, TOK, live)
        for l in range(3):
            p = v + "encoder.layer.%d." % l
            h_in, h_mid, h_out = A["h_in"][l], A["h_mid"][l], A["h_in"][l + 1]
            ln1, qk, vt, ctx, ln2, u, g = (A[n][l] for n in ("ln1", "qk", "vt", "ctx", "ln2", "u", "g"))
            be.layernorm(h_in, P[p + "layernorm_before.weight"], P[p + "layernorm_before.bias"], B, TOK, TOK, LN_EPS,
                         ln1.hi, ln1.lo, None)
            self._gemm(ln1, HID, M, HID, W["qkv%d" % l], HID, 3 * HID, bias=self.b_qkv[l], store=STORE_QKV,
                       qk_cols=2 * HID, tokens=TOK, out=qk, ldo=2 * HID, vt_hi=vt.hi, vt_lo=vt.lo)
            if self._alloc_fused:
                be.attention_lse(qk.hi, qk.lo, vt.hi, vt.lo, ctx.hi, ctx.lo, A["lse"][l], B, self.precision)
            else:
                be.attention(qk.hi, qk.lo, vt.hi, vt.lo, ctx.hi, ctx.lo, B, self.precision)
            self._gemm(ctx, HID, M, HID, W["o%d" % l], HID, HID, bias=P[p + "attention.output.dense.bias"],
                       resid=h_in, resid_ld=HID, out_f32=h_mid, ldo=HID)
            be.layernorm(h_mid, P[p + "layernorm_after.weight"], P[p + "layernorm_after.bias"], B, TOK, TOK, LN_EPS,
                         ln2.hi, ln2.lo, None)
            self._gemm(ln2, HID, M, HID, W["up%d" % l], HID, MLPD, bias=P[p + "intermediate.dense.bias"],
                       out_f32=u, ldo=MLPD)
            be.gelu_fwd(u, M * MLPD, g.hi, g.lo)
            self._gemm(g, MLPD, M, MLPD, W["down%d" % l], MLPD, HID, bias=P[p + "output.dense.bias"],
                       resid=h_mid, resid_ld=HID, out_f32=h_out, ldo=HID)
        fin = A["fin"]
        be.layernorm(A["h_in"][3], P[v + "layernorm.weight"], P[v + "layernorm.bias"], B, TOK, live, LN_EPS,
                     fin.hi, fin.lo, None)
        # ---- FC encoders: Linear -> BatchNorm1d (batch statistics) -> LeakyReLU, three times each
        for e, enc in (("p", "pos_heatmap_encoder."), ("r", "rot_heatmap_encoder.")):
            a_in, k = (fin, 16 * HID) if e == "p" else (A["a_limb"], 8192)
            for i, n in enumerate((2048, 512, EMB)):
                pre = "%sfc%d." % (enc, i + 1)
                y = A["%sy%d" % (e, i + 1)]
                self._gemm(a_in, k, R, k, W["%sfc%d" % (e, i + 1)], k, n, bias=P[pre + "fc.bias"], out_f32=y, ldo=n)
                mean, rstd, scale, shift = (A["%s%s%d" % (e, s, i + 1)] for s in ("mean", "rstd", "scale", "shift"))
                be.bn_stats(y, R, n, P[pre + "bn.weight"], P[pre + "bn.bias"], P[pre + "bn.running_mean"],
                            P[pre + "bn.running_var"], P.get(pre + "bn.num_batches_tracked"), BN_MOMENTUM, BN_EPS,
                            mean, rstd, scale, shift, self.scr)
                if i < 2:
                    a_out = A["%sa%d" % (e, i + 1)]
                    be.bn_apply(y, R, n, scale, shift, a_out.hi, a_out.lo, n, None, 0, 0, 0)
                    a_in, k = a_out, n
                elif e == "p":      # joint embedding: E[:, :256] fp32 + the x half of the layer-0 [x | b'] operand
                    be.bn_apply(y, R, n, scale, shift, A["xb"].hi, A["xb"].lo, 2 * PUX, A["E"], 2 * PUX, J, 0)
                else:               # limb embedding (the bridge): E[:, 256:] fp32 only
                    be.bn_apply(y, R, n, scale, shift, None, None, 0, A["E"], 2 * PUX, J, PUX)
        # ---- propagation chain (chain semantics; reference custom_cells.py:94-120,149-197)
        pu = "skel_sequential_layer.lstm_custom.layers."
        xb = A["xb"]
        self._gemm(xb, 2 * PUX, RJ, PUX, W["x2f0"], PUX, PUH + PUX, bias=P[pu + "0.x2f.bias"], out_f32=A["F0"],
                   ldo=PUH + PUX)
        be.pu_bridge_gate(A["F0"], PUH + PUX, PUH, A["E"], 2 * PUX, PUX, RJ, xb.hi, xb.lo)
        self._gemm(xb, 2 * PUX, RJ, 2 * PUX, W["xb0"], 2 * PUX, 4 * PUH, bias=self.b_g0, out_f32=A["G0"], ldo=4 * PUH)
        self._chain_fwd(B, A["G0"], 4 * PUH, A["F0"], PUH + PUX, A["C0"], A["H0"], A["h0b"], A["HG0"], W["hh0"])
        self._gemm(A["h0b"], PUH, RJ, PUH, W["cat1"], PUH, 5 * PUH, bias=self.b_cat1, out_f32=A["FG1"], ldo=5 * PUH)
        self._chain_fwd(B, A["FG1"][:, PUH:], 5 * PUH, A["FG1"], 5 * PUH, A["C1"], A["skel"], None, A["HG1"], W["hh1"])
        be.head(A["E"], 2 * PUX, A["skel"], P["pose_mlp.pose_fcs.0.weight"], P["pose_mlp.pose_fcs.0.bias"],
                P.get("global_mlp.pose_fcs.0.weight") if self.global_head else None,
                P.get("global_mlp.pose_fcs.0.bias") if self.global_head else None, B, J, A["pose"])
        return A["pose"]

    def _chain_fwd(self, B, G, g_ld, F, f_ld, C, H, hb, HG, Whh):
        """one propagation-unit layer over the J joints; G is completed IN PLACE to the full gate pre-activations
        (x-side term + recurrent term) so the backward can re-derive every gate from it"""
        J, be = self.J, self.be
        for t in range(J):
            if t > 0:   # gates_t += (sigmoid(F_t) * h_{t-1}) . W_hh^T ; operand rows b live at HG[b*J + t]
                gt_view = _offset(G, t * g_ld)
                self._gemm(HG.at(t), J * PUH, B, PUH, Whh, PUH, 4 * PUH, resid=gt_view, resid_ld=J * g_ld,
                           out_f32=gt_view, ldo=J * g_ld)
            be.pu_cell_fwd(G, J * g_ld, g_ld, F, J * f_ld, f_ld, C, H, None if hb is None else hb.hi,
                           None if hb is None else hb.lo, HG.hi, HG.lo, t, J, B)

    # ------------------------------------------------------------------------------------------ loss
    def loss_and_grad(self, gt):
        """total loss of the reference's backward_AutoEncoder on the last forward; fills the dpose scratch"""
        B = self.batch
        drop_first = not self.global_head
        self.be.pose_loss(self.A["pose"], gt, B, self.nj, KINEMATIC_PARENTS[self.preset], drop_first,
                          self.lambda_mpjpe, self.lambda_cos_sim, self.loss, self.S["dpose"], self.scr)
        return self.loss

    # ------------------------------------------------------------------------------------------ backward
    def backward(self, dpose=None, on_stage=None):
        """gradients of every parameter for the last forward, given d loss / d pose (default: loss_and_grad's).
        on_stage(i, start, end) is called as soon as flat_grad[start:end] is final (in stream order)."""
        be, P, A, S, WT, J, g = self.be, self.P, self.A, self.S, self.WT, self.J, self.grad
        B = self.batch
        M, R, RJ, live = B * TOK, B * self.n_hm, B * J, self.live
        rm = S["rm"]
        v = "pos_heatmap_encoder.vit."
        pu = "skel_sequential_layer.lstm_custom.layers."
        if dpose is None:
            dpose = S["dpose"]
        stage = [0]

        def done():
            if on_stage is not None:
                cb = getattr(be, "callback", None)       # recorded on the tape when a step is being recorded
                if cb is not None:
                    cb(on_stage, stage[0], *self.stages[stage[0]])
                else:
                    on_stage(stage[0], *self.stages[stage[0]])
            stage[0] += 1
        # ---- head
        gh = self.global_head
        be.head_bwd(dpose, A["E"], 2 * PUX, A["skel"], P["pose_mlp.pose_fcs.0.weight"],
                    P["global_mlp.pose_fcs.0.weight"] if gh else None, B, J, S["dE"], 2 * PUX, S["dSkel"],
                    g["pose_mlp.pose_fcs.0.weight"], g["pose_mlp.pose_fcs.0.bias"],
                    g["global_mlp.pose_fcs.0.weight"] if gh else None, g["global_mlp.pose_fcs.0.bias"] if gh else None,
                    self.scr)
        done()
        # ---- propagation chain, layer 1 then layer 0 (BPTT over the joints)
        dFG1 = S["dFG1"]
        self._chain_bwd(B, A["FG1"][:, PUH:], 5 * PUH, A["FG1"], 5 * PUH, A["C1"], A["skel"], S["dSkel"],
                        dFG1[:, PUH:], 5 * PUH, dFG1, 5 * PUH, WT["hh1"])
        bt = S["bias_tmp"]
        self._tsplit(dFG1, RJ, 5 * PUH, 5 * PUH, rm, bias=bt)
        be.copy(g[pu + "1.x2f.bias"], bt[:PUH])
        be.copy(g[pu + "1.x2h.bias"], bt[PUH:])
        be.copy(g[pu + "1.h2h.bias"], bt[PUH:])
        self._dw(rm, 5 * PUH, A["h0b"], PUH, PUH, PUH, RJ, g[pu + "1.x2f.weight"])
        self._dw(_cols(rm, PUH), 5 * PUH, A["h0b"], PUH, 4 * PUH, PUH, RJ, g[pu + "1.x2h.weight"])
        self._dw(_cols(rm, PUH), 5 * PUH, A["HG1"], PUH, 4 * PUH, PUH, RJ, g[pu + "1.h2h.weight"])
        self._gemm(rm, 5 * PUH, RJ, 5 * PUH, WT["cat1"], 5 * PUH, PUH, out_f32=S["dH0"], ldo=PUH)
        dG0, dF0, dE = S["dG0"], S["dF0"], S["dE"]
        self._chain_bwd(B, A["G0"], 4 * PUH, A["F0"], PUH + PUX, A["C0"], A["H0"], S["dH0"], dG0, 4 * PUH, dF0, PUH + PUX,
                        WT["hh0"])
        self._tsplit(dG0, RJ, 4 * PUH, 4 * PUH, rm, bias=g[pu + "0.x2h.bias"])
        be.copy(g[pu + "0.b2h.bias"], g[pu + "0.x2h.bias"])
        be.copy(g[pu + "0.h2h.bias"], g[pu + "0.x2h.bias"])
        self._dw(rm, 4 * PUH, A["xb"], 2 * PUX, 4 * PUH, PUX, RJ, g[pu + "0.x2h.weight"])                 # x half of [x | b']
        self._dw(rm, 4 * PUH, _cols(A["xb"], PUX), 2 * PUX, 4 * PUH, PUX, RJ, g[pu + "0.b2h.weight"])      # gated-bridge half
        self._dw(rm, 4 * PUH, A["HG0"], PUH, 4 * PUH, PUH, RJ, g[pu + "0.h2h.weight"])
        # dE += dG0 . [x2h | b2h]   (columns [:256] d x, [256:] d b')
        self._gemm(rm, 4 * PUH, RJ, 4 * PUH, WT["xb0"], 4 * PUH, 2 * PUX, resid=dE, resid_ld=2 * PUX, out_f32=dE,
                   ldo=2 * PUX)
        be.pu_bridge_gate_bwd(dE, 2 * PUX, A["F0"], PUH + PUX, PUH, A["E"], PUX, RJ, dF0, PUH + PUX)
        self._tsplit(dF0, RJ, PUH + PUX, PUH + PUX, rm, bias=g[pu + "0.x2f.bias"])
        self._dw(rm, PUH + PUX, A["xb"], 2 * PUX, PUH + PUX, PUX, RJ, g[pu + "0.x2f.weight"])
        self._gemm(rm, PUH + PUX, RJ, PUH + PUX, WT["x2f0"], PUH + PUX, PUX, resid=dE, resid_ld=2 * PUX, out_f32=dE,
                   ldo=2 * PUX)
        done()
        # ---- FC encoders (limb first: it has no upstream), block 3 -> 1
        for e, enc in (("r", "rot_heatmap_encoder."), ("p", "pos_heatmap_encoder.")):
            da, dn = S["db"], S["da"]            # ping-pong so that the (R x 16384) d fin lands in the big buffer
            be.regroup_gather(dE, 2 * PUX, 0 if e == "p" else PUX, B, J, EMB, da)
            for i in (2, 1, 0):
                n = (2048, 512, EMB)[i]
                pre = "%sfc%d." % (enc, i + 1)
                y = A["%sy%d" % (e, i + 1)]
                mean, rstd, scale, shift = (A["%s%s%d" % (e, s, i + 1)] for s in ("mean", "rstd", "scale", "shift"))
                be.bn_bwd(da, y, R, n, scale, shift, mean, rstd, g[pre + "bn.weight"], g[pre + "bn.bias"], self.scr)
                self._tsplit(da, R, n, n, rm, bias=g[pre + "fc.bias"])
                if i == 0:
                    x_in, k = (A["fin"], 16 * HID) if e == "p" else (A["a_limb"], 8192)
                else:
                    x_in, k = A["%sa%d" % (e, i)], (2048, 512)[i - 1]
                self._dw(rm, n, x_in, k, n, k, R, g[pre + "fc.weight"])
                if i > 0 or e == "p":
                    self._gemm(rm, n, R, n, WT["%sfc%d" % (e, i + 1)], n, k, out_f32=dn, ldo=k)
                    da, dn = dn, da
            done()
        # ---- final LayerNorm: d fin (R x 16384) == (B*live x 1024) compact rows -> dH (dummy rows get zero)
        dH = S["dH"]
        be.zero(dH)
        be.layernorm_bwd(da, A["h_in"][3], P[v + "layernorm.weight"], B, TOK, live, LN_EPS, dH, 0,
                         g[v + "layernorm.weight"], g[v + "layernorm.bias"], self.scr)
        done()
        # ---- ViT layers
        dA = S["dA"]
        for l in (2, 1, 0):
            p = v + "encoder.layer.%d." % l
            h_in, h_mid = A["h_in"][l], A["h_mid"][l]
            ln1, qk, vt, ctx, ln2, u, gl = (A[n][l] for n in ("ln1", "qk", "vt", "ctx", "ln2", "u", "g"))
            # MLP: h_out = h_mid + down(gelu(up(ln2)))
            self._tsplit(dH, M, HID, HID, rm, bias=g[p + "output.dense.bias"])
            self._dw(rm, HID, gl, MLPD, HID, MLPD, M, g[p + "output.dense.weight"])
            self._gemm(rm, HID, M, HID, WT["down%d" % l], HID, MLPD, out_f32=dA, ldo=MLPD)
            # d u = d g * gelu'(u), its bf16 copies and the bias gradient in one pass (d u is never stored in fp32)
            self._tsplit(dA, M, MLPD, MLPD, rm, bias=g[p + "intermediate.dense.bias"], gelu_u=u)
            self._dw(rm, MLPD, ln2, HID, MLPD, HID, M, g[p + "intermediate.dense.weight"])
            self._gemm(rm, MLPD, M, MLPD, WT["up%d" % l], MLPD, HID, out_f32=dA, ldo=HID)
            be.layernorm_bwd(dA, h_mid, P[p + "layernorm_after.weight"], B, TOK, TOK, LN_EPS, dH, 1,
                             g[p + "layernorm_after.weight"], g[p + "layernorm_after.bias"], self.scr)
            # attention block: h_mid = h_in + o(attn(ln1))
            self._tsplit(dH, M, HID, HID, rm, bias=g[p + "attention.output.dense.bias"])
            self._dw(rm, HID, ctx, HID, HID, HID, M, g[p + "attention.output.dense.weight"])
            dctx = S["dctx"]
            self._gemm(rm, HID, M, HID, WT["o%d" % l], HID, HID, out=dctx, ldo=HID)
            if self._alloc_fused:                            # dA <- d[Q | K | V]  (M x 3072)
                be.attn_dsum(ctx.hi, ctx.lo, dctx.hi, dctx.lo, M, S["dsum"])
                be.attention_bwd(qk.hi, vt.hi, dctx.hi, A["lse"][l], S["dsum"], dA, B)
            else:
                self._attention_bwd(B, qk, vt, dctx, dA)
            self._tsplit(dA, M, 3 * HID, 3 * HID, rm, bias=S["dpos"])     # (3072,) sums into a spare buffer
            for q, n in enumerate(("query", "key", "value")):
                be.copy(g[p + "attention.attention.%s.bias" % n], S["dpos"].view(-1)[q * HID:(q + 1) * HID])
            for q, n in enumerate(("query", "key", "value")):
                self._dw(_cols(rm, q * HID), 3 * HID, ln1, HID, HID, HID, M, g[p + "attention.attention.%s.weight" % n])
            self._gemm(rm, 3 * HID, M, 3 * HID, WT["qkv%d" % l], 3 * HID, HID, out_f32=dA, ldo=HID)
            be.layernorm_bwd(dA, h_in, P[p + "layernorm_before.weight"], B, TOK, TOK, LN_EPS, dH, 1,
                             g[p + "layernorm_before.weight"], g[p + "layernorm_before.bias"], self.scr)
            done()
        # ---- embeddings: hidden[b, t] = patch_gemm + bias + pos_perm[t] (t < live) | mask_token + pos_perm[t]
        Ml = B * live
        self._tsplit(dH, Ml, HID, HID, rm, rows_in=TOK, rows_out=live, bias=g[v + "embeddings.patch_embeddings.projection.bias"])
        self._dw(rm, HID, A["a_patch"], 256, HID, 256, Ml, g[v + "embeddings.patch_embeddings.projection.weight"])
        be.colsum(dH, B, TOK * HID, TOK * HID, 0, 0, S["dpos"], self.scr)      # sum over frames per token
        be.embed_grads(S["dpos"], self.grid, self.n_hm, g[v + "embeddings.position_embeddings"],
                       g[v + "embeddings.mask_token"])
        done()
        return self.grad

    def _chain_bwd(self, B, G, g_ld, F, f_ld, C, H, dOut, dG, dg_ld, dF, df_ld, WhhT):
        J, be, S = self.J, self.be, self.S
        if self.persistent_bptt:
            for b0 in range(0, B, 1024):              # 32 CTAs per 256 frames must be co-resident
                bc = min(1024, B - b0)
                be.pu_chain_bwd(WhhT.hi, WhhT.lo, _offset(G, b0 * J * g_ld), J * g_ld, g_ld, _offset(F, b0 * J * f_ld), J * f_ld,
                                f_ld, C[b0 * J:], H[b0 * J:], dOut[b0 * J:], _offset(dG, b0 * J * dg_ld), J * dg_ld, dg_ld,
                                _offset(dF, b0 * J * df_ld), J * df_ld, df_ld, S["dgx"].hi, S["dgx"].lo, S["counters"], bc, J,
                                self.precision)
            return
        for t in range(J - 1, -1, -1):
            be.pu_cell_bwd(G, J * g_ld, g_ld, F, J * f_ld, f_ld, C, H, dOut, S["dhg"], S["dc"], dG, J * dg_ld, dg_ld,
                           dF, J * df_ld, df_ld, S["dgp"].hi, S["dgp"].lo, t, J, B)
            if t > 0:   # d(sigmoid(F_t) * h_{t-1}) = dgates_t . W_hh
                self._gemm(S["dgp"], 4 * PUH, B, 4 * PUH, WhhT, 4 * PUH, PUH, out_f32=S["dhg"], ldo=PUH)

    def _attention_bwd(self, B, qk, vt, dctx, dqkv):
        """backward of softmax attention for all (frame, head) pairs, recomputing the probabilities:
        S = QK^T/sqrt(d); P = softmax(S); dV = P^T dctx; dP = dctx V^T; dS = P (dP - rowsum(P dP)) / sqrt(d);
        dQ = dS K; dK = dS^T Q  (reference forward: model/modeling_vit.py:233-252).  dqkv: fp32 (M x 3072)."""
        be, S = self.be, self.S
        scale = 1.0 / math.sqrt(HD)
        Bc = min(B, self.attn_chunk)
        for b0 in range(0, B, Bc):
            nb = min(Bc, B - b0)
            G = nb * HEADS
            r0 = b0 * TOK
            qk_c, dctx_c = qk.at(r0), dctx.at(r0)
            vt_c = vt.at(b0 * HEADS * HD)
            qgrp = (HEADS, HD, nb, TOK * 2 * HID)
            # scores (as the unfused forward path): S[g] = scale * Q_g K_g^T
            be.gemm(qk_c.hi, qk_c.lo, qk_c.at(0, HID).hi, qk_c.at(0, HID).lo, TOK, TOK, HD, groups=G, a_group=qgrp,
                    b_group=qgrp, a_rows=TOK, b_rows=TOK, lda=2 * HID, ldb=2 * HID, precision=self.precision,
                    alpha=scale, out_f32=S["Sc"], ldo=TOK, group_rows=TOK)
            # V row-major per (frame, head) from the stored V^T; dP[g] = dctx_g V_g^T
            self._tbf16(vt_c, HD, TOK, TOK, S["V"], HD, groups=(G, HD * TOK, 1, 0), dst_groups=(TOK * HD, 0), pad=HD)
            cgrp = (HEADS, HD, nb, TOK * HID)
            be.gemm(dctx_c.hi, dctx_c.lo, S["V"].hi, S["V"].lo, TOK, TOK, HD, groups=G, a_group=cgrp,
                    b_group=(HEADS, TOK * HD, nb, HEADS * TOK * HD), a_rows=TOK, b_rows=TOK, lda=HID, ldb=HD,
                    precision=self.precision, out_f32=S["dP"], ldo=TOK, group_rows=TOK)
            be.softmax_bwd(S["Sc"], S["dP"], G * TOK, TOK, scale, S["P"].hi, S["P"].lo, S["dS"].hi, S["dS"].lo)
            sq = (G, TOK * TOK, 1, 0)
            self._tbf16(S["P"], TOK, TOK, TOK, S["PT"], TOK, groups=sq, dst_groups=(TOK * TOK, 0))
            self._tbf16(S["dS"], TOK, TOK, TOK, S["dST"], TOK, groups=sq, dst_groups=(TOK * TOK, 0))
            tgrp = (HD * TOK, HEADS * HD * TOK)
            self._tbf16(dctx_c, TOK, HD, HID, S["dctxT"], TOK, groups=cgrp, dst_groups=tgrp)
            self._tbf16(qk_c, TOK, HD, 2 * HID, S["QT"], TOK, groups=qgrp, dst_groups=tgrp)
            self._tbf16(qk_c.at(0, HID), TOK, HD, 2 * HID, S["KT"], TOK, groups=qgrp, dst_groups=tgrp)
            out = _offset(dqkv, r0 * 3 * HID)
            sgrp = (HEADS, TOK * TOK, nb, HEADS * TOK * TOK)
            hgrp = (HEADS, HD * TOK, nb, HEADS * HD * TOK)
            for src, other, col in ((S["dS"], S["KT"], 0), (S["dST"], S["QT"], HID), (S["PT"], S["dctxT"], 2 * HID)):
                be.gemm(src.hi, src.lo, other.hi, other.lo, TOK, HD, TOK, groups=G, a_group=sgrp, b_group=hgrp,
                        a_rows=TOK, b_rows=HD, lda=TOK, ldb=TOK, precision=self.precision, store=STORE_HEAD_MERGE,
                        heads=HEADS, tokens=TOK, out_f32=out, ldo=3 * HID, col_off=col)

    # ------------------------------------------------------------------------------------------ optimiser
    def adamw_step(self, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-4, weight_decay=0.0, grad_scale=1.0):
        """torch.optim.AdamW semantics on every trained parameter, in place (reference model/network.py:72-78,
        options/train_options.py:29-37); re-packs the bf16 operand copies lazily at the next forward."""
        if self.opt_m is None:
            self.opt_m = self.be.empty((self.flat_grad.numel(),), torch.float32)
            self.opt_v = self.be.empty((self.flat_grad.numel(),), torch.float32)
            self.be.zero(self.opt_m)
            self.be.zero(self.opt_v)
        self.opt_step += 1
        self.mutation += 1
        ks = self.order
        self.be.adamw([self.P[k] for k in ks], [self.grad[k] for k in ks],
                      [self.opt_m[self.offsets[k]:] for k in ks], [self.opt_v[self.offsets[k]:] for k in ks],
                      self.opt_step, lr, beta1, beta2, eps, weight_decay, grad_scale)
        self.packed = False

    def _graph_step(self, x, gt):
        """forward + loss + backward through a CUDA graph (torch.cuda.CUDAGraph owns capture and replay; every launch
        inside is ours, enqueued on the capturing stream through the C ABI)"""
        B = x.shape[0]
        self._alloc(B)
        self.be.copy(self.x_in, x)
        self.be.copy(self.gt_in, gt)

        def body():
            self.packed = False
            self.forward(self.x_in)
            self.loss_and_grad(self.gt_in)
            self.backward()
        if self._graph is not None and self._graph_key == B:
            self._graph.replay()
        elif self._graph_warm != B:
            body()                                  # eager: also performs every lazy allocation
            self._graph_warm, self._graph = B, None
        else:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                body()
            self._graph, self._graph_key = g, B
            g.replay()

    def train_step(self, x, gt, lr=1e-3, eps=1e-4, weight_decay=0.0, reducer=None):
        """forward + loss + backward + AdamW, all on the current stream; returns the 3-element loss tensor
        (total, mpjpe term, cos-sim term) without synchronising.  ``reducer`` (ddp.StagedGradAllReduce) sums the
        gradients over the data-parallel ranks stage by stage while the backward is still running."""
        on_stage = None if reducer is None else reducer.on_stage
        if self.use_cuda_graph and reducer is None:
            self._graph_step(x, gt)
            loss = self.loss
        elif not self.use_tape:
            self.forward(x)
            loss = self.loss_and_grad(gt)
            self.backward(on_stage=on_stage)
        else:
            be = self.be
            self._alloc(x.shape[0])
            be.copy(self.x_in, x)
            be.copy(self.gt_in, gt)
            key = (x.shape[0], id(reducer))
            if self._tape is not None and self._tape_key == key and be.can_replay(self._tape):
                be.replay(self._tape)
                loss = self.loss
            else:
                self.packed = False                       # the recorded step must contain the re-pack of the weights
                be.begin_record()
                try:
                    self.forward(self.x_in)
                    loss = self.loss_and_grad(self.gt_in)
                    self.backward(on_stage=on_stage)
                finally:
                    tape = be.end_record()
                self._tape, self._tape_key = tape, key
        if reducer is None:
            self.adamw_step(lr=lr, eps=eps, weight_decay=weight_decay)
        else:
            reducer.finish()
            self.adamw_step(lr=lr, eps=eps, weight_decay=weight_decay, grad_scale=1.0 / reducer.world)
        return loss


def _offset(t, elems):
    """pointer arithmetic on a tensor: a 1-D view starting `elems` elements after t's first element"""
    n = t.untyped_storage().nbytes() // t.element_size() - t.storage_offset() - elems
    return torch.as_strided(t, (n,), (1,), t.storage_offset() + elems)


def _cols(pair, col):
    """columns [col:] of a row-major pair (pointer offset; the caller keeps using the full matrix's row stride)"""
    return Pair(_offset(pair.hi, col), None if pair.lo is None else _offset(pair.lo, col))


def cosine_warmup_lr(step, base_lr, warmup_steps, total_steps):
    """transformers.get_cosine_schedule_with_warmup as the reference uses it (model/network.py:49-52)"""
    if step < warmup_steps:
        return base_lr * step / max(1, warmup_steps)
    prog = (step - warmup_steps) / max(1, total_steps - warmup_steps)
    return base_lr * max(0.0, 0.5 * (1.0 + math.cos(math.pi * prog)))
