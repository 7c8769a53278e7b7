"""TEST INFRASTRUCTURE ONLY -- torch-CPU restatement of every op-level entry point of the C ABI
(include/egotap_b200.h), with the SAME argument meaning as the ctypes backend in egotap_b200/capi.py.

Two uses, both inside tests/:
  * ``OracleBackend`` can be injected into ``egotap_b200.training.TrainEngine`` so the host-side orchestration of the
    training step (which buffer feeds which op, strides, group layouts, split-K, gradient routing) is executed on the
    CPU and compared with torch.autograd on the restated forward (oracle/train_oracle.py);
  * each method is the per-op checker for the GPU op-level tests (run the op through the C ABI on the B200 and here
    on identical inputs).
The product never imports this file: TrainEngine's default backend is the CUDA library and it fails loudly without it.

Operand / epilogue semantics follow egotap_b200/csrc/gemm.cuh (GPU-verified in round 1 through
tests/test_gemm_gpu.py and the forward parity tests); references to the model math cite the reference tree.
"""
import math

import torch
import torch.nn.functional as F

ACT_NONE, ACT_GELU, ACT_LRELU = 0, 1, 2
STORE_ROWMAJOR, STORE_QKV, STORE_JOINT_REGROUP, STORE_HEAD_MERGE = 0, 1, 2, 3
PREC_BF16X3, PREC_BF16 = 0, 1


def flat(t):
    """1-D view of everything from t's first element to the end of its storage (pointer semantics)."""
    n = t.untyped_storage().nbytes() // t.element_size() - t.storage_offset()
    return torch.as_strided(t, (n,), (1,))


def split(x):
    """fp32 -> (hi, lo) bf16 with x ~= hi + lo (csrc/ptx.cuh split_bf16)."""
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi, lo


def store_pair(hi, lo, idx, v):
    """scatter fp32 values v to flat index idx of a bf16 hi (/lo) pair.  (In OracleBackend(exact=True) the "bf16"
    buffers are fp32 tensors: hi then carries the full value and lo zero -- orchestration check without rounding.)"""
    if hi.dtype == torch.float32:
        flat(hi)[idx] = v.float()
        if lo is not None:
            flat(lo)[idx] = 0.0
        return
    h, l = split(v)
    flat(hi)[idx] = h
    if lo is not None:
        flat(lo)[idx] = l


def pair_f32(hi, lo, size, stride):
    x = torch.as_strided(hi, size, stride).float()
    if lo is not None:
        x = x + torch.as_strided(lo, size, stride).float()
    return x


def gelu_erf(x):
    return F.gelu(x)


class OracleBackend:
    """CPU tensors in, CPU tensors out; every method mirrors one extern "C" entry (same name minus the prefix)."""

    name = "oracle"
    device = torch.device("cpu")

    def __init__(self, exact=False):
        """exact=True: operand "pairs" are stored as fp32 (no bf16 rounding anywhere), which isolates the host-side
        orchestration from the precision design when comparing with autograd"""
        self.launches = 0
        self.exact = exact

    # ------------------------------------------------------------------ memory helpers
    def empty(self, shape, dtype=torch.float32):
        # poisoned allocation so that reads of never-written memory show up in the comparison
        if dtype == torch.float32:
            return torch.full(shape if isinstance(shape, tuple) else (shape,), float("nan"))
        if dtype == torch.bfloat16:
            return torch.full(shape if isinstance(shape, tuple) else (shape,), float("nan"),
                              dtype=torch.float32 if self.exact else torch.bfloat16)
        return torch.zeros(shape, dtype=dtype)

    def zero(self, t):
        assert t.is_contiguous()
        t.zero_()

    def copy(self, dst, src):
        assert dst.is_contiguous() and src.is_contiguous() and dst.numel() == src.numel()
        dst.view(-1).copy_(src.view(-1))

    def add3(self, a, b, c, out, n):
        self.launches += 1
        v = flat(a)[:n] + flat(b)[:n]
        if c is not None:
            v = v + flat(c)[:n]
        flat(out)[:n] = v

    # ------------------------------------------------------------------ tcgen05 GEMM (csrc/gemm.cuh)
    def gemm(self, a_hi, a_lo, b_hi, b_lo, M, N, K, *, groups=1, a_group=(1, 0, 1, 0), b_group=(1, 0, 1, 0),
             a_rows=None, b_rows=None, lda=None, ldb=None, precision=PREC_BF16X3, variant=-1, tn=False, **epi):
        assert K % 64 == 0 and N % 32 == 0, (N, K)
        if precision == PREC_BF16X3:
            assert a_lo is not None and b_lo is not None
        else:
            a_lo = b_lo = None
        self.launches += 1

        def operand(hi, lo, ld, rows, group, want_rows):
            g0c, g0s, g1c, g1s = group
            g0c, g1c = max(g0c, 1), max(g1c, 1)
            if g0s <= 0:
                g0s = rows * ld
            if g1s <= 0:
                g1s = g0s * g0c
            assert g0c * g1c == groups, "operand group counts %s do not cover %d groups" % (group, groups)
            x = pair_f32(hi, lo, (g1c, g0c, rows, K), (g1s, g0s, ld, 1)).reshape(groups, rows, K)
            if rows < want_rows:            # TMA zero-fills rows beyond the tensor map's extent
                x = torch.cat([x, x.new_zeros(groups, want_rows - rows, K)], 1)
            return x[:, :want_rows]
        if tn:
            # row-major operands, contraction along the rows, group = chunk of K rows (EGOTAP_GEMM_TN); rows beyond a_rows are zero
            assert a_rows == b_rows and N % 256 == 0 and M % 64 == 0

            def chunks(hi, lo, ld, cols):
                x = pair_f32(hi, lo, (a_rows, cols), (ld, 1))
                pad = groups * K - a_rows
                if pad > 0:
                    x = torch.cat([x, x.new_zeros(pad, cols)], 0)
                return x[:groups * K].reshape(groups, K, cols).transpose(1, 2)      # (groups, cols, K)
            A, B = chunks(a_hi, a_lo, lda, M), chunks(b_hi, b_lo, ldb, N)
        else:
            A = operand(a_hi, a_lo, lda or K, a_rows or M, a_group, M)
            B = operand(b_hi, b_lo, ldb or K, b_rows or N, b_group, N)
        # bf16x3 on the GPU is Ah*Bh + Ah*Bl + Al*Bh (the Al*Bl term, 2^-16 relative, is dropped); here the full
        # (Ah + Al)(Bh + Bl) product in fp32 stands for it
        acc = torch.matmul(A, B.transpose(1, 2))
        v = acc * float(epi.pop("alpha", 1.0))
        scale, bias, resid = epi.pop("scale", None), epi.pop("bias", None), epi.pop("resid", None)
        act = int(epi.pop("act", 0))
        resid_ld, resid_mod = int(epi.pop("resid_ld", 0)), int(epi.pop("resid_mod", 0))
        rows_in, rows_out = int(epi.pop("rows_in", 0)), int(epi.pop("rows_out", 0))
        group_rows = int(epi.pop("group_rows", 0))
        out_f32, out_hi, out_lo = epi.pop("out_f32", None), epi.pop("out_hi", None), epi.pop("out_lo", None)
        ldo, col_off, store = int(epi.pop("ldo", 0)) or N, int(epi.pop("col_off", 0)), int(epi.pop("store", 0))
        qk_cols, tokens = int(epi.pop("qk_cols", 0)), int(epi.pop("tokens", 0))
        vt_hi, vt_lo = epi.pop("vt_hi", None), epi.pop("vt_lo", None)
        J, heads = int(epi.pop("J", 0)), int(epi.pop("heads", 0))
        assert not epi, "unknown epilogue fields %s" % sorted(epi)
        assert ldo % 8 == 0 and col_off % 32 == 0
        if scale is not None:
            v = v * flat(scale)[:N]
        if bias is not None:
            v = v + flat(bias)[:N]
        if act == ACT_GELU:
            v = gelu_erf(v)
        elif act == ACT_LRELU:
            v = F.leaky_relu(v, 0.2)
        g = torch.arange(groups).view(groups, 1).expand(groups, M)
        m = torch.arange(M).view(1, M).expand(groups, M)
        col_shift = torch.zeros(groups, M, dtype=torch.long)
        if store == STORE_JOINT_REGROUP:
            frame, rem = m // (2 * J), m % (2 * J)
            orow = frame * J + rem % J
            col_shift = (rem // J) * N
        elif store == STORE_HEAD_MERGE:
            orow = (g // heads) * tokens + m
            col_shift = (g % heads) * N
        else:
            orow = (m // rows_in) * rows_out + m % rows_in if rows_in > 0 else m
            orow = orow + g * group_rows
        n = torch.arange(N).view(1, 1, N)
        if resid is not None:
            rrow = m % resid_mod if resid_mod > 0 else orow
            ridx = rrow.unsqueeze(-1) * resid_ld + n + col_off + col_shift.unsqueeze(-1)
            v = v + flat(resid)[ridx]
        idx = orow.unsqueeze(-1) * ldo + n + col_off + col_shift.unsqueeze(-1)
        if store == STORE_QKV:
            vcols = N - qk_cols
            frame, tok = m // tokens, m % tokens
            vidx = (frame.unsqueeze(-1) * vcols + (n[..., qk_cols:] - qk_cols)) * tokens + tok.unsqueeze(-1)
            store_pair(vt_hi, vt_lo, vidx, v[..., qk_cols:])
            v, idx = v[..., :qk_cols], idx[..., :qk_cols]
        if out_f32 is not None:
            flat(out_f32)[idx] = v
        if out_hi is not None:
            store_pair(out_hi, out_lo, idx, v)

    # ------------------------------------------------------------------ round-1 ops (kernels.cu, attention.cu)
    def split2d(self, src, rows, cols, src_ld, hi, lo, dst_ld):
        self.launches += 1
        x = torch.as_strided(src, (rows, cols), (src_ld, 1))
        idx = torch.arange(rows).view(-1, 1) * dst_ld + torch.arange(cols).view(1, -1)
        store_pair(hi, lo, idx, x)

    def ingest(self, x, J, p_hi, p_lo, l_hi, l_lo):
        """reference model/net_architecture.py:688-694, :375-383, modeling_vit.py:195 (see kernels.cu ingest_kernel)"""
        self.launches += 1
        B = x.shape[0]
        n = 2 * J
        pos = x[:, :n]
        patches = pos.reshape(B, n, 4, 16, 4, 16).permute(0, 1, 2, 4, 3, 5).reshape(B * n * 16, 256)
        rot = x[:, n:].reshape(B, 2, 2, J, 4096).transpose(2, 3).reshape(B * n, 8192)   # (view, j, d, pix)
        store_pair(p_hi, p_lo, torch.arange(patches.numel()).view(patches.shape), patches)
        store_pair(l_hi, l_lo, torch.arange(rot.numel()).view(rot.shape), rot)

    def fill_dummy(self, hidden, dummy, B, tokens, live):
        self.launches += 1
        nd = tokens - live
        if nd:
            h = torch.as_strided(hidden, (B, nd, 1024), (tokens * 1024, 1024, 1), live * 1024)
            h.copy_(torch.as_strided(dummy, (nd, 1024), (1024, 1)).unsqueeze(0).expand(B, nd, 1024))

    def pos_permute(self, pos, mask_token, grid, n_hm, pos_perm, dummy):
        self.launches += 1
        perm = token_perm(grid)
        p = flat(pos)[:576 * 1024].view(576, 1024)[perm[:grid * grid * 16]]
        flat(pos_perm)[:p.numel()] = p.reshape(-1)
        d = p[n_hm * 16:] + flat(mask_token)[:1024]
        flat(dummy)[:d.numel()] = d.reshape(-1)

    def layernorm(self, x, w, b, frames, rows_in, rows_out, eps, out_hi, out_lo, out_f32):
        self.launches += 1
        xi = torch.as_strided(x, (frames, rows_out, 1024), (rows_in * 1024, 1024, 1)).reshape(-1, 1024)
        y = F.layer_norm(xi, (1024,), flat(w)[:1024], flat(b)[:1024], eps)
        idx = torch.arange(y.numel()).view(y.shape)
        if out_hi is not None:
            store_pair(out_hi, out_lo, idx, y)
        if out_f32 is not None:
            flat(out_f32)[idx] = y

    def attention(self, qk_hi, qk_lo, vt_hi, vt_lo, ctx_hi, ctx_lo, frames, precision):
        """reference model/modeling_vit.py:233-252; layouts as written by the QKV GEMM's STORE_QKV epilogue"""
        self.launches += 1
        if precision != PREC_BF16X3:
            qk_lo = vt_lo = None
        qk = pair_f32(qk_hi, qk_lo, (frames, 576, 2048), (576 * 2048, 2048, 1))
        vt = pair_f32(vt_hi, vt_lo, (frames, 8, 128, 576), (8 * 128 * 576, 128 * 576, 576, 1))
        q = qk[..., :1024].view(frames, 576, 8, 128).permute(0, 2, 1, 3)
        k = qk[..., 1024:].view(frames, 576, 8, 128).permute(0, 2, 1, 3)
        p = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(128.0), -1)
        if precision != PREC_BF16X3 and not self.exact:
            p = p.to(torch.bfloat16).float()          # P is a bf16 MMA operand in the 1-MMA mode
        ctx = (p @ vt.transpose(-1, -2)).permute(0, 2, 1, 3).reshape(frames * 576, 1024)
        store_pair(ctx_hi, ctx_lo, torch.arange(ctx.numel()).view(ctx.shape), ctx)

    def pu_bridge_gate(self, f, f_ld, f_col, e, e_ld, X, rows, hi, lo):
        """reference model/custom_cells.py:99-102"""
        self.launches += 1
        g = torch.as_strided(f, (rows, X), (f_ld, 1), f.storage_offset() + f_col)
        b = torch.as_strided(e, (rows, X), (e_ld, 1), e.storage_offset() + X)
        idx = torch.arange(rows).view(-1, 1) * e_ld + X + torch.arange(X).view(1, -1)
        store_pair(hi, lo, idx, torch.sigmoid(g) * b)

    def head(self, e, e_ld, skel, Wp, bp, Wg, bg, frames, J, pose):
        """reference model/net_architecture.py:732-751"""
        self.launches += 1
        pe = torch.as_strided(e, (frames * J, 256), (e_ld, 1))
        sk = flat(skel)[:frames * J * 512].view(frames * J, 512)
        w = flat(Wp)[:3 * 768].view(3, 768)
        out = (torch.cat([pe, sk], 1) @ w.t() + flat(bp)[:3]).view(frames, J, 3)
        if Wg is not None:
            o = sk.reshape(frames, J * 512) @ flat(Wg)[:6 * J * 512].view(6, J * 512).t() + flat(bg)[:6]
            out = torch.cat([out + o[:, None, :3], o[:, None, 3:]], 1)
        flat(pose)[:out.numel()] = out.reshape(-1)

    # ------------------------------------------------------------------ training ops (csrc/train.cu)
    @staticmethod
    def _src_rows(rows, rows_in, rows_out):
        r = torch.arange(rows)
        return (r // rows_out) * rows_in + r % rows_out if rows_out > 0 else r

    def transpose_split(self, src, rows, cols, src_ld, rows_in, rows_out, rm_hi, rm_lo, rm_ld, t_hi, t_lo, t_ld, pad_rows,
                        gelu_u=None, colsum_out=None, scratch=None):
        """fp32 (rows x cols; logical row r lives at source row (r/rows_out)*rows_in + r%rows_out when rows_out > 0)
        -> row-major bf16 pair rm[r][c] and/or transposed pair t[c][r], t[c][rows..pad_rows) = 0.
        gelu_u: values are first multiplied by gelu'(gelu_u) (same layout as src); colsum_out[c] = sum_r value[r][c]."""
        self.launches += 1
        sr = self._src_rows(rows, rows_in, rows_out)
        idx = sr.view(-1, 1) * src_ld + torch.arange(cols).view(1, -1)
        x = flat(src)[idx]
        if gelu_u is not None:
            u = flat(gelu_u)[idx].double()
            d = 0.5 * (1 + torch.erf(u / math.sqrt(2.0))) + u * torch.exp(-0.5 * u * u) / math.sqrt(2 * math.pi)
            x = (x.double() * d).float()
        if colsum_out is not None:
            flat(colsum_out)[:cols] = x.double().sum(0).float()
        if rm_hi is not None:
            store_pair(rm_hi, rm_lo, torch.arange(rows).view(-1, 1) * rm_ld + torch.arange(cols).view(1, -1), x)
        if t_hi is not None:
            assert pad_rows >= rows and t_ld >= pad_rows
            xt = torch.cat([x.t(), x.new_zeros(cols, pad_rows - rows)], 1)
            store_pair(t_hi, t_lo, torch.arange(cols).view(-1, 1) * t_ld + torch.arange(pad_rows).view(1, -1), xt)

    def transpose_bf16(self, s_hi, s_lo, rows, cols, s_ld, g0c, s_g0s, g1c, s_g1s, d_hi, d_lo, d_ld, d_g0s, d_g1s, pad_rows):
        """batched bf16 transpose: d[g1][g0][c][r] = s[g1][g0][r][c]; d[..][c][rows..pad_rows) = 0"""
        self.launches += 1
        assert pad_rows >= rows and d_ld >= pad_rows
        for src, dst in ((s_hi, d_hi), (s_lo, d_lo)):
            if src is None or dst is None:
                continue
            x = torch.as_strided(src, (g1c, g0c, rows, cols), (s_g1s, s_g0s, s_ld, 1))
            xt = torch.cat([x.transpose(2, 3), x.new_zeros(g1c, g0c, cols, pad_rows - rows)], 3)
            torch.as_strided(dst, (g1c, g0c, cols, pad_rows), (d_g1s, d_g0s, d_ld, 1)).copy_(xt)
        if s_lo is None and d_lo is not None:
            torch.as_strided(d_lo, (g1c, g0c, cols, pad_rows), (d_g1s, d_g0s, d_ld, 1)).zero_()

    def colsum(self, src, rows, cols, ld, rows_in, rows_out, out, scratch=None):
        """out[c] = sum_r src[srow(r)][c]"""
        self.launches += 2
        sr = self._src_rows(rows, rows_in, rows_out)
        x = flat(src)[sr.view(-1, 1) * ld + torch.arange(cols).view(1, -1)]
        flat(out)[:cols] = x.double().sum(0).float()

    def reduce_partials(self, partials, G, n, out):
        self.launches += 1
        flat(out)[:n] = flat(partials)[:G * n].view(G, n).sum(0)

    def attention_lse(self, qk_hi, qk_lo, vt_hi, vt_lo, ctx_hi, ctx_lo, lse, frames, precision):
        """attention() + the per-row log-sum-exp in the exp2 domain: lse[(frame, head, token)] = log2(sum_k exp(s_k))"""
        self.attention(qk_hi, qk_lo, vt_hi, vt_lo, ctx_hi, ctx_lo, frames, precision)
        if precision != PREC_BF16X3:
            qk_lo = None
        qk = pair_f32(qk_hi, qk_lo, (frames, 576, 2048), (576 * 2048, 2048, 1)).double()
        q = qk[..., :1024].view(frames, 576, 8, 128).permute(0, 2, 1, 3)
        k = qk[..., 1024:].view(frames, 576, 8, 128).permute(0, 2, 1, 3)
        s = q @ k.transpose(-1, -2) / math.sqrt(128.0)
        flat(lse)[:frames * 8 * 576] = (torch.logsumexp(s, -1) / math.log(2.0)).float().reshape(-1)

    def attn_dsum(self, ctx_hi, ctx_lo, dctx_hi, dctx_lo, rows, dsum):
        """D[(frame, head, token)] = sum_d dctx * ctx over the head's 128 columns"""
        self.launches += 1
        frames = rows // 576
        o = pair_f32(ctx_hi, ctx_lo, (frames, 576, 8, 128), (576 * 1024, 1024, 128, 1)).double()
        g = pair_f32(dctx_hi, dctx_lo, (frames, 576, 8, 128), (576 * 1024, 1024, 128, 1)).double()
        flat(dsum)[:frames * 8 * 576] = (o * g).sum(-1).permute(0, 2, 1).float().reshape(-1)

    def attention_bwd(self, qk_hi, vt_hi, dctx_hi, lse, dsum, dqkv, frames):
        """fused backward, bf16 operands: P = exp2(s*c - lse), dS = P (dP - D) / sqrt(128); dqkv (frames*576, 3072) fp32 =
        [dQ | dK | dV].  P and dS are bf16 MMA operands in the kernel (rounded here unless ``exact``)."""
        self.launches += 1
        qk = pair_f32(qk_hi, None, (frames, 576, 2048), (576 * 2048, 2048, 1)).double()
        vt = pair_f32(vt_hi, None, (frames, 8, 128, 576), (8 * 128 * 576, 128 * 576, 576, 1)).double()
        do = pair_f32(dctx_hi, None, (frames, 576, 8, 128), (576 * 1024, 1024, 128, 1)).double().permute(0, 2, 1, 3)
        q = qk[..., :1024].view(frames, 576, 8, 128).permute(0, 2, 1, 3)
        k = qk[..., 1024:].view(frames, 576, 8, 128).permute(0, 2, 1, 3)
        v = vt.transpose(-1, -2)
        L = flat(lse)[:frames * 8 * 576].view(frames, 8, 576, 1).double()
        D = flat(dsum)[:frames * 8 * 576].view(frames, 8, 576, 1).double()
        scale = 1.0 / math.sqrt(128.0)
        p = torch.exp2(q @ k.transpose(-1, -2) * (scale / math.log(2.0)) - L)
        ds = p * ((do @ v.transpose(-1, -2)) - D) * scale
        if not self.exact:
            p, ds = p.to(torch.bfloat16).double(), ds.to(torch.bfloat16).double()
        dq, dk, dv = ds @ k, ds.transpose(-1, -2) @ q, p.transpose(-1, -2) @ do
        out = torch.cat([t.permute(0, 2, 1, 3).reshape(frames * 576, 1024) for t in (dq, dk, dv)], 1).float()
        flat(dqkv)[:out.numel()] = out.reshape(-1)

    def gelu_fwd(self, u, n, out_hi, out_lo):
        """reference ACT2FN['gelu'] (exact erf), model/modeling_vit.py:326"""
        self.launches += 1
        store_pair(out_hi, out_lo, torch.arange(n), gelu_erf(flat(u)[:n]))

    def gelu_bwd(self, dg, u, n):
        """in place: dg <- dg * gelu'(u),  gelu'(u) = Phi(u) + u * phi(u)"""
        self.launches += 1
        x = flat(u)[:n].double()
        d = 0.5 * (1 + torch.erf(x / math.sqrt(2.0))) + x * torch.exp(-0.5 * x * x) / math.sqrt(2 * math.pi)
        flat(dg)[:n] = (flat(dg)[:n].double() * d).float()

    def layernorm_bwd(self, dy, x, w, frames, rows_in, rows_out, eps, dx, accumulate, dw, db, scratch=None):
        """backward of layernorm(): dy (frames*rows_out, 1024) compact; x and dx in the rows_in layout.
        dx[in_row] = (accumulate ? dx[in_row] : 0) + dLN/dx ; dw[c] = sum dy*xhat ; db[c] = sum dy"""
        self.launches += 2
        R = frames * rows_out
        sr = self._src_rows(R, rows_in, rows_out)
        idx = sr.view(-1, 1) * 1024 + torch.arange(1024).view(1, -1)
        xi = flat(x)[idx].double()
        g_out = flat(dy)[:R * 1024].view(R, 1024).double()
        mu = xi.mean(1, keepdim=True)
        rstd = 1.0 / torch.sqrt(xi.var(1, unbiased=False, keepdim=True) + eps)
        xhat = (xi - mu) * rstd
        g = g_out * flat(w)[:1024].double()
        d = rstd * (g - g.mean(1, keepdim=True) - xhat * (g * xhat).mean(1, keepdim=True))
        if accumulate:
            flat(dx)[idx] = (flat(dx)[idx].double() + d).float()
        else:
            flat(dx)[idx] = d.float()
        flat(dw)[:1024] = (g_out * xhat).sum(0).float()
        flat(db)[:1024] = g_out.sum(0).float()

    def softmax_bwd(self, S, dP, rows, cols, scale, p_hi, p_lo, ds_hi, ds_lo):
        """S: scaled scores, dP: gradient wrt probabilities (both rows x cols fp32).
        P = softmax(S) ; dS = P * (dP - sum(P*dP)) * scale   (scale = d scores / d (Q.K))"""
        self.launches += 1
        s = flat(S)[:rows * cols].view(rows, cols).double()
        p = torch.softmax(s, 1)
        d = flat(dP)[:rows * cols].view(rows, cols).double()
        ds = p * (d - (p * d).sum(1, keepdim=True)) * scale
        idx = torch.arange(rows * cols).view(rows, cols)
        store_pair(p_hi, p_lo, idx, p.float())
        store_pair(ds_hi, ds_lo, idx, ds.float())

    def bn_stats(self, y, rows, cols, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps,
                 mean, rstd, scale, shift, scratch=None):
        """train-mode BatchNorm1d statistics over the rows (reference model/network_utils.py:123-142;
        torch.nn.BatchNorm1d: biased variance to normalise, unbiased for the running estimate) and the folded
        per-column scale/shift; updates the running buffers in place."""
        self.launches += 2
        yy = flat(y)[:rows * cols].view(rows, cols).double()
        mu = yy.mean(0)
        var = yy.var(0, unbiased=False)
        r = 1.0 / torch.sqrt(var + eps)
        flat(mean)[:cols] = mu.float()
        flat(rstd)[:cols] = r.float()
        sc = flat(gamma)[:cols].double() * r
        flat(scale)[:cols] = sc.float()
        flat(shift)[:cols] = (flat(beta)[:cols].double() - mu * sc).float()
        rm, rv = flat(running_mean)[:cols], flat(running_var)[:cols]
        rm.copy_(((1 - momentum) * rm.double() + momentum * mu).float())
        rv.copy_(((1 - momentum) * rv.double() + momentum * var * rows / max(rows - 1, 1)).float())
        if num_batches_tracked is not None:
            num_batches_tracked += 1

    def bn_apply(self, y, rows, cols, scale, shift, out_hi, out_lo, out_ld, out_f32, f32_ld, J, col_off):
        """a = LeakyReLU_0.2(y * scale + shift); J > 0: regroup row frame*2J + view*J + j -> frame*J + j,
        column += view*cols + col_off (the [left | right] per-joint layout, reference net_architecture.py:699-705)"""
        self.launches += 1
        a = F.leaky_relu(flat(y)[:rows * cols].view(rows, cols) * flat(scale)[:cols] + flat(shift)[:cols], 0.2)
        r = torch.arange(rows)
        c = torch.arange(cols).view(1, -1)
        if J > 0:
            frame, rem = r // (2 * J), r % (2 * J)
            orow = (frame * J + rem % J).view(-1, 1)
            c = c + ((rem // J) * cols).view(-1, 1) + col_off
        else:
            orow = r.view(-1, 1)
            c = c + col_off
        if out_hi is not None:
            store_pair(out_hi, out_lo, orow * out_ld + c, a)
        if out_f32 is not None:
            flat(out_f32)[orow * f32_ld + c] = a

    def bn_bwd(self, da, y, rows, cols, scale, shift, mean, rstd, dgamma, dbeta, scratch=None):
        """in place da -> dy through LeakyReLU and train-mode BatchNorm:
        z = y*scale + shift ; dz = da * (z > 0 ? 1 : 0.2) ; xhat = (y - mean) * rstd
        dgamma = sum dz*xhat ; dbeta = sum dz ; dy = scale * (dz - dbeta/rows - xhat * dgamma/rows)"""
        self.launches += 3
        yy = flat(y)[:rows * cols].view(rows, cols).double()
        sc, sh = flat(scale)[:cols].double(), flat(shift)[:cols].double()
        z = yy * sc + sh
        dz = flat(da)[:rows * cols].view(rows, cols).double() * torch.where(z > 0, 1.0, 0.2)
        xhat = (yy - flat(mean)[:cols].double()) * flat(rstd)[:cols].double()
        dg, dbt = (dz * xhat).sum(0), dz.sum(0)
        flat(dgamma)[:cols] = dg.float()
        flat(dbeta)[:cols] = dbt.float()
        flat(da)[:rows * cols] = (sc * (dz - dbt / rows - xhat * dg / rows)).float().reshape(-1)

    def regroup_gather(self, dE, e_ld, col_off, frames, J, cols, out):
        """out[f*2J + v*J + j][c] = dE[f*J + j][col_off + v*cols + c]  (inverse of bn_apply's regrouped store)"""
        self.launches += 1
        r = torch.arange(frames * 2 * J)
        frame, rem = r // (2 * J), r % (2 * J)
        idx = ((frame * J + rem % J) * e_ld + col_off + (rem // J) * cols).view(-1, 1) + torch.arange(cols).view(1, -1)
        flat(out)[:r.numel() * cols] = flat(dE)[idx].reshape(-1)

    def pu_cell_fwd(self, G, g_rs, g_ts, F_, f_rs, f_ts, C_all, H, h_hi, h_lo, hg_hi, hg_lo, t, J, B):
        """one joint step of a propagation-unit layer (reference model/custom_cells.py:109-120, gate order f,i,g,o):
        gates = G[b*g_rs + t*g_ts + :2048] (already x-side + recurrent); c_prev = C_all[b*J+t-1] (0 at t = 0)
        c -> C_all[b*J+t]; h -> H[b*J+t] (+ bf16 pair); and the NEXT step's pre-gated recurrent operand
        hg[b*J+t+1] = sigmoid(F[b*f_rs + (t+1)*f_ts + :512]) * h   (custom_cells.py:101)."""
        self.launches += 1
        Hd = 512
        b = torch.arange(B)
        g = flat(G)[(b * g_rs + t * g_ts).view(-1, 1) + torch.arange(4 * Hd).view(1, -1)]
        fg, ig, cg, og = g.chunk(4, 1)
        row = (b * J + t).view(-1, 1) * Hd + torch.arange(Hd).view(1, -1)
        c_prev = flat(C_all)[row - Hd] if t > 0 else torch.zeros(B, Hd)
        c = c_prev * torch.sigmoid(fg) + torch.sigmoid(ig) * torch.tanh(cg)
        h = torch.sigmoid(og) * torch.tanh(c)
        flat(C_all)[row] = c
        flat(H)[row] = h
        if h_hi is not None:
            store_pair(h_hi, h_lo, row, h)
        if t + 1 < J:
            f = flat(F_)[(b * f_rs + (t + 1) * f_ts).view(-1, 1) + torch.arange(Hd).view(1, -1)]
            store_pair(hg_hi, hg_lo, row + Hd, torch.sigmoid(f) * h)

    def pu_cell_bwd(self, G, g_rs, g_ts, F_, f_rs, f_ts, C_all, H, dOut, dhg, dc, dG, dg_rs, dg_ts, dF, df_rs, df_ts,
                    dgp_hi, dgp_lo, t, J, B):
        """backward of pu_cell_fwd for step t (walks t = J-1 .. 0):
        dh = dOut[b*J+t] + (t+1 < J ? dhg[b] * sigmoid(F[b,t+1]) : 0)
        dF[b,t+1] = dhg[b] * h_t * sigmoid'(F[b,t+1])   (t+1 < J);  dF[b,0] = 0 (written at t = 0)
        dc_tot = dc[b] + dh * o * (1 - tanh(c_t)^2)
        dgates = [dc_tot*c_prev*f'(fg), dc_tot*tanh(cg)*i'(ig), dc_tot*i*(1-tanh(cg)^2), dh*tanh(c_t)*o'(og)]
        dc[b] <- dc_tot * sigmoid(fg);  dgates -> dG[b*dg_rs + t*dg_ts + :2048] (fp32) and the bf16 pair dgp[b][:2048]
        (dhg = dgates . W_hh is the GEMM the caller issues next; dc is read only for t < J-1)."""
        self.launches += 1
        Hd = 512
        b = torch.arange(B)
        cols = torch.arange(Hd).view(1, -1)
        g = flat(G)[(b * g_rs + t * g_ts).view(-1, 1) + torch.arange(4 * Hd).view(1, -1)].double()
        fg, ig, cg, og = g.chunk(4, 1)
        row = (b * J + t).view(-1, 1) * Hd + cols
        c_t = flat(C_all)[row].double()
        c_prev = flat(C_all)[row - Hd].double() if t > 0 else torch.zeros(B, Hd, dtype=torch.float64)
        h_t = flat(H)[row].double()
        dh = flat(dOut)[row].double()
        if t + 1 < J:
            f_idx = (b * f_rs + (t + 1) * f_ts).view(-1, 1) + cols
            sf = torch.sigmoid(flat(F_)[f_idx].double())
            dhg_ = flat(dhg)[:B * Hd].view(B, Hd).double()
            dh = dh + dhg_ * sf
            flat(dF)[(b * df_rs + (t + 1) * df_ts).view(-1, 1) + cols] = (dhg_ * h_t * sf * (1 - sf)).float()
        if t == 0:
            flat(dF)[(b * df_rs).view(-1, 1) + cols] = 0.0
        sfg, sig, sog = torch.sigmoid(fg), torch.sigmoid(ig), torch.sigmoid(og)
        tcg, tc = torch.tanh(cg), torch.tanh(c_t)
        dc_tot = dh * sog * (1 - tc * tc)
        if t + 1 < J:
            dc_tot = dc_tot + flat(dc)[:B * Hd].view(B, Hd).double()
        dgates = torch.cat([dc_tot * c_prev * sfg * (1 - sfg), dc_tot * tcg * sig * (1 - sig),
                            dc_tot * sig * (1 - tcg * tcg), dh * tc * sog * (1 - sog)], 1).float()
        flat(dc)[:B * Hd] = (dc_tot * sfg).float().reshape(-1)
        flat(dG)[(b * dg_rs + t * dg_ts).view(-1, 1) + torch.arange(4 * Hd).view(1, -1)] = dgates
        store_pair(dgp_hi, dgp_lo, torch.arange(B * 4 * Hd).view(B, 4 * Hd), dgates)

    def pu_chain_bwd(self, wT_hi, wT_lo, G, g_rs, g_ts, F_, f_rs, f_ts, C_all, H, dOut, dG, dg_rs, dg_ts, dF, df_rs, df_ts, x_hi,
                     x_lo, counters, B, J, precision):
        """BPTT of one layer in one call == for t = J-1 .. 0: pu_cell_bwd, then dhg = dgates . W_hh (wT = W_hh^T pair)"""
        Hd = 512
        wt = pair_f32(wT_hi, wT_lo if precision == PREC_BF16X3 else None, (Hd, 4 * Hd), (4 * Hd, 1))
        dhg, dc = torch.zeros(B, Hd), torch.zeros(B, Hd)
        for t in range(J - 1, -1, -1):
            gh = flat(x_hi)[:B * 4 * Hd].view(B, 4 * Hd)
            gl = None if (x_lo is None or precision != PREC_BF16X3) else flat(x_lo)[:B * 4 * Hd].view(B, 4 * Hd)
            self.pu_cell_bwd(G, g_rs, g_ts, F_, f_rs, f_ts, C_all, H, dOut, dhg, dc, dG, dg_rs, dg_ts, dF, df_rs, df_ts, gh, gl, t, J, B)
            if t > 0:
                dhg = pair_f32(gh, gl, (B, 4 * Hd), (4 * Hd, 1)) @ wt.t()

    def pu_bridge_gate_bwd(self, dE, e_ld, F0, f_ld, f_col, E, X, rows, dF, df_ld):
        """backward of pu_bridge_gate: in: dE[r][X + c] = d b' ; out: dE[r][X + c] = d b' * sigmoid(Fb),
        dF[r][f_col + c] = d b' * bridge * sigmoid'(Fb)"""
        self.launches += 1
        r = torch.arange(rows).view(-1, 1)
        c = torch.arange(X).view(1, -1)
        db = flat(dE)[r * e_ld + X + c].double()
        s = torch.sigmoid(flat(F0)[r * f_ld + f_col + c].double())
        bridge = flat(E)[r * e_ld + X + c].double()
        flat(dE)[r * e_ld + X + c] = (db * s).float()
        flat(dF)[r * df_ld + f_col + c] = (db * bridge * s * (1 - s)).float()

    def head_bwd(self, dpose, e, e_ld, skel, Wp, Wg, frames, J, dE, de_ld, dSkel, dWp, dbp, dWg, dbg, scratch=None):
        """backward of head(): writes dE[r][:256] (and ZEROES dE[r][256:512]), dSkel, dWp, dbp, dWg, dbg"""
        self.launches += 2
        nj = J + 1 if Wg is not None else J
        dp = flat(dpose)[:frames * nj * 3].view(frames, nj, 3).double()
        dj = dp[:, :J].reshape(frames * J, 3)
        w = flat(Wp)[:3 * 768].view(3, 768).double()
        pe = torch.as_strided(e, (frames * J, 256), (e_ld, 1)).double()
        sk = flat(skel)[:frames * J * 512].view(frames * J, 512).double()
        d_in = dj @ w                                          # (frames*J, 768)
        dsk = d_in[:, 256:]
        flat(dWp)[:3 * 768] = (dj.t() @ torch.cat([pe, sk], 1)).float().reshape(-1)
        flat(dbp)[:3] = dj.sum(0).float()
        if Wg is not None:
            do = torch.cat([dp[:, :J].sum(1), dp[:, J]], 1)    # (frames, 6)
            wg = flat(Wg)[:6 * J * 512].view(6, J * 512).double()
            dsk = dsk + (do @ wg).view(frames * J, 512)
            flat(dWg)[:6 * J * 512] = (do.t() @ sk.reshape(frames, J * 512)).float().reshape(-1)
            flat(dbg)[:6] = do.sum(0).float()
        flat(dSkel)[:frames * J * 512] = dsk.float().reshape(-1)
        r = torch.arange(frames * J).view(-1, 1)
        flat(dE)[r * de_ld + torch.arange(256).view(1, -1)] = d_in[:, :256].float()
        flat(dE)[r * de_ld + 256 + torch.arange(256).view(1, -1)] = 0.0

    def embed_grads(self, dpos_perm, grid, n_hm, dpos, dmask):
        """dpos_perm: (576, 1024) gradient per heatmap-major token summed over frames ->
        dpos (raster order, reference modeling_vit.py:153) and dmask = sum over the dummy tokens (:137-142)"""
        self.launches += 1
        perm = token_perm(grid)
        d = flat(dpos_perm)[:576 * 1024].view(576, 1024)
        out = torch.zeros(576, 1024)
        out[perm] = d
        flat(dpos)[:576 * 1024] = out.reshape(-1)
        flat(dmask)[:1024] = d[n_hm * 16:].sum(0)

    def pose_loss(self, pred, gt, frames, nj, parents, drop_first, lambda_mpjpe, lambda_cos, loss, dpose, scratch=None):
        """loss[0] = total, loss[1] = mpjpe term, loss[2] = cos-sim term (reference
        model/egotap_autoencoder_model.py:284-296, utils/loss.py:44-85); dpose = d total / d pred.
        parents: kinematic parents over the (possibly root-prepended) joint list; drop_first: EgoCap prepends a zero
        root and drops the first bone."""
        self.launches += 2
        p = flat(pred)[:frames * nj * 3].view(frames, nj, 3).double().clone().requires_grad_(True)
        g = flat(gt)[:frames * nj * 3].view(frames, nj, 3).double()
        mp = torch.linalg.norm(g - p, dim=-1).mean() * lambda_mpjpe
        pp, gg = p, g
        if drop_first:
            z = p.new_zeros(frames, 1, 3)
            pp, gg = torch.cat([z, p], 1), torch.cat([z, g], 1)
        par = torch.tensor(parents)
        pb, gb = (pp - pp[:, par])[:, 1:], (gg - gg[:, par])[:, 1:]
        cos = F.cosine_similarity(pb, gb, dim=2)
        if drop_first:
            cos = cos[:, 1:]
        cs = cos.sum(1).mean(0) * lambda_cos * lambda_mpjpe
        total = mp + cs
        (dp,) = torch.autograd.grad(total, p)
        flat(loss)[:3] = torch.stack([total, mp, cs]).detach().float()
        flat(dpose)[:frames * nj * 3] = dp.float().reshape(-1)

    def gt_heatmaps(self, pts2d, pts3d_left, frames, preset, out):
        """keypoints -> (frames, 6J, 64, 64) lifting input; semantics in oracle/gt_heatmap_oracle.py"""
        import gt_heatmap_oracle as gto
        self.launches += 1
        p2, p3 = pts2d.numpy(), pts3d_left.numpy()
        for b in range(frames):
            # the right view's own 3-D points only matter for theta, which the reference takes from the left view
            hm = gto.lifting_input(p2[b, 0], p2[b, 1], p3[b], p3[b], preset)
            out[b].copy_(torch.from_numpy(hm))

    def adamw(self, params, grads, m, v, step, lr, beta1, beta2, eps, weight_decay, grad_scale=1.0):
        """torch.optim.AdamW (reference model/network.py:72-78); params / grads / m / v: lists of tensors"""
        self.launches += 1
        bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
        for p, g, mm, vv in zip(params, grads, m, v):
            pf, gf, mf, vf = p.data.view(-1), flat(g)[:p.numel()] * grad_scale, flat(mm)[:p.numel()], flat(vv)[:p.numel()]
            pf.mul_(1 - lr * weight_decay)
            mf.mul_(beta1).add_(gf, alpha=1 - beta1)
            vf.mul_(beta2).addcmul_(gf, gf, value=1 - beta2)
            pf.sub_((lr / bc1) * mf / (vf.sqrt() / math.sqrt(bc2) + eps))


def token_perm(grid):
    """heatmap-major token -> raster token of the mosaic (reference model/net_architecture.py:397-402)"""
    side = grid * 4
    out = []
    for n in range(grid * grid):
        for pr in range(4):
            for pc in range(4):
                out.append(((n // grid) * 4 + pr) * side + (n % grid) * 4 + pc)
    return torch.tensor(out)
