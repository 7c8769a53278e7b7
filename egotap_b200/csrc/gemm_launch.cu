// Host launcher + C-ABI entry for the tcgen05 GEMM (gemm.cuh).
#include "gemm.cuh"
#include "host_util.cuh"
#include "internal.h"

#include <cstdlib>
#include <cstring>
#include <vector>

namespace eb {

static EpiMaps g_no_maps;      // kernels without the TMA epilogue never look at their map argument

template <int CG, int BN, int NS, int ST, bool COAL = false, bool TN = false, bool EW16 = false, bool TEPI = false>
static int launch_variant_maps(const CUtensorMap* tm, const GemmShape& s, const EpiParams& ep, const EpiMaps& maps, cudaStream_t stream) {
  using C = GemmCfg<CG, BN, NS, ST, COAL, TN, EW16, TEPI>;
  auto kern = gemm_tc_kernel<CG, BN, NS, ST, COAL, TN, EW16, TEPI>;
  // the opt-in to > 48 KB dynamic shared memory is a per-device function attribute
  static bool attr_done[64] = {false};
  const int dev_ = current_device();
  if (!attr_done[dev_]) {
    EB_CUDA(EB_SET_MAX_SMEM(kern, C::SMEM_BYTES));
    attr_done[dev_] = true;
  }
  const int nM = (s.M + C::BM * CG - 1) / (C::BM * CG);
  const int nN = (s.N + BN - 1) / BN;
  const long long tiles = (long long)s.groups * nM * nN;
  long long nclusters = num_sms() / CG;
  if (tiles < nclusters) nclusters = tiles;
  EB_CUDA(EB_LAUNCH_CLUSTER(kern, (unsigned)(nclusters * CG), C::THREADS, C::SMEM_BYTES, CG, stream, tm[0], tm[1], tm[2], tm[3],
                            s, ep, maps));
  EB_CHECK_LAUNCH("gemm_tc_kernel");
  return 0;
}
template <int CG, int BN, int NS, int ST, bool COAL = false, bool TN = false, bool EW16 = false>
static int launch_variant(const CUtensorMap* tm, const GemmShape& s, const EpiParams& ep, cudaStream_t stream) {
  return launch_variant_maps<CG, BN, NS, ST, COAL, TN, EW16, false>(tm, s, ep, g_no_maps, stream);
}

struct Variant {
  const char* name;
  int cg, bn, nsplit;
  int (*launch)(const CUtensorMap*, const GemmShape&, const EpiParams&, cudaStream_t);
  int (*launch_coalesced)(const CUtensorMap*, const GemmShape&, const EpiParams&, cudaStream_t);   // EGOTAP_EPI=coalesced
  int (*launch_tn)(const CUtensorMap*, const GemmShape&, const EpiParams&, cudaStream_t);          // D = A^T B (BN = 256 only)
  int (*launch_ew16)(const CUtensorMap*, const GemmShape&, const EpiParams&, cudaStream_t);        // row-domain epilogue on 16 warps
  int (*launch_tepi)(const CUtensorMap*, const GemmShape&, const EpiParams&, const EpiMaps&, cudaStream_t);   // TMA epilogue (fp32 out)
};

static const Variant kVariants[] = {
    {"cg1_bn128_bf16_s6", 1, 128, 1, &launch_variant<1, 128, 1, 6>, &launch_variant<1, 128, 1, 6, true>, nullptr, nullptr, nullptr},
    {"cg1_bn128_bf16x3_s3", 1, 128, 3, &launch_variant<1, 128, 3, 3>, &launch_variant<1, 128, 3, 3, true>, nullptr, nullptr, nullptr},
    {"cg1_bn256_bf16_s4", 1, 256, 1, &launch_variant<1, 256, 1, 4>, &launch_variant<1, 256, 1, 4, true>,
     &launch_variant<1, 256, 1, 4, true, true>, nullptr, nullptr},
    {"cg1_bn256_bf16x3_s2", 1, 256, 3, &launch_variant<1, 256, 3, 2>, &launch_variant<1, 256, 3, 2, true>,
     &launch_variant<1, 256, 3, 2, true, true>, nullptr, nullptr},
    // the TMA-epilogue forms trade operand-ring stages for the epilogue boxes (gemm.cuh): 4 stages + 3 boxes per warp in bf16
    // mode, 3 stages + 1 box in the parity mode (both fill the 227 KB exactly)
    {"cg2_bn256_bf16_s6", 2, 256, 1, &launch_variant<2, 256, 1, 6>, &launch_variant<2, 256, 1, 6, true>,
     &launch_variant<2, 256, 1, 6, true, true>, &launch_variant<2, 256, 1, 6, false, false, true>,
     &launch_variant_maps<2, 256, 1, 4, false, false, false, true>},
    {"cg2_bn256_bf16x3_s3", 2, 256, 3, &launch_variant<2, 256, 3, 3>, &launch_variant<2, 256, 3, 3, true>,
     &launch_variant<2, 256, 3, 3, true, true>, &launch_variant<2, 256, 3, 3, false, false, true>,
     &launch_variant_maps<2, 256, 3, 3, false, false, false, true>},
};
static const int kNumVariants = int(sizeof(kVariants) / sizeof(kVariants[0]));

// EGOTAP_EPI_TMA=0 keeps the load / store epilogue for fp32 outputs (A/B runs and tests; read per call)
static int tepi_mode() {
  const char* e = getenv("EGOTAP_EPI_TMA");
  return (e && e[0] == '0') ? 0 : 1;
}

static int g_prefer_cg = -1;
static int prefer_cg() {
  if (g_prefer_cg < 0) {
    const char* e = getenv("EGOTAP_GEMM_CG");
    g_prefer_cg = (e && e[0] == '1') ? 1 : 2;
  }
  return g_prefer_cg;
}

int pick_variant(int M, int N, int groups, int nsplit) {
  const bool wide = (N % 256 == 0) || N > 256;
  if (!wide) return nsplit == 3 ? 1 : 0;
  // the paired-SM tile wants enough 256-row tiles to fill 74 clusters reasonably
  const long long tiles2 = (long long)groups * ((M + 255) / 256) * ((N + 255) / 256);
  if (prefer_cg() == 2 && M >= 256 && tiles2 >= 16) return nsplit == 3 ? 5 : 4;
  return nsplit == 3 ? 3 : 2;
}

// D[g] = A[g K .. (g+1) K)^T B[g K .. (g+1) K): row-major operands with the contraction along the rows (GemmShape comment in
// gemm.cuh); a.rows / b.rows = total rows of the matrices (reads beyond are zero-filled), a.ld / b.ld their row strides
int gemm_tn_run(const GemmOperand& a, const GemmOperand& b, const GemmShape& s, const EpiParams& ep, int nsplit, int variant,
                cudaStream_t stream) {
  EB_REQUIRE(s.M > 0 && s.N > 0 && s.K > 0 && s.groups > 0, "gemm(tn): bad shape M %d N %d K %d groups %d", s.M, s.N, s.K, s.groups);
  EB_REQUIRE(s.K % 64 == 0 && s.N % 256 == 0 && s.M % 64 == 0, "gemm(tn): K %% 64, N %% 256, M %% 64 must be 0 (M %d N %d K %d)", s.M,
             s.N, s.K);
  EB_REQUIRE(nsplit == 1 || nsplit == 3, "gemm: nsplit must be 1 or 3");
  EB_REQUIRE(a.hi && b.hi && (nsplit == 1 || (a.lo && b.lo)), "gemm(tn): null operand");
  EB_REQUIRE(a.rows == b.rows && a.rows > 0, "gemm(tn): both operands have the contraction along their rows (a.rows %lld b.rows %lld)",
             a.rows, b.rows);
  if (variant < 0) {
    const long long tiles2 = (long long)s.groups * ((s.M + 255) / 256) * (s.N / 256);
    variant = (prefer_cg() == 2 && s.M >= 256 && tiles2 >= 16) ? (nsplit == 3 ? 5 : 4) : (nsplit == 3 ? 3 : 2);
  }
  EB_REQUIRE(variant >= 0 && variant < kNumVariants && kVariants[variant].launch_tn && kVariants[variant].nsplit == nsplit,
             "gemm(tn): variant %d has no transposed-operand form for nsplit %d", variant, nsplit);
  CUtensorMap tm[4];
  int rc;
  // contiguous extent = the M / N columns, rows = the contraction; boxes of 64 x 64
  if ((rc = make_operand_tmap(&tm[0], a.hi, s.M, a.rows, a.ld, 1, 0, 1, 0, 64))) return rc;
  if ((rc = make_operand_tmap(&tm[2], b.hi, s.N, b.rows, b.ld, 1, 0, 1, 0, 64))) return rc;
  if (nsplit == 3) {
    if ((rc = make_operand_tmap(&tm[1], a.lo, s.M, a.rows, a.ld, 1, 0, 1, 0, 64))) return rc;
    if ((rc = make_operand_tmap(&tm[3], b.lo, s.N, b.rows, b.ld, 1, 0, 1, 0, 64))) return rc;
  } else {
    tm[1] = tm[0];
    tm[3] = tm[2];
  }
  GemmShape sh = s;
  sh.gdiv = 1;
  ProfScope prof("gemm_tc_kernel", stream, s.M, s.N, s.K, s.groups, variant);
  return kVariants[variant].launch_tn(tm, sh, ep, stream);
}

int gemm_run(const GemmOperand& a, const GemmOperand& b, const GemmShape& s, const EpiParams& ep, int nsplit,
             int variant, cudaStream_t stream) {
  EB_REQUIRE(s.M > 0 && s.N > 0 && s.K > 0 && s.groups > 0, "gemm: bad shape M %d N %d K %d groups %d", s.M, s.N, s.K,
             s.groups);
  EB_REQUIRE(s.K % 64 == 0, "gemm: K (%d) must be a multiple of 64", s.K);
  EB_REQUIRE(s.N % 32 == 0, "gemm: N (%d) must be a multiple of 32", s.N);
  EB_REQUIRE(nsplit == 1 || nsplit == 3, "gemm: nsplit must be 1 or 3");
  EB_REQUIRE(a.hi && b.hi, "gemm: null operand");
  EB_REQUIRE(nsplit == 1 || (a.lo && b.lo), "gemm: bf16x3 mode needs lo parts of both operands");
  if (variant < 0) variant = pick_variant(s.M, s.N, s.groups, nsplit);
  EB_REQUIRE(variant < kNumVariants, "gemm: variant %d out of range", variant);
  const Variant& v = kVariants[variant];
  EB_REQUIRE(v.nsplit == nsplit, "gemm: variant %s does not match nsplit %d", v.name, nsplit);
  CUtensorMap tm[4];
  int rc;
  const int box_b = v.bn / v.cg;
  if ((rc = make_operand_tmap(&tm[0], a.hi, s.K, a.rows, a.ld, a.g0_count, a.g0_stride, a.g1_count, a.g1_stride, 128)))
    return rc;
  if ((rc = make_operand_tmap(&tm[2], b.hi, s.K, b.rows, b.ld, b.g0_count, b.g0_stride, b.g1_count, b.g1_stride, box_b)))
    return rc;
  if (nsplit == 3) {
    if ((rc = make_operand_tmap(&tm[1], a.lo, s.K, a.rows, a.ld, a.g0_count, a.g0_stride, a.g1_count, a.g1_stride, 128)))
      return rc;
    if ((rc = make_operand_tmap(&tm[3], b.lo, s.K, b.rows, b.ld, b.g0_count, b.g0_stride, b.g1_count, b.g1_stride, box_b)))
      return rc;
  } else {
    tm[1] = tm[0];
    tm[3] = tm[2];
  }
  GemmShape sh = s;
  if (sh.gdiv <= 0) sh.gdiv = int(a.g0_count > 0 ? a.g0_count : 1);
  ProfScope prof("gemm_tc_kernel", stream, s.M, s.N, s.K, s.groups, variant);
  // Epilogue form (gemm.cuh).  Measured on the B200 (profiles/r02_epilogue_ab.md): re-distributing each warp's chunk through
  // shared memory pays where the epilogue touches fp32 rows (residual stream in / out: out-projection 533 -> 754 TFLOP/s in
  // bf16 mode, patch embedding 152 -> 211) and costs where it only writes bf16 operands (MLP-up + GELU 889 -> 692: the
  // extra shared-memory round trip lands on an epilogue that is already issue-bound).  EGOTAP_EPI=coalesced / rows forces
  // one form for every GEMM (A/B runs); read per call so that one process can compare them.
  bool coalesced = ep.resid != nullptr || ep.out_f32 != nullptr;
  if (const char* epi_env = getenv("EGOTAP_EPI")) {
    if (strcmp(epi_env, "coalesced") == 0) coalesced = true;
    else if (strcmp(epi_env, "rows") == 0) coalesced = false;
  }
  if (ep.stats_out != nullptr) coalesced = true;      // the LayerNorm-fold producers: row statistics exist in these two forms only
  // Measured on the B200 (profiles/r02g_tma_epilogue.md, bf16 mode): out-projection (K = 1024) 682 -> 926 TFLOP/s, patch embedding
  // (K = 256) 185 -> 247, but MLP-down (K = 4096, main-loop bound with the 6-stage ring) 1,359 -> 1,222 on the 4-stage ring the
  // boxes leave room for: the TMA epilogue is used where the main loop of a tile is short (K <= 1024).
  if (coalesced && v.launch_tepi && tepi_mode() != 0 && s.K <= 1024) {
    // TMA epilogue (gemm.cuh, TEPI): fp32 row-major output only, and every 32-row box of a warp must map to 32 consecutive
    // output rows and 32 consecutive residual rows
    const bool rows_ok = s.M % 32 == 0 && (ep.rows_in == 0 || ep.rows_in % 32 == 0) && (ep.resid_mod == 0 || ep.resid_mod % 32 == 0);
    // (a bf16 copy of the stored rows -- the LayerNorm-fold producers -- is written with 32-byte row-domain stores)
    const bool bf_ok = (!ep.out_hi && !ep.out_lo) ||
                       (ep.out_hi && ep.ldo % 16 == 0 && ep.col_off % 16 == 0 &&
                        ((reinterpret_cast<uintptr_t>(ep.out_hi) | reinterpret_cast<uintptr_t>(ep.out_lo)) & 31) == 0);
    const bool out_ok = ep.out_f32 && bf_ok && ep.store == STORE_ROWMAJOR && ep.ldo % 4 == 0 && ep.col_off % 4 == 0 &&
                        (reinterpret_cast<uintptr_t>(ep.out_f32) & 15) == 0 && ep.col_off + s.N <= ep.ldo;
    const bool res_ok = ep.resid == nullptr || (ep.resid_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(ep.resid) & 15) == 0 &&
                                                 ep.col_off + s.N <= ep.resid_ld);
    if (rows_ok && out_ok && res_ok) {
      const long long last = s.M - 1;
      const long long out_rows = (long long)(s.groups - 1) * ep.group_rows +
                                 (ep.rows_in > 0 ? (last / ep.rows_in) * ep.rows_out + last % ep.rows_in : last) + 1;
      EpiMaps maps;
      if ((rc = make_operand_tmap(&maps.out, ep.out_f32, 2 * ep.ldo, out_rows, 2 * ep.ldo, 1, 0, 1, 0, 32))) return rc;
      if (ep.resid) {
        const long long res_rows = ep.resid_mod > 0 ? ep.resid_mod : out_rows;
        if ((rc = make_operand_tmap(&maps.res, ep.resid, 2 * ep.resid_ld, res_rows, 2 * ep.resid_ld, 1, 0, 1, 0, 32))) return rc;
      } else {
        maps.res = maps.out;
      }
      return v.launch_tepi(tm, sh, ep, maps, stream);
    }
  }
  if (coalesced) return v.launch_coalesced(tm, sh, ep, stream);
  // row-domain epilogue: 16 warps for the GELU epilogue (MLP-up), where the math per element is long enough for four warps per
  // scheduler to pay (measured, bf16 mode: 966 -> 1,060 TFLOP/s; the plain bf16 stores of QKV lose 6 % with 16 warps, the
  // parity mode is indifferent: profiles/r02e_epilogue_warps.md).  EGOTAP_EPI_WARPS=8 / 16 forces one form (A/B).
  static int ew = -1;
  if (ew < 0) { const char* e = getenv("EGOTAP_EPI_WARPS"); ew = !e ? 0 : (e[0] == '8' ? 8 : 16); }
  const bool want16 = ew == 16 || (ew == 0 && ep.act == ACT_GELU);
  if (want16 && v.launch_ew16 && ep.resid == nullptr) return v.launch_ew16(tm, sh, ep, stream);
  return v.launch(tm, sh, ep, stream);
}

}  // namespace eb

using namespace eb;

extern "C" int egotap_b200_gemm_num_variants(void) { return kNumVariants; }
extern "C" const char* egotap_b200_gemm_variant_name(int v) {
  return (v >= 0 && v < kNumVariants) ? kVariants[v].name : "";
}

extern "C" int egotap_b200_gemm(const egotap_gemm* d, void* stream) {
  if (!d) return fail(EGOTAP_E_ARG, "gemm: null descriptor");
  GemmOperand a{(const __nv_bfloat16*)d->a.hi, (const __nv_bfloat16*)d->a.lo, d->a.ld, d->a.rows,
                d->a.g0_count, d->a.g0_stride, d->a.g1_count, d->a.g1_stride};
  GemmOperand b{(const __nv_bfloat16*)d->b.hi, (const __nv_bfloat16*)d->b.lo, d->b.ld, d->b.rows,
                d->b.g0_count, d->b.g0_stride, d->b.g1_count, d->b.g1_stride};
  GemmShape s{d->M, d->N, d->K, d->groups, 0, 0};
  const egotap_epilogue& e = d->epi;
  EpiParams ep;
  memset(&ep, 0, sizeof(ep));       // the LayerNorm-fold fields are internal to the lifting plan
  ep.alpha = e.alpha; ep.scale = e.scale; ep.bias = e.bias; ep.act = e.act;
  ep.resid = e.resid; ep.resid_ld = e.resid_ld; ep.resid_mod = e.resid_mod;
  ep.rows_in = e.rows_in; ep.rows_out = e.rows_out; ep.group_rows = e.group_rows;
  ep.out_f32 = e.out_f32; ep.out_hi = (__nv_bfloat16*)e.out_hi; ep.out_lo = (__nv_bfloat16*)e.out_lo;
  ep.ldo = e.ldo; ep.col_off = e.col_off; ep.store = e.store;
  ep.qk_cols = e.qk_cols; ep.tokens = e.tokens;
  ep.vt_hi = (__nv_bfloat16*)e.vt_hi; ep.vt_lo = (__nv_bfloat16*)e.vt_lo; ep.J = e.J; ep.heads = e.heads;
  EB_REQUIRE(ep.out_f32 || ep.out_hi || (ep.store == STORE_QKV && ep.vt_hi), "gemm: no output pointer");
  EB_REQUIRE(ep.ldo % 8 == 0 && ep.col_off % 32 == 0, "gemm: ldo %% 8 and col_off %% 32 must be 0");
  const int nsplit = d->precision == EGOTAP_PREC_BF16 ? 1 : 3;
  if (d->layout == EGOTAP_GEMM_TN) return gemm_tn_run(a, b, s, ep, nsplit, d->variant, (cudaStream_t)stream);
  EB_REQUIRE(d->layout == EGOTAP_GEMM_NT, "gemm: unknown operand layout %d", d->layout);
  return gemm_run(a, b, s, ep, nsplit, d->variant, (cudaStream_t)stream);
}
