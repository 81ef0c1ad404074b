// a14 + a15: Clebsch-Gordan assembly of the nao x nao orbital block from the HamLayer coefficients,
// orbital reorder (folded into the CSR row order on the host), inverse-edge symmetrisation, +H0,
// orbital masks, and placement into the per-crystal interleaved output rows.  HBM-bound stages.
#include "hgb_common.cuh"

namespace {

constexpr int HT = 256;

struct AsmArgs {
  const int32_t* row_ptr;
  const int32_t* col;
  const float* val;
  int n_out, n_in;
  const float* coef;
  int64_t n_rows;
  float* raw;
};

// one CTA per HROWS rows; coefficients staged in shared memory, every thread owns matrix entries q
constexpr int HROWS = 8;
__global__ void __launch_bounds__(HT) ham_assemble_kernel(const __grid_constant__ AsmArgs a) {
  extern __shared__ float sc[];
  const int nc = a.n_in, nn = a.n_out;
  const int64_t r0 = (int64_t)blockIdx.x * HROWS;
  const int nr = (int)min((int64_t)HROWS, a.n_rows - r0);
  for (int idx = threadIdx.x; idx < nr * nc; idx += HT) sc[idx] = a.coef[r0 * nc + idx];
  __syncthreads();
  for (int q = threadIdx.x; q < nn; q += HT) {
    const int n0 = a.row_ptr[q], n1 = a.row_ptr[q + 1];
    float acc[HROWS];
#pragma unroll
    for (int r = 0; r < HROWS; ++r) acc[r] = 0.f;
    for (int n = n0; n < n1; ++n) {
      const float v = a.val[n];
      const int c = a.col[n];
#pragma unroll
      for (int r = 0; r < HROWS; ++r) acc[r] = fmaf(v, sc[r * nc + c], acc[r]);  // rows >= nr read stale smem, never stored
    }
#pragma unroll
    for (int r = 0; r < HROWS; ++r)
      if (r < nr) a.raw[(r0 + r) * nn + q] = acc[r];
  }
}

struct FinArgs {
  hgb_ham_plan plan;
  const float* raw;
  const int64_t* partner;
  const float* h0;
  const int64_t* z;
  const int64_t* node_a;
  const int64_t* node_b;
  const int64_t* out_row;
  int64_t n_rows;
  int symmetrize;
  float* out;
};

__global__ void __launch_bounds__(HT) ham_finalize_kernel(const __grid_constant__ FinArgs a) {
  const int nao = a.plan.nao, nn = nao * nao;
  const int64_t r = blockIdx.x;
  const int64_t pr = a.partner ? a.partner[r] : r;
  const int64_t na = a.node_a ? a.node_a[r] : r;
  const int64_t nb = a.node_b ? a.node_b[r] : r;
  const int za = (int)a.z[na], zb = (int)a.z[nb];
  const uint8_t* ma = a.plan.orb_mask + (size_t)za * nao;
  const uint8_t* mb = a.plan.orb_mask + (size_t)zb * nao;
  const int64_t orow = a.out_row ? a.out_row[r] : r;
  for (int q = threadIdx.x; q < nn; q += HT) {
    const int i = q / nao, j = q - i * nao;
    float v = a.raw[r * nn + q];
    if (a.symmetrize) v = 0.5f * (v + a.raw[pr * nn + j * nao + i]);
    if (a.h0) v += a.h0[r * nn + q];
    a.out[orow * nn + q] = (ma[i] && mb[j]) ? v : 0.f;
  }
}

// ------------------------------------------------------------------------------------------------ a16 (SOC)
// su2: raw = [rows][2 planes][M][M] (M = 2 nao; row index s1*nao + a, column s2*nao + b) from hgb_csr_rows.
// Hermitian symmetrisation with the inverse edge on the complex (2 nao)^2 matrix, orbital masks per spin block,
// then +H0 (real) / +iH0 (imaginary) -- the reference adds H0 AFTER masking in the SOC branch.
struct SocFinArgs {
  int nao;
  const uint8_t* orb_mask;
  const float* raw;
  const int64_t* partner;
  const float* h0_re;
  const float* h0_im;
  const int64_t* z;
  const int64_t* node_a;
  const int64_t* node_b;
  const int64_t* out_row;
  int64_t n_rows;
  int symmetrize;
  float* out_re;
  float* out_im;
};

__global__ void __launch_bounds__(HT) ham_finalize_su2_kernel(const __grid_constant__ SocFinArgs a) {
  const int nao = a.nao, M = 2 * nao, MM = M * M;
  const int64_t r = blockIdx.x;
  const int64_t pr = a.partner ? a.partner[r] : r;
  const int64_t na = a.node_a ? a.node_a[r] : r;
  const int64_t nb = a.node_b ? a.node_b[r] : r;
  const uint8_t* ma = a.orb_mask + (size_t)a.z[na] * nao;
  const uint8_t* mb = a.orb_mask + (size_t)a.z[nb] * nao;
  const int64_t orow = a.out_row ? a.out_row[r] : r;
  const float* mine = a.raw + (size_t)r * 2 * MM;
  const float* other = a.raw + (size_t)pr * 2 * MM;
  for (int q = threadIdx.x; q < MM; q += HT) {
    const int p = q / M, c = q - p * M;
    float vr = mine[q], vi = mine[MM + q];
    if (a.symmetrize) {
      const int qt = c * M + p;
      vr = 0.5f * (vr + other[qt]);
      vi = 0.5f * (vi - other[MM + qt]);
    }
    const int oa = (p >= nao) ? p - nao : p, ob = (c >= nao) ? c - nao : c;
    if (!(ma[oa] && mb[ob])) { vr = 0.f; vi = 0.f; }
    if (a.h0_re) vr += a.h0_re[r * MM + q];
    if (a.h0_im) vi += a.h0_im[r * MM + q];
    a.out_re[orow * MM + q] = vr;
    a.out_im[orow * MM + q] = vi;
  }
}

// so3: in place, ksi[r] <- average over the m components of every p/d/f shell, rows first, then columns
// (symmetrize_orbital_coefficients).  One CTA per row, the row lives in shared memory.
struct KsiArgs {
  int nao, n_blocks;
  int blk_lo[8], blk_hi[8];
  float* ksi;
  int64_t n_rows;
};
__global__ void __launch_bounds__(HT) ksi_shell_average_kernel(const __grid_constant__ KsiArgs a) {
  extern __shared__ float sk[];
  const int nao = a.nao, nn = nao * nao;
  float* row = a.ksi + (size_t)blockIdx.x * nn;
  for (int q = threadIdx.x; q < nn; q += HT) sk[q] = row[q];
  __syncthreads();
  // rows: for every shell [lo, hi) and every column j: mean over i in the shell
  for (int b = 0; b < a.n_blocks; ++b) {
    const int lo = a.blk_lo[b], hi = a.blk_hi[b];
    for (int j = threadIdx.x; j < nao; j += HT) {
      float s = 0.f;
      for (int i = lo; i < hi; ++i) s += sk[i * nao + j];
      s /= (float)(hi - lo);
      for (int i = lo; i < hi; ++i) sk[i * nao + j] = s;
    }
  }
  __syncthreads();
  for (int b = 0; b < a.n_blocks; ++b) {
    const int lo = a.blk_lo[b], hi = a.blk_hi[b];
    for (int i = threadIdx.x; i < nao; i += HT) {
      float s = 0.f;
      for (int j = lo; j < hi; ++j) s += sk[i * nao + j];
      s /= (float)(hi - lo);
      for (int j = lo; j < hi; ++j) sk[i * nao + j] = s;
    }
  }
  __syncthreads();
  for (int q = threadIdx.x; q < nn; q += HT) row[q] = sk[q];
}

// so3: spin blocks from the spin-less block Hns and A_c = 0.5 (ksi L_c - (ksi L_c)[partner]^T):
//   real: uu = dd = Hns, ud = du = A_1 ;  imaginary: uu = A_2, dd = -A_2, ud = A_0, du = -A_0 ; then +H0 / +iH0.
struct So3Args {
  int nao;
  const float* hns;      // [rows][nao^2], final spin-less block (symmetrised + masked, or the given Hon_nonsoc)
  const float* ksi;      // [rows][nao^2], shell-averaged
  const float* lmat;     // [rows][nao^2][3]
  const int64_t* partner;
  const float* h0_re;
  const float* h0_im;
  const int64_t* out_row;
  int64_t n_rows;
  int symmetrize, h0_offdiag_only;
  float* out_re;
  float* out_im;
};
__global__ void __launch_bounds__(HT) ham_finalize_so3_kernel(const __grid_constant__ So3Args a) {
  const int nao = a.nao, nn = nao * nao, M = 2 * nao, MM = M * M;
  const int64_t r = blockIdx.x;
  const int64_t pr = a.partner ? a.partner[r] : r;
  const int64_t orow = a.out_row ? a.out_row[r] : r;
  for (int q = threadIdx.x; q < nn; q += HT) {
    const int i = q / nao, j = q - i * nao;
    const float k = a.ksi[r * nn + q];
    const float* lm = a.lmat + ((size_t)r * nn + q) * 3;
    float A0 = k * lm[0], A1 = k * lm[1], A2 = k * lm[2];
    if (a.symmetrize) {
      const int qt = j * nao + i;
      const float kp = a.ksi[pr * nn + qt];
      const float* lp = a.lmat + ((size_t)pr * nn + qt) * 3;
      A0 = 0.5f * (A0 - kp * lp[0]); A1 = 0.5f * (A1 - kp * lp[1]); A2 = 0.5f * (A2 - kp * lp[2]);
    }
    const float h = a.hns[r * nn + q];
    const int uu = i * M + j, ud = i * M + nao + j, du = (nao + i) * M + j, dd = (nao + i) * M + nao + j;
    float re[4] = {h, A1, A1, h}, im[4] = {A2, A0, -A0, -A2};
    const int pos[4] = {uu, ud, du, dd};
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const bool diag = (s == 0 || s == 3);
      if (a.h0_re && !(a.h0_offdiag_only && diag)) re[s] += a.h0_re[r * MM + pos[s]];
      if (a.h0_im) im[s] += a.h0_im[r * MM + pos[s]];
      a.out_re[orow * MM + pos[s]] = re[s];
      a.out_im[orow * MM + pos[s]] = im[s];
    }
  }
}

}  // namespace

extern "C" int hgb_ham_assemble(const hgb_ham_plan* plan, const float* coef, int64_t n_rows, float* raw, void* stream) {
  HGB_DEVICE_GUARD(raw);
  HGB_CHECK_ARG(plan && coef && raw, "hgb_ham_assemble: NULL argument");
  hgb::TimeScope ts_(HGB_K_HAM_ASSEMBLE, stream);
  HGB_CHECK_ARG(plan->nao > 0 && plan->nao <= 64 && plan->n_coef > 0, "hgb_ham_assemble: bad plan (nao=%d)", plan->nao);
  if (n_rows == 0) return 0;
  return hgb_csr_rows(plan->row_ptr, plan->col, plan->val, plan->nao * plan->nao, plan->n_coef, coef, n_rows, raw, stream);
}

extern "C" int hgb_csr_rows(const int32_t* row_ptr, const int32_t* col, const float* val, int32_t n_out, int32_t n_in,
                            const float* x, int64_t n_rows, float* y, void* stream) {
  HGB_DEVICE_GUARD(y);
  HGB_CHECK_ARG(row_ptr && col && val && x && y, "hgb_csr_rows: NULL argument");
  HGB_CHECK_ARG(n_out > 0 && n_in > 0 && n_rows >= 0, "hgb_csr_rows: bad sizes (n_out=%d, n_in=%d)", n_out, n_in);
  const size_t smem = (size_t)HROWS * n_in * sizeof(float);
  HGB_CHECK_ARG(smem <= 220 * 1024, "hgb_csr_rows: %d input columns x %d rows do not fit in shared memory", n_in, HROWS);
  if (n_rows == 0) return 0;
  AsmArgs a;
  a.row_ptr = row_ptr; a.col = col; a.val = val; a.n_out = n_out; a.n_in = n_in; a.coef = x; a.n_rows = n_rows; a.raw = y;
  HGB_CUDA_OK(cudaFuncSetAttribute(ham_assemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ham_assemble_kernel<<<(unsigned)((n_rows + HROWS - 1) / HROWS), HT, smem, (cudaStream_t)stream>>>(a);
  HGB_LAUNCH_OK("ham_assemble_kernel");
  return 0;
}

extern "C" int hgb_ham_finalize(const hgb_ham_plan* plan, const float* raw, const int64_t* partner, const float* h0,
                                const int64_t* z, const int64_t* node_a, const int64_t* node_b, const int64_t* out_row,
                                int64_t n_rows, int32_t symmetrize, float* out, void* stream) {
  HGB_DEVICE_GUARD(out);
  HGB_CHECK_ARG(plan && raw && z && out, "hgb_ham_finalize: NULL argument");
  hgb::TimeScope ts_(HGB_K_HAM_FINALIZE, stream);
  HGB_CHECK_ARG(n_rows >= 0 && n_rows < (1ll << 31), "hgb_ham_finalize: bad row count");
  if (n_rows == 0) return 0;
  FinArgs a;
  a.plan = *plan; a.raw = raw; a.partner = partner; a.h0 = h0; a.z = z; a.node_a = node_a; a.node_b = node_b;
  a.out_row = out_row; a.n_rows = n_rows; a.symmetrize = symmetrize; a.out = out;
  ham_finalize_kernel<<<(unsigned)n_rows, HT, 0, (cudaStream_t)stream>>>(a);
  HGB_LAUNCH_OK("ham_finalize_kernel");
  return 0;
}

extern "C" int hgb_ham_finalize_su2(int32_t nao, const uint8_t* orb_mask, const float* raw, const int64_t* partner,
                                    const float* h0_re, const float* h0_im, const int64_t* z, const int64_t* node_a,
                                    const int64_t* node_b, const int64_t* out_row, int64_t n_rows, int32_t symmetrize,
                                    float* out_re, float* out_im, void* stream) {
  HGB_DEVICE_GUARD(out_re);
  HGB_CHECK_ARG(orb_mask && raw && z && out_re && out_im, "hgb_ham_finalize_su2: NULL argument");
  HGB_CHECK_ARG(nao > 0 && nao <= 64, "hgb_ham_finalize_su2: bad nao %d", nao);
  HGB_CHECK_ARG(n_rows >= 0 && n_rows < (1ll << 31), "hgb_ham_finalize_su2: bad row count");
  if (n_rows == 0) return 0;
  SocFinArgs a;
  a.nao = nao; a.orb_mask = orb_mask; a.raw = raw; a.partner = partner; a.h0_re = h0_re; a.h0_im = h0_im; a.z = z;
  a.node_a = node_a; a.node_b = node_b; a.out_row = out_row; a.n_rows = n_rows; a.symmetrize = symmetrize;
  a.out_re = out_re; a.out_im = out_im;
  ham_finalize_su2_kernel<<<(unsigned)n_rows, HT, 0, (cudaStream_t)stream>>>(a);
  HGB_LAUNCH_OK("ham_finalize_su2_kernel");
  return 0;
}

extern "C" int hgb_ksi_shell_average(int32_t nao, const int32_t* blk_lo_host, const int32_t* blk_hi_host, int32_t n_blocks,
                                     float* ksi, int64_t n_rows, void* stream) {
  HGB_DEVICE_GUARD(ksi);
  HGB_CHECK_ARG(ksi && (n_blocks == 0 || (blk_lo_host && blk_hi_host)), "hgb_ksi_shell_average: NULL argument");
  HGB_CHECK_ARG(nao > 0 && nao <= 64 && n_blocks >= 0 && n_blocks <= 8, "hgb_ksi_shell_average: bad sizes");
  if (n_rows == 0 || n_blocks == 0) return 0;
  KsiArgs a;
  memset(&a, 0, sizeof(a));
  a.nao = nao; a.n_blocks = n_blocks; a.ksi = ksi; a.n_rows = n_rows;
  for (int b = 0; b < n_blocks; ++b) {
    HGB_CHECK_ARG(blk_lo_host[b] >= 0 && blk_lo_host[b] < blk_hi_host[b] && blk_hi_host[b] <= nao, "hgb_ksi_shell_average: bad shell %d", b);
    a.blk_lo[b] = blk_lo_host[b]; a.blk_hi[b] = blk_hi_host[b];
  }
  ksi_shell_average_kernel<<<(unsigned)n_rows, HT, (size_t)nao * nao * sizeof(float), (cudaStream_t)stream>>>(a);
  HGB_LAUNCH_OK("ksi_shell_average_kernel");
  return 0;
}

extern "C" int hgb_ham_finalize_so3(int32_t nao, const float* hns, const float* ksi, const float* lmat,
                                    const int64_t* partner, const float* h0_re, const float* h0_im,
                                    const int64_t* out_row, int64_t n_rows, int32_t symmetrize, int32_t h0_offdiag_only,
                                    float* out_re, float* out_im, void* stream) {
  HGB_DEVICE_GUARD(out_re);
  HGB_CHECK_ARG(hns && ksi && lmat && out_re && out_im, "hgb_ham_finalize_so3: NULL argument");
  HGB_CHECK_ARG(nao > 0 && nao <= 64, "hgb_ham_finalize_so3: bad nao %d", nao);
  HGB_CHECK_ARG(n_rows >= 0 && n_rows < (1ll << 31), "hgb_ham_finalize_so3: bad row count");
  if (n_rows == 0) return 0;
  So3Args a;
  a.nao = nao; a.hns = hns; a.ksi = ksi; a.lmat = lmat; a.partner = partner; a.h0_re = h0_re; a.h0_im = h0_im;
  a.out_row = out_row; a.n_rows = n_rows; a.symmetrize = symmetrize; a.h0_offdiag_only = h0_offdiag_only;
  a.out_re = out_re; a.out_im = out_im;
  ham_finalize_so3_kernel<<<(unsigned)n_rows, HT, 0, (cudaStream_t)stream>>>(a);
  HGB_LAUNCH_OK("ham_finalize_so3_kernel");
  return 0;
}
