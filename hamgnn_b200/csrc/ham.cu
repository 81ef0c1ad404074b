// a14 + a15: Clebsch-Gordan assembly of the nao x nao orbital block from the HamLayer coefficients,
// orbital reorder (folded into the CSR row order on the host), inverse-edge symmetrisation, +H0,
// orbital masks, and placement into the per-crystal interleaved output rows.  HBM-bound stages.
#include "hgb_common.cuh"

namespace {

constexpr int HT = 256;

struct AsmArgs {
  hgb_ham_plan plan;
  const float* coef;
  int64_t n_rows;
  float* raw;
};

// one CTA per HROWS rows; coefficients staged in shared memory, every thread owns matrix entries q
constexpr int HROWS = 8;
__global__ void __launch_bounds__(HT) ham_assemble_kernel(const __grid_constant__ AsmArgs a) {
  extern __shared__ float sc[];
  const int nc = a.plan.n_coef, nn = a.plan.nao * a.plan.nao;
  const int64_t r0 = (int64_t)blockIdx.x * HROWS;
  const int nr = (int)min((int64_t)HROWS, a.n_rows - r0);
  for (int idx = threadIdx.x; idx < nr * nc; idx += HT) sc[idx] = a.coef[r0 * nc + idx];
  __syncthreads();
  for (int q = threadIdx.x; q < nn; q += HT) {
    const int n0 = a.plan.row_ptr[q], n1 = a.plan.row_ptr[q + 1];
    float acc[HROWS];
#pragma unroll
    for (int r = 0; r < HROWS; ++r) acc[r] = 0.f;
    for (int n = n0; n < n1; ++n) {
      const float v = a.plan.val[n];
      const int c = a.plan.col[n];
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

}  // namespace

extern "C" int hgb_ham_assemble(const hgb_ham_plan* plan, const float* coef, int64_t n_rows, float* raw, void* stream) {
  HGB_CHECK_ARG(plan && coef && raw, "hgb_ham_assemble: NULL argument");
  HGB_CHECK_ARG(plan->nao > 0 && plan->nao <= 64 && plan->n_coef > 0, "hgb_ham_assemble: bad plan (nao=%d)", plan->nao);
  if (n_rows == 0) return 0;
  AsmArgs a;
  a.plan = *plan; a.coef = coef; a.n_rows = n_rows; a.raw = raw;
  const size_t smem = (size_t)HROWS * plan->n_coef * sizeof(float);
  HGB_CUDA_OK(cudaFuncSetAttribute(ham_assemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ham_assemble_kernel<<<(unsigned)((n_rows + HROWS - 1) / HROWS), HT, smem, (cudaStream_t)stream>>>(a);
  HGB_LAUNCH_OK("ham_assemble_kernel");
  return 0;
}

extern "C" int hgb_ham_finalize(const hgb_ham_plan* plan, const float* raw, const int64_t* partner, const float* h0,
                                const int64_t* z, const int64_t* node_a, const int64_t* node_b, const int64_t* out_row,
                                int64_t n_rows, int32_t symmetrize, float* out, void* stream) {
  HGB_CHECK_ARG(plan && raw && z && out, "hgb_ham_finalize: NULL argument");
  HGB_CHECK_ARG(n_rows >= 0 && n_rows < (1ll << 31), "hgb_ham_finalize: bad row count");
  if (n_rows == 0) return 0;
  FinArgs a;
  a.plan = *plan; a.raw = raw; a.partner = partner; a.h0 = h0; a.z = z; a.node_a = node_a; a.node_b = node_b;
  a.out_row = out_row; a.n_rows = n_rows; a.symmetrize = symmetrize; a.out = out;
  ham_finalize_kernel<<<(unsigned)n_rows, HT, 0, (cudaStream_t)stream>>>(a);
  HGB_LAUNCH_OK("ham_finalize_kernel");
  return 0;
}
