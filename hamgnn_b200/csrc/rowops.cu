// Row-wise equivariant ops: o3.Linear, ResidualBlock (Linear -> e3nn Gate -> Linear + residual) and HamLayer's
// trailing Linear.  A CTA stages TR rows in shared memory and keeps every intermediate (the 1012-wide gate
// input, the gated row) on chip; HBM traffic per row is one read of x (+extra) and one write of y.
#include "hgb_common.cuh"

namespace {

constexpr int RO_THREADS = 256;
constexpr int TR = 8;  // rows per CTA

// sout[r][out_off + w*dim + k] += sum_u sin[r][in_off + u*dim + k] * W[w_off + u*mul_out + w]
__device__ __forceinline__ void lin_apply(const hgb_linblock_t* __restrict__ blocks, int nb,
                                          const float* __restrict__ w, const float* sin, int ldin, float* sout,
                                          int ldout, int nrows) {
  for (int b = 0; b < nb; ++b) {
    const hgb_linblock_t B = blocks[b];
    const int per_row = B.mul_out * B.dim;
    const int total = nrows * per_row;
    for (int idx = threadIdx.x; idx < total; idx += RO_THREADS) {
      const int r = idx / per_row;
      const int rem = idx - r * per_row;
      const int wc = rem / B.dim, k = rem - wc * B.dim;
      const float* xi = sin + (size_t)r * ldin + B.in_off + k;
      const float* wp = w + B.w_off + wc;
      float acc = 0.f;
      for (int u = 0; u < B.mul_in; ++u) acc = fmaf(xi[u * B.dim], __ldg(wp + (size_t)u * B.mul_out), acc);
      sout[(size_t)r * ldout + B.out_off + rem] += acc;
    }
    __syncthreads();  // blocks feeding the same output slot are serialised
  }
}

struct LinArgs {
  hgb_linear_plan plan;
  const float* x;
  const int64_t* rows;
  int64_t n_rows;
  float* y;
  int accumulate;
};

__global__ void __launch_bounds__(RO_THREADS) linear_kernel(const __grid_constant__ LinArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int din = a.plan.in_dim, dout = a.plan.out_dim;
  float* sx = smem;
  float* sy = smem + TR * din;
  const int64_t r0 = (int64_t)blockIdx.x * TR;
  const int nr = (int)min((int64_t)TR, a.n_rows - r0);
  for (int idx = threadIdx.x; idx < nr * din; idx += RO_THREADS) {
    const int r = idx / din, c = idx - r * din;
    const int64_t row = a.rows ? a.rows[r0 + r] : (r0 + r);
    sx[idx] = a.x[row * din + c];
  }
  for (int idx = threadIdx.x; idx < nr * dout; idx += RO_THREADS)
    sy[idx] = a.accumulate ? a.y[r0 * dout + idx] : 0.f;
  __syncthreads();
  lin_apply(a.plan.blocks, a.plan.n_blocks, a.plan.w, sx, din, sy, dout, nr);
  for (int idx = threadIdx.x; idx < nr * dout; idx += RO_THREADS) a.y[r0 * dout + idx] = sy[idx];
}

struct ResArgs {
  hgb_linear_plan lin1, lin2, post;
  hgb_gate_desc gate;
  int has_post;
  const float* x;
  const float* extra;
  int64_t n_rows;
  float* y;
  int wide;  // max(gate.in_dim, post.out_dim)
};

__global__ void __launch_bounds__(RO_THREADS) resblock_kernel(const __grid_constant__ ResArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int D = a.lin1.in_dim;
  const int DG = a.gate.in_dim;
  float* sx = smem;                 // [TR][D]   x, later y
  float* sh = sx + TR * D;          // [TR][wide] gate input, later post output
  float* sa = sh + TR * a.wide;     // [TR][D]   gated row
  const int64_t r0 = (int64_t)blockIdx.x * TR;
  const int nr = (int)min((int64_t)TR, a.n_rows - r0);
  for (int idx = threadIdx.x; idx < nr * D; idx += RO_THREADS) sx[idx] = a.x[r0 * D + idx];
  for (int idx = threadIdx.x; idx < nr * DG; idx += RO_THREADS) sh[idx] = 0.f;
  __syncthreads();
  lin_apply(a.lin1.blocks, a.lin1.n_blocks, a.lin1.w, sx, D, sh, DG, nr);
  // ---- e3nn Gate
  const hgb_gate_desc& g = a.gate;
  for (int s = 0; s < g.n_scalar_slots; ++s) {
    const int n = g.sc_n[s];
    for (int idx = threadIdx.x; idx < nr * n; idx += RO_THREADS) {
      const int r = idx / n, c = idx - r * n;
      const float v = sh[r * DG + g.sc_in_off[s] + c];
      sa[r * D + g.sc_out_off[s] + c] = (g.sc_act[s] == 0) ? hgb::ssp_f(v) * g.c_ssp : tanhf(v) * g.c_tanh;
    }
  }
  for (int s = 0; s < g.n_gated; ++s) {
    const int per = g.gd_mul[s] * g.gd_dim[s];
    for (int idx = threadIdx.x; idx < nr * per; idx += RO_THREADS) {
      const int r = idx / per, c = idx - r * per;
      const int u = c / g.gd_dim[s];
      const float gate = hgb::ssp_f(sh[r * DG + g.gd_gate_off[s] + u]) * g.c_ssp;
      sa[r * D + g.gd_out_off[s] + c] = sh[r * DG + g.gd_in_off[s] + c] * gate;
    }
  }
  if (a.extra)
    for (int idx = threadIdx.x; idx < nr * D; idx += RO_THREADS) sx[idx] += a.extra[r0 * D + idx];
  __syncthreads();
  lin_apply(a.lin2.blocks, a.lin2.n_blocks, a.lin2.w, sa, D, sx, D, nr);  // sx = x (+extra) + Lin2(gated)
  if (!a.has_post) {
    for (int idx = threadIdx.x; idx < nr * D; idx += RO_THREADS) a.y[r0 * D + idx] = sx[idx];
    return;
  }
  const int DP = a.post.out_dim;
  for (int idx = threadIdx.x; idx < nr * DP; idx += RO_THREADS) sh[idx] = 0.f;
  __syncthreads();
  lin_apply(a.post.blocks, a.post.n_blocks, a.post.w, sx, D, sh, DP, nr);
  for (int idx = threadIdx.x; idx < nr * DP; idx += RO_THREADS) a.y[r0 * DP + idx] = sh[idx];
}

}  // namespace

extern "C" int hgb_linear_forward(const hgb_linear_plan* plan, const float* x, const int64_t* rows, int64_t n_rows,
                                  float* y, int32_t accumulate, void* stream) {
  HGB_CHECK_ARG(plan && x && y, "hgb_linear_forward: NULL argument");
  HGB_CHECK_ARG(n_rows >= 0, "hgb_linear_forward: negative row count");
  if (n_rows == 0) return 0;
  LinArgs a;
  a.plan = *plan; a.x = x; a.rows = rows; a.n_rows = n_rows; a.y = y; a.accumulate = accumulate;
  const size_t smem = (size_t)TR * (plan->in_dim + plan->out_dim) * sizeof(float);
  HGB_CHECK_ARG(smem <= 200 * 1024, "hgb_linear_forward: rows of %d+%d floats do not fit in shared memory", plan->in_dim, plan->out_dim);
  HGB_CUDA_OK(cudaFuncSetAttribute(linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  linear_kernel<<<(unsigned)((n_rows + TR - 1) / TR), RO_THREADS, smem, (cudaStream_t)stream>>>(a);
  HGB_LAUNCH_OK("linear_kernel");
  return 0;
}

extern "C" int hgb_resblock_forward(const hgb_linear_plan* lin1, const hgb_gate_desc* gate, const hgb_linear_plan* lin2,
                                    const hgb_linear_plan* post, const float* x, const float* extra, int64_t n_rows,
                                    float* y, void* stream) {
  HGB_CHECK_ARG(lin1 && gate && lin2 && x && y, "hgb_resblock_forward: NULL argument");
  HGB_CHECK_ARG(lin1->out_dim == gate->in_dim && lin2->in_dim == gate->out_dim && lin2->out_dim == lin1->in_dim,
                "hgb_resblock_forward: inconsistent dims lin1 %d->%d gate %d->%d lin2 %d->%d", lin1->in_dim, lin1->out_dim,
                gate->in_dim, gate->out_dim, lin2->in_dim, lin2->out_dim);
  HGB_CHECK_ARG(gate->out_dim == lin1->in_dim, "hgb_resblock_forward: gate output dim %d != feature dim %d", gate->out_dim, lin1->in_dim);
  HGB_CHECK_ARG(!post || post->in_dim == lin1->in_dim, "hgb_resblock_forward: post Linear input dim mismatch");
  if (n_rows == 0) return 0;
  ResArgs a;
  memset(&a, 0, sizeof(a));
  a.lin1 = *lin1; a.lin2 = *lin2; a.gate = *gate; a.has_post = post ? 1 : 0;
  if (post) a.post = *post;
  a.x = x; a.extra = extra; a.n_rows = n_rows; a.y = y;
  a.wide = gate->in_dim;
  if (post && post->out_dim > a.wide) a.wide = post->out_dim;
  const size_t smem = (size_t)TR * (2 * lin1->in_dim + a.wide) * sizeof(float);
  HGB_CHECK_ARG(smem <= 200 * 1024, "hgb_resblock_forward: row tile does not fit in shared memory");
  HGB_CUDA_OK(cudaFuncSetAttribute(resblock_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  resblock_kernel<<<(unsigned)((n_rows + TR - 1) / TR), RO_THREADS, smem, (cudaStream_t)stream>>>(a);
  HGB_LAUNCH_OK("resblock_kernel");
  return 0;
}
