// Row-wise equivariant ops: o3.Linear, ResidualBlock (Linear -> e3nn Gate -> Linear + residual) and HamLayer's
// trailing Linear.  A CTA stages TR rows in shared memory and keeps every intermediate (the 1012-wide gate
// input, the gated row) on chip; HBM traffic per row is one read of x (+extra) and one write of y.
//
// Tiles are stored column-major in shared memory ([column][LDR], LDR = TR + 4 floats) so that a thread which
// owns one output column (slot, channel w, component k) reads the TR row values of an input column with two
// 128-bit loads and keeps TR accumulators in registers: per input channel u it issues 1 weight load + 2 shared
// loads for 8 FMAs (the stride of 12 floats keeps the quarter-warp 128-bit accesses on distinct banks).
#include "hgb_common.cuh"

namespace {

constexpr int RO_THREADS = 256;
constexpr int TR = 8;         // rows per CTA
constexpr int LDR = TR + 4;   // column stride in the transposed tile

// one output column (block B, flat index idx = w*dim + k) for all TR rows
__device__ __forceinline__ void lin_item(const hgb_linblock_t& B, int idx, const float* __restrict__ w, const float* sin,
                                         float* sout) {
  const int wc = idx / B.dim, k = idx - wc * B.dim;
  const float* xi = sin + (size_t)(B.in_off + k) * LDR;
  const float* wp = w + B.w_off + wc;
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
#pragma unroll 4
  for (int u = 0; u < B.mul_in; ++u) {
    const float wv = __ldg(wp + (size_t)u * B.mul_out);
    const float4 x0 = *reinterpret_cast<const float4*>(xi + (size_t)u * B.dim * LDR);
    const float4 x1 = *reinterpret_cast<const float4*>(xi + (size_t)u * B.dim * LDR + 4);
    a0.x = fmaf(x0.x, wv, a0.x); a0.y = fmaf(x0.y, wv, a0.y); a0.z = fmaf(x0.z, wv, a0.z); a0.w = fmaf(x0.w, wv, a0.w);
    a1.x = fmaf(x1.x, wv, a1.x); a1.y = fmaf(x1.y, wv, a1.y); a1.z = fmaf(x1.z, wv, a1.z); a1.w = fmaf(x1.w, wv, a1.w);
  }
  float4* o = reinterpret_cast<float4*>(sout + (size_t)(B.out_off + idx) * LDR);
  float4 c0 = o[0], c1 = o[1];
  c0.x += a0.x; c0.y += a0.y; c0.z += a0.z; c0.w += a0.w;
  c1.x += a1.x; c1.y += a1.y; c1.z += a1.z; c1.w += a1.w;
  o[0] = c0; o[1] = c1;
}

// sout[col(out_off + w*dim + k)][r] += sum_u sin[col(in_off + u*dim + k)][r] * W[w_off + u*mul_out + w]
// disjoint != 0: every block writes its own output slot -> one pass over the flattened (block, column) items;
// otherwise blocks are serialised (several input slots feeding one output slot).
__device__ __forceinline__ void lin_apply(const hgb_linblock_t* __restrict__ blocks, int nb, int disjoint,
                                          const float* __restrict__ w, const float* sin, float* sout) {
  __shared__ int spre[65];
  if (disjoint && nb <= 64) {
    if (threadIdx.x == 0) {
      int acc = 0;
      for (int b = 0; b < nb; ++b) { spre[b] = acc; acc += blocks[b].mul_out * blocks[b].dim; }
      spre[nb] = acc;
    }
    __syncthreads();
    const int total = spre[nb];
    for (int it = threadIdx.x; it < total; it += RO_THREADS) {
      int b = 0;
      while (it >= spre[b + 1]) ++b;
      lin_item(blocks[b], it - spre[b], w, sin, sout);
    }
    __syncthreads();
    return;
  }
  for (int b = 0; b < nb; ++b) {
    const hgb_linblock_t B = blocks[b];
    const int per_row = B.mul_out * B.dim;
    for (int idx = threadIdx.x; idx < per_row; idx += RO_THREADS) lin_item(B, idx, w, sin, sout);
    __syncthreads();
  }
}

// global [rows][dim] (optionally gathered) -> transposed tile; rows >= nr are zero-filled
__device__ __forceinline__ void load_tile(float* s, const float* __restrict__ g, const int64_t* rows, int64_t r0, int nr,
                                          int dim, bool add, int64_t ld = -1) {
  if (ld < 0) ld = dim;
  for (int idx = threadIdx.x; idx < TR * dim; idx += RO_THREADS) {
    const int r = idx / dim, c = idx - r * dim;
    float v = 0.f;
    if (r < nr) {
      const int64_t row = rows ? rows[r0 + r] : (r0 + r);
      v = g[row * ld + c];
    }
    if (add) s[(size_t)c * LDR + r] += v;
    else s[(size_t)c * LDR + r] = v;
  }
}
__device__ __forceinline__ void store_tile(const float* s, float* __restrict__ g, int64_t r0, int nr, int dim, int64_t ld = -1) {
  if (ld < 0) ld = dim;
  for (int idx = threadIdx.x; idx < nr * dim; idx += RO_THREADS) {
    const int r = idx / dim, c = idx - r * dim;
    g[(r0 + r) * ld + c] = s[(size_t)c * LDR + r];
  }
}
__device__ __forceinline__ void zero_tile(float* s, int dim) {
  for (int idx = threadIdx.x; idx < dim * LDR; idx += RO_THREADS) s[idx] = 0.f;
}

struct LinArgs {
  hgb_linear_plan plan;
  const float* x;
  const int64_t* rows;
  int64_t n_rows;
  float* y;
  int64_t ldy;
  int accumulate;
};

__global__ void __launch_bounds__(RO_THREADS) linear_kernel(const __grid_constant__ LinArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int din = a.plan.in_dim, dout = a.plan.out_dim;
  float* sx = smem;
  float* sy = smem + (size_t)din * LDR;
  const int64_t r0 = (int64_t)blockIdx.x * TR;
  const int nr = (int)min((int64_t)TR, a.n_rows - r0);
  load_tile(sx, a.x, a.rows, r0, nr, din, false);
  if (a.accumulate) load_tile(sy, a.y, nullptr, r0, nr, dout, false, a.ldy);
  else zero_tile(sy, dout);
  __syncthreads();
  lin_apply(a.plan.blocks, a.plan.n_blocks, a.plan.pad & 1, a.plan.w, sx, sy);
  store_tile(sy, a.y, r0, nr, dout, a.ldy);
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
  float* sx = smem;                          // [D][LDR]    x, later y
  float* sh = sx + (size_t)D * LDR;          // [wide][LDR] gate input, later post output
  float* sa = sh + (size_t)a.wide * LDR;     // [D][LDR]    gated row
  const int64_t r0 = (int64_t)blockIdx.x * TR;
  const int nr = (int)min((int64_t)TR, a.n_rows - r0);
  load_tile(sx, a.x, nullptr, r0, nr, D, false);
  zero_tile(sh, DG);
  __syncthreads();
  lin_apply(a.lin1.blocks, a.lin1.n_blocks, a.lin1.pad & 1, a.lin1.w, sx, sh);
  // ---- e3nn Gate (element-wise on [column][row])
  const hgb_gate_desc& g = a.gate;
  for (int s = 0; s < g.n_scalar_slots; ++s) {
    const int n = g.sc_n[s];
    for (int idx = threadIdx.x; idx < n * TR; idx += RO_THREADS) {
      const int c = idx / TR, r = idx - c * TR;
      const float v = sh[(size_t)(g.sc_in_off[s] + c) * LDR + r];
      sa[(size_t)(g.sc_out_off[s] + c) * LDR + r] = (g.sc_act[s] == 0) ? hgb::ssp_f(v) * g.c_ssp : tanhf(v) * g.c_tanh;
    }
  }
  for (int s = 0; s < g.n_gated; ++s) {
    const int per = g.gd_mul[s] * g.gd_dim[s];
    for (int idx = threadIdx.x; idx < per * TR; idx += RO_THREADS) {
      const int c = idx / TR, r = idx - c * TR;
      const int u = c / g.gd_dim[s];
      const float gate = hgb::ssp_f(sh[(size_t)(g.gd_gate_off[s] + u) * LDR + r]) * g.c_ssp;
      sa[(size_t)(g.gd_out_off[s] + c) * LDR + r] = sh[(size_t)(g.gd_in_off[s] + c) * LDR + r] * gate;
    }
  }
  if (a.extra) load_tile(sx, a.extra, nullptr, r0, nr, D, true);
  __syncthreads();
  lin_apply(a.lin2.blocks, a.lin2.n_blocks, a.lin2.pad & 1, a.lin2.w, sa, sx);  // sx = x (+extra) + Lin2(gated)
  if (!a.has_post) {
    store_tile(sx, a.y, r0, nr, D);
    return;
  }
  const int DP = a.post.out_dim;
  zero_tile(sh, DP);
  __syncthreads();
  lin_apply(a.post.blocks, a.post.n_blocks, a.post.pad & 1, a.post.w, sx, sh);
  store_tile(sh, a.y, r0, nr, DP);
}

}  // namespace

extern "C" int hgb_linear_forward(const hgb_linear_plan* plan, const float* x, const int64_t* rows, int64_t n_rows,
                                  float* y, int32_t accumulate, void* stream) {
  HGB_CHECK_ARG(plan, "hgb_linear_forward: NULL argument");
  return hgb_linear_forward_ld(plan, x, rows, n_rows, y, plan->out_dim, accumulate, stream);
}

extern "C" int hgb_linear_forward_ld(const hgb_linear_plan* plan, const float* x, const int64_t* rows, int64_t n_rows,
                                     float* y, int64_t ldy, int32_t accumulate, void* stream) {
  HGB_DEVICE_GUARD(y);
  HGB_CHECK_ARG(plan && x && y, "hgb_linear_forward: NULL argument");
  HGB_CHECK_ARG(n_rows >= 0, "hgb_linear_forward: negative row count");
  HGB_CHECK_ARG(ldy >= plan->out_dim, "hgb_linear_forward: output row stride %lld < out_dim %d", (long long)ldy, plan->out_dim);
  if (n_rows == 0) return 0;
  LinArgs a;
  a.plan = *plan; a.x = x; a.rows = rows; a.n_rows = n_rows; a.y = y; a.ldy = ldy; a.accumulate = accumulate;
  const size_t smem = (size_t)LDR * (plan->in_dim + plan->out_dim) * sizeof(float);
  HGB_CHECK_ARG(smem <= 220 * 1024, "hgb_linear_forward: rows of %d+%d floats do not fit in shared memory", plan->in_dim, plan->out_dim);
  HGB_CUDA_OK(cudaFuncSetAttribute(linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  linear_kernel<<<(unsigned)((n_rows + TR - 1) / TR), RO_THREADS, smem, (cudaStream_t)stream>>>(a);
  HGB_LAUNCH_OK("linear_kernel");
  return 0;
}

extern "C" int hgb_resblock_forward(const hgb_linear_plan* lin1, const hgb_gate_desc* gate, const hgb_linear_plan* lin2,
                                    const hgb_linear_plan* post, const float* x, const float* extra, int64_t n_rows,
                                    float* y, void* stream) {
  HGB_DEVICE_GUARD(y);
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
  const size_t smem = (size_t)LDR * (2 * lin1->in_dim + a.wide) * sizeof(float);
  HGB_CHECK_ARG(smem <= 220 * 1024, "hgb_resblock_forward: row tile does not fit in shared memory");
  HGB_CUDA_OK(cudaFuncSetAttribute(resblock_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  resblock_kernel<<<(unsigned)((n_rows + TR - 1) / TR), RO_THREADS, smem, (cudaStream_t)stream>>>(a);
  HGB_LAUNCH_OK("resblock_kernel");
  return 0;
}
