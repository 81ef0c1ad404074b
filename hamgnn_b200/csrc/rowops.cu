// Row-wise equivariant ops: o3.Linear, ResidualBlock (Linear -> e3nn Gate -> Linear + residual) and HamLayer's
// trailing Linear.  A CTA stages TR rows in shared memory and keeps every intermediate (the 1012-wide gate
// input, the gated row) on chip; HBM traffic per row is one read of x (+extra) and one write of y.
//
// Tiles are stored column-major in shared memory ([column][LDR], LDR = TR + 4 floats) so that a thread which
// owns one output column (slot, channel w, component k) reads the TR row values of an input column with two
// 128-bit loads and keeps TR accumulators in registers: per input channel u it issues 1 weight load + 2 shared
// loads for 8 FMAs (the stride of 12 floats keeps the quarter-warp 128-bit accesses on distinct banks).
#include <stdlib.h>
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


// ------------------------------------------------------------------------------------------------------------------
// resblock2_kernel: the same Linear -> Gate -> Linear (+ residual) [-> Linear] chain on 16-row tiles.
//
// resblock_kernel (8 rows per CTA, one thread per output column, weights read through L1/L2 per input channel) ran the
// off-site HamLayer of tbg_m8 at 51 GB/s = 0.8 % of the HBM roofline: every thread's inner loop is a
// chain of dependent weight loads with 8 warps per SM to hide them.  Here
//   * the row tiles are row-major in shared memory ([16][dim], three of them = 177 KB for D = 877);
//   * a Linear block is evaluated by items (output chunk of 8 channels, row, component): the block's weights are staged
//     once per CTA in shared memory ([u][mul_out padded to 4]) and read as warp-broadcast float4, the input value is one
//     conflict-free LDS (row stride = dim is odd for the irreps in use) -- 8 FMAs per 3 shared loads, no global load in the loop;
//   * 512 threads per CTA.
// <RB2_ROWS, RB2_THREADS, RB2_WFLOATS>: <16, 512, 8192> = 209 KB, one CTA per SM; <8, 256, 4096> = 105 KB, two CTAs per SM (one
// stages weights while the other computes)

template <int RB2_ROWS, int RB2_THREADS, int RB2_WFLOATS>
__device__ __forceinline__ void lin_apply2(const hgb_linblock_t* __restrict__ blocks, int nb, const float* __restrict__ w,
                                           const float* sin, int ldin, float* sout, int ldout, float* sw) {
  for (int b = 0; b < nb; ++b) {
    const hgb_linblock_t B = blocks[b];
    const int mo4 = (B.mul_out + 3) & ~3;
    const int rows_per_chunk = min(B.mul_in, RB2_WFLOATS / mo4);
    const int nwc = (B.mul_out + 7) >> 3;
    const int per_wc = RB2_ROWS * B.dim;
    const int items = nwc * per_wc;
    for (int u0 = 0; u0 < B.mul_in; u0 += rows_per_chunk) {
      const int rc = min(rows_per_chunk, B.mul_in - u0);
      __syncthreads();   // the previous chunk's readers are done with sw; the previous block's writers with sout
      for (int idx = threadIdx.x; idx < rc * mo4; idx += RB2_THREADS) {
        const int ur = idx / mo4, wc = idx - ur * mo4;
        sw[idx] = (wc < B.mul_out) ? __ldg(w + B.w_off + (size_t)(u0 + ur) * B.mul_out + wc) : 0.f;
      }
      __syncthreads();
      for (int it = threadIdx.x; it < items; it += RB2_THREADS) {
        const int wc = it / per_wc, rk = it - wc * per_wc;
        const int r = rk / B.dim, k = rk - r * B.dim;
        const float* xi = sin + (size_t)r * ldin + B.in_off + (size_t)u0 * B.dim + k;
        const float* wr = sw + wc * 8;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        const bool two = wc * 8 + 4 < mo4;   // the last chunk of 8 may hold one float4 only
#pragma unroll 4
        for (int ur = 0; ur < rc; ++ur) {
          const float xv = xi[(size_t)ur * B.dim];
          const float4 w0 = *reinterpret_cast<const float4*>(wr + ur * mo4);
          acc[0] = fmaf(xv, w0.x, acc[0]); acc[1] = fmaf(xv, w0.y, acc[1]);
          acc[2] = fmaf(xv, w0.z, acc[2]); acc[3] = fmaf(xv, w0.w, acc[3]);
          if (two) {
            const float4 w1 = *reinterpret_cast<const float4*>(wr + ur * mo4 + 4);
            acc[4] = fmaf(xv, w1.x, acc[4]); acc[5] = fmaf(xv, w1.y, acc[5]);
            acc[6] = fmaf(xv, w1.z, acc[6]); acc[7] = fmaf(xv, w1.w, acc[7]);
          }
        }
        float* o = sout + (size_t)r * ldout + B.out_off + (size_t)(wc * 8) * B.dim + k;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (wc * 8 + j < B.mul_out) o[(size_t)j * B.dim] += acc[j];
      }
    }
  }
  __syncthreads();
}

template <int RB2_ROWS, int RB2_THREADS>
__device__ __forceinline__ void load_rows2(float* s, int lds, const float* __restrict__ g, int64_t r0, int nr, int dim, bool add) {
  for (int idx = threadIdx.x; idx < RB2_ROWS * dim; idx += RB2_THREADS) {
    const int r = idx / dim, c = idx - r * dim;
    const float v = (r < nr) ? __ldg(g + (r0 + r) * (int64_t)dim + c) : 0.f;
    if (add) s[(size_t)r * lds + c] += v;
    else s[(size_t)r * lds + c] = v;
  }
}

template <int RB2_ROWS, int RB2_THREADS, int RB2_WFLOATS>
__global__ void __launch_bounds__(RB2_THREADS, (RB2_ROWS <= 8 ? 2 : 1)) resblock2_kernel(const __grid_constant__ ResArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int D = a.lin1.in_dim;
  const int ldh = a.wide | 1;   // odd row stride: conflict-free column access across rows
  float* sx = smem;                              // [16][D]    x, later y = x (+ extra) + Lin2(gated)
  float* sh = sx + (size_t)RB2_ROWS * D;         // [16][wide] gate input, later the post-Linear output
  float* sa = sh + (size_t)RB2_ROWS * ldh;    // [16][D]    gated row
  float* sw = sa + (size_t)RB2_ROWS * D;         // weight staging
  const int64_t r0 = (int64_t)blockIdx.x * RB2_ROWS;
  const int nr = (int)min((int64_t)RB2_ROWS, a.n_rows - r0);
  load_rows2<RB2_ROWS, RB2_THREADS>(sx, D, a.x, r0, nr, D, false);
  for (int idx = threadIdx.x; idx < RB2_ROWS * ldh; idx += RB2_THREADS) sh[idx] = 0.f;
  lin_apply2<RB2_ROWS, RB2_THREADS, RB2_WFLOATS>(a.lin1.blocks, a.lin1.n_blocks, a.lin1.w, sx, D, sh, ldh, sw);   // starts with a barrier
  // ---- e3nn Gate, element-wise
  const hgb_gate_desc& g = a.gate;
  for (int s = 0; s < g.n_scalar_slots; ++s) {
    const int n = g.sc_n[s];
    for (int idx = threadIdx.x; idx < n * RB2_ROWS; idx += RB2_THREADS) {
      const int r = idx / n, c = idx - r * n;
      const float v = sh[(size_t)r * ldh + g.sc_in_off[s] + c];
      sa[(size_t)r * D + g.sc_out_off[s] + c] = (g.sc_act[s] == 0) ? hgb::ssp_f(v) * g.c_ssp : tanhf(v) * g.c_tanh;
    }
  }
  for (int s = 0; s < g.n_gated; ++s) {
    const int per = g.gd_mul[s] * g.gd_dim[s];
    for (int idx = threadIdx.x; idx < per * RB2_ROWS; idx += RB2_THREADS) {
      const int r = idx / per, c = idx - r * per;
      const int u = c / g.gd_dim[s];
      const float gate = hgb::ssp_f(sh[(size_t)r * ldh + g.gd_gate_off[s] + u]) * g.c_ssp;
      sa[(size_t)r * D + g.gd_out_off[s] + c] = sh[(size_t)r * ldh + g.gd_in_off[s] + c] * gate;
    }
  }
  if (a.extra) load_rows2<RB2_ROWS, RB2_THREADS>(sx, D, a.extra, r0, nr, D, true);
  lin_apply2<RB2_ROWS, RB2_THREADS, RB2_WFLOATS>(a.lin2.blocks, a.lin2.n_blocks, a.lin2.w, sa, D, sx, D, sw);   // sx = x (+ extra) + Lin2(gated)
  if (!a.has_post) {
    for (int idx = threadIdx.x; idx < nr * D; idx += RB2_THREADS) a.y[r0 * D + idx] = sx[idx];
    return;
  }
  const int DP = a.post.out_dim;
  for (int idx = threadIdx.x; idx < RB2_ROWS * ldh; idx += RB2_THREADS) sh[idx] = 0.f;
  lin_apply2<RB2_ROWS, RB2_THREADS, RB2_WFLOATS>(a.post.blocks, a.post.n_blocks, a.post.w, sx, D, sh, ldh, sw);
  for (int idx = threadIdx.x; idx < nr * DP; idx += RB2_THREADS) {
    const int r = idx / DP, c = idx - r * DP;
    a.y[(r0 + r) * DP + c] = sh[(size_t)r * ldh + c];
  }
}


// out[i][c] = sum of rows[order[j]][c] for j in [ptr[i], ptr[i+1]), added serially in list order: the deterministic receiver
// reduction (torch_scatter.scatter(..., reduce='sum'), hamgnn/nn/convolution.py:147-149) for message kernels that write one row
// per edge.  One CTA per output row, a thread per column (coalesced row reads), four rows in flight per thread.
__global__ void __launch_bounds__(256) segment_sum_kernel(const float* __restrict__ rows, int n_cols, const int64_t* __restrict__ ptr,
                                                          const int64_t* __restrict__ order, float* __restrict__ out) {
  const int64_t i = blockIdx.x;
  const int64_t j0 = ptr[i], j1 = ptr[i + 1];
  for (int c = threadIdx.x; c < n_cols; c += 256) {
    float acc = 0.f;
    int64_t j = j0;
    for (; j + 4 <= j1; j += 4) {
      const float v0 = __ldg(rows + order[j] * n_cols + c), v1 = __ldg(rows + order[j + 1] * n_cols + c);
      const float v2 = __ldg(rows + order[j + 2] * n_cols + c), v3 = __ldg(rows + order[j + 3] * n_cols + c);
      acc += v0; acc += v1; acc += v2; acc += v3;   // list order
    }
    for (; j < j1; ++j) acc += __ldg(rows + order[j] * n_cols + c);
    out[i * n_cols + c] = acc;
  }
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
  hgb::TimeScope ts_(HGB_K_LINEAR, stream);
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
  hgb::TimeScope ts_(HGB_K_RESBLOCK, stream);
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
  // row tiles with staged weights when they fit in shared memory, else the 8-row kernel
  const char* force_old = getenv("HGB_RESBLOCK_V1");
  const char* rows_env = getenv("HGB_RESBLOCK_ROWS");
  const int want_rows = rows_env ? atoi(rows_env) : 16;   // measured: 8 rows x 2 CTAs per SM is no faster (profiles/README.md r03a)
  const size_t tile_floats = (size_t)(2 * lin1->in_dim + (a.wide | 1));
  if (!(force_old && force_old[0] == '1')) {
    const size_t smem8 = (8 * tile_floats + 4096) * sizeof(float), smem16 = (16 * tile_floats + 8192) * sizeof(float);
    if (want_rows == 8 && 2 * (smem8 + 1024) <= 227 * 1024) {
      HGB_CUDA_OK(cudaFuncSetAttribute(resblock2_kernel<8, 256, 4096>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem8));
      resblock2_kernel<8, 256, 4096><<<(unsigned)((n_rows + 7) / 8), 256, smem8, (cudaStream_t)stream>>>(a);
      HGB_LAUNCH_OK("resblock2_kernel");
      return 0;
    }
    if (smem16 <= 227 * 1024) {
      HGB_CUDA_OK(cudaFuncSetAttribute(resblock2_kernel<16, 512, 8192>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));
      resblock2_kernel<16, 512, 8192><<<(unsigned)((n_rows + 15) / 16), 512, smem16, (cudaStream_t)stream>>>(a);
      HGB_LAUNCH_OK("resblock2_kernel");
      return 0;
    }
  }
  const size_t smem = (size_t)LDR * (2 * lin1->in_dim + a.wide) * sizeof(float);
  HGB_CHECK_ARG(smem <= 220 * 1024, "hgb_resblock_forward: row tile does not fit in shared memory");
  HGB_CUDA_OK(cudaFuncSetAttribute(resblock_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  resblock_kernel<<<(unsigned)((n_rows + TR - 1) / TR), RO_THREADS, smem, (cudaStream_t)stream>>>(a);
  HGB_LAUNCH_OK("resblock_kernel");
  return 0;
}

extern "C" int hgb_segment_sum(const float* rows, int32_t n_cols, const int64_t* seg_ptr, const int64_t* seg_order,
                               int64_t n_out_rows, float* out, void* stream) {
  HGB_DEVICE_GUARD(out);
  HGB_CHECK_ARG(rows && seg_ptr && seg_order && out, "hgb_segment_sum: NULL argument");
  HGB_CHECK_ARG(n_cols > 0 && n_out_rows >= 0 && n_out_rows < (1ll << 31), "hgb_segment_sum: bad sizes");
  if (n_out_rows == 0) return 0;
  hgb::TimeScope ts_(HGB_K_SEGMENT_SUM, stream);
  segment_sum_kernel<<<(unsigned)n_out_rows, 256, 0, (cudaStream_t)stream>>>(rows, n_cols, seg_ptr, seg_order, out);
  HGB_LAUNCH_OK("segment_sum_kernel");
  return 0;
}
