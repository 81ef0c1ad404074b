// Fused MessagePackBlock forward (a5-a8, with a9's receiver scatter-sum or a11's skip Linear folded in).
//
// One CTA owns a tile of TE edges and walks the output slots ("types") of irreps_out.  For a type
// t = (l3, p3) with multiplicity M and d3 = 2 l3 + 1 the rows of every matrix are the (edge z, component k)
// pairs of a sub-block of the tile (R = nz * d3 <= RMAX rows), and for every tensor-product path p -> t
//
//     T_z[i][k]   = sum_j w3j(l1,l2,l3)[i,j,k] Y_l2[z][j]                       (sparse CG list, SIMT)
//     A[u][(z,k)] = sum_i x1[z][u][i] T_z[i][k]                                  (per edge d1 x d3, SIMT)
//     B[(z,k)][w] = g[z][w] * sum_u A[u][(z,k)] W_p[u][w]                        (GEMM1, K = mul1)
//     g[z][w]     = sum_h h2[z][h] W3_p[h][w]                                    (radial MLP last layer)
//     C[(z,k)][w'] += sum_w B[(z,k)][w] L'_p[w][w']                              (GEMM2, K = M)
//
// where W_p already carries the path coefficient sqrt((2 l3+1)/mul1), and L'_p is the host-folded product
// of the `mid.simplify() -> irreps_out` o3.Linear block of LinearScaleWithWeights with the following
// node/edge_linear_out block (both fan-in normalised).  Nothing of size [E, 17523] or [E, 3589] is ever
// written: per edge the kernel reads the three 877-float input rows (through L1/L2, once per path that
// uses an irrep block), 36 SH + 64 radial floats, and writes 877 floats (or scatter-adds them).
//
// All three dense contractions run through one register-tiled routine (gemm_tiles: 4x4 outputs per
// thread-tile, A operand K-major in shared memory so that a float4 covers 4 rows, weights read as
// warp-uniform float4 through the read-only path).  fp32 FMA throughout: the 1e-5 parity bar rules out
// single-pass TF32/BF16 tensor-core products (SURVEY.md section 7 "Hard parts").
#include <stdlib.h>

#include "hgb_common.cuh"

namespace {

struct MsgArgs {
  hgb_msgpack_plan plan;
  const float* src[4];
  const int64_t* src_rows[4];
  const float* sh;
  const float* rbf;
  int64_t n_edges;
  float* out;
  const int64_t* out_index;
};

constexpr int KC = 32;        // K-chunk staged per GEMM1 step
constexpr int MPAD_MAX = 64;  // widest (padded) output multiplicity
constexpr int D_MAX = 2 * HGB_MAX_L + 1;

template <int NT, int NTL>
__device__ __forceinline__ void gemm_tiles(float (&acc)[NTL][4][4], const float* __restrict__ As, int lda, int K,
                                           int R4, int M4, const float* __restrict__ W, int ldw) {
  const int tiles = R4 * M4;
#pragma unroll
  for (int n = 0; n < NTL; ++n) {
    const int ti = threadIdx.x + n * NT;
    if (ti < tiles) {
      const int c4 = ti / R4, r4 = ti - c4 * R4;
      const float* ap = As + r4 * 4;
      const float* wp = W + c4 * 4;
#pragma unroll 4
      for (int u = 0; u < K; ++u) {
        const float4 a = *reinterpret_cast<const float4*>(ap + (size_t)u * lda);
        const float4 w = __ldg(reinterpret_cast<const float4*>(wp + (size_t)u * ldw));
        hgb::fma4x4(acc[n], a, w);
      }
    }
  }
}

// Edges per sub-block of a tile for an output slot of dimension d3: the tile is cut into the smallest number of
// near-equal pieces (multiples of 4 edges) whose row count nz*d3 fits RMAX.
__host__ __device__ inline int sub_tile_edges(int te, int rmax, int d3) {
  int nsub = (te * d3 + rmax - 1) / rmax;
  int nz = ((te + nsub - 1) / nsub + 3) & ~3;
  while (nz * d3 > rmax && nz > 4) {
    ++nsub;
    nz = ((te + nsub - 1) / nsub + 3) & ~3;
  }
  return nz;
}

template <int NTL>
__device__ __forceinline__ void zero_tiles(float (&acc)[NTL][4][4]) {
#pragma unroll
  for (int n = 0; n < NTL; ++n)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[n][i][j] = 0.f;
}

// A[u][(z,k)] for u in [u0, u0+kc): per edge a d1 x D3 contraction with T_z.  Lanes run over z so that
// every shared-memory access has an odd stride.
template <int NT, int D3>
__device__ __forceinline__ void agen(float* __restrict__ sA, int lda, const float* __restrict__ xblk, int ldx,
                                     const float* __restrict__ sT, int d1, int nz, int u0, int kc) {
  constexpr int UB = 2;
  const int nub = (kc + UB - 1) / UB;
  for (int idx = threadIdx.x; idx < nz * nub; idx += NT) {
    const int ub = idx / nz, z = idx - ub * nz;
    const int ul = ub * UB;
    float a[UB][D3];
#pragma unroll
    for (int q = 0; q < UB; ++q)
#pragma unroll
      for (int k = 0; k < D3; ++k) a[q][k] = 0.f;
    const float* xz = xblk + (size_t)z * ldx + (size_t)(u0 + ul) * d1;
    const float* tz = sT + (size_t)z * d1 * D3;
    const bool has1 = (ul + 1) < kc;
    for (int i = 0; i < d1; ++i) {
      const float x0 = xz[i];
      const float x1 = has1 ? xz[d1 + i] : 0.f;
#pragma unroll
      for (int k = 0; k < D3; ++k) {
        const float t = tz[i * D3 + k];
        a[0][k] = fmaf(x0, t, a[0][k]);
        a[1][k] = fmaf(x1, t, a[1][k]);
      }
    }
#pragma unroll
    for (int k = 0; k < D3; ++k) {
      sA[(size_t)ul * lda + z * D3 + k] = a[0][k];
      if (has1) sA[(size_t)(ul + 1) * lda + z * D3 + k] = a[1][k];
    }
  }
}

template <int NT>
__device__ __forceinline__ void agen_dispatch(int d3, float* sA, int lda, const float* xblk, int ldx, const float* sT,
                                              int d1, int nz, int u0, int kc) {
  switch (d3) {
    case 1: agen<NT, 1>(sA, lda, xblk, ldx, sT, d1, nz, u0, kc); break;
    case 3: agen<NT, 3>(sA, lda, xblk, ldx, sT, d1, nz, u0, kc); break;
    case 5: agen<NT, 5>(sA, lda, xblk, ldx, sT, d1, nz, u0, kc); break;
    case 7: agen<NT, 7>(sA, lda, xblk, ldx, sT, d1, nz, u0, kc); break;
    case 9: agen<NT, 9>(sA, lda, xblk, ldx, sT, d1, nz, u0, kc); break;
    case 11: agen<NT, 11>(sA, lda, xblk, ldx, sT, d1, nz, u0, kc); break;
    case 13: agen<NT, 13>(sA, lda, xblk, ldx, sT, d1, nz, u0, kc); break;
    case 15: agen<NT, 15>(sA, lda, xblk, ldx, sT, d1, nz, u0, kc); break;
    default: agen<NT, 17>(sA, lda, xblk, ldx, sT, d1, nz, u0, kc); break;
  }
}

// Shared-memory carve-up (floats).  xblk aliases sB: the input irrep block is dead once GEMM1 is complete.
template <int TE, int RMAX>
struct Smem {
  static constexpr int A = 0;                                   // [KC][RMAX]
  static constexpr int B = A + KC * RMAX;                       // [MPAD_MAX][RMAX]  (also xblk)
  static constexpr int T = B + MPAD_MAX * RMAX;                 // [RMAX * D_MAX]
  static constexpr int G = T + RMAX * D_MAX;                    // [TE][MPAD_MAX]
  static constexpr int H2 = G + TE * MPAD_MAX;                  // [2][64][TE]
  static constexpr int Y = H2 + 2 * 64 * TE;                    // [TE][sh_dim]  (sh_dim <= 81)
  static constexpr int ROWS = Y + TE * 84;                      // int [4][TE]
  static constexpr int TOTAL = ROWS + 4 * TE;
};

template <int TE, int RMAX, int NT>
__global__ void __launch_bounds__(NT, (NT <= 256 ? 2 : 1)) msgpack_kernel(const __grid_constant__ MsgArgs a) {
  constexpr int NTL = ((RMAX / 4) * (MPAD_MAX / 4) + NT - 1) / NT;
  using L = Smem<TE, RMAX>;
  extern __shared__ __align__(16) float smem[];
  float* sA = smem + L::A;
  float* sB = smem + L::B;
  float* sT = smem + L::T;
  float* sG = smem + L::G;
  float* sH2 = smem + L::H2;
  float* sY = smem + L::Y;
  int* sRow = reinterpret_cast<int*>(smem + L::ROWS);

  const hgb_msgpack_plan& P = a.plan;
  const int tid = threadIdx.x;
  const int64_t e0 = (int64_t)blockIdx.x * TE;
  const int ne = (int)min((int64_t)TE, a.n_edges - e0);
  const int S = P.sh_dim;
  const float* __restrict__ wbuf = P.wbuf;

  // ---- tile prologue: gather rows, SH tile, radial MLP hidden layers --------------------------------
  for (int idx = tid; idx < P.n_sources * TE; idx += NT) {
    const int s = idx / TE, z = idx - s * TE;
    const int64_t e = e0 + (z < ne ? z : 0);  // tail lanes alias edge 0 of the tile; never written back
    sRow[idx] = (int)(a.src_rows[s] ? a.src_rows[s][e] : e);
  }
  for (int idx = tid; idx < TE * S; idx += NT) sY[idx] = (idx < ne * S) ? a.sh[e0 * S + idx] : 0.f;
  {
    const int Rb = P.rbf_dim;
    for (int idx = tid; idx < TE * Rb; idx += NT) {
      const int z = idx / Rb, c = idx - z * Rb;
      sA[c * TE + z] = (z < ne) ? a.rbf[e0 * Rb + idx] : 0.f;  // K-major [rbf][z]
    }
  }
  __syncthreads();
  for (int b = 0; b < P.n_branches; ++b) {
    float acc[NTL][4][4];
    zero_tiles<NTL>(acc);
    gemm_tiles<NT, NTL>(acc, sA, TE, P.rbf_dim, TE / 4, P.h1 / 4, wbuf + P.fc1_off[b], P.h1);
#pragma unroll
    for (int n = 0; n < NTL; ++n) {
      const int ti = tid + n * NT;
      if (ti < (TE / 4) * (P.h1 / 4)) {
        const int c4 = ti / (TE / 4), r4 = ti - c4 * (TE / 4);
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int i = 0; i < 4; ++i) sB[(c4 * 4 + j) * TE + r4 * 4 + i] = hgb::silu_f(acc[n][i][j]) * P.act_const;
      }
    }
    __syncthreads();
    zero_tiles<NTL>(acc);
    gemm_tiles<NT, NTL>(acc, sB, TE, P.h1, TE / 4, P.h2 / 4, wbuf + P.fc2_off[b], P.h2);
#pragma unroll
    for (int n = 0; n < NTL; ++n) {
      const int ti = tid + n * NT;
      if (ti < (TE / 4) * (P.h2 / 4)) {
        const int c4 = ti / (TE / 4), r4 = ti - c4 * (TE / 4);
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int i = 0; i < 4; ++i)
            sH2[(b * 64 + c4 * 4 + j) * TE + r4 * 4 + i] = hgb::silu_f(acc[n][i][j]) * P.act_const;
      }
    }
    __syncthreads();
  }

  // ---- main loop over output slots -----------------------------------------------------------------
  for (int t = 0; t < P.n_types; ++t) {
    const hgb_type_t ty = P.types[t];
    if (ty.path_begin == ty.path_end && a.out_index != nullptr) continue;  // nothing to scatter
    const int d3 = 2 * ty.l + 1;
    const int mpad = ty.mpad, M4 = mpad >> 2;
    const int tesub = sub_tile_edges(TE, RMAX, d3);
    for (int z0 = 0; z0 < TE; z0 += tesub) {
      const int nz = min(tesub, TE - z0);
      if (z0 >= ne) break;
      const int R4 = (nz * d3) >> 2;  // nz % 4 == 0
      float acc[NTL][4][4];
      zero_tiles<NTL>(acc);

      for (int p = ty.path_begin; p < ty.path_end; ++p) {
        const hgb_path_t pa = P.paths[p];
        const int d1 = 2 * pa.l1 + 1;
        const int K = pa.nsrc * pa.mul_in;
        const int blk = pa.mul_in * d1;   // floats per source
        const int ldx = (K * d1) | 1;
        float* xblk = sB;
        // stage the input irrep block (coalesced along the row)
        for (int idx = tid; idx < nz * pa.nsrc * blk; idx += NT) {
          const int z = idx / (pa.nsrc * blk);
          const int rem = idx - z * (pa.nsrc * blk);
          const int s = rem / blk, c = rem - s * blk;
          const int sidx = pa.src0 + s;
          const float* row = a.src[sidx] + (size_t)sRow[sidx * TE + z0 + z] * P.src_dim[sidx] + pa.in_off;
          xblk[(size_t)z * ldx + rem] = __ldg(row + c);
        }
        // T_z (one thread per (z,k) column)
        for (int idx = tid; idx < nz * d3; idx += NT) {
          const int z = idx / d3, k = idx - z * d3;
          float* tz = sT + (size_t)z * d1 * d3;
          for (int i = 0; i < d1; ++i) tz[i * d3 + k] = 0.f;
          if (pa.kind == 0) {
            const float* yz = sY + (z0 + z) * S + pa.sh_off;
            const int n0 = P.cg_kstart[pa.cg_kstart + k], n1 = P.cg_kstart[pa.cg_kstart + k + 1];
            for (int n = n0; n < n1; ++n) {
              const int ij = P.cg_ij[pa.cg_off + n];
              tz[(ij & 255) * d3 + k] += P.cg_val[pa.cg_off + n] * yz[ij >> 8];
            }
          } else {
            tz[k * d3 + k] = 1.f;
          }
        }
        float bacc[NTL][4][4];
        if (pa.kind == 0) {
          // radial gate for this path's channels: g[z][w]
          zero_tiles<NTL>(bacc);
          gemm_tiles<NT, NTL>(bacc, sH2 + pa.branch * 64 * TE + z0, TE, P.h2, nz >> 2, M4, wbuf + pa.w3_off, mpad);
#pragma unroll
          for (int n = 0; n < NTL; ++n) {
            const int ti = tid + n * NT;
            if (ti < (nz >> 2) * M4) {
              const int c4 = ti / (nz >> 2), r4 = ti - c4 * (nz >> 2);
#pragma unroll
              for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) sG[(r4 * 4 + i) * mpad + c4 * 4 + j] = bacc[n][i][j];
            }
          }
          zero_tiles<NTL>(bacc);
        }
        __syncthreads();
        for (int u0 = 0; u0 < K; u0 += KC) {
          const int kc = min(KC, K - u0);
          agen_dispatch<NT>(d3, sA, RMAX, xblk, ldx, sT, d1, nz, u0, kc);
          __syncthreads();
          if (pa.kind == 0)
            gemm_tiles<NT, NTL>(bacc, sA, RMAX, kc, R4, M4, wbuf + pa.w_off + (size_t)u0 * mpad, mpad);
          else
            gemm_tiles<NT, NTL>(acc, sA, RMAX, kc, R4, M4, wbuf + pa.lf_off + (size_t)u0 * mpad, mpad);
          __syncthreads();
        }
        if (pa.kind == 0) {
          // gate and hand B to GEMM2 (K-major in sB; xblk is dead)
#pragma unroll
          for (int n = 0; n < NTL; ++n) {
            const int ti = tid + n * NT;
            if (ti < R4 * M4) {
              const int c4 = ti / R4, r4 = ti - c4 * R4;
              int zr[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) zr[i] = (r4 * 4 + i) / d3;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float4 v;
                v.x = bacc[n][0][j] * sG[zr[0] * mpad + c4 * 4 + j];
                v.y = bacc[n][1][j] * sG[zr[1] * mpad + c4 * 4 + j];
                v.z = bacc[n][2][j] * sG[zr[2] * mpad + c4 * 4 + j];
                v.w = bacc[n][3][j] * sG[zr[3] * mpad + c4 * 4 + j];
                *reinterpret_cast<float4*>(sB + (size_t)(c4 * 4 + j) * RMAX + r4 * 4) = v;
              }
            }
          }
          __syncthreads();
          gemm_tiles<NT, NTL>(acc, sB, RMAX, mpad, R4, M4, wbuf + pa.lf_off, mpad);
          __syncthreads();
        }
      }

      // ---- write the slot: out[e][out_off + w*d3 + k]
#pragma unroll
      for (int n = 0; n < NTL; ++n) {
        const int ti = tid + n * NT;
        if (ti < R4 * M4) {
          const int c4 = ti / R4, r4 = ti - c4 * R4;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = r4 * 4 + i;
            const int z = r / d3, k = r - z * d3;
            const int zt = z0 + z;
            if (zt < ne) {
              const int64_t e = e0 + zt;
              const int64_t orow = a.out_index ? a.out_index[e] : e;
              float* o = a.out + orow * P.out_dim + ty.out_off + k;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int w = c4 * 4 + j;
                if (w < ty.mul) {
                  if (a.out_index) atomicAdd(o + w * d3, acc[n][i][j]);
                  else o[w * d3] = acc[n][i][j];
                }
              }
            }
          }
        }
      }
    }
  }
}

template <int TE, int RMAX, int NT>
int launch(const MsgArgs& a, cudaStream_t st) {
  constexpr size_t smem = (size_t)Smem<TE, RMAX>::TOTAL * sizeof(float);
  static_assert(smem <= 227 * 1024, "shared memory budget");
  auto kern = msgpack_kernel<TE, RMAX, NT>;
  HGB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned grid = (unsigned)((a.n_edges + TE - 1) / TE);
  kern<<<grid, NT, smem, st>>>(a);
  HGB_LAUNCH_OK("msgpack_kernel");
  return 0;
}

}  // namespace

extern "C" int hgb_msgpack_forward(const hgb_msgpack_plan* plan, const float* const* src,
                                   const int64_t* const* src_rows, const float* sh, const float* rbf,
                                   int64_t n_edges, float* out, const int64_t* out_index, void* stream) {
  HGB_DEVICE_GUARD(out);
  HGB_CHECK_ARG(plan && src && sh && rbf && out, "hgb_msgpack_forward: NULL argument");
  HGB_CHECK_ARG(plan->n_sources >= 1 && plan->n_sources <= 4, "hgb_msgpack_forward: n_sources=%d", plan->n_sources);
  HGB_CHECK_ARG(plan->n_branches >= 1 && plan->n_branches <= 2, "hgb_msgpack_forward: n_branches=%d", plan->n_branches);
  HGB_CHECK_ARG(plan->h1 % 4 == 0 && plan->h2 % 4 == 0 && plan->h1 <= 64 && plan->h2 <= 64 && plan->rbf_dim <= 128,
                "hgb_msgpack_forward: radial MLP [%d,%d,%d] unsupported (hidden sizes must be multiples of 4, <= 64)",
                plan->rbf_dim, plan->h1, plan->h2);
  HGB_CHECK_ARG(plan->sh_dim <= 84, "hgb_msgpack_forward: sh_dim=%d too large", plan->sh_dim);
  HGB_CHECK_ARG(n_edges >= 0 && n_edges < (1ll << 31), "hgb_msgpack_forward: bad edge count");
  HGB_CHECK_ARG(plan->types_host && plan->paths_host, "hgb_msgpack_forward: host copies of the type/path tables are required");
  // tile configuration: 32 edges x 128 rows x 256 threads (2 CTAs/SM) or, with HGB_SIMT_TILE=64, 64 x 256 x 512
  static const bool big_tile = [] { const char* e = getenv("HGB_SIMT_TILE"); return e && atoi(e) == 64; }();
  const int TE = big_tile ? 64 : 32, RMAX = big_tile ? 256 : 128;
  for (int t = 0; t < plan->n_types; ++t) {
    const hgb_type_t& ty = plan->types_host[t];
    const int d3 = 2 * ty.l + 1;
    HGB_CHECK_ARG(ty.l >= 0 && ty.l <= HGB_MAX_L, "hgb_msgpack_forward: output l=%d unsupported (max %d)", ty.l, HGB_MAX_L);
    HGB_CHECK_ARG(ty.mpad % 4 == 0 && ty.mpad >= ty.mul && ty.mpad <= MPAD_MAX,
                  "hgb_msgpack_forward: output multiplicity %d (padded %d) unsupported (max %d)", ty.mul, ty.mpad, MPAD_MAX);
    HGB_CHECK_ARG(ty.out_off >= 0 && ty.out_off + ty.mul * d3 <= plan->out_dim, "hgb_msgpack_forward: slot %d outside the output row", t);
    HGB_CHECK_ARG(ty.path_begin >= 0 && ty.path_begin <= ty.path_end && ty.path_end <= plan->n_paths, "hgb_msgpack_forward: bad path range of slot %d", t);
    const int nz = sub_tile_edges(TE, RMAX, d3);
    HGB_CHECK_ARG(nz * d3 <= RMAX, "hgb_msgpack_forward: l=%d rows do not fit the %d-row tile", ty.l, RMAX);
    for (int p = ty.path_begin; p < ty.path_end; ++p) {
      const hgb_path_t& pa = plan->paths_host[p];
      const int d1 = 2 * pa.l1 + 1;
      HGB_CHECK_ARG(pa.l3 == ty.l && pa.l1 >= 0 && pa.l1 <= HGB_MAX_L && pa.l2 >= 0 && pa.l2 <= HGB_MAX_L, "hgb_msgpack_forward: path %d has unsupported l", p);
      HGB_CHECK_ARG(pa.kind == 0 || (pa.kind == 1 && pa.l1 == pa.l3), "hgb_msgpack_forward: path %d has bad kind", p);
      HGB_CHECK_ARG(pa.nsrc >= 1 && pa.nsrc <= 2 && pa.src0 >= 0 && pa.src0 + pa.nsrc <= plan->n_sources, "hgb_msgpack_forward: path %d has bad sources", p);
      HGB_CHECK_ARG(pa.branch >= 0 && pa.branch < plan->n_branches, "hgb_msgpack_forward: path %d has bad branch", p);
      for (int s = 0; s < pa.nsrc; ++s)
        HGB_CHECK_ARG(pa.in_off >= 0 && pa.in_off + pa.mul_in * d1 <= plan->src_dim[pa.src0 + s], "hgb_msgpack_forward: path %d reads outside its source row", p);
      const long need = (long)nz * ((pa.nsrc * pa.mul_in * d1) | 1);
      HGB_CHECK_ARG(need <= (long)MPAD_MAX * RMAX, "hgb_msgpack_forward: input block of path %d (mul %d x l %d) exceeds the %d-float staging buffer",
                    p, pa.nsrc * pa.mul_in, pa.l1, MPAD_MAX * RMAX);
    }
  }
  if (n_edges == 0) return 0;
  MsgArgs a;
  memset(&a, 0, sizeof(a));
  a.plan = *plan;
  for (int s = 0; s < plan->n_sources; ++s) {
    HGB_CHECK_ARG(src[s] != nullptr, "hgb_msgpack_forward: source %d is NULL", s);
    a.src[s] = src[s];
    a.src_rows[s] = src_rows ? src_rows[s] : nullptr;
  }
  a.sh = sh; a.rbf = rbf; a.n_edges = n_edges; a.out = out; a.out_index = out_index;
  return big_tile ? launch<64, 256, 512>(a, (cudaStream_t)stream) : launch<32, 128, 256>(a, (cudaStream_t)stream);
}
