// Fused MessagePackBlock forward on tcgen05, variant "tcg": the radial gate g[e][c] = h2[e] . W3[:, c] is produced
// once per call by `radial_gate_kernel` (one [E, n_channels] fp32 tensor per branch, written once and read once),
// so the message kernel needs neither the G' MMA nor the hidden activations in TMEM.  That shrinks the TMEM
// footprint to 3 regions  C | B -> (B*g)hi | (B*g)lo  (96 or 192 columns) and the shared-memory footprint below
// 113 KB, i.e. TWO CTAs per SM for every slot class -- the measured bottleneck of msgpack_tc_kernel was exposed
// latency with one (or two) resident CTAs, not tensor or FMA throughput (profiles/r01g_*).
//
// Everything else (tables, packing, A-operand generation, 3xTF32 split, GEMM1 / GEMM2 structure) is identical to
// msgpack_tc.cu.  Price: 2 x 4 B x n_channels (28.7 KB) of extra HBM write + read per message.
#include <stdlib.h>
#include <cuda_fp16.h>
#include "hgb_common.cuh"
#include "tc_common.cuh"
#include "msgpack_tc_helpers.cuh"

namespace {
using namespace tcmsg;

constexpr int KIMG = 32;  // K-chunking of the packed W images (host side)
constexpr int NMAX = 64;

struct TgArgs {
  hgb_msgpack_plan plan;
  const float* src[4];
  const int64_t* src_rows[4];
  const float* sh;
  const float* g;       // [n_branches][E][gstride]
  int64_t n_edges;
  int gstride;
  float* out;
  const int64_t* out_index;
  int tile_start[33];
  int type_order[32];
  int n_sched;
};

template <int RW>
struct TmG {
  static constexpr uint32_t C = 0, B = RW, BGH = RW, BGL = 2 * RW;
  static constexpr int COLS = (RW == 32) ? 128 : 256;
};

// KC = channels per GEMM1 step (A image of KC/4 slabs): 32 for the wide class, 16 for the small class, which then
// fits FOUR CTAs of 128 threads per SM (50 KB shared memory, 128 TMEM columns, 128 registers each).
template <int RW, int KC>
struct SmemG {
  static constexpr int NSLAB = KC / 4;
  static constexpr int XBLK = (RW == 64) ? 6144 : 2560;
  static constexpr int TBLK = (RW == 64) ? 1536 : 3072;   // RW = 64 serves d3 == 1 (compact T: 128 * d1 floats, l1 <= 5)
  static constexpr int A = 0;
  static constexpr int W = A + 2 * NSLAB * ROWS * 4;
  static constexpr int L = W + 2 * RW * KC;
  static constexpr int X = L + 2 * RW * RW;
  static constexpr int T = X + XBLK;
  static constexpr int ROW = T + TBLK;
  static constexpr int TOTAL = ROW + 4 * ROWS;
};

// d3 == 1 specialisation of the A-operand generator with a compact T (one float per (z, i))
template <int NT>
__device__ __forceinline__ void agen_d3_1(float* __restrict__ sA, int lo_off, const float* __restrict__ sX, int ldx, int cc,
                                          const float* __restrict__ sT, int d1, int nz, uint32_t mz, int slab0,
                                          int nquad) {
  for (int item = threadIdx.x; item < nz * nquad; item += NT) {
    const int q = (int)fdiv((uint32_t)item, mz), z = item - q * nz;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* xz = sX + (size_t)z * ldx + q * 4;
    const float* tz = sT + (size_t)z * d1;
    for (int i = 0; i < d1; ++i) {
      const float4 x = *reinterpret_cast<const float4*>(xz + i * cc);
      const float t = tz[i];
      acc.x = fmaf(x.x, t, acc.x); acc.y = fmaf(x.y, t, acc.y); acc.z = fmaf(x.z, t, acc.z); acc.w = fmaf(x.w, t, acc.w);
    }
    float4 h, l;
    tc::split_tf32(acc.x, h.x, l.x); tc::split_tf32(acc.y, h.y, l.y);
    tc::split_tf32(acc.z, h.z, l.z); tc::split_tf32(acc.w, h.w, l.w);
    float* hi = sA + (size_t)(slab0 + q) * (ROWS * 4) + (size_t)z * 4;
    *reinterpret_cast<float4*>(hi) = h;
    *reinterpret_cast<float4*>(hi + lo_off) = l;
  }
}

template <int RW, int NT, int KC>
__global__ void __launch_bounds__(NT, (NT == 128 ? 4 : 2)) msgpack_tcg_kernel(const __grid_constant__ TgArgs a) {
  using SM = SmemG<RW, KC>;
  constexpr int LO = SM::NSLAB * ROWS * 4;  // float offset of the lo image of A
  constexpr int WPQ = NT / 128;
  constexpr int NGRP = (RW / 8 + WPQ - 1) / WPQ;  // 8-column groups per warp in the epilogues
  constexpr int XBLK = SM::XBLK;
  constexpr uint32_t TC_C = TmG<RW>::C, TC_B = TmG<RW>::B, TC_BGH = TmG<RW>::BGH, TC_BGL = TmG<RW>::BGL;
  extern __shared__ __align__(128) float smem[];
  float* sA = smem + SM::A;
  float* sW = smem + SM::W;
  float* sL = smem + SM::L;
  float* sX = smem + SM::X;
  float* sT = smem + SM::T;
  int* sRow = reinterpret_cast<int*>(smem + SM::ROW);
  __shared__ uint64_t mbar[2];
  __shared__ uint32_t tmem_slot;

  const hgb_msgpack_plan& P = a.plan;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* __restrict__ wbuf = P.wbuf;

  int sq = 0;
  while (sq + 1 < a.n_sched && (int)blockIdx.x >= a.tile_start[sq + 1]) ++sq;
  const int t = a.type_order[sq];
  const hgb_type_t ty = P.types[t];
  const int d3 = 2 * ty.l + 1;
  const int mp = ty.mpad;
  const int nz_full = ROWS / d3;
  const int64_t e0 = (int64_t)(blockIdx.x - a.tile_start[sq]) * nz_full;
  const int nz = (int)min((int64_t)nz_full, a.n_edges - e0);
  const int R = nz * d3;
  const int S = P.sh_dim;

  if (tid == 0) {
    tc::mbar_init(&mbar[0], 1);
    tc::mbar_init(&mbar[1], 1);
    tc::mbar_fence_init();
  }
  if (warp == 0) tc::tmem_alloc<TmG<RW>::COLS>(&tmem_slot);
  for (int idx = tid; idx < P.n_sources * nz; idx += NT) {
    const int s = idx / nz, z = idx - s * nz;
    const int64_t e = e0 + z;
    sRow[s * ROWS + z] = (int)(a.src_rows[s] ? a.src_rows[s][e] : e);
  }
  for (int idx = tid; idx < 2 * LO; idx += NT) sA[idx] = 0.f;
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const int my_row = (warp & 3) * 32 + lane;
  const bool live = my_row < R;
  const int my_z = live ? my_row / d3 : 0;
  const int wq = warp >> 2;
  const uint32_t idesc = tc::idesc_tf32_m128(mp);
  const uint32_t sbo = 128, lbo_a = ROWS * 16, lbo_n = (uint32_t)mp * 16;

  uint32_t ph0 = 0, ph1 = 0;
  bool pend0 = false, pend1 = false, c_started = false;
  const uint32_t mz = fdiv_magic((uint32_t)nz), md3 = fdiv_magic_odd(ty.l);
  const int xcap = XBLK / nz - 4;

  for (int p = ty.path_begin; p < ty.path_end; ++p) {
    const hgb_path_t pa = P.paths[p];
    const int d1 = 2 * pa.l1 + 1;
    const int K = pa.nsrc * pa.mul_in;
    const int Kpad = (K + 7) & ~7;
    if (pend1) { tc::cta_wait(&mbar[1], ph1); ph1 ^= 1; pend1 = false; tc::fence_after_sync(); }
    if (pend0) { tc::cta_wait(&mbar[0], ph0); ph0 ^= 1; pend0 = false; tc::fence_after_sync(); }

    if (pa.kind == 0) copy_f4<NT>(sL, wbuf + pa.lf_off, 2 * mp * mp);
    // T_z: padded rows (stride ldt, ldt/4 odd) for d3 > 1, compact [z][i] for d3 == 1; Y is read from global (L1/L2)
    const int d3p = (d3 + 3) & ~3;
    int ldt = d1 * d3p;
    if (((ldt >> 2) & 1) == 0) ldt += 4;
    if (d3 == 1) {
      for (int idx = tid; idx < nz * d1; idx += NT) sT[idx] = 0.f;
      __syncthreads();
      if (pa.kind == 0) {
        const int n1 = P.cg_kstart[pa.cg_kstart + 1];
        for (int idx = tid; idx < nz * n1; idx += NT) {
          const int z = idx / n1, n = idx - z * n1;
          const int ij = P.cg_ij[pa.cg_off + n];
          atomicAdd(&sT[z * d1 + (ij & 255)], P.cg_val[pa.cg_off + n] * __ldg(a.sh + (e0 + z) * S + pa.sh_off + (ij >> 8)));
        }
      } else {
        for (int z = tid; z < nz; z += NT) sT[z] = 1.f;  // d1 == d3 == 1
      }
    } else {
      for (int idx = tid; idx < nz * d3; idx += NT) {
        const int z = (int)fdiv((uint32_t)idx, md3), k = idx - z * d3;
        float* tz = sT + (size_t)z * ldt;
        for (int i = 0; i < d1; ++i) tz[i * d3p + k] = 0.f;
        if (pa.kind == 0) {
          const float* yz = a.sh + (e0 + z) * S + pa.sh_off;
          const int n0 = P.cg_kstart[pa.cg_kstart + k], n1 = P.cg_kstart[pa.cg_kstart + k + 1];
          for (int n = n0; n < n1; ++n) {
            const int ij = P.cg_ij[pa.cg_off + n];
            tz[(ij & 255) * d3p + k] += P.cg_val[pa.cg_off + n] * __ldg(yz + (ij >> 8));
          }
        } else {
          tz[k * d3p + k] = 1.f;
        }
      }
    }

    // ---- GEMM1 in K chunks
    int chunk = 0;
    for (int u0 = 0; u0 < Kpad; u0 += KC, ++chunk) {
      const int kc = min(KC, Kpad - u0);
      if (pend0) { tc::cta_wait(&mbar[0], ph0); ph0 ^= 1; pend0 = false; tc::fence_after_sync(); }
      {
        // the packed image is chunked by KIMG = 32 channels: [kimg/4 slabs][mp][4] hi | lo; take our KC-slice of both
        const int img = u0 / KIMG, kimg = min(KIMG, Kpad - img * KIMG), s0 = (u0 - img * KIMG) >> 2;
        const float* base = wbuf + (pa.kind == 0 ? pa.w_off : pa.lf_off) + (size_t)2 * mp * KIMG * img;
        copy_f4<NT>(sW, base + (size_t)s0 * mp * 4, mp * kc);
        copy_f4<NT>(sW + mp * kc, base + (size_t)mp * kimg + (size_t)s0 * mp * 4, mp * kc);
      }
      const uint32_t md1 = fdiv_magic_odd(pa.l1);
      int cch = (int)fdiv((uint32_t)xcap, md1) & ~3;
      if (cch > kc) cch = kc;
      for (int ua = u0; ua < u0 + kc; ua += cch) {
        const int cc = min(cch, u0 + kc - ua);
        int ldx = cc * d1;
        if (((ldx >> 2) & 1) == 0) ldx += 4;
        for (int z = warp; z < nz; z += NT / 32) {
          float* xrow = sX + (size_t)z * ldx;
          for (int sg = 0; sg < pa.nsrc; ++sg) {
            const int ulo = max(ua, sg * pa.mul_in), uhi = min(min(ua + cc, (sg + 1) * pa.mul_in), K);
            if (ulo >= uhi) continue;
            const int sidx = pa.src0 + sg;
            const float* gsrc = a.src[sidx] + (size_t)sRow[sidx * ROWS + z] * P.src_dim[sidx] + pa.in_off + (ulo - sg * pa.mul_in) * d1;
            const int n = (uhi - ulo) * d1, cbase = ulo - ua;
            for (int j = lane; j < n; j += 32) {
              const int ul = (int)fdiv((uint32_t)j, md1), i = j - ul * d1;
              cp_async4(xrow + i * cc + cbase + ul, gsrc + j);
            }
          }
          const int upad = max(K, ua) - ua;
          if (upad < cc)
            for (int j = lane; j < (cc - upad) * d1; j += 32) {
              const int ul = (int)fdiv((uint32_t)j, md1), i = j - ul * d1;
              xrow[i * cc + upad + ul] = 0.f;
            }
        }
        cp_async_wait_all();
        __syncthreads();
        if (d3 == 1) agen_d3_1<NT>(sA, LO, sX, ldx, cc, sT, d1, nz, mz, (ua - u0) >> 2, cc >> 2);
        else agen_tc_dispatch<NT>(d3, sA, LO, sX, ldx, cc, sT, ldt, d1, nz, mz, (ua - u0) >> 2, cc >> 2);
        if (ua + cch < u0 + kc) __syncthreads();
      }
      tc::fence_proxy_async();
      tc::fence_before_sync();
      __syncthreads();
      if (tid == 0) {
        tc::fence_after_sync();
        const uint32_t dhi = tc::smem_desc_hi(sbo);
        const uint32_t ah = tc::smem_desc_lo(tc::smem_u32(sA), lbo_a), al = ah + ((LO * 4) >> 4);
        const uint32_t wh = tc::smem_desc_lo(tc::smem_u32(sW), lbo_n), wl = wh + (((uint32_t)mp * kc * 4) >> 4);
        const uint32_t astep = (2 * lbo_a) >> 4, bstep = (2 * lbo_n) >> 4;
        const uint32_t dcol = tmem + (pa.kind == 0 ? TC_B : TC_C);
        const uint32_t acc0 = (pa.kind == 0) ? (uint32_t)(chunk > 0) : (uint32_t)(c_started || chunk > 0);
        for (int k8 = 0; k8 < (kc >> 3); ++k8) {
          const uint64_t dah = tc::desc64(ah + k8 * astep, dhi), dal = tc::desc64(al + k8 * astep, dhi);
          const uint64_t dbh = tc::desc64(wh + k8 * bstep, dhi), dbl = tc::desc64(wl + k8 * bstep, dhi);
          tc::mma_tf32(dcol, dal, dbh, idesc, acc0 | (uint32_t)(k8 > 0));
          tc::mma_tf32(dcol, dah, dbl, idesc, 1);
          tc::mma_tf32(dcol, dah, dbh, idesc, 1);
        }
        tc::mma_commit(&mbar[0]);
      }
      pend0 = true;
    }
    if (pa.kind != 0) { c_started = true; continue; }

    // ---- gate: the radial gate of this path's channels is fetched while GEMM1 drains
    float gv[NGRP][8];
    {
      const float* gp = a.g + ((size_t)pa.branch * a.n_edges + (size_t)(e0 + my_z)) * a.gstride + pa.pad0;
#pragma unroll
      for (int q = 0; q < NGRP; ++q) {
        const int c0 = (wq + q * WPQ) * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) gv[q][j] = (live && c0 + j < ty.mul) ? __ldg(gp + c0 + j) : 0.f;
      }
    }
    tc::cta_wait(&mbar[0], ph0); ph0 ^= 1; pend0 = false;
    tc::fence_after_sync();
#pragma unroll
    for (int q = 0; q < NGRP; ++q) {
      const int c0 = (wq + q * WPQ) * 8;
      if (c0 < mp) {  // warp-uniform
        uint32_t rb[8], hi[8], lo[8];
        tc::tmem_ld8(tmem + lane_base + TC_B + c0, rb);
        tc::tmem_ld_wait8(rb);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float h, l;
          tc::split_tf32(__uint_as_float(rb[j]) * gv[q][j], h, l);
          hi[j] = __float_as_uint(h); lo[j] = __float_as_uint(l);
        }
        tc::tmem_st8(tmem + lane_base + TC_BGH + c0, hi);
        tc::tmem_st8(tmem + lane_base + TC_BGL + c0, lo);
      }
    }
    tc::tmem_st_wait();
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
      const uint32_t dhi = tc::smem_desc_hi(sbo);
      const uint32_t lh = tc::smem_desc_lo(tc::smem_u32(sL), lbo_n), ll = lh + (((uint32_t)mp * mp * 4) >> 4);
      const uint32_t kstep = (2 * lbo_n) >> 4;
      for (int k8 = 0; k8 < (mp >> 3); ++k8) {
        const uint64_t bh = tc::desc64(lh + k8 * kstep, dhi), bl = tc::desc64(ll + k8 * kstep, dhi);
        tc::mma_tf32_ts(tmem + TC_C, tmem + TC_BGL + k8 * 8, bh, idesc, (uint32_t)(c_started || k8 > 0));
        tc::mma_tf32_ts(tmem + TC_C, tmem + TC_BGH + k8 * 8, bl, idesc, 1);
        tc::mma_tf32_ts(tmem + TC_C, tmem + TC_BGH + k8 * 8, bh, idesc, 1);
      }
      tc::mma_commit(&mbar[1]);
    }
    pend1 = true;
    c_started = true;
  }
  if (pend1) { tc::cta_wait(&mbar[1], ph1); ph1 ^= 1; }
  if (pend0) { tc::cta_wait(&mbar[0], ph0); ph0 ^= 1; }
  tc::fence_after_sync();

  {
    const int z = my_z, k = live ? my_row - z * d3 : 0;
    const int64_t e = e0 + z;
    const int64_t orow = (live && a.out_index) ? a.out_index[e] : e;
    float* o = a.out + orow * P.out_dim + ty.out_off + k;
    for (int c0 = wq * 8; c0 < mp; c0 += WPQ * 8) {
      if (c0 >= ty.mul) break;
      uint32_t rc[8];
      if (c_started) {
        tc::tmem_ld8(tmem + lane_base + TC_C + c0, rc);
        tc::tmem_ld_wait8(rc);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) rc[j] = 0u;
      }
      if (live) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int w = c0 + j;
          if (w < ty.mul) {
            if (a.out_index) atomicAdd(o + w * d3, __uint_as_float(rc[j]));
            else o[w * d3] = __uint_as_float(rc[j]);
          }
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<TmG<RW>::COLS>(tmem);
}

// ---------------------------------------------------------------------------------------------------------
// radial MLP, all three layers: g[b][e][c] = (act(act(rbf W1) W2) W3)[c], c < nch_b.  One CTA = 64 edges; the last
// layer is a register-tiled SGEMM (4 rows x 8 columns per thread, K = h2) streamed over 128-column tiles of W3.
struct GateArgs {
  const float* rbf;
  const float* w1[2];
  const float* w2[2];
  const float* w3[2];   // [h2][nch] row-major, pre-scaled
  int nch[2];
  float* g;             // [n_branches][E][gstride]
  int64_t n_edges;
  int n_branches, rbf_dim, h1, h2dim, gstride;
  float act_const;
  int tmajor;           // 1: g is [n_branches][tile of 128 edges][gstride][128], 2: [tile][n_branches][gstride][128] (see gtc::Args)
};
constexpr int GE = 64;    // edges per CTA
constexpr int GN = 128;   // W3 column tile

constexpr int GATE_SMEM_FLOATS = GE * 65 + 64 * (GE + 4) + 64 * GN;

__global__ void __launch_bounds__(256, 3) radial_gate_kernel(const __grid_constant__ GateArgs a) {
  extern __shared__ __align__(16) float gsm[];
  float (*sH)[GE + 4] = reinterpret_cast<float (*)[GE + 4]>(gsm);                       // h2, K-major: sH[h][z]
  float (*sB)[GN] = reinterpret_cast<float (*)[GN]>(gsm + 64 * (GE + 4));              // W3 tile
  float (*sIn)[65] = reinterpret_cast<float (*)[65]>(gsm + 64 * (GE + 4) + 64 * GN);   // rbf tile, later h1
  const int tid = threadIdx.x;
  const int64_t e0 = (int64_t)blockIdx.x * GE;
  const int ne = (int)min((int64_t)GE, a.n_edges - e0);
  const int tx = tid & 15, ty = tid >> 4;  // 16 column groups of 8, 16 row groups of 4
  for (int b = 0; b < a.n_branches; ++b) {
    __syncthreads();
    for (int idx = tid; idx < GE * a.rbf_dim; idx += 256) {
      const int z = idx / a.rbf_dim, c = idx - z * a.rbf_dim;
      sIn[z][c] = (z < ne) ? a.rbf[(e0 + z) * a.rbf_dim + c] : 0.f;
    }
    __syncthreads();
    // layer 1 -> registers -> sIn (as h1) ; rbf_dim, h1, h2 <= 64
    float h1v[16];
    {
      int cnt = 0;
      for (int idx = tid; idx < GE * a.h1; idx += 256, ++cnt) {
        const int z = idx / a.h1, c = idx - z * a.h1;
        float acc = 0.f;
        for (int r = 0; r < a.rbf_dim; ++r) acc = fmaf(sIn[z][r], __ldg(a.w1[b] + r * a.h1 + c), acc);
        h1v[cnt] = hgb::silu_f(acc) * a.act_const;
      }
      __syncthreads();
      cnt = 0;
      for (int idx = tid; idx < GE * a.h1; idx += 256, ++cnt) sIn[idx / a.h1][idx % a.h1] = h1v[cnt];
    }
    __syncthreads();
    for (int idx = tid; idx < GE * a.h2dim; idx += 256) {
      const int z = idx / a.h2dim, c = idx - z * a.h2dim;
      float acc = 0.f;
      for (int r = 0; r < a.h1; ++r) acc = fmaf(sIn[z][r], __ldg(a.w2[b] + r * a.h2dim + c), acc);
      sH[c][z] = hgb::silu_f(acc) * a.act_const;
    }
    __syncthreads();
    // layer 3: g tile [GE x GN] per column tile
    const int nch = a.nch[b];
    float* gb = a.g + ((size_t)b * a.n_edges + e0) * a.gstride;
    for (int n0 = 0; n0 < nch; n0 += GN) {
      for (int idx = tid; idx < 64 * (GN / 4); idx += 256) {
        const int h = idx / (GN / 4), c4 = idx - h * (GN / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (h < a.h2dim) {
          const int c = n0 + c4 * 4;
          const float* wp = a.w3[b] + (size_t)h * nch + c;
          if (c + 3 < nch && ((((size_t)h * nch + c) & 3) == 0)) v = __ldg(reinterpret_cast<const float4*>(wp));
          else {
            if (c < nch) v.x = __ldg(wp);
            if (c + 1 < nch) v.y = __ldg(wp + 1);
            if (c + 2 < nch) v.z = __ldg(wp + 2);
            if (c + 3 < nch) v.w = __ldg(wp + 3);
          }
        }
        *reinterpret_cast<float4*>(&sB[h][c4 * 4]) = v;
      }
      __syncthreads();
      float acc[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
      for (int h = 0; h < a.h2dim; ++h) {
        const float4 av = *reinterpret_cast<const float4*>(&sH[h][ty * 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&sB[h][tx * 4]);        // columns tx*4.. and 64+tx*4..: the 16 lanes of
        const float4 b1 = *reinterpret_cast<const float4*>(&sB[h][64 + tx * 4]);   // a half warp read 256 contiguous bytes (no bank conflict)
        const float ar[4] = {av.x, av.y, av.z, av.w};
        const float br[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
      }
      if (a.tmajor) {
        // rows ty*4 .. ty*4+3 of one column are 16 contiguous bytes of the tile-major layout (padding rows included)
        const int64_t n_tiles = (a.n_edges + 127) / 128;
        const size_t blk = (a.tmajor == 2) ? (size_t)(e0 >> 7) * a.n_branches + b : (size_t)b * n_tiles + (size_t)(e0 >> 7);   // 2: [tile][branch]
        float* gt = a.g + blk * (size_t)a.gstride * 128 + (e0 & 127) + ty * 4;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = n0 + ((j < 4) ? tx * 4 + j : 64 + tx * 4 + (j - 4));
          if (c < nch) *reinterpret_cast<float4*>(gt + (size_t)c * 128) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
        }
      } else
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int z = ty * 4 + i;
        if (z < ne) {
          float* gr = gb + (size_t)z * a.gstride + n0;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int c = (j < 4) ? tx * 4 + j : 64 + tx * 4 + (j - 4);
            if (n0 + c < nch) gr[c] = acc[i][j];
          }
        }
      }
      __syncthreads();
    }
  }
}

#include "msgpack_tcr_kernel.cuh"
#include "radial_gate_tc_kernel.cuh"
#include "msgpack_rot_kernel.cuh"
#include "msgpack_rot16_kernel.cuh"
#include "msgpack_rotf_kernel.cuh"
#include "msgpack_rot2_kernel.cuh"

// Radial gate pre-pass: the tcgen05 kernel when the host supplies the packed W3 tiles (w3img_off != NULL) and the
// MLP widths fit its tiling, the fp32-FMA kernel otherwise.
int launch_radial_gate(const hgb_msgpack_plan* plan, const float* rbf, const int32_t* w3_off, const int32_t* nch,
                       const int32_t* w3img_off, int32_t gstride, float* g_ws, int64_t n_edges, cudaStream_t st, int tmajor = 0) {
  for (int b = 0; b < plan->n_branches; ++b)
    HGB_CHECK_ARG(nch[b] > 0 && nch[b] <= gstride, "radial gate: width %d exceeds stride %d", nch[b], gstride);
  const bool tc_ok = w3img_off != nullptr && plan->h2 % 8 == 0 && plan->h2 <= gtc::KMAX && plan->h1 % 4 == 0 &&
                     plan->h1 <= 64 && plan->rbf_dim <= 64 && gstride % 4 == 0;
  if (tc_ok) {
    gtc::Args ga;
    memset(&ga, 0, sizeof(ga));
    ga.rbf = rbf; ga.g = g_ws; ga.n_edges = n_edges; ga.rbf_dim = plan->rbf_dim; ga.h1 = plan->h1; ga.h2dim = plan->h2;
    ga.gstride = gstride; ga.act_const = plan->act_const; ga.tmajor = tmajor;
    for (int b = 0; b < plan->n_branches; ++b) {
      HGB_CHECK_ARG(w3img_off[b] >= 0 && w3img_off[b] % 4 == 0, "radial gate: packed W3 tiles of branch %d are not 16-byte aligned", b);
      ga.w1[b] = plan->wbuf + plan->fc1_off[b];
      ga.w2[b] = plan->wbuf + plan->fc2_off[b];
      ga.w3img[b] = plan->wbuf + w3img_off[b];
      ga.nch[b] = nch[b];
    }
    constexpr size_t smem = (size_t)gtc::Sm::TOTAL * sizeof(float);
    static_assert(smem <= 200 * 1024, "radial_gate_tc_kernel shared memory");
    HGB_CUDA_OK(cudaFuncSetAttribute(gtc::radial_gate_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((n_edges + ROWS - 1) / ROWS), (unsigned)plan->n_branches);
    gtc::radial_gate_tc_kernel<<<grid, gtc::NT, smem, st>>>(ga);
    HGB_LAUNCH_OK("radial_gate_tc_kernel");
    return 0;
  }
  GateArgs ga;
  memset(&ga, 0, sizeof(ga));
  ga.rbf = rbf; ga.g = g_ws; ga.n_edges = n_edges; ga.n_branches = plan->n_branches; ga.rbf_dim = plan->rbf_dim;
  ga.h1 = plan->h1; ga.h2dim = plan->h2; ga.gstride = gstride; ga.act_const = plan->act_const; ga.tmajor = tmajor;
  for (int b = 0; b < plan->n_branches; ++b) {
    ga.w1[b] = plan->wbuf + plan->fc1_off[b];
    ga.w2[b] = plan->wbuf + plan->fc2_off[b];
    ga.w3[b] = plan->wbuf + w3_off[b];
    ga.nch[b] = nch[b];
  }
  constexpr size_t gate_smem = (size_t)GATE_SMEM_FLOATS * sizeof(float);
  HGB_CUDA_OK(cudaFuncSetAttribute(radial_gate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gate_smem));
  radial_gate_kernel<<<(unsigned)((n_edges + GE - 1) / GE), 256, gate_smem, st>>>(ga);
  HGB_LAUNCH_OK("radial_gate_kernel");
  return 0;
}

}  // namespace

// The radial gate alone (for tests and for callers that keep g): g_ws[b][e][c], c < nch[b], row stride gstride.
extern "C" int hgb_radial_gate(const hgb_msgpack_plan* plan, const float* rbf, const int32_t* w3_off, const int32_t* nch,
                               const int32_t* w3img_off, int32_t gstride, float* g_ws, int64_t n_edges, void* stream) {
  HGB_DEVICE_GUARD(g_ws);
  HGB_CHECK_ARG(plan && rbf && g_ws && w3_off && nch, "hgb_radial_gate: NULL argument");
  HGB_CHECK_ARG(plan->n_branches >= 1 && plan->n_branches <= 2, "hgb_radial_gate: bad branch count");
  HGB_CHECK_ARG(plan->h2 <= 64 && plan->h1 <= 64 && plan->rbf_dim <= 64,
                "hgb_radial_gate: radial MLP [%d,%d,%d] unsupported (all widths <= 64)", plan->rbf_dim, plan->h1, plan->h2);
  HGB_CHECK_ARG(n_edges >= 0 && n_edges < (1ll << 31), "hgb_radial_gate: bad edge count");
  if (n_edges == 0) return 0;
  return launch_radial_gate(plan, rbf, w3_off, nch, w3img_off, gstride, g_ws, n_edges, (cudaStream_t)stream);
}

// g_ws: device workspace of n_branches * n_edges * gstride floats (gstride >= max n_channels).  w3_off[b] / nch[b]:
// offset (floats, into plan->wbuf) of the pre-scaled last radial layer [h2][nch_b] and its width.  paths[].pad0
// holds the first gate column of the path.  Otherwise like hgb_msgpack_tc_forward.
extern "C" int hgb_msgpack_tcg_forward(const hgb_msgpack_plan* plan, const float* const* src,
                                       const int64_t* const* src_rows, const float* sh, const float* rbf,
                                       const int32_t* w3_off, const int32_t* nch, int32_t gstride, float* g_ws,
                                       int64_t n_edges, float* out, const int64_t* out_index, void* stream) {
  return hgb_msgpack_tcg_forward_v2(plan, src, src_rows, sh, rbf, w3_off, nch, nullptr, gstride, g_ws, n_edges, out, out_index, stream);
}

// w3img_off (nullable): offsets into plan->wbuf of the per-branch packed W3 tiles for the tensor-core gate pre-pass.
extern "C" int hgb_msgpack_tcg_forward_v2(const hgb_msgpack_plan* plan, const float* const* src,
                                          const int64_t* const* src_rows, const float* sh, const float* rbf,
                                          const int32_t* w3_off, const int32_t* nch, const int32_t* w3img_off,
                                          int32_t gstride, float* g_ws, int64_t n_edges, float* out,
                                          const int64_t* out_index, void* stream) {
  HGB_DEVICE_GUARD(out);
  HGB_CHECK_ARG(plan && src && sh && rbf && out && g_ws && w3_off && nch, "hgb_msgpack_tcg_forward: NULL argument");
  HGB_CHECK_ARG(plan->types_host && plan->paths_host, "hgb_msgpack_tcg_forward: host copies of the type/path tables are required");
  HGB_CHECK_ARG(plan->n_sources >= 1 && plan->n_sources <= 4 && plan->n_branches >= 1 && plan->n_branches <= 2,
                "hgb_msgpack_tcg_forward: bad source/branch count");
  HGB_CHECK_ARG(plan->h2 <= 64 && plan->h1 <= 64 && plan->rbf_dim <= 64,
                "hgb_msgpack_tcg_forward: radial MLP [%d,%d,%d] unsupported (all widths <= 64)", plan->rbf_dim, plan->h1, plan->h2);
  HGB_CHECK_ARG(plan->n_types <= 32, "hgb_msgpack_tcg_forward: too many output slots");
  HGB_CHECK_ARG(n_edges >= 0 && n_edges < (1ll << 31), "hgb_msgpack_tcg_forward: bad edge count");
  if (n_edges == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;

  for (int b = 0; b < plan->n_branches; ++b)
    HGB_CHECK_ARG(nch[b] > 0 && nch[b] <= gstride, "hgb_msgpack_tcg_forward: gate width %d exceeds stride %d", nch[b], gstride);
  // slot classes: 0 = scalar slots (d3 == 1, up to 64 channels) on msgpack_tcg_kernel<64,256,32>;
  // 1 / 2 = l >= 1 slots padded to 16 / 32 channels on the rows-in-lanes pipeline msgpack_tcr_kernel<16 / 32>
  // (HGB_TCG_ROWS=0 keeps them on msgpack_tcg_kernel<32,128,16>, the r01j kernel, for A/B measurements).
  const char* rows_env = getenv("HGB_TCG_ROWS");
  const bool use_rows = !(rows_env && rows_env[0] == '0');
  TgArgs cls[3];
  int ctas[3] = {0, 0, 0};
  double cost[32];
  int order[32], klass[32];
  for (int t = 0; t < plan->n_types; ++t) {
    const hgb_type_t& ty = plan->types_host[t];
    HGB_CHECK_ARG(ty.l >= 0 && ty.l <= HGB_MAX_L && ty.mpad % 16 == 0 && ty.mpad >= ty.mul && ty.mpad <= NMAX,
                  "hgb_msgpack_tcg_forward: slot %d (mul %d, padded %d, l %d) unsupported", t, ty.mul, ty.mpad, ty.l);
    const int d3 = 2 * ty.l + 1;
    klass[t] = (ty.mpad <= 32 && d3 >= 3) ? ((use_rows && ty.mpad == 32) ? 2 : 1) : 0;
    HGB_CHECK_ARG(klass[t] != 0 || d3 == 1, "hgb_msgpack_tcg_forward: slot %d (multiplicity %d > 32 with l = %d) unsupported", t, ty.mul, ty.l);
    double c = 0;
    for (int p = ty.path_begin; p < ty.path_end; ++p) {
      const hgb_path_t& pa = plan->paths_host[p];
      const int d1 = 2 * pa.l1 + 1;
      HGB_CHECK_ARG(pa.l3 == ty.l && pa.l1 >= 0 && pa.l1 <= HGB_MAX_L && pa.l2 >= 0 && pa.l2 <= HGB_MAX_L, "hgb_msgpack_tcg_forward: bad path %d", p);
      HGB_CHECK_ARG(pa.nsrc >= 1 && pa.nsrc <= 2 && pa.src0 >= 0 && pa.src0 + pa.nsrc <= plan->n_sources, "hgb_msgpack_tcg_forward: path %d sources", p);
      HGB_CHECK_ARG(pa.kind != 0 || (pa.pad0 >= 0 && pa.pad0 + ty.mul <= nch[pa.branch]), "hgb_msgpack_tcg_forward: gate columns of path %d out of range", p);
      const int xb = klass[t] ? SmemG<32, 16>::XBLK : SmemG<64, 32>::XBLK, tb = klass[t] ? SmemG<32, 16>::TBLK : SmemG<64, 32>::TBLK;
      const int tneed = (d3 == 1) ? (ROWS * d1) : (ROWS / d3) * (d1 * ((d3 + 3) & ~3) + 4);
      if (use_rows && klass[t] != 0) {
        HGB_CHECK_ARG(d1 <= tcr::D1MAX && (31 / d3 + 2) * ((pa.nsrc * pa.mul_in > 4) ? 8 : 4) * d1 <= tcr::XW && ROWS / d3 <= tcr::ZSTR, "hgb_msgpack_tcg_forward: staging buffers too small for path %d", p);
      } else {
        HGB_CHECK_ARG((ROWS / d3) * (d1 * 4 + 4) <= xb && tneed <= tb, "hgb_msgpack_tcg_forward: staging buffers too small for path %d", p);
      }
      c += (double)(pa.nsrc * pa.mul_in + ty.mpad) * ty.mpad * d3;
    }
    cost[t] = c;
    order[t] = t;
  }
  for (int i = 0; i < plan->n_types; ++i)
    for (int j = i + 1; j < plan->n_types; ++j)
      if (cost[order[j]] > cost[order[i]]) { int tmp = order[i]; order[i] = order[j]; order[j] = tmp; }
  for (int k = 0; k < 3; ++k) {
    TgArgs& a = cls[k];
    memset(&a, 0, sizeof(a));
    a.plan = *plan;
    int ns = 0;
    for (int q = 0; q < plan->n_types; ++q) {
      const int t = order[q];
      const hgb_type_t& ty = plan->types_host[t];
      if (klass[t] != k) continue;
      if (ty.path_begin == ty.path_end && out_index != nullptr) continue;
      const int nzf = ROWS / (2 * ty.l + 1);
      a.type_order[ns] = t;
      a.tile_start[ns] = ctas[k];
      ctas[k] += (int)((n_edges + nzf - 1) / nzf);
      ++ns;
    }
    a.tile_start[ns] = ctas[k];
    a.n_sched = ns;
    for (int s = 0; s < plan->n_sources; ++s) {
      HGB_CHECK_ARG(src[s] != nullptr, "hgb_msgpack_tcg_forward: source %d is NULL", s);
      a.src[s] = src[s];
      a.src_rows[s] = src_rows ? src_rows[s] : nullptr;
    }
    a.sh = sh; a.g = g_ws; a.gstride = gstride; a.n_edges = n_edges; a.out = out; a.out_index = out_index;
  }
  {
    const int rc = launch_radial_gate(plan, rbf, w3_off, nch, w3img_off, gstride, g_ws, n_edges, st);
    if (rc != 0) return rc;
  }

  if (ctas[0] > 0) {
    constexpr size_t smem = (size_t)SmemG<64, 32>::TOTAL * sizeof(float);
    static_assert(smem <= 113 * 1024, "2 CTAs/SM budget");
    HGB_CUDA_OK(cudaFuncSetAttribute(msgpack_tcg_kernel<64, 256, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    msgpack_tcg_kernel<64, 256, 32><<<(unsigned)ctas[0], 256, smem, st>>>(cls[0]);
    HGB_LAUNCH_OK("msgpack_tcg_kernel<64,256,32>");
  }
  if (use_rows) {
    if (ctas[1] > 0) {
      constexpr size_t smem = (size_t)tcr::Sm<16>::TOTAL * sizeof(float);
      static_assert(smem <= 55 * 1024 + 512, "4 CTAs/SM budget");
      HGB_CUDA_OK(cudaFuncSetAttribute(tcr::msgpack_tcr_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      tcr::msgpack_tcr_kernel<16><<<(unsigned)ctas[1], tcr::NTHR, smem, st>>>(cls[1]);
      HGB_LAUNCH_OK("msgpack_tcr_kernel<16>");
    }
    if (ctas[2] > 0) {
      constexpr size_t smem = (size_t)tcr::Sm<32>::TOTAL * sizeof(float);
      static_assert(smem <= 112 * 1024, "2 CTAs/SM budget");
      HGB_CUDA_OK(cudaFuncSetAttribute(tcr::msgpack_tcr_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      tcr::msgpack_tcr_kernel<32><<<(unsigned)ctas[2], tcr::NTHR, smem, st>>>(cls[2]);
      HGB_LAUNCH_OK("msgpack_tcr_kernel<32>");
    }
  } else if (ctas[1] > 0) {
    constexpr size_t smem = (size_t)SmemG<32, 16>::TOTAL * sizeof(float);
    static_assert(smem <= 56 * 1024, "4 CTAs/SM budget");
    HGB_CUDA_OK(cudaFuncSetAttribute(msgpack_tcg_kernel<32, 128, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    msgpack_tcg_kernel<32, 128, 16><<<(unsigned)ctas[1], 128, smem, st>>>(cls[1]);
    HGB_LAUNCH_OK("msgpack_tcg_kernel<32,128,16>");
  }
  return 0;
}


// ---------------------------------------------------------------------------------------------------------
// Rotated-frame path (msgpack_rot_kernel.cuh)
extern "C" int hgb_wigner(const hgb_rot_plan* rp, const float* edge_vec, int64_t n_edges, float* dw, void* stream) {
  HGB_DEVICE_GUARD(dw);
  HGB_CHECK_ARG(rp && edge_vec && dw && rp->wigner_j, "hgb_wigner: NULL argument");
  HGB_CHECK_ARG(rp->lmax >= 0 && rp->lmax <= rot::LMAX, "hgb_wigner: lmax %d unsupported (<= %d)", rp->lmax, rot::LMAX);
  HGB_CHECK_ARG(n_edges >= 0 && n_edges < (1ll << 31), "hgb_wigner: bad edge count");
  hgb::TimeScope ts_(HGB_K_WIGNER, stream);
  for (int l = 0; l <= rp->lmax; ++l)
    HGB_CHECK_ARG(rp->doff[l] >= 0 && rp->doff[l] + (2 * l + 1) * (2 * l + 1) <= rp->dstride, "hgb_wigner: D^%d block outside the row", l);
  if (n_edges == 0) return 0;
  rot::WigArgs wa;
  memset(&wa, 0, sizeof(wa));
  wa.vec = edge_vec; wa.J = rp->wigner_j; wa.dw = dw; wa.n_edges = n_edges; wa.dstride = rp->dstride;
  for (int l = 0; l <= rp->lmax; ++l) wa.doff[l] = rp->doff[l];
  dim3 grid((unsigned)((n_edges + 127) / 128), (unsigned)(rp->lmax + 1));
  rot::wigner_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(wa);
  HGB_LAUNCH_OK("wigner_kernel");
  return 0;
}

namespace {
template <int RW, int NST>
int launch_rot_class(const rot::RotArgs& ra, int n_tiles, cudaStream_t st) {
  constexpr size_t smem = rot::rot_smem_bytes<RW, NST>();
  static_assert(smem <= 227 * 1024, "msgpack_rot_kernel shared memory");
  HGB_CUDA_OK(cudaFuncSetAttribute(rot::msgpack_rot_kernel<RW, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rot::msgpack_rot_kernel<RW, NST><<<(unsigned)(n_tiles * ra.n_slots), rot::NTHR2, smem, st>>>(ra);
  HGB_LAUNCH_OK("msgpack_rot_kernel");
  return 0;
}
}  // namespace

namespace {
template <int RW, int NST, int NWG, int NMW = 1, bool F16 = false, bool SPLIT = false>
int launch_rotf_class(const typename rotf::ArgsOf<F16>::type& ra, int n_tiles, cudaStream_t st) {
  constexpr size_t smem = rotf::rotf_smem_bytes<RW, NST>();
  static_assert(smem <= 113 * 1024, "msgpack_rotf_kernel shared memory (2 CTAs / SM)");
  HGB_CUDA_OK(cudaFuncSetAttribute(rotf::msgpack_rotf_kernel<RW, NST, NWG, NMW, F16, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rotf::msgpack_rotf_kernel<RW, NST, NWG, NMW, F16, SPLIT><<<(unsigned)(n_tiles * ra.n_slots), 128 * NWG + 128, smem, st>>>(ra);
  HGB_LAUNCH_OK("msgpack_rotf_kernel");
  return 0;
}
}  // namespace

extern "C" int hgb_msgpack_rot_forward(const hgb_msgpack_plan* plan, const hgb_rot_plan* rp, const float* const* src,
                                       const int64_t* const* src_rows, const float* dw, const float* rbf,
                                       const int32_t* w3_off, const int32_t* nch, const int32_t* w3img_off, int32_t gstride,
                                       float* g_ws, float* xp_ws, int64_t chunk_edges, int64_t n_edges, float* out,
                                       const int64_t* out_index, void* stream) {
  HGB_DEVICE_GUARD(out);
  HGB_CHECK_ARG(plan && rp && src && dw && rbf && out && g_ws && xp_ws && w3_off && nch, "hgb_msgpack_rot_forward: NULL argument");
  HGB_CHECK_ARG(plan->types_host && plan->paths_host && rp->blocks_host && rp->steps_host && rp->blocks && rp->steps,
                "hgb_msgpack_rot_forward: host and device copies of the tables are required");
  HGB_CHECK_ARG(plan->n_sources >= 1 && plan->n_sources <= 4 && plan->n_branches >= 1 && plan->n_branches <= 2,
                "hgb_msgpack_rot_forward: bad source/branch count");
  HGB_CHECK_ARG(plan->h2 <= 64 && plan->h1 <= 64 && plan->rbf_dim <= 64,
                "hgb_msgpack_rot_forward: radial MLP [%d,%d,%d] unsupported (all widths <= 64)", plan->rbf_dim, plan->h1, plan->h2);
  HGB_CHECK_ARG(plan->n_types <= 32 && rp->lmax >= 0 && rp->lmax <= rot::LMAX, "hgb_msgpack_rot_forward: too many output slots or l > %d", rot::LMAX);
  HGB_CHECK_ARG(n_edges >= 0 && n_edges < (1ll << 31), "hgb_msgpack_rot_forward: bad edge count");
  HGB_CHECK_ARG(chunk_edges >= rot::TILE && chunk_edges % rot::TILE == 0, "hgb_msgpack_rot_forward: chunk_edges must be a positive multiple of %d", rot::TILE);
  HGB_CHECK_ARG(rp->tile_stride > 0 && rp->tile_stride % 4 == 0 && rp->n_blocks >= 1, "hgb_msgpack_rot_forward: bad packed-input layout");
  if (n_edges == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  for (int b = 0; b < plan->n_branches; ++b)
    HGB_CHECK_ARG(nch[b] > 0 && nch[b] <= gstride, "hgb_msgpack_rot_forward: gate width %d exceeds stride %d", nch[b], gstride);
  for (int s = 0; s < plan->n_sources; ++s) HGB_CHECK_ARG(src[s] != nullptr, "hgb_msgpack_rot_forward: source %d is NULL", s);

  // ---- validate the tables (host copies)
  for (int i = 0; i < rp->n_blocks; ++i) {
    const hgb_rot_block_t& b = rp->blocks_host[i];
    HGB_CHECK_ARG(b.l1 >= 0 && b.l1 <= rp->lmax && b.nsrc >= 1 && b.nsrc <= 2 && b.src0 >= 0 && b.src0 + b.nsrc <= plan->n_sources,
                  "hgb_msgpack_rot_forward: bad block %d", i);
    HGB_CHECK_ARG(b.kpad % 8 == 0 && b.kpad >= b.nsrc * b.mul && b.mul >= 1 && b.in_off >= 0 &&
                      b.in_off + b.mul * (2 * b.l1 + 1) <= plan->src_dim[b.src0],
                  "hgb_msgpack_rot_forward: block %d outside its source row", i);
    HGB_CHECK_ARG(b.xoff >= 0 && b.xoff % 4 == 0 && (int64_t)b.xoff + (int64_t)(2 * b.l1 + 1) * 2 * b.kpad * rot::TILE <= rp->tile_stride,
                  "hgb_msgpack_rot_forward: block %d outside the packed tile", i);
  }
  int klass[32], order[32];
  double cost[32];
  // slot classes 16 / 32 run msgpack_rotf_kernel (L' on the FMA pipes, four GEMM1 accumulators in TMEM) when every slot of the
  // class fits its 256 TMEM columns and every step carries the un-split fp32 L' image; HGB_ROT_FMA = bit mask of the classes
  // (default 3; 0 = msgpack_rot_kernel everywhere)
  // bit 0 / 1: classes 16 / 32 on msgpack_rotf_kernel; bit 2: two gate warpgroups for class 16; bit 3: two GEMM1 issuer warps;
  // bit 4: two ring stages instead of three.  Measured on tbg_m8 (profiles/README.md r05): message kernels 57.5 ms with 0,
  // 50.0 with 1, 68.4 with 2 (class 32 on the FMA pipes spills at 128 registers), 47.4 with 5, 48.8 with 13, 46.2 with 21.
  const int fma_env = getenv("HGB_ROT_FMA") ? atoi(getenv("HGB_ROT_FMA")) : 21;
  bool fma_ok[3] = {(fma_env & 1) != 0, (fma_env & 2) != 0, false};
  bool fma_wg2 = (fma_env & 4) != 0;
  for (int t = 0; t < plan->n_types; ++t) {
    const hgb_type_t& ty = plan->types_host[t];
    const int d3 = 2 * ty.l + 1;
    HGB_CHECK_ARG(ty.l >= 0 && ty.l <= rp->lmax && ty.mpad % 16 == 0 && ty.mpad >= ty.mul && ty.mpad <= NMAX,
                  "hgb_msgpack_rot_forward: slot %d (mul %d, padded %d, l %d) unsupported", t, ty.mul, ty.mpad, ty.l);
    klass[t] = ty.mpad <= 16 ? 0 : (ty.mpad <= 32 ? 1 : 2);
    if (klass[t] < 2) {
      if (rotf::NB * ty.mpad + d3 * ty.mul > 256 || ty.mpad != (klass[t] == 0 ? 16 : 32)) fma_ok[klass[t]] = false;
      if (klass[t] == 0 && rotf::NB * ty.mpad + 2 * d3 * ty.mul > 256) fma_wg2 = false;
      for (int si = rp->step_begin[t]; si < rp->step_begin[t + 1]; ++si)
        if (rp->steps_host[si].pad2 <= 0 || rp->steps_host[si].pad2 % 4 != 0) fma_ok[klass[t]] = false;
    }
    {
      const int dbl = (klass[t] == 1) ? 0 : 1;   // TMEM: B0 B1 GL (GL) S (S) + d3 x mul
      HGB_CHECK_ARG((4 + 2 * dbl) * ty.mpad + d3 * ty.mul <= 512, "hgb_msgpack_rot_forward: slot %d needs more than 512 TMEM columns", t);
    }
    HGB_CHECK_ARG(rp->step_begin[t] >= 0 && rp->step_begin[t] <= rp->step_begin[t + 1], "hgb_msgpack_rot_forward: bad step range of slot %d", t);
    double c = 0;
    int last_m3 = -1, open_group = 0;
    for (int si = rp->step_begin[t]; si < rp->step_begin[t + 1]; ++si) {
      const hgb_rot_step_t& s = rp->steps_host[si];
      HGB_CHECK_ARG(s.kpad >= 8 && s.kpad % 8 == 0 && s.m3 >= 0 && s.m3 < d3 && s.kind == 0 && s.a_off >= 0 && s.a_off % 4 == 0 &&
                        (int64_t)s.a_off + (int64_t)2 * s.kpad * rot::TILE <= rp->tile_stride && s.w_off >= 0 && s.w_off % 4 == 0 &&
                        s.lf_off >= 0 && s.lf_off % 4 == 0 && (s.new_path & 1),
                    "hgb_msgpack_rot_forward: bad step %d", si);
      HGB_CHECK_ARG(s.branch < plan->n_branches && (s.branch < 0 || (s.g_off >= 0 && s.g_off + ty.mul <= nch[s.branch])),
                    "hgb_msgpack_rot_forward: gate columns of step %d out of range", si);
      // steps are grouped by output component (the kernel keeps one component in registers at a time)
      HGB_CHECK_ARG(open_group ? (s.m3 == last_m3) : (s.m3 > last_m3), "hgb_msgpack_rot_forward: step %d breaks the m3 grouping", si);
      last_m3 = s.m3;
      open_group = (s.new_path & 4) ? 0 : 1;
      c += (double)(s.kpad + ty.mpad) * ty.mpad + 600.0;
    }
    HGB_CHECK_ARG(open_group == 0, "hgb_msgpack_rot_forward: slot %d ends inside an m3 group", t);
    cost[t] = c;
    order[t] = t;
  }
  for (int i = 0; i < plan->n_types; ++i)
    for (int j = i + 1; j < plan->n_types; ++j)
      if (cost[order[j]] > cost[order[i]]) { int tmp = order[i]; order[i] = order[j]; order[j] = tmp; }

  rot::RotArgs cls[3];
  for (int k = 0; k < 3; ++k) {
    rot::RotArgs& a = cls[k];
    memset(&a, 0, sizeof(a));
    a.plan = *plan;
    a.steps = rp->steps;
    for (int t = 0; t <= plan->n_types; ++t) a.step_begin[t] = rp->step_begin[t];
    a.tile_stride = rp->tile_stride; a.dstride = rp->dstride;
    for (int l = 0; l <= rp->lmax; ++l) a.doff[l] = rp->doff[l];
    a.gstride = gstride; a.out = out; a.out_index = out_index; a.xp = xp_ws; a.g = g_ws; a.dw = dw;
    int ns = 0;
    for (int q = 0; q < plan->n_types; ++q) {
      const int t = order[q];
      if (klass[t] != k) continue;
      if (rp->step_begin[t] == rp->step_begin[t + 1] && out_index != nullptr) continue;   // nothing to add
      a.slot[ns++] = t;
    }
    a.n_slots = ns;
    a.dbl = (k == 1) ? 0 : 1;
    if (k == 0 && getenv("HGB_ROT_DBL0") && atoi(getenv("HGB_ROT_DBL0"))) a.dbl = 0;   // experiment: single-buffered GL / S for class 16
  }
  rot::RpArgs pa;
  memset(&pa, 0, sizeof(pa));
  pa.blocks = rp->blocks; pa.n_blocks = rp->n_blocks; pa.tile_stride = rp->tile_stride; pa.dstride = rp->dstride;
  for (int l = 0; l <= rp->lmax; ++l) pa.doff[l] = rp->doff[l];
  for (int s = 0; s < plan->n_sources; ++s) {
    pa.src[s] = src[s];
    pa.src_rows[s] = src_rows ? src_rows[s] : nullptr;
    pa.src_dim[s] = plan->src_dim[s];
  }
  pa.dw = dw; pa.xp = xp_ws;
  pa.blocks_per_cta = (rp->n_blocks + 3) / 4;
  const unsigned rp_gy = (unsigned)((rp->n_blocks + pa.blocks_per_cta - 1) / pa.blocks_per_cta);

  for (int64_t e_lo = 0; e_lo < n_edges; e_lo += chunk_edges) {
    const int64_t n = (n_edges - e_lo < chunk_edges) ? (n_edges - e_lo) : chunk_edges;
    const int n_tiles = (int)((n + rot::TILE - 1) / rot::TILE);
    {
      hgb::TimeScope ts(HGB_K_RADIAL_GATE, stream);
      const int rc = launch_radial_gate(plan, rbf + e_lo * plan->rbf_dim, w3_off, nch, w3img_off, gstride, g_ws, n, st, 1);
      if (rc != 0) return rc;
    }
    pa.e_lo = e_lo; pa.n_chunk = n;
    {
      hgb::TimeScope ts(HGB_K_ROTATE_PACK, stream);
      rot::rotate_pack_kernel<<<dim3((unsigned)n_tiles, rp_gy), rot::TILE, 0, st>>>(pa);
      HGB_LAUNCH_OK("rotate_pack_kernel");
    }
    hgb::TimeScope ts_msg(HGB_K_MSGPACK_ROT, stream);   // the three slot classes of msgpack_rot_kernel together
    for (int k = 0; k < 3; ++k) {
      if (cls[k].n_slots == 0) continue;
      cls[k].e_lo = e_lo; cls[k].n_chunk = n;
      int rc = 0;
      if (k == 0) rc = !fma_ok[0] ? launch_rot_class<16, 3>(cls[k], n_tiles, st)
                       : !fma_wg2 ? launch_rotf_class<16, 3, 1>(cls[k], n_tiles, st)
                       : (fma_env & 16) ? launch_rotf_class<16, 2, 2>(cls[k], n_tiles, st)      // experiment: two ring stages
                       : (fma_env & 8) ? launch_rotf_class<16, 3, 2, 2>(cls[k], n_tiles, st)    // experiment: two GEMM1 issuer warps
                                       : launch_rotf_class<16, 3, 2>(cls[k], n_tiles, st);
      else if (k == 1) rc = !fma_ok[1] ? launch_rot_class<32, 2>(cls[k], n_tiles, st)
                                       : launch_rotf_class<32, 2, 2, 1, false, true>(cls[k], n_tiles, st);   // both warpgroups on every step
      else rc = launch_rot_class<64, 2>(cls[k], n_tiles, st);
      if (rc != 0) return rc;
    }
  }
  return 0;
}

// ============================================================================================== rot16 (fp16 x 2 split) host side
namespace {
template <int RW, int NST>
int launch_rot16_class(const rot16::Rot16Args& ra, int n_tiles, cudaStream_t st) {
  constexpr size_t smem = rot16::rot16_smem_bytes<RW, NST>();
  static_assert(smem <= 227 * 1024, "msgpack_rot16_kernel shared memory");
  HGB_CUDA_OK(cudaFuncSetAttribute(rot16::msgpack_rot16_kernel<RW, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rot16::msgpack_rot16_kernel<RW, NST><<<(unsigned)(n_tiles * ra.n_slots), rot::NTHR2, smem, st>>>(ra);
  HGB_LAUNCH_OK("msgpack_rot16_kernel");
  return 0;
}
}  // namespace

extern "C" int hgb_msgpack_rot16_forward(const hgb_msgpack_plan* plan, const hgb_rot_plan* rp, const float* const* src,
                                         const int64_t* const* src_rows, const float* dw, const float* rbf,
                                         const int32_t* w3_off, const int32_t* nch, const int32_t* w3img_off, int32_t gstride,
                                         float* g_ws, float* xp_ws, float* sx_ws, const float* wbuf16, int64_t wbuf16_words,
                                         const float* img_inv, int32_t n_images, int64_t chunk_edges, int64_t n_edges,
                                         float* out, const int64_t* out_index, int32_t flags, void* stream) {
  HGB_DEVICE_GUARD(out);
  HGB_CHECK_ARG(plan && rp && src && dw && rbf && out && g_ws && xp_ws && sx_ws && wbuf16 && img_inv && w3_off && nch,
                "hgb_msgpack_rot16_forward: NULL argument");
  HGB_CHECK_ARG(plan->types_host && plan->paths_host && rp->blocks_host && rp->steps_host && rp->blocks && rp->steps,
                "hgb_msgpack_rot16_forward: host and device copies of the tables are required");
  HGB_CHECK_ARG(plan->n_sources >= 1 && plan->n_sources <= 4 && plan->n_branches >= 1 && plan->n_branches <= 2,
                "hgb_msgpack_rot16_forward: bad source/branch count");
  HGB_CHECK_ARG(plan->h2 <= 64 && plan->h1 <= 64 && plan->rbf_dim <= 64,
                "hgb_msgpack_rot16_forward: radial MLP [%d,%d,%d] unsupported (all widths <= 64)", plan->rbf_dim, plan->h1, plan->h2);
  HGB_CHECK_ARG(plan->n_types <= 32 && rp->lmax >= 0 && rp->lmax <= rot::LMAX, "hgb_msgpack_rot16_forward: too many output slots or l > %d", rot::LMAX);
  HGB_CHECK_ARG(n_edges >= 0 && n_edges < (1ll << 31), "hgb_msgpack_rot16_forward: bad edge count");
  HGB_CHECK_ARG(chunk_edges >= rot::TILE && chunk_edges % rot::TILE == 0, "hgb_msgpack_rot16_forward: chunk_edges must be a positive multiple of %d", rot::TILE);
  HGB_CHECK_ARG(rp->tile_stride > 0 && rp->tile_stride % 4 == 0 && rp->n_blocks >= 1 && rp->n_blocks < 32768, "hgb_msgpack_rot16_forward: bad packed-input layout");
  HGB_CHECK_ARG(n_images >= 1 && n_images <= 65536 && wbuf16_words > 0, "hgb_msgpack_rot16_forward: bad image table");
  if (n_edges == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  for (int b = 0; b < plan->n_branches; ++b)
    HGB_CHECK_ARG(nch[b] > 0 && nch[b] <= gstride, "hgb_msgpack_rot16_forward: gate width %d exceeds stride %d", nch[b], gstride);
  for (int s = 0; s < plan->n_sources; ++s) HGB_CHECK_ARG(src[s] != nullptr, "hgb_msgpack_rot16_forward: source %d is NULL", s);

  // ---- validate the tables (host copies); offsets are in 32-bit words, kpad of a block in channels, kpad of a step in words
  for (int i = 0; i < rp->n_blocks; ++i) {
    const hgb_rot_block_t& b = rp->blocks_host[i];
    HGB_CHECK_ARG(b.l1 >= 0 && b.l1 <= rp->lmax && b.nsrc >= 1 && b.nsrc <= 2 && b.src0 >= 0 && b.src0 + b.nsrc <= plan->n_sources,
                  "hgb_msgpack_rot16_forward: bad block %d", i);
    HGB_CHECK_ARG(b.kpad % 16 == 0 && b.kpad >= b.nsrc * b.mul && b.mul >= 1 && b.in_off >= 0 &&
                      b.in_off + b.mul * (2 * b.l1 + 1) <= plan->src_dim[b.src0],
                  "hgb_msgpack_rot16_forward: block %d outside its source row", i);
    HGB_CHECK_ARG(b.xoff >= 0 && b.xoff % 4 == 0 && (int64_t)b.xoff + (int64_t)(2 * b.l1 + 1) * b.kpad * rot::TILE <= rp->tile_stride,
                  "hgb_msgpack_rot16_forward: block %d outside the packed tile", i);
  }
  int klass[32], order[32];
  double cost[32];
  for (int t = 0; t < plan->n_types; ++t) {
    const hgb_type_t& ty = plan->types_host[t];
    const int d3 = 2 * ty.l + 1;
    HGB_CHECK_ARG(ty.l >= 0 && ty.l <= rp->lmax && ty.mpad % 16 == 0 && ty.mpad >= ty.mul && ty.mpad <= NMAX,
                  "hgb_msgpack_rot16_forward: slot %d (mul %d, padded %d, l %d) unsupported", t, ty.mul, ty.mpad, ty.l);
    klass[t] = ty.mpad <= 16 ? 0 : (ty.mpad <= 32 ? 1 : 2);
    if (klass[t] == 0) {   // msgpack_rotf_kernel<16, 3, 2, 1, F16>: four GEMM1 accumulators + one C' per gate warpgroup
      HGB_CHECK_ARG(rotf::NB * ty.mpad + 2 * d3 * ty.mul <= 256, "hgb_msgpack_rot16_forward: slot %d needs more than 256 TMEM columns", t);
    } else {
      const int dbl = (klass[t] == 1) ? 0 : 1;
      HGB_CHECK_ARG((4 + 2 * dbl) * ty.mpad + d3 * ty.mul <= 512, "hgb_msgpack_rot16_forward: slot %d needs more than 512 TMEM columns", t);
    }
    HGB_CHECK_ARG(rp->step_begin[t] >= 0 && rp->step_begin[t] <= rp->step_begin[t + 1], "hgb_msgpack_rot16_forward: bad step range of slot %d", t);
    double c = 0;
    int last_m3 = -1, open_group = 0;
    for (int si = rp->step_begin[t]; si < rp->step_begin[t + 1]; ++si) {
      const hgb_rot_step_t& s = rp->steps_host[si];
      const uint32_t wimg = (uint32_t)s.pad2 & 0xffffu, limg = (uint32_t)s.pad2 >> 16;
      HGB_CHECK_ARG(s.kpad >= 8 && s.kpad % 8 == 0 && s.m3 >= 0 && s.m3 < d3 && s.kind == 0 && s.a_off >= 0 && s.a_off % 4 == 0 &&
                        (int64_t)s.a_off + (int64_t)2 * s.kpad * rot::TILE <= rp->tile_stride && s.w_off >= 0 && s.w_off % 4 == 0 &&
                        (int64_t)s.w_off + (int64_t)2 * ty.mpad * s.kpad <= wbuf16_words && s.lf_off >= 0 && s.lf_off % 4 == 0 &&
                        (klass[t] == 0 || (int64_t)s.lf_off + (int64_t)ty.mpad * ty.mpad <= wbuf16_words) && (s.new_path & 1) && s.pad >= 0 &&
                        s.pad < rp->n_blocks && (int)wimg < n_images && (int)limg < n_images,
                    "hgb_msgpack_rot16_forward: bad step %d", si);
      HGB_CHECK_ARG(s.branch < plan->n_branches && (s.branch < 0 || (s.g_off >= 0 && s.g_off + ty.mul <= nch[s.branch])),
                    "hgb_msgpack_rot16_forward: gate columns of step %d out of range", si);
      HGB_CHECK_ARG(open_group ? (s.m3 == last_m3) : (s.m3 > last_m3), "hgb_msgpack_rot16_forward: step %d breaks the m3 grouping", si);
      last_m3 = s.m3;
      open_group = (s.new_path & 4) ? 0 : 1;
      c += (double)(2 * s.kpad + ty.mpad) * ty.mpad + 600.0;
    }
    HGB_CHECK_ARG(open_group == 0, "hgb_msgpack_rot16_forward: slot %d ends inside an m3 group", t);
    cost[t] = c;
    order[t] = t;
  }
  for (int i = 0; i < plan->n_types; ++i)
    for (int j = i + 1; j < plan->n_types; ++j)
      if (cost[order[j]] > cost[order[i]]) { int tmp = order[i]; order[i] = order[j]; order[j] = tmp; }

  rot16::Rot16Args cls[3];
  for (int k = 0; k < 3; ++k) {
    rot16::Rot16Args& a = cls[k];
    memset(&a, 0, sizeof(a));
    a.plan = *plan;
    a.steps = rp->steps;
    for (int t = 0; t <= plan->n_types; ++t) a.step_begin[t] = rp->step_begin[t];
    a.tile_stride = rp->tile_stride; a.dstride = rp->dstride;
    for (int l = 0; l <= rp->lmax; ++l) a.doff[l] = rp->doff[l];
    a.gstride = gstride; a.out = out; a.out_index = out_index; a.xp = reinterpret_cast<const uint32_t*>(xp_ws); a.g = g_ws; a.dw = dw;
    a.sx = sx_ws; a.n_blocks = rp->n_blocks; a.img_inv = img_inv; a.wbuf16 = wbuf16; a.swap_halves = flags & 1;
    int ns = 0;
    for (int q = 0; q < plan->n_types; ++q) {
      const int t = order[q];
      if (klass[t] != k) continue;
      if (rp->step_begin[t] == rp->step_begin[t + 1] && out_index != nullptr) continue;
      a.slot[ns++] = t;
    }
    a.n_slots = ns;
    a.dbl = (k == 1) ? 0 : 1;
  }
  rot16::Rp16Args pa;
  memset(&pa, 0, sizeof(pa));
  pa.blocks = rp->blocks; pa.n_blocks = rp->n_blocks; pa.tile_stride = rp->tile_stride; pa.dstride = rp->dstride;
  for (int l = 0; l <= rp->lmax; ++l) pa.doff[l] = rp->doff[l];
  for (int s = 0; s < plan->n_sources; ++s) {
    pa.src[s] = src[s];
    pa.src_rows[s] = src_rows ? src_rows[s] : nullptr;
    pa.src_dim[s] = plan->src_dim[s];
  }
  pa.dw = dw; pa.xp = reinterpret_cast<uint32_t*>(xp_ws); pa.sx = sx_ws;
  pa.blocks_per_cta = (rp->n_blocks + 3) / 4;
  const unsigned rp_gy = (unsigned)((rp->n_blocks + pa.blocks_per_cta - 1) / pa.blocks_per_cta);

  for (int64_t e_lo = 0; e_lo < n_edges; e_lo += chunk_edges) {
    const int64_t n = (n_edges - e_lo < chunk_edges) ? (n_edges - e_lo) : chunk_edges;
    const int n_tiles = (int)((n + rot::TILE - 1) / rot::TILE);
    {
      hgb::TimeScope ts(HGB_K_RADIAL_GATE, stream);
      const int rc = launch_radial_gate(plan, rbf + e_lo * plan->rbf_dim, w3_off, nch, w3img_off, gstride, g_ws, n, st, 1);
      if (rc != 0) return rc;
    }
    pa.e_lo = e_lo; pa.n_chunk = n;
    {
      hgb::TimeScope ts(HGB_K_ROTATE_PACK, stream);
      rot16::rotate_pack16_kernel<<<dim3((unsigned)n_tiles, rp_gy), rot::TILE, 0, st>>>(pa);
      HGB_LAUNCH_OK("rotate_pack16_kernel");
    }
    hgb::TimeScope ts_msg(HGB_K_MSGPACK_ROT, stream);
    for (int k = 0; k < 3; ++k) {
      if (cls[k].n_slots == 0) continue;
      cls[k].e_lo = e_lo; cls[k].n_chunk = n;
      int rc = 0;
      if (k == 0) rc = launch_rotf_class<16, 3, 2, 1, true>(cls[k], n_tiles, st);   // steps of these slots carry the fp32 L' offset
      else if (k == 1) rc = launch_rot16_class<32, 2>(cls[k], n_tiles, st);
      else rc = launch_rot16_class<64, 2>(cls[k], n_tiles, st);
      if (rc != 0) return rc;
    }
  }
  return 0;
}

// ============================================================================================== rot2 (A-stationary) host side
extern "C" int hgb_msgpack_rot2_forward(const hgb_msgpack_plan* plan, const hgb_rot_plan* rp, const hgb_rot2_plan* r2,
                                        const float* const* src, const int64_t* const* src_rows, const float* dw,
                                        const float* rbf, const int32_t* w3_off, const int32_t* nch, const int32_t* w3img_off,
                                        int32_t gstride, float* g_ws, float* xp_ws, float* cp_ws, int64_t chunk_edges,
                                        int64_t n_edges, float* out, const int64_t* seg_ptr, const int64_t* seg_order,
                                        int64_t n_out_rows, void* stream) {
  HGB_DEVICE_GUARD(out);
  HGB_CHECK_ARG(plan && rp && r2 && src && dw && rbf && out && g_ws && xp_ws && cp_ws && w3_off && nch, "hgb_msgpack_rot2_forward: NULL argument");
  HGB_CHECK_ARG(rp->blocks_host && rp->blocks && r2->passes && r2->pieces && r2->batches && r2->dsts && r2->passes_host && r2->pieces_host &&
                    r2->batches_host && r2->dsts_host && (r2->n_gpf == 0 || (r2->gpf && r2->gpf_host)),
                "hgb_msgpack_rot2_forward: host and device copies of the tables are required");
  HGB_CHECK_ARG(plan->n_sources >= 1 && plan->n_sources <= 4 && plan->n_branches >= 1 && plan->n_branches <= 2,
                "hgb_msgpack_rot2_forward: bad source/branch count");
  HGB_CHECK_ARG(plan->h2 <= 64 && plan->h1 <= 64 && plan->rbf_dim <= 64,
                "hgb_msgpack_rot2_forward: radial MLP [%d,%d,%d] unsupported (all widths <= 64)", plan->rbf_dim, plan->h1, plan->h2);
  HGB_CHECK_ARG(rp->lmax >= 0 && rp->lmax <= rot::LMAX && r2->n_slots >= 1 && r2->n_slots <= 32, "hgb_msgpack_rot2_forward: too many output slots or l > %d", rot::LMAX);
  HGB_CHECK_ARG(n_edges >= 0 && n_edges < (1ll << 31), "hgb_msgpack_rot2_forward: bad edge count");
  HGB_CHECK_ARG(chunk_edges >= rot::TILE && chunk_edges % rot::TILE == 0, "hgb_msgpack_rot2_forward: chunk_edges must be a positive multiple of %d", rot::TILE);
  HGB_CHECK_ARG(rp->tile_stride > 0 && rp->tile_stride % 4 == 0 && rp->n_blocks >= 1, "hgb_msgpack_rot2_forward: bad packed-input layout");
  HGB_CHECK_ARG((seg_ptr == nullptr) == (seg_order == nullptr), "hgb_msgpack_rot2_forward: seg_ptr and seg_order go together");
  HGB_CHECK_ARG(seg_ptr ? n_out_rows >= 0 : n_out_rows == n_edges, "hgb_msgpack_rot2_forward: bad output row count");
  HGB_CHECK_ARG(r2->n_passes >= 1 && r2->rowstride >= 1 && r2->rowstride <= plan->out_dim, "hgb_msgpack_rot2_forward: bad pass table");
  cudaStream_t st = (cudaStream_t)stream;
  for (int b = 0; b < plan->n_branches; ++b)
    HGB_CHECK_ARG(nch[b] > 0 && nch[b] + 3 <= gstride && gstride % 4 == 0, "hgb_msgpack_rot2_forward: gate width %d exceeds stride %d", nch[b], gstride);
  for (int s = 0; s < plan->n_sources; ++s) HGB_CHECK_ARG(src[s] != nullptr, "hgb_msgpack_rot2_forward: source %d is NULL", s);

  // ---- validate the tables (host copies): every offset the kernel dereferences
  for (int i = 0; i < rp->n_blocks; ++i) {
    const hgb_rot_block_t& b = rp->blocks_host[i];
    HGB_CHECK_ARG(b.l1 >= 0 && b.l1 <= rp->lmax && b.nsrc >= 1 && b.nsrc <= 2 && b.src0 >= 0 && b.src0 + b.nsrc <= plan->n_sources,
                  "hgb_msgpack_rot2_forward: bad block %d", i);
    HGB_CHECK_ARG(b.kpad % 8 == 0 && b.kpad >= b.nsrc * b.mul && b.mul >= 1 && b.in_off >= 0 &&
                      b.in_off + b.mul * (2 * b.l1 + 1) <= plan->src_dim[b.src0],
                  "hgb_msgpack_rot2_forward: block %d outside its source row", i);
    HGB_CHECK_ARG(b.xoff >= 0 && b.xoff % 4 == 0 && (int64_t)b.xoff + (int64_t)(2 * b.l1 + 1) * 2 * b.kpad * rot::TILE <= rp->tile_stride,
                  "hgb_msgpack_rot2_forward: block %d outside the packed tile", i);
  }
  int col_cover = 0;
  for (int p = 0; p < r2->n_passes; ++p) {
    const hgb_rot2_pass_t& ps = r2->passes_host[p];
    HGB_CHECK_ARG(ps.piece_begin >= 0 && ps.piece_begin < ps.piece_end && ps.piece_end <= r2->n_pieces && ps.ncols >= 1 &&
                      ps.ncols <= rot2::ACC_COLS && ps.out_col0 == col_cover,
                  "hgb_msgpack_rot2_forward: bad pass %d", p);
    for (int h = 0; h < rot2::NH; ++h)
      HGB_CHECK_ARG(ps.stream_begin[h] >= 0 && ps.stream_begin[h] <= ps.stream_end[h] && ps.stream_end[h] <= r2->n_batches,
                    "hgb_msgpack_rot2_forward: bad gate stream %d of pass %d", h, p);
    col_cover += ps.ncols;
    for (int q = ps.piece_begin; q < ps.piece_end; ++q) {
      const hgb_rot2_piece_t& pc = r2->pieces_host[q];
      HGB_CHECK_ARG(pc.kpad >= 8 && pc.kpad % 8 == 0 && pc.ncols >= 16 && pc.ncols % 16 == 0 && pc.ncols <= rot2::NB && pc.a_off >= 0 &&
                        pc.a_off % 4 == 0 && (int64_t)pc.a_off + (int64_t)2 * pc.kpad * rot::TILE <= rp->tile_stride && pc.w_off >= 0 &&
                        pc.w_off % 4 == 0 && pc.l_off >= 0 && pc.l_off % 4 == 0 && pc.l_floats > 0 && pc.l_floats % 4 == 0 &&
                        pc.l_floats <= rot2::LBUF && pc.ndst >= 0 && pc.dst_begin >= 0 && pc.dst_begin + pc.ndst <= r2->n_dsts &&
                        pc.gpf_n >= 0 && pc.gpf_n <= 32 && pc.gpf_begin >= 0 && (int64_t)pc.gpf_begin + pc.gpf_n <= r2->n_gpf,
                    "hgb_msgpack_rot2_forward: bad piece %d", q);
      for (int k = pc.gpf_begin; k < pc.gpf_begin + pc.gpf_n; ++k) {
        const hgb_rot2_gpf_t& gr = r2->gpf_host[k];
        HGB_CHECK_ARG(gr.off % 4 == 0 && gr.bytes % 16 == 0 && gr.bytes > 0 &&
                          (uint64_t)gr.off * 4 + gr.bytes <= (uint64_t)plan->n_branches * gstride * 128 * 4,
                      "hgb_msgpack_rot2_forward: gate prefetch run %d out of range", k);
      }
      int s_used = 0;
      for (int d = pc.dst_begin; d < pc.dst_begin + pc.ndst; ++d) {
        const hgb_rot2_dst_t& ds = r2->dsts_host[d];
        HGB_CHECK_ARG(ds.col0 >= 0 && ds.col0 % 8 == 0 && ds.kcols >= 8 && ds.kcols % 8 == 0 && ds.col0 + ds.kcols <= pc.ncols &&
                          ds.mp >= 16 && ds.mp % 16 == 0 && ds.s_off == s_used && ds.s_off + ds.mp <= rot2::SW && ds.mul >= 1 &&
                          ds.mul <= ds.mp && ds.acc_col0 >= 0 && ds.acc_col0 + ds.mul <= ps.ncols && ds.l_rel >= 0 && ds.l_rel % 4 == 0 &&
                          ds.l_rel + 2 * ds.kcols * ds.mp <= pc.l_floats,
                      "hgb_msgpack_rot2_forward: bad destination group %d", d);
        s_used += ds.mp;
      }
    }
    // the two gate streams: each visits every piece of the pass once (first ... last flags), in piece order
    for (int h = 0; h < rot2::NH; ++h) {
      const int sb = ps.stream_begin[h], se = ps.stream_end[h];
      int q = ps.piece_begin, open = 0;
      for (int b = sb; b < se; ++b) {
        const hgb_rot2_batch_t& bt = r2->batches_host[b];
        const uint32_t meta = (uint32_t)bt.meta;
        const int kind = (int)(meta & 3u);
        HGB_CHECK_ARG(q < ps.piece_end && (((meta >> 2) & 1u) != 0) == (open == 0), "hgb_msgpack_rot2_forward: gate stream %d of pass %d is not piece-aligned at %d", h, p, b);
        open = 1;
        const hgb_rot2_piece_t& pc = r2->pieces_host[q];
        const uint32_t gmax = (uint32_t)plan->n_branches * (uint32_t)gstride * 128u;
        HGB_CHECK_ARG(kind <= 2 && (bt.goff_a == 0xFFFFFFFFu || (bt.goff_a % 128u == 0 && bt.goff_a + 4u * 128u <= gmax)) &&
                          (bt.goff_b == 0xFFFFFFFFu || (bt.goff_b % 128u == 0 && bt.goff_b + 4u * 128u <= gmax)),
                      "hgb_msgpack_rot2_forward: gate batch %d out of range", b);
        if (kind != 2) HGB_CHECK_ARG((int)((meta >> 8) & 0xFFu) * 8 + 8 <= pc.ncols, "hgb_msgpack_rot2_forward: gate batch %d outside its piece", b);
        if (kind == 1) {
          const int mul = (int)((meta >> 16) & 0x1Fu), acc0 = (int)((meta >> 21) & 0xFFu), m4 = (mul + 3) & ~3;
          HGB_CHECK_ARG(mul >= 1 && mul <= 16 && acc0 + mul <= ps.ncols && bt.l_off >= 0 && bt.l_off % 4 == 0 && bt.l_off + 8 * m4 <= pc.l_floats,
                        "hgb_msgpack_rot2_forward: bad FMA-pipe batch %d", b);
        }
        if ((meta >> 3) & 1u) { open = 0; ++q; }
      }
      HGB_CHECK_ARG(open == 0 && q == ps.piece_end, "hgb_msgpack_rot2_forward: gate stream %d of pass %d does not cover its pieces", h, p);
    }
  }
  HGB_CHECK_ARG(col_cover == r2->rowstride, "hgb_msgpack_rot2_forward: passes cover %d of %d columns", col_cover, r2->rowstride);

  rot2::UnrotArgs ua;
  memset(&ua, 0, sizeof(ua));
  ua.cp = cp_ws; ua.rowstride = r2->rowstride; ua.dw = dw; ua.dstride = rp->dstride;
  HGB_CHECK_ARG(rp->dstride <= 480, "hgb_msgpack_rot2_forward: Wigner row of %d floats exceeds the staging buffer", rp->dstride);
  for (int l = 0; l <= rp->lmax; ++l) ua.doff[l] = rp->doff[l];
  int n_items = 0;
  for (int t = 0; t < r2->n_slots; ++t) {
    HGB_CHECK_ARG(r2->slot_l[t] >= 0 && r2->slot_l[t] <= rp->lmax && r2->slot_mul[t] >= 1 && r2->slot_out_off[t] >= 0 &&
                      r2->slot_out_off[t] + r2->slot_mul[t] * (2 * r2->slot_l[t] + 1) <= plan->out_dim,
                  "hgb_msgpack_rot2_forward: bad slot %d", t);
    ua.slot_l[t] = r2->slot_l[t]; ua.slot_mul[t] = r2->slot_mul[t]; ua.slot_out_off[t] = r2->slot_out_off[t];
    for (int m = 0; m < 13; ++m) {
      ua.ccol[t][m] = r2->ccol[t][m];
      HGB_CHECK_ARG(m < 2 * r2->slot_l[t] + 1 ? (r2->ccol[t][m] >= -1 && r2->ccol[t][m] + r2->slot_mul[t] <= r2->rowstride) : true,
                    "hgb_msgpack_rot2_forward: bad cp column of slot %d", t);
    }
    for (int w0 = 0; w0 < r2->slot_mul[t]; w0 += 32) {
      HGB_CHECK_ARG(n_items < 64, "hgb_msgpack_rot2_forward: more than 64 (slot, 32-channel) groups");
      ua.item_slot[n_items] = t; ua.item_w0[n_items] = w0; ++n_items;
    }
  }
  ua.n_items = n_items;
  ua.n_warps = n_items < 32 ? n_items : 32;
  ua.seg_ptr = seg_ptr; ua.seg_order = seg_order; ua.n_rows = n_out_rows; ua.out = out; ua.out_dim = plan->out_dim;
  if (n_edges == 0) {
    if (n_out_rows > 0) HGB_CUDA_OK(cudaMemsetAsync(out, 0, (size_t)n_out_rows * plan->out_dim * sizeof(float), st));
    return 0;
  }

  rot::RpArgs pa;
  memset(&pa, 0, sizeof(pa));
  pa.blocks = rp->blocks; pa.n_blocks = rp->n_blocks; pa.tile_stride = rp->tile_stride; pa.dstride = rp->dstride;
  for (int l = 0; l <= rp->lmax; ++l) pa.doff[l] = rp->doff[l];
  for (int s = 0; s < plan->n_sources; ++s) {
    pa.src[s] = src[s];
    pa.src_rows[s] = src_rows ? src_rows[s] : nullptr;
    pa.src_dim[s] = plan->src_dim[s];
  }
  pa.dw = dw; pa.xp = xp_ws;
  pa.blocks_per_cta = (rp->n_blocks + 3) / 4;
  const unsigned rp_gy = (unsigned)((rp->n_blocks + pa.blocks_per_cta - 1) / pa.blocks_per_cta);

  rot2::Args ka;
  memset(&ka, 0, sizeof(ka));
  ka.wbuf = plan->wbuf; ka.passes = r2->passes; ka.pieces = r2->pieces; ka.batches = r2->batches; ka.dsts = r2->dsts; ka.gpf = r2->gpf;
  ka.n_passes = r2->n_passes; ka.xp = xp_ws; ka.tile_stride = rp->tile_stride; ka.g = g_ws; ka.gtile_floats = plan->n_branches * gstride * 128;
  ka.cp = cp_ws; ka.rowstride = r2->rowstride;
  { const char* fl = getenv("HGB_ROT2_FLAGS"); ka.flags = fl ? atoi(fl) : 0; }
  static_assert(rot2::SMEM_BYTES <= 227 * 1024, "msgpack_rot2_kernel: shared memory budget");
  HGB_CUDA_OK(cudaFuncSetAttribute(rot2::msgpack_rot2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rot2::SMEM_BYTES));

  for (int64_t e_lo = 0; e_lo < n_edges; e_lo += chunk_edges) {
    const int64_t n = (n_edges - e_lo < chunk_edges) ? (n_edges - e_lo) : chunk_edges;
    const int n_tiles = (int)((n + rot::TILE - 1) / rot::TILE);
    {
      hgb::TimeScope ts(HGB_K_RADIAL_GATE, stream);
      const int rc = launch_radial_gate(plan, rbf + e_lo * plan->rbf_dim, w3_off, nch, w3img_off, gstride, g_ws, n, st, 2);
      if (rc != 0) return rc;
    }
    pa.e_lo = e_lo; pa.n_chunk = n;
    {
      hgb::TimeScope ts(HGB_K_ROTATE_PACK, stream);
      rot::rotate_pack_kernel<<<dim3((unsigned)n_tiles, rp_gy), rot::TILE, 0, st>>>(pa);
      HGB_LAUNCH_OK("rotate_pack_kernel");
    }
    ka.e_lo = e_lo; ka.n_chunk = n;
    // HGB_ROT2_TRACE=<file>: per-role clock64 stamps of one CTA (pass HGB_ROT2_TRACE_PASS of tile 0) appended to <file>;
    // a diagnosis aid (one extra sync per launch), off unless the variable is set
    const char* trace_path = getenv("HGB_ROT2_TRACE");
    long long* d_trace = nullptr;
    const int trace_pieces = 128;
    if (trace_path && trace_path[0]) {
      const char* tp = getenv("HGB_ROT2_TRACE_PASS");
      ka.trace_cta = tp ? atoi(tp) : 5;
      if (ka.trace_cta >= r2->n_passes) ka.trace_cta = r2->n_passes - 1;
      ka.trace_pieces = trace_pieces;
      HGB_CUDA_OK(cudaMalloc(&d_trace, sizeof(long long) * (10 * trace_pieces + 1024)));
      HGB_CUDA_OK(cudaMemsetAsync(d_trace, 0, sizeof(long long) * (10 * trace_pieces + 1024), st));
      ka.trace = d_trace;
    }
    {
      hgb::TimeScope ts(HGB_K_MSGPACK_ROT2, stream);
      rot2::msgpack_rot2_kernel<<<(unsigned)(n_tiles * r2->n_passes), rot2::NTHR, rot2::SMEM_BYTES, st>>>(ka);
      HGB_LAUNCH_OK("msgpack_rot2_kernel");
    }
    if (d_trace) {
      static long long h_trace[10 * 128 + 1024];
      HGB_CUDA_OK(cudaStreamSynchronize(st));
      HGB_CUDA_OK(cudaMemcpy(h_trace, d_trace, sizeof(h_trace), cudaMemcpyDeviceToHost));
      cudaFree(d_trace);
      ka.trace = nullptr;
      FILE* f = fopen(trace_path, "a");
      if (f) {
        const hgb_rot2_pass_t& tps = r2->passes_host[ka.trace_cta];
        fprintf(f, "# launch n_tiles %d pass %d pieces %d\n", n_tiles, ka.trace_cta, tps.piece_end - tps.piece_begin);
        long long t0 = 0;
        for (int i = 0; i < 5 * trace_pieces * 2; ++i)
          if (h_trace[i] && (!t0 || h_trace[i] < t0)) t0 = h_trace[i];
        for (int n2 = 0; n2 < trace_pieces && n2 < tps.piece_end - tps.piece_begin; ++n2) {
          const hgb_rot2_piece_t& pc = r2->pieces_host[tps.piece_begin + n2];
          fprintf(f, "%d kpad %d ncols %d ndst %d |", n2, pc.kpad, pc.ncols, pc.ndst);
          for (int role = 0; role < 5; ++role)
            fprintf(f, " %lld %lld |", h_trace[(role * trace_pieces + n2) * 2] - t0, h_trace[(role * trace_pieces + n2) * 2 + 1] - t0);
          fprintf(f, "\n");
        }
        fprintf(f, "# gate warp 0 batches: start | after tmem ld | after gate values | end (kind)\n");
        for (int b2 = 0; b2 < 256; ++b2) {
          const long long* q4 = h_trace + 10 * trace_pieces + b2 * 4;
          if (!q4[0]) break;
          fprintf(f, "b%d %lld %lld %lld %lld k%lld\n", b2, q4[0] - t0, q4[1] ? q4[1] - t0 : 0, q4[2] ? q4[2] - t0 : 0,
                  (q4[3] & 0x0fffffffffffffffll) - t0, (q4[3] >> 60) & 3);
        }
        fclose(f);
      }
    }
  }
  if (n_out_rows > 0) {
    hgb::TimeScope ts(HGB_K_UNROTATE, stream);
    rot2::unrotate_kernel<<<(unsigned)n_out_rows, (unsigned)(32 * ua.n_warps), 0, st>>>(ua);
    HGB_LAUNCH_OK("unrotate_kernel");
  }
  return 0;
}
