// Edge-aligned fused MessagePackBlock with fp16 x 2 split operands ("rot16").  Included by msgpack_tcg.cu after
// msgpack_rot_kernel.cuh (same step structure, same role layout, same radial-gate pre-pass, same Wigner matrices).
//
// Why: the tf32 kernel is paced by the NUMBER of tcgen05.mma instructions -- ~55 cycles each for any N <= 96
// (profiles/r02g_mma_probe.txt), ~60 000 of them per 128-edge tile and message -- and one kind::tf32 instruction covers K = 8.
// A kind::f16 instruction covers K = 16 in the same 32 operand bytes per row.  fp32 accuracy is kept the same way as 3xTF32:
// every operand is split a = hi + lo with hi = fp16(a), lo = fp16(a - hi) and three MMAs accumulate lo*hi + hi*lo + hi*hi in
// the fp32 TMEM accumulator: 22 significand bits, like the tf32 split.  fp16 has a 5-bit exponent, so every operand is first
// multiplied by a power of two (exact) that puts its largest element just below 2^15:
//   X'   per (edge, input block): max |x| of the block's row segment, bound sqrt(d1) on the rotation -> rotate_pack16_kernel,
//        inverse scale kept per edge in `sx`;
//   W, L' per image, on the host side of the ABI (the packing program), inverse scales in `img_inv`;
//   (X'W).g  per (edge, step): row maximum over the mp columns, taken by the gate thread that owns the TMEM lane.
// Elements more than 2^18 below the largest of their row / image lose relative (not absolute) precision: absolute error
// <= 2^-40 of the largest.  All inverse scales are folded back in fp32 (gate factor, accumulate factor).
//
// In units of 32-bit words an fp16 image of K channels is laid out exactly like a tf32 image of K / 2 channels (word j of a
// row = channels 2j | 2j+1 << 16), so descriptors, chunking (32 words = 64 channels per chunk) and TMEM addressing are those of
// msgpack_rot_kernel with kpad -> kpad / 2; the packed gated product occupies mp / 2 TMEM columns.
#pragma once


namespace rot16 {
using namespace tcmsg;
using rot::TILE;
using rot::KC;      // 32-bit words per operand chunk row = 64 channels
using rot::NTHR2;
using rot::wait_a;
using rot::warp_wait_a;
using rot::arrive_a;
using rot::expect_tx_a;
using rot::bulk_g2s_a;
using rot::commit_a;
using rot::elect_one;
using rot::bulk_prefetch_l2;
using rot::tmem_alloc_dyn;
using rot::tmem_dealloc_dyn;
using rot::tmem_st1;

// kind::f16 (fp16 operands), fp32 accumulate, A and B K-major, M = 128
__host__ __device__ constexpr uint32_t idesc_f16_m128(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); }

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from TMEM: lane = row, 32-bit column j = (k = 2j | k = 2j+1 << 16)
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// (a, b) -> hi word = fp16(a) | fp16(b) << 16, lo word = fp16(a - hi_a) | fp16(b - hi_b) << 16
__device__ __forceinline__ void split_f16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// power of two s with amax * bound_factor * s < 2^15, and its inverse, from the biased exponent of amax (amax >= 0):
// amax < 2^(E-126)  ->  s = 2^(141 - E - shift) with 2^shift >= bound_factor.  amax == 0 (or denormal): s = 1.
__device__ __forceinline__ void pow2_scale(float amax, int shift, float& s, float& inv) {
  const int E = (int)((__float_as_uint(amax) >> 23) & 0xffu);
  int sb = 127 + 141 - E - shift;           // biased exponent of s
  sb = (E == 0) ? 127 : min(max(sb, 1), 253);
  s = __uint_as_float((uint32_t)sb << 23);
  inv = __uint_as_float((uint32_t)(254 - sb) << 23);
}

// ================================================================================================ rotate + pack
struct Rp16Args {
  const hgb_rot_block_t* blocks;
  int n_blocks, blocks_per_cta;
  int tile_stride, dstride;   // tile_stride in 32-bit words
  int doff[12];
  const float* src[4];
  const int64_t* src_rows[4];
  int src_dim[4];
  const float* dw;
  int64_t e_lo, n_chunk;
  uint32_t* xp;               // [tile][tile_stride] words
  float* sx;                  // [tile][n_blocks][128] inverse scale of the block's operand rows
};

template <int L1>
__device__ __forceinline__ void rotpack16_block(const Rp16Args& a, const hgb_rot_block_t& b, int bi, int tile, int z, int64_t e, bool live) {
  constexpr int d1 = 2 * L1 + 1;
  const int K = b.nsrc * b.mul, kpad = b.kpad, kw = kpad >> 1;
  const float* r0 = nullptr;
  const float* r1 = nullptr;
  const float* Dz = nullptr;
  float s = 1.f, inv = 1.f;
  if (live) {
    const int s0 = b.src0, s1 = b.src0 + b.nsrc - 1;
    const int64_t row0 = a.src_rows[s0] ? a.src_rows[s0][e] : e;
    const int64_t row1 = a.src_rows[s1] ? a.src_rows[s1][e] : e;
    r0 = a.src[s0] + row0 * a.src_dim[s0] + b.in_off;
    r1 = a.src[s1] + row1 * a.src_dim[s1] + b.in_off;
    Dz = a.dw + e * a.dstride + a.doff[L1];
    // row maximum of the block: |x'_{u,m}| <= ||x_u|| <= sqrt(d1) max|x| < 4 max|x|
    const int n = b.mul * d1;
    float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
    int i = 0;
    for (; i + 4 <= n; i += 4) {
      m0 = fmaxf(m0, fabsf(__ldg(r0 + i))); m1 = fmaxf(m1, fabsf(__ldg(r0 + i + 1)));
      m2 = fmaxf(m2, fabsf(__ldg(r0 + i + 2))); m3 = fmaxf(m3, fabsf(__ldg(r0 + i + 3)));
    }
    for (; i < n; ++i) m0 = fmaxf(m0, fabsf(__ldg(r0 + i)));
    if (b.nsrc > 1) {
      i = 0;
      for (; i + 4 <= n; i += 4) {
        m0 = fmaxf(m0, fabsf(__ldg(r1 + i))); m1 = fmaxf(m1, fabsf(__ldg(r1 + i + 1)));
        m2 = fmaxf(m2, fabsf(__ldg(r1 + i + 2))); m3 = fmaxf(m3, fabsf(__ldg(r1 + i + 3)));
      }
      for (; i < n; ++i) m0 = fmaxf(m0, fabsf(__ldg(r1 + i)));
    }
    pow2_scale(fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)), (L1 == 0) ? 0 : 2, s, inv);
  }
  a.sx[((size_t)tile * a.n_blocks + bi) * TILE + z] = inv;
  uint32_t* xo = a.xp + (size_t)tile * a.tile_stride + b.xoff + z * 4;
  const size_t per_m = (size_t)2 * kw * TILE;
  for (int q = 0; q < (kpad >> 2); ++q) {   // 4 channels = 2 words
    float x[4][d1];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int u = 4 * q + c;
      const bool okc = live && u < K;
      const bool second = u >= b.mul;
      const float* p = okc ? ((second ? r1 : r0) + (u - (second ? b.mul : 0)) * d1) : nullptr;
#pragma unroll
      for (int i = 0; i < d1; ++i) x[c][i] = okc ? __ldg(p + i) * s : 0.f;
    }
    const int w0 = 2 * q;                                   // first word-channel of the quad
    const int chunk = w0 / KC, wl = w0 - chunk * KC, kc = min(KC, kw - chunk * KC);
    uint32_t* base = xo + (size_t)chunk * 2 * KC * TILE + (size_t)(wl >> 2) * (TILE * 4) + (wl & 3);
#pragma unroll
    for (int m = 0; m < d1; ++m) {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (live) {
        if (L1 == 0) {
#pragma unroll
          for (int c = 0; c < 4; ++c) v[c] = x[c][0];
        } else {
#pragma unroll
          for (int i = 0; i < d1; ++i) {
            const float dmi = __ldg(Dz + m * d1 + i);
#pragma unroll
            for (int c = 0; c < 4; ++c) v[c] = fmaf(dmi, x[c][i], v[c]);
          }
        }
      }
      uint2 h, l;
      split_f16x2(v[0], v[1], h.x, l.x);
      split_f16x2(v[2], v[3], h.y, l.y);
      uint32_t* dst = base + m * per_m;
      *reinterpret_cast<uint2*>(dst) = h;
      *reinterpret_cast<uint2*>(dst + (size_t)kc * TILE) = l;
    }
  }
}

__global__ void __launch_bounds__(TILE) rotate_pack16_kernel(const __grid_constant__ Rp16Args a) {
  const int tile = blockIdx.x, z = threadIdx.x;
  const int64_t el = (int64_t)tile * TILE + z;
  const bool live = el < a.n_chunk;
  const int64_t e = a.e_lo + el;
  const int b0 = blockIdx.y * a.blocks_per_cta, b1 = min(a.n_blocks, b0 + a.blocks_per_cta);
  for (int bi = b0; bi < b1; ++bi) {
    const hgb_rot_block_t b = a.blocks[bi];
    switch (b.l1) {
      case 0: rotpack16_block<0>(a, b, bi, tile, z, e, live); break;
      case 1: rotpack16_block<1>(a, b, bi, tile, z, e, live); break;
      case 2: rotpack16_block<2>(a, b, bi, tile, z, e, live); break;
      case 3: rotpack16_block<3>(a, b, bi, tile, z, e, live); break;
      case 4: rotpack16_block<4>(a, b, bi, tile, z, e, live); break;
      case 5: rotpack16_block<5>(a, b, bi, tile, z, e, live); break;
      default: rotpack16_block<6>(a, b, bi, tile, z, e, live); break;
    }
  }
}

// ================================================================================================= message kernel
struct Rot16Args {
  hgb_msgpack_plan plan;
  const hgb_rot_step_t* steps;   // the fp16 program: kpad = WORDS per operand row, pad = input block, pad2 = W image | L' image << 16
  int step_begin[33];
  const uint32_t* xp;
  int tile_stride;
  const float* dw;
  int dstride;
  int doff[12];
  const float* g;
  int gstride;
  const float* sx;               // [tile][n_blocks][128]
  int n_blocks;
  const float* img_inv;          // inverse power-of-two scale per packed image
  const float* wbuf16;           // packed fp16 images (word offsets)
  int64_t e_lo, n_chunk;
  float* out;
  const int64_t* out_index;
  int n_slots;
  int slot[32];
  int dbl;
  int swap_halves;               // debug: pack the gated product as (k odd | k even << 16)
};

// Same roles and barriers as msgpack_rot_kernel: warps 0-3 gate / accumulate, 4 GEMM1, 5 TMA (A chunks), 6 GEMM2,
// 7 TMA (W chunks), 8 TMA (L' images) + gate L2 prefetch.
template <int RW, int NST>
__global__ void __launch_bounds__(NTHR2, (RW == 64 ? 1 : 2)) msgpack_rot16_kernel(const __grid_constant__ Rot16Args a) {
  constexpr int STG = 2 * KC * TILE + 2 * RW * KC;   // words per ring stage: A chunk (hi | lo) + W chunk (hi | lo)
  extern __shared__ __align__(128) float smem[];
  __shared__ uint64_t bars[2 * NST + 8];
  __shared__ uint32_t tmem_slot;
  const uint32_t bar0 = tc::smem_u32(bars);
  const uint32_t B_FULL = bar0, B_EMPTY = bar0 + 8 * NST, B_LFULL = bar0 + 16 * NST, B_BFULL = B_LFULL + 16, B_GFULL = B_LFULL + 32,
                 B_S2 = B_LFULL + 48;
  const uint32_t stage0 = tc::smem_u32(smem);
  const uint32_t sl0 = stage0 + (uint32_t)(NST * STG) * 4u;     // 2 x (hi | lo) L' images, RW^2 words each (RW/2 word rows x RW x 2)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x / a.n_slots;
  const int t = a.slot[blockIdx.x - tile * a.n_slots];
  const hgb_type_t ty = a.plan.types[t];
  const int d3 = 2 * ty.l + 1, mp = ty.mpad, mul = ty.mul, mw = mp >> 1;
  const int sb = a.step_begin[t], se = a.step_begin[t + 1];
  const int dbl = a.dbl;
  // TMEM columns: B0 | B1 (fp32 GEMM1 results; the packed hi product overwrites the first mp/2 columns) | GL0 (| GL1) | S0 (| S1) | C'
  const uint32_t TB0 = 0, TGL0 = 2 * mp, TS0 = (uint32_t)((3 + dbl) * mp), TC = (uint32_t)((4 + 2 * dbl) * mp);
  uint32_t ncols = 32;
  while (ncols < TC + (uint32_t)(d3 * mul)) ncols <<= 1;

  if (tid == 0) {
    for (int i = 0; i < 2 * NST + 8; ++i) tc::mbar_init(&bars[i], (i >= 2 * NST + 4 && i < 2 * NST + 6) ? 4 : (i < NST ? 2 : 1));
    tc::mbar_fence_init();
  }
  if (warp == 4) tmem_alloc_dyn(&tmem_slot, ncols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const float* __restrict__ wbuf = a.wbuf16;
  const uint32_t idesc = idesc_f16_m128(mp);
  const uint32_t dhi = tc::smem_desc_hi(128);
  const uint32_t lbo_a = TILE * 16, lbo_n = (uint32_t)mp * 16;
  const uint32_t astep = (2 * lbo_a) >> 4, bstep = (2 * lbo_n) >> 4;   // one MMA = 8 words = 16 channels

  if (warp == 5 || warp == 7 || warp == 8) {
    if (lane == 0) {
      const uint32_t* xt = a.xp + (size_t)tile * a.tile_stride;
      if (warp == 8) {
        constexpr int GPF = 3;
        const size_t g_bstride = (size_t)((a.n_chunk + TILE - 1) / TILE) * a.gstride * TILE;
        const float* gt = a.g + (size_t)tile * a.gstride * TILE;
        auto prefetch_gate = [&](int sj) {
          if (sj < se) {
            const hgb_rot_step_t* ps = a.steps + sj;
            if (ps->branch >= 0) bulk_prefetch_l2(gt + (size_t)ps->branch * g_bstride + (size_t)ps->g_off * TILE, (uint32_t)(mul * TILE) * 4u);
          }
        };
        for (int j = 0; j < GPF; ++j) prefetch_gate(sb + j);
        int n = 0;
        const uint32_t lbytes = (uint32_t)(mp * mp) * 4u;   // hi | lo, mp/2 word rows x mp x 4 bytes each
        for (int si = sb; si < se; ++si, ++n) {
          const hgb_rot_step_t st = a.steps[si];
          prefetch_gate(si + GPF);
          const int lb = n & 1;
          if (n >= 2) wait_a(B_S2 + 8 * lb, (uint32_t)(((n >> 1) - 1) & 1));
          expect_tx_a(B_LFULL + 8 * lb, lbytes);
          bulk_g2s_a(sl0 + (uint32_t)(lb * RW * RW) * 4u, wbuf + st.lf_off, lbytes, B_LFULL + 8 * lb);
        }
      } else {
        const bool isA = warp == 5;
        int c_all = 0;
        for (int si = sb; si < se; ++si) {
          const hgb_rot_step_t st = a.steps[si];
          const int kw = st.kpad;
          for (int u0 = 0, c = 0; u0 < kw; u0 += KC, ++c, ++c_all) {
            const int kc = min(KC, kw - u0), s = c_all % NST;
            if (c_all >= NST) wait_a(B_EMPTY + 8 * s, (uint32_t)(((c_all / NST) - 1) & 1));
            const uint32_t sa = stage0 + (uint32_t)(s * STG) * 4u;
            const uint32_t ab = (uint32_t)(kc * TILE * 2) * 4u, wb = (uint32_t)(2 * mp * kc) * 4u;
            if (isA) {
              expect_tx_a(B_FULL + 8 * s, ab);
              bulk_g2s_a(sa, xt + st.a_off + (size_t)c * (2 * KC * TILE), ab, B_FULL + 8 * s);
            } else {
              expect_tx_a(B_FULL + 8 * s, wb);
              bulk_g2s_a(sa + 2 * KC * TILE * 4, wbuf + st.w_off + (size_t)c * (2 * mp * KC), wb, B_FULL + 8 * s);
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 4) {
    // =============================== GEMM1 issuer ===============================
    int n = 0, c_all = 0;
    int kw = (sb < se) ? a.steps[sb].kpad : 0;
    for (int si = sb; si < se; ++si, ++n) {
      const int kw_next = (si + 1 < se) ? a.steps[si + 1].kpad : 0;
      if (n >= 2) warp_wait_a(B_S2 + 8 * (n & 1), (uint32_t)(((n >> 1) - 1) & 1));
      const uint32_t dcol = tmem + TB0 + (uint32_t)((n & 1) * mp);
      for (int u0 = 0, c = 0; u0 < kw; u0 += KC, ++c, ++c_all) {
        const int kc = min(KC, kw - u0), s = c_all % NST;
        warp_wait_a(B_FULL + 8 * s, (uint32_t)((c_all / NST) & 1));
        tc::fence_after_sync();
        if (elect_one()) {
          const uint32_t sa = stage0 + (uint32_t)(s * STG) * 4u;
          const uint32_t ah = tc::smem_desc_lo(sa, lbo_a), al = ah + (((uint32_t)kc * TILE * 4) >> 4);
          const uint32_t wh = tc::smem_desc_lo(sa + 2 * KC * TILE * 4, lbo_n), wl = wh + (((uint32_t)mp * kc * 4) >> 4);
          for (int k8 = 0; k8 < (kc >> 3); ++k8) {
            const uint64_t dah = tc::desc64(ah + k8 * astep, dhi), dal = tc::desc64(al + k8 * astep, dhi);
            const uint64_t dbh = tc::desc64(wh + k8 * bstep, dhi), dbl_ = tc::desc64(wl + k8 * bstep, dhi);
            mma_f16(dcol, dal, dbh, idesc, (uint32_t)(c > 0) | (uint32_t)(k8 > 0));
            mma_f16(dcol, dah, dbl_, idesc, 1);
            mma_f16(dcol, dah, dbh, idesc, 1);
          }
          commit_a(B_EMPTY + 8 * s);
          if (u0 + KC >= kw) commit_a(B_BFULL + 8 * (n & 1));
        }
        __syncwarp();
      }
      kw = kw_next;
    }
  } else if (warp == 6) {
    // =============================== GEMM2 issuer ===============================
    int n = 0;
    for (int si = sb; si < se; ++si, ++n) {
      const int gi = dbl ? (n & 1) : 0;
      warp_wait_a(B_GFULL + 8 * gi, (uint32_t)((dbl ? (n >> 1) : n) & 1));
      warp_wait_a(B_LFULL + 8 * (n & 1), (uint32_t)((n >> 1) & 1));
      tc::fence_after_sync();
      if (elect_one()) {
        const uint32_t bq = tmem + TB0 + (uint32_t)((n & 1) * mp);
        const uint32_t gl = tmem + TGL0 + (uint32_t)(gi * mp);
        const uint32_t sc = tmem + TS0 + (uint32_t)(gi * mp);
        const uint32_t lh = tc::smem_desc_lo(sl0 + (uint32_t)((n & 1) * RW * RW) * 4u, lbo_n), ll = lh + (((uint32_t)mp * mw * 4) >> 4);
        for (int k8 = 0; k8 < (mw >> 3); ++k8) {
          const uint64_t bh = tc::desc64(lh + k8 * bstep, dhi), bl = tc::desc64(ll + k8 * bstep, dhi);
          mma_f16_ts(sc, gl + k8 * 8, bh, idesc, (uint32_t)(k8 > 0));
          mma_f16_ts(sc, bq + k8 * 8, bl, idesc, 1);
          mma_f16_ts(sc, bq + k8 * 8, bh, idesc, 1);
        }
        commit_a(B_S2 + 8 * (n & 1));
      }
      __syncwarp();
    }
  } else {
    // =============================== gate, accumulate, final rotation (thread = edge = TMEM lane) ===============================
    const int64_t el = (int64_t)tile * TILE + tid;
    const bool live = el < a.n_chunk;
    const int64_t e = a.e_lo + el;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const size_t g_bstride = (size_t)((a.n_chunk + TILE - 1) / TILE) * a.gstride * TILE;
    const float* grow = a.g + (size_t)tile * a.gstride * TILE + (live ? tid : 0);
    const float* sxrow = a.sx + (size_t)tile * a.n_blocks * TILE + tid;
    const float* __restrict__ img_inv = a.img_inv;
    float gv[RW], acc[RW];
#pragma unroll
    for (int j = 0; j < RW; ++j) { gv[j] = 0.f; acc[j] = 0.f; }
    float gA = 0.f, gB = 0.f;        // gate factor of the step whose values sit in gv: B *= gv * gA + gB (inverse X' and W scales folded in)
    float linv_next = 1.f;           // inverse L' scale of the step whose values sit in gv
    float ainv0 = 1.f, ainv1 = 1.f;  // accumulate factor of step n (slot n & 1): inverse row scale x inverse L' scale
    uint32_t cmask = 0;
    const uint4* steps4 = reinterpret_cast<const uint4*>(a.steps);
    auto load_gate = [&](const uint4& w0, const uint4& w1) {
      const int br = (int)(int8_t)(w1.y >> 24);
      const int blk = (int)(w1.z >> 16);
      const float sc = __uint_as_float(w1.x) * __ldg(img_inv + (w1.w & 0xffffu)) * __ldg(sxrow + (size_t)blk * TILE);
      linv_next = __ldg(img_inv + (w1.w >> 16));
      gA = (br < 0) ? 0.f : sc;
      gB = (br < 0) ? sc : 0.f;
      const float* gp = grow + (size_t)max(br, 0) * g_bstride + (size_t)((br < 0) ? 0 : (int)w0.w) * TILE;
#pragma unroll
      for (int j = 0; j < RW; ++j)
        if (j < mul) gv[j] = __ldg(gp + j * TILE);
    };
    auto accumulate = [&](int n, int flags, int m3) {
      const int gi = dbl ? (n & 1) : 0;
      const float f = (n & 1) ? ainv1 : ainv0;
      warp_wait_a(B_S2 + 8 * (n & 1), (uint32_t)((n >> 1) & 1));
      tc::fence_after_sync();
      const uint32_t sc = tmem + lane_base + TS0 + (uint32_t)(gi * mp);
#pragma unroll
      for (int c0 = 0; c0 < RW; c0 += 8) {
        if (c0 < mp) {
          uint32_t rs[8];
          tc::tmem_ld8(sc + c0, rs);
          tc::tmem_ld_wait8(rs);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[c0 + j] = fmaf(__uint_as_float(rs[j]), f, acc[c0 + j]);
        }
      }
      if (flags & 4) {
        const uint32_t cc = tmem + lane_base + TC + (uint32_t)(m3 * mul);
#pragma unroll
        for (int j = 0; j < RW; ++j) {
          if (j < mul) tmem_st1(cc + j, __float_as_uint(acc[j]));
          acc[j] = 0.f;
        }
        tc::tmem_st_wait();
      }
      tc::fence_before_sync();
    };
    auto gate = [&](int n) {
      const int gi = dbl ? (n & 1) : 0;
      const float fa = gA, fb = gB;
      warp_wait_a(B_BFULL + 8 * (n & 1), (uint32_t)((n >> 1) & 1));
      tc::fence_after_sync();
      const uint32_t bq = tmem + lane_base + TB0 + (uint32_t)((n & 1) * mp);
      const uint32_t gl = tmem + lane_base + TGL0 + (uint32_t)(gi * mp);
      // pass 1: row maximum of the gated product
      float tmax = 0.f;
#pragma unroll
      for (int c0 = 0; c0 < RW; c0 += 8) {
        if (c0 < mp) {
          uint32_t rb[8];
          tc::tmem_ld8(bq + c0, rb);
          tc::tmem_ld_wait8(rb);
#pragma unroll
          for (int j = 0; j < 8; ++j) tmax = fmaxf(tmax, fabsf(__uint_as_float(rb[j]) * fmaf(gv[c0 + j], fa, fb)));
        }
      }
      float s, inv;
      pow2_scale(tmax, 0, s, inv);
      if (n & 1) ainv1 = inv * linv_next; else ainv0 = inv * linv_next;
      // pass 2: scale, split, pack pairs of columns; the packed hi words overwrite columns [0, mp/2) of B (already read)
#pragma unroll
      for (int c0 = 0; c0 < RW; c0 += 16) {
        if (c0 < mp) {
          uint32_t ra[8], rb[8], hi[8], lo[8];
          tc::tmem_ld8(bq + c0, ra);
          tc::tmem_ld8(bq + c0 + 8, rb);
          tc::tmem_ld_wait8(ra);
          tc::tmem_ld_wait8(rb);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float v0 = __uint_as_float(ra[2 * j]) * (fmaf(gv[c0 + 2 * j], fa, fb) * s);
            const float v1 = __uint_as_float(ra[2 * j + 1]) * (fmaf(gv[c0 + 2 * j + 1], fa, fb) * s);
            const float v2 = __uint_as_float(rb[2 * j]) * (fmaf(gv[c0 + 8 + 2 * j], fa, fb) * s);
            const float v3 = __uint_as_float(rb[2 * j + 1]) * (fmaf(gv[c0 + 8 + 2 * j + 1], fa, fb) * s);
            if (a.swap_halves) {
              split_f16x2(v1, v0, hi[j], lo[j]);
              split_f16x2(v3, v2, hi[4 + j], lo[4 + j]);
            } else {
              split_f16x2(v0, v1, hi[j], lo[j]);
              split_f16x2(v2, v3, hi[4 + j], lo[4 + j]);
            }
          }
          tc::tmem_st8(bq + (c0 >> 1), hi);
          tc::tmem_st8(gl + (c0 >> 1), lo);
        }
      }
      tc::tmem_st_wait();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) arrive_a(B_GFULL + 8 * gi);
    };
    int n = 0, pflags = 0, pm3 = 0;
    uint32_t cur_fm = 0;
    if (se > sb) {
      const uint4 w0 = __ldg(steps4 + 2 * sb), w1 = __ldg(steps4 + 2 * sb + 1);
      cur_fm = w1.z;
      load_gate(w0, w1);
    }
    for (int si = sb; si < se; ++si, ++n) {
      uint4 n0 = make_uint4(0, 0, 0, 0), n1 = n0;
      const bool more = si + 1 < se;
      if (more) { n0 = __ldg(steps4 + 2 * (si + 1)); n1 = __ldg(steps4 + 2 * (si + 1) + 1); }
      const int m3 = (int)(cur_fm & 0xff), flags = (int)((cur_fm >> 8) & 0xff);
      cmask |= 1u << m3;
      if (!dbl && n > 0) accumulate(n - 1, pflags, pm3);
      gate(n);
      if (more) load_gate(n0, n1);
      if (dbl && n > 0) accumulate(n - 1, pflags, pm3);
      pflags = flags; pm3 = m3;
      cur_fm = n1.z;
    }
    if (n > 0) accumulate(n - 1, pflags, pm3);
    tc::fence_after_sync();
    {
      const int64_t orow = (live && a.out_index) ? a.out_index[e] : e;
      float* op = a.out + (live ? orow : 0) * a.plan.out_dim + ty.out_off;
      const float* Dz = a.dw + (live ? e : 0) * a.dstride + a.doff[ty.l];
      const uint32_t tc0 = tmem + lane_base + TC;
      const bool atomic = a.out_index != nullptr;
      switch (ty.l) {
        case 0: rot::rot_epilogue<0>(tc0, mul, cmask, Dz, op, live, atomic); break;
        case 1: rot::rot_epilogue<1>(tc0, mul, cmask, Dz, op, live, atomic); break;
        case 2: rot::rot_epilogue<2>(tc0, mul, cmask, Dz, op, live, atomic); break;
        case 3: rot::rot_epilogue<3>(tc0, mul, cmask, Dz, op, live, atomic); break;
        case 4: rot::rot_epilogue<4>(tc0, mul, cmask, Dz, op, live, atomic); break;
        case 5: rot::rot_epilogue<5>(tc0, mul, cmask, Dz, op, live, atomic); break;
        default: rot::rot_epilogue<6>(tc0, mul, cmask, Dz, op, live, atomic); break;
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc_dyn(tmem, ncols);
}

template <int RW, int NST>
constexpr size_t rot16_smem_bytes() { return (size_t)(NST * (2 * KC * TILE + 2 * RW * KC) + 2 * RW * RW) * sizeof(float); }

}  // namespace rot16
