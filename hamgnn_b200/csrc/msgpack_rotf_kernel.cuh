// Edge-aligned fused MessagePackBlock, variant with L' on the fp32 FMA pipes ("rotf").  Included by msgpack_tcg.cu after
// msgpack_rot_kernel.cuh: same step tables, same packed operands, same rotate-pack / radial-gate pre-passes, same epilogue.
//
// Why (profiles/README.md, r05): in msgpack_rot_kernel a step of a slot with multiplicity <= 32 takes 3 100 - 4 700 cycles per
// CTA although its tensor-core work is ~100 cycles: the gate / accumulate warps execute a ~350-instruction serial chain per step
// (wait B -> tcgen05.ld -> gate -> tf32 split -> 2 x tcgen05.st -> fence -> arrive -> [GEMM2: 3 MMAs per 8 channels] ->
// wait S -> tcgen05.ld -> add) at ~8 cycles per instruction, and everything that lengthens that chain (fp16 packing: +34 %,
// deeper gate prefetch with spills: +10 % per level) lengthens the kernel by the same factor.  For these slots the second
// contraction C'_{m3} += (B.g) L'_p is tiny (mul^2 <= 1 024 FMAs per edge and step), so it runs here on the FMA pipes straight
// from the registers that hold the gated product: no tf32 split, no write-back to TMEM, no GEMM2 issue / commit / wait, no S
// round trip.  L'_p is the un-split fp32 row-major [mpad][mpad] image (hgb_rot_step_t.pad2) in a 4-deep shared-memory ring,
// read as warp-broadcast float4.  With GEMM2 gone TMEM holds FOUR GEMM1 accumulators (B0..B3), so the GEMM1 warp runs up to
// three steps ahead of the gate warps.  fp32 accuracy: GEMM1 as before (3xTF32, chains <= 48 MMAs), everything after it is
// plain fp32 FMA (better than the split path).
//
// NWG = 1: warps 0-3 gate + L' + accumulate (thread = edge = TMEM lane), 4 GEMM1 issuer, 5 TMA A chunks, 6 TMA W chunks,
// 7 TMA L' images + gate L2 prefetch: 256 threads, 2 CTAs / SM (128 registers per thread).
// NWG = 2: TWO gate warpgroups (warps 0-3 take the even steps, warps 4-7 the odd steps; warp w works on TMEM lane quadrant
// w % 4), issuer / producers are warps 8-11: 384 threads, 2 CTAs / SM.  ncu of NWG = 1 (profiles/r05d): the gate warps execute
// ~290 instructions per step at ~8 cycles per instruction (two gate warps per scheduler: no latency hiding) and wait on their
// barriers only ~25 % of the time -- the step rate is the gate warps' own instruction stream, so the steps are dealt to twice
// as many warps.  Each warpgroup keeps its own partial sums and its own C' region; the epilogue adds the two regions.
#pragma once

namespace rotf {
using namespace tcmsg;
using rot::TILE;
using rot::KC;
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

// acc[0 .. 4 MC) += sum_{w < mul} p[w] L'[w][0 .. 4 MC): rows leave by a branch (the executed work follows the multiplicity),
// L' rows are warp-broadcast float4 loads from shared memory
template <int RW, int MC>
__device__ __forceinline__ void fma_rows(float (&acc)[RW], const float (&p)[RW], const float4* __restrict__ Ls, int rq, int mul) {
#pragma unroll
  for (int w = 0; w < RW; ++w) {
    if (w >= mul) break;   // warp-uniform
    const float pw = p[w];
#pragma unroll
    for (int q = 0; q < MC; ++q) {
      const float4 l = Ls[w * rq + q];
      acc[4 * q + 0] = fmaf(pw, l.x, acc[4 * q + 0]);
      acc[4 * q + 1] = fmaf(pw, l.y, acc[4 * q + 1]);
      acc[4 * q + 2] = fmaf(pw, l.z, acc[4 * q + 2]);
      acc[4 * q + 3] = fmaf(pw, l.w, acc[4 * q + 3]);
    }
  }
}

constexpr int NB = 4;    // GEMM1 accumulators in flight
constexpr int NLB = 4;   // L' buffers in flight

// acc[0 .. 4 MC) += sum_{j < nw} p[j] L'[j][0 .. 4 MC) with Ls already advanced to the first of the nw rows
template <int HW, int RW, int MC>
__device__ __forceinline__ void fma_rows_part(float (&acc)[RW], const float (&p)[HW], const float4* __restrict__ Ls, int rq, int nw) {
#pragma unroll
  for (int j = 0; j < HW; ++j) {
    if (j >= nw) break;   // warp-uniform
    const float pw = p[j];
#pragma unroll
    for (int q = 0; q < MC; ++q) {
      const float4 l = Ls[j * rq + q];
      acc[4 * q + 0] = fmaf(pw, l.x, acc[4 * q + 0]);
      acc[4 * q + 1] = fmaf(pw, l.y, acc[4 * q + 1]);
      acc[4 * q + 2] = fmaf(pw, l.z, acc[4 * q + 2]);
      acc[4 * q + 3] = fmaf(pw, l.w, acc[4 * q + 3]);
    }
  }
}

// C[z][w][k] = sum_m3 D^{l3}_z[m3][k] (C'_A + C'_B)[z][m3][w] for the channel quads c_first, c_first + c_step, ...; C'_X[m3][w] is
// TMEM column tcX + m3 * mul + w, bit m3 of cmX: region X holds that component (otherwise it counts as zero).
template <int L3>
__device__ __forceinline__ void rotf_epilogue(uint32_t tcA, uint32_t tcB, int mul, uint32_t cmA, uint32_t cmB, const float* __restrict__ Dz,
                                              float* __restrict__ op, bool live, bool atomic, int c_first, int c_step) {
  constexpr int d3 = 2 * L3 + 1;
  for (int c0 = c_first; c0 < mul; c0 += c_step) {   // warp-uniform
    float v[d3][4];
#pragma unroll
    for (int m = 0; m < d3; ++m) {
      uint32_t ca[4], cb[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        ca[j] = 0u; cb[j] = 0u;
        if (((cmA >> m) & 1u) && c0 + j < mul) rot::tmem_ld1(tcA + m * mul + c0 + j, ca[j]);
        if (((cmB >> m) & 1u) && c0 + j < mul) rot::tmem_ld1(tcB + m * mul + c0 + j, cb[j]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) { rot::tmem_ld_wait1(ca[j]); rot::tmem_ld_wait1(cb[j]); }
#pragma unroll
      for (int j = 0; j < 4; ++j) v[m][j] = __uint_as_float(ca[j]) + __uint_as_float(cb[j]);
    }
    if (!live) continue;
#pragma unroll
    for (int k = 0; k < d3; ++k) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int m = 0; m < d3; ++m) {
        const float dmk = (L3 == 0) ? 1.f : __ldg(Dz + m * d3 + k);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = fmaf(dmk, v[m][j], acc[j]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int w = c0 + j;
        if (w < mul) {
          if (atomic) atomicAdd(op + w * d3 + k, acc[j]);
          else op[w * d3 + k] = acc[j];
        }
      }
    }
  }
}

template <bool F16> struct ArgsOf { using type = rot::RotArgs; };
template <> struct ArgsOf<true> { using type = rot16::Rot16Args; };

// F16 = true: GEMM1 on fp16 x 2 split operands (the rot16 packing: rotate_pack16_kernel images with a power-of-two scale per
// (edge, input block), W images with a scale per image, kind::f16 MMAs with K = 16 per instruction, offsets / kpad in 32-bit
// words) -- half the tensor-core instructions and half the operand bytes of the tf32 form; the inverse scales are folded
// into the gate factor.  The step's lf_off is then the offset of the un-split fp32 L' image in plan.wbuf.
// SPLIT = true (NWG = 2, slot class 32): both gate warpgroups work on EVERY step, each on half of the mid channels w (rows of
// L'): 16 gate values, 16 gated products and 32 partial sums per thread instead of 32 / 32 / 32 -- the class-32 form of the
// kernel with one warpgroup spilled at 128 registers and ran 1.7x slower than msgpack_rot_kernel<32,2>.  At the end of an m3
// group warpgroup 0 parks its partial sums in C', warpgroup 1 adds its own (one named barrier per group).
template <int RW, int NST, int NWG, int NMW, bool F16, bool SPLIT = false>
__global__ void __launch_bounds__(128 * NWG + 128, 2) msgpack_rotf_kernel(const __grid_constant__ typename ArgsOf<F16>::type a) {
  static_assert(!SPLIT || (NWG == 2 && RW == 32), "SPLIT: two gate warpgroups, slot class 32");
  // NMW = GEMM1 issuer warps; the CTA always has four non-gate warps: NMW = 1: issuer | A | W | L' + gate prefetch,
  // NMW = 2: issuer 0 | issuer 1 | A + gate prefetch | W + L'
  constexpr int W_MMA = 4 * NWG, W_A = W_MMA + NMW, W_W = W_A + 1, W_L = (NMW == 2) ? W_W : W_A + 2;
  constexpr int STG = 2 * KC * TILE + 2 * RW * KC;   // floats per ring stage: A chunk (hi | lo) + W chunk (hi | lo)
  constexpr int NBAR = 2 * NST + 2 * NLB + 2 * NB;
  extern __shared__ __align__(128) float smem[];
  // barriers: full[NST] | empty[NST] | lfull[NLB] | lfree[NLB] | bfull[NB] | bfree[NB]
  __shared__ uint64_t bars[NBAR];
  __shared__ uint32_t tmem_slot;
  __shared__ uint32_t cm_sh[2];
  __shared__ volatile int full_turn;   // NMW == 2: chunks whose ring-stage "full" wait has been passed, in chunk order
  const uint32_t bar0 = tc::smem_u32(bars);
  const uint32_t B_FULL = bar0, B_EMPTY = bar0 + 8 * NST, B_LFULL = bar0 + 16 * NST, B_LFREE = B_LFULL + 8 * NLB,
                 B_BFULL = B_LFREE + 8 * NLB, B_BFREE = B_BFULL + 8 * NB;
  const uint32_t stage0 = tc::smem_u32(smem);
  float* const lsm = smem + NST * STG;                           // NLB x RW^2 floats
  const uint32_t sl0 = stage0 + (uint32_t)(NST * STG) * 4u;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x / a.n_slots;
  const int t = a.slot[blockIdx.x - tile * a.n_slots];
  const hgb_type_t ty = a.plan.types[t];
  const int d3 = 2 * ty.l + 1, mp = ty.mpad, mul = ty.mul;
  const int sb = a.step_begin[t], se = a.step_begin[t + 1];
  // TMEM columns: B0 | B1 | B2 | B3 | C' of warpgroup 0 (d3 x mul, exact stride) (| C' of warpgroup 1)
  const uint32_t TC = (uint32_t)(NB * mp);
  uint32_t ncols = 32;
  while (ncols < TC + (uint32_t)((SPLIT ? 1 : NWG) * d3 * mul)) ncols <<= 1;

  if (tid == 0) {
    full_turn = 0;
    for (int i = 0; i < NBAR; ++i) {
      const bool four = (i >= 2 * NST + NLB && i < 2 * NST + 2 * NLB) || i >= 2 * NST + 2 * NLB + NB;   // lfree, bfree: one arrival per gate warp
      tc::mbar_init(&bars[i], four ? (SPLIT ? 8 : 4) : (i < NST ? 2 : 1));                                           // full: A + W producers
    }
    tc::mbar_fence_init();
  }
  if (warp == W_MMA) tmem_alloc_dyn(&tmem_slot, ncols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const float* __restrict__ wbuf = a.plan.wbuf;   // un-split fp32 L' images (and the tf32 W images)
  const float* __restrict__ wimg = wbuf;          // GEMM1 W images
  if constexpr (F16) wimg = a.wbuf16;
  const uint32_t idesc = F16 ? rot16::idesc_f16_m128(mp) : tc::idesc_tf32_m128(mp);
  const uint32_t dhi = tc::smem_desc_hi(128);
  const uint32_t lbo_a = TILE * 16, lbo_n = (uint32_t)mp * 16;
  const uint32_t astep = (2 * lbo_a) >> 4, bstep = (2 * lbo_n) >> 4;

  if (warp >= W_A) {
    // =============================== TMA producers ===============================
    if (lane == 0) {
      const bool doA = warp == W_A, doW = warp == W_W, doL = warp == W_L, doG = (NMW == 2) ? doA : doL;
      const float* xt = reinterpret_cast<const float*>(a.xp) + (size_t)tile * a.tile_stride;
      constexpr int GPF = 4;
      const size_t g_bstride = (size_t)((a.n_chunk + TILE - 1) / TILE) * a.gstride * TILE;
      const float* gt = a.g + (size_t)tile * a.gstride * TILE;
      // The step records are read one iteration ahead (and the record of the gate block to prefetch one more): a dependent
      // global load at the top of every iteration would put an L2 round trip (~700 cycles) into this thread's per-step time
      const uint4* rec4 = reinterpret_cast<const uint4*>(a.steps);
      auto gate_block = [&](const uint4& w0, const uint4& w1) {   // L2 prefetch of the gate block of a step
        const int br = (int)(int8_t)(w1.y >> 24);
        if (br >= 0) bulk_prefetch_l2(gt + (size_t)br * g_bstride + (size_t)(int)w0.w * TILE, (uint32_t)(mul * TILE) * 4u);
      };
      if (doG)
        for (int j = 0; j < GPF && sb + j < se; ++j) gate_block(__ldg(rec4 + 2 * (sb + j)), __ldg(rec4 + 2 * (sb + j) + 1));
      const uint32_t lbytes = (uint32_t)(mp * mp) * 4u;
      int n = 0, c_all = 0;
      uint4 c0 = make_uint4(0, 0, 0, 0), c1 = c0, g0 = c0, g1 = c0;
      if (sb < se) { c0 = __ldg(rec4 + 2 * sb); c1 = __ldg(rec4 + 2 * sb + 1); }
      if (doG && sb + GPF < se) { g0 = __ldg(rec4 + 2 * (sb + GPF)); g1 = __ldg(rec4 + 2 * (sb + GPF) + 1); }
      for (int si = sb; si < se; ++si, ++n) {
        uint4 n0 = c0, n1 = c1, h0 = g0, h1 = g1;
        if (si + 1 < se) { n0 = __ldg(rec4 + 2 * (si + 1)); n1 = __ldg(rec4 + 2 * (si + 1) + 1); }
        if (doG && si + 1 + GPF < se) { h0 = __ldg(rec4 + 2 * (si + 1 + GPF)); h1 = __ldg(rec4 + 2 * (si + 1 + GPF) + 1); }
        if (doG && si + GPF < se) gate_block(g0, g1);
        const int st_a_off = (int)c0.x, st_w_off = (int)c0.y, st_lf_off = (int)c0.z, st_kpad = (int)(int16_t)(c1.y & 0xffffu), st_pad2 = (int)c1.w;
        if (doA || doW) {
          const int kpad = st_kpad;
          for (int u0 = 0, c = 0; u0 < kpad; u0 += KC, ++c, ++c_all) {
            const int kc = min(KC, kpad - u0), s = c_all % NST;
            if (c_all >= NST) wait_a(B_EMPTY + 8 * s, (uint32_t)(((c_all / NST) - 1) & 1));
            const uint32_t sa = stage0 + (uint32_t)(s * STG) * 4u;
            const uint32_t ab = (uint32_t)(kc * TILE * 2) * 4u, wb = (uint32_t)(2 * mp * kc) * 4u;
            if (doA) {
              expect_tx_a(B_FULL + 8 * s, ab);
              bulk_g2s_a(sa, xt + st_a_off + (size_t)c * (2 * KC * TILE), ab, B_FULL + 8 * s);
            } else {
              expect_tx_a(B_FULL + 8 * s, wb);
              bulk_g2s_a(sa + 2 * KC * TILE * 4, wimg + st_w_off + (size_t)c * (2 * mp * KC), wb, B_FULL + 8 * s);
            }
          }
        }
        if (doL) {
          const int lb = n % NLB;
          if (n >= NLB) wait_a(B_LFREE + 8 * lb, (uint32_t)(((n / NLB) - 1) & 1));   // the gate warps have read L'(n - NLB)
          expect_tx_a(B_LFULL + 8 * lb, lbytes);
          bulk_g2s_a(sl0 + (uint32_t)(lb * RW * RW) * 4u, wbuf + (F16 ? st_lf_off : st_pad2), lbytes, B_LFULL + 8 * lb);
        }
        c0 = n0; c1 = n1; g0 = h0; g1 = h1;
      }
    }
    __syncwarp();
  } else if (warp >= W_MMA) {
    // =============================== GEMM1 issuer(s) ===============================
    // NWG = 2: two issuer warps take alternate steps (ncu r05f: the single issuer's ~190 scalar instructions per step at ~8 cycles
    // each paced the CTA once the gate work was dealt to two warpgroups); both walk all steps to keep the ring's chunk count
    const int mw = warp - W_MMA;
    int n = 0, c_all = 0;
    int kpad = (sb < se) ? a.steps[sb].kpad : 0;
    for (int si = sb; si < se; ++si, ++n) {
      const int kpad_next = (si + 1 < se) ? a.steps[si + 1].kpad : 0;
      if (NMW == 2 && (n & 1) != mw) {   // the other issuer's step
        c_all += (kpad + KC - 1) / KC;
        kpad = kpad_next;
        continue;
      }
      const int b = n % NB;
      if (n >= NB) warp_wait_a(B_BFREE + 8 * b, (uint32_t)(((n / NB) - 1) & 1));   // the gate warps have read B(n - NB)
      const uint32_t dcol = tmem + (uint32_t)(b * mp);
      for (int u0 = 0, c = 0; u0 < kpad; u0 += KC, ++c, ++c_all) {
        const int kc = min(KC, kpad - u0), s = c_all % NST;
        if (NMW == 2) {
          // a parity wait is only valid one phase ahead: the two issuers pass the ring's "full" waits strictly in chunk order
          // (without this the issuer of step n + 1 could test a stage two uses early and see the phase of use u - 2)
          if (lane == 0) {
            uint32_t spins = 0;
            while (full_turn != c_all)
              if (++spins > 0x4000000u) __trap();
            wait_a(B_FULL + 8 * s, (uint32_t)((c_all / NST) & 1));
            full_turn = c_all + 1;
          }
          __syncwarp();
        } else {
          warp_wait_a(B_FULL + 8 * s, (uint32_t)((c_all / NST) & 1));
        }
        tc::fence_after_sync();
        if (elect_one()) {
          const uint32_t sa = stage0 + (uint32_t)(s * STG) * 4u;
          const uint32_t ah = tc::smem_desc_lo(sa, lbo_a), al = ah + (((uint32_t)kc * TILE * 4) >> 4);
          const uint32_t wh = tc::smem_desc_lo(sa + 2 * KC * TILE * 4, lbo_n), wl = wh + (((uint32_t)mp * kc * 4) >> 4);
          for (int k8 = 0; k8 < (kc >> 3); ++k8) {
            const uint64_t dah = tc::desc64(ah + k8 * astep, dhi), dal = tc::desc64(al + k8 * astep, dhi);
            const uint64_t dbh = tc::desc64(wh + k8 * bstep, dhi), dbl_ = tc::desc64(wl + k8 * bstep, dhi);
            if constexpr (F16) {
              rot16::mma_f16(dcol, dal, dbh, idesc, (uint32_t)(c > 0) | (uint32_t)(k8 > 0));
              rot16::mma_f16(dcol, dah, dbl_, idesc, 1);
              rot16::mma_f16(dcol, dah, dbh, idesc, 1);
            } else {
              tc::mma_tf32(dcol, dal, dbh, idesc, (uint32_t)(c > 0) | (uint32_t)(k8 > 0));
              tc::mma_tf32(dcol, dah, dbl_, idesc, 1);
              tc::mma_tf32(dcol, dah, dbh, idesc, 1);
            }
          }
          commit_a(B_EMPTY + 8 * s);
          if (u0 + KC >= kpad) commit_a(B_BFULL + 8 * b);
        }
        __syncwarp();
      }
      kpad = kpad_next;
    }
  } else {
    if constexpr (SPLIT) {
      // =============================== gate halves (both warpgroups on every step) ===============================
      constexpr int HW = RW / 2;
      const int wg = warp >> 2, zt = (warp & 3) * 32 + lane;
      const int64_t el = (int64_t)tile * TILE + zt;
      const bool live = el < a.n_chunk;
      const int64_t e = a.e_lo + el;
      const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
      const size_t g_bstride = (size_t)((a.n_chunk + TILE - 1) / TILE) * a.gstride * TILE;
      const float* grow = a.g + (size_t)tile * a.gstride * TILE + (live ? zt : 0);
      const int mc = (mul + 3) >> 2;
      const int h0 = min(mul, ((mul >> 1) + 3) & ~3);        // mid channels [0, h0) -> warpgroup 0, [h0, mul) -> warpgroup 1
      const int w0 = wg ? h0 : 0, nw = wg ? mul - h0 : h0;   // this warpgroup's rows of L' (nw <= 16)
      const int nq = (nw + 3) >> 2;
      const int nsteps = se - sb;
      float gv[HW], acc[RW];
#pragma unroll
      for (int j = 0; j < HW; ++j) gv[j] = 0.f;
#pragma unroll
      for (int j = 0; j < RW; ++j) acc[j] = 0.f;
      float gA = 0.f, gB = 0.f;
      uint32_t cmask = 0;
      const uint4* steps4 = reinterpret_cast<const uint4*>(a.steps) + 2 * sb;
      auto load_gate = [&](const uint4& r0, const uint4& r1) {
        float sc = __uint_as_float(r1.x);
        const int br = (int)(int8_t)(r1.y >> 24);
        gA = (br < 0) ? 0.f : sc;
        gB = (br < 0) ? sc : 0.f;
        if (br >= 0) {   // warp-uniform
          const float* gp = grow + (size_t)br * g_bstride + (size_t)((int)r0.w + w0) * TILE;
#pragma unroll
          for (int q = 0; q < HW / 4; ++q) {
            if (q >= nq) break;
#pragma unroll
            for (int j = 4 * q; j < 4 * q + 4; ++j) gv[j] = __ldg(gp + j * TILE);   // columns past the multiplicity meet B == 0 / unused rows
          }
        }
      };
      uint32_t cur_fm = 0;
      if (nsteps > 0) {
        const uint4 r0 = __ldg(steps4), r1 = __ldg(steps4 + 1);
        cur_fm = r1.z;
        load_gate(r0, r1);
      }
      for (int n = 0; n < nsteps; ++n) {
        uint4 n0 = make_uint4(0, 0, 0, 0), n1 = n0;
        const bool more = n + 1 < nsteps;
        if (more) { n0 = __ldg(steps4 + 2 * (n + 1)); n1 = __ldg(steps4 + 2 * (n + 1) + 1); }
        const int m3 = (int)(cur_fm & 0xff);
        const bool park = ((cur_fm >> 8) & 4u) != 0;
        cmask |= 1u << m3;
        const int b = n % NB, lb = n % NLB;
        warp_wait_a(B_BFULL + 8 * b, (uint32_t)((n / NB) & 1));
        tc::fence_after_sync();
        const uint32_t bq = tmem + lane_base + (uint32_t)(b * mp + w0);
        float p[HW];
#pragma unroll
        for (int c0 = 0; c0 < HW; c0 += 8) {
          uint32_t rb[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) rb[j] = 0u;
          if (c0 < nw) {   // warp-uniform; w0 + c0 + 7 <= 31 < mp
            tc::tmem_ld8(bq + c0, rb);
            tc::tmem_ld_wait8(rb);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) p[c0 + j] = __uint_as_float(rb[j]) * fmaf(gv[c0 + j], gA, gB);
        }
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) arrive_a(B_BFREE + 8 * b);
        if (more) load_gate(n0, n1);
        wait_a(B_LFULL + 8 * lb, (uint32_t)((n / NLB) & 1));
        const int rq = mp >> 2;
        const float4* Ls = reinterpret_cast<const float4*>(lsm + lb * RW * RW) + w0 * rq;
        if (mc <= 6) fma_rows_part<HW, RW, 6>(acc, p, Ls, rq, nw);
        else fma_rows_part<HW, RW, RW / 4>(acc, p, Ls, rq, nw);
        __syncwarp();
        if (lane == 0) arrive_a(B_LFREE + 8 * lb);
        if (park) {   // warp-uniform, the same decision in both warpgroups
          const uint32_t cc = tmem + lane_base + TC + (uint32_t)(m3 * mul);
          if (wg == 0) {
#pragma unroll
            for (int j = 0; j < RW; ++j)
              if (j < mul) tmem_st1(cc + j, __float_as_uint(acc[j]));
            tc::tmem_st_wait();
            tc::fence_before_sync();
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (wg == 1) {
            tc::fence_after_sync();
#pragma unroll
            for (int j = 0; j < RW; ++j) {
              if (j < mul) {
                uint32_t c = 0u;
                rot::tmem_ld1(cc + j, c);
                rot::tmem_ld_wait1(c);
                tmem_st1(cc + j, __float_as_uint(acc[j] + __uint_as_float(c)));
              }
            }
            tc::tmem_st_wait();
          }
#pragma unroll
          for (int j = 0; j < RW; ++j) acc[j] = 0.f;
        }
        cur_fm = n1.z;
      }
      tc::fence_before_sync();
      asm volatile("bar.sync 1, 256;" ::: "memory");   // warpgroup 1 has added its last partial sums
      tc::fence_after_sync();
      {
        const int64_t orow = (live && a.out_index) ? a.out_index[e] : e;
        float* op = a.out + (live ? orow : 0) * a.plan.out_dim + ty.out_off;
        const float* Dz = a.dw + (live ? e : 0) * a.dstride + a.doff[ty.l];
        const uint32_t tcA = tmem + lane_base + TC;
        const bool atomic = a.out_index != nullptr;
        const int cf = 4 * wg, cs = 8;
        switch (ty.l) {
          case 0: rotf_epilogue<0>(tcA, tcA, mul, cmask, 0u, Dz, op, live, atomic, cf, cs); break;
          case 1: rotf_epilogue<1>(tcA, tcA, mul, cmask, 0u, Dz, op, live, atomic, cf, cs); break;
          case 2: rotf_epilogue<2>(tcA, tcA, mul, cmask, 0u, Dz, op, live, atomic, cf, cs); break;
          case 3: rotf_epilogue<3>(tcA, tcA, mul, cmask, 0u, Dz, op, live, atomic, cf, cs); break;
          case 4: rotf_epilogue<4>(tcA, tcA, mul, cmask, 0u, Dz, op, live, atomic, cf, cs); break;
          case 5: rotf_epilogue<5>(tcA, tcA, mul, cmask, 0u, Dz, op, live, atomic, cf, cs); break;
          default: rotf_epilogue<6>(tcA, tcA, mul, cmask, 0u, Dz, op, live, atomic, cf, cs); break;
        }
      }
    } else {
    // =============================== gate, L', accumulate, final rotation (thread = edge = TMEM lane) ===============================
    const int wg = warp >> 2, zt = (warp & 3) * 32 + lane;   // warpgroup (its steps: n = wg, wg + NWG, ...), edge inside the tile
    const int64_t el = (int64_t)tile * TILE + zt;
    const bool live = el < a.n_chunk;
    const int64_t e = a.e_lo + el;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const size_t g_bstride = (size_t)((a.n_chunk + TILE - 1) / TILE) * a.gstride * TILE;
    const float* grow = a.g + (size_t)tile * a.gstride * TILE + (live ? zt : 0);
    const int mc = (mul + 3) >> 2;   // output quads that carry data
    const int nsteps = se - sb;
    const uint32_t tcw = tmem + lane_base + TC + (uint32_t)(wg * d3 * mul);   // this warpgroup's C'
    float gv[RW], acc[RW];
#pragma unroll
    for (int j = 0; j < RW; ++j) { gv[j] = 0.f; acc[j] = 0.f; }
    float gA = 0.f, gB = 0.f;
    uint32_t cmask = 0;
    const uint4* steps4 = reinterpret_cast<const uint4*>(a.steps) + 2 * sb;
    // gate values of a step straight into registers (quads past the multiplicity are skipped by a branch); un-gated steps
    // (branch < 0) use the factor (gA, gB) = (0, scale) instead of (scale, 0) on whatever the registers hold
    auto load_gate = [&](const uint4& w0, const uint4& w1) {
      float sc = __uint_as_float(w1.x);
      const int br = (int)(int8_t)(w1.y >> 24);
      if constexpr (F16)   // inverse scales of the W image and of this edge's rows of the input block
        sc *= __ldg(a.img_inv + (w1.w & 0xffffu)) * __ldg(a.sx + ((size_t)tile * a.n_blocks + (w1.z >> 16)) * TILE + zt);
      gA = (br < 0) ? 0.f : sc;
      gB = (br < 0) ? sc : 0.f;
      if (br >= 0) {   // warp-uniform
        const float* gp = grow + (size_t)br * g_bstride + (size_t)(int)w0.w * TILE;
#pragma unroll
        for (int q = 0; q < RW / 4; ++q) {
          if (q >= mc) break;
#pragma unroll
          for (int j = 4 * q; j < 4 * q + 4; ++j) gv[j] = __ldg(gp + j * TILE);   // columns in [mul, 4 mc) exist (gstride % 4 == 0) and meet B == 0
        }
      }
    };
    uint32_t cur_fm = 0;
    if (wg < nsteps) {
      const uint4 w0 = __ldg(steps4 + 2 * wg), w1 = __ldg(steps4 + 2 * wg + 1);
      cur_fm = w1.z;
      load_gate(w0, w1);
    }
    for (int n = wg; n < nsteps; n += NWG) {
      uint4 n0 = make_uint4(0, 0, 0, 0), n1 = n0;
      const bool more = n + NWG < nsteps;
      if (more) { n0 = __ldg(steps4 + 2 * (n + NWG)); n1 = __ldg(steps4 + 2 * (n + NWG) + 1); }
      uint32_t fm_other = 0;   // NWG == 2: the step in between belongs to the other warpgroup; it may close this m3 group
      if (NWG == 2 && n + 1 < nsteps) fm_other = __ldg(reinterpret_cast<const uint32_t*>(steps4 + 2 * (n + 1) + 1) + 2);
      const int m3 = (int)(cur_fm & 0xff);
      const bool park = (((cur_fm | fm_other) >> 8) & 4u) != 0;
      cmask |= 1u << m3;
      const int b = n % NB, lb = n % NLB;
      // ---- B(n) -> registers, gated
      warp_wait_a(B_BFULL + 8 * b, (uint32_t)((n / NB) & 1));
      tc::fence_after_sync();
      const uint32_t bq = tmem + lane_base + (uint32_t)(b * mp);
      float p[RW];
#pragma unroll
      for (int c0 = 0; c0 < RW; c0 += 8) {
        uint32_t rb[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) rb[j] = 0u;
        if (c0 < mul) {   // warp-uniform
          tc::tmem_ld8(bq + c0, rb);
          tc::tmem_ld_wait8(rb);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) p[c0 + j] = __uint_as_float(rb[j]) * fmaf(gv[c0 + j], gA, gB);
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) arrive_a(B_BFREE + 8 * b);   // GEMM1(n + NB) may overwrite the accumulator
      if (more) load_gate(n0, n1);                // the gate values of this warpgroup's next step travel while the FMA block runs
      // ---- acc[w'] += sum_w p[w] L'[w][w']
      wait_a(B_LFULL + 8 * lb, (uint32_t)((n / NLB) & 1));   // every lane observes the phase: it reads the TMA-written bytes itself
      const float4* Ls = reinterpret_cast<const float4*>(lsm + lb * RW * RW);
      const int rq = mp >> 2;   // float4 per L' row
      if (RW == 16) {
        switch (mc) {
          case 1: fma_rows<RW, 1>(acc, p, Ls, rq, mul); break;
          case 2: fma_rows<RW, 2>(acc, p, Ls, rq, mul); break;
          case 3: fma_rows<RW, 3>(acc, p, Ls, rq, mul); break;
          default: fma_rows<RW, 4>(acc, p, Ls, rq, mul); break;
        }
      } else {
        if (mc <= 6) fma_rows<RW, (RW >= 24 ? 6 : RW / 4)>(acc, p, Ls, rq, mul);
        else fma_rows<RW, RW / 4>(acc, p, Ls, rq, mul);
      }
      __syncwarp();
      if (lane == 0) arrive_a(B_LFREE + 8 * lb);
      // ---- this warpgroup's last step of an m3 group: the registers go to its C'[m3]
      if (park) {
        const uint32_t cc = tcw + (uint32_t)(m3 * mul);
#pragma unroll
        for (int j = 0; j < RW; ++j) {
          if (j < mul) tmem_st1(cc + j, __float_as_uint(acc[j]));   // warp-uniform predicate
          acc[j] = 0.f;
        }
        tc::tmem_st_wait();
      }
      cur_fm = n1.z;
    }
    tc::fence_before_sync();
    uint32_t cmA = cmask, cmB = 0;
    if (NWG == 2) {
      if ((warp & 3) == 0 && lane == 0) cm_sh[wg] = cmask;
      asm volatile("bar.sync 1, 256;" ::: "memory");   // both warpgroups have parked their sums
      cmA = cm_sh[0]; cmB = cm_sh[1];
    }
    tc::fence_after_sync();
    {
      const int64_t orow = (live && a.out_index) ? a.out_index[e] : e;
      float* op = a.out + (live ? orow : 0) * a.plan.out_dim + ty.out_off;
      const float* Dz = a.dw + (live ? e : 0) * a.dstride + a.doff[ty.l];
      const uint32_t tcA = tmem + lane_base + TC, tcB = tcA + (uint32_t)(d3 * mul);
      const bool atomic = a.out_index != nullptr;
      const int cf = 4 * wg, cs = 4 * NWG;   // the warpgroups share the channel quads of the final rotation
      switch (ty.l) {
        case 0: rotf_epilogue<0>(tcA, tcB, mul, cmA, cmB, Dz, op, live, atomic, cf, cs); break;
        case 1: rotf_epilogue<1>(tcA, tcB, mul, cmA, cmB, Dz, op, live, atomic, cf, cs); break;
        case 2: rotf_epilogue<2>(tcA, tcB, mul, cmA, cmB, Dz, op, live, atomic, cf, cs); break;
        case 3: rotf_epilogue<3>(tcA, tcB, mul, cmA, cmB, Dz, op, live, atomic, cf, cs); break;
        case 4: rotf_epilogue<4>(tcA, tcB, mul, cmA, cmB, Dz, op, live, atomic, cf, cs); break;
        case 5: rotf_epilogue<5>(tcA, tcB, mul, cmA, cmB, Dz, op, live, atomic, cf, cs); break;
        default: rotf_epilogue<6>(tcA, tcB, mul, cmA, cmB, Dz, op, live, atomic, cf, cs); break;
      }
    }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc_dyn(tmem, ncols);
}

template <int RW, int NST>
constexpr size_t rotf_smem_bytes() { return (size_t)(NST * (2 * KC * TILE + 2 * RW * KC) + NLB * RW * RW) * sizeof(float); }

}  // namespace rotf
