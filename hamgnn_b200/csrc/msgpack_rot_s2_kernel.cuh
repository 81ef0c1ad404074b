// EXPERIMENTAL (opt-in with HGB_ROT_S2=1, not yet run on a GPU): variant of msgpack_rot_kernel for the slots with padded
// multiplicity 16 (48 % of the step, profiles/r01v_launches_m8.csv) that applies L' on the fp32 FMA pipes.
//
// Why: r01v shows the step of msgpack_rot_kernel bound by its sync points (bfull -> gate -> gfull -> GEMM2 -> s2done ->
// accumulate: ~2000 warp instructions at 1.55 IPC), not by arithmetic.  For mp = 16 the second GEMM is only 16 x 16 per
// row: 256 FFMA per thread and step.  Doing it in registers removes the hi/lo split of the gated product, both TMEM
// write-backs, the GEMM2 issuer warp and two of the three barrier hops; the step becomes
//     GEMM1  B[n&1] = X'_{m1} W_p     warp 4 (tcgen05 3xTF32, operands from the TMA ring filled by warp 5)
//     warps 0-3 (thread = edge = TMEM lane):  b = tcgen05.ld(B) -> release B -> b *= scale * g -> acc += b . L'   (fp32 FMA)
// with L' as a plain fp32 [16][16] tile in shared memory (broadcast 128-bit loads).  TMEM: B0 B1 | C' <= 128 columns, shared
// memory 74 KB  =>  3 CTAs per SM.  Arithmetic is fp32 throughout after GEMM1 (no 3xTF32 in the second contraction).
#pragma once

namespace rot {

constexpr int S2_MP = 16;
constexpr int S2_NST = 2;          // operand ring stages
constexpr int S2_NTHR = 192;       // 4 gate warps + GEMM1 issuer + TMA producer
constexpr int S2_STG = 2 * KC * TILE + 2 * S2_MP * KC;   // floats per ring stage

constexpr size_t rot_s2_smem_bytes() { return (size_t)(S2_NST * S2_STG + 2 * S2_MP * S2_MP) * sizeof(float); }

__global__ void __launch_bounds__(S2_NTHR, 3) msgpack_rot_s2_kernel(const __grid_constant__ RotArgs a) {
  constexpr int MP = S2_MP, NST = S2_NST, STG = S2_STG;
  extern __shared__ __align__(128) float smem[];
  // barriers: full[NST] | empty[NST] | lfull[2] | lfree[2] | bfull[2] | bfree[2]
  __shared__ uint64_t bars[2 * NST + 8];
  __shared__ uint32_t tmem_slot;
  const uint32_t bar0 = tc::smem_u32(bars);
  const uint32_t B_FULL = bar0, B_EMPTY = bar0 + 8 * NST, B_LFULL = bar0 + 16 * NST, B_LFREE = B_LFULL + 16, B_BFULL = B_LFULL + 32,
                 B_BFREE = B_LFULL + 48;
  const uint32_t stage0 = tc::smem_u32(smem);
  float* sL = smem + NST * STG;                                   // 2 x fp32 L' [16][16]
  const uint32_t sl0 = stage0 + (uint32_t)(NST * STG) * 4u;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x / a.n_slots;
  const int t = a.slot[blockIdx.x - tile * a.n_slots];
  const hgb_type_t ty = a.plan.types[t];
  const int d3 = 2 * ty.l + 1, mul = ty.mul;   // ty.mpad == 16 (host-checked)
  const int sb = a.step_begin[t], se = a.step_begin[t + 1];
  // TMEM columns: B0 | B1 | C' (d3 x mul, exact stride)
  const uint32_t TB0 = 0, TC = 2 * MP;
  uint32_t ncols = 32;
  while (ncols < TC + (uint32_t)(d3 * mul)) ncols <<= 1;

  if (tid == 0) {
    // lfree / bfree collect one arrival per gate warp, everything else one arrival (TMA transaction or tcgen05.commit)
    for (int i = 0; i < 2 * NST + 8; ++i) {
      const int k = i - 2 * NST;
      tc::mbar_init(&bars[i], (k == 2 || k == 3 || k == 6 || k == 7) ? 4 : 1);
    }
    tc::mbar_fence_init();
  }
  if (warp == 4) tmem_alloc_dyn(&tmem_slot, ncols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const float* __restrict__ wbuf = a.plan.wbuf;

  if (warp == 5) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      const float* xt = a.xp + (size_t)tile * a.tile_stride;
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
      int n = 0, c_all = 0;
      constexpr uint32_t lbytes = (uint32_t)(MP * MP) * 4u;
      for (int si = sb; si < se; ++si, ++n) {
        const hgb_rot_step_t st = a.steps[si];
        prefetch_gate(si + GPF);
        {
          const int lb = n & 1;
          if (n >= 2) wait_a(B_LFREE + 8 * lb, (uint32_t)(((n >> 1) - 1) & 1));   // the gate warps are done with L'(n-2)
          expect_tx_a(B_LFULL + 8 * lb, lbytes);
          bulk_g2s_a(sl0 + (uint32_t)(lb * MP * MP) * 4u, wbuf + st.pad2, lbytes, B_LFULL + 8 * lb);   // pad2: plain fp32 L'
        }
        const int kpad = st.kpad;
        for (int u0 = 0, c = 0; u0 < kpad; u0 += KC, ++c, ++c_all) {
          const int kc = min(KC, kpad - u0), s = c_all % NST;
          if (c_all >= NST) wait_a(B_EMPTY + 8 * s, (uint32_t)(((c_all / NST) - 1) & 1));
          const uint32_t sa = stage0 + (uint32_t)(s * STG) * 4u;
          const uint32_t ab = (uint32_t)(kc * TILE * 2) * 4u, wb = (uint32_t)(2 * MP * kc) * 4u;
          expect_tx_a(B_FULL + 8 * s, ab + wb);
          bulk_g2s_a(sa, xt + st.a_off + (size_t)c * (2 * KC * TILE), ab, B_FULL + 8 * s);
          bulk_g2s_a(sa + 2 * KC * TILE * 4, wbuf + st.w_off + (size_t)c * (2 * MP * KC), wb, B_FULL + 8 * s);
        }
      }
    }
    __syncwarp();
  } else if (warp == 4) {
    // =============================== GEMM1 issuer ===============================
    const uint32_t idesc = tc::idesc_tf32_m128(MP);
    const uint32_t dhi = tc::smem_desc_hi(128);
    constexpr uint32_t lbo_a = TILE * 16, lbo_n = (uint32_t)MP * 16;
    constexpr uint32_t astep = (2 * lbo_a) >> 4, bstep = (2 * lbo_n) >> 4;
    int n = 0, c_all = 0;
    int kpad = (sb < se) ? a.steps[sb].kpad : 0;
    for (int si = sb; si < se; ++si, ++n) {
      const int kpad_next = (si + 1 < se) ? a.steps[si + 1].kpad : 0;
      if (n >= 2) warp_wait_a(B_BFREE + 8 * (n & 1), (uint32_t)(((n >> 1) - 1) & 1));   // the gate warps have read B[n&1] of step n-2
      const uint32_t dcol = tmem + TB0 + (uint32_t)((n & 1) * MP);
      for (int u0 = 0, c = 0; u0 < kpad; u0 += KC, ++c, ++c_all) {
        const int kc = min(KC, kpad - u0), s = c_all % NST;
        warp_wait_a(B_FULL + 8 * s, (uint32_t)((c_all / NST) & 1));
        tc::fence_after_sync();
        if (elect_one()) {
          const uint32_t sa = stage0 + (uint32_t)(s * STG) * 4u;
          const uint32_t ah = tc::smem_desc_lo(sa, lbo_a), al = ah + (((uint32_t)kc * TILE * 4) >> 4);
          const uint32_t wh = tc::smem_desc_lo(sa + 2 * KC * TILE * 4, lbo_n), wl = wh + (((uint32_t)MP * kc * 4) >> 4);
          for (int k8 = 0; k8 < (kc >> 3); ++k8) {
            const uint64_t dah = tc::desc64(ah + k8 * astep, dhi), dal = tc::desc64(al + k8 * astep, dhi);
            const uint64_t dbh = tc::desc64(wh + k8 * bstep, dhi), dbl_ = tc::desc64(wl + k8 * bstep, dhi);
            tc::mma_tf32(dcol, dal, dbh, idesc, (uint32_t)(c > 0) | (uint32_t)(k8 > 0));
            tc::mma_tf32(dcol, dah, dbl_, idesc, 1);
            tc::mma_tf32(dcol, dah, dbh, idesc, 1);
          }
          commit_a(B_EMPTY + 8 * s);
          if (u0 + KC >= kpad) commit_a(B_BFULL + 8 * (n & 1));
        }
        __syncwarp();
      }
      kpad = kpad_next;
    }
  } else {
    // =============================== gate, L' and accumulation in registers, final rotation ===============================
    const int64_t el = (int64_t)tile * TILE + tid;
    const bool live = el < a.n_chunk;
    const int64_t e = a.e_lo + el;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const size_t g_bstride = (size_t)((a.n_chunk + TILE - 1) / TILE) * a.gstride * TILE;
    const float* grow = a.g + (size_t)tile * a.gstride * TILE + (live ? tid : 0);
    float gv[MP], acc[MP];
#pragma unroll
    for (int j = 0; j < MP; ++j) { gv[j] = 0.f; acc[j] = 0.f; }
    float gA = 0.f, gB = 0.f;
    uint32_t cmask = 0;
    const uint4* steps4 = reinterpret_cast<const uint4*>(a.steps);
    auto load_gate = [&](const uint4& w0, const uint4& w1) {
      const float sc = __uint_as_float(w1.x);
      const int br = (int)(int8_t)(w1.y >> 24);
      gA = (br < 0) ? 0.f : sc;
      gB = (br < 0) ? sc : 0.f;
      const float* gp = grow + (size_t)max(br, 0) * g_bstride + (size_t)((br < 0) ? 0 : (int)w0.w) * TILE;
#pragma unroll
      for (int j = 0; j < MP; ++j)
        if (j < mul) gv[j] = __ldg(gp + j * TILE);
    };
    int n = 0;
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
      // ---- B of this step -> registers, then B[n&1] belongs to GEMM1 of step n + 2 again
      warp_wait_a(B_BFULL + 8 * (n & 1), (uint32_t)((n >> 1) & 1));
      tc::fence_after_sync();
      uint32_t rb0[8], rb1[8];
      const uint32_t bq = tmem + lane_base + TB0 + (uint32_t)((n & 1) * MP);
      tc::tmem_ld8(bq, rb0);
      tc::tmem_ld8(bq + 8, rb1);
      tc::tmem_ld_wait8(rb0);
      tc::tmem_ld_wait8(rb1);
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) arrive_a(B_BFREE + 8 * (n & 1));
      float bg[MP];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        bg[j] = __uint_as_float(rb0[j]) * fmaf(gv[j], gA, gB);
        bg[8 + j] = __uint_as_float(rb1[j]) * fmaf(gv[8 + j], gA, gB);
      }
      if (more) load_gate(n0, n1);   // next step's gate values travel while L' is applied
      // ---- acc += bg . L'  (L' rows >= mul and bg columns >= mul are zero by construction of the images)
      wait_a(B_LFULL + 8 * (n & 1), (uint32_t)((n >> 1) & 1));   // every lane observes the TMA completion itself before its generic reads
      const float4* lp = reinterpret_cast<const float4*>(sL + (n & 1) * (MP * MP));
#pragma unroll
      for (int w = 0; w < MP; ++w) {
        const float4 l0 = lp[4 * w], l1 = lp[4 * w + 1], l2 = lp[4 * w + 2], l3 = lp[4 * w + 3];   // broadcast reads
        const float bw = bg[w];
        acc[0] = fmaf(bw, l0.x, acc[0]);   acc[1] = fmaf(bw, l0.y, acc[1]);   acc[2] = fmaf(bw, l0.z, acc[2]);   acc[3] = fmaf(bw, l0.w, acc[3]);
        acc[4] = fmaf(bw, l1.x, acc[4]);   acc[5] = fmaf(bw, l1.y, acc[5]);   acc[6] = fmaf(bw, l1.z, acc[6]);   acc[7] = fmaf(bw, l1.w, acc[7]);
        acc[8] = fmaf(bw, l2.x, acc[8]);   acc[9] = fmaf(bw, l2.y, acc[9]);   acc[10] = fmaf(bw, l2.z, acc[10]); acc[11] = fmaf(bw, l2.w, acc[11]);
        acc[12] = fmaf(bw, l3.x, acc[12]); acc[13] = fmaf(bw, l3.y, acc[13]); acc[14] = fmaf(bw, l3.z, acc[14]); acc[15] = fmaf(bw, l3.w, acc[15]);
      }
      __syncwarp();
      if (lane == 0) arrive_a(B_LFREE + 8 * (n & 1));
      if (flags & 4) {   // end of the m3 group: park the component in TMEM
        const uint32_t cc = tmem + lane_base + TC + (uint32_t)(m3 * mul);
#pragma unroll
        for (int j = 0; j < MP; ++j) {
          if (j < mul) tmem_st1(cc + j, __float_as_uint(acc[j]));
          acc[j] = 0.f;
        }
        tc::tmem_st_wait();
      }
      cur_fm = n1.z;
    }
    tc::fence_after_sync();
    {
      const int64_t orow = (live && a.out_index) ? a.out_index[e] : e;
      float* op = a.out + (live ? orow : 0) * a.plan.out_dim + ty.out_off;
      const float* Dz = a.dw + (live ? e : 0) * a.dstride + a.doff[ty.l];
      const uint32_t tc0 = tmem + lane_base + TC;
      const bool atomic = a.out_index != nullptr;
      switch (ty.l) {
        case 0: rot_epilogue<0>(tc0, mul, cmask, Dz, op, live, atomic); break;
        case 1: rot_epilogue<1>(tc0, mul, cmask, Dz, op, live, atomic); break;
        case 2: rot_epilogue<2>(tc0, mul, cmask, Dz, op, live, atomic); break;
        case 3: rot_epilogue<3>(tc0, mul, cmask, Dz, op, live, atomic); break;
        case 4: rot_epilogue<4>(tc0, mul, cmask, Dz, op, live, atomic); break;
        case 5: rot_epilogue<5>(tc0, mul, cmask, Dz, op, live, atomic); break;
        default: rot_epilogue<6>(tc0, mul, cmask, Dz, op, live, atomic); break;
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc_dyn(tmem, ncols);
}

}  // namespace rot
