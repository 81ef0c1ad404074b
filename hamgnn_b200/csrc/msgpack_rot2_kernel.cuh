// A-stationary edge-aligned fused MessagePackBlock ("rot2").  Included by msgpack_tcg.cu after msgpack_rot_kernel.cuh
// (same Wigner / rotate-pack / radial-gate pre-passes, same packed X' images).
//
// msgpack_rot_kernel evaluates one (path, m1) step at a time: 2 800 steps per tile of 128 edges and message, every
// step a GEMM1 with N = one padded multiplicity (16 on the slots that hold half of the time), its own copy of the
// input image from L2 (53 MB of A traffic per tile and message) and a gate -> GEMM2 -> accumulate hand-off
// (profiles/r01v_rot_msgpack_ncu_summary.md: tensor pipe 7 %, 3 000 cycles per step, bound by the hand-offs).
// Here the steps are regrouped on the host (plan.MessagePackOp._build_rot2_program):
//
//   CTA   = (tile of 128 edges, pass); pass = (output component m3, output slots with <= 128 accumulator columns)
//   piece = one image X'_{block, m1} x the CONCATENATED weights of every path of the pass that reads it
//             GEMM1  B[n&1][128 x ncols] = X' [W_p1 | W_p2 | ...]          warp 13, A / W chunks of 16 channels from the TMA ring (warp 12)
//             gate   b = B * g_p (8-column batches, thread = edge)          warps 0-7, two per TMEM lane quadrant, each walks its own
//                                                                           host-built stream; the w3j scale is folded into W.
//                      multiplicity > 16: B <- hi(b), GL[n&1] <- lo        -> GEMM2 on the tensor core
//                      multiplicity <= 16: s += b L' on the fp32 FMA pipes, C' += s  (no GEMM2: a tcgen05.mma costs ~55 cycles
//                                          whatever its N, profiles/r02g_mma_probe.txt -- 57 % of the GEMM2 instructions had N = 16)
//             GEMM2  S[n&1][s_off ..] = (B.g)[:, col0 : col0 + kcols] L'stack   warp 14, one K-concatenated chain per destination slot
//             acc    C'[acc_col0 + w] += S[s_off + w]                        warps 8-11, fp32 round-to-nearest in shared memory
//   end   : the pass's columns of the aligned-frame row cp[e][.] are stored (coalesced), unrotate_kernel finishes.
//
// ~700 pieces instead of ~2 800 steps, mean N 62 instead of 16-32, every image fetched once per pass: 15 MB of A per tile and
// message instead of 53 MB.  The tensor core's truncating accumulation still only spans one piece (<= 16 + 36 MMAs); the
// sum over pieces is fp32 round-to-nearest (see msgpack_rot_kernel.cuh on why).
// TMEM (512 columns, 1 CTA / SM): B0 B1 | GL0 GL1 (96 each) | S0 S1 (64 each) -- everything double buffered, so a gate
// warp only ever waits for GEMM1 and the GEMM2 issuer only for the gate.
#pragma once

namespace rot2 {
using namespace tcmsg;
using rot::arrive_a;
using rot::bulk_g2s_a;
using rot::bulk_prefetch_l2;
using rot::commit_a;
using rot::elect_one;
using rot::expect_tx_a;
using rot::tmem_alloc_dyn;
using rot::tmem_dealloc_dyn;

// Barrier wait that SLEEPS: try_wait with a suspend-time hint parks the thread in hardware until the phase completes (or
// the hint expires), instead of re-issuing the poll every few cycles.  With 15 warps per CTA of which ~10 are waiting at any
// time, the polling loops of msgpack_rot_kernel's wait_a took more than half of all issued instructions and starved the
// gate warps of issue slots (profiles/r02p).  Bounded: ~4 s, then trap.
__device__ __forceinline__ void wait_a(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .u32 n;\n\t"
      "mov.u32 n, 0;\n\t"
      "mov.u32 %0, 1;\n"
      "HGB_R2_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x4000;\n\t"
      "@p bra HGB_R2_DONE;\n\t"
      "add.u32 n, n, 1;\n\t"
      "setp.lt.u32 p, n, 0x400000;\n\t"
      "@p bra HGB_R2_WAIT;\n\t"
      "mov.u32 %0, 0;\n"
      "HGB_R2_DONE:\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(addr), "r"(parity)
      : "memory");
  if (!ok) __trap();   // a pipeline bug must surface as a kernel error, never as a hung GPU
}
__device__ __forceinline__ void warp_wait_a(uint32_t addr, uint32_t parity) {   // one lane waits, the warp re-converges
  if ((threadIdx.x & 31) == 0) wait_a(addr, parity);
  __syncwarp();
}

constexpr int TILE = 128;
constexpr int KC32 = 32;   // channel chunk of the packed X' images (rot::KC)
constexpr int KC2 = 16;    // channels per ring stage (host: MessagePackOp.R2_KC)
constexpr int NB = 96;     // B columns per piece (R2_NB)
constexpr int SW = 64;     // S columns per piece (R2_SW)
constexpr int NST = 3;
constexpr int ACC_COLS = 128, ACC_LD = 129;
constexpr int A_LO = KC2 * TILE;                       // float offset of the A lo image inside a stage
constexpr int W_AT = 2 * KC2 * TILE;                   // float offset of the W chunk inside a stage
constexpr int STG = 2 * KC2 * TILE + 2 * KC2 * NB;     // floats per ring stage (28 KB)
constexpr int LBUF = 8192;                             // floats per L' buffer (R2_LMAX_FLOATS)
constexpr int NH = HGB_ROT2_GATE_GROUPS;               // gate-warp groups (each = 4 warps, one per TMEM lane quadrant)
constexpr int NGW = 4 * NH;                            // gate warps
// A cp.async.bulk costs its issuing THREAD ~400 cycles whatever its size, and the cost does not add up across warps
// (profiles/r02v_tma_probe.txt): with one producer thread and ~14 bulk operations per piece the producer paced the whole kernel
// (~6 000 cycles per piece, profiles/r02r trace).  The three operands of a stage are therefore issued by three warps, and the
// L2 prefetches of later pieces by the (otherwise mostly idle) accumulate warps.
#ifndef HGB_ROT2_PRODUCERS
#define HGB_ROT2_PRODUCERS 1
#endif
constexpr int NPROD = HGB_ROT2_PRODUCERS;              // 1: one warp issues A hi, A lo, W and L'; 3: one warp each (W warp also L')
constexpr int NTHR = 32 * (NGW + 6 + NPROD);           // warps 0..NGW-1 gate, then 4 accumulate, NPROD TMA, GEMM1, GEMM2
constexpr int W_ACC = NGW, W_TMA = NGW + 4, W_MMA1 = NGW + 4 + NPROD, W_MMA2 = NGW + 5 + NPROD;
constexpr uint32_t TB = 0, TGL = 2 * NB, TS = 4 * NB;  // TMEM columns
constexpr size_t SMEM_BYTES = (size_t)(NST * STG + 2 * LBUF + ACC_COLS * ACC_LD) * sizeof(float);

struct Args {
  const float* wbuf;
  const hgb_rot2_pass_t* passes;
  const hgb_rot2_piece_t* pieces;
  const hgb_rot2_batch_t* batches;
  const hgb_rot2_dst_t* dsts;
  const hgb_rot2_gpf_t* gpf;
  int n_passes;
  const float* xp;      // packed rotated inputs of the chunk [tile][tile_stride]
  int tile_stride;
  const float* g;       // radial gate of the chunk [tile][branch][gstride][128]
  int gtile_floats;     // n_branches * gstride * 128
  int64_t e_lo, n_chunk;
  float* cp;            // aligned-frame messages of ALL edges [E][rowstride]
  int rowstride;
  long long* trace;     // optional (HGB_ROT2_TRACE): clock64 stamps of CTA trace_cta, [role][piece][2]
  int trace_cta, trace_pieces;
  int flags;            // experiment switches (HGB_ROT2_FLAGS), unused at present
};

// role: 0 TMA, 1 GEMM1, 2 gate (warp 0), 3 GEMM2, 4 accumulate (warp 8)
#define R2_TRACE(role, n, k)                                                                                         \
  do {                                                                                                               \
    if (a.trace != nullptr && (int)blockIdx.x == a.trace_cta && (threadIdx.x & 31) == 0 && (n) < a.trace_pieces)     \
      a.trace[((role) * a.trace_pieces + (n)) * 2 + (k)] = clock64();                                                \
  } while (0)

struct PieceRec {   // hgb_rot2_piece_t as two 16-byte words
  int a_off, w_off, l_off, l_floats, gpf_begin, dst_begin, kpad, ncols, ndst, gpf_n;
};
__device__ __forceinline__ PieceRec load_piece(const hgb_rot2_piece_t* p) {
  const uint4 w0 = __ldg(reinterpret_cast<const uint4*>(p)), w1 = __ldg(reinterpret_cast<const uint4*>(p) + 1);
  PieceRec r;
  r.a_off = (int)w0.x; r.w_off = (int)w0.y; r.l_off = (int)w0.z; r.l_floats = (int)w0.w;
  r.gpf_begin = (int)w1.x; r.dst_begin = (int)w1.y;
  r.kpad = (int)(w1.z & 0xffffu); r.ncols = (int)(w1.z >> 16); r.ndst = (int)(w1.w & 0xffffu); r.gpf_n = (int)(w1.w >> 16);
  return r;
}

__global__ void __launch_bounds__(NTHR, 1) msgpack_rot2_kernel(const __grid_constant__ Args a) {
  extern __shared__ __align__(128) float smem[];
  // barriers: full[3] | empty[3] | lfull[2] | bfull[2] | gfull[2] | s2done[2] | sfree[2] | simtdone
  __shared__ uint64_t bars[17];
  __shared__ uint32_t tmem_slot;
  const uint32_t bar0 = tc::smem_u32(bars);
  const uint32_t B_FULL = bar0, B_EMPTY = bar0 + 8 * NST, B_LFULL = bar0 + 16 * NST, B_BFULL = B_LFULL + 16, B_GFULL = B_LFULL + 32,
                 B_S2 = B_LFULL + 48, B_SFREE = B_LFULL + 64, B_SIMT = B_LFULL + 80;
  const uint32_t stage0 = tc::smem_u32(smem);
  const uint32_t lbuf0 = stage0 + (uint32_t)(NST * STG) * 4u;
  float* accs = smem + NST * STG + 2 * LBUF;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x / a.n_passes;
  const int pass = blockIdx.x - tile * a.n_passes;
  const hgb_rot2_pass_t ps = a.passes[pass];

  if (tid == 0) {
    for (int i = 0; i < 17; ++i) tc::mbar_init(&bars[i], i < NST ? 3 : (((i >= 10 && i < 12) || i == 16) ? NGW : (i >= 14 ? 4 : 1)));   // gfull / simtdone / sfree: one arrival per warp
    tc::mbar_fence_init();
  }
  if (warp == W_MMA1) tmem_alloc_dyn(&tmem_slot, 512);
  for (int i = tid; i < ps.ncols * ACC_LD; i += NTHR) accs[i] = 0.f;   // C' of the pass
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const float* __restrict__ wbuf = a.wbuf;
  const uint32_t dhi = tc::smem_desc_hi(128);

  if (warp >= W_TMA && warp < W_TMA + NPROD) {
    // =============================== TMA producers: A hi | A lo | W and L' ===============================
    const int role = (NPROD == 1) ? -1 : warp - W_TMA;   // -1: every operand
    if (lane == 0) {
      const float* xt = a.xp + (size_t)tile * a.tile_stride;
      int n = 0, c_all = 0;
      for (int qi = ps.piece_begin; qi < ps.piece_end; ++qi, ++n) {
        const PieceRec pc = load_piece(a.pieces + qi);
        if (role <= 0) R2_TRACE(0, n, 0);
        for (int u0 = 0, c = 0; u0 < pc.kpad; u0 += KC2, ++c, ++c_all) {
          const int kc = min(KC2, pc.kpad - u0), s = c_all % NST;
          if (c_all >= NST) wait_a(B_EMPTY + 8 * s, (uint32_t)(((c_all / NST) - 1) & 1));
          const uint32_t sa = stage0 + (uint32_t)(s * STG) * 4u;
          const uint32_t ab = (uint32_t)(kc * TILE) * 4u, wb = (uint32_t)(2 * kc * pc.ncols) * 4u;
          // the packed image is chunked by 32 channels, (hi | lo) per chunk: this stage is half of such a chunk
          const int c32 = c >> 1, kc32 = min(KC32, pc.kpad - c32 * KC32);
          const float* ahi = xt + pc.a_off + (size_t)c32 * (2 * KC32 * TILE) + (size_t)(c & 1) * (KC2 * TILE);
          if (role <= 0) {
            expect_tx_a(B_FULL + 8 * s, ab);
            bulk_g2s_a(sa, ahi, ab, B_FULL + 8 * s);
          }
          if (role < 0 || role == 1) {
            expect_tx_a(B_FULL + 8 * s, ab);
            bulk_g2s_a(sa + A_LO * 4, ahi + (size_t)kc32 * TILE, ab, B_FULL + 8 * s);
          }
          if (role < 0 || role == 2) {
            expect_tx_a(B_FULL + 8 * s, wb);
            bulk_g2s_a(sa + W_AT * 4, wbuf + pc.w_off + (size_t)c * (2 * KC2 * pc.ncols), wb, B_FULL + 8 * s);
          }
        }
        if (role < 0 || role == 2) {
          // the L' operands are needed last (GEMM2 / FMA-pipe batches of this piece)
          const int lb = n & 1;
          if (n >= 2) wait_a(B_S2 + 8 * lb, (uint32_t)(((n >> 1) - 1) & 1));   // GEMM2(n-2) has read this L' buffer
          const uint32_t lbytes = (uint32_t)pc.l_floats * 4u;
          expect_tx_a(B_LFULL + 8 * lb, lbytes);
          bulk_g2s_a(lbuf0 + (uint32_t)(lb * LBUF) * 4u, wbuf + pc.l_off, lbytes, B_LFULL + 8 * lb);
        }
        if (role <= 0) R2_TRACE(0, n, 1);
      }
    }
    __syncwarp();
  } else if (warp == W_MMA1) {
    // =============================== GEMM1 issuer ===============================
    const uint32_t lbo_a = TILE * 16, astep = (2 * lbo_a) >> 4;
    int n = 0, c_all = 0;
    for (int qi = ps.piece_begin; qi < ps.piece_end; ++qi, ++n) {
      const PieceRec pc = load_piece(a.pieces + qi);
      if (n >= 2) warp_wait_a(B_S2 + 8 * (n & 1), (uint32_t)(((n >> 1) - 1) & 1));   // GEMM2(n-2) has read B / GL [n&1]
      R2_TRACE(1, n, 0);
      const uint32_t dcol = tmem + TB + (uint32_t)((n & 1) * NB);
      const uint32_t idesc = tc::idesc_tf32_m128(pc.ncols);
      const uint32_t lbo_n = (uint32_t)pc.ncols * 16, bstep = (2 * lbo_n) >> 4;
      for (int u0 = 0, c = 0; u0 < pc.kpad; u0 += KC2, ++c, ++c_all) {
        const int kc = min(KC2, pc.kpad - u0), s = c_all % NST;
        warp_wait_a(B_FULL + 8 * s, (uint32_t)((c_all / NST) & 1));
        tc::fence_after_sync();
        if (elect_one()) {
          const uint32_t sa = stage0 + (uint32_t)(s * STG) * 4u;
          const uint32_t ah = tc::smem_desc_lo(sa, lbo_a), al = tc::smem_desc_lo(sa + A_LO * 4, lbo_a);
          const uint32_t wh = tc::smem_desc_lo(sa + W_AT * 4, lbo_n), wl = wh + (((uint32_t)pc.ncols * kc * 4) >> 4);
          for (int k8 = 0; k8 < (kc >> 3); ++k8) {
            const uint64_t dah = tc::desc64(ah + k8 * astep, dhi), dal = tc::desc64(al + k8 * astep, dhi);
            const uint64_t dbh = tc::desc64(wh + k8 * bstep, dhi), dbl_ = tc::desc64(wl + k8 * bstep, dhi);
            tc::mma_tf32(dcol, dal, dbh, idesc, (uint32_t)(c > 0) | (uint32_t)(k8 > 0));
            tc::mma_tf32(dcol, dah, dbl_, idesc, 1);
            tc::mma_tf32(dcol, dah, dbh, idesc, 1);
          }
          commit_a(B_EMPTY + 8 * s);
          if (u0 + KC2 >= pc.kpad) commit_a(B_BFULL + 8 * (n & 1));
        }
        __syncwarp();
      }
      R2_TRACE(1, n, 1);
    }
  } else if (warp == W_MMA2) {
    // =============================== GEMM2 issuer ===============================
    int n = 0;
    for (int qi = ps.piece_begin; qi < ps.piece_end; ++qi, ++n) {
      const uint4 w1 = __ldg(reinterpret_cast<const uint4*>(a.pieces + qi) + 1);
      const int dst_begin = (int)w1.y, ndst = (int)(w1.w & 0xffffu);
      uint4 drec = make_uint4(0, 0, 0, 0);   // lane d holds destination group d of the piece
      if (lane < ndst) drec = __ldg(reinterpret_cast<const uint4*>(a.dsts + dst_begin + lane));
      const int nb = n & 1;
      const uint32_t par = (uint32_t)((n >> 1) & 1);
      if (n >= 2) warp_wait_a(B_SFREE + 8 * nb, par ^ 1u);   // the accumulate warps have drained S[n&1] of piece n-2
      warp_wait_a(B_LFULL + 8 * nb, par);
      warp_wait_a(B_GFULL + 8 * nb, par);
      tc::fence_after_sync();
      R2_TRACE(3, n, 0);
      const uint32_t bq = tmem + TB + (uint32_t)(nb * NB), gl = tmem + TGL + (uint32_t)(nb * NB), sc0 = tmem + TS + (uint32_t)(nb * SW);
      const uint32_t lb = lbuf0 + (uint32_t)(nb * LBUF) * 4u;
      for (int di = 0; di < ndst; ++di) {
        const uint32_t dx = __shfl_sync(0xffffffffu, drec.x, di), dy = __shfl_sync(0xffffffffu, drec.y, di),
                       dwv = __shfl_sync(0xffffffffu, drec.w, di);
        if (elect_one()) {
          const uint32_t col0 = dx & 0xffffu, kcols = dx >> 16, mp = dy & 0xffffu, s_off = dy >> 16, l_rel = dwv;
          const uint32_t idesc = tc::idesc_tf32_m128((int)mp);
          const uint32_t lbo_l = mp * 16, lstep = (2 * lbo_l) >> 4;
          const uint32_t lh = tc::smem_desc_lo(lb + l_rel * 4u, lbo_l), ll = lh + ((kcols * mp * 4u) >> 4);
          const uint32_t sc = sc0 + s_off;
          for (uint32_t k8 = 0; k8 < (kcols >> 3); ++k8) {
            const uint64_t bh = tc::desc64(lh + k8 * lstep, dhi), bl = tc::desc64(ll + k8 * lstep, dhi);
            tc::mma_tf32_ts(sc, gl + col0 + k8 * 8, bh, idesc, (uint32_t)(k8 > 0));
            tc::mma_tf32_ts(sc, bq + col0 + k8 * 8, bl, idesc, 1);
            tc::mma_tf32_ts(sc, bq + col0 + k8 * 8, bh, idesc, 1);
          }
          if (di == ndst - 1) commit_a(B_S2 + 8 * nb);
        }
        __syncwarp();
      }
      if (ndst == 0) {   // every destination of the piece was applied on the FMA pipes: nothing to wait for, B / GL / L' are free
        if (elect_one()) commit_a(B_S2 + 8 * nb);
        __syncwarp();
      }
      R2_TRACE(3, n, 1);
    }
  } else if (warp < NGW) {
    // =============================== gate (thread = edge = TMEM lane) ===============================
    // Two warps per lane quadrant; half h walks its own stream of 8-column batches (host: FMA-pipe slots belong to one
    // half, tensor batches alternate).  Two-stage software pipeline: the gate values of batch k + 1 are in flight while
    // batch k is processed (the body exists twice -- no register shuffling, ~10 KB of code).
    const int q = warp & 3, h = warp >> 2;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int zt = q * 32 + lane;
    const bool live = (int64_t)tile * TILE + zt < a.n_chunk;
    // gate tensor of the chunk: [tile][branch][gstride][128]; a batch record holds the float offset of its two 4-column blocks
    const float* gz = a.g + (size_t)tile * a.gtile_floats + (live ? zt : 0);
    float* acc = accs + zt;
    const float* lbase = smem + NST * STG;
    const int sb = ps.stream_begin[h], se = ps.stream_end[h];
    const uint4* recs = reinterpret_cast<const uint4*>(a.batches);
    auto load_gates = [&](const uint4& r, float (&gq)[8]) {
      if (r.y == 0xFFFFFFFFu) {
        gq[0] = gq[1] = gq[2] = gq[3] = 1.f;
      } else {
        const float* gp = gz + r.y;
#pragma unroll
        for (int j = 0; j < 4; ++j) gq[j] = __ldg(gp + j * TILE);   // a warp reads 128 contiguous bytes per column
      }
      if (r.z == 0xFFFFFFFFu) {
        gq[4] = gq[5] = gq[6] = gq[7] = 1.f;
      } else {
        const float* gp = gz + r.z;
#pragma unroll
        for (int j = 0; j < 4; ++j) gq[4 + j] = __ldg(gp + j * TILE);
      }
    };
    float s[16];
#pragma unroll
    for (int w = 0; w < 16; ++w) s[w] = 0.f;
    int n = 0, tb = 0;
    auto process = [&](const uint4& r, const float (&gq)[8]) {
      const uint32_t meta = r.x;
      const uint32_t kind = meta & 3u;
      if (meta & 4u) {   // first batch of this half in piece n
        warp_wait_a(B_BFULL + 8 * (n & 1), (uint32_t)((n >> 1) & 1));
        warp_wait_a(B_LFULL + 8 * (n & 1), (uint32_t)((n >> 1) & 1));   // FMA-pipe batches read the L' rows of the piece
        tc::fence_after_sync();
        if (warp == 0) R2_TRACE(2, n, 0);
      }
      const bool tr = a.trace != nullptr && (int)blockIdx.x == a.trace_cta && warp == 0 && lane == 0 && tb < 256;
      if (tr) a.trace[10 * a.trace_pieces + tb * 4 + 0] = clock64();
      if (kind != 2u) {
        const uint32_t col = (meta >> 5) & 0x7F8u;   // (meta >> 8 & 0xFF) * 8
        const uint32_t bq = tmem + lane_base + TB + (uint32_t)((n & 1) * NB) + col;
        uint32_t rb[8];
        tc::tmem_ld8(bq, rb);
        tc::tmem_ld_wait8(rb);
        if (tr) a.trace[10 * a.trace_pieces + tb * 4 + 1] = clock64();
        float bg[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) bg[j] = __uint_as_float(rb[j]) * gq[j];
        if (tr) { float sum = 0.f; for (int j = 0; j < 8; ++j) sum += bg[j]; if (sum == 123.456f) a.trace[0] = 0; a.trace[10 * a.trace_pieces + tb * 4 + 2] = clock64(); }
        if (kind == 0u) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float hh, ll;
            tc::split_tf32(bg[j], hh, ll);
            hi[j] = __float_as_uint(hh); lo[j] = __float_as_uint(ll);
          }
          tc::tmem_st8(bq, hi);
          tc::tmem_st8(tmem + lane_base + TGL + (uint32_t)((n & 1) * NB) + col, lo);
        } else {
          // s[w'] += sum_j b_j L'[j][w'] : L' rows broadcast from shared memory (padding rows are zero, b = 0 there)
          const int mul = (int)((meta >> 16) & 0x1Fu), m4 = (mul + 3) & ~3;
          if (meta & 16u) {
#pragma unroll
            for (int w = 0; w < 16; ++w) s[w] = 0.f;
          }
          const float* lr = lbase + (n & 1) * LBUF + r.w;
#pragma unroll
          for (int w4 = 0; w4 < 4; ++w4) {
            if (4 * w4 < m4) {   // warp-uniform
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 lv = *reinterpret_cast<const float4*>(lr + j * m4 + 4 * w4);
                s[4 * w4 + 0] = fmaf(bg[j], lv.x, s[4 * w4 + 0]);
                s[4 * w4 + 1] = fmaf(bg[j], lv.y, s[4 * w4 + 1]);
                s[4 * w4 + 2] = fmaf(bg[j], lv.z, s[4 * w4 + 2]);
                s[4 * w4 + 3] = fmaf(bg[j], lv.w, s[4 * w4 + 3]);
              }
            }
          }
          if (meta & 32u) {   // last batch of the destination group: C' += s (this warp owns these columns)
            float* ap = acc + ((meta >> 21) & 0xFFu) * ACC_LD;
#pragma unroll
            for (int w = 0; w < 16; ++w)
              if (w < mul) ap[w * ACC_LD] += s[w];
          }
        }
      }
      if (tr) a.trace[10 * a.trace_pieces + tb * 4 + 3] = clock64() | ((long long)kind << 60);
      ++tb;
      if (meta & 8u) {   // last batch of this half in piece n: hand the piece to GEMM2
        tc::tmem_st_wait();
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) arrive_a(B_GFULL + 8 * (n & 1));
        if (warp == 0) R2_TRACE(2, n, 1);
        ++n;
      }
    };
    // records travel four batches ahead (a record load used right away cost 10 % of all stall samples, profiles/r02p),
    // gate values one batch ahead
    const uint4 rdummy = make_uint4(2u, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u);
    uint4 r0 = rdummy, r1 = rdummy, r2 = rdummy, r3 = rdummy;
    float g0[8], g1[8];
    if (sb < se) r0 = __ldg(recs + sb);
    if (sb + 1 < se) r1 = __ldg(recs + sb + 1);
    if (sb + 2 < se) r2 = __ldg(recs + sb + 2);
    if (sb + 3 < se) r3 = __ldg(recs + sb + 3);
    if (sb < se) load_gates(r0, g0);
#pragma unroll 1
    for (int bb = sb; bb < se; bb += 2) {
      uint4 r4 = rdummy, r5 = rdummy;
      if (bb + 4 < se) r4 = __ldg(recs + bb + 4);
      if (bb + 5 < se) r5 = __ldg(recs + bb + 5);
      if (bb + 1 < se) load_gates(r1, g1);
      process(r0, g0);
      if (bb + 1 < se) {
        if (bb + 2 < se) load_gates(r2, g0);
        process(r1, g1);
      }
      r0 = r2; r1 = r3; r2 = r4; r3 = r5;
    }
    __syncwarp();
    if (lane == 0) arrive_a(B_SIMT);   // every FMA-pipe contribution of this warp is in C'
  } else {
    // =============================== accumulate: C' += S, finally store the pass's columns of cp ===============================
    const int q = warp - W_ACC, zl = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    float* acc = accs + zl;
    // L2 prefetch of the operands of piece qi (the ring holds about one piece; the packed inputs and the gate tensor of a
    // chunk are GBs, so the first CTA of a tile to touch a block reads it from DRAM): one operand per accumulate warp
    const float* xt = a.xp + (size_t)tile * a.tile_stride;
    const float* gtile = a.g + (size_t)tile * a.gtile_floats;
    auto prefetch_piece = [&](int qi) {
      if (qi >= ps.piece_end) return;
      const PieceRec pp = load_piece(a.pieces + qi);
      if (q == 0 && lane == 0) bulk_prefetch_l2(xt + pp.a_off, (uint32_t)(2 * pp.kpad * TILE) * 4u);
      if (q == 1 && lane == 0) bulk_prefetch_l2(wbuf + pp.w_off, (uint32_t)(2 * pp.kpad * pp.ncols) * 4u);
      if (q == 2 && lane == 0) bulk_prefetch_l2(wbuf + pp.l_off, (uint32_t)pp.l_floats * 4u);
      if (q == 3 && lane < pp.gpf_n) {
        const uint2 gr = __ldg(reinterpret_cast<const uint2*>(a.gpf + pp.gpf_begin + lane));
        bulk_prefetch_l2(gtile + gr.x, gr.y);
      }
    };
    prefetch_piece(ps.piece_begin + 1);
    prefetch_piece(ps.piece_begin + 2);
    prefetch_piece(ps.piece_begin + 3);
    int n = 0;
    for (int qi = ps.piece_begin; qi < ps.piece_end; ++qi, ++n) {
      const uint4 w1 = __ldg(reinterpret_cast<const uint4*>(a.pieces + qi) + 1);
      const int dst_begin = (int)w1.y, ndst = (int)(w1.w & 0xffffu);
      uint4 drec = make_uint4(0, 0, 0, 0);   // lane d holds destination group d of the piece
      if (lane < ndst) drec = __ldg(reinterpret_cast<const uint4*>(a.dsts + dst_begin + lane));
      prefetch_piece(qi + 4);
      warp_wait_a(B_S2 + 8 * (n & 1), (uint32_t)((n >> 1) & 1));
      tc::fence_after_sync();
      if (q == 0) R2_TRACE(4, n, 0);
      const uint32_t sc0 = tmem + lane_base + TS + (uint32_t)((n & 1) * SW);
      for (int di = 0; di < ndst; ++di) {
        const uint32_t dy = __shfl_sync(0xffffffffu, drec.y, di), dz = __shfl_sync(0xffffffffu, drec.z, di);
        const int s_off = (int)(dy >> 16), acc_col0 = (int)(dz & 0xffffu), mul = (int)(dz >> 16);
        float* ap = acc + acc_col0 * ACC_LD;
        for (int c0 = 0; c0 < mul; c0 += 32) {   // <= 32 columns in flight per wait
          uint32_t rs[4][8];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (c0 + 8 * i < mul) tc::tmem_ld8(sc0 + (uint32_t)(s_off + c0 + 8 * i), rs[i]);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (c0 + 8 * i < mul) {
              tc::tmem_ld_wait8(rs[i]);
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (c0 + 8 * i + j < mul) ap[(c0 + 8 * i + j) * ACC_LD] += __uint_as_float(rs[i][j]);
            }
          }
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) arrive_a(B_SFREE + 8 * (n & 1));
      if (q == 0) R2_TRACE(4, n, 1);
    }
    warp_wait_a(B_SIMT, 0);   // the gate warps have added the last FMA-pipe contributions
    __syncwarp();
    // warp q owns edges [32 q, 32 q + 32) of the tile: rows of cp, the pass's columns are contiguous
    for (int zz = 0; zz < 32; ++zz) {
      const int64_t el = (int64_t)tile * TILE + q * 32 + zz;
      if (el >= a.n_chunk) break;
      float* row = a.cp + (size_t)(a.e_lo + el) * a.rowstride + ps.out_col0;
      const float* as = accs + q * 32 + zz;
      for (int c = lane; c < ps.ncols; c += 32) row[c] = as[c * ACC_LD];
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == W_MMA1) tmem_dealloc_dyn(tmem, 512);
}

// ===================================================================================================== unrotate (+ segment sum)
struct UnrotArgs {
  const float* cp;
  int rowstride;
  const float* dw;
  int dstride;
  int doff[12];
  int n_warps;                     // warps per sweep = blockDim.x / 32
  int n_items;                     // (slot, 32-channel group) items; item i is handled by warp i % n_warps in sweep i / n_warps
  int item_slot[64], item_w0[64];
  int slot_l[32], slot_mul[32], slot_out_off[32];
  int ccol[32][13];
  const int64_t* seg_ptr;          // NULL: row r = edge r
  const int64_t* seg_order;
  int64_t n_rows;
  float* out;
  int out_dim;
};

__device__ __forceinline__ void unrot_bar() { asm volatile("bar.sync 1;" ::: "memory"); }   // same barrier from every instantiation

// warp = (slot t, 32 channels); thread: o[k] = sum over the segment's edges of sum_m3 D^{l3}_e[m3][k] C'_e[t][m3][w], summed in
// list order (deterministic).  Every warp of the CTA walks the same segment and meets the others at unrot_bar() twice per
// edge (D^l of the edge staged once in shared memory); warps differ in L3, lanes of a warp do not.
template <int L3>
__device__ __forceinline__ void unrot_run(const UnrotArgs& a, float (*sD)[480], int t, int w, bool active, int64_t j0, int64_t j1,
                                          int64_t row) {
  constexpr int d3 = 2 * L3 + 1;
  float o[d3];
#pragma unroll
  for (int k = 0; k < d3; ++k) o[k] = 0.f;
  int cc[d3];
#pragma unroll
  for (int m = 0; m < d3; ++m) cc[m] = active ? a.ccol[t][m] : -1;
  const int dof = a.doff[L3];
  int buf = 0;
  for (int64_t j = j0; j < j1; ++j, buf ^= 1) {
    const int64_t e = a.seg_order ? a.seg_order[j] : j;
    const float* drow = a.dw + (size_t)e * a.dstride;
    for (int i = threadIdx.x; i < a.dstride; i += blockDim.x) sD[buf][i] = __ldg(drow + i);
    float c[d3];
    const float* crow = a.cp + (size_t)e * a.rowstride + w;
#pragma unroll
    for (int m = 0; m < d3; ++m) c[m] = (cc[m] >= 0) ? __ldg(crow + cc[m]) : 0.f;
    unrot_bar();   // sD[buf] complete; sD[buf ^ 1] was last read before the previous barrier
    if (L3 == 0) {
      o[0] += c[0];
    } else {
      const float* D = sD[buf] + dof;
#pragma unroll
      for (int m = 0; m < d3; ++m)
#pragma unroll
        for (int k = 0; k < d3; ++k) o[k] = fmaf(D[m * d3 + k], c[m], o[k]);
    }
  }
  if (active) {
    float* op = a.out + (size_t)row * a.out_dim + a.slot_out_off[t] + w * d3;
#pragma unroll
    for (int k = 0; k < d3; ++k) op[k] = o[k];
  }
}

__global__ void __launch_bounds__(1024) unrotate_kernel(const __grid_constant__ UnrotArgs a) {
  __shared__ float sD[2][480];
  const int64_t row = blockIdx.x;
  const int64_t j0 = a.seg_ptr ? a.seg_ptr[row] : row, j1 = a.seg_ptr ? a.seg_ptr[row + 1] : row + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i0 = 0; i0 < a.n_items; i0 += a.n_warps) {
    const int it = i0 + warp;
    const bool have = it < a.n_items;
    const int t = have ? a.item_slot[it] : 0;
    const int w = have ? a.item_w0[it] + lane : 0;
    const bool active = have && w < a.slot_mul[t];
    const int l3 = have ? a.slot_l[t] : 0;
    switch (l3) {   // warp-uniform
      case 0: unrot_run<0>(a, sD, t, active ? w : 0, active, j0, j1, row); break;
      case 1: unrot_run<1>(a, sD, t, active ? w : 0, active, j0, j1, row); break;
      case 2: unrot_run<2>(a, sD, t, active ? w : 0, active, j0, j1, row); break;
      case 3: unrot_run<3>(a, sD, t, active ? w : 0, active, j0, j1, row); break;
      case 4: unrot_run<4>(a, sD, t, active ? w : 0, active, j0, j1, row); break;
      case 5: unrot_run<5>(a, sD, t, active ? w : 0, active, j0, j1, row); break;
      default: unrot_run<6>(a, sD, t, active ? w : 0, active, j0, j1, row); break;
    }
    unrot_bar();
  }
}

}  // namespace rot2
