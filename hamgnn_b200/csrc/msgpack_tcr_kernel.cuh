// "Rows in lanes" message kernel for the slots with l3 >= 1 (d3 >= 3) and padded multiplicity <= 32 -- the slots
// that hold 96 % of the message time (profiles/r01k_launches_m8.csv).  Included by msgpack_tcg.cu (same tables,
// same packed images, same radial-gate pre-pass).
//
// Why: msgpack_tcg_kernel walks (path, K-chunk) steps in lock-step -- stage x -> barrier -> A -> barrier -> MMA ->
// wait -- and each step carries only ~1 k elements, so a CTA spends ~20 k cycles per path on exposed latency.  Here
// the CTA is a warp-specialised pipeline with no CTA-wide barrier in the steady state:
//
//   * warps 0-3 (producers): thread r owns accumulator row r = (edge z, component k) = TMEM lane r.  Per "oct"
//     (8 input channels of one path) a warp stages the 8*d1 contiguous floats of each of its <= 12 edges with
//     cp.async (natural layout, double buffered, one oct ahead), each thread contracts them with its own
//     T_z[.][k] (11 floats, thread-private) into 8 values of A, splits them hi/lo and writes them straight into a
//     TMEM A stage with tcgen05.st (ring of 4 stages x 16 columns) -- there is no shared-memory A image, no
//     transposition and no generic->async proxy hand-off for A.
//   * warp 4 (MMA): prefetches the oct's W slice (ring of 8) and the path's L' image (ring of 2) with cp.async,
//     waits for the stage, issues the 3xTF32 MMAs with A from TMEM (B = A.W into B_q, q = parity of the gated-path
//     counter; direct paths accumulate into C) and commits to the stage's "empty" barrier.
//   * the gate of path p (B_q *= g, split, written back in place as (B.g)hi plus a (B.g)lo block) is run by the
//     producers after they have produced the first oct of the next path, so GEMM1(p) drains and GEMM1(p+1) starts
//     while it runs; GEMM2 (C += (B.g) L') is issued by the MMA warp right after.
//
// TMEM: A ring 64 | C mp | B0 mp | B1 mp | (B.g)lo mp  = 128 columns for mp = 16 (4 CTAs/SM), 192 -> 256 for mp = 32.
#pragma once

namespace tcr {
using namespace tcmsg;

constexpr int NPROD = 128;   // producer threads = accumulator rows
constexpr int NTHR = 160;    // + the MMA warp
constexpr int ASTAGES = 4, WSTAGES = 8;
constexpr int XW = 1056;     // floats per warp per staging buffer: 12 edges x 8 channels x d1 = 11 (host-checked per path)
constexpr int D1MAX = 13;
constexpr int ZSTR = 48;     // row-index stride: at most 128 / 3 = 42 edges per tile

template <int RW>
struct Tm {
  static constexpr uint32_t A = 0, C = 64, B0 = 64 + RW, B1 = 64 + 2 * RW, BGL = 64 + 3 * RW;
  static constexpr int COLS = (RW == 16) ? 128 : 256;
};
template <int RW>
struct Sm {
  static constexpr int X = 0;                          // [4 warps][2][XW]
  static constexpr int T = X + 8 * XW;                 // [D1MAX][128] thread-private columns
  static constexpr int W = T + D1MAX * NPROD;          // ring of WSTAGES x (hi | lo) x [2 slabs][RW][4]
  static constexpr int L = W + WSTAGES * 16 * RW;      // ring of 2 x (hi | lo) x [RW/4 slabs][RW][4]
  static constexpr int ROW = L + 4 * RW * RW;          // int [4 sources][ZSTR]
  static constexpr int TOTAL = ROW + 4 * ZSTR;
};

__device__ __forceinline__ void cp_async4_s(uint32_t dst_smem, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(mbar)) : "memory");
}
// Wait with try_wait (suspend hint) + nanosleep back-off instead of a busy poll: a waiting warp must not eat the
// issue slots of the producer warps on its scheduler.  Bounded: 2^22 expiries (> 0.25 s), then trap.
__device__ __forceinline__ void mbar_wait_suspend(uint64_t* mbar, uint32_t parity) {
  const uint32_t addr = tc::smem_u32(mbar);
  uint32_t done;
  int spins = 0;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(10000u)
        : "memory");
    if (!done) {
      __nanosleep(64);
      if (++spins > (1 << 22)) __trap();
    }
  } while (!done);
}
// one lane waits, the warp re-converges
__device__ __forceinline__ void warp_wait(uint64_t* mbar, uint32_t parity) {
  if ((threadIdx.x & 31) == 0) mbar_wait_suspend(mbar, parity);
  __syncwarp();
}

// A for one row and one oct: acc[c] = sum_i x[c*D1 + i] * T[i]; x = nchan*D1 contiguous floats (16-byte aligned),
// nchan = 4 or 8 staged channels (an oct whose upper half is padding is staged as 4 channels)
template <int D1>
__device__ __forceinline__ void agen_oct(const float* __restrict__ xrow, const float* __restrict__ tcol, int nchan, float (&acc)[8]) {
  float T[D1];
#pragma unroll
  for (int i = 0; i < D1; ++i) T[i] = tcol[i * NPROD];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (h == 1 && nchan <= 4) break;
#pragma unroll
    for (int q = 0; q < D1; ++q) {
      const float4 v = *reinterpret_cast<const float4*>(xrow + h * 4 * D1 + 4 * q);
      const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = 4 * q + c;
        acc[h * 4 + j / D1] = fmaf(e[c], T[j % D1], acc[h * 4 + j / D1]);
      }
    }
  }
}
__device__ __forceinline__ void agen_oct_dispatch(int d1, const float* xrow, const float* tcol, int nchan, float (&acc)[8]) {
  switch (d1) {
    case 1: agen_oct<1>(xrow, tcol, nchan, acc); break;
    case 3: agen_oct<3>(xrow, tcol, nchan, acc); break;
    case 5: agen_oct<5>(xrow, tcol, nchan, acc); break;
    case 7: agen_oct<7>(xrow, tcol, nchan, acc); break;
    case 9: agen_oct<9>(xrow, tcol, nchan, acc); break;
    case 11: agen_oct<11>(xrow, tcol, nchan, acc); break;
    default: agen_oct<13>(xrow, tcol, nchan, acc); break;
  }
}

template <int RW>
__global__ void __launch_bounds__(NTHR, (RW == 16 ? 4 : 2)) msgpack_tcr_kernel(const __grid_constant__ TgArgs a) {
  using SM = Sm<RW>;
  using TM = Tm<RW>;
  constexpr int mp = RW;
  extern __shared__ __align__(128) float smem[];
  float* sX = smem + SM::X;
  float* sTp = smem + SM::T;
  float* sW = smem + SM::W;
  float* sL = smem + SM::L;
  int* sRow = reinterpret_cast<int*>(smem + SM::ROW);
  __shared__ uint64_t afull[ASTAGES], aempty[ASTAGES], bfull[2], ldone[2], gfull, gdone, cdone;
  __shared__ uint32_t tmem_slot;

  const hgb_msgpack_plan& P = a.plan;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* __restrict__ wbuf = P.wbuf;

  int sq = 0;
  while (sq + 1 < a.n_sched && (int)blockIdx.x >= a.tile_start[sq + 1]) ++sq;
  const int t = a.type_order[sq];
  const hgb_type_t ty = P.types[t];
  const int d3 = 2 * ty.l + 1;
  const int nz_full = ROWS / d3;
  const int64_t e0 = (int64_t)(blockIdx.x - a.tile_start[sq]) * nz_full;
  const int nz = (int)min((int64_t)nz_full, a.n_edges - e0);
  const int R = nz * d3;
  const int S = P.sh_dim;
  const int npaths = ty.path_end - ty.path_begin;

  if (tid == 0) {
    for (int i = 0; i < ASTAGES; ++i) { tc::mbar_init(&afull[i], 4); tc::mbar_init(&aempty[i], 1); }
    tc::mbar_init(&bfull[0], 1); tc::mbar_init(&bfull[1], 1);
    tc::mbar_init(&ldone[0], 1); tc::mbar_init(&ldone[1], 1);
    tc::mbar_init(&gfull, 4); tc::mbar_init(&gdone, 1); tc::mbar_init(&cdone, 1);
    tc::mbar_fence_init();
  }
  if (warp == 4) tc::tmem_alloc<TM::COLS>(&tmem_slot);
  for (int idx = tid; idx < P.n_sources * nz; idx += NTHR) {
    const int s = idx / nz, z = idx - s * nz;
    const int64_t e = e0 + z;
    sRow[s * ZSTR + z] = (int)(a.src_rows[s] ? a.src_rows[s][e] : e);
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = tc::idesc_tf32_m128(mp);

  if (warp == 4) {
    // =============================== MMA warp ===============================
    const uint32_t dhi = tc::smem_desc_hi(128);
    const uint32_t lbo_n = (uint32_t)mp * 16, kstep = (2 * lbo_n) >> 4;
    // prefetch iterator (runs WSTAGES/2 octs ahead of the issue iterator)
    int fp = ty.path_begin, fo = 0, fn = 0, fKpad = 0;
    const float* fbase = nullptr;
    auto load_prefetch_path = [&]() {
      const hgb_path_t& pa = P.paths[fp];
      fKpad = (pa.nsrc * pa.mul_in + 7) & ~7;
      fbase = wbuf + (pa.kind == 0 ? pa.w_off : pa.lf_off);
    };
    if (npaths > 0) load_prefetch_path();
    auto prefetch_next = [&]() {
      if (fp < ty.path_end) {
        const int u0 = 8 * fo, img = u0 / KIMG, kimg = min(KIMG, fKpad - img * KIMG), s0 = (u0 - img * KIMG) >> 2;
        const float* base = fbase + (size_t)2 * mp * KIMG * img + (size_t)s0 * mp * 4;
        float* dst = sW + (fn & (WSTAGES - 1)) * (16 * mp);
#pragma unroll
        for (int i = lane; i < 2 * mp; i += 32) {
          cp_async16(dst + 4 * i, base + 4 * i);
          cp_async16(dst + 8 * mp + 4 * i, base + (size_t)mp * kimg + 4 * i);
        }
        ++fn;
        if (8 * (++fo) >= fKpad) {
          ++fp; fo = 0;
          if (fp < ty.path_end) load_prefetch_path();
        }
      }
    };
    for (int i = 0; i < 4; ++i) { prefetch_next(); cp_async_commit(); }

    int n = 0, q = 0, pend_q = -1;
    bool c_started = false;
    auto issue_mma2 = [&]() {
      warp_wait(&gfull, (uint32_t)(pend_q & 1));
      cp_async_wait_all();
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tc::fence_after_sync();
        const uint32_t bq = tmem + ((pend_q & 1) ? TM::B1 : TM::B0);
        const uint32_t lh = tc::smem_desc_lo(tc::smem_u32(sL + (pend_q & 1) * (2 * mp * mp)), lbo_n), ll = lh + (((uint32_t)mp * mp * 4) >> 4);
#pragma unroll
        for (int k8 = 0; k8 < mp / 8; ++k8) {
          const uint64_t bh = tc::desc64(lh + k8 * kstep, dhi), bl = tc::desc64(ll + k8 * kstep, dhi);
          tc::mma_tf32_ts(tmem + TM::C, tmem + TM::BGL + k8 * 8, bh, idesc, (uint32_t)(c_started || k8 > 0));
          tc::mma_tf32_ts(tmem + TM::C, bq + k8 * 8, bl, idesc, 1);
          tc::mma_tf32_ts(tmem + TM::C, bq + k8 * 8, bh, idesc, 1);
        }
        tc::mma_commit(&gdone);
        tc::mma_commit(&ldone[pend_q & 1]);
      }
      c_started = true;
      pend_q = -1;
    };
    for (int p = ty.path_begin; p < ty.path_end; ++p) {
      const hgb_path_t pa = P.paths[p];
      const int nocts = (pa.nsrc * pa.mul_in + 7) >> 3;
      for (int o = 0; o < nocts; ++o, ++n) {
        const int s = n & (ASTAGES - 1);
        warp_wait(&afull[s], (uint32_t)((n >> 2) & 1));
        if (o == 0 && pa.kind == 0) {
          // L' of this path -> sL[q & 1]; its previous user was GEMM2 of gated path q - 2
          if (q >= 2) warp_wait(&ldone[q & 1], (uint32_t)(((q >> 1) - 1) & 1));
          const float* lsrc = wbuf + pa.lf_off;
          float* ldst = sL + (q & 1) * (2 * mp * mp);
          for (int i = lane; i < (mp * mp) / 2; i += 32) cp_async16(ldst + 4 * i, lsrc + 4 * i);
        }
        prefetch_next();   // W of oct n + 4 -> slot (n + 4) % 8, last read by the MMAs of oct n - 4 (complete: afull[n] seen)
        cp_async_commit();
        cp_async_wait_group<4>();   // the group that carried W of oct n has landed
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tc::fence_after_sync();
          const uint32_t acol = tmem + TM::A + s * 16;
          const uint32_t wh = tc::smem_desc_lo(tc::smem_u32(sW + (n & (WSTAGES - 1)) * (16 * mp)), lbo_n), wl = wh + (((uint32_t)mp * 8 * 4) >> 4);
          const uint64_t dbh = tc::desc64(wh, dhi), dbl = tc::desc64(wl, dhi);
          uint32_t dcol, acc0;
          if (pa.kind == 0) { dcol = tmem + ((q & 1) ? TM::B1 : TM::B0); acc0 = (uint32_t)(o > 0); }
          else { dcol = tmem + TM::C; acc0 = (uint32_t)(c_started || o > 0); }
          tc::mma_tf32_ts(dcol, acol + 8, dbh, idesc, acc0);   // lo . hi
          tc::mma_tf32_ts(dcol, acol, dbl, idesc, 1);          // hi . lo
          tc::mma_tf32_ts(dcol, acol, dbh, idesc, 1);          // hi . hi
          tc::mma_commit(&aempty[s]);
          if (o == nocts - 1 && pa.kind == 0) tc::mma_commit(&bfull[q & 1]);
        }
        if (pa.kind != 0) c_started = true;
        if (o == 0 && pend_q >= 0) issue_mma2();
      }
      if (pa.kind == 0) { pend_q = q; ++q; }
    }
    if (pend_q >= 0) issue_mma2();
    if (lane == 0 && npaths > 0) tc::mma_commit(&cdone);
  } else {
    // =============================== producers ===============================
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const int my_row = tid;
    const bool live = my_row < R;
    const int my_z = live ? my_row / d3 : 0;
    const int my_k = live ? my_row - my_z * d3 : 0;
    // edges touched by this warp's 32 rows
    const int zlo = (warp * 32) / d3;
    const int nzw = (warp * 32 < R) ? (min((warp * 32 + 31) / d3, nz - 1) - zlo + 1) : 0;
    float* xbuf = sX + warp * (2 * XW);
    float* tcol = sTp + tid;

    // Staging geometry: G = 2 / 4 / 8 lanes per edge row (the warp touches <= 16 / 8 / 4 edges), lane -> (edge zl_l, g).
    // A lane's row offsets in the input sources are fixed for the tile.
    const int gsh = (nzw <= 4) ? 3 : (nzw <= 8 ? 2 : 1);
    const int G = 1 << gsh;
    const int zl_l = lane >> gsh, g = lane & (G - 1);
    const bool rowok = zl_l < nzw;
    size_t roff[4];
#pragma unroll
    for (int sidx = 0; sidx < 4; ++sidx)
      roff[sidx] = (rowok && sidx < P.n_sources) ? (size_t)sRow[sidx * ZSTR + zlo + zl_l] * (size_t)P.src_dim[sidx] : 0;
    auto row_ptr = [&](int sidx) -> const float* {
      const float* b = sidx == 0 ? a.src[0] : (sidx == 1 ? a.src[1] : (sidx == 2 ? a.src[2] : a.src[3]));
      const size_t o = sidx == 0 ? roff[0] : (sidx == 1 ? roff[1] : (sidx == 2 ? roff[2] : roff[3]));
      return b + o;
    };
    // gate of gated path number qg: B_q <- (B_q . g) hi in place, (B.g) lo block; then hand over to the MMA warp
    auto do_gate = [&](int branch, int goff, int qg) {
      // the gate values are fetched BEFORE the barrier waits so that their latency overlaps the drain of GEMM1
      // (the row segment was pulled into L2 when the path started, see the prefetch below)
      const float* gp = a.g + ((size_t)branch * a.n_edges + (size_t)(e0 + my_z)) * a.gstride + goff;
      float gv[mp];
#pragma unroll
      for (int j = 0; j < mp; ++j) gv[j] = (live && j < ty.mul) ? __ldg(gp + j) : 0.f;
      warp_wait(&bfull[qg & 1], (uint32_t)((qg >> 1) & 1));
      if (qg >= 1) warp_wait(&gdone, (uint32_t)((qg - 1) & 1));   // GEMM2 of the previous gated path has read (B.g) lo
      tc::fence_after_sync();
      const uint32_t bq = tmem + lane_base + ((qg & 1) ? TM::B1 : TM::B0);
#pragma unroll
      for (int c0 = 0; c0 < mp; c0 += 8) {
        uint32_t rb[8], hi[8], lo[8];
        tc::tmem_ld8(bq + c0, rb);
        tc::tmem_ld_wait8(rb);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float h, l;
          tc::split_tf32(__uint_as_float(rb[j]) * gv[c0 + j], h, l);
          hi[j] = __float_as_uint(h); lo[j] = __float_as_uint(l);
        }
        tc::tmem_st8(bq + c0, hi);
        tc::tmem_st8(tmem + lane_base + TM::BGL + c0, lo);
      }
      tc::tmem_st_wait();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&gfull);
    };

    // ---- staging cursor (one oct ahead of the compute cursor) and its per-path constants
    int s_p = ty.path_begin, s_o = 0, s_d1 = 1, s_m = 0, s_K = 0, s_nsrc = 1;
    const float* s_pA = nullptr;
    const float* s_pB = nullptr;
    auto load_stage_path = [&]() {
      const hgb_path_t& pa = P.paths[s_p];
      s_d1 = 2 * pa.l1 + 1; s_m = pa.mul_in; s_nsrc = pa.nsrc; s_K = pa.nsrc * pa.mul_in;
      s_pA = row_ptr(pa.src0) + pa.in_off;
      s_pB = row_ptr(pa.src0 + pa.nsrc - 1) + pa.in_off;
    };
    if (npaths > 0) load_stage_path();

    // ---- compute cursor
    int p = ty.path_begin, o = 0, q = 0, buf = 0;
    int pend_q = -1, pend_branch = 0, pend_goff = 0;
    int c_d1 = 1, c_K = 0, c_nocts = 0, c_kind = 0, c_branch = 0, c_goff = 0, c_tkey = -1;

    for (int n = -1; p < ty.path_end; ++n) {
      // -- stage the next oct: channels [8 s_o, 8 s_o + 8) of path s_p -> xbuf[buf ^ 1][zl][nchan * d1]
      const bool more = s_p < ty.path_end;
      if (more) {
        const int u0 = 8 * s_o;
        const int nchan = (min(8, s_K - u0) + 3) & ~3;      // 4 or 8 staged channels (rows of nchan*d1 floats back to back:
        const int L = nchan * s_d1;                          // d1 or 2*d1 quads -> distinct bank groups for <= 4 edges)
        const int ua_end = (s_nsrc == 2) ? min(u0 + 8, s_m) : min(u0 + 8, s_K);
        const int nA = max(0, ua_end - u0) * s_d1;
        const int ub = max(u0, s_m);
        const int nAB = nA + ((s_nsrc == 2) ? max(0, min(u0 + 8, s_K) - ub) * s_d1 : 0);
        if (rowok) {
          float* d = xbuf + ((n < 0) ? buf : (buf ^ 1)) * XW + zl_l * L;
          const uint32_t d32 = tc::smem_u32(d);
          const float* pa_ = s_pA + u0 * s_d1;
          const float* pb_ = s_pB + (ub - s_m) * s_d1 - nA;
#pragma unroll 4
          for (int j = g; j < L; j += G) {
            if (j < nAB) cp_async4_s(d32 + 4 * j, (j < nA ? pa_ : pb_) + j);
            else d[j] = 0.f;
          }
        }
        cp_async_commit();
        if (8 * (++s_o) >= s_K) {
          ++s_p; s_o = 0;
          if (s_p < ty.path_end) load_stage_path();
        }
      }
      if (n < 0) continue;

      if (o == 0) {
        const hgb_path_t& pa = P.paths[p];
        c_d1 = 2 * pa.l1 + 1; c_K = pa.nsrc * pa.mul_in; c_nocts = (c_K + 7) >> 3;
        c_kind = pa.kind; c_branch = pa.branch; c_goff = pa.pad0;
        // T_z[.][k] of this path, thread-private column of sTp; consecutive gated paths with the same (l1, l2) --
        // the two branches of one CG path -- share it
        const int tkey = (pa.kind == 0) ? (pa.l1 * 16 + pa.l2) : -1;
        const bool t_same = (tkey >= 0 && tkey == c_tkey);
        c_tkey = tkey;
        if (!t_same)
          for (int i = 0; i < c_d1; ++i) tcol[i * NPROD] = 0.f;
        if (live && c_kind == 0 && my_k == 0) {
          // pull this path's gate row segment (mp floats of the [E, n_channels] gate tensor, streamed from HBM)
          // into L2 now; do_gate reads it one path later
          const float* gp = a.g + ((size_t)c_branch * a.n_edges + (size_t)(e0 + my_z)) * a.gstride + c_goff;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(gp));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(gp + ty.mul - 1));
        }
        if (live && !t_same) {
          if (c_kind == 0) {
            const float* yz = a.sh + (e0 + my_z) * S + pa.sh_off;
            const int* cij = P.cg_ij + pa.cg_off;
            const float* cval = P.cg_val + pa.cg_off;
            const int n0 = __ldg(P.cg_kstart + pa.cg_kstart + my_k), n1 = __ldg(P.cg_kstart + pa.cg_kstart + my_k + 1);
            // four table entries per round: all index/value loads, then all Y loads, then the accumulations, so a
            // round costs one dependent-load chain instead of four
            for (int c = n0; c < n1; c += 4) {
              int ij[4];
              float cv[4], yv[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const bool ok = c + u < n1;
                ij[u] = ok ? __ldg(cij + c + u) : 0;
                cv[u] = ok ? __ldg(cval + c + u) : 0.f;
              }
#pragma unroll
              for (int u = 0; u < 4; ++u) yv[u] = __ldg(yz + (ij[u] >> 8));
#pragma unroll
              for (int u = 0; u < 4; ++u) tcol[(ij[u] & 255) * NPROD] += cv[u] * yv[u];
            }
          } else {
            tcol[my_k * NPROD] = 1.f;   // d1 == d3
          }
        }
      }
      if (more) cp_async_wait_group<1>(); else cp_async_wait_all();
      __syncwarp();
      float acc[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = 0.f;
      const int nchan = (min(8, c_K - 8 * o) + 3) & ~3;
      if (live) agen_oct_dispatch(c_d1, xbuf + buf * XW + (my_z - zlo) * (nchan * c_d1), tcol, nchan, acc);
      __syncwarp();   // the buffer may be refilled by the next staging pass
      const int s = n & (ASTAGES - 1);
      if (n >= ASTAGES) warp_wait(&aempty[s], (uint32_t)(((n >> 2) - 1) & 1));
      tc::fence_after_sync();
      {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float h, l;
          tc::split_tf32(acc[c], h, l);
          hi[c] = __float_as_uint(h); lo[c] = __float_as_uint(l);
        }
        tc::tmem_st8(tmem + lane_base + TM::A + s * 16, hi);
        tc::tmem_st8(tmem + lane_base + TM::A + s * 16 + 8, lo);
      }
      tc::tmem_st_wait();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&afull[s]);
      buf ^= 1;
      // gate of the previous gated path, one oct into this path; and of the last path right after its last oct
      const bool last_oct = (o + 1 == c_nocts);
      const bool final_oct = last_oct && (p + 1 == ty.path_end);
      const bool g0 = (o == 0 && pend_q >= 0);
      const int gq = pend_q, gb = pend_branch, gg = pend_goff;
      if (g0) pend_q = -1;
      if (last_oct) {
        if (c_kind == 0) { pend_q = q; pend_branch = c_branch; pend_goff = c_goff; ++q; }
        ++p; o = 0;
      } else {
        ++o;
      }
#pragma unroll 1
      for (int rep = 0; rep < 2; ++rep) {
        const bool run = (rep == 0) ? g0 : (final_oct && pend_q >= 0);
        if (run) do_gate(rep == 0 ? gb : pend_branch, rep == 0 ? gg : pend_goff, rep == 0 ? gq : pend_q);
      }
    }

    // ---- epilogue: C -> global
    if (npaths > 0) { warp_wait(&cdone, 0); tc::fence_after_sync(); }
    {
      const int64_t e = e0 + my_z;
      const int64_t orow = (live && a.out_index) ? a.out_index[e] : e;
      float* op = a.out + orow * P.out_dim + ty.out_off + my_k;
#pragma unroll
      for (int c0 = 0; c0 < mp; c0 += 8) {
        if (c0 >= ty.mul) break;
        uint32_t rc[8];
        if (npaths > 0) {
          tc::tmem_ld8(tmem + lane_base + TM::C + c0, rc);
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
              if (a.out_index) atomicAdd(op + w * d3, __uint_as_float(rc[j]));
              else op[w * d3] = __uint_as_float(rc[j]);
            }
          }
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc<TM::COLS>(tmem);
}

}  // namespace tcr
