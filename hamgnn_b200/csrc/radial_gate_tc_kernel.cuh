// Radial gate pre-pass on the tensor cores: g[b][e][c] = (act(act(rbf W1_b) W2_b) W3_b)[c] for every mid channel c of
// branch b (a7: e3nn FullyConnectedNet [R, h1, h2, n_channels], hamgnn/nn/message_passing.py:173-189).  The last layer
// is a real GEMM -- [E x h2] . [h2 x 3589] per branch, 0.92 MFLOP/edge, 79 % of the radial MLP -- and the only dense
// contraction of the message path whose N is large, so it maps onto full-width tcgen05 tiles:
//
//   CTA = (128 edges, one branch), 160 threads.
//   * prologue (warps 0-3, thread = edge row): layers 1 and 2 in fp32 FMA from shared memory (weights broadcast,
//     inputs in a stride-65 tile), the activated h2 row is split hi/lo (3xTF32) and written as the A operand
//     images [h2/4][128][4] (K-major, no swizzle) -- A stays resident for the whole CTA.
//   * warp 5 (one thread): streams the host-packed W3 tiles (64 gate columns each, hi | lo images [h2/4][64][4], L2 resident,
//     32 KB contiguous) through a ring of 3 with ONE cp.async.bulk per tile (r05: the per-lane cp.async loop of rounds 1-2 --
//     64 x 16 B per lane and tile, issued by the MMA warp itself -- was ~1 500 cycles of the issuing warp per tile).
//   * warp 4: issues 3 x h2/8 tcgen05.mma.kind::tf32 (M 128, N 64) per tile into one of two TMEM accumulators and commits
//     to the tile's barriers.
//   * warps 0-3: drain the other accumulator (tcgen05.ld, 32 columns per wait) and store 256 contiguous bytes per
//     row and tile; every 32-byte sector of g is written by one thread in two back-to-back stores.
//
// The prologue scratch (rbf tile, h1 tile, W1/W2) aliases the W3 ring, which is idle until the A images are complete.
// HBM: reads 256 B/edge (rbf) per branch, writes 4 n_channels B/edge; bound by the gate write (14.4 KB/edge/branch).
#pragma once

namespace gtc {
using namespace tcmsg;

constexpr int NT = 192;   // warps 0-3 prologue + epilogue, warp 4 MMA issuer, warp 5 W3-tile loader (one cp.async.bulk per tile)
constexpr int TN = 128;    // gate columns per MMA tile = MMA N (plan.MessagePackOp.GATE_TILE_COLS); r05: 64 -> 128 halves the MMAs per column
constexpr int KMAX = 64;   // h2 <= 64
constexpr int WRING = 2;   // two 64 KB tiles (hi | lo images of 128 columns)

// Barrier wait of this kernel: a plain try_wait loop (try_wait itself parks the thread for a hardware time slice).  The
// tcr::warp_wait it used in round 1 adds a try_wait suspend hint of 10 us and a nanosleep back-off: 36 % of this kernel's
// stall samples were `stall_sleep` (profiles/r02y), every MMA <-> epilogue hand-off of its 57 tiles overslept.
__device__ __forceinline__ void spin_wait(uint64_t* mbar, uint32_t parity) {
  if ((threadIdx.x & 31) == 0) {
    const uint32_t addr = tc::smem_u32(mbar);
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .u32 n;\n\t"
        "mov.u32 n, 0;\n\t"
        "mov.u32 %0, 1;\n"
        "HGB_GTC_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "@p bra HGB_GTC_DONE;\n\t"
        "add.u32 n, n, 1;\n\t"
        "setp.lt.u32 p, n, 0x4000000;\n\t"
        "@p bra HGB_GTC_WAIT;\n\t"
        "mov.u32 %0, 0;\n"
        "HGB_GTC_DONE:\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!ok) __trap();
  }
  __syncwarp();
}

__device__ __forceinline__ bool elect_lane() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

struct Sm {
  static constexpr int AHI = 0;
  static constexpr int ALO = AHI + KMAX * ROWS;
  static constexpr int RING = ALO + KMAX * ROWS;
  static constexpr int TILE = 2 * KMAX * TN;
  static constexpr int TOTAL = RING + WRING * TILE;
  static constexpr int LDI = 65;   // odd row stride of the thread-private input rows
  static constexpr int IN = RING, H1 = IN + ROWS * LDI, W = H1 + ROWS * LDI;
};
static_assert(Sm::W + 64 * 64 <= Sm::TOTAL, "prologue scratch must fit inside the W3 ring");

struct Args {
  const float* rbf;
  const float* w1[2];
  const float* w2[2];
  const float* w3img[2];   // per branch: tiles of (hi | lo) images, each [h2/4][TN][4]
  int nch[2];
  float* g;                // [n_branches][E][gstride]
  int64_t n_edges;
  int rbf_dim, h1, h2dim, gstride;
  float act_const;
  int tmajor;              // 2: [tile][n_branches][gstride][128] (rot2: a gate block has ONE offset inside its tile);
                           // 1: g is [n_branches][tile of 128 edges][gstride][128] (edge-minor inside a tile: the layout the
                           // rotated-frame message kernel reads, coalesced on both sides); 0: [n_branches][E][gstride]
};

// one dense layer for one row: out[j] = act(sum_r xin[r] * W[r][j]) * c, j in [j0, j0 + 16); W row-major [n_in][ldw]
__device__ __forceinline__ void layer16(const float* __restrict__ xin, const float* __restrict__ sW, int n_in, int ldw, int j0,
                                        float c, float (&out)[16]) {
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
  for (int r = 0; r < n_in; ++r) {
    const float x = xin[r];
    const float4* w = reinterpret_cast<const float4*>(sW + r * ldw + j0);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 wv = w[q];
      acc[4 * q + 0] = fmaf(x, wv.x, acc[4 * q + 0]);
      acc[4 * q + 1] = fmaf(x, wv.y, acc[4 * q + 1]);
      acc[4 * q + 2] = fmaf(x, wv.z, acc[4 * q + 2]);
      acc[4 * q + 3] = fmaf(x, wv.w, acc[4 * q + 3]);
    }
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) out[j] = hgb::silu_f(acc[j]) * c;
}

__global__ void __launch_bounds__(NT, 1) radial_gate_tc_kernel(const __grid_constant__ Args a) {
  extern __shared__ __align__(128) float smem[];
  __shared__ uint64_t dfull[2], dempty[2], wdone[WRING], wfull[WRING];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  const int64_t e0 = (int64_t)blockIdx.x * ROWS;
  const int ne = (int)min((int64_t)ROWS, a.n_edges - e0);
  const int K = a.h2dim;
  const int nch = a.nch[b];
  const int ntiles = (nch + TN - 1) / TN;
  float* sAhi = smem + Sm::AHI;
  float* sAlo = smem + Sm::ALO;
  float* sIn = smem + Sm::IN;
  float* sH1 = smem + Sm::H1;
  float* sW = smem + Sm::W;

  if (tid == 0) {
    tc::mbar_init(&dfull[0], 1); tc::mbar_init(&dfull[1], 1);
    tc::mbar_init(&dempty[0], 4); tc::mbar_init(&dempty[1], 4);
    for (int i = 0; i < WRING; ++i) { tc::mbar_init(&wdone[i], 1); tc::mbar_init(&wfull[i], 1); }
    tc::mbar_fence_init();
  }
  if (warp == 4) tc::tmem_alloc<2 * TN>(&tmem_slot);

  // ---- prologue: h2 = act(act(rbf W1) W2), split, A images
  for (int idx = tid; idx < ROWS * a.rbf_dim; idx += NT) {
    const int z = idx / a.rbf_dim, c = idx - z * a.rbf_dim;
    sIn[z * Sm::LDI + c] = (z < ne) ? __ldg(a.rbf + (e0 + z) * a.rbf_dim + c) : 0.f;
  }
  for (int idx = tid; idx < a.rbf_dim * a.h1; idx += NT) sW[idx] = __ldg(a.w1[b] + idx);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (tid < ROWS) {
    for (int j0 = 0; j0 < a.h1; j0 += 16) {
      float h[16];
      layer16(sIn + tid * Sm::LDI, sW, a.rbf_dim, a.h1, j0, a.act_const, h);
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j0 + j < a.h1) sH1[tid * Sm::LDI + j0 + j] = h[j];
    }
  }
  __syncthreads();
  for (int idx = tid; idx < a.h1 * K; idx += NT) sW[idx] = __ldg(a.w2[b] + idx);
  __syncthreads();
  if (tid < ROWS) {
    for (int j0 = 0; j0 < K; j0 += 16) {
      float h[16];
      layer16(sH1 + tid * Sm::LDI, sW, a.h1, K, j0, a.act_const, h);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = j0 + 4 * q;
        if (k < K) {   // K % 8 == 0: whole quads
          float4 hi, lo;
          tc::split_tf32(h[4 * q + 0], hi.x, lo.x); tc::split_tf32(h[4 * q + 1], hi.y, lo.y);
          tc::split_tf32(h[4 * q + 2], hi.z, lo.z); tc::split_tf32(h[4 * q + 3], hi.w, lo.w);
          *reinterpret_cast<float4*>(sAhi + (k >> 2) * (ROWS * 4) + tid * 4) = hi;
          *reinterpret_cast<float4*>(sAlo + (k >> 2) * (ROWS * 4) + tid * 4) = lo;
        }
      }
    }
  }
  tc::fence_proxy_async();   // A images (generic-proxy writes) -> visible to the tensor core's operand fetch
  __syncthreads();           // from here on the scratch region belongs to the W3 ring

  const int img = K * TN;    // floats per operand image of one tile
  if (warp == 5) {
    // =============================== W3-tile loader ===============================
    if (lane == 0) {
      const float* wsrc = a.w3img[b];
      const uint32_t bytes = (uint32_t)(2 * img) * 4u;
      for (int t = 0; t < ntiles; ++t) {
        const int slot = t % WRING;
        if (t >= WRING) {   // the MMAs of tile t - WRING have read the slot
          const uint32_t addr = tc::smem_u32(&wdone[slot]), parity = (uint32_t)(((t / WRING) - 1) & 1);
          uint32_t ok = 0, spins = 0;
          while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
            if (++spins > 0x4000000u) __trap();
          }
        }
        const uint32_t mb = tc::smem_u32(&wfull[slot]);
        const uint32_t dst = tc::smem_u32(smem + Sm::RING + slot * Sm::TILE);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                     "l"(wsrc + (size_t)t * 2 * img), "r"(bytes), "r"(mb)
                     : "memory");
      }
    }
    __syncwarp();
  } else if (warp == 4) {
    // =============================== MMA warp ===============================
    const uint32_t idesc = tc::idesc_tf32_m128(TN);
    const uint32_t dhi = tc::smem_desc_hi(128);
    const uint32_t lbo_a = ROWS * 16, lbo_b = TN * 16;
    const uint32_t astep = (2 * lbo_a) >> 4, bstep = (2 * lbo_b) >> 4;
    for (int t = 0; t < ntiles; ++t) {
      spin_wait(&wfull[t % WRING], (uint32_t)((t / WRING) & 1));                 // tile t has landed (async-proxy writes)
      if (t >= 2) spin_wait(&dempty[t & 1], (uint32_t)(((t >> 1) - 1) & 1));   // accumulator drained (tile t - 2)
      tc::fence_after_sync();
      if (elect_lane()) {   // convergent warp + elect.sync: descriptors stay in uniform registers, the UTCHMMA issue back to back
        const float* wt = smem + Sm::RING + (t % WRING) * Sm::TILE;
        const uint32_t ah = tc::smem_desc_lo(tc::smem_u32(sAhi), lbo_a), al = tc::smem_desc_lo(tc::smem_u32(sAlo), lbo_a);
        const uint32_t wh = tc::smem_desc_lo(tc::smem_u32(wt), lbo_b), wl = tc::smem_desc_lo(tc::smem_u32(wt + img), lbo_b);
        const uint32_t dcol = tmem + (uint32_t)((t & 1) * TN);
        for (int k8 = 0; k8 < (K >> 3); ++k8) {
          const uint64_t dah = tc::desc64(ah + k8 * astep, dhi), dal = tc::desc64(al + k8 * astep, dhi);
          const uint64_t dbh = tc::desc64(wh + k8 * bstep, dhi), dbl = tc::desc64(wl + k8 * bstep, dhi);
          tc::mma_tf32(dcol, dal, dbh, idesc, (uint32_t)(k8 > 0));
          tc::mma_tf32(dcol, dah, dbl, idesc, 1);
          tc::mma_tf32(dcol, dah, dbh, idesc, 1);
        }
        tc::mma_commit(&dfull[t & 1]);
        tc::mma_commit(&wdone[t % WRING]);
      }
      __syncwarp();
    }
  } else {
    // =============================== epilogue warps: TMEM -> g ===============================
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const bool live = tid < ne;
    float* grow = a.g + ((size_t)b * a.n_edges + (size_t)(e0 + (live ? tid : 0))) * a.gstride;
    for (int t = 0; t < ntiles; ++t) {
      spin_wait(&dfull[t & 1], (uint32_t)((t >> 1) & 1));
      tc::fence_after_sync();
      const int n0 = t * TN;
#pragma unroll
      for (int h = 0; h < TN / 32; ++h) {
        uint32_t r[4][8];
        const uint32_t col = tmem + lane_base + (uint32_t)((t & 1) * TN + h * 32);
        tc::tmem_ld8(col, r[0]); tc::tmem_ld8(col + 8, r[1]); tc::tmem_ld8(col + 16, r[2]); tc::tmem_ld8(col + 24, r[3]);
        tc::tmem_ld_wait8(r[0]); tc::tmem_ld_wait8(r[1]); tc::tmem_ld_wait8(r[2]); tc::tmem_ld_wait8(r[3]);
        if (a.tmajor) {
          // every thread (padding rows of the last tile included: the workspace is padded) writes one float per column;
          // a warp covers 128 contiguous bytes
          const size_t blk = (a.tmajor == 2) ? (size_t)blockIdx.x * gridDim.y + b : (size_t)b * gridDim.x + blockIdx.x;
          float* gt = a.g + blk * (size_t)a.gstride * ROWS + tid;
#pragma unroll
          for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int c = n0 + h * 32 + q * 8 + j;
              if (c < nch) gt[(size_t)c * ROWS] = __uint_as_float(r[q][j]);
            }
        } else if (live) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int c = n0 + h * 32 + q * 8;
            if (c + 8 <= nch) {
              float4* o = reinterpret_cast<float4*>(grow + c);
              o[0] = make_float4(__uint_as_float(r[q][0]), __uint_as_float(r[q][1]), __uint_as_float(r[q][2]), __uint_as_float(r[q][3]));
              o[1] = make_float4(__uint_as_float(r[q][4]), __uint_as_float(r[q][5]), __uint_as_float(r[q][6]), __uint_as_float(r[q][7]));
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (c + j < nch) grow[c + j] = __uint_as_float(r[q][j]);
            }
          }
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tcr::mbar_arrive(&dempty[t & 1]);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc<2 * TN>(tmem);
}

}  // namespace gtc
