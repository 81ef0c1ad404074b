// Edge-aligned ("rotated frame") fused MessagePackBlock.  Included by msgpack_tcg.cu (same packed W / L' images, same
// radial-gate pre-pass).
//
// Why: in the frame where the edge vector is the polar axis the real spherical harmonics reduce to
// Y_l2 = sqrt(2 l2 + 1) delta_{m2,0}, so the per-edge CG contraction T[i,k] = sum_j w3j[i,j,k] Y[j] has ONE non-zero per
// output component m3 (at m1 = m3 for even l1+l2+l3, at m1 = -m3 for odd).  The SIMT "A generation" of the other
// message kernels (K d1 d3 FMAs per edge and path, 95 % of their instructions, profiles/r01l_*) disappears: a path is
// <= min(d1, d3) steps
//     C'_{m3} += ((X'_{m1} W_p) * (scale * g_p)) L'_p
// whose A operand X'_{m1}[z, u] = (D^{l1}(R_z) x_z)[u, m1] is plain data.  Three kernels:
//   wigner_kernel       D^l(R_z) per edge, l <= 6, fp64 arithmetic (once per forward)
//   rotate_pack_kernel  gathers the input rows, rotates every irrep block, splits hi/lo and writes ready-made UMMA
//                       operand images (tile of 128 edges = 128 MMA rows)
//   msgpack_rot_kernel  one CTA per (tile, output slot), steps ordered by output component m3: warp 5 streams A / W chunks
//                       and L' images with TMA bulk copies into an mbarrier ring and prefetches the gate blocks into L2,
//                       warp 4 issues GEMM1 (tcgen05 3xTF32), warps 0-3 (thread = edge = TMEM lane) apply the gate in
//                       TMEM, warp 6 issues GEMM2 into a fresh accumulator, warps 0-3 add it into fp32 registers, park
//                       each finished component in TMEM and finally rotate the message back, C = D^{l3}(R_z)^T C',
//                       before the store / receiver scatter-add.
#pragma once

namespace rot {
using namespace tcmsg;

constexpr int TILE = 128;   // edges per tile = MMA rows
constexpr int KC = 32;      // channels per operand chunk (host: MessagePackOp.ROT_KC)
constexpr int NTHR = 192;
constexpr int LMAX = 6;

using tcr::mbar_arrive;

// Barrier waits sit on the critical path of every step here (three hops per step), so they spin on try_wait -- which
// itself suspends the thread for a hardware time slice -- without the nanosleep back-off of the rows-in-lanes kernel
// (measured: 14 % of stall samples in nanosleep, ~3000 cycles per step, profiles/r01o_*).  Bounded: ~2 s, then trap.
__device__ __forceinline__ void mbar_wait_suspend(uint64_t* mbar, uint32_t parity) {
  const uint32_t addr = tc::smem_u32(mbar);
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .u32 n;\n\t"
      "mov.u32 n, 0;\n\t"
      "mov.u32 %0, 1;\n"
      "HGB_ROT_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "@p bra HGB_ROT_DONE;\n\t"
      "add.u32 n, n, 1;\n\t"
      "setp.lt.u32 p, n, 0x4000000;\n\t"
      "@p bra HGB_ROT_WAIT;\n\t"
      "mov.u32 %0, 0;\n"
      "HGB_ROT_DONE:\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(addr), "r"(parity)
      : "memory");
  if (!ok) __trap();   // a pipeline bug must surface as a kernel error, never as a hung GPU
}
__device__ __forceinline__ void warp_wait(uint64_t* mbar, uint32_t parity) {
  if ((threadIdx.x & 31) == 0) mbar_wait_suspend(mbar, parity);
  __syncwarp();
}

__device__ __forceinline__ void tmem_alloc_dyn(uint32_t* smem_slot, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(smem_slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_dyn(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(mbar)), "r"(bytes) : "memory");
}
// TMA bulk prefetch of a contiguous global block into L2 (16-byte aligned, size % 16 == 0)
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on `mbar` (16-byte aligned, size % 16 == 0)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc::smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(tc::smem_u32(mbar))
               : "memory");
}

// ===================================================================================================== wigner
struct WigArgs {
  const float* vec;
  const double* J;
  float* dw;
  int64_t n_edges;
  int dstride;
  int doff[12];
};

// D = J Z(-theta) J^T Z(-phi);  Z(psi): Y'_{+m} = cos(m psi) Y_{+m} - sin(m psi) Y_{-m}, Y'_{-m} = sin(m psi) Y_{+m} + cos(m psi) Y_{-m}
template <int L>
__device__ void wigner_l(const double* __restrict__ J, double cp, double sp, double ct, double st, float* __restrict__ out) {
  constexpr int d = 2 * L + 1;
  double ca[L + 1], sa[L + 1], cb[L + 1], sb[L + 1];   // cos / sin of m * (-phi), m * (-theta)
  ca[0] = 1.0; sa[0] = 0.0; cb[0] = 1.0; sb[0] = 0.0;
#pragma unroll
  for (int m = 1; m <= L; ++m) {
    ca[m] = ca[m - 1] * cp + sa[m - 1] * sp;
    sa[m] = sa[m - 1] * cp - ca[m - 1] * sp;
    cb[m] = cb[m - 1] * ct + sb[m - 1] * st;
    sb[m] = sb[m - 1] * ct - cb[m - 1] * st;
  }
  double M[d * d];
  // N = J^T Z(-phi)  (column mixing)
  for (int i = 0; i < d; ++i) {
    M[i * d + L] = __ldg(J + L * d + i);
#pragma unroll
    for (int m = 1; m <= L; ++m) {
      const double jp = __ldg(J + (L + m) * d + i), jm = __ldg(J + (L - m) * d + i);
      M[i * d + L + m] = jp * ca[m] + jm * sa[m];
      M[i * d + L - m] = -jp * sa[m] + jm * ca[m];
    }
  }
  // M2 = Z(-theta) N  (row mixing)
#pragma unroll
  for (int m = 1; m <= L; ++m)
    for (int j = 0; j < d; ++j) {
      const double u = M[(L + m) * d + j], v = M[(L - m) * d + j];
      M[(L + m) * d + j] = cb[m] * u - sb[m] * v;
      M[(L - m) * d + j] = sb[m] * u + cb[m] * v;
    }
  // D = J M2
  for (int r = 0; r < d; ++r) {
    double acc[d];
#pragma unroll
    for (int j = 0; j < d; ++j) acc[j] = 0.0;
    for (int i = 0; i < d; ++i) {
      const double jr = __ldg(J + r * d + i);
#pragma unroll
      for (int j = 0; j < d; ++j) acc[j] = fma(jr, M[i * d + j], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < d; ++j) out[r * d + j] = (float)acc[j];
  }
}

__global__ void __launch_bounds__(128) wigner_kernel(const __grid_constant__ WigArgs a) {
  const int64_t e = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (e >= a.n_edges) return;
  const int l = blockIdx.y;
  double x = a.vec[3 * e], y = a.vec[3 * e + 1], z = a.vec[3 * e + 2];
  const double n = sqrt(x * x + y * y + z * z);
  x /= n; y /= n; z /= n;
  const double rho = sqrt(x * x + y * y);
  const bool ok = rho > 1e-30;
  const double cp = ok ? x / rho : 1.0, sp = ok ? y / rho : 0.0;
  float* out = a.dw + e * a.dstride + a.doff[l];
  const double* J = a.J + a.doff[l];
  switch (l) {
    case 0: out[0] = 1.f; break;
    case 1: wigner_l<1>(J, cp, sp, z, rho, out); break;
    case 2: wigner_l<2>(J, cp, sp, z, rho, out); break;
    case 3: wigner_l<3>(J, cp, sp, z, rho, out); break;
    case 4: wigner_l<4>(J, cp, sp, z, rho, out); break;
    case 5: wigner_l<5>(J, cp, sp, z, rho, out); break;
    default: wigner_l<6>(J, cp, sp, z, rho, out); break;
  }
}

// ================================================================================================ rotate + pack
struct RpArgs {
  const hgb_rot_block_t* blocks;
  int n_blocks, blocks_per_cta;
  int tile_stride, dstride;
  int doff[12];
  const float* src[4];
  const int64_t* src_rows[4];
  int src_dim[4];
  const float* dw;
  int64_t e_lo, n_chunk;
  float* xp;
};

template <int L1>
__device__ __forceinline__ void rotpack_block(const RpArgs& a, const hgb_rot_block_t& b, int tile, int z, int64_t e, bool live) {
  constexpr int d1 = 2 * L1 + 1;
  const int K = b.nsrc * b.mul, kpad = b.kpad;
  const float* r0 = nullptr;
  const float* r1 = nullptr;
  const float* Dz = nullptr;
  if (live) {
    const int s0 = b.src0, s1 = b.src0 + b.nsrc - 1;
    const int64_t row0 = a.src_rows[s0] ? a.src_rows[s0][e] : e;
    const int64_t row1 = a.src_rows[s1] ? a.src_rows[s1][e] : e;
    r0 = a.src[s0] + row0 * a.src_dim[s0] + b.in_off;
    r1 = a.src[s1] + row1 * a.src_dim[s1] + b.in_off;
    Dz = a.dw + e * a.dstride + a.doff[L1];
  }
  float* xo = a.xp + (size_t)tile * a.tile_stride + b.xoff + z * 4;
  const size_t per_m = (size_t)2 * kpad * TILE;
  for (int q = 0; q < (kpad >> 2); ++q) {
    float x[4][d1];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int u = 4 * q + c;
      const bool okc = live && u < K;
      const bool second = u >= b.mul;
      const float* p = okc ? ((second ? r1 : r0) + (u - (second ? b.mul : 0)) * d1) : nullptr;
#pragma unroll
      for (int i = 0; i < d1; ++i) x[c][i] = okc ? __ldg(p + i) : 0.f;
    }
    const int chunk = (4 * q) / KC, ul = (4 * q) - chunk * KC, kc = min(KC, kpad - chunk * KC);
    float* base = xo + (size_t)chunk * 2 * KC * TILE + (size_t)(ul >> 2) * (TILE * 4);
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
      float4 h, l;
      tc::split_tf32(v[0], h.x, l.x); tc::split_tf32(v[1], h.y, l.y);
      tc::split_tf32(v[2], h.z, l.z); tc::split_tf32(v[3], h.w, l.w);
      float* dst = base + m * per_m;
      *reinterpret_cast<float4*>(dst) = h;
      *reinterpret_cast<float4*>(dst + (size_t)kc * TILE) = l;
    }
  }
}

__global__ void __launch_bounds__(TILE) rotate_pack_kernel(const __grid_constant__ RpArgs a) {
  const int tile = blockIdx.x, z = threadIdx.x;
  const int64_t el = (int64_t)tile * TILE + z;
  const bool live = el < a.n_chunk;
  const int64_t e = a.e_lo + el;
  const int b0 = blockIdx.y * a.blocks_per_cta, b1 = min(a.n_blocks, b0 + a.blocks_per_cta);
  for (int bi = b0; bi < b1; ++bi) {
    const hgb_rot_block_t b = a.blocks[bi];
    switch (b.l1) {
      case 0: rotpack_block<0>(a, b, tile, z, e, live); break;
      case 1: rotpack_block<1>(a, b, tile, z, e, live); break;
      case 2: rotpack_block<2>(a, b, tile, z, e, live); break;
      case 3: rotpack_block<3>(a, b, tile, z, e, live); break;
      case 4: rotpack_block<4>(a, b, tile, z, e, live); break;
      case 5: rotpack_block<5>(a, b, tile, z, e, live); break;
      default: rotpack_block<6>(a, b, tile, z, e, live); break;
    }
  }
}

// ================================================================================================= message kernel
struct RotArgs {
  hgb_msgpack_plan plan;
  const hgb_rot_step_t* steps;
  int step_begin[33];
  const float* xp;
  int tile_stride;
  const float* dw;
  int dstride;
  int doff[12];
  const float* g;       // [n_branches][tile][gstride][128]: tile-major radial gate of this chunk
  int gstride;
  int64_t e_lo, n_chunk;
  float* out;
  const int64_t* out_index;
  int n_slots;
  int slot[32];
  int dbl;              // 1: the gated-product lo block and the per-step GEMM2 accumulator are double buffered
};

__device__ __forceinline__ bool elect_one() {
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
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, uint32_t& r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n" : "=r"(r) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait1(uint32_t& r) { asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r)::"memory"); }
__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};\n" ::"r"(taddr), "r"(r) : "memory");
}

// C[z][w][k] = sum_m3 D^{l3}_z[m3][k] C'[z][m3][w]; thread = edge z = TMEM lane; C'[m3][w] is TMEM column tc0 + m3 * mul + w
// (exact stride, written by the accumulate phase); bit m3 of cmask: the component received contributions.
template <int L3>
__device__ __forceinline__ void rot_epilogue(uint32_t tc0, int mul, uint32_t cmask, const float* __restrict__ Dz, float* __restrict__ op,
                                             bool live, bool atomic) {
  constexpr int d3 = 2 * L3 + 1;
  for (int c0 = 0; c0 < mul; c0 += 4) {   // warp-uniform
    uint32_t c[d3][4];
#pragma unroll
    for (int m = 0; m < d3; ++m)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        c[m][j] = 0u;
        if (((cmask >> m) & 1u) && c0 + j < mul) tmem_ld1(tc0 + m * mul + c0 + j, c[m][j]);
      }
#pragma unroll
    for (int m = 0; m < d3; ++m)
#pragma unroll
      for (int j = 0; j < 4; ++j) tmem_ld_wait1(c[m][j]);
    if (!live) continue;
#pragma unroll
    for (int k = 0; k < d3; ++k) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int m = 0; m < d3; ++m) {
        const float dmk = (L3 == 0) ? 1.f : __ldg(Dz + m * d3 + k);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = fmaf(dmk, __uint_as_float(c[m][j]), acc[j]);
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

// One CTA = one (tile of 128 edges, output slot).  Steps are ordered by output component m3; every step is
//   GEMM1  B[n&1]  = X'_{m1} W_p            warp 4 (A / W chunks from the TMA ring filled by warp 5)
//   gate   B <- (B * scale * g) hi, GL <- lo warps 0-3, thread = edge = TMEM lane
//   GEMM2  S       = (B.g) L'_p             warp 6 (fresh accumulator: accumulate = 0)
//   acc   += S                              warps 0-3, fp32 round-to-nearest in registers
// and at the end of an m3 group the registers go to C'[m3] in TMEM.  The tensor core's own accumulation chains stay
// short (K/8 and mp/8 instructions x 3): its accumulator adds truncate, and a chain over all ~50 paths of a slot
// cost 4x the fp32 rounding error of the whole network (scripts/error_budget.py, profiles/r01o_error_budget.log).
// RW = padded-multiplicity capacity of the class (registers, stage size); ty.mpad <= RW is the MMA N.
// warps 0-3 gate / accumulate, 4 GEMM1, 5 TMA (A chunks), 6 GEMM2, 7 TMA (W chunks), 8 TMA (L' images) + gate L2 prefetch.
// Measured on tbg_m28 (profiles/README.md r03): one producer warp 871 ms, A | W+L' 854 ms, A | W | L' 838 ms, L' issued by the
// GEMM2 warp 893 ms (the ~400-cycle issue lands on the gate -> GEMM2 -> accumulate chain).
// One cp.async.bulk costs its issuing thread ~400 cycles whatever its size, but the cost does not add up across warps
// (profiles/r02v_tma_probe.txt): with a single producer thread the 3-5 bulk copies of a step paced the kernel.
constexpr int NTHR2 = 288;

// barrier helpers on precomputed 32-bit shared addresses (no generic -> shared conversion per use)
__device__ __forceinline__ void wait_a(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .u32 n;\n\t"
      "mov.u32 n, 0;\n\t"
      "mov.u32 %0, 1;\n"
      "HGB_ROT_WAIT_A:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "@p bra HGB_ROT_DONE_A;\n\t"
      "add.u32 n, n, 1;\n\t"
      "setp.lt.u32 p, n, 0x4000000;\n\t"
      "@p bra HGB_ROT_WAIT_A;\n\t"
      "mov.u32 %0, 0;\n"
      "HGB_ROT_DONE_A:\n\t"
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
__device__ __forceinline__ void arrive_a(uint32_t addr) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory"); }
__device__ __forceinline__ void expect_tx_a(uint32_t addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_a(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void commit_a(uint32_t addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(addr) : "memory");
}

template <int RW, int NST>
__global__ void __launch_bounds__(NTHR2, (RW == 64 ? 1 : 2)) msgpack_rot_kernel(const __grid_constant__ RotArgs a) {
  constexpr int STG = 2 * KC * TILE + 2 * RW * KC;   // floats per ring stage: A chunk (hi | lo) + W chunk (hi | lo)
  extern __shared__ __align__(128) float smem[];
  // barriers: full[NST] | empty[NST] | lfull[2] | bfull[2] | gfull[2] | s2done[2]
  __shared__ uint64_t bars[2 * NST + 8];
  __shared__ uint32_t tmem_slot;
  const uint32_t bar0 = tc::smem_u32(bars);
  const uint32_t B_FULL = bar0, B_EMPTY = bar0 + 8 * NST, B_LFULL = bar0 + 16 * NST, B_BFULL = B_LFULL + 16, B_GFULL = B_LFULL + 32,
                 B_S2 = B_LFULL + 48;
  const uint32_t stage0 = tc::smem_u32(smem);                   // ring stages, STG floats each
  const uint32_t sl0 = stage0 + (uint32_t)(NST * STG) * 4u;     // 2 x (hi | lo) L' images, 2 RW^2 floats each

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x / a.n_slots;
  const int t = a.slot[blockIdx.x - tile * a.n_slots];
  const hgb_type_t ty = a.plan.types[t];
  const int d3 = 2 * ty.l + 1, mp = ty.mpad, mul = ty.mul;
  const int sb = a.step_begin[t], se = a.step_begin[t + 1];
  const int dbl = a.dbl;
  // TMEM columns: B0 | B1 | GL0 (| GL1) | S0 (| S1) | C' (d3 x mul, exact stride)
  const uint32_t TB0 = 0, TGL0 = 2 * mp, TS0 = (uint32_t)((3 + dbl) * mp), TC = (uint32_t)((4 + 2 * dbl) * mp);
  uint32_t ncols = 32;
  while (ncols < TC + (uint32_t)(d3 * mul)) ncols <<= 1;

  if (tid == 0) {
    for (int i = 0; i < 2 * NST + 8; ++i) tc::mbar_init(&bars[i], (i >= 2 * NST + 4 && i < 2 * NST + 6) ? 4 : (i < NST ? 2 : 1));   // gfull: one arrival per gate warp; full: A + W producers
    tc::mbar_fence_init();
  }
  if (warp == 4) tmem_alloc_dyn(&tmem_slot, ncols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const float* __restrict__ wbuf = a.plan.wbuf;
  const uint32_t idesc = tc::idesc_tf32_m128(mp);
  const uint32_t dhi = tc::smem_desc_hi(128);
  const uint32_t lbo_a = TILE * 16, lbo_n = (uint32_t)mp * 16;
  const uint32_t astep = (2 * lbo_a) >> 4, bstep = (2 * lbo_n) >> 4;

  if (warp == 5 || warp == 7 || warp == 8) {
    // =============================== TMA producers: A chunks | W chunks | L' images + gate prefetch ===============================
    if (lane == 0) {
      const float* xt = a.xp + (size_t)tile * a.tile_stride;
      if (warp == 8) {
        // gate block of a step: mul columns x 128 edges, contiguous in the tile-major gate tensor -> pulled into L2 GPF steps
        // before the epilogue warps load it (the gate tensor of a chunk is GBs, written by the pre-pass: not L2 resident)
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
        const uint32_t lbytes = (uint32_t)(2 * mp * mp) * 4u;
        for (int si = sb; si < se; ++si, ++n) {
          const hgb_rot_step_t st = a.steps[si];
          prefetch_gate(si + GPF);
          const int lb = n & 1;
          if (n >= 2) wait_a(B_S2 + 8 * lb, (uint32_t)(((n >> 1) - 1) & 1));   // GEMM2(n-2) has read the L' buffer
          expect_tx_a(B_LFULL + 8 * lb, lbytes);
          bulk_g2s_a(sl0 + (uint32_t)(lb * 2 * RW * RW) * 4u, wbuf + st.lf_off, lbytes, B_LFULL + 8 * lb);
        }
      } else {
        const bool isA = warp == 5;
        int c_all = 0;
        for (int si = sb; si < se; ++si) {
          const hgb_rot_step_t st = a.steps[si];
          const int kpad = st.kpad;
          for (int u0 = 0, c = 0; u0 < kpad; u0 += KC, ++c, ++c_all) {
            const int kc = min(KC, kpad - u0), s = c_all % NST;
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
    int kpad = (sb < se) ? a.steps[sb].kpad : 0;
    for (int si = sb; si < se; ++si, ++n) {
      const int kpad_next = (si + 1 < se) ? a.steps[si + 1].kpad : 0;   // in flight while this step is issued
      if (n >= 2) warp_wait_a(B_S2 + 8 * (n & 1), (uint32_t)(((n >> 1) - 1) & 1));   // GEMM2(n-2) has read B[n&1]
      const uint32_t dcol = tmem + TB0 + (uint32_t)((n & 1) * mp);
      for (int u0 = 0, c = 0; u0 < kpad; u0 += KC, ++c, ++c_all) {
        const int kc = min(KC, kpad - u0), s = c_all % NST;
        warp_wait_a(B_FULL + 8 * s, (uint32_t)((c_all / NST) & 1));
        tc::fence_after_sync();
        if (elect_one()) {
          const uint32_t sa = stage0 + (uint32_t)(s * STG) * 4u;
          const uint32_t ah = tc::smem_desc_lo(sa, lbo_a), al = ah + (((uint32_t)kc * TILE * 4) >> 4);
          const uint32_t wh = tc::smem_desc_lo(sa + 2 * KC * TILE * 4, lbo_n), wl = wh + (((uint32_t)mp * kc * 4) >> 4);
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
        const uint32_t lh = tc::smem_desc_lo(sl0 + (uint32_t)((n & 1) * 2 * RW * RW) * 4u, lbo_n), ll = lh + (((uint32_t)mp * mp * 4) >> 4);
        for (int k8 = 0; k8 < (mp >> 3); ++k8) {
          const uint64_t bh = tc::desc64(lh + k8 * bstep, dhi), bl = tc::desc64(ll + k8 * bstep, dhi);
          tc::mma_tf32_ts(sc, gl + k8 * 8, bh, idesc, (uint32_t)(k8 > 0));
          tc::mma_tf32_ts(sc, bq + k8 * 8, bl, idesc, 1);
          tc::mma_tf32_ts(sc, bq + k8 * 8, bh, idesc, 1);
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
    const size_t g_bstride = (size_t)((a.n_chunk + TILE - 1) / TILE) * a.gstride * TILE;   // floats per branch
    const float* grow = a.g + (size_t)tile * a.gstride * TILE + (live ? tid : 0);            // column c of this edge: grow[c * TILE]
    float gv[RW], acc[RW];
#pragma unroll
    for (int j = 0; j < RW; ++j) { gv[j] = 0.f; acc[j] = 0.f; }
    float gA = 0.f, gB = 0.f;   // gate factor of the step whose values sit in gv: B *= gv * gA + gB
    uint32_t cmask = 0;         // output components some step writes
    // One step record = two 16-byte words (hgb_rot_step_t): {a_off, w_off, lf_off, g_off} {scale, kpad|kind|branch, m3|flags|pad, pad2}
    const uint4* steps4 = reinterpret_cast<const uint4*>(a.steps);
    // gv = g_p[z, :] of a step: raw predicated loads straight into the registers, nothing here consumes them, so they
    // stay in flight while the accumulate phase runs.  Un-gated steps (branch < 0, the direct Linear of the edge
    // features) load the same way (any valid column) and use the factor (gA, gB) = (0, scale) instead of (scale, 0).
    // Padding rows of the last tile read row 0 of the tile: their B rows are zero and are never stored.
    auto load_gate = [&](const uint4& w0, const uint4& w1) {
      const float sc = __uint_as_float(w1.x);
      const int br = (int)(int8_t)(w1.y >> 24);
      gA = (br < 0) ? 0.f : sc;
      gB = (br < 0) ? sc : 0.f;
      const float* gp = grow + (size_t)max(br, 0) * g_bstride + (size_t)((br < 0) ? 0 : (int)w0.w) * TILE;
#pragma unroll
      for (int j = 0; j < RW; ++j)
        if (j < mul) gv[j] = __ldg(gp + j * TILE);   // warp-uniform predicate; a warp reads 128 contiguous bytes
    };
    // acc += S of step (n, flags, m3); at the end of an m3 group the registers move to C'[m3]
    auto accumulate = [&](int n, int flags, int m3) {
      const int gi = dbl ? (n & 1) : 0;
      warp_wait_a(B_S2 + 8 * (n & 1), (uint32_t)((n >> 1) & 1));
      tc::fence_after_sync();
      const uint32_t sc = tmem + lane_base + TS0 + (uint32_t)(gi * mp);
#pragma unroll
      for (int c0 = 0; c0 < RW; c0 += 8) {
        if (c0 < mp) {   // warp-uniform
          uint32_t rs[8];
          tc::tmem_ld8(sc + c0, rs);
          tc::tmem_ld_wait8(rs);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[c0 + j] += __uint_as_float(rs[j]);
        }
      }
      if (flags & 4) {
        const uint32_t cc = tmem + lane_base + TC + (uint32_t)(m3 * mul);
#pragma unroll
        for (int j = 0; j < RW; ++j) {
          if (j < mul) tmem_st1(cc + j, __float_as_uint(acc[j]));   // warp-uniform predicate
          acc[j] = 0.f;
        }
        tc::tmem_st_wait();
      }
      tc::fence_before_sync();   // the S block may be overwritten by a later GEMM2 once gfull of a later step is signalled
    };
    auto gate = [&](int n) {
      const int gi = dbl ? (n & 1) : 0;
      const float fa = gA, fb = gB;
      warp_wait_a(B_BFULL + 8 * (n & 1), (uint32_t)((n >> 1) & 1));
      tc::fence_after_sync();
      const uint32_t bq = tmem + lane_base + TB0 + (uint32_t)((n & 1) * mp);
      const uint32_t gl = tmem + lane_base + TGL0 + (uint32_t)(gi * mp);
#pragma unroll
      for (int c0 = 0; c0 < RW; c0 += 8) {
        if (c0 < mp) {   // warp-uniform
          uint32_t rb[8], hi[8], lo[8];
          tc::tmem_ld8(bq + c0, rb);
          tc::tmem_ld_wait8(rb);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float h, l;
            tc::split_tf32(__uint_as_float(rb[j]) * fmaf(gv[c0 + j], fa, fb), h, l);
            hi[j] = __float_as_uint(h); lo[j] = __float_as_uint(l);
          }
          tc::tmem_st8(bq + c0, hi);
          tc::tmem_st8(gl + c0, lo);
        }
      }
      tc::tmem_st_wait();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) arrive_a(B_GFULL + 8 * gi);
    };
    int n = 0, pflags = 0, pm3 = 0;
    uint32_t cur_fm = 0;   // m3 | flags << 8 of the current step
    if (se > sb) {
      const uint4 w0 = __ldg(steps4 + 2 * sb), w1 = __ldg(steps4 + 2 * sb + 1);
      cur_fm = w1.z;
      load_gate(w0, w1);
    }
    for (int si = sb; si < se; ++si, ++n) {
      uint4 n0 = make_uint4(0, 0, 0, 0), n1 = n0;
      const bool more = si + 1 < se;
      if (more) { n0 = __ldg(steps4 + 2 * (si + 1)); n1 = __ldg(steps4 + 2 * (si + 1) + 1); }   // in flight during the gate phase
      const int m3 = (int)(cur_fm & 0xff), flags = (int)((cur_fm >> 8) & 0xff);
      cmask |= 1u << m3;
      if (!dbl && n > 0) accumulate(n - 1, pflags, pm3);
      gate(n);
      if (more) load_gate(n0, n1);   // next step's gate values travel while the accumulate phase runs
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

template <int RW, int NST>
constexpr size_t rot_smem_bytes() { return (size_t)(NST * (2 * KC * TILE + 2 * RW * KC) + 2 * 2 * RW * RW) * sizeof(float); }

}  // namespace rot
