// SIMT helpers shared by the tensor-core message kernels (msgpack_tc.cu, msgpack_tcg.cu): asynchronous copies,
// division-free index arithmetic and the A-operand generator (per-edge d1 x d3 CG contraction + tf32 split into the
// UMMA interleaved K-major image).
#pragma once
#include "hgb_common.cuh"
#include "tc_common.cuh"

namespace tcmsg {

constexpr int ROWS = 128;  // MMA M: accumulator rows per CTA tile

// asynchronous global -> shared copies (LDGSTS): issued by all threads, completed by cp_async_wait_all()
__device__ __forceinline__ void cp_async16(float* dst, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tc::smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(float* dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(tc::smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// group form: commit what this thread has issued so far; wait until at most N of its newest groups are in flight
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
template <int NT>
__device__ __forceinline__ void copy_f4(float* dst, const float* __restrict__ src, int nfloats) {
  for (int i = threadIdx.x; i < (nfloats >> 2); i += NT) cp_async16(dst + 4 * i, src + 4 * i);
}

// n / d for n < 2^16 by multiply-high; d == 1 is encoded as magic 0
__device__ __forceinline__ uint32_t fdiv_magic(uint32_t d) { return d == 1u ? 0u : 0xFFFFFFFFu / d + 1u; }
// magic of 2l+1 without a division (l <= 8)
__device__ __forceinline__ uint32_t fdiv_magic_odd(int l) {
  constexpr uint32_t tab[9] = {0u, 0xFFFFFFFFu / 3 + 1, 0xFFFFFFFFu / 5 + 1, 0xFFFFFFFFu / 7 + 1, 0xFFFFFFFFu / 9 + 1,
                               0xFFFFFFFFu / 11 + 1, 0xFFFFFFFFu / 13 + 1, 0xFFFFFFFFu / 15 + 1, 0xFFFFFFFFu / 17 + 1};
  return tab[l];
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, uint32_t magic) { return magic ? __umulhi(n, magic) : n; }

// A[(z,k)][u] for the 4-channel groups of a staged sub-chunk, written hi/lo into the interleaved K-major image.
// sX is channel-minor ([z][i][c], 16-byte aligned quads), T_z rows are padded to D3P = round4(D3) floats, both with
// (row stride / 4) odd so that the 128-bit loads of a quarter warp (lanes = consecutive z) hit distinct banks.
template <int NT, int D3>
__device__ __forceinline__ void agen_tc(float* __restrict__ sA, int lo_off, const float* __restrict__ sX, int ldx, int cc,
                                        const float* __restrict__ sT, int ldt, int d1, int nz, uint32_t mz, int slab0,
                                        int nquad) {
  constexpr int D3P = (D3 + 3) & ~3;
  for (int item = threadIdx.x; item < nz * nquad; item += NT) {
    const int q = (int)fdiv((uint32_t)item, mz), z = item - q * nz;
    float a[4][D3];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int k = 0; k < D3; ++k) a[c][k] = 0.f;
    const float* xz = sX + (size_t)z * ldx + q * 4;
    const float* tz = sT + (size_t)z * ldt;
    for (int i = 0; i < d1; ++i) {
      const float4 x = *reinterpret_cast<const float4*>(xz + i * cc);
#pragma unroll
      for (int k4 = 0; k4 < D3P; k4 += 4) {
        const float4 t = *reinterpret_cast<const float4*>(tz + i * D3P + k4);
        const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (k4 + j < D3) {
            a[0][k4 + j] = fmaf(x.x, tv[j], a[0][k4 + j]);
            a[1][k4 + j] = fmaf(x.y, tv[j], a[1][k4 + j]);
            a[2][k4 + j] = fmaf(x.z, tv[j], a[2][k4 + j]);
            a[3][k4 + j] = fmaf(x.w, tv[j], a[3][k4 + j]);
          }
      }
    }
    float* hi = sA + (size_t)(slab0 + q) * (ROWS * 4) + (size_t)z * D3 * 4;
    float* lo = hi + lo_off;
#pragma unroll
    for (int k = 0; k < D3; ++k) {
      float4 h, l;
      tc::split_tf32(a[0][k], h.x, l.x);
      tc::split_tf32(a[1][k], h.y, l.y);
      tc::split_tf32(a[2][k], h.z, l.z);
      tc::split_tf32(a[3][k], h.w, l.w);
      *reinterpret_cast<float4*>(hi + k * 4) = h;
      *reinterpret_cast<float4*>(lo + k * 4) = l;
    }
  }
}

template <int NT>
__device__ __forceinline__ void agen_tc_dispatch(int d3, float* sA, int lo_off, const float* sX, int ldx, int cc, const float* sT,
                                                 int ldt, int d1, int nz, uint32_t mz, int slab0, int nquad) {
  switch (d3) {
    case 1: agen_tc<NT, 1>(sA, lo_off, sX, ldx, cc, sT, ldt, d1, nz, mz, slab0, nquad); break;
    case 3: agen_tc<NT, 3>(sA, lo_off, sX, ldx, cc, sT, ldt, d1, nz, mz, slab0, nquad); break;
    case 5: agen_tc<NT, 5>(sA, lo_off, sX, ldx, cc, sT, ldt, d1, nz, mz, slab0, nquad); break;
    case 7: agen_tc<NT, 7>(sA, lo_off, sX, ldx, cc, sT, ldt, d1, nz, mz, slab0, nquad); break;
    case 9: agen_tc<NT, 9>(sA, lo_off, sX, ldx, cc, sT, ldt, d1, nz, mz, slab0, nquad); break;
    case 11: agen_tc<NT, 11>(sA, lo_off, sX, ldx, cc, sT, ldt, d1, nz, mz, slab0, nquad); break;
    case 13: agen_tc<NT, 13>(sA, lo_off, sX, ldx, cc, sT, ldt, d1, nz, mz, slab0, nquad); break;
    case 15: agen_tc<NT, 15>(sA, lo_off, sX, ldx, cc, sT, ldt, d1, nz, mz, slab0, nquad); break;
    default: agen_tc<NT, 17>(sA, lo_off, sX, ldx, cc, sT, ldt, d1, nz, mz, slab0, nquad); break;
  }
}


}  // namespace tcmsg
