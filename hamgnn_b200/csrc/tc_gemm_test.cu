// Self-test of the tcgen05 3xTF32 building block: C[128 x N] = A[128 x K] . B[K x N] per CTA, fp32 in/out,
// operands staged in the interleaved K-major layout of tc_common.cuh, accumulator in TMEM.  Exposed through the
// C ABI (hgb_tc_gemm_selftest) so the GPU test-suite can validate descriptors/layouts against a plain matmul
// before the fused message kernel relies on them.
#include "hgb_common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TM = 128;

// MODE 0: A from shared memory (SS).  MODE 1: A staged into TMEM with tcgen05.st (TS), K <= 64.
template <int NPAD, int MODE>
__global__ void __launch_bounds__(128) tc_gemm_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                      float* __restrict__ C, int K, int N) {
  extern __shared__ __align__(128) float smem[];
  // layout: Ahi[K/4][128][4], Alo, Bhi[K/4][NPAD][4], Blo
  float* Ahi = smem;
  float* Alo = Ahi + (size_t)K * TM;
  float* Bhi = Alo + (size_t)K * TM;
  float* Blo = Bhi + (size_t)K * NPAD;
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const float* Ab = A + (size_t)blockIdx.x * TM * K;
  float* Cb = C + (size_t)blockIdx.x * TM * N;

  if (tid == 0) {
    tc::mbar_init(&mbar, 1);
    tc::mbar_fence_init();
  }
  constexpr int TCOLS = (MODE == 0) ? (NPAD <= 32 ? 32 : 64) : 256;  // power of two >= 32  // TS: D at col 0, A_hi at 64, A_lo at 128
  if (warp == 0) tc::tmem_alloc<TCOLS>(&tmem_base);
  // stage + split A: element (r, k) -> (k/4)*(128*4) + r*4 + k%4
  for (int idx = tid; idx < TM * K; idx += 128) {
    const int r = idx / K, k = idx - r * K;
    float hi, lo;
    tc::split_tf32(Ab[idx], hi, lo);
    const int o = (k >> 2) * (TM * 4) + r * 4 + (k & 3);
    Ahi[o] = hi; Alo[o] = lo;
  }
  // stage + split B (global [K][N]) as [N][K] K-major: (n, k) -> (k/4)*(NPAD*4) + n*4 + k%4 ; zero padding
  for (int idx = tid; idx < NPAD * K; idx += 128) {
    const int k = idx / NPAD, n = idx - k * NPAD;
    float hi = 0.f, lo = 0.f;
    if (n < N) tc::split_tf32(B[(size_t)k * N + n], hi, lo);
    const int o = (k >> 2) * (NPAD * 4) + n * 4 + (k & 3);
    Bhi[o] = hi; Blo[o] = lo;
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base;
  if (MODE == 1) {
    // every thread owns TMEM lane (= row) warp*32 + lane: write its A row, split, 8 columns at a time
    const int rowa = warp * 32 + (tid & 31);
    for (int c0 = 0; c0 < K; c0 += 8) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float h, l;
        tc::split_tf32(Ab[(size_t)rowa * K + c0 + j], h, l);
        hi[j] = __float_as_uint(h); lo[j] = __float_as_uint(l);
      }
      tc::tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + 64 + c0, hi);
      tc::tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + 128 + c0, lo);
    }
    tc::tmem_st_wait();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
  }
  if (tid == 0 && MODE == 1) {
    constexpr uint32_t idesc = tc::idesc_tf32_m128(NPAD);
    const uint32_t b_hi = tc::smem_u32(Bhi), b_lo = tc::smem_u32(Blo);
    const uint32_t lbo_b = NPAD * 16, sbo = 128;
    for (int k8 = 0; k8 < K / 8; ++k8) {
      const uint32_t ob = k8 * 2 * lbo_b;
      tc::mma_tf32_ts(tmem, tmem + 128 + k8 * 8, tc::smem_desc(b_hi + ob, lbo_b, sbo), idesc, k8 > 0);
      tc::mma_tf32_ts(tmem, tmem + 64 + k8 * 8, tc::smem_desc(b_lo + ob, lbo_b, sbo), idesc, 1);
      tc::mma_tf32_ts(tmem, tmem + 64 + k8 * 8, tc::smem_desc(b_hi + ob, lbo_b, sbo), idesc, 1);
    }
    tc::mma_commit(&mbar);
  }
  if (tid == 0 && MODE == 0) {
    constexpr uint32_t idesc = tc::idesc_tf32_m128(NPAD);
    const uint32_t a_hi = tc::smem_u32(Ahi), a_lo = tc::smem_u32(Alo), b_hi = tc::smem_u32(Bhi), b_lo = tc::smem_u32(Blo);
    const uint32_t lbo_a = TM * 16, lbo_b = NPAD * 16, sbo = 128;
    for (int k8 = 0; k8 < K / 8; ++k8) {
      const uint32_t oa = k8 * 2 * lbo_a, ob = k8 * 2 * lbo_b;
      tc::mma_tf32(tmem, tc::smem_desc(a_lo + oa, lbo_a, sbo), tc::smem_desc(b_hi + ob, lbo_b, sbo), idesc, k8 > 0);
      tc::mma_tf32(tmem, tc::smem_desc(a_hi + oa, lbo_a, sbo), tc::smem_desc(b_lo + ob, lbo_b, sbo), idesc, 1);
      tc::mma_tf32(tmem, tc::smem_desc(a_hi + oa, lbo_a, sbo), tc::smem_desc(b_hi + ob, lbo_b, sbo), idesc, 1);
    }
    tc::mma_commit(&mbar);
  }
  tc::mbar_wait(&mbar, 0);
  tc::fence_after_sync();
  const int row = warp * 32 + (tid & 31);
  for (int c0 = 0; c0 < NPAD; c0 += 8) {
    uint32_t r[8];
    tc::tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
    tc::tmem_ld_wait8(r);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (c0 + j < N) Cb[(size_t)row * N + c0 + j] = __uint_as_float(r[j]);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<TCOLS>(tmem);
}

template <int NPAD>
int launch(const float* A, const float* B, float* C, int tiles, int K, int N, int mode, cudaStream_t st) {
  const size_t smem = (size_t)(2 * K * TM + 2 * K * NPAD) * sizeof(float);
  HGB_CHECK_ARG(smem <= 200 * 1024, "hgb_tc_gemm_selftest: K=%d too large for the staging buffers", K);
  if (mode == 0) {
    HGB_CUDA_OK(cudaFuncSetAttribute(tc_gemm_kernel<NPAD, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_gemm_kernel<NPAD, 0><<<tiles, 128, smem, st>>>(A, B, C, K, N);
  } else {
    HGB_CHECK_ARG(K <= 64, "hgb_tc_gemm_selftest: TMEM-A mode supports K <= 64");
    HGB_CUDA_OK(cudaFuncSetAttribute(tc_gemm_kernel<NPAD, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_gemm_kernel<NPAD, 1><<<tiles, 128, smem, st>>>(A, B, C, K, N);
  }
  HGB_LAUNCH_OK("tc_gemm_kernel");
  return 0;
}

}  // namespace

// C[t] (128 x N) = A[t] (128 x K) . B (K x N) for t < tiles.  K % 8 == 0, N <= 64.
extern "C" int hgb_tc_gemm_selftest(const float* A, const float* B, float* C, int32_t tiles, int32_t K, int32_t N,
                                    int32_t a_from_tmem, void* stream) {
  HGB_DEVICE_GUARD(C);
  HGB_CHECK_ARG(A && B && C, "hgb_tc_gemm_selftest: NULL argument");
  HGB_CHECK_ARG(K > 0 && K % 8 == 0, "hgb_tc_gemm_selftest: K=%d must be a positive multiple of 8", K);
  HGB_CHECK_ARG(N > 0 && N <= 64, "hgb_tc_gemm_selftest: N=%d out of range (1..64)", N);
  cudaStream_t st = (cudaStream_t)stream;
  if (N <= 16) return launch<16>(A, B, C, tiles, K, N, a_from_tmem, st);
  if (N <= 32) return launch<32>(A, B, C, tiles, K, N, a_from_tmem, st);
  if (N <= 48) return launch<48>(A, B, C, tiles, K, N, a_from_tmem, st);
  return launch<64>(A, B, C, tiles, K, N, a_from_tmem, st);
}
