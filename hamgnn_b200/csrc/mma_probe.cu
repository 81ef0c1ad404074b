// Micro-probe of tcgen05.mma.kind::tf32 issue / completion cost on sm_100a (diagnosis tool, not on the product path):
// one CTA issues `count` MMAs (M = 128, N = n, K = 8) from one thread, round-robin over `ndest` accumulators,
// A from shared memory (ts = 0) or from TMEM (ts = 1), and reports the cycles until the commit barrier fires.
#include "hgb_common.cuh"
#include "tc_common.cuh"

namespace {
struct ProbeArgs {
  int n, count, ndest, ts, same_ab;
  long long* out;   // [0] issue cycles, [1] total cycles until completion
};

__global__ void __launch_bounds__(128, 1) mma_probe_kernel(const ProbeArgs a) {
  extern __shared__ __align__(128) float smem[];   // A: 2 x [2 slabs][128][4], B: [2 slabs][256][4] x 2
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ int flag;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 16384; i += 128) smem[i] = 0.f;
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); flag = 0; }
  if (warp == 0) tc::tmem_alloc<512>(&tmem_slot);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = tc::idesc_tf32_m128(a.n);
    const uint32_t dhi = tc::smem_desc_hi(128);
    const uint32_t sa = tc::smem_u32(smem), sb = sa + 16384;
    const uint32_t ah = tc::smem_desc_lo(sa, 128 * 16), bh = tc::smem_desc_lo(sb, (uint32_t)a.n * 16);
    const long long t0 = clock64();
    const uint32_t dmask = (uint32_t)(a.ndest - 1);   // ndest is a power of two
    const uint64_t da0 = tc::desc64(ah, dhi), db0 = tc::desc64(bh, dhi);
    const uint64_t da1 = tc::desc64(ah + (a.same_ab ? 0u : 512u), dhi), db1 = tc::desc64(bh + (a.same_ab ? 0u : 512u), dhi);
    if (a.ts) {
#pragma unroll 4
      for (int i = 0; i < a.count; i += 2) {
        tc::mma_tf32_ts(tmem + ((uint32_t)i & dmask) * (uint32_t)a.n, tmem + 384, db0, idesc, (uint32_t)(i >= a.ndest));
        tc::mma_tf32_ts(tmem + ((uint32_t)(i + 1) & dmask) * (uint32_t)a.n, tmem + 392, db1, idesc, (uint32_t)(i + 1 >= a.ndest));
      }
    } else {
#pragma unroll 4
      for (int i = 0; i < a.count; i += 2) {
        tc::mma_tf32(tmem + ((uint32_t)i & dmask) * (uint32_t)a.n, da0, db0, idesc, (uint32_t)(i >= a.ndest));
        tc::mma_tf32(tmem + ((uint32_t)(i + 1) & dmask) * (uint32_t)a.n, da1, db1, idesc, (uint32_t)(i + 1 >= a.ndest));
      }
    }
    const long long t1 = clock64();
    *reinterpret_cast<volatile int*>(&flag) = 1;   // the MMAs are queued: warp 1 now times a TMEM load
    tc::mma_commit(&bar);
    tc::mbar_wait(&bar, 0);
    const long long t2 = clock64();
    a.out[0] = t1 - t0;
    a.out[1] = t2 - t0;
  }
  if (warp == 1) {
    // TMEM load round trip while `count` MMAs are queued on the tensor pipe (columns 448.. are not touched by the MMAs)
    while (*reinterpret_cast<volatile int*>(&flag) == 0) {}
    const long long t0 = clock64();
    uint32_t r[8];
    tc::tmem_ld8(tmem + (32u << 16) + 448, r);
    tc::tmem_ld_wait8(r);
    const long long t1 = clock64();
    if (tid == 32) { a.out[2] = t1 - t0; a.out[3] = (long long)r[0]; }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem);
}
}  // namespace

extern "C" int hgb_mma_probe(int32_t n, int32_t count, int32_t ndest, int32_t ts, int32_t same_ab, long long* out_dev, void* stream) {
  HGB_DEVICE_GUARD(out_dev);
  HGB_CHECK_ARG(out_dev && n >= 16 && n % 16 == 0 && n <= 256 && ndest >= 1 && ndest * n <= 384 && count >= 1, "hgb_mma_probe: bad arguments");
  ProbeArgs a{n, count, ndest, ts, same_ab, out_dev};
  HGB_CUDA_OK(cudaFuncSetAttribute(mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  mma_probe_kernel<<<1, 128, 65536, (cudaStream_t)stream>>>(a);
  HGB_LAUNCH_OK("mma_probe_kernel");
  return 0;
}

// ---- cp.async.bulk probe: lane 0 of each of `nwarps` warps issues `count` bulk copies of `bytes` from a (L2-resident after the
// first sweep) global buffer into its own ring of `depth` shared-memory slots, each with its own mbarrier; prefetch != 0 issues
// cp.async.bulk.prefetch.L2 instead.  out[0] = cycles per copy seen by warp 0 (issue + completion with `depth` in flight),
// out[1] = cycles of the issue instructions alone (first `depth` copies).
namespace {
struct TmaArgs { const float* src; int bytes, count, depth, prefetch; long long* out; };
__global__ void __launch_bounds__(256, 1) tma_probe_kernel(const TmaArgs a) {
  extern __shared__ __align__(128) float smem[];
  __shared__ uint64_t bars[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 64; ++i) tc::mbar_init(&bars[i], 1);
    tc::mbar_fence_init();
  }
  __syncthreads();
  if (lane == 0) {
    const int slot_floats = a.bytes / 4;
    float* ring = smem + (size_t)warp * a.depth * slot_floats;
    uint64_t* wb = bars + warp * 8;
    const float* src = a.src + (size_t)warp * 8 * slot_floats;
    long long t_issue = 0;
    const long long t0 = clock64();
    int s = 0, ph = 0;
    for (int i = 0; i < a.count; ++i) {
      if (a.prefetch) {
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + (size_t)(i & 7) * slot_floats), "r"((uint32_t)a.bytes) : "memory");
      } else {
        if (i >= a.depth) tc::mbar_wait(&wb[s], (uint32_t)(ph ^ 1));
        const uint32_t bar = tc::smem_u32(&wb[s]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)a.bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc::smem_u32(ring + s * slot_floats)),
                     "l"(src + (size_t)(i & 7) * slot_floats), "r"((uint32_t)a.bytes), "r"(bar)
                     : "memory");
      }
      if (i == a.depth - 1) t_issue = clock64() - t0;
      if (++s == a.depth) { s = 0; ph ^= 1; }
    }
    if (!a.prefetch) {
      // drain: the last phase of every slot
      int ss = s, pp = ph;
      for (int k = 0; k < a.depth; ++k) {
        // slot ss was last used in phase (pp ^ 1) if ss >= s ... simply wait for the most recent completed phase of each slot
        const int last_phase = (ss < s) ? ph : (ph ^ 1);
        tc::mbar_wait(&wb[ss], (uint32_t)last_phase);
        if (++ss == a.depth) ss = 0;
        (void)pp;
      }
    }
    const long long t1 = clock64();
    if (warp == 0) {
      a.out[0] = (t1 - t0) / a.count;
      a.out[1] = t_issue / a.depth;
    }
    (void)nw;
  }
}
}  // namespace

extern "C" int hgb_tma_probe(const float* src_dev, int32_t bytes, int32_t count, int32_t depth, long long* out_dev, void* stream) {
  HGB_DEVICE_GUARD(out_dev);
  // depth encodes: low 8 bits ring depth, bits 8-15 issuing warps (0 = 1), bit 16 = prefetch.L2 instead of a copy
  const int nwarps = ((depth >> 8) & 0xFF) ? ((depth >> 8) & 0xFF) : 1, prefetch = (depth >> 16) & 1;
  depth &= 0xFF;
  HGB_CHECK_ARG(src_dev && out_dev && bytes >= 16 && bytes % 16 == 0 && depth >= 1 && depth <= 8 && nwarps <= 8 &&
                    (size_t)bytes * depth * nwarps <= 200 * 1024 && count >= depth && count % depth == 0,
                "hgb_tma_probe: bad arguments");
  TmaArgs a{src_dev, bytes, count, depth, prefetch, out_dev};
  HGB_CUDA_OK(cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  tma_probe_kernel<<<1, 32 * nwarps, 200 * 1024, (cudaStream_t)stream>>>(a);
  HGB_LAUNCH_OK("tma_probe_kernel");
  return 0;
}
