// Band-energy head, reciprocal-space assembly: H(k), S(k) of one crystal from the real-space blocks,
//   X(k)[(i,o1),(j,o2)] = delta_ij Xon[i][o1,o2] + sum_{edges e = (i -> j, R_e)} exp(2 pi i k.R_e) Xoff[e][o1,o2],
// restricted to the orbitals the basis defines for each species.  Replaces the dense [num_k, Na, Na, nao, nao] scatter of
// hamgnn/models/hamgnn_output.py:1775-1909 (index_put_ with accumulate per k point, swapaxes, masked_select): the compact
// [num_k][n_orb][n_orb] complex matrices are written directly.  Edges are grouped by (i, j) on the host side of the ABI (stable
// sort), so one CTA owns one (i, j) block of one k point and adds the lattice images of the pair serially in a fixed order: no
// atomics, bit-reproducible (the reference's index_put accumulate is an atomic add on the GPU).  The generalized eigenproblem
// itself (Cholesky, triangular inverses, eigh) stays with cuSOLVER through torch.linalg (hamgnn_b200/band.py): plain library
// linear algebra.  HBM-bound: every off-site block is read n_k times (L2-resident per segment), n_k n_orb^2 complex written once.
#include "hgb_common.cuh"

namespace {

struct BandArgs {
  const float* hon; const float* hoff; const float* son; const float* soff;
  const int64_t* seg_ptr; const int64_t* seg_edge; const int64_t* src; const int64_t* dst;
  const float* shift; const float* kvec; const int32_t* orb_index;
  int64_t n_atoms; int32_t nao, n_k, n_orb;
  float2* hk; float2* sk;
};

// on-site blocks: grid (atoms, k points); plain stores (every diagonal block has exactly one writer)
__global__ void __launch_bounds__(128) band_onsite_kernel(const __grid_constant__ BandArgs a) {
  const int64_t i = blockIdx.x;
  const int k = blockIdx.y, nn2 = a.nao * a.nao;
  const int32_t* oi = a.orb_index + i * a.nao;
  float2* hk = a.hk + (size_t)k * a.n_orb * a.n_orb;
  float2* sk = a.sk + (size_t)k * a.n_orb * a.n_orb;
  for (int idx = threadIdx.x; idx < nn2; idx += blockDim.x) {
    const int o1 = idx / a.nao, o2 = idx - o1 * a.nao;
    const int r = oi[o1], c = oi[o2];
    if (r < 0 || c < 0) continue;
    hk[(size_t)r * a.n_orb + c] = make_float2(a.hon[i * nn2 + idx], 0.f);
    sk[(size_t)r * a.n_orb + c] = make_float2(a.son[i * nn2 + idx], 0.f);
  }
}

// off-site blocks: grid ((i, j) segments, k points); the images of a pair are summed in list order, then added to the block
__global__ void __launch_bounds__(128) band_offsite_kernel(const __grid_constant__ BandArgs a) {
  const int64_t s = blockIdx.x;
  const int k = blockIdx.y, nn2 = a.nao * a.nao;
  const int64_t e0 = a.seg_ptr[s], e1 = a.seg_ptr[s + 1];
  if (e1 <= e0) return;
  const int64_t first = a.seg_edge[e0];
  const int64_t i = a.src[first], j = a.dst[first];
  const int32_t* oi = a.orb_index + i * a.nao;
  const int32_t* oj = a.orb_index + j * a.nao;
  const float kx = a.kvec[3 * k], ky = a.kvec[3 * k + 1], kz = a.kvec[3 * k + 2];
  float2* hk = a.hk + (size_t)k * a.n_orb * a.n_orb;
  float2* sk = a.sk + (size_t)k * a.n_orb * a.n_orb;
  for (int idx = threadIdx.x; idx < nn2; idx += blockDim.x) {
    const int o1 = idx / a.nao, o2 = idx - o1 * a.nao;
    const int r = oi[o1], c = oj[o2];
    if (r < 0 || c < 0) continue;
    float hr = 0.f, hi = 0.f, sr = 0.f, si = 0.f;
    for (int64_t q = e0; q < e1; ++q) {
      const int64_t e = a.seg_edge[q];
      const float d = a.shift[3 * e] * kx + a.shift[3 * e + 1] * ky + a.shift[3 * e + 2] * kz;
      float sn, cs;
      sincospif(2.f * d, &sn, &cs);   // exp(2 pi i k.R)
      const float h = a.hoff[e * nn2 + idx], sv = a.soff[e * nn2 + idx];
      hr = fmaf(cs, h, hr); hi = fmaf(sn, h, hi);
      sr = fmaf(cs, sv, sr); si = fmaf(sn, sv, si);
    }
    float2& ph = hk[(size_t)r * a.n_orb + c];
    float2& ps = sk[(size_t)r * a.n_orb + c];
    ph = make_float2(ph.x + hr, ph.y + hi);
    ps = make_float2(ps.x + sr, ps.y + si);
  }
}

}  // namespace

extern "C" int hgb_band_kspace(const float* hon, const float* hoff, const float* son, const float* soff, int64_t n_atoms,
                               int32_t nao, const int64_t* seg_ptr, int64_t n_segs, const int64_t* seg_edge, const int64_t* src,
                               const int64_t* dst, const float* nbr_shift, const float* kvec, int32_t n_k,
                               const int32_t* orb_index, int32_t n_orb, float* hk, float* sk, void* stream) {
  HGB_DEVICE_GUARD(hk);
  HGB_CHECK_ARG(hon && son && orb_index && kvec && hk && sk, "hgb_band_kspace: NULL argument");
  HGB_CHECK_ARG(n_segs == 0 || (hoff && soff && seg_ptr && seg_edge && src && dst && nbr_shift), "hgb_band_kspace: NULL edge argument");
  HGB_CHECK_ARG(n_atoms >= 1 && n_atoms < (1ll << 31) && nao >= 1 && nao <= 64 && n_k >= 1 && n_k <= 65535 && n_orb >= 1 &&
                    (int64_t)n_orb <= n_atoms * nao && n_segs >= 0 && n_segs < (1ll << 31),
                "hgb_band_kspace: bad sizes (atoms %lld, nao %d, k points %d, orbitals %d)", (long long)n_atoms, nao, n_k, n_orb);
  cudaStream_t st = (cudaStream_t)stream;
  HGB_CUDA_OK(cudaMemsetAsync(hk, 0, (size_t)n_k * n_orb * n_orb * sizeof(float2), st));
  HGB_CUDA_OK(cudaMemsetAsync(sk, 0, (size_t)n_k * n_orb * n_orb * sizeof(float2), st));
  BandArgs a;
  a.hon = hon; a.hoff = hoff; a.son = son; a.soff = soff; a.seg_ptr = seg_ptr; a.seg_edge = seg_edge; a.src = src; a.dst = dst;
  a.shift = nbr_shift; a.kvec = kvec; a.orb_index = orb_index; a.n_atoms = n_atoms; a.nao = nao; a.n_k = n_k; a.n_orb = n_orb;
  a.hk = reinterpret_cast<float2*>(hk); a.sk = reinterpret_cast<float2*>(sk);
  {
    hgb::TimeScope ts(HGB_K_OTHER, stream);
    band_onsite_kernel<<<dim3((unsigned)n_atoms, (unsigned)n_k), 128, 0, st>>>(a);
    HGB_LAUNCH_OK("band_onsite_kernel");
    if (n_segs > 0) {
      band_offsite_kernel<<<dim3((unsigned)n_segs, (unsigned)n_k), 128, 0, st>>>(a);
      HGB_LAUNCH_OK("band_offsite_kernel");
    }
  }
  return 0;
}
