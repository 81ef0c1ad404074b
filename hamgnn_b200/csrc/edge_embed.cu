// a1 + a2: edge vector, real spherical harmonics (component normalisation, reference axis order),
// Bessel basis x cosine cutoff.  HBM-bound: ~28 B of indices/shifts in, (S + R + 4) * 4 B out per edge;
// each thread evaluates one edge into a shared-memory row, the CTA then stores whole rows coalesced.
//
// SH evaluation: sqrt(4 pi) * standard real Y_lm of the physical unit vector (x,y,z)
//   Y_{l,+m} = N_lm sqrt2 Q_l^m(z) Re (x+iy)^m,  Y_{l,-m} = N_lm sqrt2 Q_l^m(z) Im (x+iy)^m,  Y_l0 = N_l0 Q_l^0(z)
// with Q_l^m the polynomial part of the associated Legendre function (no Condon-Shortley phase), which is
// what e3nn's o3.SphericalHarmonics(normalize=True, 'component') returns when it is fed v[:, [1,2,0]]
// as the reference does (toolbox/nequip/nn/embedding/_edge.py:45,65).
#include <math.h>

#include "hgb_common.cuh"

namespace {

constexpr int EE_THREADS = 128;

struct EEArgs {
  const float* pos;
  const float* nbr_shift;
  const int64_t* edge_index;
  int64_t n_edges;
  float* sh;
  float* rbf;
  float* edge_vec;
  float* edge_len;
  float cutoff;
  int num_radial;
  int sh_dim;
  int want_mask;            // bit l set -> irrep l is emitted
  int sh_off[HGB_MAX_L + 1];  // column of Y_l in the output row
  float norm[HGB_MAX_L + 1][HGB_MAX_L + 1];  // N_lm (times sqrt2 for m>0)
  float freq[128];          // n*pi/cutoff evaluated in fp32 exactly like the reference buffer
};

template <int LMAX>
__global__ void __launch_bounds__(EE_THREADS) edge_embed_kernel(const __grid_constant__ EEArgs a) {
  extern __shared__ float srow[];
  const int ld = (a.sh_dim + a.num_radial) | 1;  // odd stride: conflict-free row-per-thread writes
  const int t = threadIdx.x;
  const int64_t e = (int64_t)blockIdx.x * EE_THREADS + t;
  float* my = srow + (size_t)t * ld;
  if (e < a.n_edges) {
    const int64_t j = a.edge_index[e];
    const int64_t i = a.edge_index[a.n_edges + e];
    // (pos[i] + nbr_shift) - pos[j], fp32, same association as the reference
    float vx = __fsub_rn(__fadd_rn(a.pos[3 * i + 0], a.nbr_shift[3 * e + 0]), a.pos[3 * j + 0]);
    float vy = __fsub_rn(__fadd_rn(a.pos[3 * i + 1], a.nbr_shift[3 * e + 1]), a.pos[3 * j + 1]);
    float vz = __fsub_rn(__fadd_rn(a.pos[3 * i + 2], a.nbr_shift[3 * e + 2]), a.pos[3 * j + 2]);
    float r2 = __fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz));
    float r = sqrtf(r2);
    float rn = fmaxf(r, 1e-12f);
    float x = vx / rn, y = vy / rn, z = vz / rn;
    a.edge_vec[3 * e + 0] = __fdiv_rn(vx, r);
    a.edge_vec[3 * e + 1] = __fdiv_rn(vy, r);
    a.edge_vec[3 * e + 2] = __fdiv_rn(vz, r);
    a.edge_len[e] = r;

    // ---- spherical harmonics, all indices static
    float cm[LMAX + 1], sm[LMAX + 1];  // Re/Im (x+iy)^m
    cm[0] = 1.f; sm[0] = 0.f;
#pragma unroll
    for (int m = 1; m <= LMAX; ++m) {
      cm[m] = cm[m - 1] * x - sm[m - 1] * y;
      sm[m] = sm[m - 1] * x + cm[m - 1] * y;
    }
#pragma unroll
    for (int m = 0; m <= LMAX; ++m) {
      // Q_m^m = (2m-1)!!
      float qmm = 1.f;
#pragma unroll
      for (int k = 1; k <= m; ++k) qmm *= (float)(2 * k - 1);
      float q_prev2 = 0.f, q_prev = qmm;
#pragma unroll
      for (int l = m; l <= LMAX; ++l) {
        float q;
        if (l == m) q = qmm;
        else if (l == m + 1) q = (float)(2 * m + 1) * z * qmm;
        else q = ((float)(2 * l - 1) * z * q_prev - (float)(l + m - 1) * q_prev2) / (float)(l - m);
        if (l > m) { q_prev2 = q_prev; q_prev = q; }
        if ((a.want_mask >> l) & 1) {
          const float nq = a.norm[l][m] * q;
          if (m == 0) my[a.sh_off[l] + l] = nq;
          else {
            my[a.sh_off[l] + l + m] = nq * cm[m];
            my[a.sh_off[l] + l - m] = nq * sm[m];
          }
        }
      }
    }
    // ---- radial basis
    const float cut = (r < a.cutoff) ? 0.5f * (cosf(r * 3.14159265358979323846f / a.cutoff) + 1.0f) : 0.f;
    float* rb = my + a.sh_dim;
    for (int n = 0; n < a.num_radial; ++n) {
      float ax = __fmul_rn(r, a.freq[n]);
      rb[n] = __fmul_rn(__fdiv_rn(sinf(ax), r), cut);
    }
  }
  __syncthreads();
  // coalesced row stores
  const int64_t e0 = (int64_t)blockIdx.x * EE_THREADS;
  const int ne = (int)min((int64_t)EE_THREADS, a.n_edges - e0);
  for (int idx = t; idx < ne * a.sh_dim; idx += EE_THREADS) {
    int rr = idx / a.sh_dim, c = idx - rr * a.sh_dim;
    a.sh[e0 * a.sh_dim + idx] = srow[(size_t)rr * ld + c];
  }
  for (int idx = t; idx < ne * a.num_radial; idx += EE_THREADS) {
    int rr = idx / a.num_radial, c = idx - rr * a.num_radial;
    a.rbf[e0 * a.num_radial + idx] = srow[(size_t)rr * ld + a.sh_dim + c];
  }
}

template <int LMAX>
int launch(const EEArgs& a, cudaStream_t st) {
  const int ld = (a.sh_dim + a.num_radial) | 1;
  const size_t smem = (size_t)EE_THREADS * ld * sizeof(float);
  HGB_CUDA_OK(cudaFuncSetAttribute(edge_embed_kernel<LMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned grid = (unsigned)((a.n_edges + EE_THREADS - 1) / EE_THREADS);
  edge_embed_kernel<LMAX><<<grid, EE_THREADS, smem, st>>>(a);
  HGB_LAUNCH_OK("edge_embed_kernel");
  return 0;
}

}  // namespace

extern "C" int hgb_edge_embed(const float* pos, const float* nbr_shift, const int64_t* edge_index, int64_t n_edges,
                              const int32_t* sh_ls_host, int32_t n_ls, float cutoff, const float* bessel_freqs_host,
                              int32_t num_radial, float* sh, float* rbf, float* edge_vec, float* edge_len,
                              void* stream) {
  HGB_DEVICE_GUARD(sh);
  hgb::TimeScope ts_(HGB_K_EDGE_EMBED, stream);
  HGB_CHECK_ARG(n_edges >= 0, "hgb_edge_embed: negative edge count");
  HGB_CHECK_ARG(num_radial >= 1 && num_radial <= 128, "hgb_edge_embed: num_radial %d out of range [1,128]", num_radial);
  HGB_CHECK_ARG(n_ls >= 1 && n_ls <= HGB_MAX_L + 1, "hgb_edge_embed: bad number of SH irreps %d", n_ls);
  HGB_CHECK_ARG(cutoff > 0.f, "hgb_edge_embed: cutoff must be positive");
  if (n_edges == 0) return 0;
  EEArgs a;
  memset(&a, 0, sizeof(a));
  a.pos = pos; a.nbr_shift = nbr_shift; a.edge_index = edge_index; a.n_edges = n_edges;
  a.sh = sh; a.rbf = rbf; a.edge_vec = edge_vec; a.edge_len = edge_len;
  a.cutoff = cutoff; a.num_radial = num_radial;
  int off = 0, lmax = 0;
  for (int q = 0; q < n_ls; ++q) {
    int l = sh_ls_host[q];
    HGB_CHECK_ARG(l >= 0 && l <= HGB_MAX_L, "hgb_edge_embed: l=%d unsupported (max %d)", l, HGB_MAX_L);
    HGB_CHECK_ARG(!((a.want_mask >> l) & 1), "hgb_edge_embed: irreps_edge_sh lists l=%d twice", l);
    a.want_mask |= 1 << l;
    a.sh_off[l] = off;
    off += 2 * l + 1;
    if (l > lmax) lmax = l;
  }
  a.sh_dim = off;
  for (int l = 0; l <= HGB_MAX_L; ++l)
    for (int m = 0; m <= l; ++m) {
      double f = 1.0;  // (l-m)!/(l+m)!
      for (int k = l - m + 1; k <= l + m; ++k) f /= (double)k;
      double n = sqrt((2.0 * l + 1.0) * f);
      a.norm[l][m] = (float)(m == 0 ? n : n * sqrt(2.0));
    }
  // the reference's registered buffer BesselBasis.freqs = arange(1, n+1) * pi / cutoff (basis_functions.py:188-189),
  // passed in so that its fp32 rounding (and any state_dict override) is honoured bit for bit
  HGB_CHECK_ARG(bessel_freqs_host != nullptr, "hgb_edge_embed: bessel_freqs_host is NULL");
  for (int n = 0; n < num_radial; ++n) a.freq[n] = bessel_freqs_host[n];
  cudaStream_t st = (cudaStream_t)stream;
  switch (lmax) {
    case 0: return launch<0>(a, st);
    case 1: return launch<1>(a, st);
    case 2: return launch<2>(a, st);
    case 3: return launch<3>(a, st);
    case 4: return launch<4>(a, st);
    case 5: return launch<5>(a, st);
    case 6: return launch<6>(a, st);
    case 7: return launch<7>(a, st);
    default: return launch<8>(a, st);
  }
}
