// Graph construction on the device: the reference builds its internal message-passing graph on the
// CPU with ASE's neighbour list (hamgnn/models/base_model.py:87-178, generate_graph :237-288: "neighborlists require a round
// trip to the CPU") -- directed edge i -> (j, S) over all periodic images iff 0 < |r_j + S.cell - r_i| < rc_i + rc_j with
// per-atom cutoff radii -- and matches the DFT edges of the data against it column by column (find_matching_columns_of_A_in_B,
// :180-233).  Here: one thread per source atom walks (j, S) in lexicographic order, in fp64 like ASE, first counting, then
// (after an exclusive scan on the host side of the ABI) writing its edges; edges are therefore sorted by (i, j, S) and a
// (i, j, S) lookup is a binary search inside the segment of i.  10^4 atoms x 9 images: 8.6e8 distance tests, ~6 ms.
#include "hgb_common.cuh"

namespace {

struct NbrArgs {
  const double* pos;     // [N][3]
  const double* radius;  // [N]
  double cell[9];        // rows = lattice vectors
  int reps[3];           // images -reps .. +reps along each lattice vector
  int64_t n;
  int64_t* deg;          // count pass: [N]
  const int64_t* offset; // fill pass: exclusive scan of deg, [N + 1]
  int64_t* edge_index;   // [2][E]
  int64_t* cell_shift;   // [E][3]
  float* nbr_shift;      // [E][3] = cell_shift @ cell (fp32, what the models read)
  int64_t n_edges;
};

template <bool FILL>
__global__ void __launch_bounds__(128) neighbor_kernel(const __grid_constant__ NbrArgs a) {
  const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (i >= a.n) return;
  const double xi = a.pos[3 * i], yi = a.pos[3 * i + 1], zi = a.pos[3 * i + 2], ri = a.radius[i];
  int64_t cnt = 0, w = FILL ? a.offset[i] : 0;
  for (int64_t j = 0; j < a.n; ++j) {
    const double rc = ri + a.radius[j];
    const double dx0 = a.pos[3 * j] - xi, dy0 = a.pos[3 * j + 1] - yi, dz0 = a.pos[3 * j + 2] - zi;
    for (int sa = -a.reps[0]; sa <= a.reps[0]; ++sa)
      for (int sb = -a.reps[1]; sb <= a.reps[1]; ++sb)
        for (int sc = -a.reps[2]; sc <= a.reps[2]; ++sc) {
          const double tx = sa * a.cell[0] + sb * a.cell[3] + sc * a.cell[6];
          const double ty = sa * a.cell[1] + sb * a.cell[4] + sc * a.cell[7];
          const double tz = sa * a.cell[2] + sb * a.cell[5] + sc * a.cell[8];
          const double dx = dx0 + tx, dy = dy0 + ty, dz = dz0 + tz;
          const double d2 = dx * dx + dy * dy + dz * dz;
          if (d2 < rc * rc && d2 > 1e-16) {
            if (FILL) {
              a.edge_index[w] = i;
              a.edge_index[a.n_edges + w] = j;
              a.cell_shift[3 * w] = sa; a.cell_shift[3 * w + 1] = sb; a.cell_shift[3 * w + 2] = sc;
              a.nbr_shift[3 * w] = (float)tx; a.nbr_shift[3 * w + 1] = (float)ty; a.nbr_shift[3 * w + 2] = (float)tz;
              ++w;
            } else {
              ++cnt;
            }
          }
        }
  }
  if (!FILL) a.deg[i] = cnt;
}

// index of the edge (src, dst, shift) in a graph whose edges are sorted by (src, dst, shift) with segment offsets `offset`
// (-1 if absent); sign = -1 looks up the inverse edge (dst, src, -shift) instead.
__global__ void __launch_bounds__(128) edge_lookup_kernel(const int64_t* __restrict__ q_index, const int64_t* __restrict__ q_shift, int64_t nq,
                                                          int sign, const int64_t* __restrict__ g_index, const int64_t* __restrict__ g_shift,
                                                          const int64_t* __restrict__ offset, int64_t ng, int64_t* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (e >= nq) return;
  int64_t s = q_index[e], d = q_index[nq + e];
  int64_t k0 = q_shift[3 * e], k1 = q_shift[3 * e + 1], k2 = q_shift[3 * e + 2];
  if (sign < 0) { const int64_t t = s; s = d; d = t; k0 = -k0; k1 = -k1; k2 = -k2; }
  int64_t lo = offset[s], hi = offset[s + 1];
  while (lo < hi) {   // lower bound of (d, k0, k1, k2) in the segment
    const int64_t mid = (lo + hi) >> 1;
    const int64_t md = g_index[ng + mid], m0 = g_shift[3 * mid], m1 = g_shift[3 * mid + 1], m2 = g_shift[3 * mid + 2];
    const bool less = md < d || (md == d && (m0 < k0 || (m0 == k0 && (m1 < k1 || (m1 == k1 && m2 < k2)))));
    if (less) lo = mid + 1;
    else hi = mid;
  }
  const bool hit = lo < offset[s + 1] && g_index[ng + lo] == d && g_shift[3 * lo] == k0 && g_shift[3 * lo + 1] == k1 && g_shift[3 * lo + 2] == k2;
  out[e] = hit ? lo : -1;
}

}  // namespace

extern "C" int hgb_neighbor_list(const double* pos, const double* radius, const double* cell_host, const int32_t* reps_host, int64_t n_atoms,
                                 int64_t* deg, const int64_t* offset, int64_t n_edges, int64_t* edge_index, int64_t* cell_shift,
                                 float* nbr_shift, void* stream) {
  HGB_DEVICE_GUARD(pos);
  HGB_CHECK_ARG(pos && radius && cell_host && reps_host, "hgb_neighbor_list: NULL argument");
  HGB_CHECK_ARG(n_atoms >= 0 && n_atoms < (1ll << 31), "hgb_neighbor_list: bad atom count");
  HGB_CHECK_ARG((deg != nullptr) != (offset != nullptr), "hgb_neighbor_list: pass deg (count pass) or offset (fill pass)");
  HGB_CHECK_ARG(reps_host[0] >= 0 && reps_host[1] >= 0 && reps_host[2] >= 0 && reps_host[0] <= 64 && reps_host[1] <= 64 && reps_host[2] <= 64,
                "hgb_neighbor_list: bad image range");
  if (n_atoms == 0) return 0;
  NbrArgs a;
  memset(&a, 0, sizeof(a));
  a.pos = pos; a.radius = radius; a.n = n_atoms;
  for (int k = 0; k < 9; ++k) a.cell[k] = cell_host[k];
  for (int k = 0; k < 3; ++k) a.reps[k] = reps_host[k];
  const unsigned grid = (unsigned)((n_atoms + 127) / 128);
  if (deg) {
    a.deg = deg;
    neighbor_kernel<false><<<grid, 128, 0, (cudaStream_t)stream>>>(a);
  } else {
    HGB_CHECK_ARG(n_edges >= 0 && (n_edges == 0 || (edge_index && cell_shift && nbr_shift)), "hgb_neighbor_list: fill pass needs the output buffers");
    a.offset = offset; a.n_edges = n_edges; a.edge_index = edge_index; a.cell_shift = cell_shift; a.nbr_shift = nbr_shift;
    neighbor_kernel<true><<<grid, 128, 0, (cudaStream_t)stream>>>(a);
  }
  HGB_LAUNCH_OK("neighbor_kernel");
  return 0;
}

extern "C" int hgb_edge_lookup(const int64_t* q_index, const int64_t* q_shift, int64_t n_query, int32_t inverse, const int64_t* g_index,
                               const int64_t* g_shift, const int64_t* g_offset, int64_t n_graph_edges, int64_t* out, void* stream) {
  HGB_DEVICE_GUARD(out);
  HGB_CHECK_ARG(out && (n_query == 0 || (q_index && q_shift && g_index && g_shift && g_offset)), "hgb_edge_lookup: NULL argument");
  HGB_CHECK_ARG(n_query >= 0 && n_graph_edges >= 0, "hgb_edge_lookup: bad sizes");
  if (n_query == 0) return 0;
  edge_lookup_kernel<<<(unsigned)((n_query + 127) / 128), 128, 0, (cudaStream_t)stream>>>(q_index, q_shift, n_query, inverse ? -1 : 1, g_index, g_shift,
                                                                                        g_offset, n_graph_edges, out);
  HGB_LAUNCH_OK("edge_lookup_kernel");
  return 0;
}
