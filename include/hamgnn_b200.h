/* hamgnn_b200.h -- C ABI of the B200-native HamGNN hot path (libhamgnn_b200.so).
 *
 * The reference (QuantumLab-ZY/HamGNN) has no FFI: its hot path is Python calling e3nn / torch_scatter.
 * Each entry point below replaces the arithmetic of the cited reference function(s); a maintainer binds
 * them with ctypes from inside the two nn.Modules `HamGNNConvE3` / `HamGNNPlusPlusOut`.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`; row-major, contiguous;
 *     float = fp32, indices int64 (edge_index, inv_edge_idx, z) exactly as the reference's Data fields.
 *   - `stream` is a cudaStream_t passed as void*.
 *   - return value: 0 on success, non-zero on failure; hgb_last_error() gives the message
 *     (thread-local).  No entry point synchronises the stream.
 *   - plans (`hgb_*_plan`) are plain structs of device pointers to constant tables built by the host
 *     (irreps instruction tables, sparse Clebsch-Gordan lists, packed weights); the library never
 *     allocates or frees device memory.
 */
#ifndef HAMGNN_B200_H
#define HAMGNN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HGB_ABI_VERSION 1
#define HGB_MAX_L 8

int hgb_abi_version(void);
const char* hgb_last_error(void);
/* number of kernels launched by this library in the calling process (for bench.py's gpu_launches) */
int64_t hgb_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * a1+a2: edge geometry, spherical harmonics, Bessel x cosine-cutoff radial embedding.
 * Replaces SphericalHarmonicEdgeAttrs.forward (hamgnn/toolbox/nequip/nn/embedding/_edge.py:59-67),
 * RadialBasisEdgeEncoding.forward (hamgnn/nn/embeddings.py:73-100), BesselBasis.forward
 * (hamgnn/utils/basis_functions.py:193-208), CosineCutoff.forward (hamgnn/utils/cutoff_functions.py:50-61).
 *   edge_index [2,E] (row 0 = j, row 1 = i): vec = pos[i] + nbr_shift - pos[j]
 *   sh_ls[n_ls]: the l of every irrep of irreps_edge_sh (mul 1 each), output sh [E, sum(2l+1)],
 *   'component' normalisation of the unit vector (||Y_l||^2 = 2l+1), reference axis order (y,z,x).
 *   rbf [E,num_radial] = sin(freq_n r)/r * 0.5(cos(pi r/rc)+1) [r<rc], freq = the reference's BesselBasis.freqs
 *   buffer (HOST pointer, num_radial floats);  edge_vec [E,3] unit; edge_len [E].
 */
int hgb_edge_embed(const float* pos, const float* nbr_shift, const int64_t* edge_index, int64_t n_edges,
                   const int32_t* sh_ls_host, int32_t n_ls, float cutoff, const float* bessel_freqs_host,
                   int32_t num_radial, float* sh, float* rbf, float* edge_vec, float* edge_len, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a5-a8 (+a9 scatter, +a11 skip): fused MessagePackBlock.
 * Replaces MessagePackBlock.forward (hamgnn/nn/message_passing.py:216-231), LinearScaleWithWeights
 * (hamgnn/nn/tensor_products.py:25-47), the FullyConnectedNet radial weight generators
 * (message_passing.py:173-189), AttentionHeadsToVector (hamgnn/nn/attention_utils.py:85-120),
 * torch_scatter.scatter in ConvBlockE3.forward (hamgnn/nn/convolution.py:147-149) and the
 * `edge_feats_mix + skip_linear(edge_feats)` of PairInteractionBlock.forward
 * (hamgnn/nn/interaction_blocks.py:154-155).  The same kernel with one branch evaluates
 * TensorProductWithMemoryOptimizationWithWeight.forward (hamgnn/nn/tensor_products.py:170-189).
 */
typedef struct {
  int32_t mul;      /* multiplicity of the output slot                              */
  int32_t mpad;     /* mul rounded up to a multiple of 4 (weight tables are padded) */
  int32_t l;        /* l of the output irrep                                        */
  int32_t out_off;  /* first column of the slot in the output row                   */
  int32_t path_begin, path_end; /* paths feeding this slot: [begin, end)            */
  int32_t pad0, pad1;
} hgb_type_t;

typedef struct {
  int32_t kind;      /* 0: CG path (A W) * gate -> L';  1: direct linear (no SH, no gate)           */
  int32_t branch;    /* which radial MLP / input group the path belongs to                         */
  int32_t src0;      /* index of the first input source                                            */
  int32_t nsrc;      /* 1, or 2 for the fused (src|dst) node input: channel u -> source src0+u/mul  */
  int32_t in_off;    /* first column of the input irrep block inside its source row                */
  int32_t mul_in;    /* multiplicity per source; K = nsrc * mul_in                                 */
  int32_t l1, l2, l3;
  int32_t sh_off;    /* first column of Y_l2 in the SH row                                          */
  int32_t cg_off;    /* offset into cg_ij / cg_val (entries sorted by k)                            */
  int32_t cg_kstart; /* offset into cg_kstart table (2*l3+2 entries)                                */
  int32_t w_off;     /* packed TP weight  [K][mpad]   (scaled by the path coefficient)              */
  int32_t w3_off;    /* packed radial last layer [H2][mpad] (scaled by 1/sqrt(H2))                   */
  int32_t lf_off;    /* packed folded (mid->D Linear) x (out Linear): [mpad][mpad]; kind 1: [K][mpad] */
  int32_t pad0;
} hgb_path_t;

typedef struct {
  int32_t n_types, n_paths, n_branches, n_sources;
  int32_t sh_dim, rbf_dim, h1, h2;   /* radial MLP sizes [rbf_dim, h1, h2, *]                   */
  int32_t out_dim;
  int32_t src_dim[4];                /* row length of every input source                          */
  int32_t fc1_off[2], fc2_off[2];    /* per branch: packed [rbf_dim][h1] and [h1][h2] (pre-scaled) */
  float act_const;                   /* normalize2mom constant of silu                             */
  const hgb_type_t* types;           /* device copies (read by the kernel)                         */
  const hgb_path_t* paths;
  const hgb_type_t* types_host;      /* host copies of the same tables (validated on every call)   */
  const hgb_path_t* paths_host;
  const int32_t* cg_ij;              /* i | (j << 8)                                               */
  const float* cg_val;
  const int32_t* cg_kstart;
  const float* wbuf;                 /* all packed weights                                         */
} hgb_msgpack_plan;

/* src_rows[s]: row-gather index (int64, length E) for source s or NULL for identity (row e).
 * out_index: NULL -> out[e, :] = message (+ `out` is overwritten);
 *            else  -> out[out_index[e], :] += message  (out must be pre-zeroed; receiver scatter-sum).
 */
int hgb_msgpack_forward(const hgb_msgpack_plan* plan_host, const float* const* src_host,
                        const int64_t* const* src_rows_host, const float* sh, const float* rbf,
                        int64_t n_edges, float* out, const int64_t* out_index, void* stream);

/* Same operation on the tcgen05 tensor cores (3xTF32, fp32-accurate; csrc/msgpack_tc.cu).  `plan->wbuf` holds
 * the tensor-core packing (hi|lo operand images in the UMMA interleaved K-major layout, multiplicities padded to
 * 16 in plan->types[].mpad); h2_ws is a device workspace of n_branches * n_edges * h2 floats for the radial-MLP
 * hidden activations.  Replaces the same reference functions as hgb_msgpack_forward. */
int hgb_msgpack_tc_forward(const hgb_msgpack_plan* plan_host, const float* const* src_host,
                           const int64_t* const* src_rows_host, const float* sh, const float* rbf, float* h2_ws,
                           int64_t n_edges, float* out, const int64_t* out_index, void* stream);

/* Variant of the tensor-core kernel that evaluates the radial gate g = FCN(rbf) [n_branches][E][gstride] once per
 * call into the workspace g_ws (so that two CTAs fit per SM in the message kernel; csrc/msgpack_tcg.cu).
 * w3_off_host[b] / nch_host[b]: offset into plan->wbuf of the pre-scaled last radial layer [h2][nch_b], and nch_b;
 * plan->paths[].pad0 = first gate column of the path.  Replaces the same reference functions. */
int hgb_msgpack_tcg_forward(const hgb_msgpack_plan* plan_host, const float* const* src_host,
                            const int64_t* const* src_rows_host, const float* sh, const float* rbf,
                            const int32_t* w3_off_host, const int32_t* nch_host, int32_t gstride, float* g_ws,
                            int64_t n_edges, float* out, const int64_t* out_index, void* stream);
/* v2: w3img_off_host (nullable) = per-branch offsets into plan->wbuf of the last radial layer packed as tensor-core
 * tiles -- per tile of 64 gate columns a (hi | lo) pair of images [h2/4][64][4] (K-major, tf32 split, scaled by
 * 1/sqrt(h2), zero padded).  When given (and h2 % 8 == 0, h1 % 4 == 0) the gate pre-pass runs as a tcgen05 GEMM
 * (radial_gate_tc_kernel) instead of the fp32-FMA kernel. */
int hgb_msgpack_tcg_forward_v2(const hgb_msgpack_plan* plan_host, const float* const* src_host,
                               const int64_t* const* src_rows_host, const float* sh, const float* rbf,
                               const int32_t* w3_off_host, const int32_t* nch_host, const int32_t* w3img_off_host,
                               int32_t gstride, float* g_ws, int64_t n_edges, float* out, const int64_t* out_index,
                               void* stream);
/* the radial gate alone: g_ws[b][e][c] = FullyConnectedNet_b(rbf[e])[c] (hamgnn/nn/message_passing.py:173-189),
 * c < nch_host[b], row stride gstride; same selection rule between the two kernels as above. */
int hgb_radial_gate(const hgb_msgpack_plan* plan_host, const float* rbf, const int32_t* w3_off_host,
                    const int32_t* nch_host, const int32_t* w3img_off_host, int32_t gstride, float* g_ws,
                    int64_t n_edges, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Edge-aligned ("rotated frame") evaluation of the same fused MessagePackBlock (csrc/msgpack_rot_kernel.cuh).
 *
 * In the frame where the edge vector is the polar axis, Y_l2 = sqrt(2 l2+1) delta_{m2,0} and the per-edge CG
 * contraction T[i,k] = sum_j w3j[i,j,k] Y[j] of o3.TensorProduct (hamgnn/nn/message_passing.py:81-96) has a single
 * non-zero per output component: a path splits into <= min(2 l1+1, 2 l3+1) "steps"
 *     C'_{m3} += ((X'_{m1} W_p) * (scale * g_p)) L'_p,       X'_{m1}[z,u] = (D^{l1}(R_z) x_z)[u, m1],
 * all of them tensor-core GEMMs on plain data.  hgb_wigner builds D^l(R_z) per edge (R_z takes the edge vector to
 * the polar axis), the message call rotates + packs the inputs into (hi | lo) operand images (rotate_pack_kernel),
 * runs the steps (msgpack_rot_kernel: TMA bulk copies -> tcgen05.mma 3xTF32 -> gate in TMEM -> tcgen05.mma) and
 * rotates the message back, C = D^{l3}(R_z)^T C', before the store / receiver scatter-add.  Equivariance of the
 * Wigner 3j symbols makes the result identical to the reference's up to fp32 rounding.
 */
typedef struct {
  int32_t src0, nsrc;   /* input sources of the block: channel u -> source src0 + u / mul            */
  int32_t in_off, mul;  /* first column / multiplicity of the irrep block inside a source row         */
  int32_t l1;
  int32_t kpad;         /* nsrc * mul rounded up to 8                                                 */
  int32_t xoff;         /* float offset of the block inside a tile of the packed rotated input:
                           [m1 = -l1..l1][chunk of <= 32 channels][hi | lo][k/4][128 edges][k%4]       */
  int32_t pad;
} hgb_rot_block_t;

typedef struct {
  int32_t a_off;        /* float offset of the step's A operand X'_{m1} inside a tile                  */
  int32_t w_off;        /* plan->wbuf offset of the W_p images (kind 1: of the direct Linear images)   */
  int32_t lf_off;       /* plan->wbuf offset of the L'_p image                                        */
  int32_t g_off;        /* first gate column of the path                                              */
  float scale;          /* w3j(l1,l2,l3)[m1, 0, m3] * sqrt(2 l2 + 1)                                  */
  int16_t kpad;
  int8_t kind;          /* 0 (every step runs GEMM1 -> gate -> GEMM2)                                 */
  int8_t branch;        /* radial-MLP branch of the gate; -1: un-gated (direct Linear of the edge features:
                           w_off = its images, lf_off = an identity image)                              */
  int8_t m3;            /* output component index l3 + m3                                             */
  int8_t new_path;      /* flags: bit 0 = the L' image is to be loaded (set on every step), bit 2 = last step of its
                           output-component (m3) group; steps of a slot are ordered by m3                 */
  int16_t pad;
  int32_t pad2;         /* plan->wbuf offset of the un-split fp32 copy of L'_p, row-major [mpad][mpad]
                           (read only by the experimental msgpack_rot_s2_kernel)                       */
} hgb_rot_step_t;

typedef struct {
  int32_t n_blocks, tile_stride, lmax, dstride;
  int32_t doff[12];        /* float offset of D^l inside a per-edge Wigner row (row-major d x d)        */
  int32_t step_begin[33];  /* steps of output slot t: [step_begin[t], step_begin[t+1])                 */
  int32_t pad;
  const hgb_rot_block_t* blocks;  /* device */
  const hgb_rot_step_t* steps;    /* device */
  const hgb_rot_block_t* blocks_host;
  const hgb_rot_step_t* steps_host;
  const double* wigner_j;         /* device: J^l = D^l(R_x(-90 deg)) at doff[l], fp64                  */
} hgb_rot_plan;

/* dw[e][doff[l] + a * (2l+1) + b] = D^l(R_e)[a][b] for l <= lmax, R_e = R_y(-theta) R_z(-phi) with (theta, phi) the
 * polar angles of edge_vec[e] (physical unit vector, as written by hgb_edge_embed); evaluated in fp64 as
 * J Z(-theta) J^T Z(-phi) from the vector components, stored fp32. */
int hgb_wigner(const hgb_rot_plan* rot_host, const float* edge_vec, int64_t n_edges, float* dw, void* stream);

/* The fused MessagePackBlock through the rotated frame.  plan: the tensor-core packing (as hgb_msgpack_tcg_forward);
 * dw: hgb_wigner output for the same edges; g_ws / xp_ws: workspaces of n_branches * chunk * gstride and
 * ceil(chunk / 128) * rot->tile_stride floats, chunk = min(chunk_edges, n_edges) (edges are processed chunk by
 * chunk, chunk_edges a multiple of 128).  Other arguments as hgb_msgpack_tcg_forward_v2. */
int hgb_msgpack_rot_forward(const hgb_msgpack_plan* plan_host, const hgb_rot_plan* rot_host, const float* const* src_host,
                            const int64_t* const* src_rows_host, const float* dw, const float* rbf,
                            const int32_t* w3_off_host, const int32_t* nch_host, const int32_t* w3img_off_host,
                            int32_t gstride, float* g_ws, float* xp_ws, int64_t chunk_edges, int64_t n_edges,
                            float* out, const int64_t* out_index, void* stream);

/* The same fused MessagePackBlock (hamgnn/nn/message_passing.py:26-231) through the rotated frame with fp16 x 2 split
 * operands (csrc/msgpack_rot16_kernel.cuh): every operand a = hi + lo, hi = fp16(a s), lo = fp16(a s - hi), s a power of
 * two per packed image / per (edge, input block) / per (edge, step) row, three kind::f16 MMAs (K = 16 per instruction
 * instead of K = 8 for kind::tf32) accumulate lo.hi + hi.lo + hi.hi in fp32.  rot_host: the fp16 program -- block kpad in
 * channels (multiple of 16), block xoff / step a_off, w_off, lf_off / tile_stride in 32-bit WORDS, step kpad = words per
 * operand row, step pad = input block, step pad2 = W image | L' image << 16 (indices into img_inv).  wbuf16: the packed
 * fp16 images (wbuf16_words 32-bit words), img_inv[n_images]: inverse scale per image.  xp_ws: ceil(chunk / 128) *
 * tile_stride words; sx_ws: ceil(chunk / 128) * n_blocks * 128 floats (inverse row scales written by the rotate-pack
 * kernel).  flags bit 0: debug, swap the two halves of every packed TMEM word.  Other arguments as
 * hgb_msgpack_rot_forward. */
int hgb_msgpack_rot16_forward(const hgb_msgpack_plan* plan_host, const hgb_rot_plan* rot_host, const float* const* src_host,
                              const int64_t* const* src_rows_host, const float* dw, const float* rbf,
                              const int32_t* w3_off_host, const int32_t* nch_host, const int32_t* w3img_off_host,
                              int32_t gstride, float* g_ws, float* xp_ws, float* sx_ws, const float* wbuf16,
                              int64_t wbuf16_words, const float* img_inv, int32_t n_images, int64_t chunk_edges,
                              int64_t n_edges, float* out, const int64_t* out_index, int32_t flags, void* stream);

/* ---------------------------------------------------------------------------------------------
 * f3: band-energy head, reciprocal-space assembly.  Replaces the dense per-k scatter of calculate_band_energies
 * (hamgnn/models/hamgnn_output.py:1775-1909: phase factors exp(2 pi i k.R), index_put_ accumulate into
 * [num_k, Na, Na, nao, nao], swapaxes, masked_select of the defined orbitals) for ONE crystal:
 *   hk / sk [n_k][n_orb][n_orb] complex64 (interleaved re, im), rows / columns = the defined orbitals of the crystal's atoms in
 *   atom-major order, orb_index[atom][o] = compact index of orbital o of that atom or -1.
 * hon / son [n_atoms][nao^2], hoff / soff [E][nao^2] fp32; src / dst: atom indices LOCAL to the crystal per edge; nbr_shift
 * [E][3] and kvec [n_k][3] in reciprocal units of each other (phase = 2 pi k.R).  seg_edge lists the edges grouped by (src, dst)
 * (any fixed order inside a group), seg_ptr[n_segs + 1] the group boundaries: one CTA adds the images of a pair in list order
 * (no atomics, bit-reproducible).  The generalized eigenproblem is solved by the caller (hamgnn_b200/band.py, cuSOLVER via
 * torch.linalg). */
int hgb_band_kspace(const float* hon, const float* hoff, const float* son, const float* soff, int64_t n_atoms, int32_t nao,
                    const int64_t* seg_ptr, int64_t n_segs, const int64_t* seg_edge, const int64_t* src, const int64_t* dst,
                    const float* nbr_shift, const float* kvec, int32_t n_k, const int32_t* orb_index, int32_t n_orb, float* hk,
                    float* sk, void* stream);

/* ---------------------------------------------------------------------------------------------
 * f4: graph construction on the device.  hgb_neighbor_list replaces neighbor_list_and_relative_vec
 * (hamgnn/models/base_model.py:87-178: ASE primitive_neighbor_list on the CPU with per-atom cutoffs) for one crystal:
 * directed edges i -> (j, S), S in [-reps, reps]^3, iff 0 < |pos_j + S.cell - pos_i| < radius_i + radius_j (fp64), sorted by
 * (i, j, S).  Two passes: deg != NULL counts the edges of every atom; after an exclusive scan (offset[N+1]) the second call
 * (deg == NULL) writes edge_index [2][E], cell_shift [E][3] (int64) and nbr_shift [E][3] = S.cell (fp32).
 * hgb_edge_lookup replaces find_matching_columns_of_A_in_B (:180-233) and the inverse-edge search of the data generator
 * (DFT_interfaces/openmx/graph_data_gen.py:293-295): out[e] = index of query edge e -- or of its inverse (dst, src, -S) -- in
 * the sorted graph, -1 if absent. */
int hgb_neighbor_list(const double* pos, const double* radius, const double* cell_host, const int32_t* reps_host, int64_t n_atoms,
                      int64_t* deg, const int64_t* offset, int64_t n_edges, int64_t* edge_index, int64_t* cell_shift, float* nbr_shift,
                      void* stream);
int hgb_edge_lookup(const int64_t* q_index, const int64_t* q_shift, int64_t n_query, int32_t inverse, const int64_t* g_index,
                    const int64_t* g_shift, const int64_t* g_offset, int64_t n_graph_edges, int64_t* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a9, deterministic receiver reduction for message kernels that write one row per edge (the 'rot' backend): replaces
 * torch_scatter.scatter(messages, receiver, reduce='sum') (hamgnn/nn/convolution.py:147-149).  out[i][:] = sum over
 * j in [seg_ptr[i], seg_ptr[i+1]) of rows[seg_order[j]][:], added in list order (no atomics: bit-reproducible). */
int hgb_segment_sum(const float* rows, int32_t n_cols, const int64_t* seg_ptr, const int64_t* seg_order, int64_t n_out_rows,
                    float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Measurement aid: per-kernel device time of the library's launches, CUDA events on the launching
 * stream.  hgb_timing_enable(1) starts recording, hgb_timing_collect adds elapsed ms / launch counts per kernel id to the
 * caller's arrays of HGB_N_KERNEL_IDS entries (synchronises the recorded events) and clears the records. */
enum {
  HGB_K_RADIAL_GATE = 0, HGB_K_ROTATE_PACK = 1, HGB_K_MSGPACK_ROT2 = 2, HGB_K_UNROTATE = 3, HGB_K_WIGNER = 4,
  HGB_K_EDGE_EMBED = 5, HGB_K_LINEAR = 6, HGB_K_RESBLOCK = 7, HGB_K_HAM_ASSEMBLE = 8, HGB_K_HAM_FINALIZE = 9,
  HGB_K_OTHER = 10, HGB_K_MSGPACK_ROT = 11, HGB_K_SEGMENT_SUM = 12, HGB_N_KERNEL_IDS = 16
};
int hgb_timing_enable(int32_t on);
int hgb_timing_collect(float* ms_out, int64_t* count_out);
/* Diagnosis tool (csrc/mma_probe.cu): cycles per tcgen05.mma.kind::tf32 (M 128, N n, K 8) issued back to back by one thread,
 * out_dev[0] = issue cycles, [1] = cycles until the commit barrier fires, [2] = a concurrent tcgen05.ld round trip. */
int hgb_mma_probe(int32_t n, int32_t count, int32_t ndest, int32_t ts, int32_t same_ab, long long* out_dev, void* stream);
/* cycles per cp.async.bulk of `bytes` (one thread, `depth` copies in flight, L2-resident source): out_dev[0] pipelined, [1] issue. */
int hgb_tma_probe(const float* src_dev, int32_t bytes, int32_t count, int32_t depth, long long* out_dev, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a5-a9, A-stationary form of the rotated frame ("rot2", the default message path).  Same mathematics and same call
 * sites as hgb_msgpack_rot_forward (MessagePackBlock.forward hamgnn/nn/message_passing.py:191-231, ConvBlockE3's
 * scatter-sum hamgnn/nn/convolution.py:147-149, PairInteractionBlock hamgnn/nn/interaction_blocks.py:133-164); the
 * (path, m1) steps are regrouped so that every rotated input image is multiplied ONCE by the concatenated weights of all
 * paths that consume it:
 *   pass   = (output component m3, subset of output slots with sum of multiplicities <= 128): one CTA per (tile, pass),
 *            accumulators C'[slot][m3][w] of the pass in shared memory, written as columns [out_col0, out_col0 + ncols)
 *            of the per-edge aligned-frame row cp[e][rowstride];
 *   piece  = image X'_{block, m1} x [W_p1 | W_p2 | ...] -> B[128 x ncols] (GEMM1), gate per 8-column batch,
 *            per destination slot S = (B.g)[:, col0 : col0 + kcols] [L'_p1; L'_p2; ...] (GEMM2, K-concatenated);
 *   unrotate: out[row][slot][w][k] = sum over the row's edge segment, sum_m3 D^{l3}_e[m3][k] cp[e][ccol[slot][m3] + w]
 *            -- the receiver reduction is a serial sum over a receiver-sorted segment (deterministic, no atomics).
 */
#ifndef HGB_ROT2_GATE_GROUPS
#define HGB_ROT2_GATE_GROUPS 2
#endif
typedef struct {
  int32_t piece_begin, piece_end;     /* pieces of the pass                                                  */
  int32_t ncols;                      /* accumulator columns (<= 128)                                       */
  int32_t out_col0;                   /* first column of the pass in a cp row                               */
  int32_t stream_begin[HGB_ROT2_GATE_GROUPS]; /* gate stream of each gate-warp group (entries of `batches`)  */
  int32_t stream_end[HGB_ROT2_GATE_GROUPS];
} hgb_rot2_pass_t;

typedef struct {
  int32_t a_off;       /* float offset of the image X'_{block, m1} inside a packed tile (hgb_rot_block_t layout)   */
  int32_t w_off;       /* wbuf offset of the concatenated W: per chunk of 16 channels (hi | lo) [kc/4][ncols][4];
                          the w3j scale of every (path, m1, m3) step is folded into its columns                   */
  int32_t l_off;       /* wbuf offset of the piece's L' operands (all destination groups, contiguous)              */
  int32_t l_floats;    /* their size (<= 8192 floats)                                                              */
  int32_t gpf_begin;   /* gate blocks of the piece (hgb_rot2_gpf_t entries), pulled into L2 ahead of the gate warps         */
  int32_t dst_begin;   /* tensor-core destination groups (GEMM2)                                                   */
  int16_t kpad;        /* K of GEMM1 (multiple of 8)                                                               */
  int16_t ncols;       /* N of GEMM1 (multiple of 16, <= 96)                                                       */
  int16_t ndst;        /* may be 0: every destination of the piece applies L' on the FMA pipes                     */
  int16_t gpf_n;
} hgb_rot2_piece_t;

typedef struct {
  uint32_t off;        /* float offset inside the tile's gate block [branch][gstride][128]                          */
  uint32_t bytes;      /* contiguous bytes (columns x 512)                                                          */
} hgb_rot2_gpf_t;

/* Gate-stream entry: 8 consecutive B columns of a piece, processed by one gate-warp half (thread = edge):
 *   b_j = B[z][8 col8 + j] * g[z][block(j)]   (g = 1 where the block offset is 0xFFFFFFFF: un-gated or padding columns)
 *   kind 0 (tensor): b is split hi/lo and written back to TMEM, GEMM2 multiplies by the (hi | lo) L' stack;
 *   kind 1 (FMA pipes, slots with multiplicity <= 16, paths packed at 4-column granularity): s[w'] += b_j L'[j][w'] with
 *            plain fp32 L' rows [8][round4(mul)] at l_off inside the piece's L' block; group-first zeroes s, group-last
 *            adds s into the accumulator columns acc_col0 .. acc_col0 + mul (each such slot belongs to ONE warp half, so its
 *            columns are updated in a fixed order);
 *   kind 2 (dummy): the half has no columns in this piece but still waits for GEMM1 and signals GEMM2.
 * The radial gate of a chunk is laid out [tile][branch][gstride][128 edges] (hgb_radial_gate layout 2). */
typedef struct {
  int32_t meta;        /* kind | first-of-piece << 2 | last-of-piece << 3 | group-first << 4 | group-last << 5 | col8 << 8 |
                          mul << 16 | acc_col0 << 21                                                               */
  uint32_t goff_a;     /* float offset of the gate block of columns 0-3 inside the tile's gate block                */
  uint32_t goff_b;     /* ... of columns 4-7                                                                       */
  int32_t l_off;
} hgb_rot2_batch_t;

typedef struct {
  int16_t col0, kcols; /* B columns of the group = K of GEMM2 (multiples of 8)                                     */
  int16_t mp;          /* N of GEMM2 (padded multiplicity of the slot, multiple of 16, <= 64)                      */
  int16_t s_off;       /* column of the result inside the S buffer (64 columns)                                   */
  int16_t acc_col0;    /* accumulator column of the slot inside the pass                                           */
  int16_t mul;
  int32_t l_rel;       /* offset of the group's (hi | lo) L' stack [kcols/4][mp][4] inside the piece's L' block    */
} hgb_rot2_dst_t;

typedef struct {
  int32_t n_passes, n_pieces, n_batches, n_dsts;
  int32_t rowstride;                  /* floats per cp row = sum over slots of mul * (2l+1)                        */
  int32_t n_slots;
  int32_t slot_l[32], slot_mul[32], slot_out_off[32];
  int32_t ccol[32][13];               /* cp column of (slot, l3 + m3), -1: no contribution                        */
  const hgb_rot2_pass_t* passes;      /* device */
  const hgb_rot2_piece_t* pieces;     /* device */
  const hgb_rot2_batch_t* batches;    /* device */
  const hgb_rot2_dst_t* dsts;         /* device */
  const hgb_rot2_gpf_t* gpf;          /* device */
  int64_t n_gpf;
  const hgb_rot2_pass_t* passes_host;
  const hgb_rot2_piece_t* pieces_host;
  const hgb_rot2_batch_t* batches_host;
  const hgb_rot2_dst_t* dsts_host;
  const hgb_rot2_gpf_t* gpf_host;
} hgb_rot2_plan;

/* g_ws / xp_ws: per-chunk workspaces as in hgb_msgpack_rot_forward; cp_ws: n_edges * rowstride floats (all edges).
 * seg_ptr == NULL: out has n_edges rows, row e = message of edge e.  Otherwise out has n_out_rows rows and row i is
 * the sum of the messages of edges seg_order[seg_ptr[i] .. seg_ptr[i+1]) in that order (see receiver_segments). */
int hgb_msgpack_rot2_forward(const hgb_msgpack_plan* plan_host, const hgb_rot_plan* rot_host, const hgb_rot2_plan* rot2_host,
                             const float* const* src_host, const int64_t* const* src_rows_host, const float* dw,
                             const float* rbf, const int32_t* w3_off_host, const int32_t* nch_host,
                             const int32_t* w3img_off_host, int32_t gstride, float* g_ws, float* xp_ws, float* cp_ws,
                             int64_t chunk_edges, int64_t n_edges, float* out, const int64_t* seg_ptr,
                             const int64_t* seg_order, int64_t n_out_rows, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a10/a13 and the o3.Linear's: row-wise equivariant Linear -> Gate -> Linear (+residual) [-> Linear].
 * Replaces o3.Linear call sites (hamgnn/nn/convolution.py:112, interaction_blocks.py:126,306-309,
 * embeddings.py:286, toolbox/nequip/nn/_atomwise.py:51, hamgnn_output.py:49), ResidualBlock.forward
 * (hamgnn/nn/interaction_blocks.py:332-358, e3nn Gate with ssp/tanh), HamLayer.forward
 * (hamgnn/models/hamgnn_output.py:38-58).
 */
typedef struct {
  int32_t in_off, out_off, mul_in, mul_out, dim, w_off; /* out[r, out_off + w*dim + k] += sum_u in[r, in_off + u*dim + k] W[w_off + u*mul_out + w] */
} hgb_linblock_t;

typedef struct {
  int32_t n_blocks, in_dim, out_dim;
  int32_t pad;                   /* bit 0: every block writes its own output slot (single-pass evaluation) */
  const hgb_linblock_t* blocks;  /* device */
  const float* w;                /* device, pre-scaled by 1/sqrt(fan_in) */
} hgb_linear_plan;

/* y[r,:] = (accumulate ? y[r,:] : 0) + Linear(x[rows ? rows[r] : r, :]) */
int hgb_linear_forward(const hgb_linear_plan* plan_host, const float* x, const int64_t* rows, int64_t n_rows,
                       float* y, int32_t accumulate, void* stream);
/* same with an explicit output row stride ldy >= plan->out_dim (floats): writes columns [0, out_dim) of rows of a
 * wider matrix -- used to evaluate a wide head Linear (SOC: hamgnn_output.py:190-198) in column chunks. */
int hgb_linear_forward_ld(const hgb_linear_plan* plan_host, const float* x, const int64_t* rows, int64_t n_rows,
                          float* y, int64_t ldy, int32_t accumulate, void* stream);

typedef struct {
  /* e3nn Gate: input row = sorted/simplified (scalars | gates | gated), output row = scalars + gated */
  int32_t n_scalar_slots;
  int32_t sc_in_off[4], sc_out_off[4], sc_n[4], sc_act[4];  /* act: 0 = ssp (even), 1 = tanh (odd) */
  int32_t n_gated;
  int32_t gd_in_off[16], gd_out_off[16], gd_mul[16], gd_dim[16], gd_gate_off[16]; /* gate columns in the input row */
  float c_ssp, c_tanh;             /* normalize2mom constants */
  int32_t in_dim, out_dim;
} hgb_gate_desc;

/* y = x + Lin2(Gate(Lin1(x))) [+ extra]  ; if post != NULL: y = post(y).  x,y: [n_rows, D]. */
int hgb_resblock_forward(const hgb_linear_plan* lin1_host, const hgb_gate_desc* gate_host,
                         const hgb_linear_plan* lin2_host, const hgb_linear_plan* post_host,
                         const float* x, const float* extra, int64_t n_rows, float* y, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a14+a15: Clebsch-Gordan assembly of orbital blocks, orbital reorder, (anti-)symmetrisation with the
 * inverse edge, +H0, orbital masks, interleaved per-crystal output rows.
 * Replaces merge_tensor_components (hamgnn/models/hamgnn_output.py:851-891), reorder_matrix (:1056-1096),
 * symmetrize_hamiltonian (:1231-1285), apply_orbital_masks_to_hamiltonians (:2288-2365),
 * concatenate_hamiltonians_by_crystal (:1187-1229) and the +Hon0/+Hoff0 of forward (:3782-3795).
 */
typedef struct {
  int32_t nao, n_coef, nnz, pad;
  const int32_t* row_ptr;   /* device [nao*nao+1], CSR over the reordered flattened (a,b) entries */
  const int32_t* col;       /* device [nnz] coefficient column                                   */
  const float* val;         /* device [nnz] sqrt(2L+1) * w3j                                     */
  const uint8_t* orb_mask;  /* device [128][nao] 1 if the orbital exists for element Z            */
} hgb_ham_plan;

/* raw[r, a*nao+b] = sum_nnz val * coef[r, col]  (CG merge + reorder) */
int hgb_ham_assemble(const hgb_ham_plan* plan_host, const float* coef, int64_t n_rows, float* raw, void* stream);

/* out[out_row[r], :] = mask(z[na[r]], z[nb[r]]) * ( (sym ? 0.5*(raw[r] + raw[partner[r]]^T) : raw[r]) + (h0 ? h0[r] : 0) )
 * partner == NULL -> on-site (partner = r); node_a/node_b == NULL -> identity (row r is atom r). */
int hgb_ham_finalize(const hgb_ham_plan* plan_host, const float* raw, const int64_t* partner, const float* h0,
                     const int64_t* z, const int64_t* node_a, const int64_t* node_b, const int64_t* out_row,
                     int64_t n_rows, int32_t symmetrize, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a16: spin-orbit-coupling assembly (complex H, (2 nao)^2 spin-orbital blocks, real and imaginary planes).
 *
 * su2 basis -- replaces E3TensorDecomposition.get_H (hamgnn/nn/tensor_decomposition.py:575-627, spinful) +
 * reorder_matrix + the (spin, orbital) interleave, symmetrize_{on,off}site_hamiltonian_soc, the per-spin-block
 * orbital masks and +H0/+iH0 of HamGNNPlusPlusOut.forward (hamgnn/models/hamgnn_output.py:3146-3178, 3603-3608):
 *   1. hgb_csr_rows applies the host-built sparse map (w3j(L,1,L') recoupling x w3j(l1,l2,L) x oyzx2spin, reorder
 *      folded in) to the used head outputs:  y[r, o] = sum_nnz val * x[r, col],  y = [rows][2][2nao][2nao];
 *   2. hgb_ham_finalize_su2:  H = sym ? 0.5 (raw[r] + conj(raw[partner[r]])^T) : raw[r];  zero where an orbital is
 *      absent for z[node_a[r]] (rows) / z[node_b[r]] (columns);  += h0_re / h0_im;  written to row out_row[r] of
 *      out_re / out_im ([*, (2nao)^2] each).
 * so3 basis -- replaces the xi.L construction (hamgnn_output.py:3026-3144) and symmetrize_orbital_coefficients
 * (:2367-2431): hgb_ksi_shell_average averages ksi over the m components of each shell [lo, hi) (rows, then
 * columns, in place); hgb_ham_finalize_so3 builds  re: uu = dd = hns, ud = du = A_1;  im: uu = A_2, dd = -A_2,
 * ud = A_0, du = -A_0  with A_c = sym ? 0.5 (ksi L_c - (ksi L_c)[partner]^T) : ksi L_c, lmat = [rows][nao^2][3];
 * h0_offdiag_only != 0 skips h0_re on the uu/dd blocks (add_H_nonsoc, :3028-3049).
 */
int hgb_csr_rows(const int32_t* row_ptr, const int32_t* col, const float* val, int32_t n_out, int32_t n_in,
                 const float* x, int64_t n_rows, float* y, void* stream);
int hgb_ham_finalize_su2(int32_t nao, const uint8_t* orb_mask, const float* raw, const int64_t* partner,
                         const float* h0_re, const float* h0_im, const int64_t* z, const int64_t* node_a,
                         const int64_t* node_b, const int64_t* out_row, int64_t n_rows, int32_t symmetrize,
                         float* out_re, float* out_im, void* stream);
int hgb_ksi_shell_average(int32_t nao, const int32_t* blk_lo_host, const int32_t* blk_hi_host, int32_t n_blocks,
                          float* ksi, int64_t n_rows, void* stream);
int hgb_ham_finalize_so3(int32_t nao, const float* hns, const float* ksi, const float* lmat, const int64_t* partner,
                         const float* h0_re, const float* h0_im, const int64_t* out_row, int64_t n_rows,
                         int32_t symmetrize, int32_t h0_offdiag_only, float* out_re, float* out_im, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Self-test of the tcgen05 3xTF32 GEMM building block (TMEM accumulator, interleaved K-major shared-memory
 * operands, or A staged in TMEM when a_from_tmem != 0): C[t] (128 x N) = A[t] (128 x K) . B (K x N), t < tiles;
 * K % 8 == 0, N <= 64.  No reference
 * counterpart -- it validates the tensor-core path that replaces the dense per-path contractions of
 * o3.TensorProduct / o3.Linear (hamgnn/nn/message_passing.py:81-134) against a plain matmul.
 */
int hgb_tc_gemm_selftest(const float* A, const float* B, float* C, int32_t tiles, int32_t K, int32_t N,
                         int32_t a_from_tmem, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HAMGNN_B200_H */
