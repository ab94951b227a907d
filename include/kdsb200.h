/*
 * kdsb200.h -- C ABI of libkdsb200.so, the B200 (sm_100a) kernel-density hot
 * path of KDSource.  Plain pointers and sizes only.  Unless stated otherwise
 * every `d_*` pointer is a DEVICE pointer owned by the caller, `h_*` is a host
 * pointer, `stream` is a cudaStream_t passed as void*, and every function
 * returns 0 on success / non-zero on failure with the reason available from
 * kdsb200_last_error() (thread-local).  No function synchronises the stream
 * unless documented.
 *
 * Each entry point cites the reference interface (paths relative to the
 * KDSource tree) whose arithmetic it replaces.
 */
#ifndef KDSB200_H
#define KDSB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ---------------------------------------------------------- */
const char* kdsb200_last_error(void);
int kdsb200_version(void);
int kdsb200_device_info(int* sm_count, int* cc_major, int* cc_minor, int* clock_khz);
/* Pipe-rate calibration (mode 0 FFMA, 1 FFMA2, 2 MUFU.EX2): lane-ops/s measured
 * with CUDA events; the roofline denominators of bench.py.  Synchronises. */
int kdsb200_calibrate_pipe(int mode, float* d_scratch, int iters, double* ops_per_s,
                           void* stream);

/* ---- source / query packing ------------------------------------------ */
/* Number of floats of the packed fp32 source set: tiles of 512 sources x (D+2)
 * rows (see csrc/common.cuh). */
uint64_t kdsb200_packed_floats(uint64_t N, int D);
/* Query count padded to the evaluation kernel's CTA tile. */
uint64_t kdsb200_eval_mpad(uint64_t M);
/* d_data (N,D) row-major fp64 = KDEpy `kde.data` (vecs/scaling, kdsource.py:212);
 * d_w (N) or NULL; d_bw (N) or NULL (-> bw_scalar) = `kde.bw`;
 * mode 0: kde_eval layout, mode 1: mlcv layout. */
int kdsb200_pack_sources_f32(const double* d_data, const double* d_w, const double* d_bw,
                             double bw_scalar, uint64_t N, int D, const double* h_center,
                             double w_scale, double lc_ref, int mode, float* d_packed,
                             void* stream);
int kdsb200_pack_queries_f32(const double* d_pts, uint64_t M, uint64_t Mpad, int D,
                             const double* h_center, float* d_qT, void* stream);

/* ---- evaluation: KDEpy TreeKDE.evaluate (kdsource.py:239, kde.py:146-151) */
/* Source-dimension split that fills the GPU for (N, M). */
int kdsb200_eval_nsplit(uint64_t N, uint64_t M);
/* out[m] = out_scale * sum_i 2^-nlc_i exp(-|q_m-x_i|^2/(2 h_i^2)); cutoff > 0
 * drops sources farther than `cutoff` (hard support radius).  d_partial holds
 * nsplit*Mpad doubles. */
int kdsb200_kde_eval_f32(const float* d_packed, uint64_t N, int D, const float* d_qT, uint64_t M,
                         uint64_t Mpad, double cutoff, double out_scale, int nsplit,
                         double* d_partial, double* d_out, void* stream);
/* Same with a kernel selector -- the kernel names the reference hands to KDEpy
 * (kdsource.py:60-65): 0 gaussian (as above), 1 epa, 2 box.  For 1 and 2 the sources
 * are packed with mode 1 and the bandwidth replaced by the support radius
 * R_i = h_i*sqrt(5) (epa) / h_i*sqrt(3) (box), KDEpy's unit-variance scaling; then
 * out[m] = out_scale * sum_i c_i * max(0, 1-u) (epa) or c_i*[u<1] (box),
 * u = |q_m-x_i|^2/R_i^2; `cutoff` is ignored (the support is finite). */
int kdsb200_kde_eval_kernel_f32(const float* d_packed, uint64_t N, int D, const float* d_qT,
                                uint64_t M, uint64_t Mpad, int kernel, double cutoff,
                                double out_scale, int nsplit, double* d_partial, double* d_out,
                                void* stream);
/* Split-coordinate fp32 mode (Gaussian kernel): every recentred coordinate is stored as
 * hi + lo with hi a multiple of 1/grid (grid a power of two with max|x - c| * grid < 2^22),
 * so the difference q - x is formed like in fp64 and the error no longer grows with
 * |x - c| / h.  Sources: tiles of 512 x (2D+2) rows; queries d_qT [2D][Mpad]. */
uint64_t kdsb200_packed_split_floats(uint64_t N, int D);
int kdsb200_pack_sources_split_f32(const double* d_data, const double* d_w, const double* d_bw,
                                   double bw_scalar, uint64_t N, int D, const double* h_center,
                                   double w_scale, double lc_ref, double grid, float* d_packed,
                                   void* stream);
int kdsb200_pack_queries_split_f32(const double* d_pts, uint64_t M, uint64_t Mpad, int D,
                                   const double* h_center, double grid, float* d_qT,
                                   void* stream);
int kdsb200_kde_eval_split_f32(const float* d_packed, uint64_t N, int D, const float* d_qT,
                               uint64_t M, uint64_t Mpad, double out_scale, int nsplit,
                               double* d_partial, double* d_out, void* stream);
/* ---- spatially ordered tiles: exact tile culling + tile-local coordinates --------------
 * The work saving of the reference's TreeKDE (only the sources inside the support radius
 * are summed, kdsource.py:238-240, kde.py:146-151) without its approximation: sources are
 * ordered along a Hilbert curve, each tile of 512 keeps its bounding box and a centre,
 * coordinates are stored relative to that centre, and the kernel skips every tile whose
 * box is farther than the cutoff from the box of a CTA's 1024 (equally ordered) queries. */
uint64_t kdsb200_minmax_scratch(void); /* doubles */
/* d_minmax[0..D) = column minima, [D..2D) = maxima of d_data (N,D) */
int kdsb200_column_minmax(const double* d_data, uint64_t N, int D, double* d_scratch,
                          double* d_minmax, void* stream);
uint64_t kdsb200_order_scratch_bytes(uint64_t N);
int kdsb200_hilbert_bits(int D);
/* d_perm[i] = index of the i-th point along the Hilbert curve over the box [h_lo, h_hi]
 * quantised to kdsb200_hilbert_bits(D) bits per dimension (points outside are clamped);
 * stable radix sort, ties keep the input order.  N < 2^32. */
int kdsb200_hilbert_order(const double* d_data, uint64_t N, int D, const double* h_lo,
                          const double* h_hi, void* d_scratch, uint32_t* d_perm, void* stream);
/* Sources in d_perm order, tiles as kdsb200_pack_sources_f32 (same `mode`) but with
 * coordinates relative to the tile centre d_tilecen[t][D] (multiples of 1/grid, grid a
 * power of two); d_tilebox[t] (kdsb200_sorted_tilebox_floats(D) floats per tile) =
 * {lo[D], hi[D]} of x - h_center over the tile, then {lo[D], hi[D], hmax} of each of its
 * sixteen runs of 32 sources; d_tilehmax[t] = the tile's largest bandwidth (as passed in
 * d_bw / bw_scalar). */
int kdsb200_sorted_tilebox_floats(int D);
int kdsb200_pack_sources_sorted_f32(const double* d_data, const double* d_w, const double* d_bw,
                                    double bw_scalar, uint64_t N, int D, const uint32_t* d_perm,
                                    const double* h_center, double grid, double w_scale,
                                    double lc_ref, int mode, float* d_packed, float* d_tilebox,
                                    float* d_tilecen, float* d_tilehmax, void* stream);
uint64_t kdsb200_sorted_qpad(uint64_t M); /* query count padded to the CTA tile (1024) */
/* queries in d_perm order as hi + lo floats, hi on the grid: d_qT [2D][Mpad] */
int kdsb200_pack_queries_sorted_f32(const double* d_pts, uint64_t M, uint64_t Mpad, int D,
                                    const uint32_t* d_perm, const double* h_center, double grid,
                                    float* d_qT, void* stream);
int kdsb200_sorted_ctas(void);
int kdsb200_sorted_nsplit(uint64_t N, uint64_t M);
uint64_t kdsb200_sorted_scratch_bytes(uint64_t N, uint64_t Mpad, int nsplit);
/* d_out[d_perm_q[m]] = out_scale * sum over sources, as kdsb200_kde_eval_kernel_f32.
 * kernel 0: cutoff > 0 = one hard radius for every source (KDEpy's TreeKDE derives it from
 * the largest bandwidth); cutoff_rel > 0 = a hard radius cutoff_rel * h_i per source, tiles
 * culled at cutoff_rel times their largest bandwidth; both <= 0 = the exact sum over every
 * tile.  kernel 1/2: culled at the tile's largest support radius.  Culling is exact: a
 * skipped tile holds no pair inside the radius and pairs inside visited tiles are still
 * masked one by one; inside a visited tile every cell of 32 consecutive sources x 256
 * consecutive queries is tested (and skipped) the same way.  d_tiles_visited (device, two
 * uint64, may be NULL) = {tiles loaded, cells computed}, summed over (1024-query block,
 * source split) work items: executed pairs = cells * 32 * 256. */
int kdsb200_kde_eval_sorted_f32(const float* d_packed, const float* d_tilebox,
                                const float* d_tilecen, const float* d_tilehmax, uint64_t N,
                                int D, const float* d_qT, const uint32_t* d_perm_q, uint64_t M,
                                uint64_t Mpad, int kernel, double cutoff, double cutoff_rel,
                                double out_scale, int nsplit, void* d_scratch, double* d_out,
                                uint64_t* d_tiles_visited, void* stream);
/* fp64 mode (1e-12 parity).  d_scratch: 2N + nsplit*M doubles. */
int kdsb200_kde_eval_kernel_f64(const double* d_data, const double* d_w, const double* d_bw,
                                double bw_scalar, uint64_t N, int D, double w_scale,
                                const double* d_pts, uint64_t M, int kernel, double cutoff,
                                int nsplit, double* d_scratch, double* d_out, void* stream);
int kdsb200_kde_eval_f64(const double* d_data, const double* d_w, const double* d_bw,
                         double bw_scalar, uint64_t N, int D, double w_scale,
                         const double* d_pts, uint64_t M, double cutoff, int nsplit,
                         double* d_scratch, double* d_out, void* stream);

/* ---- MLCV: _kde_cv_score x grid (kde.py:106-153, :204-211) --------------- */
int kdsb200_mlcv_gtile(int G);          /* grid slots the kernel computes for G factors */
uint64_t kdsb200_mlcv_mpad(uint64_t M); /* test-point count padded to the CTA tile */
int kdsb200_mlcv_nsplit(uint64_t N, uint64_t M);
/* how many of the gtile(G) slots evaluate 2^x on the FMA pipe instead of MUFU.EX2
 * (for roofline accounting; the result does not depend on it beyond 1e-6) */
int kdsb200_mlcv_poly_slots(int G);
/* Sources packed with mode 1.  Test points d_qT [D][Mpad]; per test point the
 * excluded source interval [lo,hi) (its fold; [i,i+1) for leave-one-out), the
 * coefficient coefw = w_i / (n_fold * K) and lwtr = ln(sum of train weights).
 * h_nk[g] = -log2(e) / (2 grid[g]^2), G <= 32 per call.
 * d_partial[(Mpad/256)][gtile]: per-CTA sums of coefw*(ln S_ig - lwtr), in fixed
 * order (deterministic).  d_sbuf: nsplit*Mpad*gtile doubles, needed if nsplit>1.
 * d_flagmask (Mpad uint32, may be NULL): bit g of entry m is set when S_mg fell into
 * the window where ex2.approx.ftz has flushed terms to zero (S < 2^-70 in the scaled
 * units, i.e. the nearest training source is ~10 bandwidths or more away); such a
 * (row, g) adds 0 to d_partial and must be recomputed by kdsb200_mlcv_rows_f64. */
int kdsb200_mlcv_scores_f32(const float* d_packed, uint64_t N, int D, const float* d_qT,
                            uint64_t M, uint64_t Mpad, const int32_t* d_excl_lo,
                            const int32_t* d_excl_hi, const double* d_coefw,
                            const double* d_lwtr, const float* h_nk, int G, int nsplit,
                            double* d_sbuf, double* d_partial, uint32_t* d_flagmask,
                            void* stream);
/* fp64 rows (kde.py:106-153 is fp64 in the reference): the 1e-12 mode of the search and
 * the recomputation of flagged rows.  Test rows are rows [row0, row0+M) of d_data (N,D);
 * S_mg = sum_{j not in [lo_m,hi_m)} w_j w_scale h_j^-D exp(-|x_m-x_j|^2/(2 grid_g^2 h_j^2));
 * d_partial[kdsb200_mlcv_f64_blocks(M)][kdsb200_mlcv_f64_gtile(G)] receives per-CTA sums
 * of coefw_m (ln S_mg + log_shift - lwtr_m), restricted to the bits of d_rowmask[m] when
 * d_rowmask != NULL.  d_coef2n: 2N doubles filled by kdsb200_mlcv_coef_f64 (w_j w_scale h_j^-D,
 * then 1/(2 h_j^2)). */
uint64_t kdsb200_mlcv_f64_blocks(uint64_t M);
int kdsb200_mlcv_f64_gtile(int G);
int kdsb200_mlcv_coef_f64(const double* d_w, const double* d_bw, double bw_scalar, uint64_t N,
                          int D, double w_scale, double* d_coef2n, void* stream);
/* d_fix_scratch: kdsb200_mlcv_fix_scratch_bytes(M) bytes, needed when d_rowmask != NULL
 * (the first 256 row blocks with flagged rows are recomputed by 64 CTAs each, over
 * slices of the list, and combined in slice order). */
uint64_t kdsb200_mlcv_fix_scratch_bytes(uint64_t M);
int kdsb200_mlcv_rows_f64(const double* d_data, const double* d_coef2n, uint64_t N, int D,
                          uint64_t row0, uint64_t M, const int32_t* d_excl_lo,
                          const int32_t* d_excl_hi, const double* d_coefw, const double* d_lwtr,
                          const double* h_grid, int G, double log_shift,
                          const uint32_t* d_rowmask, void* d_fix_scratch, double* d_partial,
                          void* stream);
/* Fold bookkeeping of the test rows [row0, row0+M) built on the device: sklearn
 * KFold(n_splits=K) without shuffle (kde.py:134-136).  d_cwf[f] = 1/(n_f K),
 * d_lwt[f] = ln(training weight of fold f), d_w (N) or NULL.  Outputs have Mpad entries. */
int kdsb200_mlcv_rows_setup(uint64_t N, uint64_t K, uint64_t row0, uint64_t M, uint64_t Mpad,
                            const double* d_w, const double* d_cwf, const double* d_lwt,
                            int32_t* d_lo, int32_t* d_hi, double* d_coefw, double* d_lwtr,
                            void* stream);
/* d_out[g] (g < G) = sum of the rows of d_a[na][gt_a] and, if given, d_b[nb][gt_b]; fixed
 * summation order (128 CTAs over contiguous shares, then one): the on-stream reduction
 * ahead of the NCCL all-reduce of the G partial log-likelihoods.  d_tmp:
 * kdsb200_mlcv_reduce_scratch() doubles. */
uint64_t kdsb200_mlcv_reduce_scratch(void);
int kdsb200_mlcv_reduce(const double* d_a, uint64_t na, int gt_a, const double* d_b, uint64_t nb,
                        int gt_b, int G, double* d_tmp, double* d_out, void* stream);

/* ---- kNN: NearestNeighbors(k+1).kneighbors inside batches (kde.py:91-98) -- */
/* d_data (N,D) fp64 row-major; d_batch_off: nb+1 int64 offsets (np.array_split
 * bounds); kk = k+1 (self included).  use_f64 != 0 reproduces numpy/sklearn
 * distances bit for bit; 0 is the fp32 fast mode.  d_scratch: N*D elements of
 * the compute type.  Outputs (each may be NULL): d_kth (N) distance to the
 * kk-th neighbour, d_dist (N,kk) ascending, d_idx (N,kk) index inside the batch,
 * ordered by (distance, index). */
int kdsb200_knn(const double* d_data, uint64_t N, int D, const int64_t* d_batch_off, int nb,
                int64_t max_batch, int kk, int use_f64, void* d_scratch, double* d_kth,
                double* d_dist, int32_t* d_idx, void* stream);

/* ---- geometry: Geometry.transform / Geometry.jac (python/kdsource/geom.py:158-205)
 * d_parts (M,8) fp64 [ekin,x,y,z,dx,dy,dz,t] -> d_vecs (M,D) fp64 = transformed
 * variables times h_inv_scaling (NULL: 1), and d_jac (M) (may be NULL).  The
 * geometry's translation is taken from h_geom; its rotation from h_rot_inv, the
 * row-major inverse rotation matrix (required when h_geom->has_rot). */
int kdsb200_geom_transform(const double* d_parts, uint64_t M, const struct kdsb200_geom_s* h_geom,
                           const double* h_rot_inv, const double* h_inv_scaling, double* d_vecs,
                           double* d_jac, void* stream);

/* Weighted column moments on the device (Metric.mean / Metric.std, geom.py:61-70; N_eff,
 * plist.py:349-395): d_out[k] = sum_i w_i (v_ik - c_k)^power (power 1 or 2; h_center NULL:
 * c = 0), d_out[D] = sum w_i, d_out[D+1] = sum w_i^2; d_w NULL: unit weights.  Fixed grid
 * and summation order (reproducible).  d_scratch: kdsb200_moments_scratch() doubles. */
uint64_t kdsb200_moments_scratch(void);
int kdsb200_weighted_moments(const double* d_vecs, const double* d_w, uint64_t N, int D,
                             const double* h_center, int power, double* d_scratch, double* d_out,
                             void* stream);

/* ---- resampling: KDS_rand_sample2 loop (kdsource.c:282-330, resamplemcpl.c:40-43)
 * with the metric perturbations of geom.c:192-650 ------------------------- */
#define KDSB200_MAX_METRICS 8
typedef struct {
  int32_t type;      /* index into _metric_names, geom.c:13-16 */
  int32_t dim;
  float scaling[6];  /* float, as Metric_create stores it (geom.c:35-38) */
  int32_t nps;
  double params[8];
} kdsb200_metric;
typedef struct kdsb200_geom_s {
  int32_t order;
  kdsb200_metric ms[KDSB200_MAX_METRICS];
  char kernel;       /* 'g' | 'e' | 'b' */
  int32_t has_trasl, has_rot;
  double trasl[3], rot[3];
} kdsb200_geom;

uint64_t kdsb200_cdf_scratch_words(uint64_t N);
/* Fixed-point CDF: c_i = floor(w_i * 2^62 / sum_w) (>=1 if w_i>0), inclusive
 * uint64 prefix sums (order independent, hence bit-reproducible). */
int kdsb200_weights_cdf(const double* d_w, uint64_t N, double sum_w, uint64_t* d_scratch,
                        uint64_t* d_cdf, void* stream);
/* Optional search accelerator: d_guide[b] (2^bits + 1 uint32) = first i with
 * cdf[i] > (b << (62-bits)); narrows the binary search window, same result. */
int kdsb200_cdf_guide(const uint64_t* d_cdf, uint64_t N, int bits, uint32_t* d_guide, void* stream);
/* Pick table (guide_kind 1 of kdsb200_resample): the window table with the answer inlined.
 * 2^bits entries of 128 bytes (one HBM line, fetched by eight lanes as one request) --
 * {cdf[i0], cdf[i0+1], i0, i1 - i0, the fp32 particles i0 and i0+1 and their bandwidths}
 * -- so a pick whose target lies below the bucket's second CDF boundary costs one line
 * in all; other targets continue in cdf[i0+2 .. i1].  Same index as the plain search.
 * bits ~ log2(2 N) keeps the search to ~1 pick in 30.  d_src8_f32 is the fp32 store of
 * kdsb200_convert_particles; d_bw may be NULL (bw_scalar is stored). */
uint64_t kdsb200_pick_table_bytes(int bits);
int kdsb200_pick_table(const uint64_t* d_cdf, uint64_t N, int bits, const float* d_src8_f32,
                       const float* d_bw, double bw_scalar, void* d_table, void* stream);
/* (N,8) fp64 particles [ekin,x,y,z,ux,uy,uz,t] -> fp32 or fp64 device store */
int kdsb200_convert_particles(const double* d_src8, uint64_t N, int to_f32, void* d_out,
                              void* stream);
/* Output particle n (global index first+n): Philox4x32-10 counter (index, block),
 * key = seed; pick = first i with cdf[i] > mulhi64(R64, total); then Geom_perturb.
 * d_guide/guide_bits/guide_kind: table from kdsb200_cdf_guide (kind 0) or
 * kdsb200_pick_table (kind 1: fp32 store, Philox mode, specialised geometries), or NULL
 * for the plain search.
 * seq_start >= 0 replaces the pick by the list walk i = (seq_start + first + n) % N
 * (the w_crit <= 0 mode of KDS_rand_sample2); pass -1 for weighted sampling.
 * d_ext_uniforms (count x slots doubles) replaces Philox (test mode).
 * contract: how the Philox words turn into Gaussian deviates.  1 = the reference's own
 * map (Marsaglia polar method, second value discarded, utils.c:9-29; ~12 words per
 * particle, rejection loops); 2 = Box-Muller with both values of a pair used (8 words
 * per particle of the usual geometries, no rejection).  External uniforms always follow
 * the reference's map.  Same distribution, different particles for a given seed.
 * Outputs (each may be NULL): 36-byte MCPL-3 single-precision records, picked
 * source index, full-precision particle (8 doubles). */
/* 1 if the geometry has a specialised fp32 kernel (Lethargy+SurfXY+Isotrop or
 * Energy+Vol+Isotrop, Gaussian kernel, no translation/rotation) whose maps hold every
 * coordinate within 1e-5 of its scale; other geometries go through acos/atan2-based
 * maps that fp32-rounded sources cannot hold to 1e-5 and should use the fp64 store. */
int kdsb200_resample_fp32_ok(const kdsb200_geom* h_geom);
int kdsb200_resample(const void* d_src8, int src_is_f32, const uint64_t* d_cdf,
                     const void* d_guide, int guide_bits, int guide_kind, const float* d_bw,
                     double bw_scalar, int64_t N, const kdsb200_geom* h_geom, int32_t pdgcode,
                     uint64_t seed, uint64_t first, uint64_t count, int64_t seq_start,
                     const double* d_ext_uniforms, int slots, int perturb, int contract,
                     void* d_out36, int64_t* d_out_idx, double* d_out_parts, void* stream);

/* ---- MCPL-3 records -> particles on the device: the decode step of PList_get /
 * mcpl_read (clib/src/plist.c:69-89, python/kdsource/plist.py:349-395) ----------- */
typedef struct {
  int32_t record_bytes;     /* particle size from the file header */
  int32_t single_precision; /* 1: float fields, 0: double */
  int32_t off_pos;          /* byte offsets inside a record: position[3] */
  int32_t off_dir;          /* packed direction (2 values) followed by signed ekin */
  int32_t off_time;
  int32_t off_weight;       /* -1: universal weight */
  int32_t off_pdg;          /* -1: universal pdg code */
  int32_t universal_pdg;
  double universal_weight;
} kdsb200_mcpl_layout;
/* d_raw (n records, little endian) -> d_parts8 (n,8) fp64 [ekin,x,y,z,ux,uy,uz,t],
 * d_weights (n) fp64, d_pdg (n) int32.  Bit-identical to the host decoder. */
int kdsb200_mcpl_unpack(const void* d_raw, uint64_t n, const kdsb200_mcpl_layout* h_layout,
                        double* d_parts8, double* d_weights, int32_t* d_pdg, void* stream);

/* ---- gzip on the device: the output side of kdsource_resample_to_mcpl, which
 * always ends in mcpl_closeandgzip_outfile (clib/src/resamplemcpl.c:30-46) ------
 * The byte stream is cut into members of `member_bytes`; each becomes one gzip
 * member (RFC 1952) with a single dynamic-Huffman deflate block of literals
 * (RFC 1951).  Members are laid out back to back (each a multiple of 4 bytes), which
 * is a valid multi-member .gz file for zlib's gzread (libmcpl), gzip(1), Python. */
typedef struct {
  uint32_t code[257]; /* bit-reversed canonical Huffman code of byte b / end-of-block (256) */
  uint8_t len[257];   /* code lengths, 1..15 */
  uint32_t hdr_bits;  /* length of the deflate block header (BFINAL..code lengths) */
  uint32_t hdr[64];   /* the block header, packed LSB first */
  /* record mode: when the last tail_bytes of a record equal the previous record's, they
   * are sent as one LZ77 match (length tail_bytes, distance rec_bytes) of match_nbits bits */
  uint32_t rec_bytes, tail_bytes, match_bits, match_nbits;
} kdsb200_gz_table;
/* Histogram of the symbols the encoder will emit for d_data (258 uint64 counters on the
 * device): bytes 0..255, [256] unused, [257] = records whose tail repeats.  rec_bytes = 0:
 * plain byte stream (no matches); otherwise rec_bytes and tail_bytes are multiples of 4,
 * member_bytes a multiple of 256 records and nbytes whole records. */
int kdsb200_gz_histogram(const void* d_data, uint64_t nbytes, uint32_t member_bytes,
                         uint32_t rec_bytes, uint32_t tail_bytes, uint64_t* d_hist258,
                         void* stream);
/* host: Huffman table + block header from that histogram (every byte value stays
 * encodable, so the table is valid for data the histogram has not seen) */
int kdsb200_gz_build_table(const uint64_t* h_hist258, uint32_t rec_bytes, uint32_t tail_bytes,
                           kdsb200_gz_table* out);
uint32_t kdsb200_gz_member_bytes(void); /* default member size: 2048 MCPL-3 records */
uint64_t kdsb200_gz_scratch_bytes(uint64_t nbytes, uint32_t member_bytes);
/* d_data (nbytes, multiple of 4) -> d_out (<= out_capacity bytes, multiple of 4);
 * *d_total = compressed size (device); if it exceeds out_capacity the output is
 * incomplete and the caller must retry with a larger buffer. */
int kdsb200_gz_compress(const void* d_data, uint64_t nbytes, uint32_t member_bytes,
                        const kdsb200_gz_table* h_table, void* d_scratch, void* d_out,
                        uint64_t out_capacity, uint64_t* d_total, void* stream);
/* Streaming writer: device batches are compressed on the GPU, copied to pinned host
 * buffers and written by a worker thread while the caller produces the next batch;
 * host pieces (the MCPL header) are deflated with zlib as members of their own. */
typedef struct kdsb200_gz_writer_s kdsb200_gz_writer;
kdsb200_gz_writer* kdsb200_gz_writer_open(const char* path, uint64_t max_batch_bytes,
                                          uint32_t rec_bytes, uint32_t tail_bytes, void* stream);
int kdsb200_gz_writer_put_host(kdsb200_gz_writer* w, const void* h_data, uint64_t nbytes);
int kdsb200_gz_writer_put_device(kdsb200_gz_writer* w, const void* d_data, uint64_t nbytes);
int kdsb200_gz_writer_close(kdsb200_gz_writer* w, uint64_t* bytes_written);

#ifdef __cplusplus
}
#endif
#endif /* KDSB200_H */
