/*
 * nimrud_b200.h -- C ABI of the B200-native multiscale neighborhood eigenfeature path.
 *
 * The reference (grayhem/nimrud) is pure Python and has no FFI of its own: its boundary for this
 * path is the Python call signature of nimrud/minimal/multiscale.py.  Every entry point below names
 * the reference interface it replaces (file:line relative to the reference tree).  INTEGRATION.md
 * shows the ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - all array arguments are DEVICE pointers unless the name ends in `_host`; row-major, C order.
 *   - point clouds are (n,3); `dtype` is NBR_F32 or NBR_F64.  float32 inputs are promoted to float64
 *     exactly, so both dtypes describe the same points when the values are float32-representable.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  calls are
 *     stream-ordered and return without synchronising unless stated.
 *   - every function returns NBR_OK (0) or an error code; nbr_last_error() gives the message of the
 *     last failure on the calling thread.
 *   - scratch memory comes from the library's own stream-ordered CUDA memory pool (see nbr_trim_memory);
 *     the caller owns every input and output buffer.
 *   - the device that holds the buffers (and owns `stream`) must be the CURRENT device of the calling
 *     thread (cudaSetDevice) for every call, including nbr_lattice_destroy / nbr_lattice_info: kernels,
 *     scratch and cached tables are created on the current device.  one process may drive several
 *     devices in turn (kernel attributes and table caches are kept per device).
 */
#ifndef NIMRUD_B200_H
#define NIMRUD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NBR_OK 0
#define NBR_ERR_INVALID 1        /* bad argument                                                  */
#define NBR_ERR_TOO_FEW_POINTS 2 /* reference: ValueError, utils/geometry.py:34-35                */
#define NBR_ERR_ADDRESS_BITS 3   /* reference: ValueError, utils/geometry.py:59-60 (> 64 bits)    */
#define NBR_ERR_CUDA 4
#define NBR_ERR_UNSUPPORTED 5
#define NBR_ERR_OUT_OF_BOUNDS 6  /* reference: ValueError, utils/geometry.py:96-97                */
#define NBR_ERR_CAPACITY 7       /* nbr_tile_step_gather: the gather buffers are too small; grow them and call again */

#define NBR_F32 0
#define NBR_F64 1

/* descriptor_mask bits for the feature kernels.  0 = the reference's 4 columns per scale
 * [population, centroid distance, l_max/sum, l_mid/sum] (minimal/features.py:21-57).
 * NBR_DESC_EXTENDED appends 22 columns per scale (extension, not in the reference):
 * linearity, planarity, sphericity, omnivariance, anisotropy, eigenentropy, change of curvature,
 * verticality, normal x, y, z (nz >= 0), sum of eigenvalues (covariance trace, ddof = 1), the
 * upper triangle of the covariance xx, xy, xz, yy, yz, zz (ddof = 1; the legacy C_MSO output,
 * nimrud/prototypes/mso.py:1735-1746), and x, y of the unit eigenvectors of the largest and of the
 * middle eigenvalue (sign: x > 0, else y > 0, else z > 0; the legacy OG_MSO keeps the first two
 * components of two eigenvectors, nimrud/prototypes/mso.py:1498-1539). */
#define NBR_DESC_REFERENCE 0
#define NBR_DESC_EXTENDED 1
#define NBR_COLS_REFERENCE 4
#define NBR_COLS_EXTENDED 26

const char *nbr_last_error(void);
int nbr_version(void);
/* scratch comes from a PRIVATE stream-ordered memory pool per device (the application's default pool is never
 * touched); it keeps up to NBR_POOL_KEEP_MB (default 5 % of the device memory, at least 2 GB) of freed scratch for
 * the next call.  nbr_trim_memory releases what is cached on the current device (synchronises the device). */
int nbr_trim_memory(void);

/* ------------------------------------------------------------------------------------------------
 * voxel grid (replaces VoxelFilter.__init__ / _calculate_shift, utils/geometry.py:23-62)
 * ---------------------------------------------------------------------------------------------- */
typedef struct nbr_grid {
    double min_corner[3];  /* points.min(0) - e/2            geometry.py:37 */
    double max_corner[3];  /* points.max(0) + e/2            geometry.py:38 */
    double edge;
    int32_t widths[3];     /* ceil(log2(span/e)) per axis    geometry.py:55 */
    int32_t shifts[3];     /* bit offset of each axis in the packed address; shifts[0] = 0 */
    int32_t ndim;          /* 2 or 3 (2: z is ignored, width 0) */
    int32_t reserved;
} nbr_grid;

/* min and max of an (n, ndim<=3) cloud -> 6 doubles on the DEVICE (lo[3], hi[3]); unused axes 0. */
int nbr_bbox(const void *xyz, int dtype, int64_t n, int ndim, double *lohi_dev, void *stream);

/* HOST helper: grid parameters from a bounding box.  NBR_ERR_ADDRESS_BITS if more than 64 bits. */
int nbr_grid_from_bbox(const double lo[3], const double hi[3], double edge, int ndim, nbr_grid *out);

/* ------------------------------------------------------------------------------------------------
 * spatial index primitives
 * ---------------------------------------------------------------------------------------------- */
/* VoxelFilter.coordinate_to_address, geometry.py:103-116.  *oob_dev (device int32, may be NULL) is
 * set non-zero if a point lies outside [min_corner, max_corner] (geometry.py:94-97). */
int nbr_voxel_addresses(const void *xyz, int dtype, int64_t n, const nbr_grid *grid,
                        int64_t *addresses, int32_t *oob_dev, void *stream);

/* ascending LSD radix sort of 64-bit keys on bits [begin_bit, end_bit); result in `keys`.
 * `tmp` holds n keys.  replaces the sort inside np.unique, geometry.py:150. */
int nbr_sort_u64(uint64_t *keys, uint64_t *tmp, int64_t n, int begin_bit, int end_bit, void *stream);

/* same, carrying a 32-bit payload (stable). */
int nbr_sort_pairs_u64_u32(uint64_t *keys, uint64_t *keys_tmp, uint32_t *vals, uint32_t *vals_tmp,
                           int64_t n, int begin_bit, int end_bit, void *stream);

/* dedup of a sorted key array (the other half of np.unique).  out may alias nothing; *n_out_dev is
 * a device int64. */
int nbr_unique_u64(const uint64_t *sorted, int64_t n, uint64_t *out, int64_t *n_out_dev, void *stream);

/* VoxelFilter.address_to_coordinate, geometry.py:120-138: (k*e + min_corner) + e*0.5 -> (n,ndim) f64 */
int nbr_voxel_centres(const int64_t *addresses, int64_t n, const nbr_grid *grid, double *xyz_out,
                      void *stream);

/* exclusive prefix sum (the cell-offset scan). */
int nbr_exclusive_scan_u32(const uint32_t *in, uint32_t *out, int64_t n, void *stream);
int nbr_exclusive_scan_i64(const int64_t *in, int64_t *out, int64_t n, void *stream);

/* ------------------------------------------------------------------------------------------------
 * lattice index: the voxel-filtered search cloud of one edge length, as occupancy bit bricks
 * (replaces VoxelFilter.unique_voxels + cKDTree(search_voxels), minimal/multiscale.py:75-87)
 * ---------------------------------------------------------------------------------------------- */
typedef struct nbr_lattice nbr_lattice;

#define NBR_LATTICE_INDEXED 1 /* also keep the sorted unique addresses (np.unique order), so that
                                 neighbor INDICES can be reported; needs the radix sort */

int nbr_lattice_create(nbr_lattice **out, const void *search_xyz, int dtype, int64_t n_search,
                       const nbr_grid *grid, int flags, void *stream);
void nbr_lattice_destroy(nbr_lattice *lattice);
/* synchronises the stream; n_voxels = number of unique voxels (== len(np.unique(addresses))). */
int nbr_lattice_info(const nbr_lattice *lattice, int64_t *n_voxels, int64_t *n_bricks);
/* sorted unique addresses and/or their centres (either pointer may be NULL); INDEXED lattices only. */
int nbr_lattice_export(const nbr_lattice *lattice, int64_t *addresses, double *centres, void *stream);

/* ------------------------------------------------------------------------------------------------
 * queries
 * ---------------------------------------------------------------------------------------------- */
/* fused radius query + covariance + eigensolve + feature emission for the scales that share this
 * lattice's edge (replaces minimal/multiscale.py:94-122 and minimal/features.py:14-57).
 * for radius j, row i, writes cols [col_offset + j*C, col_offset + (j+1)*C) of out, C = 4 or 26 by
 * descriptor_mask; out has `out_row_stride` elements per row and dtype out_dtype.
 * algorithm: 0 = automatic, 1 = exact per-candidate kernel, 2 = row-interval kernel. */
int nbr_radius_features(const nbr_lattice *lattice, const void *query_xyz, int dtype, int64_t n_query,
                        const double *radii_host, int32_t n_radii, void *out, int out_dtype,
                        int64_t out_row_stride, int32_t col_offset, int32_t descriptor_mask,
                        int32_t algorithm, void *stream);

/* neighbor index sets in CSR form (parity path; replaces query_ball_tree, minimal/multiscale.py:103).
 * indices are positions in the sorted unique address array, ascending within each query.
 * pass 1: indices == NULL, fills offsets[n_query+1] (device int64).  pass 2: fills indices. */
int nbr_radius_sets(const nbr_lattice *lattice, const void *query_xyz, int dtype, int64_t n_query,
                    double radius, int64_t *offsets, int32_t *indices, void *stream);

/* k nearest voxels, total order (squared distance as float64, index).  no reference counterpart
 * (extension).  idx_out (n_query,k) int32 padded with -1, d2_out (n_query,k) f64 padded
 * with +inf; either may be NULL.  if feats_out is non-NULL also writes the 4 (or 26) feature
 * columns for each k in ks_host (ascending, ks[n_k-1] == k). */
int nbr_knn(const nbr_lattice *lattice, const void *query_xyz, int dtype, int64_t n_query, int32_t k,
            int32_t *idx_out, double *d2_out, const int32_t *ks_host, int32_t n_k, void *feats_out,
            int out_dtype, int64_t out_row_stride, int32_t col_offset, int32_t descriptor_mask,
            void *stream);

/* k nearest RAW POINTS of the search cloud (no voxel filter; the legacy pipeline's sspedge = 0,
 * nimrud/prototypes/mso.py:277,303-308; no counterpart in nimrud/minimal).  total order (float64 squared distance,
 * index in search_xyz).  builds its own cell index per call (cell_edge: hint for the cell size, <= 0 = automatic).
 * outputs as nbr_knn; the feature columns use the float64 moments of the k nearest points themselves. */
int nbr_knn_points(const void *search_xyz, int s_dtype, int64_t n_search, const void *query_xyz, int q_dtype,
                   int64_t n_query, int32_t k, double cell_edge, int32_t *idx_out, double *d2_out,
                   const int32_t *ks_host, int32_t n_k, void *feats_out, int out_dtype, int64_t out_row_stride,
                   int32_t col_offset, int32_t descriptor_mask, void *stream);

/* ------------------------------------------------------------------------------------------------
 * whole path
 * ---------------------------------------------------------------------------------------------- */
/* process_single_core, minimal/multiscale.py:27-67: out is (n_query, C*n_scales), scale-major.
 * global_lohi_host: optional 6 doubles (lo, hi) of the search cloud's bounding box to anchor the
 * lattices (multi-GPU: the all-reduced box); NULL = computed from `search`.
 * n_voxels_host: optional int64[n_scales] receiving the unique-voxel count per scale (forces a
 * stream synchronise). */
int nbr_multiscale_features(const void *query_xyz, int q_dtype, int64_t n_query, const void *search_xyz,
                            int s_dtype, int64_t n_search, const double *edges_host,
                            const double *radii_host, int32_t n_scales, void *out, int out_dtype,
                            int32_t descriptor_mask, const double *global_lohi_host,
                            int64_t *n_voxels_host, void *stream);

/* same with HOST buffers: copies the clouds in, runs, copies the features out; synchronous.
 * this is the call the reference-facing Python shim makes for numpy arguments.  the rows cross PCIe as float32 and,
 * for out_dtype NBR_F64, are widened into out_host by host threads (NBR_HOST_WIRE=f64: float64 on the wire);
 * pageable buffers are staged through pinned rings, pinned ones are used in place; the queries run in batches whose
 * copies overlap the kernels.  NBR_HOST_THREADS bounds the host threads (default: all cores, at most 64). */
int nbr_multiscale_features_host(const void *query_host, int q_dtype, int64_t n_query,
                                 const void *search_host, int s_dtype, int64_t n_search,
                                 const double *edges_host, const double *radii_host, int32_t n_scales,
                                 void *out_host, int out_dtype, int32_t descriptor_mask,
                                 int64_t *n_voxels_host);

/* host buffers for clouds / result rows.  pinned = 1: page-locked (cudaHostAlloc), used in place by
 * nbr_multiscale_features_host; pinned = 0: 2 MB-aligned pageable memory with huge pages requested (the Python shim
 * recycles its result arrays through these: a recycled buffer costs no page faults). */
int nbr_host_alloc(size_t bytes, int pinned, void **out);
int nbr_host_free(void *ptr);

/* vector-field multiscale operator (extension; legacy precedent V_MSO, nimrud/prototypes/mso.py:12-257).
 * nbr_voxel_vector_means: vectors (n_search, n_components) float32 carried by the search points are averaged
 *   per voxel of an INDEXED lattice -> voxvec_out (n_voxels, n_components) float32, rows in np.unique order.
 * nbr_radius_vector_means: for every query the mean of the voxel vectors over the voxels within `radius`
 *   (inclusive float64 test, as the eigenfeature path) -> out[q * out_row_stride + col_offset + f]; 0 where the
 *   neighborhood is empty. */
int nbr_voxel_vector_means(const nbr_lattice *lattice, const void *search_xyz, int dtype, int64_t n_search,
                           const float *vectors, int32_t n_components, float *voxvec_out, void *stream);
int nbr_radius_vector_means(const nbr_lattice *lattice, const void *query_xyz, int dtype, int64_t n_query,
                            double radius, const float *voxvec, int32_t n_components, void *out, int out_dtype,
                            int64_t out_row_stride, int32_t col_offset, void *stream);

/* multi-GPU halo selection (no reference counterpart; the precedent is nested_regions,
 * nimrud/utils/geometry.py:203-253: inclusive box +- buffer radius).  boxes_host: ndst <= 8 boxes as
 * [lo x,y,z, hi x,y,z] float64, already grown by the halo width.
 * nbr_halo_count: counts_dev[d] (uint64, device) += number of points of xyz inside box d.
 * nbr_halo_fill:  the points inside box d are written to rows [offsets_host[d], offsets_host[d] + count_d)
 *                 of `out` (rows of 3, same dtype as xyz, any order); cursors_dev[ndst] must be zero on entry. */
int nbr_halo_count(const void *xyz, int dtype, int64_t n, const double *boxes_host, int32_t ndst,
                   uint64_t *counts_dev, void *stream);
int nbr_halo_fill(const void *xyz, int dtype, int64_t n, const double *boxes_host, int32_t ndst,
                  const int64_t *offsets_host, uint64_t *cursors_dev, void *out, void *stream);

/* multi-GPU tile path in three steps, so that ordering the tile can overlap the halo exchange:
 * nbr_brick_origin: corner of brick (0,0,0) of the lattice of edge `finest_edge` anchored on the global box whose
 *                   directory covers local_lohi_host (NULL: the global box itself);
 * nbr_order_cloud:  perm_out[n] / sorted_out (n,3): the cloud in the spatially coherent order of the feature kernels
 *                   (cells aligned with that origin); lohi_host = bounding box of the cloud;
 * nbr_multiscale_features_tile: search cloud = the ordered tile (sorted_xyz, perm) + halo_xyz, queries = the tile;
 *                   local_lohi_host must contain every point of both; out rows in the tile's ORIGINAL order. */
int nbr_brick_origin(const double *global_lohi_host, const double *local_lohi_host, double finest_edge,
                     double *origin_out);
int nbr_order_cloud(const void *xyz, int dtype, int64_t n, const double *lohi_host, const double *origin_host,
                    double finest_edge, uint32_t *perm_out, void *sorted_out, void *stream);
int nbr_multiscale_features_tile(const void *sorted_xyz, const uint32_t *perm, int dtype, int64_t n,
                                 const void *halo_xyz, int64_t n_halo, const double *local_lohi_host,
                                 const double *global_lohi_host, const double *edges_host,
                                 const double *radii_host, int32_t n_scales, void *out, int out_dtype,
                                 int32_t descriptor_mask, int64_t *n_voxels_host, void *stream);

/* ------------------------------------------------------------------------------------------------
 * multi-GPU halo exchange over peer-mapped memory (NVLink / NVSwitch): no NCCL call, no count hand-shake.
 * no reference counterpart; what it selects is nested_regions' search-space rule (inclusive box +- buffer
 * radius, nimrud/utils/geometry.py:203-253, pinned by utils/tests/geometry_tests.py:353-389).
 * every rank owns a MAILBOX (one cudaMalloc allocation: header + capacity_rows rows of 3 coordinates) that
 * the other ranks map: through CUDA IPC between processes (nbr_mailbox_ipc_handle -> 64 bytes, exchanged by
 * the host code -> nbr_mailbox_connect_ipc), or directly when several tiles live in one process
 * (nbr_mailbox_connect_local; this is how the single-GPU tests drive the path).  world <= 16.
 * one step, every call collective over the ranks and stream-ordered:
 *   nbr_tile_box_publish  bounding box of this rank's tile (n may be 0) -> every peer's box table
 *   nbr_tile_boxes_wait   all boxes of the step -> boxes_host[world][8] = lo[3], hi[3], n_points, 0.
 *                         SYNCHRONISES the stream (the only host synchronisation of the exchange)
 *   nbr_halo_push         one pass over the tile: the points inside peer d's box grown by h (inclusive) are
 *                         stored into d's mailbox (remote atomic cursor + remote stores), then d is signalled
 *   nbr_halo_wait         stream-ordered wait for every peer's signal; afterwards nbr_mailbox_rows() holds
 *                         *nbr_mailbox_count_dev() rows (device-side count).  nbr_multiscale_features_tile_mb
 *                         does this wait itself, after it has marked the tile's own bricks.
 * rows that do not fit the destination's capacity are dropped and counted: the next nbr_tile_boxes_wait fails
 * with NBR_ERR_UNSUPPORTED, nbr_mailbox_status reports {timeout, rows dropped, rows pushed last step}. */
typedef struct nbr_mailbox nbr_mailbox;
int nbr_mailbox_create(nbr_mailbox **out, int32_t rank, int32_t world, int dtype, int64_t capacity_rows);
void nbr_mailbox_destroy(nbr_mailbox *mailbox);
/* closes this rank's mappings of the peers' allocations (staging_only != 0: only the gather staging buffers).  memory that
 * another process maps must not be freed: before teardown or a collective re-allocation every rank disconnects, the ranks
 * meet at a barrier, then the owners free (nimrud_b200/distributed.py: release_mailboxes, HaloMailbox.ensure_gather). */
int nbr_mailbox_disconnect(nbr_mailbox *mailbox, int32_t staging_only);
int nbr_mailbox_ipc_handle(const nbr_mailbox *mailbox, void *handle_out_64);
int nbr_mailbox_connect_ipc(nbr_mailbox *mailbox, int32_t peer, const void *handle_64);
int nbr_mailbox_connect_local(nbr_mailbox *mailbox, int32_t peer, const nbr_mailbox *peer_mailbox);
int nbr_mailbox_set_peer_capacity(nbr_mailbox *mailbox, int32_t peer, int64_t capacity_rows);
int nbr_tile_box_publish(nbr_mailbox *mailbox, const void *xyz, int dtype, int64_t n, void *stream);
int nbr_tile_boxes_wait(nbr_mailbox *mailbox, double *boxes_host, void *stream);
int nbr_halo_push(nbr_mailbox *mailbox, const void *xyz, int dtype, int64_t n, const double *boxes_host, double h,
                  void *stream);
int nbr_halo_wait(nbr_mailbox *mailbox, void *stream);
const void *nbr_mailbox_rows(const nbr_mailbox *mailbox);
const uint64_t *nbr_mailbox_count_dev(const nbr_mailbox *mailbox);
int nbr_mailbox_status(const nbr_mailbox *mailbox, uint64_t *status3_host);
/* tests / debugging: copies min(count, max_rows) rows of the last completed step to dst_dev, *n_rows_host = count.
 * synchronises the stream. */
int nbr_mailbox_read(const nbr_mailbox *mailbox, void *dst_dev, int64_t max_rows, int64_t *n_rows_host, void *stream);
/* nbr_multiscale_features_tile with the halo taken from this rank's mailbox */
int nbr_multiscale_features_tile_mb(const void *sorted_xyz, const uint32_t *perm, int dtype, int64_t n,
                                    nbr_mailbox *mailbox, const double *local_lohi_host,
                                    const double *global_lohi_host, const double *edges_host,
                                    const double *radii_host, int32_t n_scales, void *out, int out_dtype,
                                    int32_t descriptor_mask, int64_t *n_voxels_host, void *stream);

/* one whole step of a rank in one call: publish + wait (host synchronisation) -> push -> query order of the tile ->
 * lattices from tile + mailbox -> features of the tile's points, rows in the tile's own order.
 * boxes_host_out: optional [world][8] (lo, hi, n_points, 0 of every tile). */
int nbr_tile_step(nbr_mailbox *mailbox, const void *xyz, int dtype, int64_t n, const double *edges_host,
                  const double *radii_host, int32_t n_scales, void *out, int out_dtype, int32_t descriptor_mask,
                  double *boxes_host_out, int64_t *n_voxels_host, void *stream);

/* FEATURE ALL-GATHER WITHOUT A COLLECTIVE CALL (north_star: "final feature all-gather"; no reference counterpart, the
 * reference is one process).  every rank owns a STAGING BUFFER (cudaMalloc'ed by nbr_mailbox_gather_alloc, mapped by the
 * peers like the mailbox: nbr_mailbox_gather_ipc_handle -> nbr_mailbox_gather_connect_ipc, or
 * nbr_mailbox_gather_connect_local inside one process) of at least nbr_gather_staging_bytes(rows of all ranks, row
 * bytes).  (re)allocation is collective on the caller's side: every rank allocates the same size, the handles are
 * exchanged, everybody connects, and only then the next step runs.
 * nbr_tile_step_gather is nbr_tile_step with the all-gather inside: the fused feature kernel writes each finished row
 * to its place in out_all AND -- in processing order, a warp's 32 rows as one contiguous piece, with the row numbers --
 * into the staging buffer of every other rank (remote stores over NVLink from the kernel's own write-out; scale sets
 * the 7x7x7 kernel does not cover alone are pushed after their kernels); a stream-ordered signal + wait follows, then
 * the staged rows of the peers are put in place.  when the call's work on `stream` is done, out_all (ordinary device
 * memory, out_rows_capacity rows) holds the rows of all ranks in rank order, every share in its tile's own order.
 * row_offsets_host[world + 1]: first row of every rank's share.  NBR_ERR_CAPACITY (after the step's halo exchange has
 * completed, on every rank alike if the ranks pass results of one size): staging buffers or result too small for
 * sum n rows; row_offsets_host is valid, grow them collectively and call again. */
int nbr_mailbox_gather_alloc(nbr_mailbox *mailbox, uint64_t bytes);
int nbr_mailbox_gather_ipc_handle(const nbr_mailbox *mailbox, void *handle_out_64);
int nbr_mailbox_gather_connect_ipc(nbr_mailbox *mailbox, int32_t peer, const void *handle_64, uint64_t bytes);
int nbr_mailbox_gather_connect_local(nbr_mailbox *mailbox, int32_t peer, const nbr_mailbox *peer_mailbox);
void *nbr_mailbox_gather_ptr(const nbr_mailbox *mailbox, uint64_t *bytes_out);
uint64_t nbr_gather_staging_bytes(int64_t total_rows, int64_t row_bytes);
/* step-wise form: nbr_multiscale_features_tile_mb with the staging writes (this rank's rows are rows [row_offset,
 * row_offset + n) of total_rows), then the signal + wait, then the staged rows into out_all */
int nbr_multiscale_features_tile_mb_gather(const void *sorted_xyz, const uint32_t *perm, int dtype, int64_t n,
                                           nbr_mailbox *mailbox, const double *local_lohi_host,
                                           const double *global_lohi_host, const double *edges_host,
                                           const double *radii_host, int32_t n_scales, void *out_all, int out_dtype,
                                           int32_t descriptor_mask, int64_t row_offset, int64_t total_rows,
                                           int64_t *n_voxels_host, void *stream);
int nbr_gather_finish(nbr_mailbox *mailbox, void *stream);
int nbr_gather_unpermute(nbr_mailbox *mailbox, const int64_t *row_offsets_host, int64_t row_bytes, void *out_all, void *stream);
int nbr_tile_step_gather(nbr_mailbox *mailbox, const void *xyz, int dtype, int64_t n, const double *edges_host,
                         const double *radii_host, int32_t n_scales, void *out_all, int64_t out_rows_capacity,
                         int out_dtype, int32_t descriptor_mask, double *boxes_host_out, int64_t *n_voxels_host,
                         int64_t *row_offsets_host, void *stream);

/* nbr_tile_step with HOST buffers (the tile goes up, the rows come down in batches whose copies overlap the kernels;
 * float32 on the wire, widened by host threads for out_dtype NBR_F64, like nbr_multiscale_features_host). */
int nbr_tile_step_host(nbr_mailbox *mailbox, const void *xyz_host, int dtype, int64_t n, const double *edges_host,
                       const double *radii_host, int32_t n_scales, void *out_host, int out_dtype,
                       int32_t descriptor_mask, double *boxes_host_out);

/* debug builds only (-DNBR_BOUNDS_CHECK=1, scripts/bounds_check.sh): number of out-of-range shared-memory indices the
 * fused kernel has seen on the current device; always 0 in a normal build. */
int64_t nbr_debug_bounds_violations(void);

/* counters for tests and benches: number of kernels this library has launched in this process. */
int64_t nbr_kernel_launches(void);

/* optional device timing of the whole-path drivers, CUDA events on the launching stream.
 * nbr_timing_read synchronises on the recorded events, writes the accumulated milliseconds of the
 * NBR_TIMING_PHASES phases [bounding box, index build, query ordering, feature kernels, tile box exchange, halo
 * push, wait for the peers' halo (also inside the index build's span), reserved] since the
 * previous read, and clears them. */
#define NBR_TIMING_PHASES 8
void nbr_timing_enable(int on);
int nbr_timing_read(double *ms_out);

#ifdef __cplusplus
}
#endif
#endif /* NIMRUD_B200_H */
