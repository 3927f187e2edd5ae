/*
 * g2p.h - C ABI of libg2pgpu.so: the B200 (sm_100a) implementation of pyg2p's
 * data-parallel interpolation path (intertable build + intertable apply).
 *
 * The reference (ec-jrc/pyg2p 3.2.8) is pure Python and has no FFI on this path; each
 * entry point below names the reference code it replaces (paths relative to
 * src/pyg2p/main/).  The Python host in pyg2p_b200/ binds these with ctypes
 * (pyg2p_b200/_lib.py); INTEGRATION.md shows the stub a pyg2p maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; `stream` is a cudaStream_t passed as void*
 *     (NULL = legacy default stream);
 *   - every pointer is a DEVICE pointer on the current CUDA device unless the parameter
 *     is documented as host; buffers are caller-owned (allocate them with torch,
 *     cudaMalloc, ...); only `g2p_index` is library-owned;
 *   - calls are asynchronous on `stream` unless they return a scalar to the host;
 *   - return value: 0 = G2P_OK, negative = error, text via g2p_last_error()
 *     (thread-local).  There is no CPU fallback: without a CUDA device every compute
 *     entry point returns G2P_ERR_CUDA / G2P_ERR_NO_DEVICE.
 *
 * Device table layout ("SoA table"): indexes int32 [k][m], coeffs float64 [k][m]
 * (neighbour-major so that a warp reads consecutive targets).  The on-disk intertable
 * layout of the reference (array of structs {int64 indexes, float64|float32 coeffs},
 * shape (m,) or (m,k); interpolation/__init__.py:251) is produced/consumed by
 * g2p_table_pack_rec / g2p_table_unpack_rec.
 */
#ifndef G2P_H
#define G2P_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define G2P_VERSION 100

enum {
  G2P_OK = 0,
  G2P_ERR_INVALID = -1,     /* bad argument */
  G2P_ERR_CUDA = -2,        /* CUDA runtime error (message has the CUDA string) */
  G2P_ERR_NOMEM = -3,
  G2P_ERR_NO_DEVICE = -4,
  G2P_ERR_UNSUPPORTED = -5
};

enum { G2P_F64 = 0, G2P_F32 = 1 };                 /* dtype of lat/lon maps, distances, coeffs */
enum { G2P_MODE_RAW = 0, G2P_MODE_NEAREST = 1, G2P_MODE_INVDIST = 2 };
enum {
  G2P_MASK_AS_REFERENCE = 0,
  G2P_MASK_PROPAGATE = 1,
  G2P_MASK_FILL = 2,
  G2P_MASK_PROPAGATE_BITS = 3, /* PROPAGATE with bit-packed masks, in and out (see g2p_apply) */
  G2P_MASK_MARKED = 4          /* masked source values carry G2P_MASK_MARKER; out_mask (nullable) is bit-packed */
};

/* Summation order of the k = 4 weighted sum, OR-ed into the `mask_policy` argument of the apply entry points.
 * Default: ((w0*v0 + w1*v1) + w2*v2) + w3*v3 - what numpy's einsum does on the strided float64 field views of a table
 * loaded from file (reference interpolation/__init__.py:156).  G2P_SUM_PAIRWISE:
 * (w0*v0 + w2*v2) + (w1*v1 + w3*v3) - what the same einsum does when it buffers its operands: tables with float32
 * coefficients (PCRaster REAL4 target maps) and freshly built contiguous tables.  For other k (g2p_apply only) the flag
 * selects the same routine's order: two lanes taking the even / odd products, eight products per trip in the order
 * 6,7 | 4,5 | 2,3 | 0,1, then pairs, lane 0 + lane 1 (numpy `sum_of_products_contig_two`, SSE2).  Separate multiplies
 * and adds throughout. */
#define G2P_SUM_PAIRWISE 0x100

/* Bit-packed masks (G2P_MASK_PROPAGATE_BITS / G2P_MASK_MARKED): bit j of a row is bit (j & 7) of byte (j >> 3),
 * i.e. numpy.packbits(mask, bitorder='little') - one bit per point, as the GRIB bitmap section itself stores
 * missing points.  A source row of `src_stride` values has src_stride / 8 mask bytes (src_stride % 8 == 0); an
 * output row of `out_stride` values has G2P_BITS_ROW_BYTES(out_stride) mask bytes (whole 32-bit words). */
#define G2P_BITS_ROW_BYTES(stride) ((((stride) + 31) / 32) * 4)

/* The float64 quiet NaN that stands for "masked" inside a source field under G2P_MASK_MARKED (and inside the
 * shared-memory windows of the apply kernels under every mask policy).  Producers that touch every value anyway -
 * g2p_manip, g2p_mark_masked - write it where the mask is set, so that the apply kernel needs no mask input.
 * Consumers test the upper 32 bits only: any NaN whose upper word is 0x7ff86d61 is masked. */
#define G2P_MASK_MARKER 0x7ff86d61736b0001ull

/* per-target flags written by g2p_knn */
enum {
  G2P_FLAG_TIE = 1,        /* two of the best k+1 squared distances are exactly equal */
  G2P_FLAG_NEAR_TIE = 2,   /* ... or differ by less than 2^-44 relative */
  G2P_FLAG_SHORT = 4       /* fewer than k source points exist: idx = n, dist = +inf (as cKDTree) */
};

/* `to_regular` rotation of scipy_interpolation_lib.py:673-678 (unit-sphere output); NULL = default branch */
typedef struct g2p_rotation {
  double lat_south_pole_deg;
  double lon_south_pole_deg;
} g2p_rotation;

typedef struct g2p_index g2p_index; /* source spatial index on one device (replaces the cKDTree, :559) */

typedef struct g2p_index_info {
  int64_t n;          /* source points */
  int32_t cells_per_face_side; /* G: the cube-sphere grid is 6 x G x G cells */
  int64_t n_cells;
  double radius;      /* sphere radius of the source cloud */
  double mean_spacing;/* estimated mean point spacing (same unit as radius) */
  int64_t bytes;      /* device memory held by the index */
} g2p_index_info;

int g2p_version(void);
const char* g2p_last_error(void);
/* number of CUDA devices visible, or a negative error */
int g2p_device_count(void);
/* make `device` current for the calling thread (the library links the CUDA runtime statically;
 * the host calls this with torch's current device before using the other entry points) */
int g2p_set_device(int device);
/* kernels launched by this library in the calling process since load (bench.py's gpu_launches) */
int64_t g2p_launch_count(void);

/* K1  lat/lon -> xyz.  Replaces ScipyInterpolation.to_3d (scipy_interpolation_lib.py:665-696).
 * lon, lat: n values of `dtype`; xyz_out: (n,3) float64 row-major (float32-rounded values
 * when dtype == G2P_F32, as the reference casts x,y,z to float32, :691-694). */
int g2p_latlon_to_xyz(const void* lon, const void* lat, int64_t n, int dtype, double radius,
                      const g2p_rotation* rot /* host, nullable */, double* xyz_out, void* stream);

/* K1+K2  build the source index.  Replaces `KDTree(source_locations, leafsize=30)`
 * (scipy_interpolation_lib.py:551-559).  lon/lat: float64[n].  Synchronises `stream`. */
int g2p_index_create(const double* lon, const double* lat, int64_t n, double radius,
                     const g2p_rotation* rot /* host, nullable */, g2p_index** out, void* stream);
/* same, from given xyz (n,3) float64; the points must lie on a sphere centred at the origin */
int g2p_index_create_xyz(const double* xyz, int64_t n, g2p_index** out, void* stream);
int g2p_index_destroy(g2p_index* index);
int g2p_index_get_info(const g2p_index* index, g2p_index_info* info /* host */);
/* copy the source xyz (n,3), original order, into a caller buffer */
int g2p_index_get_xyz(const g2p_index* index, double* xyz_out, void* stream);

/* K3+K4  self query k=2 and max-reduce.  Replaces `tree.query(source_locations, k=2)` + `np.max`
 * (scipy_interpolation_lib.py:568-570).  dmax is a HOST pointer; synchronises `stream`.
 * min_upper_bound = dmax + dmax*4/Nj is computed by the host (Python float arithmetic). */
int g2p_index_max_nn_dist(g2p_index* index, double* dmax /* host */, void* stream);
/* the same over the source records [first, last) of the (cell-sorted) index only: with the sources split over
 * ranks, max over ranks of the partial results is the full result (one all-reduce of a double) */
int g2p_index_max_nn_dist_part(g2p_index* index, int64_t first, int64_t last, double* dmax /* host */, void* stream);

/* K1+K3(+K5)  k nearest source points of every target.
 * Replaces `tree.query(efas_locations, k=nnear)` (scipy_interpolation_lib.py:624-628) and,
 * with mode NEAREST / INVDIST, also `_build_nn` (:698-727) / `_build_weights_invdist`
 * (:774-812, :1018-1026) in the same kernel.
 *   tlon, tlat : m target coordinates of `dtype` (degrees), or NULL when `txyz` is given
 *   txyz       : (m,3) float64 target xyz, or NULL
 *   radius     : `r` of to_3d for the targets (grid_details['radius']); <= 0 = the index's
 *   k          : 1..16
 *   mode       : RAW      idx_out = neighbour ids, coef_out = distances (float32-rounded if dtype F32)
 *                NEAREST  idx_out = id or n_src where dist > mub; coef_out = distance (all rows)
 *                INVDIST  idx_out/coef_out = ids / normalised 1/d^2 weights, NaN + n_src beyond mub
 *   idx_out    : int32 [k][m];  coef_out: float64 [k][m];  flags_out: uint8 [m] (nullable)
 *   xyz_out    : (m,3) float64 target xyz as used by the search (nullable)
 * Distances use the cKDTree arithmetic ((dx*dx + dy*dy) + dz*dz, sqrt, float64, no FMA); ties are
 * ordered by lower source index and flagged. */
int g2p_knn(const g2p_index* index, const void* tlon, const void* tlat, const double* txyz,
            int64_t m, int dtype, double radius, const g2p_rotation* rot /* host, nullable */, int k, int mode,
            double mub, int32_t* idx_out, double* coef_out, uint8_t* flags_out, double* xyz_out,
            void* stream);

/* g2p_knn writing a BLOCK of a larger table: the m targets of this call are columns [out_offset, out_offset + m) of
 * idx_out / coef_out rows of `out_stride` values (flags_out / xyz_out stay per call, m entries).  This is how a rank
 * of a multi-GPU build writes its target block straight into its copy of the full table. */
int g2p_knn_block(const g2p_index* index, const void* tlon, const void* tlat, const double* txyz,
                  int64_t m, int dtype, double radius, const g2p_rotation* rot /* host, nullable */, int k, int mode,
                  double mub, int32_t* idx_out, double* coef_out, int64_t out_stride, int64_t out_offset,
                  uint8_t* flags_out, double* xyz_out, void* stream);

/* Multi-GPU table assembly over NVLink peer memory (replaces the NCCL all-gather of table shards): the
 * byte range [col_off, col_off + seg_bytes) of each of n_rows rows (pitch row_bytes) of the local buffer is stored
 * at the same place in n_peers peer buffers.  peer_bases: HOST array of device pointers of the peers' buffers as mapped
 * into this process (CUDA IPC / torch symmetric memory); all sizes multiples of 4 bytes (16 for full-width stores).
 * Asynchronous on `stream`; the caller's inter-rank barrier after it makes the table complete everywhere. */
int g2p_peer_push(const void* local_base, void* const* peer_bases /* host */, int n_peers, int64_t row_bytes,
                  int64_t col_off, int64_t seg_bytes, int n_rows, void* stream);

/* The same through the NVSwitch: `multicast_base` is the multicast (NVLS) address of the symmetric buffer; every 16-byte
 * store is replicated by the switch into all GPUs of the group (this one included), so a rank sends its block once
 * instead of once per peer (`multimem.st`).  16-byte aligned bases, rows and segments. */
int g2p_multicast_push(const void* local_base, void* multicast_base, int64_t row_bytes, int64_t col_off,
                       int64_t seg_bytes, int n_rows, void* stream);

/* K5  weights from raw (idx, dist).  Same arithmetic as the fused epilogue of g2p_knn; kept as
 * its own entry point so the stage can be checked in isolation.  In place on idx / coef. */
int g2p_weights(int mode, int k, int64_t m, int64_t n_src, double mub, int dtype,
                int32_t* idx_inout, double* coef_inout, void* stream);

/* K5b  Angular-distance weighting, mode `adw` (Shepard 1968) with k = 11.  Replaces the Shepard branch of
 * `_build_weights_invdist` (scipy_interpolation_lib.py:817-838, :953-1001, :1019-1021, :1036-1052) and its numba kernel
 * `adw_compute_weights_from_cutoff_distances` (:322-381).  Input is the RAW result of g2p_knn (k = 11).
 *
 * g2p_adw_distance_sum: the reference radius is the mean of the 7th distances over ALL targets (`np.mean(distances[:, 6])`,
 * :339, or the k=7 query of a split build, :576-585).  Returns the sum of dist[column][0..m) as an unevaluated
 * double-double {hi, lo} in sum_out (HOST, 2 doubles), accumulated exactly enough that hi does not depend on the launch
 * shape; with the targets split over ranks the partial sums are added before dividing by the global m.  Synchronises.
 * (The reference's own value comes from a numba parallel reduction and moves by a few ulp with the thread count.)
 *
 * g2p_weights_adw: in place on coef (distances [k][m] in -> normalised weights out; float32-representable values when
 * dtype == G2P_F32, as the reference computes them in float32 arrays for float32 target maps).  tlat / tlon are the m
 * ORIGINAL target coordinates (`dtype`), slat / slon the float64 source coordinates: the angular term is evaluated on the
 * lat/lon plane (:955-966).  No min_upper_bound, no sentinel rows (:561-562); d(1) <= 1e-10 -> [1, 0, ..., 0]. */
int g2p_adw_distance_sum(const double* dist, int64_t m, int k, int column, int dtype, double* sum_out /* host[2] */,
                         void* stream);
int g2p_weights_adw(int64_t m, int k, int dtype, double ref_radius, const void* tlat, const void* tlon,
                    const double* slat, const double* slon, int64_t n_src, const int32_t* idx, double* coef_inout,
                    void* stream);

/* K6  SoA table <-> the reference's record layout.  rec: m*k records of
 * {int64 indexes; float64 coeffs} (16 B) or, coeff_dtype F32, {int64; float32} (12 B, packed). */
int g2p_table_pack_rec(const int32_t* idx, const double* coef, int64_t m, int k, int coeff_dtype,
                       void* rec_out, void* stream);
int g2p_table_unpack_rec(const void* rec, int64_t m, int k, int coeff_dtype, int32_t* idx_out,
                         double* coef_out, void* stream);

/* The source points a table references ("touched"), sorted, and the table renumbered onto them: only those
 * values of a message are needed downstream (4 % of an O1280 field for the EFAS grid), so g2p_manip gathers
 * exactly them - also straight out of pinned host memory - and the apply kernels read compact fields.
 *   touched_out: int32, capacity min(n, m*k); cidx_out: int32 [k][m] (sentinel = nc); nc_out: host.  Synchronises. */
int g2p_table_compact(const int32_t* idx, int64_t m, int k, int64_t n, int32_t* touched_out, int32_t* cidx_out,
                      int64_t* nc_out /* host */, void* stream);

/* Host-side staging helper of the ingest path - the one entry point that takes HOST pointers for data and does no
 * device work: out[c] = src[touched[c]] for the nc source points a table reads (g2p_table_compact's list, copied to
 * the host), straight out of the pageable array a GRIB decoder returned (readers/grib.py:191-201) into a small
 * page-locked buffer - 2.2 MB instead of staging the 52.8 MB field.  mask (nullable): one byte per source value,
 * nonzero = missing; out_mask (nullable): the mask bytes of the picked points; fold_marker != 0: masked points get
 * G2P_MASK_MARKER as their value (input of G2P_MASK_MARKED).  The caller runs it on a decoder-side thread (ctypes
 * releases the GIL) while the device works on the previous message.  threads: the pick is split over that many threads
 * (the caller + persistent helpers, at most 4; <= 0: the default, the caller alone). */
int g2p_host_pick(const double* src /* host */, const uint8_t* mask /* host */, const int32_t* touched /* host */,
                  int64_t nc, double* out /* host */, uint8_t* out_mask /* host */, int fold_marker, int threads);

/* K7  batched intertable apply.  Replaces Interpolator._interpolate_scipy_append_mv
 * (interpolation/__init__.py:142-162) for `b` messages in one launch.
 *   idx, coef  : SoA table, int32 [k][m], float64 [k][m] (coef may be NULL when k == 1)
 *   src        : float64, message i at src + i*src_stride, n values each
 *   src_mask   : uint8 (1 = missing), message i at src_mask + i*src_stride; nullable
 *   out        : float64, message i at out + i*out_stride, m values each
 *   out_mask   : uint8, same striding as out; required iff mask_policy == PROPAGATE (ignored otherwise)
 *   k == 1     : out = src[idx], or mv_out where idx == n (the appended slot, :147)
 *   k  > 1     : out = ((w0*v0 + w1*v1) + w2*v2) + ..., separate multiply and add
 *   mask_policy: AS_REFERENCE ignores src_mask (the reference drops it);
 *                PROPAGATE is the intended semantics of :148-161: out_mask = OR of the k source masks and
 *                the masked outputs hold mv_out (`ma.filled(result, mv_out)`, what the writers do with a
 *                masked result: util/numeric.py:21-27); FILL is PROPAGATE without the out_mask array;
 *                PROPAGATE_BITS is PROPAGATE with src_mask and out_mask bit-packed (layout above; 8x less mask
 *                traffic, and out_mask costs one memset plus a store per MASKED output only);
 *                MARKED takes no src_mask: a source value equal to G2P_MASK_MARKER is masked; out_mask is
 *                bit-packed and may be NULL (then it is FILL for marked fields) */
int g2p_apply(const int32_t* idx, const double* coef, int64_t m, int k, const double* src,
              const uint8_t* src_mask, int64_t n, int b, int64_t src_stride, int64_t out_stride,
              double mv_out, int mask_policy, double* out, uint8_t* out_mask, void* stream);

/* In place on b rows of n values: v = mask ? G2P_MASK_MARKER : v (src_mask uint8, or bit-packed when
 * mask_bits != 0; strides as in g2p_apply).  The producer-side half of G2P_MASK_MARKED for fields that reach the
 * device with a separate mask. */
int g2p_mark_masked(double* src, const uint8_t* src_mask, int mask_bits, int64_t n, int b, int64_t src_stride,
                    void* stream);

/* K7w  windowed apply: the same operation as g2p_apply with the source values staged through shared
 * memory by TMA bulk copies (see pyg2p_b200/csrc/g2p_apply_win.cu).  It needs a plan, built once per
 * table: the targets are cut into 2-D tiles (rows x cols is the target map shape as in
 * `result.reshape(lonefas.shape)`, interpolation/__init__.py:263; pass rows = cols = 0 for a flat
 * table) and every tile gets the list of contiguous source index ranges its neighbours fall in.
 * g2p_plan_create returns G2P_ERR_UNSUPPORTED when the table's indexes are too scattered for a
 * shared-memory window; callers then use g2p_apply (still on the GPU). */
typedef struct g2p_plan g2p_plan;

typedef struct g2p_plan_info {
  int64_t m, n;
  int32_t k, tile_rows, tile_cols, n_tiles;
  int32_t max_window;   /* largest per-tile window, in source values */
  int32_t max_ranges;   /* largest number of index ranges of one tile */
  double mean_window;   /* mean per-tile window, in source values */
  int64_t bytes;        /* device memory held by the plan */
  int32_t block_rows, block_cols; /* targets owned by one half-warp (1x16, 2x8 or 4x4), chosen per table */
  double gather_wavefronts;       /* mean shared-memory wavefronts per half-warp gather (1 = conflict-free) */
} g2p_plan_info;

/* idx: int32 [k][m] device (the SoA table indexes; values >= n are "no value").  Synchronises `stream`. */
int g2p_plan_create(const int32_t* idx, int64_t m, int k, int64_t n, int64_t rows, int64_t cols, g2p_plan** out,
                    void* stream);
int g2p_plan_destroy(g2p_plan* plan);
int g2p_plan_get_info(const g2p_plan* plan, g2p_plan_info* info /* host */);
/* arguments as g2p_apply; the indexes come from the plan.  Requirements (else G2P_ERR_UNSUPPORTED):
 * k in {1, 4}; src 16-byte aligned with src_stride a multiple of 2 and >= n rounded up to 16;
 * under PROPAGATE / FILL also src_mask 16-byte aligned and src_stride a multiple of 16; under PROPAGATE_BITS
 * src_stride a multiple of 8 and out_mask 4-byte aligned. */
int g2p_apply_planned(const g2p_plan* plan, const double* coef, const double* src, const uint8_t* src_mask, int b,
                      int64_t src_stride, int64_t out_stride, double mv_out, int mask_policy, double* out,
                      uint8_t* out_mask, void* stream);

/* K8  the stages the reference runs around the interpolation, for all output messages of a run.
 *
 * g2p_manip: unit conversion -> step aggregation -> cut-off of negatives on the source grid
 * (controller.py:106-132; manipulation/conversion.py:21,33-52; manipulation/aggregator.py:72-279), one
 * launch for all outputs.  Conversion is `where(x != mv_grib, op2(op1(x)), mv_grib)` with op in
 * {none, x*c, x+c, x/c} - this covers every function of configuration/global/parameters.json (`x*1000`,
 * `-x*1000` = x*(-1000), `x-273.15` = x+(-273.15), `x*(-0.0353)`, ...) and the gem formula `(z/9.81)*0.0065`.
 * Each output message is one g2p_manip_item:
 *   PICK     out = 0.0 + c(V[in0])                         (instantaneous; in0 < 0: zeros)
 *   ACCUM    out = ((0.0 + c(V[in0])) - c(V[in1])) * unit_time / divisor      (in1 < 0: zeros at step 0)
 *   AVERAGE  out = (0.0 + c(V[in0]) + c(V[in1]) + ...) / divisor              (one add per hourly iterator)
 *   WAVERAGE out = (0.0 + c(V[in0])*w0 + c(V[in1])*w1 + ...) / divisor        (`@halfweights` average, aggregator.py:
 *            186-207: w = w_edge for the first / last input when edge_mask bit 0 / 1 is set, else unit_time (the
 *            mid weight); a masked input value is added unweighted, as numpy.ma's multiply leaves it; n_in = 0 and
 *            divisor = 0 give the reference's 0/0 = NaN field when fewer than two steps fall in the window)
 *   COPY     out = c(V[in0]), mask = mask(V[in0])   (no aggregation: the plain compaction of a message;
 *            `src` / `src_mask` may then be PINNED HOST memory - the kernel reads the touched values
 *            over PCIe and nothing else of the field ever crosses the bus)
 * With src_mask and NO out_mask, a masked output holds G2P_MASK_MARKER instead (input for G2P_MASK_MARKED).
 * out_mask follows the reference literally: ACCUM keeps only the mask of V[in1] (the current step loses its
 * mask in `v_iter_ma += V[t]`, aggregator.py:142-148), PICK none (:264-276); AVERAGE / WAVERAGE by mask_all:
 * -1 none - what _average really returns, because `numeric.get_masks(v_ord.values())` (:232) passes the dict view
 * as ONE argument and gets no mask -, 0 the OR of the item's inputs, 1 the OR of ALL b_in messages (what :231
 * says it intends).  With `touched` (device, int32[nc], sorted source indexes) the
 * outputs are compact fields out[o][c] = manip(V[..][touched[c]]); NULL = whole grid (nc = n). */
#define G2P_MANIP_MAX_IN 32
enum { G2P_OP_NONE = 0, G2P_OP_MUL = 1, G2P_OP_ADD = 2, G2P_OP_DIV = 3 };
enum { G2P_AGG_PICK = 0, G2P_AGG_ACCUM = 1, G2P_AGG_AVERAGE = 2, G2P_AGG_COPY = 3, G2P_AGG_WAVERAGE = 4 };

typedef struct g2p_manip_item {
  int32_t kind;                     /* G2P_AGG_* */
  int32_t n_in;                     /* PICK / COPY 1, ACCUM 2, AVERAGE 1..G2P_MANIP_MAX_IN, WAVERAGE 0..G2P_MANIP_MAX_IN */
  int32_t in[G2P_MANIP_MAX_IN];     /* input message numbers, in the order they are added */
  int32_t mask_all;
  int32_t edge_mask;                /* WAVERAGE: bit 0 / 1 = the first / last input carries w_edge */
  double unit_time;                 /* ACCUM: unit time; WAVERAGE: the mid weight dt1 */
  double divisor;                   /* ACCUM: aggregation step; AVERAGE: count; WAVERAGE: sum of the weights */
  double w_edge;                    /* WAVERAGE: the half weight dt1/2 */
} g2p_manip_item;

typedef struct g2p_manip_desc {
  int32_t conv_op1;
  double conv_c1;
  int32_t conv_op2;
  double conv_c2;
  double mv_grib;                   /* GRIB missingValue guarded by the conversion */
  int32_t cut_off_negative;
  int32_t n_out;
  const g2p_manip_item* items;      /* host, n_out entries */
} g2p_manip_desc;

int g2p_manip(const g2p_manip_desc* desc /* host */, const double* src, const uint8_t* src_mask /* nullable */,
              int64_t n, int b_in, int64_t src_stride, const int32_t* touched /* nullable */, int64_t nc,
              int64_t out_stride, double* out, uint8_t* out_mask /* iff src_mask */, void* stream);

/* V[a] + (V[b] - V[a]) * frac: a missing step synthesised linearly in time (aggregator.py:105,128) */
int g2p_time_interp(const double* va, const double* vb, double frac, int64_t n, double* out, void* stream);

/* Corrector (manipulation/correction.py:46,59-82): `where((dem!=dem_mv) & (p!=mv) & (gem!=gem_mv),
 * p + gem - dem*coeff, mv)` with mv = the netCDF f8 fill value.  g2p_correction_prepare computes the
 * per-target terms once (dem_term = dem*coeff, valid = dem and gem present); g2p_apply_planned_corrected
 * applies the formula in the epilogue of the windowed apply kernel, before the store. */
typedef struct g2p_correction {
  const double* gem;        /* device, m */
  const double* dem_term;   /* device, m */
  const uint8_t* valid;     /* device, m */
  double mv;                /* 9.969209968386869e36 */
  int32_t mode;             /* 0: (p + gem) - dem_term   [`p+gem-dem*c`, dem_term = dem*c]
                               1: (p / gem) * dem_term   [`p/gem*(10**((c)*dem/d))`, dem_term = 10**((c*dem)/d), which the
                                  host evaluates once per target grid with the libm `pow` the reference's numexpr calls] */
} g2p_correction;

int g2p_correction_prepare(const void* dem, int dem_dtype, double dem_mv, const double* gem, double gem_mv, double coeff,
                           int64_t m, double* dem_term_out, uint8_t* valid_out, void* stream);
/* the same formula in place on [b][m] outputs (after g2p_apply, for tables without a plan); mask nullable */
int g2p_correct(double* values, const uint8_t* mask, int64_t m, int b, int64_t stride, const g2p_correction* corr,
                void* stream);
int g2p_apply_planned_corrected(const g2p_plan* plan, const double* coef, const double* src, const uint8_t* src_mask,
                                int b, int64_t src_stride, int64_t out_stride, double mv_out, int mask_policy,
                                const g2p_correction* corr /* host struct of device pointers */, double* out,
                                uint8_t* out_mask, void* stream);

/* What the writers do to an interpolated (and corrected) field before it reaches the file, per target
 * (writers/__init__.py:66 `out_v[out_v == mv_output] = nan`; writers/netcdf.py:149-151,170-183 and writers/pcr.py:
 * 32-34,44-50 `_mask_values` + `values[isnan(values)] = mv`):
 *   out = (clone_mask | value mask | isnan(v) | (use_match & v == mv_match)) ? mv_write : v
 * clone_mask = missing cells of the clone map (nullable), mv_match = the Interpolator's mv_output (netCDF path
 * only: use_match), mv_write = missing value of the file (netCDF: fill value * scale_factor + offset; PCRaster:
 * nodata of the clone map).  g2p_writer_post_apply runs it in place on [b][m] fields (mask nullable = the
 * PROPAGATE out_mask); g2p_apply_planned_post runs it - after the optional correction - in the epilogue of the
 * windowed apply kernel, so the fields leave the device ready for the writer and no mask array has to follow them. */
typedef struct g2p_writer_post {
  const uint8_t* clone_mask;  /* device, m; nullable */
  double mv_match;
  int32_t use_match;
  double mv_write;
  int32_t out_f32;            /* g2p_apply_planned_post only: `out` is float32 [b][out_stride] - what a PCRaster REAL4 band
                                 (writers/pcr.py:38: GDAL WriteArray into the Float32 clone) or a netCDF `f4` variable
                                 stores anyway; values are rounded to nearest like a C / numpy cast, half the D2H bytes */
} g2p_writer_post;

int g2p_writer_post_apply(double* values, const uint8_t* mask, int64_t m, int b, int64_t stride,
                          const g2p_writer_post* post /* host struct */, void* stream);
/* as g2p_apply_planned_corrected; corr and post are both optional (NULL) */
int g2p_apply_planned_post(const g2p_plan* plan, const double* coef, const double* src, const uint8_t* src_mask, int b,
                           int64_t src_stride, int64_t out_stride, double mv_out, int mask_policy,
                           const g2p_correction* corr, const g2p_writer_post* post, double* out, uint8_t* out_mask,
                           void* stream);

/* The whole pipeline in ONE launch: convert -> aggregate -> cut-off (in the staging step of the windowed
 * apply kernel, on exactly the source values the tile gathers) -> interpolate -> correct.  `desc` as for
 * g2p_manip, with at most two inputs per output item (PICK, COPY, ACCUM, AVERAGE of <= 2) and no mask_all items
 * under a mask policy - otherwise G2P_ERR_UNSUPPORTED and the caller uses g2p_manip + g2p_apply_planned.
 * src / src_mask hold the b_in INPUT messages (whole source fields, rows as for g2p_apply_planned); out has
 * desc->n_out rows; corr may be NULL. */
int g2p_apply_planned_fused(const g2p_plan* plan, const double* coef, const g2p_manip_desc* desc /* host */,
                            const double* src, const uint8_t* src_mask, int b_in, int64_t src_stride,
                            int64_t out_stride, double mv_out, int mask_policy,
                            const g2p_correction* corr /* nullable */, double* out, uint8_t* out_mask, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* G2P_H */
