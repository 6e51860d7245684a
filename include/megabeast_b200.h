/*
 * megabeast_b200 -- C ABI of the B200-native ensemble likelihood (libmegabeast_b200.so).
 *
 * This is the drop-in boundary for ONE path of megaBEAST: the per-step ensemble likelihood
 *
 *     /root/reference/megabeast/singlepop_dust_model.py:110-229   MB_Model.lnlike
 *     /root/reference/megabeast/helpers.py:105-164                get_predicted_num_stars
 *
 * The reference is pure Python and has no FFI; what a maintainer would bind (ctypes stub in
 * INTEGRATION.md) is exactly the entry points below.  Plain pointers and sizes only -- no
 * torch, numpy or CUDA types in any signature (a `void* stream` is an opaque cudaStream_t).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; mb_last_error() then holds a
 *     thread-local, human-readable message (CUDA failures, invalid arguments).
 *   - NUMERICAL conditions the reference turns into exceptions are NOT errors here: they come
 *     back per walker in `status` (MB_STATUS_*), and the host language raises.
 *   - host pointers passed to mb_create are only read during the call; nothing is retained.
 *   - one handle = one CUDA device = one shard of stars.  Calls on a handle must be
 *     serialised by the caller (the reference is single-threaded).
 *   - there is no CPU fallback: without a CUDA device mb_create fails.
 */
#ifndef MEGABEAST_B200_H
#define MEGABEAST_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MB_ABI_VERSION 2

/* Ensemble parameters that can carry a free prior model, in phi order.  This is
 * MB_Model.params (singlepop_dust_model.py:29) without M_ini: a free IMF is unreachable in
 * the reference (KeyError 'nsubvars' at :137 and :257), so it stays a host-side constant. */
#define MB_NPARAM 4
enum mb_param { MB_P_LOGA = 0, MB_P_AV = 1, MB_P_RV = 2, MB_P_FA = 3 };

/* Prior-model names the device evaluates analytically (beast.physicsmodel.priormodel names,
 * docs/megabeast/settings.rst:39,78).  MB_FIXED: prior name "fixed" -> the parameter
 * contributes no factor at all (singlepop_dust_model.py:135).  MB_TABULATED: the host
 * evaluates any callable on the parameter's distinct grid values and passes the table. */
enum mb_prior_kind {
  MB_FIXED = 0,
  MB_FLAT = 1,          /* no variables            */
  MB_LOGNORMAL = 2,     /* mean, sigma             */
  MB_TWO_LOGNORMAL = 3, /* mean1, mean2, sigma1, sigma2, N1_to_N2 */
  MB_BINS_HISTO = 4,    /* nnodes heights (left-constant steps over nodes) */
  MB_BINS_INTERP = 5,   /* nnodes values (linear interpolation over nodes) */
  MB_EXPONENTIAL = 6,   /* tau [Gyr]; x is log10(age/yr) */
  MB_TABULATED = 7
};
#define MB_MAX_NODES 64

/* phi layout: for p in (logA, Av, Rv, fA), if kind[p] is free, its variables in the order
 * listed above; ndim = sum.  Same order as MB_Model.start_params (:82-108). */
typedef struct {
  int32_t kind[MB_NPARAM];
  int32_t nnodes[MB_NPARAM];
  double nodes[MB_NPARAM][MB_MAX_NODES];
} mb_model_desc;

/* beast_moddata (fit_single.py:37-57): 1-D float64 columns of length G.  col[p] may be NULL
 * when kind[p] == MB_FIXED.  logA and Z are needed only for the N-stars term. */
typedef struct {
  int64_t G;
  const double* col[MB_NPARAM];
  const double* logA;
  const double* Z;
  const double* prior_weight;
  const double* grid_weight;
  const double* completeness;
} mb_grid_desc;

/* star_lnpgriddata (helpers.py:15-43): indxs (K,S) int64 and vals (K,S) float64, C order
 * (star is the fast axis, singlepop_dust_model.py:157,196-198).  Non-finite vals mark padded
 * entries whose index is never dereferenced.  This handle owns stars [star_begin, star_end);
 * n_stars_total is the S of the Poisson term (:183-185), i.e. the full column count of the fit. */
typedef struct {
  int64_t K;
  int64_t S;
  const int64_t* indxs;
  const double* vals;
  int64_t star_begin;
  int64_t star_end;
  int64_t n_stars_total;  /* 0 -> S; set it when the arrays hold only this rank's columns */
} mb_stars_desc;

/* Output of precompute_mass_multipliers (helpers.py:46-102), a once-per-fit host constant.
 * NULL pointer to mb_create <=> no N-stars term (logA prior fixed, :71-74). */
typedef struct {
  int32_t n_ages;
  int32_t n_mets;
  const double* ages;     /* sorted unique logA */
  const double* mets;     /* sorted unique Z    */
  const double* massmult; /* [n_ages][n_mets]   */
  double avemass;
} mb_massmult_desc;

enum mb_mode {
  MB_MODE_AUTO = 0,        /* FACTOR when the factor tables fit shared memory, else TABLE_UNIQUE; ELEMENTWISE for
                              grids whose columns cannot be class-coded (> 4096 distinct values) */
  MB_MODE_FACTOR = 1,      /* per-(star,point) class codes + per-walker factor tables in smem: an OUTER table (age
                              classes, times the dust parameters that do not fit inside) read once per run of a
                              star's entries and an INNER table (the other dust parameters) read once per entry.
                              The split is chosen at mb_create: all dust parameters inside when that fits (the
                              classic age | dust split), else e.g. A(V) inside, age x R(V) outside */
  MB_MODE_TABLE_UNIQUE = 2,/* physmod materialised over the distinct parameter combinations */
  MB_MODE_TABLE_GRID = 3,  /* physmod[G][W] materialised over every grid row, gathered by index */
  MB_MODE_ELEMENTWISE = 4  /* physmod[G][W] by evaluating the prior models ELEMENTWISE over the grid's physics
                              columns (singlepop_dust_model.py:131-148 literally), N-stars sums over the grid
                              rows, gathered by index: no class coding, works for any grid.  AUTO falls back to
                              it when a free parameter's column has too many distinct values to class-code. */
};

typedef struct {
  int32_t device;          /* CUDA ordinal */
  int32_t mode;            /* enum mb_mode */
  int32_t keep_indices;    /* retain the gathered grid rows for mb_gathered_indices */
  int32_t merge_duplicates;/* pre-sum a star's entries that share all class codes (FACTOR/TABLE_UNIQUE) */
  int32_t segments_per_warp;/* scheduling granularity of K2: entry-stream segments per resident warp (0 -> 1) */
  int32_t tail_percent;    /* percent of the entries cut into small dynamically claimed segments (0 -> none) */
  int32_t table_precision; /* FACTOR mode: 0 = fp64 factor tables (default, the 1e-9 parity mode); 1 = 32-bit
                              tables in shared memory (the high word of each fp64 factor, rounded to 21
                              significant bits; all sums still fp64): half the shared-memory traffic, ~8 %
                              faster, measured error 7e-9 relative on lnlike, stated tolerance 1e-7 */
  int32_t split_inner_mask;/* FACTOR mode: 0 = choose the table split automatically; else the dust parameters kept in
                              the INNER table as a bit mask (1 = Av, 2 = Rv, 4 = fA), the others go to the outer
                              table with the age classes.  For tests and measurements. */
  int32_t reserved[8];
} mb_options;

typedef struct {
  int32_t abi_version;
  int32_t device;
  int32_t mode;            /* the mode actually in use */
  int32_t ndim;
  int32_t poisson;         /* N-stars term on */
  int32_t n_classes[MB_NPARAM];
  int32_t n_age_classes;   /* factor table A rows */
  int32_t n_dust_classes;  /* factor table B rows */
  int32_t n_bins;
  int32_t sm_count;
  int32_t walkers_per_pass;/* of the last evaluation */
  int32_t launches_last_eval;
  int64_t G;
  int64_t n_stars_local;
  int64_t n_stars_total;
  int64_t n_entries;       /* finite (star, point) entries held on the device */
  int64_t n_entries_raw;   /* before merging duplicates */
  int64_t device_bytes;
  int64_t code_bytes;      /* bytes of an entry's code word on the device: 4, or 2 (FACTOR mode, few enough classes) */
  int64_t n_outer_rows;    /* FACTOR: rows of the outer factor table (age classes x outer dust combinations) */
  int64_t n_inner_rows;    /* FACTOR: rows of the inner factor table */
  int64_t split_inner_mask;/* dust parameters in the inner table (1 = Av, 2 = Rv, 4 = fA) */
  int64_t n_runs;          /* runs of equal outer class in the entry stream (one outer-row read each) */
  int64_t reserved[3];
} mb_info;

/* Device result layout of mb_lnlike_device: four rows of W doubles.  The sum over stars of ln(star
 * prob) comes as TWO doubles, STARSUM (an exact multiple of 2^-24) and STARSUM_LO (the remainder, an
 * exact multiple of 2^-68): the kernel accumulates it as two 64-bit integers, and each word converts
 * to a double without rounding.  Adding the words of several shards in floating point (NCCL, or the
 * host for several handles) is therefore exact, and STARSUM + STARSUM_LO -- the only rounding -- gives
 * the same bits however the shards were added.  The three rows that add over star shards come first
 * so that one all-reduce of 3*W doubles covers them. */
enum { MB_OUT_STARSUM = 0, MB_OUT_STARSUM_LO = 1, MB_OUT_NONFINITE = 2, MB_OUT_POISSON = 3, MB_OUT_ROWS = 4 };

#define MB_STATUS_OK 0
#define MB_STATUS_NONFINITE_STAR 1 /* ValueError("invidual integrated star prob is not finite"), :216-217 */
#define MB_STATUS_TOTAL_ZERO 2     /* print + exit(), :225-227 */
#define MB_STATUS_PEER_TIMEOUT 4   /* a peer GPU never delivered its shard (mb_peer_connect path) */

typedef struct mb_handle mb_handle;

int mb_abi_version(void);
const char* mb_last_error(void);

/* Number of visible CUDA devices, or -1 (with mb_last_error) if the runtime cannot start. */
int mb_device_count(void);

/* replaces: the per-call arguments of MB_Model.lnlike (singlepop_dust_model.py:110) -- data is
 * uploaded once and kept resident, keyed by the caller on the identity of the two dicts. */
int mb_create(const mb_grid_desc* grid, const mb_stars_desc* stars, const mb_model_desc* model,
              const mb_massmult_desc* massmult, const mb_options* opts, mb_handle** out);
void mb_destroy(mb_handle* h);
int mb_get_info(const mb_handle* h, mb_info* out);

/* replaces: MB_Model.lnlike(phi, ...) for a batch of W walkers (singlepop_dust_model.py:110-229).
 * phi: host [W][ndim]; lnlike: host [W]; status: host [W].  Synchronous.  Covers the whole
 * likelihood for this handle's stars including the Poisson term. */
int mb_lnlike(mb_handle* h, const double* phi, int32_t W, int32_t ndim, double* lnlike,
              int32_t* status);

/* replaces: lnprob(phi, ...) (singlepop_dust_model.py:277-304) for a batch when every prior is a
 * flat box (the only kind MB_Model.lnprior accepts, :255-273): lnprob[w] = -inf for a walker with
 * any phi outside [lo, hi] (its likelihood is not evaluated, :302-303), else 0 + lnlike.  lo/hi:
 * [ndim] in phi order.  last_inside (optional): index of the last walker that was evaluated, -1 if
 * none -- the walker whose phi the reference leaves in the model dictionaries. */
int mb_lnprob_box(mb_handle* h, const double* phi, int32_t W, int32_t ndim, const double* lo,
                  const double* hi, double* lnprob, int32_t* status, int32_t* last_inside);

/* Same with host-evaluated tables for MB_TABULATED parameters: tab[p] is [W][n_classes[p]]
 * (values of the parameter's callable at mb_class_values), NULL for analytic parameters. */
int mb_lnlike_tabulated(mb_handle* h, const double* phi, int32_t W, int32_t ndim,
                        const double* const tab[MB_NPARAM], double* lnlike, int32_t* status);

/* Split form for driving several handles (GPUs) from one host thread: begin() enqueues the
 * host->device copy, the kernels and the device->host copy on the handle's stream and returns;
 * end() waits and hands back the raw [MB_OUT_ROWS][W] result rows.  mb_combine() turns rows
 * (summed over shards by the caller: rows STARSUM, STARSUM_LO and NONFINITE add, POISSON is shared) into
 * lnlike/status exactly as mb_lnlike does: lnlike = POISSON + (STARSUM + STARSUM_LO). */
int mb_lnlike_begin(mb_handle* h, const double* phi, int32_t W, int32_t ndim,
                    const double* const tab[MB_NPARAM]);
int mb_lnlike_end(mb_handle* h, double* rows);
void mb_combine(const double* rows, int32_t W, double* lnlike, int32_t* status);

/* Multi-GPU building block: device pointers, asynchronous on `stream`.  d_out is
 * [MB_OUT_ROWS][W] doubles: partial sum of ln(star prob) over this handle's stars (two words), the
 * count of non-finite stars, and the Poisson term (identical on every shard).  The caller all-reduces
 * the first 3*W doubles (rows STARSUM, STARSUM_LO, NONFINITE: one NCCL allreduce) and adds POISSON once. */
int mb_lnlike_device(mb_handle* h, const double* d_phi, int32_t W, int32_t ndim, double* d_out,
                     void* stream);

/* Star shards on several GPUs, one process per GPU: fuse the reduction over the shards into the
 * likelihood kernel itself, over NVLink peer memory (no separate collective launch).
 *   mb_peer_export   writes an opaque MB_PEER_HANDLE_BYTES token for this handle's exchange buffer
 *   mb_peer_connect  takes the tokens of all `world` ranks in rank order (exchanged by the caller,
 *                    e.g. torch.distributed.all_gather) and maps the peers' buffers (CUDA IPC)
 * Afterwards every mb_lnlike* call on the handle returns rows STARSUM and NONFINITE already summed
 * over all ranks, identical on every rank, provided all ranks issue the same sequence of calls.
 * A rank that waits longer than the timeout for a peer (20 s; environment MB_PEER_TIMEOUT_MS at
 * mb_create) reports MB_STATUS_PEER_TIMEOUT for every walker of that call. */
#define MB_MAX_PEERS 16
#define MB_PEER_HANDLE_BYTES 64
int mb_peer_export(mb_handle* h, void* token);
int mb_peer_connect(mb_handle* h, int32_t rank, int32_t world, const void* tokens);
int mb_peer_disconnect(mb_handle* h);
/* Device-side rendezvous of the connected ranks on `stream` (a one-warp kernel: every rank raises a flag
 * in all peers over NVLink and waits for theirs): all ranks' next kernels start within one NVLink
 * latency of each other.  A measurement aid (bench.py aligns the ranks ahead of every timed launch);
 * the likelihood never needs it.  No-op without peers.  All ranks must call it the same number of times. */
int mb_peer_rendezvous(mb_handle* h, void* stream);

/* Pixel batches (origin old/megabeast_fit.py:97-182, a serial loop of one
 * fit per sky pixel): n_pixels independent ensembles over ONE shared grid.  Pixel p owns the stars
 * (columns of indxs/vals) [star_offsets[p], star_offsets[p+1]); every pixel has its own W walkers
 * and its own Poisson term (S = its star count).  phi is [n_pixels][W][ndim], lnlike/status are
 * [n_pixels][W]; ALL pixels and walkers are evaluated by one launch.  The device form writes rows
 * [n_pixels][MB_OUT_ROWS][W].  1 <= W <= 128; FACTOR mode only.  To use several GPUs give each
 * rank a subset of the pixels: pixels are independent, no collective is needed. */
int mb_create_pixels(const mb_grid_desc* grid, const mb_stars_desc* stars, const mb_model_desc* model,
                     const mb_massmult_desc* massmult, const mb_options* opts, int64_t n_pixels,
                     const int64_t* star_offsets, mb_handle** out);
int mb_lnlike_pixels(mb_handle* h, const double* phi, int32_t W, int32_t ndim, double* lnlike,
                     int32_t* status);
/* Same with host-evaluated tables for MB_TABULATED parameters: tab[p] is [n_pixels * W][n_classes[p]]
 * (pixel-major, then walker), NULL for analytic parameters. */
int mb_lnlike_pixels_tabulated(mb_handle* h, const double* phi, int32_t W, int32_t ndim,
                               const double* const tab[MB_NPARAM], double* lnlike, int32_t* status);
int mb_lnlike_pixels_device(mb_handle* h, const double* d_phi, int32_t W, int32_t ndim, double* d_out,
                            void* stream);
/* replaces: lnprob(phi, ...) (singlepop_dust_model.py:277-304) for every walker of every pixel when every
 * prior is a flat box (mb_lnprob_box for pixel batches): lnprob[p][w] = -inf, status 0 for a walker with any
 * phi outside [lo, hi] -- it still occupies a column of the launch, evaluated at `park` ([ndim], inside the
 * box, e.g. the start parameters) and discarded, as the reference never evaluates it (:302-303) -- else
 * 0 + lnlike.  Analytic models only. */
int mb_lnprob_pixels_box(mb_handle* h, const double* phi, int32_t W, int32_t ndim, const double* lo,
                         const double* hi, const double* park, double* lnprob, int32_t* status);
int64_t mb_pixel_count(const mb_handle* h);

/* Guard against stale device data.  The data dictionaries arrive as arguments on every call of the
 * reference's lnlike (singlepop_dust_model.py:110) while the device copy is made once; a caller that
 * mutates an array IN PLACE would otherwise be evaluated on old data.  mb_watch_host registers a host
 * array (up to 16 per handle; ptr = NULL clears the list): the library fingerprints 16 words spread over
 * it now and again at the start of every mb_lnlike* / mb_lnprob_box call, and fails with a message
 * starting "MB_STALE" when they differ (the host side then rebuilds the handle).  This is the one
 * exception to "host pointers are not retained": the CALLER keeps watched arrays alive. */
int mb_watch_host(mb_handle* h, const void* ptr, int64_t nbytes);

/* Distinct values ("classes") of a free parameter's grid column, ascending; the x at which a
 * host callable must be tabulated.  Returns the count; fills at most `cap` values.
 * ELEMENTWISE mode (a column that cannot be class-coded) with an MB_TABULATED dust parameter: every
 * grid row is its own class -- the G column values in row order, tab[p] is [W][G] (the cost the
 * reference itself pays, singlepop_dust_model.py:146-148). */
int mb_class_values(const mb_handle* h, int32_t param, double* out, int32_t cap);

/* Parity hook: the grid rows this handle dereferences, star-major (needs keep_indices).
 * offsets has n_stars_local+1 entries.  Must equal indxs[:, s][isfinite(vals[:, s])]
 * (singlepop_dust_model.py:196-198) bit for bit. */
int mb_gathered_indices(const mb_handle* h, int64_t* rows, int64_t rows_cap, int64_t* offsets);

/* Parity hook: what the DEVICE holds.  Copies the class-coded entry stream back from device memory
 * and decodes it: entry e of local star s (offsets[s] <= e < offsets[s+1], star-major, in the device's
 * own order: sorted by class, duplicates merged unless merge_duplicates = 0) carries the class index
 * cls[e*MB_NPARAM + p] of parameter p (index into mb_class_values(p); 0 for a fixed parameter) and the
 * folded weight vals*prior_weight*grid_weight*completeness in weight[e].  Star boundaries come from
 * the new-star flags of the device's code words.  FACTOR and TABLE_UNIQUE modes; needs keep_indices.
 * cap = capacity of cls (in entries) and weight; offsets has n_stars_local+1 entries. */
int mb_device_entries(const mb_handle* h, int64_t cap, int32_t* cls, double* weight, int64_t* offsets);

/* Per-kernel device time (ms) of the last evaluation, measured with CUDA events when
 * profiling was enabled with mb_set_profiling(h, 1).  Order: K0 factor tables, K1 expand (TABLE
 * modes, else ~0), K2 stars (Poisson prologue + star loop + its tail, the K3 reduction). */
#define MB_NKERNELS 3
int mb_set_profiling(mb_handle* h, int32_t on);
int mb_last_kernel_ms(const mb_handle* h, float ms[MB_NKERNELS]);
/* With profiling on, the last K2 launch also leaves per-block phase stamps (%globaltimer, in
 * microseconds since the first block started): us[block][8] = start, tables staged, N-stars bins
 * done, star loop done, block reduced, and for the last block (-1 elsewhere): totals read + Poisson
 * term + shard totals sent to the peers, peers' totals received, end.  Returns the number of blocks
 * written (at most cap_blocks), 0 if there is no timeline, -1 on error. */
#define MB_TIMELINE_POINTS 8
int mb_last_k2_timeline(const mb_handle* h, double* us, int32_t cap_blocks);

/* One-time ingest (helpers.py:33-41 get_likelihoods): vals[k,s] = exp(lnp[k,s]) /
 * prior_weight[indxs[k,s]] on finite entries, in place on the host array, computed on `device`. */
int mb_likelihoods_from_lnp(int32_t device, int64_t K, int64_t S, const int64_t* indxs,
                            double* vals, int64_t G, const double* prior_weight);

#ifdef __cplusplus
}
#endif
#endif /* MEGABEAST_B200_H */
