/* mpp_b200.h -- C ABI of the B200-native Marked-Point-Process RJMCMC hot path.
 *
 * The reference (Ayana-Inria/MPP_CNN_RS_object_detection) is pure Python and has no native ABI; its boundary
 * for this path is a set of Python call signatures.  This header is the C-ABI a
 * binding for that path would target: every entry point names the reference interface it stands behind
 * (paths relative to the reference root).  The Python facade in mpp_cnn_rs_object_detection_b200/ binds
 * it with ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no torch types.  All array arguments are DEVICE pointers unless
 *    the name ends in _host.  The caller (PyTorch) owns every buffer; the library owns only the opaque ctx
 *    (cell lists, per-cell density sums, scratch) bound to one device and one cudaStream_t.
 *  - Every call returns 0 on success or a negative mpp_status; mpp_last_error() gives the message of the
 *    last failure on the calling thread.  Nothing throws across the boundary.
 *  - A ctx is not thread-safe; different ctxs are independent.
 *  - Calls are asynchronous on the ctx stream unless documented as synchronising.
 */
#ifndef MPP_B200_H
#define MPP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPP_ABI_VERSION 1
#define MPP_CELL_SIZE 32      /* models/mpp/point_set/point_set.py:9,58  (cell = max(r_max, 32) px) */
#define MPP_CELL_CAPACITY 32  /* object slots per cell == lanes of a warp */
#define MPP_N_CLASSES 32      /* models/shape_net/shape_net_model.py:80-85 */
#define MPP_MAX_TERMS 8       /* energy_setup_no_calibration.py:39-50 */
#define MPP_NO_OBJECT 0xFFFFFFFFu

typedef struct mpp_ctx mpp_ctx;

typedef enum {
    MPP_OK = 0,
    MPP_ERR_INVALID = -1,      /* bad argument (python: AssertionError / ValueError) */
    MPP_ERR_CUDA = -2,         /* CUDA runtime failure */
    MPP_ERR_STATE = -3,        /* maps / model not set */
    MPP_ERR_OUT_OF_BOUNDS = -4,/* object outside the support: point_set.py:99 assert */
    MPP_ERR_CELL_FULL = -5,    /* more than MPP_CELL_CAPACITY objects in one 32x32 cell */
    MPP_ERR_NEIGHBOURHOOD = -6,/* more than the scratch capacity of candidates around one perturbation */
    MPP_ERR_NOT_FOUND = -7,    /* removal of an unknown object: energy_point_set.py:88-100 KeyError */
    MPP_ERR_TIMEOUT = -8       /* split / batched sampler: a dependency never completed (a neighbour rank is not running) */
} mpp_status;

typedef enum { MPP_PRECISION_FP32 = 0, MPP_PRECISION_FP64 = 1 } mpp_precision;
typedef enum {
    MPP_SETUP_LEGACY = 0,         /* energy_setup_legacy.py:52-86 */
    MPP_SETUP_NO_CALIBRATION = 1, /* energy_setup_no_calibration.py:56-110 */
    MPP_SETUP_TOY = 2             /* the toy terms of the reference's own tests (test/test_energy_graph.py:15-35,
                                     test/test_interacting_points_set.py:24-43): a constant unit energy and a pair
                                     energy `value if distance <(=) max_dist else 0`, max-reduced; no maps needed */
} mpp_setup;
typedef enum {
    MPP_COMB_RAW_SUM = 0,      /* energy_graph.py:132-133 (energy_combinator is None) */
    MPP_COMB_HIERARCHICAL = 1, /* combination/hierarchical.py:21-32 */
    MPP_COMB_LOGISTIC = 2,     /* combination/logistic.py:20-26 */
    MPP_COMB_MANUAL_HIERARCHICAL = 3 /* combination/hierarchical.py:41-48 */
} mpp_combinator;

/* Energy model of one chain: what EnergySetup.make_energies builds (energy_setup_legacy.py:52-86,
 * energy_setup_no_calibration.py:56-110) plus the combinator (custom_types/energy.py:8-11).
 * Term order (columns of every energy-vector output):
 *   legacy : Position, Shape, RectangleOverlap, ShapeAlignment, AreaPrior
 *   nocalib: Position, Size, Ratio, Angle, OverlapPrior, AlignmentPrior, AreaPrior[, RatioPrior] */
typedef struct {
    int32_t setup;               /* mpp_setup */
    int32_t combinator;          /* mpp_combinator */
    int32_t ratio_prior;         /* nocalib only: append RatioPriorEnergy(target_ratio) */
    int32_t rewarding;           /* ShapeAlignmentEnergy.rewarding (prior_energies.py:30) */
    double pos_threshold;        /* PositionEnergy.threshold (data_energies.py:15) */
    double remap_coef[3];        /* legacy: calibration.json param_dist_remap_coefs */
    double remap_intercept[3];   /* legacy: param_dist_remap_intercepts */
    double min_area, max_area;   /* AreaPriorEnergy (prior_energies.py:54-67) */
    double target_ratio;         /* RatioPriorEnergy (prior_energies.py:71-78) */
    double overlap_max_dist;     /* 32: energy_setup_legacy.py:70 */
    double align_max_dist;       /* 16: energy_setup_legacy.py:75 */
    /* combinator parameters.
     *  hierarchical: w[0..1]=weights_data, w[2..4]=weights_prior, w[5..6]=data_prior_weights
     *  logistic / manual: w[k] = weight of term k (term order above) */
    double comb_w[MPP_MAX_TERMS];
    double comb_bias;
    double comb_threshold;       /* detection_threshold of the indicator */
    /* MPP_SETUP_TOY only (term order: Unit, Pair).  The pair exists iff distance <= overlap_max_dist
     * (energy_graph.py:70-74); its value is toy_pair_value when distance <= toy_pair_dist (toy_pair_strict: <). */
    double toy_unit_value;
    double toy_pair_value;
    double toy_pair_dist;
    int32_t toy_pair_strict;
    /* != 0: the mark maps passed to mpp_set_maps already hold energies (the reference's pre-computed
     * ShapeEnergy.parameter_energy_map / SingleMarkEnergy.parameter_energy_map, data_energies.py:30,51): they are
     * gathered as they are, without the legacy remap / the no-calibration negation. */
    int32_t marks_are_energies;
} mpp_model_params;

/* Proposal-kernel parameters: make_kernels (rjmcmc_sampler/kernels/make_kernels.py:50-177). */
typedef struct {
    double p_kernel[8];          /* [UnifBirth, UnifDeath, DataBirth, DataDeath, GaussTrl, DataTrl, GaussTrf, DataTrf] */
    double intensity;            /* Lambda = max(1, len(init_config)): sample_rjmcmc.py:68 */
    double gauss_translation_sigma; /* 2   : make_kernels.py:124 */
    int32_t data_translation_max_delta; /* 8 : make_kernels.py:130 */
    int32_t reserved;
    double gauss_transform_sigma;   /* 0.1 (fraction of each mark range): make_kernels.py:136 */
} mpp_kernel_params;

/* One recorded proposal == models/mpp/custom_types/perturbation.py:8-12 + the kernels' `data` dicts. */
typedef struct {
    int32_t kernel;              /* 0..7, order of make_kernels.py:88-144 */
    int32_t rem_x, rem_y;        /* removed object (ignored when rem_uid == MPP_NO_OBJECT) */
    uint32_t rem_uid;
    int32_t add_x, add_y;        /* added object (ignored when add_uid == MPP_NO_OBJECT) */
    uint32_t add_uid;
    uint32_t add_cls;            /* packed classes size | ratio<<8 | angle<<16 (mappings.py:45-61, host float64) */
    double add_size, add_ratio, add_angle;
    double delta0, delta1;       /* gaussian translation delta (x,y) / gaussian transform delta */
    int32_t param_id;            /* transform kernels: 0 size, 1 ratio, 2 angle */
    int32_t new_class;           /* data-driven transform: drawn class */
    double u;                    /* accept uniform: rng.random() of rjmcmc.py:113 */
} mpp_proposal;

/* What RJMCMC.step produced for one proposal (custom_types/rjmcmc.py:6-14). */
typedef struct {
    double delta_e;              /* EPointsSet.energy_delta */
    double fwd, bwd;             /* Kernel.forward_probability / backward_probability */
    double log_alpha;            /* rjmcmc.py:105-107 */
    double temperature;
    int32_t accepted;
    int32_t n_after;
} mpp_step_result;

/* One split (kind 8) or merge (kind 9) perturbation (Perturbation of SplitKernel / MergeKernel,
 * rjmcmc_sampler/kernels/split_and_merge_kernels.py:39-178: one removal and two additions, or two removals and one addition,
 * plus the kernels' `data` dict: pos_delta, shape_delta, n_neighbors). */
typedef struct {
    int32_t kind;            /* 8 split, 9 merge */
    int32_t n_neighbors;     /* merge: objects within the radius of the first removal (-1: not drawn) */
    int32_t n_add;           /* additions: 2 (split), 1 (merge), 0 (empty perturbation) */
    int32_t reserved;
    uint32_t rem_uid[2];     /* MPP_NO_OBJECT when absent */
    int32_t rem_x[2], rem_y[2];
    int32_t add_x[2], add_y[2];
    double add_size[2], add_ratio[2], add_angle[2];
    double pos_delta[2], shape_delta[3];
    double u;                /* a fresh uniform for the accept test */
} mpp_split_merge;

/* ---------------------------------------------------------------------------------------------- library */
int mpp_abi_version(void);
const char *mpp_last_error(void);
/* sizeof of the ABI structs as compiled: 0 mpp_model_params, 1 mpp_kernel_params, 2 mpp_proposal, 3 mpp_step_result,
 * 4 mpp_window_trace, 5 mpp_split_merge */
int mpp_abi_struct_size(int which);

/* ---------------------------------------------------------------------------------------------- context
 * Stands behind PointsSet.__init__ (point_set.py:50-63) / EPointsSet.__init__ (energy_point_set.py:20-47):
 * a uniform grid of 32-px cells over a (height, width) support.  `stream` is a cudaStream_t (0 = default). */
int mpp_ctx_create(mpp_ctx **out, int device, int height, int width, int precision, void *stream);
int mpp_ctx_destroy(mpp_ctx *ctx);
/* Returns a context to its freshly created state (no objects, no maps / model / kernels, counters and uid source
 * reset) and rebinds it to `stream`, keeping its device allocations: contexts are pooled by the host layer because
 * allocating and freeing the index per image costs tens of milliseconds. */
int mpp_ctx_reset(mpp_ctx *ctx, void *stream);

/* ImageWMaps.detection_map (H,W) f32 and param_dist_maps 3x(H,W,32) f32 (custom_types/image_w_maps.py:12-22).
 * det_sum <= 0: the library reduces the map itself (float64); otherwise the caller passes
 * float(np.sum(detection_map)) so that normalised densities match shape_samplers.py:87 bit for bit.
 * Builds the per-cell density sums used by the data-driven birth sampler (utils/sampler2d.py:5-48). */
int mpp_set_maps(mpp_ctx *ctx, const float *det, const float *marks, double det_sum);
/* Band-local maps of a scene split across GPUs: det_band (rows, W) and marks_band 3x(rows, W, 32) hold the scene rows
 * [row0, row0 + rows) only (a rank needs its band plus 64 rows either side); det_sum_scene = the sum of the WHOLE detection
 * map (the data-driven birth density and the birth intensity of a window are normalised by it, shape_samplers.py:87).  Only
 * the window sampler may run on such a context (the global kernels of mpp_run_chain draw from the whole map). */
int mpp_set_maps_band(mpp_ctx *ctx, const float *det_band, const float *marks_band, int row0, int rows, double det_sum_scene);
int mpp_set_model(mpp_ctx *ctx, const mpp_model_params *model_host);
int mpp_set_kernels(mpp_ctx *ctx, const mpp_kernel_params *kernels_host);

/* ---------------------------------------------------------------------------------------------- objects
 * PointsSet.add / remove / __len__ / __iter__ (point_set.py:65-109), EPointsSet.add/remove (:72-78).
 * xy [n][2] int32 (x=row, y=col), marks [n][3] float64 (size, ratio, angle), cls [n] packed classes or NULL
 * (then classes are computed on device from the marks), uid [n] or NULL (then uids are assigned).
 * out_handle [n] uint32 (cell*32 + slot) may be NULL.  Synchronises (reports CELL_FULL / OUT_OF_BOUNDS). */
int mpp_add_objects(mpp_ctx *ctx, const int32_t *xy, const double *marks, const uint32_t *cls, const uint32_t *uid,
                    int n, uint32_t *out_handle);
int mpp_remove_objects(mpp_ctx *ctx, const uint32_t *handle, int n);
int mpp_clear_objects(mpp_ctx *ctx);
int mpp_num_objects(mpp_ctx *ctx, int *n_host);  /* synchronises */
/* Cell-major enumeration (the order of PointsSetIterator, point_set.py:12-42).  Buffers hold `capacity`
 * entries; any may be NULL.  *n_host receives the object count.  Synchronises. */
int mpp_read_objects(mpp_ctx *ctx, int capacity, uint32_t *handle, int32_t *xy, double *marks, uint32_t *uid,
                     int *n_host);

/* PointsSet.get_potential_neighbors (point_set.py:111-145): every object of the cells within ceil(radius / 32)
 * cell offsets of the cell of (x, y), except `exclude_handle` (MPP_NO_OBJECT: none).  euclidean != 0 adds the
 * distance filter of get_neighbors (point_set.py:147-149).  out_handle holds `capacity` entries; *n_host receives
 * the number found (may exceed capacity: then only the first `capacity` were written).  Synchronises. */
int mpp_query_neighbors(mpp_ctx *ctx, int x, int y, double radius, int euclidean, uint32_t exclude_handle, int capacity,
                        uint32_t *out_handle, int *n_host);
/* Device-to-device copy of the object state (PointsSet.__copy__ point_set.py:74-82, EnergyGraph.__copy__
 * energy_graph.py:92-100).  Both contexts must have the same support shape, precision and device. */
int mpp_copy_state(mpp_ctx *dst, const mpp_ctx *src);

/* ---------------------------------------------------------------------------------------------- energies
 * EnergyGraph.compute_subset(return_vector=True) (energy_graph.py:108-137) for the objects named by `handle`
 * (n of them): out_vectors [n][MPP_MAX_TERMS] (unused columns 0), out_combined [n] = combinator value of each
 * object alone (both shipped combinators are sums over objects), out_totals [2] = {raw sum (energy_graph.py:105),
 * combinator total}.  Any output may be NULL. */
int mpp_energy_vectors(mpp_ctx *ctx, const uint32_t *handle, int n, double *out_vectors, double *out_combined,
                       double *out_totals);

/* PairEnergy.compute (base_energies.py:77-80) for n pairs of stored objects: out [n][2] = {overlap kind value,
 * alignment kind value}; an entry is NaN when the pair does not exist for that kind (distance > max_dist,
 * energy_graph.py:70-74). */
int mpp_pair_values(mpp_ctx *ctx, const uint32_t *handle_a, const uint32_t *handle_b, int n, double *out);

/* EPointsSet.energy_delta (energy_point_set.py:83-100 -> energy_graph.py:139-225) for m independent
 * perturbations against the current state (none is applied).  Only rem_* / add_* of each proposal are read.
 * out_delta [m]. */
int mpp_delta_batch(mpp_ctx *ctx, const mpp_proposal *props, int m, double *out_delta);

/* ---------------------------------------------------------------------------------------------- chains
 * Replays a recorded proposal stream strictly in order: RJMCMC.step (rjmcmc.py:83-164) with the kernel draws
 * and the accept uniform taken from `props`.  Temperature starts at t0 and is multiplied by alpha_t after every
 * step while > t_target (rjmcmc.py:158-159).  out [m]. */
int mpp_replay(mpp_ctx *ctx, const mpp_proposal *props, int m, double t0, double alpha_t, double t_target,
               mpp_step_result *out);

/* Device-resident sequential chain: RJMCMC.run (rjmcmc.py:83-181) with the reference's eight global kernels
 * (make_kernels.py:88-144: uniform pick among all objects point_set.py:176-185, global births, Lambda = intensity)
 * and Philox4x32-10 in place of the numpy Generator.  One proposal at a time, strictly in order, on one warp.
 * trace (device, n_steps entries) may be NULL; counters_host[8] as in mpp_run_sweeps (may be NULL). */
int mpp_run_chain(mpp_ctx *ctx, int n_steps, double t0, double alpha_t, double t_target, uint64_t seed,
                  uint64_t step_offset, mpp_step_result *trace, unsigned long long *counters_host);

/* Kernel.sample_perturbation (base_kernels.py:18-20 and subclasses) for m independent draws against the current
 * state (nothing is applied): kernel_ids [m] (device; entry < 0: the kernel itself is drawn with p_kernel,
 * rjmcmc.py:88) -> out [m] proposals (device) whose `u` field holds a fresh accept uniform. */
int mpp_sample_proposals(mpp_ctx *ctx, const int32_t *kernel_ids, int m, uint64_t seed, uint64_t offset, mpp_proposal *out);

/* SplitKernel / MergeKernel.sample_perturbation (split_and_merge_kernels.py:51-74, 125-149) against the current state, on the
 * device: uniform pick of an object; split: position delta uniform on the quarter disc of `radius` (the reference's rejection
 * loop), three normal shape deltas of standard deviation shape_sigmas_host[i] * range_i, the two children clipped to the
 * support and to the mark ranges; merge: a uniformly drawn neighbour within `radius` (Euclidean; neighbours taken in
 * (x, y, uid) order), the merged object being the average.  Philox keyed by (seed, offset).  out: ONE record (device). */
int mpp_sample_split_merge(mpp_ctx *ctx, int kind, double radius, const double *shape_sigmas_host, uint64_t seed, uint64_t offset,
                           mpp_split_merge *out);
/* SplitKernel / MergeKernel.forward_probability / backward_probability (split_and_merge_kernels.py:76-107, 151-178) of ONE
 * perturbation (device) against the current state: out [2] = {forward, backward} (device). */
int mpp_split_merge_probs(mpp_ctx *ctx, const mpp_split_merge *perturbation, double p_split, double p_merge, double radius,
                          const double *shape_sigmas_host, double *out);

/* Kernel.forward_probability / backward_probability (base_kernels.py:22-28) of m proposals against the current
 * state: out [m][2] = {forward, backward}. */
int mpp_proposal_probs(mpp_ctx *ctx, const mpp_proposal *props, int m, double *out);

/* utils/sampler2d.py:5-48 sample_point_2d(density=...): n pixel draws (with replacement) with probability proportional to a
 * non-negative density map (height, width) f32 on the device: row prefix sums + inverse CDF (row, then column), Philox4x32-10
 * keyed by (seed, draw index).  out_xy [n][2] int32 (row, column).  scratch: device doubles, height * (width + 1) + height.
 * No context needed (the data-driven birth of a context uses its own per-cell CDF: mpp_sample_births). */
int mpp_sample_points_2d(const float *density, int height, int width, int n, uint64_t seed, int32_t *out_xy, double *scratch,
                         int device, void *stream);

/* EnergyCombinationModel.compute (custom_types/energy.py:8-11) on device for n per-object energy vectors
 * [n][MPP_MAX_TERMS] (term order of the setup): out_per_object [n] (may be NULL), out_total [1].  No ctx needed. */
int mpp_combine(const mpp_model_params *model_host, const double *vectors, int n, double *out_per_object,
                double *out_total, int device, void *stream);

/* Parallel sampler (new; replaces the sequential loop RJMCMC.run rjmcmc.py:172-181).  Each sweep visits the
 * `stride`^2 colour classes of the cell grid once; every active cell performs `proposals_per_visit` local
 * Metropolis-Hastings-Green proposals with Philox4x32-10 randomness keyed by (seed, cell, sweep).  Temperature
 * is multiplied by alpha_t once per *sweep*.  counters_host[8] (may be NULL) receives
 * {proposals, accepted, births accepted, deaths accepted, proposals whose Delta-energy was evaluated, 0, 0, 0} and
 * synchronises. */
int mpp_run_sweeps(mpp_ctx *ctx, int n_sweeps, int proposals_per_visit, int stride, double t0, double alpha_t,
                   double t_target, uint64_t seed, uint64_t sweep_offset, unsigned long long *counters_host);

/* Parallel sampler, second generation (the production path).  Sampling windows are the 32-px grid cells shifted by a
 * per-sweep pseudo-random offset, coloured 3x3; one CTA per window stages everything within 64 px of its window in
 * shared memory once and runs `proposals_per_visit` (<= 128) proposals from there; its `n_warps` (1, 2, 4, 8) warps
 * evaluate consecutive proposals speculatively (the chain does not depend on n_warps).  The kernel mixture is the
 * reference's when a window holds objects and births-only when it is empty.  n_warps = 0 selects the (slower)
 * lane-per-proposal mode.  `alpha_t` is the temperature factor of one SWEEP (the reference's per-step factor,
 * rjmcmc.py:158-159, to the power of the proposals of a sweep); inside a visit the temperature decays by
 * alpha_t^(1/proposals_per_visit) per proposal index, so the schedule as a function of the number of proposals made is
 * the reference's for any proposals_per_visit; it stops at t_target.  A proposal that maps the configuration onto itself
 * (own class / own pixel drawn) is accepted, as in the reference, without touching the state.  debug_maxdiff (device
 * float, may be NULL): every Delta-energy is also recomputed by brute force and the largest |difference| is written there.
 * schedule 0: one launch per colour class (9 per sweep, a device-wide barrier between colours).  schedule 1: one
 * persistent kernel for the whole call; window visits are claimed in (sweep, colour, window) order and each starts as soon
 * as the earlier visits within 64 px of it have completed (dataflow; same chain as schedule 0, bit for bit).
 * counters_host as in mpp_run_sweeps.  Objects born in sweep s get the uid 0x80000000 | ((s * (nx + 2) * (ny + 2) + window) * 128 +
 * proposal index); calls whose sweep numbers would take that product beyond 2^31 are refused (MPP_ERR_INVALID): restart
 * the numbering (sweep_offset) with another seed. */
int mpp_run_windows(mpp_ctx *ctx, int n_sweeps, int proposals_per_visit, int n_warps, int schedule, double t0,
                    double alpha_t, double t_target, uint64_t seed, uint64_t sweep_offset,
                    unsigned long long *counters_host, float *debug_maxdiff);

/* Window sampler over a BATCH of independent scenes of equal shape on one device, or over the band of one scene that this
 * rank owns (mpp_split_attach*), in ONE persistent dataflow launch (csrc/mpp_multi.cuh).  The reference maps independent
 * 256x256 patches over a process pool (mpp_model.py:231-264, train_utils.py:11-18); here the window visits of all scenes are
 * claimed from one queue in (sweep, colour, scene, window) order, so a batch of small tiles fills the GPU like one large
 * scene.  grid_seed fixes the per-sweep grid offsets (shared by the batch), seeds_host[k] the random streams of scene k: a
 * scene follows exactly the chain mpp_run_windows(seed) gives it when grid_seed == seeds_host[k].  The scenes must have run
 * the same sweeps before (completion stamps are monotone and shared: reset the contexts together).  n_warps: 4 or 8.
 * max_ctas > 0 caps the persistent grid (several batches / bands running concurrently on one device must fit together).
 * counters_host[8]: totals over the scenes (as in mpp_run_sweeps; synchronises), may be NULL.
 * ctxs_host: host array of contexts; the launch and its scratch live on the first context's stream. */
int mpp_run_windows_batch(mpp_ctx **ctxs_host, const uint64_t *seeds_host, int n_scenes, uint64_t grid_seed, int n_sweeps,
                          int proposals_per_visit, int n_warps, double t0, double alpha_t, double t_target,
                          uint64_t sweep_offset, int max_ctas, unsigned long long *counters_host, float *debug_maxdiff);

/* One scene split into row bands across GPUs (BASELINE configs[3]; the reference has no counterpart: its largest unit of
 * work is one 256^2 patch).  Every rank creates a context of the WHOLE scene's shape, holds the objects whose 32-px cell
 * row lies in its band [row_lo, row_hi) and samples the windows that start in it; the cells just across a band boundary are
 * read and written in the neighbour's context through peer-mapped memory (NVLink P2P), and window completions are stamped
 * into the neighbour's completion grid, so mpp_run_windows_batch(one scene) on every rank at once is the single-GPU chain
 * of mpp_run_windows, bit for bit, without any exchange step.
 *  mpp_split_export: CUDA IPC handles of this context's {occupancy masks, records, completion grid} (3 x 64 bytes) for the
 *                    neighbour ranks (other processes).
 *  mpp_split_attach: row band + the handles exported by the upper / lower neighbour (NULL: none, image border).
 *  mpp_split_attach_local: the same for neighbour contexts of the same process (one process driving several GPUs with
 *                    peer access, or several bands on one GPU in the tests).
 * Bands are whole 32-px cell rows, at least 384 rows.  All ranks must have attached (host barrier) before any of them runs,
 * and all must issue the same sequence of mpp_run_windows_batch calls. */
#define MPP_IPC_HANDLE_BYTES 64
int mpp_split_export(mpp_ctx *ctx, unsigned char *handles_host);
int mpp_split_attach(mpp_ctx *ctx, int row_lo, int row_hi, const unsigned char *up_handles_host,
                     const unsigned char *down_handles_host);
int mpp_split_attach_local(mpp_ctx *ctx, int row_lo, int row_hi, mpp_ctx *up, mpp_ctx *down);
int mpp_split_detach(mpp_ctx *ctx);

/* Per-kernel statistics of the window sampler since the last call (read and reset; synchronises).  The reference keeps the
 * same tallies per kernel in RJMCMC.run's log (rjmcmc.py:115-156: kernel name, accepted).  out_host[MPP_WINDOW_STATS]:
 *   [0..7]   proposals evaluated from an EMPTY window (births-only mixture), by kernel id
 *   [8..15]  proposals evaluated from an occupied window (the reference mixture), by kernel id
 *   [16..31] accepted, same layout
 *   [32] accepted proposals that map the configuration onto itself, [33] window visits, [34] visits that found the window empty
 * Kernel ids: 0 uniform birth, 1 uniform death, 2 data-driven birth, 3 data-driven death, 4 gaussian translation,
 * 5 data-driven translation, 6 gaussian mark transform, 7 data-driven mark transform (make_kernels.py:88-144). */
#define MPP_WINDOW_STATS 35
int mpp_window_stats(mpp_ctx *ctx, unsigned long long *out_host);

/* Per-proposal trace of the window sampler: what RJMCMC.step (rjmcmc.py:83-164) computes for one step -- the kernel
 * drawn (:88), the perturbation (:90), the Delta-energy (:93-100), the proposal densities (:102-103), the temperature and
 * the accept test (:105-113) -- written by the debug instantiation of the production kernels (mpp_run_windows with
 * debug_maxdiff != NULL) so that a CPU oracle can re-derive every Green ratio from first principles.  One 64-byte record
 * per proposal of the chain; a proposal that was evaluated speculatively and discarded is overwritten when it is re-drawn,
 * so the buffer ends up holding exactly the chain.  Record of proposal `it` of the visit of window (wi, wj) in sweep s:
 *   index = (s - sweep0) * (nx + 2) * (ny + 2) * proposals_per_visit + (wi * (ny + 2) + wj) * proposals_per_visit + it
 * with nx = ceil(H / 32), ny = ceil(W / 32).  Records past `capacity` are dropped. */
#define MPP_TRACE_WRITTEN 1u      /* the proposal was drawn */
#define MPP_TRACE_EVALUATED 2u    /* its Delta-energy was evaluated (otherwise: rejected before, see the REASON bits) */
#define MPP_TRACE_ACCEPT 4u
#define MPP_TRACE_IDENTITY 8u     /* the proposal maps the configuration onto itself (accepted, nothing committed) */
#define MPP_TRACE_HAS_ADD 16u
#define MPP_TRACE_HAS_REM 32u
#define MPP_TRACE_LEFT_WINDOW 64u /* a move whose end point lies outside the window (or a data-driven kernel over zero mass) */
#define MPP_TRACE_CELL_FULL 128u  /* destination storage cell has no free slot */
typedef struct {
    uint32_t flags;          /* MPP_TRACE_* | kernel << 8 | min(n_window_objects, 255) << 16 | param_id << 24 */
    uint32_t rem_uid;        /* uid of the removed object (HAS_REM) */
    uint32_t add_uid;        /* uid the added object gets if accepted (HAS_ADD) */
    uint32_t add_cls;        /* packed classes of the added object: size | ratio << 8 | angle << 16 */
    int32_t add_x, add_y;
    float add_size, add_ratio, add_angle;
    float delta_e;           /* Delta-energy as the kernel computed it (top-2 partner reductions, float32) */
    float log_ratio;         /* log(backward density) - log(forward density) */
    float temperature;
    float u_accept;          /* the accept uniform: accepted iff log(u + 1e-16) < -delta_e / temperature + log_ratio */
    uint32_t q[3];           /* random words: kernel choice, object choice, accept (Philox4x32-10 keyed by seed, counter
                                (0|1, wi * 65536 + wj, sweep, it ^ 0x77000000)) */
} mpp_window_trace;

/* Arms (buf != NULL) or disarms the trace of the next mpp_run_windows(debug_maxdiff != NULL) calls on this context.
 * buf: device memory for `capacity` records, zero-filled by the caller; sweep0: sweep id of the first traced sweep. */
int mpp_set_window_trace(mpp_ctx *ctx, mpp_window_trace *buf, uint64_t capacity, uint64_t sweep0);

/* One third of a sweep of the window sampler, restricted to a band of rows: the three colour launches (cj = 0, 1, 2) of
 * the window rows wi = ci (mod 3) whose first pixel row lies in [row_lo, row_hi).  Window rows of equal ci are >= 65 px
 * apart, so ranks that own disjoint row bands of one scene can run this concurrently and only need to exchange their
 * boundary objects (mpp_pack_rows / mpp_unpack_rows) between two values of ci.  The grid offset, the random streams and
 * the uids of born objects depend only on (seed, sweep_id, window): a split scene follows the same chain as
 * mpp_run_windows(schedule 0) on one GPU (one temperature per call).  mpp_window_grid returns the grid offset of a sweep. */
int mpp_run_window_rows(mpp_ctx *ctx, int proposals_per_visit, int n_warps, double temperature, uint64_t seed,
                        uint64_t sweep_id, int ci, int row_lo, int row_hi);
int mpp_window_grid(mpp_ctx *ctx, uint64_t seed, uint64_t sweep_id, int *ox_host, int *oy_host);

/* Draws n pixels from the normalised detection map (sample_point_2d, utils/sampler2d.py:39-46, as used by
 * RectangleSampler.sample shape_samplers.py:90-94) and their three mark classes (shape_samplers.py:113-117):
 * out [n][5] int32 = x, y, class_size, class_ratio, class_angle. */
int mpp_sample_births(mpp_ctx *ctx, int n, uint64_t seed, int32_t *out);

/* ---------------------------------------------------------------------------------------------- init
 * naive_detection (sample_rjmcmc.py:23-35): threshold -> greedy distance NMS (utils/nms.py:68-109, 6 px) ->
 * argmax marks; the surviving objects are inserted into the ctx.  *n_host receives their number. */
int mpp_naive_init(mpp_ctx *ctx, double detection_threshold, double nms_distance, int *n_host);

/* ---------------------------------------------------------------------------------------------- multi-GPU
 * Scene split in row strips: objects whose row lies in [row_lo, row_hi) are packed as 8-double records
 * (x, y, size, ratio, angle, cls, uid, 0) into `buf` (capacity records); *n_host receives the count.
 * mpp_unpack_halo first deletes every object with row in [row_lo, row_hi) and then inserts the records. */
int mpp_pack_rows(mpp_ctx *ctx, int row_lo, int row_hi, double *buf, int capacity, int *n_host);
int mpp_unpack_rows(mpp_ctx *ctx, int row_lo, int row_hi, const double *buf, int n);

#ifdef __cplusplus
}
#endif
#endif /* MPP_B200_H */
