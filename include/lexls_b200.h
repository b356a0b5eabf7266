/* lexls_b200.h — C-ABI of the B200-native batched lexicographic least-squares solver.
 *
 * The reference (jrl-umi3218/lexls) is a header-only C++ library with no FFI layer: its boundary is the class API of
 * LexLS::internal::LexLSE (include/lexls/lexlse.h) and LexLS::internal::LexLSI (include/lexls/lexlsi.h).  This header
 * is the boundary underneath a drop-in re-implementation of those classes (include/lexls/lexlse.h, lexlsi.h in this
 * repository are batch-of-1 facades over it) and adds the batch entry point `lexls_b200_solve_batch`.
 * Every function cites the reference member it replaces as <file>:<line> (relative to the reference's include/lexls/).
 *
 * Conventions
 *   - plain C, no exceptions cross the boundary: every call returns 0 on success or a LEXLS_B200_E* code;
 *     lexls_b200_last_error() gives the message of the last failure on the calling thread.
 *   - a batch handle owns all device memory; inputs are COPIED by the set_* calls (as the reference copies in setData,
 *     objective.h:793-796, lexlse.h:1529), so caller buffers may be released right after a call returns.
 *   - calls on one handle must be serialised by the caller; distinct handles are independent.
 *   - all problems of a batch share one shape (nVar, nObj, ObjDim[], ObjType[]); data differ per problem.
 *   - "caller layout" of level k: problem p starts at data + p*problem_stride (in doubles) and is a column-major
 *     ObjDim[k] x cols matrix with leading dimension ld (cols = nVar+1 [A|b] for LexLSE, nVar+2 [A|lb|ub] for a general
 *     LexLSI level, 2 [lb|ub] for a simple-bounds level) — i.e. what Eigen::Map<dMatrixType> over the reference's
 *     setData argument would see (typedefs.h:19,30).
 *   - *_dev variants take DEVICE pointers (same layouts) and enqueue on the handle's stream without synchronising.
 *   - input validation that needs the data (lb > ub, duplicate VarIndex) runs on the device inside the set_* calls, which do
 *     NOT wait for its verdict: the host variants return as soon as the caller's buffer has been copied.  The verdict is
 *     reported by lexls_b200_validate() (synchronises; the batch-of-1 facades call it where the reference throws, lexlsi.h:430),
 *     or else by the first synchronising call after lexls_b200_solve_batch (the solve itself is refused on the device, every
 *     status stays LEXLS_B200_STATUS_UNKNOWN).  A whole batch therefore costs one synchronisation, not one per set_* call.
 *   - a handle created by lexls_b200_create_batch_multi spans several GPUs of one box: contiguous shards, one stream and one
 *     host thread per device, no inter-device traffic; every entry point below works on it unchanged (host arrays are
 *     [batch]-major, each device reads / writes its own slice), except the *_dev variants, set_stream and device_results,
 *     which address one device: fetch the shard's own handle with lexls_b200_get_shard.
 *   - the CUDA path is the only path: there is no CPU fallback; creating a handle without a usable GPU fails.
 */
#ifndef LEXLS_B200_H
#define LEXLS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LEXLS_B200_ABI_VERSION 3

/* error codes */
#define LEXLS_B200_OK 0
#define LEXLS_B200_EINVAL 1       /* bad argument (the reference would throw LexLS::Exception, typedefs.h:300-314) */
#define LEXLS_B200_ECUDA 2        /* CUDA runtime failure */
#define LEXLS_B200_EUNSUPPORTED 3 /* feature outside the hot-path scope (regularization, least-norm solves, ...) */
#define LEXLS_B200_EBOUNDS 4      /* lb > ub somewhere in the batch (lexlsi.h:430, :474) */
#define LEXLS_B200_ECAPACITY 5    /* problem shape does not fit the device resources of any kernel variant */

/* solver kind */
#define LEXLS_B200_LSE 0 /* LexLSE: equalities only, one factorization (lexlse.h) */
#define LEXLS_B200_LSI 1 /* LexLSI: active-set loop (lexlsi.h) */

/* objective types, typedefs.h:60-64 */
#define LEXLS_B200_GENERAL_OBJECTIVE 0
#define LEXLS_B200_SIMPLE_BOUNDS_OBJECTIVE 1

/* constraint activation types, typedefs.h:69-76 */
#define LEXLS_B200_CTR_INACTIVE 0
#define LEXLS_B200_CTR_ACTIVE_LB 1
#define LEXLS_B200_CTR_ACTIVE_UB 2
#define LEXLS_B200_CTR_ACTIVE_EQ 3

/* termination status, typedefs.h:49-55 */
#define LEXLS_B200_STATUS_UNKNOWN (-1)
#define LEXLS_B200_PROBLEM_SOLVED 0
#define LEXLS_B200_PROBLEM_SOLVED_CYCLING_HANDLING 1
#define LEXLS_B200_MAX_NUMBER_OF_FACTORIZATIONS_EXCEEDED 2

/* POD mirror of ParametersLexLSI (typedefs.h:127-295; defaults :268-294) and ParametersLexLSE (:78-125).
 * regularization_type (typedefs.h:32-44): 0 NONE, 1 TIKHONOV, 2 TIKHONOV_CG, 3 R, 4 R_NO_Z, 5 RT_NO_Z, 6 RT_NO_Z_CG,
 * 7 TIKHONOV_1, 8 TIKHONOV_2, 9 TEST (lexlse.h:274-411, :1700-1760, :2078-2251, :2256-2357, :2367-2550, :2592-2625).  The two
 * conjugate-gradient types run CGLS from x = 0 for at most max_number_of_cg_iterations steps (tolerance 1e-12, as the
 * reference); type 7 regularizes exactly as the reference's regularize_tikhonov_1_test does (its X_mu / residual_mu
 * diagnostics are not produced).  Regularized batches run on the shared-memory kernels (two or more warps per problem). */
typedef struct lexls_b200_params
{
    uint32_t max_number_of_factorizations; /* 200   */
    double tol_linear_dependence;          /* 1e-12 */
    double tol_wrong_sign_lambda;          /* 1e-08 */
    double tol_correct_sign_lambda;        /* 1e-12 */
    double tol_feasibility;                /* 1e-13 */
    int32_t cycling_handling_enabled;      /* 0     */
    uint32_t cycling_max_counter;          /* 50    */
    double cycling_relax_step;             /* 1e-08 */
    int32_t regularization_type;           /* 0     */
    int32_t modify_x_guess_enabled;        /* 0     */
    int32_t modify_type_active_enabled;    /* 0     */
    int32_t modify_type_inactive_enabled;  /* 0     */
    int32_t set_min_init_ctr_violation;    /* 1     */
    int32_t use_phase1_v0;                 /* 0     */
    int32_t log_working_set_enabled;       /* 0     */
    int32_t deactivate_first_wrong_sign;   /* 0; 1 = lexlsi.h:1063-1105 (EUNSUPPORTED with a simple-bounds objective: undefined in the reference) */
    double variable_regularization_factor; /* 0.0: use the objectives' factors as they are; else typedefs.h:170-176 (lexlse.h:281-310) */
    uint32_t max_number_of_cg_iterations;  /* 10 (typedefs.h:91, :167): regularization types 2 and 6 */
} lexls_b200_params_t;

typedef struct lexls_b200_batch lexls_b200_batch_t; /* opaque */

/* per-problem counters returned by lexls_b200_get_counters (lexlsi.h:635-683, :747) */
#define LEXLS_B200_NCOUNTERS 8
#define LEXLS_B200_CNT_FACTORIZATIONS 0
#define LEXLS_B200_CNT_ACTIVATIONS 1
#define LEXLS_B200_CNT_DEACTIVATIONS 2
#define LEXLS_B200_CNT_ITERATIONS 3
#define LEXLS_B200_CNT_CYCLING 4
#define LEXLS_B200_CNT_TOTAL_RANK 5
#define LEXLS_B200_CNT_ACTIVE 6
#define LEXLS_B200_CNT_FLOPS_K 7 /* algorithmic kflop actually executed by the factorizations of this problem */

int lexls_b200_abi_version(void);
const char *lexls_b200_last_error(void);
void lexls_b200_default_params(lexls_b200_params_t *p); /* ParametersLexLSI::setDefaults, typedefs.h:268-294 */
int lexls_b200_device_count(int *count);

/* ---- life cycle -----------------------------------------------------------------------------------------------
 * create: LexLSE(nVar,nObj,ObjDim) lexlse.h:50-54 / LexLSI(nVar,nObj,ObjDim,ObjType) lexlsi.h:51-58 (+ resize :77, :923),
 * for `batch` problems on CUDA device `device`.  ObjType may be NULL (all general) and must be NULL/all-general for LSE.
 * Only level 0 may be SIMPLE_BOUNDS (lexlsi.h:926).  reset: LexLSI::reset lexlsi.h:88-104 (clears guesses, x0, v0, results;
 * keeps data and parameters). */
int lexls_b200_create_batch(int kind, uint32_t nVar, uint32_t nObj, const uint32_t *ObjDim, const int32_t *ObjType,
                            uint64_t batch, int device, lexls_b200_batch_t **out);
/* the same batch sharded over `ndev` devices (ordinals in `devices`, distinct): problems
 * [first_i, first_i + count_i) live on devices[i]; the first batch % ndev devices hold one problem more. */
int lexls_b200_create_batch_multi(int kind, uint32_t nVar, uint32_t nObj, const uint32_t *ObjDim, const int32_t *ObjType,
                                  uint64_t batch, const int *devices, int ndev, lexls_b200_batch_t **out);
int lexls_b200_shard_count(lexls_b200_batch_t *h, int *count); /* 1 for a single-device handle */
int lexls_b200_get_shard(lexls_b200_batch_t *h, int index, lexls_b200_batch_t **shard, uint64_t *first, uint64_t *count, int *device);
int lexls_b200_destroy(lexls_b200_batch_t *h);
int lexls_b200_reset(lexls_b200_batch_t *h);
int lexls_b200_set_params(lexls_b200_batch_t *h, const lexls_b200_params_t *p); /* setParameters lexlsi.h:325, lexlse.h:1467 */
int lexls_b200_set_stream(lexls_b200_batch_t *h, void *cuda_stream);             /* NULL = the handle's own stream */
/* LexLSI::setRegularizationFactor lexlsi.h:497 / LexLSE::setRegularizationFactor lexlse.h:1478: damping factor of objective
 * `level` (one value for the whole batch; ignored for a simple-bounds objective, objective.h:489) */
int lexls_b200_set_regularization_factor(lexls_b200_batch_t *h, uint32_t level, double factor);

/* ---- problem data ---------------------------------------------------------------------------------------------
 * set_data: LexLSE::setData lexlse.h:1522-1530 ([A|b]) / LexLSI::setData lexlsi.h:350-435 ([A|lb|ub], equalities are
 * detected and activated inside the solve exactly as the reference does at :374-380).  lb > ub -> EBOUNDS.
 * set_simple_bounds: LexLSI::setData(ObjIndex, VarIndex, data) lexlsi.h:444-491; var_index is 0-based, [batch][ObjDim[0]]
 * with index_stride elements between problems (0 = one index vector shared by the batch); duplicates -> EINVAL (:485).
 * (lb > ub and duplicates are found on the device: see lexls_b200_validate.)
 * set_fixed_variables: LexLSE::setFixedVariablesCount + fixVariable lexlse.h:1449-1462, :1381-1388 (LSE kind only);
 * type may be NULL (CTR_ACTIVE_UB, the reference default). */
int lexls_b200_set_data(lexls_b200_batch_t *h, uint32_t level, const double *data, uint32_t ld, uint64_t problem_stride);
int lexls_b200_set_data_dev(lexls_b200_batch_t *h, uint32_t level, const double *data, uint32_t ld, uint64_t problem_stride);
int lexls_b200_set_simple_bounds(lexls_b200_batch_t *h, const uint32_t *var_index, uint64_t index_stride, const double *lbub,
                                 uint32_t ld, uint64_t problem_stride);
int lexls_b200_set_simple_bounds_dev(lexls_b200_batch_t *h, const uint32_t *var_index, uint64_t index_stride,
                                     const double *lbub, uint32_t ld, uint64_t problem_stride);
int lexls_b200_set_fixed_variables(lexls_b200_batch_t *h, uint32_t count, const uint32_t *var_index, const double *value,
                                   const uint8_t *type);

/* ---- warm start (LSI) -----------------------------------------------------------------------------------------
 * set_active_guess: api_activate lexlsi.h:120-136 for every row whose type is LB or UB, applied level-major / row
 * ascending — the order the reference's own front-end uses (interfaces/matlab-octave/lexlsi.cpp:591-622); any other call
 * order: set_active_guess_ordered.
 * types is [batch][ObjDim[level]]; NULL clears the guess of that level.
 * set_x0 / set_v0: lexlsi.h:306-320 (x0 [batch][nVar]; v0 [batch][ObjDim[level]]); NULL clears. */
int lexls_b200_set_active_guess(lexls_b200_batch_t *h, uint32_t level, const uint8_t *types);
int lexls_b200_set_active_guess_dev(lexls_b200_batch_t *h, uint32_t level, const uint8_t *types);
/* api_activate lexlsi.h:120-136 in the caller's CALL ORDER (the order fixes active[] of workingset.h:67-77, hence the row order
 * of the LexLSE problem, its rounding and its ties): entry e of problem p is (obj, ctr, type)[p * problem_stride + e], e < count;
 * type CTR_INACTIVE pads a ragged list, CTR_ACTIVE_EQ is ignored as the reference ignores it.  Replaces any set_active_guess
 * (and vice versa); count = 0 clears. */
int lexls_b200_set_active_guess_ordered(lexls_b200_batch_t *h, uint32_t count, const uint32_t *obj, const uint32_t *ctr,
                                        const uint8_t *type, uint64_t problem_stride);
int lexls_b200_set_x0(lexls_b200_batch_t *h, const double *x0);
int lexls_b200_set_x0_dev(lexls_b200_batch_t *h, const double *x0);
int lexls_b200_set_v0(lexls_b200_batch_t *h, uint32_t level, const double *v0);

/* ---- solve ----------------------------------------------------------------------------------------------------
 * LSE: factorize + solve (+ get_v), lexlse.h:117-506, :1015-1045, :1560-1582.
 * LSI: LexLSI::solve lexlsi.h:205-246 — phase 1, then the active-set loop, entirely on the device.
 * solve_batch enqueues on the handle's stream; solve_batch_sync also waits.  Results stay on the device until read. */
int lexls_b200_solve_batch(lexls_b200_batch_t *h);
int lexls_b200_solve_batch_sync(lexls_b200_batch_t *h);
int lexls_b200_synchronize(lexls_b200_batch_t *h);
/* waits for the set_* calls issued so far and reports what their device-side checks found: EBOUNDS for lb > ub
 * (lexlsi.h:430, :474), EINVAL for duplicate / out-of-range VarIndex (lexlsi.h:485); the offending data must be set again */
int lexls_b200_validate(lexls_b200_batch_t *h);
/* By default a host set_* call returns once the caller's buffer has been consumed (the reference copies in setData).  With
 * async copies enabled the calls only ENQUEUE their copy: the caller keeps every input buffer valid and unmodified until
 * the next synchronising call (synchronize, solve_batch_sync, get_*).  For pinned host buffers this removes every host
 * wait between the set_* calls of a batch (lexls_b200.PipelinedBatch uses it). */
int lexls_b200_set_async_copies(lexls_b200_batch_t *h, int enabled);
/* Asynchronous-copies mode also DEFERS the device-side repack of every host input to the next call that needs the device
 * records (solve_batch, validate, synchronize, a result read, a *_dev setter): the inputs of a round cross the host link back
 * to back, each into its own slice of a staging arena (extra device memory: one round of inputs).  wait_copies returns when
 * every copy enqueued so far has left the caller's buffers — the point at which a pipelined caller hands the link to the next
 * handle — without waiting for repacks or solves. */
int lexls_b200_wait_copies(lexls_b200_batch_t *h);

/* ---- results (host destinations; each call synchronises the handle's stream) -----------------------------------
 * get_x: get_x lexlse.h:1587 / lexlsi.h:508 -> [batch][nVar]
 * get_v: LSE get_v lexlse.h:1560 (level rows of the factorization residual); LSI Objective::v lexlsi.h:528 -> [batch][ObjDim[level]]
 * get_constraint_violation: lexlsi.h:542-545 -> [batch][ObjDim[level]]
 * get_active_set: getActiveCtr lexlsi.h:688-698 -> [batch][ObjDim[level]] activation types
 * get_active_order: getActiveCtr_order lexlsi.h:703-716 -> [batch][ObjDim[level]] constraint indices in working-set
 *                   order, padded with 0xFFFFFFFF
 * get_status: solve() return value lexlsi.h:245 -> [batch]
 * get_counters: lexlsi.h:635-683, :747 -> [batch][LEXLS_B200_NCOUNTERS]
 * get_ranks: getRank lexlse.h:1603 -> [batch][number of LexLSE levels] (for LSI: of the last factorization) */
int lexls_b200_get_x(lexls_b200_batch_t *h, double *x);
int lexls_b200_get_v(lexls_b200_batch_t *h, uint32_t level, double *v);
int lexls_b200_get_constraint_violation(lexls_b200_batch_t *h, uint32_t level, double *viol);
int lexls_b200_get_active_set(lexls_b200_batch_t *h, uint32_t level, uint8_t *types);
int lexls_b200_get_active_order(lexls_b200_batch_t *h, uint32_t level, uint32_t *order);
int lexls_b200_get_status(lexls_b200_batch_t *h, int32_t *status);
int lexls_b200_get_counters(lexls_b200_batch_t *h, uint32_t *counters);
int lexls_b200_get_ranks(lexls_b200_batch_t *h, uint32_t *ranks);
/* LSE only: column_permutations of factorize(), lexlse.h:222 -> [batch][nVar]; entries >= TotalRank are 0 */
int lexls_b200_get_column_permutations(lexls_b200_batch_t *h, uint32_t *colperm);

/* ---- optional outputs: requested BEFORE the solve, written by the solve kernel ------------------
 * request_outputs: mask of LEXLS_B200_OUT_*; 0 switches them off again.
 * get_lexqr: get_lexqr lexlse.h:1626 / lexlsi.h:607 — the factorized work matrix of the (last) factorization
 *            -> lexqr [batch][(nVar+1) x M] column-major with leading dimension M = rows of all general levels (LSI: the
 *            active rows in working-set order, zero rows below), hh [batch][M] Householder scalars (may be NULL).
 * get_lambda: getLambda lexlsi.h:552-605 — Lagrange multipliers of every objective for the constraints of `level`
 *            -> [batch][nObj][ObjDim[level]] (per problem the reference's ObjDim x nObj column-major matrix); rows of
 *            inactive constraints are zero.  As in the reference, a problem that did not end PROBLEM_SOLVED is
 *            re-factorized with its final working set first. */
#define LEXLS_B200_OUT_LEXQR 1u
#define LEXLS_B200_OUT_LAMBDA 2u
#define LEXLS_B200_OUT_XSTAR 4u
#define LEXLS_B200_OUT_LEAST_NORM 8u
int lexls_b200_request_outputs(lexls_b200_batch_t *h, uint32_t mask);
int lexls_b200_get_lexqr(lexls_b200_batch_t *h, double *lexqr, double *hh);
int lexls_b200_get_lambda(lexls_b200_batch_t *h, uint32_t level, double *lambda);
/* get_xStar lexlsi.h:519-526 (LSI, LEXLS_B200_OUT_XSTAR): the LexLSE problem of the FINAL working set is formed, factorized and
 * solved once more at the end of the solve -> xstar [batch][nVar]; data [batch][(nVar+1) x M] = LexLSI::get_data() lexlsi.h:612
 * (its constraints [A | rhs] in working-set order, zero rows below); lexqr / hh = what get_lexqr() lexlsi.h:607 returns AFTER
 * get_xStar (tests/test_numerical_error.cpp:83-86 reads all three).  Any destination may be NULL. */
int lexls_b200_get_xstar(lexls_b200_batch_t *h, double *xstar, double *data, double *lexqr, double *hh);
/* solveLeastNorm_1 / _2 / _3 lexlse.h:1046-1262 (LSE, LEXLS_B200_OUT_LEAST_NORM): next to the basic solution of solve() (free
 * variables zero) the solve kernel computes the solution of least norm among all minimisers, by the normal equations of
 * solveLeastNorm_2 (the reference's three variants reach the same vector with different roundings) -> x [batch][nVar]. */
int lexls_b200_get_x_least_norm(lexls_b200_batch_t *h, double *x);
/* getWorkingSetLog lexlsi.h:739 (entries typedefs.h:380-432), written when params.log_working_set_enabled != 0:
 * entries [batch][capacity][6] = {obj_index, ctr_index, ctr_type (0 = removed), alpha_or_lambda, rank, cycling_detected};
 * count [batch] = valid rows of entries[p]: what the solve produced (<= max_number_of_factorizations + 1), cut at `capacity`. */
int lexls_b200_get_working_set_log(lexls_b200_batch_t *h, uint32_t capacity, double *entries, uint32_t *count);

/* device views of the result buffers (valid until destroy; contents valid after the solve completes on the stream):
 * x [batch][nVar]; status [batch]; counters [batch][NCOUNTERS] */
int lexls_b200_device_results(lexls_b200_batch_t *h, const double **x, const int32_t **status, const uint32_t **counters);

/* ---- introspection / measurement --------------------------------------------------------------------------------
 * kernel_info: which kernel variant a solve will launch ("lse_w1", "lsi_w1_smem", ...), its grid, block and dynamic
 * shared memory, resident problems per SM, and how many kernel launches one solve_batch issues.
 * last_solve_ms: duration of the last solve_batch measured with CUDA events on the handle's stream (waits for it).
 * measure_fp64_peak: FMA-chain micro-benchmarks on `device` — DFMA (CUDA cores) and DMMA (mma.sync m8n8k4 f64)
 * TFLOP/s — used as the FP64 roofline denominator (MEASURED_PEAKS.json has no FP64 entry). */
typedef struct lexls_b200_kernel_info
{
    char name[64];
    uint32_t grid, block, smem_bytes, problems_per_sm, launches_per_solve, warps_per_problem;
} lexls_b200_kernel_info_t;
int lexls_b200_kernel_info(lexls_b200_batch_t *h, lexls_b200_kernel_info_t *info);
int lexls_b200_last_solve_ms(lexls_b200_batch_t *h, float *ms);
int lexls_b200_measure_fp64_peak(int device, double *dfma_tflops, double *dmma_tflops, double *sm_clock_mhz);
/* debug aid: with LEXLS_B200_CANARY=1 in the environment every device buffer of a handle is followed by a 256-byte guard band;
 * this counts the bands that were written into (0 = no kernel wrote past the end of a buffer). */
int lexls_b200_check_guard_bands(lexls_b200_batch_t *h, uint32_t *violations);

#ifdef __cplusplus
}
#endif
#endif /* LEXLS_B200_H */
