/*
 * sgdnet_b200.h — C ABI of libsgdnet_b200.so, the B200 (sm_100a) SAGA backend for sgdnet.
 *
 * This is the drop-in boundary: the entry points below are what sgdnet's Rcpp layer
 * (reference src/sgdnet.cpp:359-375 `SgdnetDense` / `SgdnetSparse`, registered in
 * src/RcppExports.cpp:11-45 and called from R/RcppExports.R:68-75, R/sgdnet.R:362-366)
 * binds instead of running the CPU templates in src/saga-dense.h / src/saga-sparse.h.
 * Plain pointers and sizes only; no R, Rcpp, Eigen or torch types. All floating point is
 * FP64, all indices are 32-bit (the reference uses `unsigned`/`int`, SURVEY.md header).
 *
 * Conventions
 *   - dense x : n x p column-major (an R numeric matrix; Eigen::MatrixXd)
 *   - sparse x: CSC of the n x p matrix, 0-based int32 (a dgCMatrix: slots i, p, x;
 *               Eigen::SparseMatrix<double>), row indices ascending inside each column
 *   - y       : n x y_cols column-major (reference R/sgdnet.R:277-344): gaussian n x 1,
 *               binomial n x 1 coded 0/1, multinomial n x 1 of class ids 0..K-1,
 *               mgaussian n x K
 *   - every function returns 0 on success and non-zero on failure; the message is in
 *     sgdnet_last_error() (per thread). The library never keeps an input pointer after a
 *     call returns (the reference copies its inputs, src/sgdnet.cpp:121-122).
 *   - there is NO CPU fallback: without a usable CUDA device every compute entry point fails
 *     with SGDNET_ERR_CUDA.
 */
#ifndef SGDNET_B200_H_
#define SGDNET_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGDNET_ABI_VERSION 3

/* status codes */
#define SGDNET_OK            0
#define SGDNET_ERR_ARG       1   /* bad argument (the R front end validates first, R/sgdnet.R:211-263) */
#define SGDNET_ERR_CUDA      2   /* CUDA runtime / launch / no device */
#define SGDNET_ERR_ALLOC     3
#define SGDNET_ERR_RNG       4   /* index source exhausted / missing callback */
#define SGDNET_ERR_INTERNAL  5

/* control$family (reference src/sgdnet.cpp:298-335) */
#define SGDNET_GAUSSIAN      0
#define SGDNET_BINOMIAL      1
#define SGDNET_MULTINOMIAL   2
#define SGDNET_MGAUSSIAN     3

/* type.measure of score() (R/score.R:55-178) */
#define SGDNET_MEASURE_DEVIANCE 0
#define SGDNET_MEASURE_MSE      1
#define SGDNET_MEASURE_MAE      2
#define SGDNET_MEASURE_CLASS    3
#define SGDNET_MEASURE_AUC      4

/* The `control` list built at R/sgdnet.R:346-359 and read at src/sgdnet.cpp:76-78,129-138. */
typedef struct sgdnet_control {
  int32_t  family;               /* control$family                                      */
  int32_t  intercept;            /* control$intercept                                   */
  int32_t  standardize;          /* control$standardize                                 */
  int32_t  standardize_response; /* control$standardize_response (mgaussian only)       */
  int32_t  n_lambda;             /* control$n_lambda                                    */
  int32_t  n_classes;            /* control$n_classes (1, #classes, or #responses)      */
  int32_t  debug;                /* control$debug: per-epoch unpenalised mean loss      */
  int32_t  grouped_multinomial;  /* control$type_multinomial == "grouped" (R always sends 0) */
  uint32_t max_iter;             /* control$max_iter                                    */
  int32_t  lambda_len;           /* length(control$lambda); 0 => automatic path         */
  double   elasticnet_mix;       /* control$elasticnet_mix (glmnet's alpha)             */
  double   lambda_min_ratio;     /* control$lambda_min_ratio                            */
  double   tol;                  /* control$tol                                         */
  const double* lambda;          /* control$lambda, used in the given order (utils.h:157-180) */
} sgdnet_control;

/*
 * Source of the sampling sequence. The reference draws `floor(R::runif(0, n))` once per
 * sample-update from R's global RNG (src/saga-dense.h:152, src/saga-sparse.h:261; RNGScope at
 * src/RcppExports.cpp:14,27). The library consumes exactly n * epochs draws per lambda, in order.
 */
#define SGDNET_RNG_MT        0   /* R's default Mersenne-Twister; mt[]/mti are R's .Random.seed[3..626] / [2],
                                    read at entry, advanced in place (epochs that were not run are not consumed) */
#define SGDNET_RNG_CALLBACK  1   /* unif_rand(ctx) called on the calling thread, once per draw, in order      */
#define SGDNET_RNG_SEQUENCE  2   /* explicit sample indices (testing): seq[seq_pos...] consumed in order      */

typedef struct sgdnet_rng {
  int32_t  kind;
  int32_t  mti;
  uint32_t mt[624];
  double (*unif_rand)(void* ctx);
  void*    ctx;
  const uint32_t* seq;
  int64_t  seq_len;
  int64_t  seq_pos;
} sgdnet_rng;

/* R's set.seed(seed) for the Mersenne-Twister (R core RNG.c: 50 LCG scrambles, 625 LCG fills, mti=624). */
void   sgdnet_rng_set_seed(sgdnet_rng* rng, uint32_t seed);
/* One R `unif_rand()` draw from an SGDNET_RNG_MT state (MT19937 genrand * 2^-32, with R's fixup). */
double sgdnet_rng_unif(sgdnet_rng* rng);
/*
 * The sampling sequence of `n_epochs` epochs over n samples, floor(runif(0, n)) per update, as the solver's launches
 * consume it. With SGDNET_RNG_MT the library generates it ON THE DEVICE (block-parallel MT19937 regeneration), so no
 * index ever crosses PCIe; this entry point exposes that generator for verification: seq receives n * n_epochs
 * indices, states (may be NULL) n_epochs + 1 generator states, states[e] = the generator after e epochs, and *rng is
 * left advanced by n * n_epochs draws. on_host != 0 runs the same regeneration schedule on the CPU (no device needed).
 */
int    sgdnet_rng_indices(sgdnet_rng* rng, uint32_t n, int32_t n_epochs, uint32_t* seq, sgdnet_rng* states, int32_t on_host);

/*
 * The list returned at src/sgdnet.cpp:275-284. Buffers are allocated by the library and released
 * with sgdnet_result_free(). Layouts equal `unlist()` of the reference's lists.
 */
typedef struct sgdnet_result {
  int32_t   n_lambda;
  int32_t   n_classes;
  int64_t   n_features;
  double*   a0;            /* [n_lambda][K]       res$a0                                       */
  double*   beta;          /* [n_lambda][p][K]    res$beta: K x p column-major per lambda      */
  double*   lambda;        /* [n_lambda]          res$lambda (original y scale)                */
  double*   dev_ratio;     /* [n_lambda]          res$dev.ratio                                */
  uint32_t* return_codes;  /* [n_lambda]          0 converged, 1 hit max_iter                  */
  uint32_t* epochs;        /* [n_lambda]          epochs run per lambda (extension; sums to npasses) */
  double*   losses;        /* debug only: concatenated per-epoch losses                        */
  int64_t*  losses_ptr;    /* [n_lambda+1]        offsets into losses                          */
  double    nulldev;       /* res$nulldev                                                      */
  uint32_t  npasses;       /* res$npasses                                                      */
  /* measurement extensions (CUDA-event timed; seconds) */
  double    seconds_total;     /* whole call                                                   */
  double    seconds_setup;     /* host preprocessing + upload                                  */
  double    seconds_solver;    /* sum of SAGA epoch kernels                                    */
  double    seconds_deviance;  /* sum of per-lambda deviance kernels                           */
  uint64_t  kernel_launches;   /* kernels launched by this call                                */
} sgdnet_result;

void        sgdnet_result_free(sgdnet_result* r);
const char* sgdnet_last_error(void);
int         sgdnet_abi_version(void);

/* device plumbing (one process per GPU; the harness picks LOCAL_RANK) */
int sgdnet_device_count(int* count);
int sgdnet_set_device(int device);

/* ---- one path fit: replaces SgdnetDense / SgdnetSparse (src/sgdnet.cpp:359-375) ---- */
int sgdnet_fit_dense(const double* x, int64_t n, int64_t p,
                     const double* y, int32_t y_cols,
                     const sgdnet_control* control, sgdnet_rng* rng, sgdnet_result* out);

int sgdnet_fit_sparse(const int32_t* csc_i, const int32_t* csc_p, const double* csc_x,
                      int64_t n, int64_t p,
                      const double* y, int32_t y_cols,
                      const sgdnet_control* control, sgdnet_rng* rng, sgdnet_result* out);

/*
 * ---- stepping interface over device-resident data ----
 * A session is SetupSgdnet (src/sgdnet.cpp:119-215) done once: preprocessing, lambda path, step
 * sizes, upload; the warm-start state (weights, intercept, g_memory, g_sum, g_sum_intercept,
 * src/sgdnet.cpp:186-198) then lives in HBM. sgdnet_fit_* is create -> for each lambda
 * {epochs until converged; finish_lambda} -> result -> destroy.
 */
typedef struct sgdnet_session sgdnet_session;

int sgdnet_session_create_dense(const double* x, int64_t n, int64_t p,
                                const double* y, int32_t y_cols,
                                const sgdnet_control* control, sgdnet_session** out);
int sgdnet_session_create_sparse(const int32_t* csc_i, const int32_t* csc_p, const double* csc_x,
                                 int64_t n, int64_t p,
                                 const double* y, int32_t y_cols,
                                 const sgdnet_control* control, sgdnet_session** out);
/* Run `n_epochs` SAGA epochs at lambda index `lambda_ind` without host round trips in between
   (no convergence stop; for measurement). Indices come from rng. */
int sgdnet_session_run_epochs(sgdnet_session* s, int32_t lambda_ind, int32_t n_epochs, sgdnet_rng* rng,
                              float* device_ms);
/* Saga() for one lambda (src/saga-*.h do/while loop): epochs until converged or max_iter. */
int sgdnet_session_fit_lambda(sgdnet_session* s, int32_t lambda_ind, sgdnet_rng* rng,
                              uint32_t* epochs, int32_t* converged);
/* Deviance + Rescale + archive for the lambda just fitted (src/sgdnet.cpp:246-269). */
int sgdnet_session_finish_lambda(sgdnet_session* s, int32_t lambda_ind, float* device_ms);
int sgdnet_session_result(sgdnet_session* s, sgdnet_result* out);
void sgdnet_session_destroy(sgdnet_session* s);

/*
 * ---- many independent fits on one GPU: the cv_sgdnet double loop (R/cv_sgdnet.R:160-200) ----
 * Each spec is one sgdnet() call on a row subset (`x[train_ind, ]`, R/cv_sgdnet.R:182-186) followed,
 * when n_test > 0, by score(fit, x_test, y_test, "deviance") (R/cv_sgdnet.R:197-198, R/score.R:55-178).
 * All fits run concurrently, one CTA each, sharing the read-only X in HBM.
 */
typedef struct sgdnet_fit_spec {
  const int32_t* train_rows;   /* ascending 0-based row ids; NULL => all rows                  */
  int64_t        n_train;
  const int32_t* test_rows;    /* rows scored after the fit; NULL/0 => no score                */
  int64_t        n_test;
  int32_t        lambda_from;  /* -1, or the index of an EARLIER spec of the batch whose lambda path (automatic or
                                  given) this fit uses: `lambda = lambda[[i]]` of R/cv_sgdnet.R:164, 186, known after
                                  that fit's setup, so full fits and their fold fits can share one batch      */
  int32_t        measure;      /* SGDNET_MEASURE_*: type.measure of the held-out score (0 = deviance)          */
  int32_t        path_only;    /* non-zero: SetupSgdnet up to the lambda path only (src/sgdnet.cpp:119-184): the
                                  result carries `lambda` and `nulldev`, nothing is fitted. A rank of a sharded
                                  cv_sgdnet run uses it for the alphas whose full-data fit another rank owns   */
  int32_t        pad_;
  sgdnet_control control;
  sgdnet_rng     rng;
} sgdnet_fit_spec;

int sgdnet_fit_batch_dense(const double* x, int64_t n, int64_t p,
                           const double* y, int32_t y_cols,
                           sgdnet_fit_spec* specs, int32_t n_fits,
                           sgdnet_result* results, double* scores /* [n_fits][n_lambda] or NULL */);
int sgdnet_fit_batch_sparse(const int32_t* csc_i, const int32_t* csc_p, const double* csc_x,
                            int64_t n, int64_t p,
                            const double* y, int32_t y_cols,
                            sgdnet_fit_spec* specs, int32_t n_fits,
                            sgdnet_result* results, double* scores);

/*
 * ---- prediction / held-out deviance: predict.sgdnet + score.* ----
 * link[s][k][l] = a0[l][k] + sum_j x[s][j] * beta[l][j][k]   (R/predict.sgdnet.R:377, 507-510),
 * laid out [n_lambda][K][n] (n fastest). score[l] follows R/score.R:66-69 (gaussian), :103-110
 * (binomial, probabilities clamped to [1e-5, 1-1e-5]), :145-151 (multinomial), :175 (mgaussian).
 */
int sgdnet_predict_dense(const double* x, int64_t n, int64_t p,
                         const double* a0, const double* beta, int32_t n_lambda, int32_t n_classes,
                         double* link);
int sgdnet_predict_sparse(const int32_t* csc_i, const int32_t* csc_p, const double* csc_x,
                          int64_t n, int64_t p,
                          const double* a0, const double* beta, int32_t n_lambda, int32_t n_classes,
                          double* link);
/* score() with any type.measure the family accepts (R/score.R:55-178): SGDNET_MEASURE_DEVIANCE / MSE / MAE for every
   family, CLASS for binomial and multinomial, AUC for binomial. y as for the fits (binomial 0/1, multinomial class
   ids). X * beta and the per-sample measures run on the device. AUC (R/score.R:203-232) is a rank statistic whose ties
   R breaks with stats::runif, 2n draws per lambda: they are taken from `rng` (SGDNET_RNG_MT or SGDNET_RNG_CALLBACK;
   ignored and may be NULL for the other measures) on the calling thread, and the sort runs on the host. */
int sgdnet_score_dense(const double* x, int64_t n, int64_t p,
                       const double* y, int32_t y_cols, int32_t family, int32_t measure,
                       const double* a0, const double* beta, int32_t n_lambda, int32_t n_classes,
                       sgdnet_rng* rng, double* score);
int sgdnet_score_sparse(const int32_t* csc_i, const int32_t* csc_p, const double* csc_x,
                        int64_t n, int64_t p,
                        const double* y, int32_t y_cols, int32_t family, int32_t measure,
                        const double* a0, const double* beta, int32_t n_lambda, int32_t n_classes,
                        sgdnet_rng* rng, double* score);
int sgdnet_score_deviance_dense(const double* x, int64_t n, int64_t p,
                                const double* y, int32_t y_cols, int32_t family,
                                const double* a0, const double* beta, int32_t n_lambda, int32_t n_classes,
                                double* score);
int sgdnet_score_deviance_sparse(const int32_t* csc_i, const int32_t* csc_p, const double* csc_x,
                                 int64_t n, int64_t p,
                                 const double* y, int32_t y_cols, int32_t family,
                                 const double* a0, const double* beta, int32_t n_lambda, int32_t n_classes,
                                 double* score);

#ifdef __cplusplus
}
#endif
#endif /* SGDNET_B200_H_ */
