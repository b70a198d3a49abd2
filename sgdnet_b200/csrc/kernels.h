// kernels.h — host-callable launchers of the sgdnet_b200 CUDA kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace sgd {

// Per-fit arguments that change with every round of launches.
struct RoundArgs {
  const uint32_t* seq;   // n * n_epochs sample indices for this launch
  uint64_t* dep;         // sparse K == 1: [n * n_epochs][32] conflict codes written by wave_deps_kernel
  uint8_t* dup;          // sparse K == 1: [n * n_epochs] distance to the last in-window row of the same sample
  int32_t n_epochs;      // epochs this launch may run (0 => the fit sits this round out)
  int32_t flags;         // bit 0: measurement mode - run exactly n_epochs, ignore convergence, stay kRunning
};

// SAGA epochs, one CTA per fit (saga_dense.cu / saga_sparse.cu).
size_t dense_smem_bytes(int K, int p, int ld, int* state_in_smem);
size_t dense_smem_budget();
cudaError_t launch_saga_dense(int n_fits, unsigned kts, unsigned pens, size_t smem, FitDev* fits, Progress* prog,
                              const RoundArgs* args, cudaStream_t st);
int dense_kt_bucket(int K);   // 1, 4, 8, 16 or 32: the class-count bucket a fit's kernel instantiation is compiled for
cudaError_t launch_saga_sparse(int n_fits, bool fast_k1, FitDev* fits, Progress* prog, const RoundArgs* args,
                               cudaStream_t st);

// Conflict codes of the staged sequences for the wavefront kernel (sparse K == 1); must precede launch_saga_sparse.
cudaError_t launch_wave_deps(int n_fits, const FitDev* fits, const Progress* prog, const RoundArgs* args,
                             int64_t max_rows, int sms, cudaStream_t st);
int wave_warps();

// passes.cu
cudaError_t launch_lag_scaling(int n_fits, FitDev* fits, Progress* prog, cudaStream_t st);
cudaError_t launch_finish_lambda(int n_fits, FitDev* fits, Progress* prog, int blocks_per_fit, cudaStream_t st);
cudaError_t launch_epoch_loss(int n_fits, FitDev* fits, Progress* prog, const RoundArgs* args, int blocks_per_fit,
                              cudaStream_t st);

struct PredictArgs {
  int32_t sparse, family, K, Ky, p, ld, n_lambda, pad_;
  int64_t n;                 // rows to score
  const int32_t* row_ids;    // optional subset of the design's rows (NULL => 0..n-1)
  const double* xd;          // dense raw [n_total][ld]
  const RowInfo* rows;       // sparse raw padded CSR
  const int32_t* ci;
  const double* cv;
  const double* y;           // [n_total][Ky] sample-major, caller's scale; NULL => no score
  const double* a0;          // [L][K]
  const double* beta;        // [L][p][K]
  double* link;              // [L][K][n] or NULL
  double* partials;          // [blocks][L]
  double* score;             // [L] or NULL
};
cudaError_t launch_predict_score(const PredictArgs& a, double* bt_scratch /* [p][L*K] */, int blocks, cudaStream_t st);

}  // namespace sgd
