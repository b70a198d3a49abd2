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
  uint32_t round_id;     // published with the fit's Progress when the launch is over (publish_progress)
  uint32_t pad_;
};

// Every launcher below works on ONE fit (its FitDev / Progress in device memory) on the stream it is given: the engine
// runs each fit of a batch as its own asynchronous pipeline (engine.cu).

// SAGA epochs, one persistent CTA per fit (saga_dense.cu / saga_sparse.cu).
size_t dense_smem_bytes(int K, int p, int ld, int* state_in_smem);
size_t dense_smem_budget();
cudaError_t launch_saga_dense(int K, int pen, size_t smem, FitDev* fit, Progress* prog, const RoundArgs& ra, cudaStream_t st);
int dense_kt_bucket(int K);
// 1, 4, 8, 16 or 32: the class-count bucket a fit's kernel instantiation is compiled for
// designs with p >= SGD_WIDE_P: one thread-block cluster of 8 CTAs per fit (saga_dense_cluster.cu; shapes without a
// compile-time instantiation there go to saga_dense_cluster_generic.cu)
size_t dense_cluster_smem_bytes(int K, int p, int pen);
cudaError_t launch_saga_dense_cluster(int K, int p, int pen, size_t smem, FitDev* fit, Progress* prog, const RoundArgs& ra, cudaStream_t st);
cudaError_t launch_saga_sparse(bool fast_k1, FitDev* fit, Progress* prog, const RoundArgs& ra, cudaStream_t st);
// sparse, K == 1, virtual centring (standardize = TRUE), rows of at most kCentCap nonzeros: the owner-computes kernel
// (saga_sparse_centred.cu). pos_global: 2 * p uint16 of scratch, used when the state does not fit shared memory.
constexpr int kCentCap = 256;
size_t centred_smem_bytes(int p, bool* state_in_smem);
// mode: 0 elastic net with alpha * gamma == 0 on the whole path (the lasso), 1 elastic net, 2 ridge, 3 general
cudaError_t launch_saga_sparse_centred(int p, int mode, FitDev* fit, Progress* prog, const RoundArgs& ra, uint16_t* pos_global, cudaStream_t st);

// Conflict codes of a staged sequence for the wavefront kernel (sparse K == 1); a function of the sequence alone, so it
// runs ahead of the solver launch that consumes it (on the fit's second stream, while the previous launch solves).
cudaError_t launch_wave_deps(const FitDev* fit, const RoundArgs& ra, int64_t rows, int ctas, cudaStream_t st);
int wave_warps();

// R's Mersenne-Twister on the device (rng.cu): the sampling sequence of `n_epochs` epochs, floor(n * unif_rand()) per
// update (src/saga-sparse.h:261, src/saga-dense.h:152), from the generator state `src`; snaps[e] is the generator after
// e epochs (snaps[0] == *src), which is where the next launch resumes when the solver stopped after e epochs.
struct MtState {
  uint32_t mt[624];
  int32_t mti;
  int32_t pad_[3];
};
cudaError_t launch_mt_indices(const MtState* src, uint32_t n, int n_epochs, uint32_t* seq, MtState* snaps, cudaStream_t st);
// the same block regeneration on the host, for the CPU unit test of the parallel schedule (tests/test_abi_cpu.py)
void mt_indices_host(const MtState* src, uint32_t n, int n_epochs, uint32_t* seq, MtState* snaps);

// passes.cu
cudaError_t launch_lag_scaling(FitDev* fit, Progress* prog, cudaStream_t st);
// mask_words: 32-bit words of the nonzero-coefficient bitmap the deviance pass may stage in shared memory (0: none)
// tiles > 0 (sparse, K == 1, no virtual centring): the bulk-copy tile form of the loss pass on that many CTAs
cudaError_t launch_finish_lambda(FitDev* fit, Progress* prog, int blocks, int tiles, int mask_words, uint32_t round_id, cudaStream_t st);
cudaError_t launch_epoch_loss(FitDev* fit, Progress* prog, int blocks, int tiles, cudaStream_t st);
int loss_mask_words_max(bool tiles);

struct PredictArgs {
  int32_t sparse, family, K, Ky, p, ld, n_lambda, measure;   // measure: SGDNET_MEASURE_* (include/sgdnet_b200.h)
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
  const int32_t* remap;      // multinomial "class": predicted class -> id among the classes predicted anywhere, or NULL
  uint32_t* present;         // multinomial "class": bitmap of the classes predicted anywhere (atomicOr), or NULL
};
// transpose: fill bt_scratch from a.beta first (false when a second pass reuses it)
cudaError_t launch_predict_score(const PredictArgs& a, double* bt_scratch /* [p][L*K] */, int blocks, cudaStream_t st, bool transpose);

}  // namespace sgd
