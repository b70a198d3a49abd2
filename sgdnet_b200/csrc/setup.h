// setup.h — launchers of the design-level setup kernels (setup.cu): the device side of SetupSgdnet before the lambda loop.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace sgd {

// ---- CSC (the caller's dgCMatrix, on the device) -> padded CSR
// 1. row counts + padded totals: totals_dev[0] = entries of the padded CSR, totals_dev[1] = longest row
cudaError_t csc_to_csr_counts(const int32_t* csc_i, int64_t nnz, int64_t n, int32_t* counts, int64_t* block_tot,
                              int64_t* totals_dev, cudaStream_t st);
int csc_to_csr_scan_blocks(int64_t n);      // length of block_tot
// 2. row descriptors, fill, per-row sort by column id (scratch_*: same length as ci / cv, only read when a row has more
//    than 128 entries; may be null otherwise)
cudaError_t csc_to_csr_fill(const int32_t* csc_i, const int32_t* csc_p, const double* csc_x, int64_t n, int32_t p,
                            const int32_t* counts, const int64_t* block_off, RowInfo* rows, int32_t* cursor, int32_t* ci,
                            double* cv, int32_t* scratch_k, double* scratch_v, cudaStream_t st);

// ---- row subsets (`x[train_ind, ]`)
cudaError_t subset_local(const int32_t* subset_dev, int64_t n_sub, int64_t n_src, int32_t* local, cudaStream_t st);
cudaError_t subset_counts(const RowInfo* src, const int32_t* subset_dev, int64_t n_sub, int32_t* counts, int64_t* block_tot,
                          int64_t* totals_dev, cudaStream_t st);
cudaError_t gather_rows(const RowInfo* src_rows, const int32_t* src_ci, const double* src_cv, const int32_t* subset_dev,
                        int64_t n_sub, const int32_t* counts, const int64_t* block_off, RowInfo* rows, int32_t* cursor,
                        int32_t* ci, double* cv, const double* x_scale, cudaStream_t st);

// ---- sparse statistics (sequential per column / per row, the reference's order)
cudaError_t sparse_col_stats(const int32_t* csc_i, const int32_t* csc_p, const double* csc_x, int32_t p, const int32_t* local,
                             int64_t n_sub, double* x_center, double* x_scale, double* c, cudaStream_t st);
cudaError_t sparse_xt_times(const int32_t* csc_i, const int32_t* csc_p, const double* csc_x, int32_t p, const int32_t* local,
                            int64_t n_sub, const double* x_scale, const double* ymap, int m, double* out, cudaStream_t st);
cudaError_t sparse_norm_max(const RowInfo* rows, const int32_t* ci, const double* cv, int64_t n, int32_t p, const double* c,
                            unsigned long long* out_bits, cudaStream_t st);

// ---- dense: column statistics, standardise + transpose to sample-major, largest row norm, X^T y
cudaError_t dense_design(const double* x, int64_t n_src, int32_t p, const int32_t* subset_dev, int64_t n_sub, int32_t ld,
                         bool standardize, double* x_center, double* x_scale, double* xd, unsigned long long* norm_bits,
                         cudaStream_t st);
cudaError_t dense_xt_times(const double* xd, int64_t n, int32_t p, int32_t ld, const double* ymap, int m, double* out,
                           cudaStream_t st);

}  // namespace sgd
