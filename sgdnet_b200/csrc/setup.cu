// setup.cu — the design-level part of SetupSgdnet on the device.
//
//   AdaptiveTranspose        src/utils.h:276-288    csc_to_csr_*: the caller's CSC -> padded CSR (row counts, exclusive
//                                                   scan, fill, per-row sort by column id)
//   PreprocessFeatures       src/utils.h:99-121     column means / standard deviations (src/math.h:66-160), sparse:
//                                                   scale only, centring stays virtual (src/sgdnet.cpp:150-151)
//   ColNormsMax              src/utils.h:60-85      largest squared row norm (of x_s - c when centring is virtual)
//   X^T y of LambdaMax       src/families.h:119-126, 203-220, 300-325, 387-406
//   row subsets              R/cv_sgdnet.R:182-186  `x[train_ind, ]` gathered on the device
//
// Every floating point reduction here is a SEQUENTIAL sum in the reference's order (a column's entries in ascending
// row order, a row's in ascending column order), one thread per column or per row, so means, scales, step sizes and
// the lambda path keep the bits of the reference's CPU code; the parallelism is across columns / rows. Integer work
// (counts, scans, the row sort) is order-free.
#include <algorithm>

#include "common.cuh"
#include "setup.h"

namespace sgd {

namespace {
constexpr int kT = 256;
inline int blocks_for(int64_t items, int per_block = kT) { return static_cast<int>(std::max<int64_t>(1, (items + per_block - 1) / per_block)); }
}  // namespace

// ------------------------------------------------------------------------------------------ CSC -> padded CSR
__global__ void row_count_kernel(const int32_t* __restrict__ csc_i, int64_t nnz, int32_t* __restrict__ counts) {
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < nnz; e += int64_t(gridDim.x) * blockDim.x)
    atomicAdd(&counts[csc_i[e]], 1);
}

// padded row starts: exclusive scan of (count + 3) & ~3. Three phases: per-block totals, scan of the block totals by one
// block, per-block scan with the block's offset.
constexpr int kScanItems = 2048;      // rows per block
__global__ void scan_block_totals_kernel(const int32_t* __restrict__ counts, int64_t n, int64_t* __restrict__ block_tot,
                                         int64_t* __restrict__ max_count) {
  __shared__ int64_t red[kT / 32];
  const int64_t base = int64_t(blockIdx.x) * kScanItems;
  int64_t a = 0;
  int32_t mx = 0;
  for (int i = threadIdx.x; i < kScanItems; i += kT) {
    const int64_t r = base + i;
    if (r < n) {
      a += (counts[r] + 3) & ~3;
      mx = max(mx, counts[r]);
    }
  }
  mx = __reduce_max_sync(0xffffffffu, mx);
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned long long*>(max_count), static_cast<unsigned long long>(mx));
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    int64_t t = 0;
    for (int w = 0; w < kT / 32; ++w) t += red[w];
    block_tot[blockIdx.x] = t;
  }
}
__global__ void scan_of_totals_kernel(int64_t* __restrict__ block_tot, int n_blocks, int64_t* __restrict__ total_out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int64_t run = 0;
  for (int b = 0; b < n_blocks; ++b) {
    const int64_t t = block_tot[b];
    block_tot[b] = run;
    run += t;
  }
  *total_out = run;
}
__global__ void scan_rows_kernel(const int32_t* __restrict__ counts, int64_t n, const int64_t* __restrict__ block_off,
                                 RowInfo* __restrict__ rows, int32_t* __restrict__ cursor) {
  // one warp scans the block's rows in order, 32 at a time
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  int64_t run = block_off[blockIdx.x];
  const int64_t base = int64_t(blockIdx.x) * kScanItems;
  for (int i0 = 0; i0 < kScanItems; i0 += 32) {
    const int64_t r = base + i0 + lane;
    const int32_t c = (r < n) ? counts[r] : 0;
    int64_t v = (c + 3) & ~3;
    int64_t incl = v;
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    if (r < n) {
      RowInfo ri;
      ri.start = run + incl - v;
      ri.nnz = c;
      ri.pad_ = 0;
      rows[r] = ri;
      cursor[r] = 0;
    }
    run += __shfl_sync(0xffffffffu, incl, 31);
  }
}

// one warp per column: every entry to its row, in whatever order the atomics resolve (the row sort below fixes it)
__global__ void fill_rows_kernel(const int32_t* __restrict__ csc_i, const int32_t* __restrict__ csc_p,
                                 const double* __restrict__ csc_x, int32_t p, const RowInfo* __restrict__ rows,
                                 int32_t* __restrict__ cursor, int32_t* __restrict__ ci, double* __restrict__ cv) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t j = warp; j < p; j += n_warps) {
    const int64_t lo = csc_p[j], hi = csc_p[j + 1];
    for (int64_t e = lo + lane; e < hi; e += 32) {
      const int32_t i = csc_i[e];
      const int64_t dst = rows[i].start + atomicAdd(&cursor[i], 1);
      ci[dst] = static_cast<int32_t>(j);
      cv[dst] = csc_x[e];
    }
  }
}

// one warp per row: sort the row's (column id, value) pairs by column id (ids are distinct inside a row) and zero the
// pad entries. Up to 128 entries: bitonic network over 4 registers x 32 lanes; longer rows: rank by counting.
__device__ __forceinline__ void cmp_swap(int32_t& ka, double& va, int32_t& kb, double& vb, bool up) {
  const bool sw = (ka > kb) == up;
  const int32_t tk = sw ? kb : ka;
  const double tv = sw ? vb : va;
  kb = sw ? ka : kb;
  vb = sw ? va : vb;
  ka = tk;
  va = tv;
}
__global__ void sort_rows_kernel(const RowInfo* __restrict__ rows, int64_t n, int32_t* __restrict__ ci, double* __restrict__ cv,
                                 int32_t* __restrict__ long_scratch_k, double* __restrict__ long_scratch_v) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = warp; r < n; r += n_warps) {
    const RowInfo ri = rows[r];
    int32_t* __restrict__ k = ci + ri.start;
    double* __restrict__ v = cv + ri.start;
    const int nnz = ri.nnz;
    const int padded = (nnz + 3) & ~3;
    if (nnz <= 128) {
      // element index of (register q, lane l) is q * 32 + l; missing entries sort to the end (key INT_MAX)
      int32_t kk[4];
      double vv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int e = q * 32 + lane;
        kk[q] = e < nnz ? k[e] : 0x7fffffff;
        vv[q] = e < nnz ? v[e] : 0.0;
      }
      // bitonic sort of 128 elements: for size = 2..128, stride = size/2..1
#pragma unroll
      for (int size = 2; size <= 128; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
          if (stride >= 32) {
            // partner differs in the register index
            const int qs = stride >> 5;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if ((q & qs) == 0) {
                const int e = q * 32 + lane;
                const bool up = (e & size) == 0;
                cmp_swap(kk[q], vv[q], kk[q | qs], vv[q | qs], up);
              }
            }
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int e = q * 32 + lane;
              const bool up = (e & size) == 0;
              const int32_t pk = __shfl_xor_sync(0xffffffffu, kk[q], stride);
              const double pv = __shfl_xor_sync(0xffffffffu, vv[q], stride);
              const bool lower = (lane & stride) == 0;
              const bool take = lower ? ((kk[q] > pk) == up) : ((pk > kk[q]) == up);
              kk[q] = take ? pk : kk[q];
              vv[q] = take ? pv : vv[q];
            }
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int e = q * 32 + lane;
        if (e < nnz) {
          k[e] = kk[q];
          v[e] = vv[q];
        } else if (e < padded) {
          k[e] = 0;
          v[e] = 0.0;
        }
      }
    } else {
      // long row: stage in scratch (same offsets), then every entry goes to the position given by its rank
      for (int e = lane; e < nnz; e += 32) {
        long_scratch_k[ri.start + e] = k[e];
        long_scratch_v[ri.start + e] = v[e];
      }
      __syncwarp();
      for (int e = lane; e < nnz; e += 32) {
        const int32_t me = long_scratch_k[ri.start + e];
        int rank = 0;
        for (int q = 0; q < nnz; ++q) rank += long_scratch_k[ri.start + q] < me ? 1 : 0;
        k[rank] = me;
        v[rank] = long_scratch_v[ri.start + e];
      }
      for (int e = nnz + lane; e < padded; e += 32) {
        k[e] = 0;
        v[e] = 0.0;
      }
      __syncwarp();
    }
  }
}

// totals_dev[0] = padded entries in all, totals_dev[1] = longest row
cudaError_t csc_to_csr_counts(const int32_t* csc_i, int64_t nnz, int64_t n, int32_t* counts, int64_t* block_tot,
                              int64_t* totals_dev, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(int32_t) * n, st);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(totals_dev, 0, 2 * sizeof(int64_t), st);
  if (e != cudaSuccess) return e;
  row_count_kernel<<<std::min(blocks_for(nnz), 148 * 16), kT, 0, st>>>(csc_i, nnz, counts);
  const int nb = blocks_for(n, kScanItems);
  scan_block_totals_kernel<<<nb, kT, 0, st>>>(counts, n, block_tot, totals_dev + 1);
  scan_of_totals_kernel<<<1, 32, 0, st>>>(block_tot, nb, totals_dev);
  return cudaGetLastError();
}
int csc_to_csr_scan_blocks(int64_t n) { return blocks_for(n, kScanItems); }

cudaError_t csc_to_csr_fill(const int32_t* csc_i, const int32_t* csc_p, const double* csc_x, int64_t n, int32_t p,
                            const int32_t* counts, const int64_t* block_off, RowInfo* rows, int32_t* cursor, int32_t* ci,
                            double* cv, int32_t* scratch_k, double* scratch_v, cudaStream_t st) {
  scan_rows_kernel<<<blocks_for(n, kScanItems), 32, 0, st>>>(counts, n, block_off, rows, cursor);
  fill_rows_kernel<<<std::min(blocks_for(int64_t(p) * 32), 148 * 16), kT, 0, st>>>(csc_i, csc_p, csc_x, p, rows, cursor, ci, cv);
  sort_rows_kernel<<<std::min(blocks_for(n * 32), 148 * 16), kT, 0, st>>>(rows, n, ci, cv, scratch_k, scratch_v);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------ row subsets
// local[r] = position of source row r in the subset, or -1
__global__ void subset_local_kernel(const int32_t* __restrict__ subset, int64_t n_sub, int32_t* __restrict__ local) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_sub; i += int64_t(gridDim.x) * blockDim.x)
    local[subset[i]] = static_cast<int32_t>(i);
}
__global__ void subset_counts_kernel(const RowInfo* __restrict__ src, const int32_t* __restrict__ subset, int64_t n_sub,
                                     int32_t* __restrict__ counts) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_sub; i += int64_t(gridDim.x) * blockDim.x)
    counts[i] = src[subset ? subset[i] : i].nnz;
}
// one warp per row: copy (and, when standardising, scale: v / x_scale[j], src/utils.h:118-120)
__global__ void gather_rows_kernel(const RowInfo* __restrict__ src_rows, const int32_t* __restrict__ src_ci,
                                   const double* __restrict__ src_cv, const int32_t* __restrict__ subset, int64_t n_sub,
                                   const RowInfo* __restrict__ rows, int32_t* __restrict__ ci, double* __restrict__ cv,
                                   const double* __restrict__ x_scale) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t i = warp; i < n_sub; i += n_warps) {
    const RowInfo s = src_rows[subset ? subset[i] : i];
    const RowInfo d = rows[i];
    const int padded = (d.nnz + 3) & ~3;
    for (int e = lane; e < padded; e += 32) {
      if (e < d.nnz) {
        const int32_t j = src_ci[s.start + e];
        const double v = src_cv[s.start + e];
        ci[d.start + e] = j;
        cv[d.start + e] = x_scale ? v / x_scale[j] : v;
      } else {
        ci[d.start + e] = 0;
        cv[d.start + e] = 0.0;
      }
    }
  }
}

cudaError_t subset_local(const int32_t* subset_dev, int64_t n_sub, int64_t n_src, int32_t* local, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(local, 0xff, sizeof(int32_t) * n_src, st);
  if (e != cudaSuccess) return e;
  subset_local_kernel<<<std::min(blocks_for(n_sub), 148 * 8), kT, 0, st>>>(subset_dev, n_sub, local);
  return cudaGetLastError();
}
cudaError_t subset_counts(const RowInfo* src, const int32_t* subset_dev, int64_t n_sub, int32_t* counts, int64_t* block_tot,
                          int64_t* totals_dev, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(totals_dev, 0, 2 * sizeof(int64_t), st);
  if (e != cudaSuccess) return e;
  subset_counts_kernel<<<std::min(blocks_for(n_sub), 148 * 8), kT, 0, st>>>(src, subset_dev, n_sub, counts);
  const int nb = blocks_for(n_sub, kScanItems);
  scan_block_totals_kernel<<<nb, kT, 0, st>>>(counts, n_sub, block_tot, totals_dev + 1);
  scan_of_totals_kernel<<<1, 32, 0, st>>>(block_tot, nb, totals_dev);
  return cudaGetLastError();
}
cudaError_t gather_rows(const RowInfo* src_rows, const int32_t* src_ci, const double* src_cv, const int32_t* subset_dev,
                        int64_t n_sub, const int32_t* counts, const int64_t* block_off, RowInfo* rows, int32_t* cursor,
                        int32_t* ci, double* cv, const double* x_scale, cudaStream_t st) {
  scan_rows_kernel<<<blocks_for(n_sub, kScanItems), 32, 0, st>>>(counts, n_sub, block_off, rows, cursor);
  gather_rows_kernel<<<std::min(blocks_for(n_sub * 32), 148 * 16), kT, 0, st>>>(src_rows, src_ci, src_cv, subset_dev, n_sub, rows, ci,
                                                                           cv, x_scale);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------ sparse column statistics
// One thread per column walks the column's entries (ascending rows, the order of the reference's per-column running
// sums) restricted to the subset: Mean, StandardDeviation (src/math.h:66-79, 89-112), then c = center / scale.
__global__ void sparse_col_stats_kernel(const int32_t* __restrict__ csc_i, const int32_t* __restrict__ csc_p,
                                        const double* __restrict__ csc_x, int32_t p, const int32_t* __restrict__ local,
                                        int64_t n_sub, double* __restrict__ x_center, double* __restrict__ x_scale,
                                        double* __restrict__ c) {
  const double nd = static_cast<double>(n_sub);
  for (int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; j < p; j += int64_t(gridDim.x) * blockDim.x) {
    const int64_t lo = csc_p[j], hi = csc_p[j + 1];
    double total = 0.0;
    int64_t count = 0;
    for (int64_t e = lo; e < hi; ++e) {
      if (local && local[csc_i[e]] < 0) continue;
      total += csc_x[e];
      ++count;
    }
    const double center = total / nd;
    double var = 0.0;
    for (int64_t e = lo; e < hi; ++e) {
      if (local && local[csc_i[e]] < 0) continue;
      const double dev = csc_x[e] - center;
      var += (dev * dev) / nd;
    }
    const int64_t zeros = n_sub - count;
    var += static_cast<double>(zeros) * center * center / nd;
    const double scale = (var == 0.0) ? 1.0 : sqrt(var);
    x_center[j] = center;
    x_scale[j] = scale;
    c[j] = center / scale;
  }
}

// (X^T ymap)[col][j]: one thread per (column, response); entries in ascending row order, v / x_scale_j when standardised
__global__ void sparse_xt_times_kernel(const int32_t* __restrict__ csc_i, const int32_t* __restrict__ csc_p,
                                       const double* __restrict__ csc_x, int32_t p, const int32_t* __restrict__ local,
                                       int64_t n_sub, const double* __restrict__ x_scale, const double* __restrict__ ymap,
                                       int m, double* __restrict__ out) {
  const int64_t total = int64_t(p) * m;
  for (int64_t q = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; q < total; q += int64_t(gridDim.x) * blockDim.x) {
    const int64_t j = q % p;
    const int col = static_cast<int>(q / p);
    const double* __restrict__ yc = ymap + int64_t(col) * n_sub;
    const double sc = x_scale ? x_scale[j] : 1.0;
    double acc = 0.0;
    for (int64_t e = csc_p[j]; e < csc_p[j + 1]; ++e) {
      const int32_t src = csc_i[e];
      const int32_t i = local ? local[src] : src;
      if (i < 0) continue;
      const double v = x_scale ? csc_x[e] / sc : csc_x[e];
      acc += v * yc[i];
    }
    out[int64_t(col) * p + j] = acc;
  }
}

// ColNormsMax (src/utils.h:60-85): one thread per row, its squared norm as one running sum; with virtual centring the
// sum runs over ALL features (x_sj - c_j)^2, zeros included. The maximum is order-free: integer max of the bit
// patterns of non-negative doubles.
__global__ void sparse_norm_max_kernel(const RowInfo* __restrict__ rows, const int32_t* __restrict__ ci,
                                       const double* __restrict__ cv, int64_t n, int32_t p, const double* __restrict__ c,
                                       unsigned long long* __restrict__ out_bits) {
  double best = 0.0;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const RowInfo ri = rows[i];
    double sq = 0.0;
    if (c) {
      int32_t e = 0;
      for (int32_t j = 0; j < p; ++j) {
        double v = 0.0;
        if (e < ri.nnz && ci[ri.start + e] == j) v = cv[ri.start + e++];
        const double dev = v - c[j];
        sq += dev * dev;
      }
    } else {
      for (int32_t e = 0; e < ri.nnz; ++e) {
        const double v = cv[ri.start + e];
        sq += v * v;
      }
    }
    best = fmax(best, sq);
  }
  best = warp_max(best);
  if ((threadIdx.x & 31) == 0) atomicMax(out_bits, static_cast<unsigned long long>(__double_as_longlong(best)));
}

cudaError_t sparse_col_stats(const int32_t* csc_i, const int32_t* csc_p, const double* csc_x, int32_t p, const int32_t* local,
                             int64_t n_sub, double* x_center, double* x_scale, double* c, cudaStream_t st) {
  sparse_col_stats_kernel<<<blocks_for(p, 128), 128, 0, st>>>(csc_i, csc_p, csc_x, p, local, n_sub, x_center, x_scale, c);
  return cudaGetLastError();
}
cudaError_t sparse_xt_times(const int32_t* csc_i, const int32_t* csc_p, const double* csc_x, int32_t p, const int32_t* local,
                            int64_t n_sub, const double* x_scale, const double* ymap, int m, double* out, cudaStream_t st) {
  sparse_xt_times_kernel<<<blocks_for(int64_t(p) * m, 128), 128, 0, st>>>(csc_i, csc_p, csc_x, p, local, n_sub, x_scale, ymap, m, out);
  return cudaGetLastError();
}
cudaError_t sparse_norm_max(const RowInfo* rows, const int32_t* ci, const double* cv, int64_t n, int32_t p, const double* c,
                            unsigned long long* out_bits, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(out_bits, 0, sizeof(unsigned long long), st);
  if (e != cudaSuccess) return e;
  sparse_norm_max_kernel<<<std::min(blocks_for(n, 128), 148 * 16), 128, 0, st>>>(rows, ci, cv, n, p, c, out_bits);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------ dense
// x: the caller's n_src x p column-major matrix on the device. One thread per column: Mean / StandardDeviation over
// the subset rows in order (src/math.h:66-79, 114-130).
__global__ void dense_col_stats_kernel(const double* __restrict__ x, int64_t n_src, int32_t p, const int32_t* __restrict__ subset,
                                       int64_t n_sub, double* __restrict__ x_center, double* __restrict__ x_scale) {
  const double nd = static_cast<double>(n_sub);
  for (int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; j < p; j += int64_t(gridDim.x) * blockDim.x) {
    const double* __restrict__ col = x + j * n_src;
    double total = 0.0;
    for (int64_t i = 0; i < n_sub; ++i) total += col[subset ? subset[i] : i];
    const double mean = total / nd;
    double ss = 0.0;
    for (int64_t i = 0; i < n_sub; ++i) {
      const double dev = col[subset ? subset[i] : i] - mean;
      ss += dev * dev;
    }
    const double var = ss / nd;
    x_center[j] = mean;
    x_scale[j] = (var == 0.0) ? 1.0 : sqrt(var);
  }
}
// column-major caller matrix -> [n_sub][ld] sample-major, centred and scaled when standardising (src/math.h:139-150),
// through a 32 x 32 shared-memory tile so that both sides are coalesced
__global__ void dense_transpose_kernel(const double* __restrict__ x, int64_t n_src, int32_t p, const int32_t* __restrict__ subset,
                                       int64_t n_sub, int32_t ld, const double* __restrict__ x_center,
                                       const double* __restrict__ x_scale, double* __restrict__ xd) {
  __shared__ double tile[32][33];
  const int64_t i0 = int64_t(blockIdx.x) * 32;
  const int32_t j0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8 threads
  for (int r = ty; r < 32; r += 8) {
    const int32_t j = j0 + r;
    const int64_t i = i0 + tx;
    double v = 0.0;
    if (j < p && i < n_sub) {
      v = x[int64_t(j) * n_src + (subset ? subset[i] : i)];
      if (x_center) v = (v - x_center[j]) / x_scale[j];
    }
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int64_t i = i0 + r;
    const int32_t j = j0 + tx;
    if (i < n_sub && j < ld) xd[i * ld + j] = (j < p) ? tile[tx][r] : 0.0;
  }
}
__global__ void dense_norm_max_kernel(const double* __restrict__ xd, int64_t n, int32_t p, int32_t ld,
                                      unsigned long long* __restrict__ out_bits) {
  double best = 0.0;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const double* __restrict__ row = xd + i * ld;
    double sq = 0.0;
    for (int32_t j = 0; j < p; ++j) sq += row[j] * row[j];
    best = fmax(best, sq);
  }
  best = warp_max(best);
  if ((threadIdx.x & 31) == 0) atomicMax(out_bits, static_cast<unsigned long long>(__double_as_longlong(best)));
}
// out[col][j] = sum_i xd[i][j] * ymap[col][i], i ascending: one thread per (j, col), coalesced across j
__global__ void dense_xt_times_kernel(const double* __restrict__ xd, int64_t n, int32_t p, int32_t ld,
                                      const double* __restrict__ ymap, int m, double* __restrict__ out) {
  const int64_t total = int64_t(p) * m;
  for (int64_t q = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; q < total; q += int64_t(gridDim.x) * blockDim.x) {
    const int64_t j = q % p;
    const int col = static_cast<int>(q / p);
    const double* __restrict__ yc = ymap + int64_t(col) * n;
    double acc = 0.0;
    for (int64_t i = 0; i < n; ++i) acc += xd[i * ld + j] * yc[i];
    out[int64_t(col) * p + j] = acc;
  }
}

cudaError_t dense_design(const double* x, int64_t n_src, int32_t p, const int32_t* subset_dev, int64_t n_sub, int32_t ld,
                         bool standardize, double* x_center, double* x_scale, double* xd, unsigned long long* norm_bits,
                         cudaStream_t st) {
  if (standardize) dense_col_stats_kernel<<<blocks_for(p, 64), 64, 0, st>>>(x, n_src, p, subset_dev, n_sub, x_center, x_scale);
  dim3 grid(static_cast<unsigned>((n_sub + 31) / 32), static_cast<unsigned>((ld + 31) / 32));
  dense_transpose_kernel<<<grid, 256, 0, st>>>(x, n_src, p, subset_dev, n_sub, ld, standardize ? x_center : nullptr, x_scale, xd);
  cudaError_t e = cudaMemsetAsync(norm_bits, 0, sizeof(unsigned long long), st);
  if (e != cudaSuccess) return e;
  dense_norm_max_kernel<<<std::min(blocks_for(n_sub, 128), 148 * 16), 128, 0, st>>>(xd, n_sub, p, ld, norm_bits);
  return cudaGetLastError();
}
cudaError_t dense_xt_times(const double* xd, int64_t n, int32_t p, int32_t ld, const double* ymap, int m, double* out,
                           cudaStream_t st) {
  dense_xt_times_kernel<<<blocks_for(int64_t(p) * m, 64), 64, 0, st>>>(xd, n, p, ld, ymap, m, out);
  return cudaGetLastError();
}

}  // namespace sgd
