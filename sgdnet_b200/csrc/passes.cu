// passes.cu — the streaming (HBM-bound) passes around the SAGA loop.
//
//   lag_scaling_kernel     running geometric sum of src/saga-sparse.h:229-240 (one thread per fit: it is a serial
//                          floating-point recurrence that must be reproduced in order; rebuilt at every lambda)
//   loss_pass_kernel       per-lambda Deviance (src/utils.h:304-329) and the debug EpochLoss (src/utils.h:199-227):
//                          one warp per sample, X streamed once, block partial sums, fixed-order final sum
//   finish_lambda_kernel   dev.ratio (src/sgdnet.cpp:246-258), Rescale + archive (src/utils.h:352-378), and the
//                          per-fit state machine step to the next lambda
//   predict / score        X * [a0; beta] for all lambda at once (R/predict.sgdnet.R:377, 507-510) and the held-out
//                          deviance of R/score.R:55-178, lanes across lambda so the coefficient reads are coalesced
//
// Algorithmic HBM bytes per loss pass: sparse 12*nnz + 16*n (row info) + 8*n*K_y; dense 8*n*ld + 8*n*K_y.
#include "../../include/sgdnet_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace sgd {

constexpr int kPassThreads = 256;

// ------------------------------------------------------------------------------------------ lag scaling
__global__ void lag_scaling_kernel(FitDev* __restrict__ fit, const Progress* __restrict__ prog) {
  const Progress& pg = *prog;
  if (pg.status != kRunning || pg.it_outer != 0) return;
  const FitDev& f = *fit;
  if (!f.sparse || threadIdx.x != 0) return;
  const int li = pg.lambda_ind;
  const double r = 1.0 - f.alpha[li] * f.gamma[li];
  if (r == 1.0) return;   // table would hold exact integers; the solver uses (double)m directly
  double* __restrict__ ls = f.lag_scaling;
  ls[0] = 0.0;
  ls[1] = 1.0;
  double geo = 1.0, last = 1.0;
  const uint32_t n1 = static_cast<uint32_t>(f.n) + 1u;
  for (uint32_t i = 2; i < n1; ++i) {
    geo *= r;
    last = last + geo;
    ls[i] = last;
  }
}

cudaError_t launch_lag_scaling(FitDev* fit, Progress* prog, cudaStream_t st) {
  lag_scaling_kernel<<<1, 32, 0, st>>>(fit, prog);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------ loss pass
// mode 0: deviance for fits with status == kLambdaDone; mode 1: epoch loss (debug) for fits that just ran an epoch.
__device__ __forceinline__ double sample_loss_warp(const FitDev& f, int K, int Ky, double lp_lane, int64_t s, int lane) {
  // lanes 0..K-1 hold lp[k]; returns the loss in every lane
  if (K == 1) {
    const double y = f.yt[s];
    return loss_scalar(f.family, __shfl_sync(0xffffffffu, lp_lane, 0), y);
  }
  const bool valid = lane < K;
  if (f.family == kMultinomial) {
    const double lse = lse_warp(lp_lane, valid);
    const unsigned c = static_cast<unsigned>(f.yt[s] + 0.5);
    const double lpc = __shfl_sync(0xffffffffu, lp_lane, static_cast<int>(c));
    return lse - lpc;
  }
  double d = 0.0;
  if (valid) {
    d = lp_lane - f.yt[s * Ky + lane];
    d = d * d;
  }
  return 0.5 * warp_sum(d);
}

// Sparse, K == 1, no virtual centring (configs 2 and 5): the bandwidth-bound form. A warp takes tiles of 32
// consecutive rows: the 32 row descriptors and responses are one coalesced load each; the rows' index / value runs are
// streamed four rows at a time (up to sixteen independent coalesced loads and as many 32-wide gathers of W in flight
// per warp); a row's dot product keeps the association of the solver's (position e -> running sum e mod 32, then the
// xor butterfly); the 32 linear predictors end up one per lane, so exp / log of the loss run ONCE per row instead of
// 32 times, and the tile's 32 losses are added with one more butterfly. Partial sums: per warp in tile order, per
// block in warp order, then the blocks in order (finish_lambda_kernel) - fixed, hence reproducible.
__device__ __forceinline__ double loss_tiles_sparse_k1(const FitDev& f, int mode, int lane, int64_t warp_global,
                                                       int64_t warp_stride, const uint32_t* nz_smem, bool use_mask) {
  const int64_t n = f.n;
  const RowInfo* __restrict__ rows = f.rows;
  const int32_t* __restrict__ ci = f.ci;
  const double* __restrict__ cv = f.cv;
  const double* __restrict__ W = f.W;
  const double* __restrict__ yt = f.yt;
  const double b0 = f.b[0];
  const int family = f.family;
  // mode 0 (deviance): the bitmap of nonzero coefficients written by rescale_kernel just before this pass sits in
  // shared memory, and a lane gathers W[j] only when it is set (the gathers - one 32-byte sector per 8-byte weight,
  // through L1 from L2 - are what bounds this kernel; along a lasso path most weights are zero most of the time).
  // A zero weight contributes +0.0 * v exactly as before (v is finite), so the sums keep their bits.
  const uint32_t* __restrict__ nzm = use_mask ? nz_smem : nullptr;
  double acc = 0.0;
  const int64_t n_tiles = (n + 31) / 32;
  for (int64_t tile = warp_global; tile < n_tiles; tile += warp_stride) {
    const int64_t s0 = tile * 32;
    const int64_t s_mine = s0 + lane;
    const bool have = s_mine < n;
    RowInfo ri_mine{};
    double y_mine = 0.0;
    if (have) {
      ri_mine = rows[s_mine];
      y_mine = yt[s_mine];
    }
    double lp_mine = 0.0;
#pragma unroll 1
    for (int r0 = 0; r0 < 32; r0 += 4) {
      if (s0 + r0 >= n) break;
      int64_t st[4];
      int nz[4];
      double a[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        st[u] = __shfl_sync(0xffffffffu, ri_mine.start, r0 + u);
        nz[u] = __shfl_sync(0xffffffffu, ri_mine.nnz, r0 + u);      // 0 for rows past the end
        a[u] = 0.0;
      }
      // the first four chunks of the four rows: sixteen independent load pairs
      int32_t jj[4][4];
      double vv[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int e = c * 32 + lane;
          const bool ok = e < nz[u];
          jj[u][c] = ok ? ci[st[u] + e] : 0;
          vv[u][c] = ok ? cv[st[u] + e] : 0.0;
        }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int j = jj[u][c];
          const bool live = !use_mask || ((nzm[j >> 5] >> (j & 31)) & 1u);
          const double wj = live ? W[j] : 0.0;
          const double a2 = a[u] + vv[u][c] * wj;
          a[u] = (c * 32 + lane < nz[u]) ? a2 : a[u];
        }
      // rows longer than 128 nonzeros (rare): the rest of the row, same association
#pragma unroll
      for (int u = 0; u < 4; ++u)
        for (int e = 128 + lane; e < nz[u]; e += 32) a[u] += cv[st[u] + e] * W[ci[st[u] + e]];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int u = 0; u < 4; ++u) a[u] += __shfl_xor_sync(0xffffffffu, a[u], o);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) lp_mine = (lane == r0 + u) ? a[u] : lp_mine;
    }
    double loss = have ? loss_scalar(family, lp_mine + b0, y_mine) : 0.0;
    if (mode == 1) loss = loss / static_cast<double>(static_cast<uint32_t>(n));
    acc += warp_sum(loss);
  }
  return acc;
}

// ---- The same pass with the design streamed by bulk copies (configs 2 and 5 at size): one persistent CTA per SM, a
// producer warp and 16 consumer warps around a 4-stage shared-memory ring of 32-row tiles. The rows of a tile are
// consecutive in the padded CSR, so a tile's index run and value run are ONE contiguous range each: two 1-D bulk
// copies (cp.async.bulk -> UBLKCP, complete_tx on the stage's mbarrier) of about 13 KB + 26 KB bring them in, three
// tiles ahead of the warps that consume them - about 115 KB in flight per SM without a register or an L1 line spent
// on it. The producer derives each row's offset inside the tile from the 32 row descriptors it reads (a warp scan; the
// descriptors themselves are fetched two tiles ahead, so a freed stage is refilled at once) and leaves it, with the
// row's response, in the stage. A consumer warp takes two rows of every tile, interleaved:
// position e of a row goes to the running sum of lane e mod 32, then the xor butterfly - the solver's association
// (sgdnet_arith.h item 2). Linear predictors and responses collect one per lane over 16 tiles, so exp / log of the
// loss run once per 32 rows per warp. Partial sums: per lane over its rows in order, butterfly per warp, warps in order
// per block, blocks in order (finish_lambda_kernel) - fixed, hence reproducible.
// A tile that is not one contiguous range or has more than kTileCap padded entries (very long rows) is read straight
// from global memory by the consumers (kind 1).
constexpr int kTileRows = 32;
constexpr int kTileCap = 3584;          // padded entries per staged tile
constexpr int kTileStages = 4;
constexpr int kTileWarps = 16;          // consumer warps; warp kTileWarps is the producer
constexpr int kTileThreads = (kTileWarps + 1) * 32;

struct __align__(16) TileRow {
  int64_t start;        // first entry of the row in ci / cv
  int32_t off;          // first entry of the row inside the staged tile
  int32_t nnz;
};
struct __align__(128) TileStage {
  double cv[kTileCap];
  int32_t ci[kTileCap];
  TileRow row[kTileRows];
  double y[kTileRows];
  int32_t kind;         // 0: staged, 1: read the rows from global memory
  int32_t pad_[31];
};
struct __align__(128) TileSmem {
  TileStage st[kTileStages];
  uint64_t full[kTileStages];
  uint64_t empty[kTileStages];
  double red[kTileWarps];
};

size_t loss_tiles_fixed_smem() { return sizeof(TileSmem); }

__global__ void __launch_bounds__(kTileThreads, 1)
loss_pass_tiles_kernel(FitDev* __restrict__ fit, const Progress* __restrict__ prog, int mode, int mask_words) {
  extern __shared__ __align__(128) unsigned char tile_smem_raw[];
  const Progress& pg = *prog;
  const FitDev& f = *fit;
  if (mode == 0 ? (pg.status != kLambdaDone) : !f.debug) return;
  TileSmem& sm = *reinterpret_cast<TileSmem*>(tile_smem_raw);
  uint32_t* const nz_smem = reinterpret_cast<uint32_t*>(tile_smem_raw + sizeof(TileSmem));

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n = f.n;
  const int p = f.p;
  const int words = (p + 31) / 32;
  const bool use_mask = mode == 0 && f.nz_mask != nullptr && mask_words >= words;
  if (tid == 0) {
    for (int i = 0; i < kTileStages; ++i) {
      mbar_init(&sm.full[i], 1);
      mbar_init(&sm.empty[i], kTileWarps);
    }
    fence_barrier_init();
  }
  if (use_mask)
    for (int i = tid; i < words; i += kTileThreads) nz_smem[i] = f.nz_mask[i];
  __syncthreads();

  const int64_t n_tiles = (n + kTileRows - 1) / kTileRows;
  const RowInfo* __restrict__ rows = f.rows;
  const int32_t* __restrict__ ci = f.ci;
  const double* __restrict__ cv = f.cv;
  double acc = 0.0;

  if (warp == kTileWarps) {
    // ------------------------------------------------------------ producer
    int it = 0;
    auto fetch = [&](int64_t tile, RowInfo& ri, double& y) {
      const int64_t s = tile * kTileRows + lane;
      ri = RowInfo{};
      y = 0.0;
      if (tile < n_tiles && s < n) {
        ri = rows[s];
        y = f.yt[s];
      }
    };
    RowInfo ri_a, ri_b;
    double y_a, y_b;
    fetch(blockIdx.x, ri_a, y_a);
    fetch(int64_t(blockIdx.x) + gridDim.x, ri_b, y_b);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int stage = it % kTileStages;
      const RowInfo ri = ri_a;
      const double y = y_a;
      ri_a = ri_b;
      y_a = y_b;
      fetch(tile + 2 * int64_t(gridDim.x), ri_b, y_b);
      if (it >= kTileStages) mbar_wait(&sm.empty[stage], static_cast<uint32_t>((it / kTileStages - 1) & 1));
      const bool have = tile * kTileRows + lane < n;
      const int padded = (ri.nnz + 3) & ~3;
      int incl = padded;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      const int64_t start0 = __shfl_sync(0xffffffffu, ri.start, 0);
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      const bool in_place = !have || ri.nnz == 0 || ri.start == start0 + (incl - padded);
      const bool staged = __all_sync(0xffffffffu, in_place) && total <= kTileCap;
      TileStage& st = sm.st[stage];
      st.row[lane] = TileRow{ri.start, incl - padded, have ? ri.nnz : 0};
      st.y[lane] = y;
      if (lane == 0) st.kind = staged ? 0 : 1;
      __syncwarp();
      if (lane == 0) {
        if (staged && total > 0) {
          mbar_expect_tx(&sm.full[stage], static_cast<uint32_t>(total) * 12u);
          bulk_g2s(st.ci, ci + start0, static_cast<uint32_t>(total) * 4u, &sm.full[stage]);
          bulk_g2s(st.cv, cv + start0, static_cast<uint32_t>(total) * 8u, &sm.full[stage]);
        } else {
          mbar_arrive(&sm.full[stage]);
        }
      }
    }
  } else {
    // ------------------------------------------------------------ consumers: rows 2 * warp, 2 * warp + 1 of every tile
    const double* __restrict__ W = f.W;
    const double b0 = f.b[0];
    const int family = f.family;
    const double nd = static_cast<double>(static_cast<uint32_t>(n));
    double lp_mine = 0.0, y_mine = 0.0;
    bool row_mine = false;
    int it = 0, held = 0;
    auto flush = [&]() {
      double loss = row_mine ? loss_scalar(family, lp_mine + b0, y_mine) : 0.0;
      if (mode == 1) loss = loss / nd;
      acc += loss;
      row_mine = false;
      held = 0;
    };
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int stage = it % kTileStages;
      mbar_wait(&sm.full[stage], static_cast<uint32_t>((it / kTileStages) & 1));
      const TileStage& st = sm.st[stage];
      const TileRow ra = st.row[2 * warp], rb = st.row[2 * warp + 1];
      double a0 = 0.0, a1 = 0.0;
      const int nmax = ra.nnz > rb.nnz ? ra.nnz : rb.nnz;
      const bool staged = st.kind == 0;
      // warp-uniform trip count (four 32-entry chunks of both rows per pass, predicated per lane): eight independent
      // index -> bitmap -> gather chains in flight
      auto two_rows = [&](const int32_t* __restrict__ cia, const int32_t* __restrict__ cib, const double* __restrict__ cva,
                          const double* __restrict__ cvb) {
#pragma unroll 1
        for (int e0 = 0; e0 < nmax; e0 += 128) {
          int ja[4], jb[4];
          double va[4], vb[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int e = e0 + 32 * u + lane;
            const bool oka = e < ra.nnz, okb = e < rb.nnz;
            ja[u] = oka ? cia[e] : -1;
            jb[u] = okb ? cib[e] : -1;
            va[u] = oka ? cva[e] : 0.0;
            vb[u] = okb ? cvb[e] : 0.0;
          }
          double wa[4], wb[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const bool la = ja[u] >= 0 && (!use_mask || ((nz_smem[ja[u] >> 5] >> (ja[u] & 31)) & 1u));
            const bool lb = jb[u] >= 0 && (!use_mask || ((nz_smem[jb[u] >> 5] >> (jb[u] & 31)) & 1u));
            wa[u] = la ? __ldg(W + ja[u]) : 0.0;
            wb[u] = lb ? __ldg(W + jb[u]) : 0.0;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const double a0n = a0 + va[u] * wa[u], a1n = a1 + vb[u] * wb[u];
            a0 = ja[u] >= 0 ? a0n : a0;
            a1 = jb[u] >= 0 ? a1n : a1;
          }
        }
      };
      if (staged) two_rows(st.ci + ra.off, st.ci + rb.off, st.cv + ra.off, st.cv + rb.off);     // shared memory (LDS)
      else two_rows(ci + ra.start, ci + rb.start, cv + ra.start, cv + rb.start);                 // global memory
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
      }
      // lane 2 * held + r keeps row r of this tile
      const int64_t s_mine = tile * kTileRows + 2 * warp + (lane & 1);
      if ((lane >> 1) == held) {
        lp_mine = (lane & 1) ? a1 : a0;
        y_mine = st.y[2 * warp + (lane & 1)];
        row_mine = s_mine < n;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.empty[stage]);
      if (++held == 16) flush();
    }
    if (held > 0) flush();
    acc = warp_sum(acc);
    if (lane == 0) sm.red[warp] = acc;
  }
  __syncthreads();
  if (tid == 0) {
    double a = 0.0;
    for (int w = 0; w < kTileWarps; ++w) a += sm.red[w];
    f.partials[blockIdx.x] = a;
  }
}

__global__ void __launch_bounds__(kPassThreads)
loss_pass_kernel(FitDev* __restrict__ fit, const Progress* __restrict__ prog, int mode, int mask_words) {
  __shared__ double wc_s[32];
  __shared__ double red_s[kPassThreads / 32];
  __shared__ double red2[2 * (kPassThreads / 32) * 32];
  const Progress& pg = *prog;
  const FitDev& f = *fit;
  if (mode == 0 ? (pg.status != kLambdaDone) : !f.debug) return;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int K = f.K, Ky = f.Ky, p = f.p, ld = f.ld;
  const int64_t n = f.n;
  const double* __restrict__ W = f.W;
  const bool stdz = f.sparse && f.standardize;
  const int64_t warp_global = int64_t(blockIdx.x) * nwarps + warp;
  const int64_t warp_stride = int64_t(gridDim.x) * nwarps;
  double acc = 0.0;

  if (f.sparse && K == 1 && !stdz) {
    extern __shared__ uint32_t nz_smem[];
    const int words = (p + 31) / 32;
    const bool use_mask = mode == 0 && f.nz_mask != nullptr && mask_words >= words;
    if (use_mask) {
      for (int i = tid; i < words; i += blockDim.x) nz_smem[i] = f.nz_mask[i];
      __syncthreads();
    }
    acc = loss_tiles_sparse_k1(f, mode, lane, warp_global, warp_stride, nz_smem, use_mask);
  } else {
    // W . c per class (virtual centring), once per block
    if (stdz) {
      for (int k = 0; k < K; ++k) {
        double a = 0.0;
        for (int j = tid; j < p; j += blockDim.x) a += W[size_t(k) * p + j] * f.c[j];
        a = warp_sum(a);
        if (lane == 0) red2[warp * 32 + k] = a;
      }
      __syncthreads();
      if (tid < K) {
        double a = 0.0;
        for (int w = 0; w < nwarps; ++w) a += red2[w * 32 + tid];
        wc_s[tid] = a;
      }
      __syncthreads();
    }
    const double bk = (lane < K) ? f.b[lane] : 0.0;
    for (int64_t s = warp_global; s < n; s += warp_stride) {
      double lp_lane = 0.0;
      if (f.sparse) {
        const RowInfo ri = f.rows[s];
        const int32_t* __restrict__ ci = f.ci + ri.start;
        const double* __restrict__ cv = f.cv + ri.start;
        for (int k = 0; k < K; ++k) {
          double a = 0.0;
          for (int e = lane; e < ri.nnz; e += 32) a += cv[e] * W[size_t(k) * p + ci[e]];
          a = warp_sum(a);
          if (lane == k) lp_lane = a;
        }
      } else {
        // the row is read once per group of four classes (config 4: once; config 3: three times, from L1 after the first)
        const double* __restrict__ xr = f.xd + size_t(s) * ld;
        for (int k0 = 0; k0 < K; k0 += 4) {
          double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
          const double* __restrict__ W0 = W + size_t(k0) * p;
          const double* __restrict__ W1 = W + size_t(k0 + 1 < K ? k0 + 1 : k0) * p;
          const double* __restrict__ W2 = W + size_t(k0 + 2 < K ? k0 + 2 : k0) * p;
          const double* __restrict__ W3 = W + size_t(k0 + 3 < K ? k0 + 3 : k0) * p;
          for (int j = lane; j < p; j += 32) {
            const double xj = xr[j];
            a0 += W0[j] * xj;
            a1 += W1[j] * xj;
            a2 += W2[j] * xj;
            a3 += W3[j] * xj;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            a0 += __shfl_xor_sync(0xffffffffu, a0, o);
            a1 += __shfl_xor_sync(0xffffffffu, a1, o);
            a2 += __shfl_xor_sync(0xffffffffu, a2, o);
            a3 += __shfl_xor_sync(0xffffffffu, a3, o);
          }
          if (lane == k0) lp_lane = a0;
          if (lane == k0 + 1 && k0 + 1 < K) lp_lane = a1;
          if (lane == k0 + 2 && k0 + 2 < K) lp_lane = a2;
          if (lane == k0 + 3 && k0 + 3 < K) lp_lane = a3;
        }
      }
      if (lane < K) {
        lp_lane += bk;
        if (stdz) lp_lane -= wc_s[lane];
      }
      double loss = sample_loss_warp(f, K, Ky, lp_lane, s, lane);
      if (mode == 1) loss = loss / static_cast<double>(static_cast<uint32_t>(n));
      acc += loss;
    }
  }
  if (lane == 0) red_s[warp] = acc;
  __syncthreads();
  if (tid == 0) {
    double a = 0.0;
    for (int w = 0; w < nwarps; ++w) a += red_s[w];
    f.partials[blockIdx.x] = a;
  }
}

// ------------------------------------------------------------------------------------------ finish lambda
// Rescale (src/utils.h:363-377): beta_j = w_j * y_scale / x_scale_j into the lambda's archive slot; each block also
// leaves its share of sum_j x_center_j * beta_j (thread-strided running sums, butterfly, warps in order) for the
// intercept, and - sparse K == 1 - the bitmap of nonzero coefficients the deviance pass uses to skip gathers of zeros.
// Runs BEFORE the deviance pass of the same lambda.
__global__ void __launch_bounds__(kPassThreads)
rescale_kernel(FitDev* __restrict__ fit, const Progress* __restrict__ prog) {
  __shared__ double red[(kPassThreads / 32) * 32];
  const Progress& pg = *prog;
  const FitDev& f = *fit;
  if (pg.status != kLambdaDone) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int K = f.K, p = f.p;
  const int li = pg.lambda_ind;
  double* __restrict__ ba = f.beta_arch + size_t(li) * p * K;
  // features in chunks of 32 * gridDim.x ... : block b owns j = b*T + tid, stride gridDim.x*T
  const int stride = gridDim.x * blockDim.x;
  for (int k = 0; k < K; ++k) {
    double a = 0.0;
    const double ys = f.y_scale[k];
    for (int j = blockIdx.x * blockDim.x + tid; j < p; j += stride) {
      const double v = f.W[size_t(k) * p + j] * (ys / f.x_scale[j]);
      ba[size_t(j) * K + k] = v;
      a += f.x_center[j] * v;
    }
    a = warp_sum(a);
    if (lane == 0) red[warp * 32 + k] = a;
  }
  __syncthreads();
  if (tid < K) {
    double a = 0.0;
    for (int w = 0; w < nwarps; ++w) a += red[w * 32 + tid];
    f.xb_partials[size_t(blockIdx.x) * K + tid] = a;
  }
  if (f.nz_mask != nullptr) {
    // one warp-wide ballot per 32 consecutive features
    const int words = (p + 31) / 32;
    for (int wd = blockIdx.x * nwarps + warp; wd < words; wd += gridDim.x * nwarps) {
      const int j = wd * 32 + lane;
      const uint32_t m = __ballot_sync(0xffffffffu, j < p && f.W[j] != 0.0);
      if (lane == 0) f.nz_mask[wd] = m;
    }
  }
}

__global__ void __launch_bounds__(kPassThreads)
finish_lambda_kernel(FitDev* __restrict__ fit, Progress* __restrict__ prog, int n_partials, uint32_t round_id) {
  Progress& pg = *prog;
  const FitDev& f = *fit;
  if (pg.status != kLambdaDone) {
    if (threadIdx.x == 0) publish_progress(f.mirror, pg, round_id);
    return;
  }
  const int tid = threadIdx.x;
  const int K = f.K;
  const int li = pg.lambda_ind;
  if (tid < K) {
    double a = 0.0;
    for (int b = 0; b < kRescaleBlocks; ++b) a += f.xb_partials[size_t(b) * K + tid];
    double b0 = f.b[tid];
    if (f.fit_intercept) b0 = b0 * f.y_scale[tid] + f.y_center[tid] - a;
    f.a0_arch[size_t(li) * K + tid] = b0;
  }
  __syncthreads();
  if (tid == 0) {
    double dev = 0.0;
    for (int i = 0; i < n_partials; ++i) dev += f.partials[i];
    dev = 2.0 * dev;
    f.dev_ratio[li] = 1.0 - dev / f.null_deviance_scaled;
    pg.lambda_ind = li + 1;
    pg.it_outer = 0;
    pg.status = (li + 1 >= f.n_lambda) ? kFitDone : kRunning;
    publish_progress(f.mirror, pg, round_id);
  }
}

// `tiles` > 0: the fit is sparse, K == 1, without virtual centring - the bulk-copy form on that many CTAs (<= blocks)
static cudaError_t launch_loss(FitDev* fit, Progress* prog, int blocks, int tiles, int mode, int mask_words, cudaStream_t st) {
  if (tiles > 0) {
    const size_t smem = loss_tiles_fixed_smem() + size_t(mask_words) * 4;
    cudaError_t e = cudaFuncSetAttribute(loss_pass_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    loss_pass_tiles_kernel<<<tiles, kTileThreads, smem, st>>>(fit, prog, mode, mask_words);
  } else {
    loss_pass_kernel<<<blocks, kPassThreads, size_t(mask_words) * 4, st>>>(fit, prog, mode, mask_words);
  }
  return cudaGetLastError();
}

int loss_mask_words_max(bool tiles) {
  return tiles ? static_cast<int>((227 * 1024 - loss_tiles_fixed_smem()) / 4) : 40 * 1024 / 4;
}

cudaError_t launch_finish_lambda(FitDev* fit, Progress* prog, int blocks, int tiles, int mask_words, uint32_t round_id, cudaStream_t st) {
  rescale_kernel<<<kRescaleBlocks, kPassThreads, 0, st>>>(fit, prog);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  e = launch_loss(fit, prog, blocks, tiles, 0, mask_words, st);
  if (e != cudaSuccess) return e;
  finish_lambda_kernel<<<1, kPassThreads, 0, st>>>(fit, prog, tiles > 0 ? tiles : blocks, round_id);
  return cudaGetLastError();
}

// debug: per-epoch mean loss appended to f.losses[lambda_ind * max_iter + it_outer - 1]
__global__ void store_epoch_loss_kernel(FitDev* __restrict__ fit, const Progress* __restrict__ prog, int n_partials) {
  const Progress& pg = *prog;
  const FitDev& f = *fit;
  if (!f.debug || threadIdx.x != 0) return;
  double loss = 0.0;
  for (int i = 0; i < n_partials; ++i) loss += f.partials[i];
  // after an epoch the fit is either still running at lambda_ind, or it just finished lambda_ind (status LambdaDone)
  f.losses[size_t(pg.lambda_ind) * f.max_iter + (pg.it_outer - 1)] = loss;
}

cudaError_t launch_epoch_loss(FitDev* fit, Progress* prog, int blocks, int tiles, cudaStream_t st) {
  cudaError_t e = launch_loss(fit, prog, blocks, tiles, 1, 0, st);
  if (e != cudaSuccess) return e;
  store_epoch_loss_kernel<<<1, 32, 0, st>>>(fit, prog, tiles > 0 ? tiles : blocks);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------ predict / score
// Bt[j][l*K + k] = beta[l][j][k]: coefficient rows contiguous across lambda for coalesced SpMM reads.
__global__ void transpose_beta_kernel(const double* __restrict__ beta, double* __restrict__ bt, int L, int p, int K) {
  const size_t total = size_t(L) * p * K;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % K);
    const size_t jl = i / K;
    const int j = static_cast<int>(jl % p);
    const int l = static_cast<int>(jl / p);
    bt[size_t(j) * L * K + size_t(l) * K + k] = beta[i];
  }
}

// One warp per row; lanes across lambda (chunks of 32). For each class k the row is re-walked (it sits in L1).
// Measures (R/score.R:55-178; `measure` = SGDNET_MEASURE_*): per (row, lambda) contributions, summed per warp in row
// order, then warps and blocks in order; score_finalize_kernel divides (by n, or by K for mgaussian: R takes colMeans
// over the responses of colSums over the samples).
//   gaussian      deviance = mse: (eta - y)^2                    mae: |eta - y|
//   binomial      p = 1 / (1 + exp(-eta)), y2 = y, y1 = 1 - y2   (y <- diag(2)[as.numeric(y), ])
//                 deviance: -2 log of the clamped probability of the observed class
//                 mse: (p + y1 - 1)^2 + (p - y2)^2   mae: |p + y1 - 1| + |p - y2|   class: y1 (p > 0.5) + y2 (p <= 0.5)
//   multinomial   p_k = exp(eta_k) / sum_k exp(eta_k)            (R/predict.sgdnet.R:534-538: no max subtraction)
//                 deviance: -2 log clamp(p_true)   mse: sum_k (y_k - p_k)^2   mae: sum_k |y_k - p_k|
//                 class: 1 - y[remap(argmax_k p_k)], first maximum wins (softmax(), R/predict.sgdnet.R:104-128); remap is
//                 the rank of the predicted class among the classes predicted anywhere (as.numeric(as.factor(...)),
//                 R/score.R:153) - the identity unless some class is never predicted; `present` collects those classes
//   mgaussian     deviance = mse: sum_k (eta_k - y_k)^2          mae: sum_k |eta_k - y_k|
__global__ void __launch_bounds__(kPassThreads)
predict_score_kernel(PredictArgs a, const double* __restrict__ bt) {
  extern __shared__ double acc_s[];   // [nwarps][L]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int L = a.n_lambda, K = a.K, Ky = a.Ky, p = a.p;
  const int LK = L * K;
  const int measure = a.measure;
  double* my_acc = acc_s + size_t(warp) * L;
  for (int l = lane; l < L; l += 32) my_acc[l] = 0.0;
  __syncwarp();
  const double pmin = 1e-5, pmax = 1.0 - 1e-5;
  uint32_t seen = 0;                  // multinomial "class": classes this lane predicted for some (row, lambda)

  const int64_t warp_global = int64_t(blockIdx.x) * nwarps + warp;
  const int64_t warp_stride = int64_t(gridDim.x) * nwarps;
  for (int64_t i = warp_global; i < a.n; i += warp_stride) {
    const int64_t s = a.row_ids ? a.row_ids[i] : i;
    int nnz;
    const int32_t* ci = nullptr;
    const double* cv = nullptr;
    const double* xr = nullptr;
    if (a.sparse) {
      const RowInfo ri = a.rows[s];
      nnz = ri.nnz;
      ci = a.ci + ri.start;
      cv = a.cv + ri.start;
    } else {
      nnz = p;
      xr = a.xd + size_t(s) * a.ld;
    }
    for (int l0 = 0; l0 < L; l0 += 32) {
      const int l = l0 + lane;
      const bool valid = l < L;
      // eta_k of this row at lambda l
      auto eta = [&](int k) {
        double lp = a.a0[size_t(l) * K + k];
        const double* __restrict__ col = bt + size_t(l) * K + k;
        if (a.sparse) {
          for (int e = 0; e < nnz; ++e) lp += cv[e] * col[size_t(ci[e]) * LK];
        } else {
          for (int j = 0; j < p; ++j) lp += xr[j] * col[size_t(j) * LK];
        }
        return lp;
      };
      if (!valid) continue;
      double contrib = 0.0;
      if (K == 1) {
        const double lp0 = eta(0);
        if (a.link) a.link[size_t(l) * a.n + i] = lp0;
        if (a.y) {
          const double yv = a.y[s];
          if (a.family != kBinomial) {
            const double rr = lp0 - yv;
            contrib = (measure == SGDNET_MEASURE_MAE) ? fabs(rr) : rr * rr;
          } else {
            const double pr = 1.0 / (1.0 + sgd_exp(-lp0));
            const double y2 = (yv > 0.5) ? 1.0 : 0.0, y1 = 1.0 - y2;
            if (measure == SGDNET_MEASURE_DEVIANCE) {
              const double pc = fmin(fmax(pr, pmin), pmax);
              contrib = 2.0 * (0.0 - ((yv > 0.5) ? sgd_log(pc) : sgd_log(1.0 - pc)));
            } else if (measure == SGDNET_MEASURE_CLASS) {
              contrib = y1 * ((pr > 0.5) ? 1.0 : 0.0) + y2 * ((pr <= 0.5) ? 1.0 : 0.0);
            } else {
              const double u = (pr + y1) - 1.0, v = pr - y2;
              contrib = (measure == SGDNET_MEASURE_MAE) ? fabs(u) + fabs(v) : u * u + v * v;
            }
          }
        }
      } else if (a.family != kMultinomial || !a.y) {      // mgaussian, and plain prediction for any K > 1
        for (int k = 0; k < K; ++k) {
          const double lp = eta(k);
          if (a.link) a.link[(size_t(l) * K + k) * a.n + i] = lp;
          if (a.y) {
            const double rr = lp - a.y[s * Ky + k];
            contrib += (measure == SGDNET_MEASURE_MAE) ? fabs(rr) : rr * rr;
          }
        }
      } else {
        const unsigned cls = a.y ? static_cast<unsigned>(a.y[s * Ky] + 0.5) : 0u;
        double tot = 0.0, lp_true = 0.0;
        for (int k = 0; k < K; ++k) {
          const double lp = eta(k);
          if (a.link) a.link[(size_t(l) * K + k) * a.n + i] = lp;
          tot += sgd_exp(lp);
          if (static_cast<unsigned>(k) == cls) lp_true = lp;
        }
        if (a.y) {
          if (measure == SGDNET_MEASURE_DEVIANCE) {
            double pr = sgd_exp(lp_true) / tot;
            pr = fmin(fmax(pr, pmin), pmax);
            contrib = 2.0 * (0.0 - sgd_log(pr));
          } else {
            // a second walk over the classes, now that the normaliser is known
            double best = 0.0;
            int best_k = 0;
            for (int k = 0; k < K; ++k) {
              const double pk = sgd_exp(eta(k)) / tot;
              const double yk = (static_cast<unsigned>(k) == cls) ? 1.0 : 0.0;
              if (measure == SGDNET_MEASURE_MSE) contrib += (yk - pk) * (yk - pk);
              else if (measure == SGDNET_MEASURE_MAE) contrib += fabs(yk - pk);
              if (k == 0 || pk > best) {
                best = pk;
                best_k = k;
              }
            }
            if (measure == SGDNET_MEASURE_CLASS) {
              seen |= 1u << best_k;
              const unsigned code = a.remap ? static_cast<unsigned>(a.remap[best_k]) : static_cast<unsigned>(best_k);
              contrib = 1.0 - ((code == cls) ? 1.0 : 0.0);
            }
          }
        }
      }
      if (a.y) my_acc[l] += contrib;
    }
  }
  if (a.present != nullptr) {
    seen = __reduce_or_sync(0xffffffffu, seen);
    if (lane == 0 && seen != 0u) atomicOr(a.present, seen);
  }
  __syncthreads();
  if (a.partials) {
    for (int l = tid; l < L; l += blockDim.x) {
      double t = 0.0;
      for (int w = 0; w < nwarps; ++w) t += acc_s[size_t(w) * L + l];
      a.partials[size_t(blockIdx.x) * L + l] = t;
    }
  }
}

__global__ void score_finalize_kernel(PredictArgs a, int blocks) {
  for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < a.n_lambda; l += gridDim.x * blockDim.x) {
    double t = 0.0;
    for (int b = 0; b < blocks; ++b) t += a.partials[size_t(b) * a.n_lambda + l];
    a.score[l] = (a.family == kMGaussian) ? t / a.K : t / static_cast<double>(a.n);
  }
}

cudaError_t launch_predict_score(const PredictArgs& a, double* bt_scratch, int blocks, cudaStream_t st, bool transpose) {
  cudaError_t e = cudaSuccess;
  if (transpose) {
    transpose_beta_kernel<<<296, 256, 0, st>>>(a.beta, bt_scratch, a.n_lambda, a.p, a.K);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  const size_t smem = sizeof(double) * (kPassThreads / 32) * a.n_lambda;
  predict_score_kernel<<<blocks, kPassThreads, smem, st>>>(a, bt_scratch);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (a.score) {
    score_finalize_kernel<<<1, 128, 0, st>>>(a, blocks);
    e = cudaGetLastError();
  }
  return e;
}

}  // namespace sgd
