// microbench2.cu — throughput of the scattered coefficient-state gather/scatter of one sparse row (100 features)
// from ONE SM, for the candidate state layouts: three arrays in L2, one packed 32-byte record in L2 (256-bit
// LDG/STG), packed records in distributed shared memory of a 16-CTA cluster.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

struct __align__(32) St { double w, g; unsigned lag, pad; unsigned long long pad2; };

__device__ __forceinline__ unsigned rnd(unsigned& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

// mode 0: separate arrays; mode 1: packed 256-bit
__global__ void k_l2(double* W, double* G, unsigned* L, St* st, int p, int rows, int mode, long long* cyc, double* sink) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  unsigned seed = 1234567u + threadIdx.x * 7919u;
  double acc = 0.0;
  __syncthreads();
  long long t0 = clock64();
  for (int r = warp; r < rows; r += nw) {
    int j[4]; double w[4], g[4]; unsigned l[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) j[c] = (c * 32 + lane < 100) ? int(rnd(seed) % unsigned(p)) : -1;
    if (mode == 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) if (j[c] >= 0) { w[c] = W[j[c]]; g[c] = G[j[c]]; l[c] = L[j[c]]; }
#pragma unroll
      for (int c = 0; c < 4; ++c) if (j[c] >= 0) { acc += w[c] * g[c]; W[j[c]] = w[c] + 1e-9; G[j[c]] = g[c] + 1e-9; L[j[c]] = l[c] + 1; }
    } else {
      unsigned long long a[4], b[4], cc[4], d[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) if (j[c] >= 0)
        asm volatile("ld.global.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(a[c]), "=l"(b[c]), "=l"(cc[c]), "=l"(d[c]) : "l"(st + j[c]));
#pragma unroll
      for (int c = 0; c < 4; ++c) if (j[c] >= 0) {
        double ww = __longlong_as_double(a[c]), gg = __longlong_as_double(b[c]);
        acc += ww * gg;
        asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" :: "l"(st + j[c]), "l"(__double_as_longlong(ww + 1e-9)),
                     "l"(__double_as_longlong(gg + 1e-9)), "l"(cc[c] + 1), "l"(d[c]) : "memory");
      }
    }
    __syncwarp();
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  sink[threadIdx.x] = acc;
}

// packed records spread over the shared memory of a cluster; only CTA 0 computes
__global__ void __cluster_dims__(16, 1, 1) k_dsmem(int per_cta, int rows, long long* cyc, double* sink) {
  extern __shared__ __align__(32) unsigned char smraw[];
  St* mine = reinterpret_cast<St*>(smraw);
  cg::cluster_group cl = cg::this_cluster();
  for (int i = threadIdx.x; i < per_cta; i += blockDim.x) { mine[i].w = 1.0; mine[i].g = 0.5; mine[i].lag = 0; }
  cl.sync();
  if (cl.block_rank() == 0) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    unsigned seed = 1234567u + threadIdx.x * 7919u;
    double acc = 0.0;
    const int p = per_cta * 16;
    long long t0 = clock64();
    for (int r = warp; r < rows; r += nw) {
      St* ptr[4]; double w[4], g[4]; unsigned l[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        ptr[c] = nullptr;
        if (c * 32 + lane < 100) { int j = int(rnd(seed) % unsigned(p)); ptr[c] = cl.map_shared_rank(mine + (j % per_cta), j / per_cta); }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) if (ptr[c]) { double2 v = *reinterpret_cast<double2*>(ptr[c]); w[c] = v.x; g[c] = v.y; l[c] = ptr[c]->lag; }
#pragma unroll
      for (int c = 0; c < 4; ++c) if (ptr[c]) { acc += w[c] * g[c]; *reinterpret_cast<double2*>(ptr[c]) = make_double2(w[c] + 1e-9, g[c] + 1e-9); ptr[c]->lag = l[c] + 1; }
      __syncwarp();
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    sink[threadIdx.x] = acc;
  }
  cl.sync();
}

int main() {
  const int p = 100000, rows = 40000;
  double *W, *G, *sink; unsigned* L; St* st; long long* cyc;
  cudaMalloc(&W, p * 8); cudaMalloc(&G, p * 8); cudaMalloc(&L, p * 4); cudaMalloc(&st, size_t(p) * 32);
  cudaMemset(W, 0, p * 8); cudaMemset(G, 0, p * 8); cudaMemset(L, 0, p * 4); cudaMemset(st, 0, size_t(p) * 32);
  cudaMalloc(&sink, 1024 * 8); cudaMalloc(&cyc, 64);
  long long h;
  for (int mode = 0; mode < 2; ++mode)
    for (int nw : {1, 4, 8, 16}) {
      for (int rep = 0; rep < 2; ++rep) k_l2<<<1, nw * 32>>>(W, G, L, st, p, rows, mode, cyc, sink);
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("L2 state %-22s warps=%2d  %8.1f cycles/row\n", mode == 0 ? "3 arrays (8+8+4 B)" : "packed 32 B (256-bit)", nw, double(h) / rows);
    }
  const int per_cta = 6272;   // 16 * 6272 = 100352 records of 32 B = 200704 B per CTA
  cudaFuncSetAttribute(k_dsmem, cudaFuncAttributeMaxDynamicSharedMemorySize, per_cta * 32);
  cudaFuncSetAttribute(k_dsmem, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int nw : {1, 4, 8, 16}) {
    for (int rep = 0; rep < 2; ++rep) k_dsmem<<<16, nw * 32, per_cta * 32>>>(per_cta, rows, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DSMEM state packed (16-CTA cluster)   warps=%2d  %8.1f cycles/row  (%s)\n", nw, double(h) / rows, cudaGetErrorString(e));
  }
  return 0;
}
