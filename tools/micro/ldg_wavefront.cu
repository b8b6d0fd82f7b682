// Microbenchmark behind the epilogue fragment layout: how fast can the 512 worker threads of one CTA per SM pull a
// 128-row x 1 KB tile (rows 1 KB apart, as the edge / node state rows are) out of L2 / HBM, depending on what one warp-level load
// instruction touches?
//   mode 0: LDG.128, a lane quad reads 64 B of one row, 8 rows per instruction    (16x256b.x2 fragment: 4 rows x 4 columns per thread)
//   mode 1: LDG.256, a lane quad reads 128 B of one row, 8 rows per instruction   (16x256b.x4 fragment: 2 rows x 8 columns per thread)
//   mode 2: LDG.128, 8 lanes read 128 B of one row, 4 rows per instruction        (not a TMEM fragment: reference for full-line access)
//   mode 3/4: the same as 0/1 with stores instead of loads
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldg_wavefront ldg_wavefront.cu && ./ldg_wavefront
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld128(const char* p, float* v) {
  asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p));
}
__device__ __forceinline__ void ld256(const char* p, float* v) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}
__device__ __forceinline__ void st128(char* p, const float* v) {
  asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
}
__device__ __forceinline__ void st256(char* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(char* base, long long n_tiles, float* sink, long long* cyc) {
  extern __shared__ char smem[];  // occupy the SM like the chain kernel does (one CTA per SM)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, hq = warp >> 2, lr = lane >> 2, lc = lane & 3;
  float acc = 0.f;
  long long t0 = clock64();
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    char* tb = base + tile * (128LL * 1024);
    float v[64];
    if (MODE == 0 || MODE == 3) {  // rows 32q + {lr, lr+8, lr+16, lr+24}; columns (bytes) 256 s + 64 hq + 16 lc
#pragma unroll
      for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          char* p = tb + (32 * q + lr + 8 * kk) * 1024 + 256 * s + 64 * hq + 16 * lc;
          if (MODE == 0) ld128(p, v + 16 * s + 4 * kk);
          else { float w[4] = {acc, 1.f, 2.f, 3.f}; st128(p, w); }
        }
    } else if (MODE == 1 || MODE == 4) {  // rows 32q + 16 (hq&1) + {lr, lr+8}; columns 256 s + 128 (hq>>1) + 32 lc
#pragma unroll
      for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          char* p = tb + (32 * q + 16 * (hq & 1) + lr + 8 * kk) * 1024 + 256 * s + 128 * (hq >> 1) + 32 * lc;
          if (MODE == 1) ld256(p, v + 16 * s + 8 * kk);
          else { float w[8] = {acc, 1.f, 2.f, 3.f, 4.f, 5.f, 6.f, 7.f}; st256(p, w); }
        }
    } else {  // mode 2: 8 lanes per row, 4 rows per instruction
#pragma unroll
      for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          char* p = tb + (32 * q + (lane >> 3) + 4 * (2 * kk + (hq & 1))) * 1024 + 256 * s + 128 * (hq >> 1) + 16 * (lane & 7);
          ld128(p, v + 16 * s + 4 * kk);
        }
    }
    if (MODE <= 2) {
#pragma unroll
      for (int i = 0; i < 64; ++i) acc += v[i];
    }
  }
  long long t1 = clock64();
  if (acc == 12345.678f) sink[0] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(char* buf, long long n_tiles, float* sink, long long* cyc, const char* what) {
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t a, b;
  cudaEventCreate(&a), cudaEventCreate(&b);
  for (int it = 0; it < 3; ++it) {
    cudaEventRecord(a);
    k<MODE><<<148, 512, 200 * 1024>>>(buf, n_tiles, sink, cyc);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
  }
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < 148; ++i) mean += (double)h[i] / 148;
  const double tiles_per_cta = (double)n_tiles / 148;
  printf("%-58s %8.3f ms  %7.1f GB/s  %8.0f cycles per 128 KB tile per SM  (%5.1f B/cycle/SM)\n", what, ms, n_tiles * 131072.0 / ms * 1e-6, mean / tiles_per_cta,
         131072.0 / (mean / tiles_per_cta));
}

int main() {
  const long long n_tiles = 2573;  // the 1-degree latent edge state at batch 8: 337 MB, partly L2 resident like the real pass
  char* buf;
  float* sink;
  long long* cyc;
  cudaMalloc(&buf, n_tiles * 131072LL);
  cudaMemset(buf, 0, n_tiles * 131072LL);
  cudaMalloc(&sink, 4);
  cudaMalloc(&cyc, 148 * 8);
  for (long long n : {2573LL, 592LL}) {  // 337 MB (L2 + HBM) and 78 MB (L2 resident after the first pass)
    printf("tiles = %lld (%.0f MB)\n", n, n * 131072.0 / 1e6);
    run<0>(buf, n, sink, cyc, "LDG.128  8 rows x 64 B per instruction (now)");
    run<1>(buf, n, sink, cyc, "LDG.256  8 rows x 128 B per instruction");
    run<2>(buf, n, sink, cyc, "LDG.128  4 rows x 128 B per instruction");
    run<3>(buf, n, sink, cyc, "STG.128  8 rows x 64 B per instruction (now)");
    run<4>(buf, n, sink, cyc, "STG.256  8 rows x 128 B per instruction");
  }
  return cudaDeviceSynchronize() != cudaSuccess;
}
