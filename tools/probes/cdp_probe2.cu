// Which ingredient makes a device-side tail launch not run?  (GPU box)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
struct P { float* g; const float* gout; float fixed; int n; char pad[2500]; };
struct S { float* g; int n; };
__global__ void cA(float* g) { g[threadIdx.x] = 1.f; }
__global__ void cB(S p) { p.g[100 + threadIdx.x] = 2.f; }
__global__ void cC(P p) { p.g[200 + threadIdx.x] = 3.f; }
__global__ void cD(const __grid_constant__ P p) { p.g[300 + threadIdx.x] = 4.f; }
__global__ void __launch_bounds__(544, 1) cE(const __grid_constant__ P p) {
  extern __shared__ unsigned char smem[];
  smem[threadIdx.x] = 5; __syncthreads();
  p.g[400 + (threadIdx.x & 31)] = smem[3];
}
__global__ void gate(const __grid_constant__ P p, int which, unsigned smem, int mode, unsigned threads, unsigned blocks) {
  if (threadIdx.x != 0) return;
  cudaStream_t st = mode == 0 ? cudaStreamTailLaunch : cudaStreamFireAndForget;
  if (which == 0) cA<<<1, 32, 0, st>>>(p.g);
  if (which == 1) { S s{p.g, p.n}; cB<<<1, 32, 0, st>>>(s); }
  if (which == 2) cC<<<1, 32, 0, st>>>(p);
  if (which == 3) cD<<<1, 32, 0, st>>>(p);
  if (which == 4) cE<<<blocks, threads, smem, st>>>(p);
  p.g[1000 + which] = 100.f + (float)(int)cudaGetLastError();
}
int main(int argc, char** argv) {
  const int only_mode = argc > 1 ? atoi(argv[1]) : -1, only_which = argc > 2 ? atoi(argv[2]) : -1;
  const unsigned only_smem = argc > 3 ? (unsigned)atoi(argv[3]) : 0u;
  const unsigned threads = argc > 4 ? (unsigned)atoi(argv[4]) : 544u, blocks = argc > 5 ? (unsigned)atoi(argv[5]) : 148u;
  float *g, *go;
  cudaMalloc(&g, 1 << 20); cudaMalloc(&go, 4);
  cudaFuncSetAttribute(cE, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  size_t lim = 0; cudaDeviceGetLimit(&lim, cudaLimitDevRuntimePendingLaunchCount); printf("pending launch limit %zu\n", lim);
  P p; p.g = g; p.gout = go; p.fixed = 1.f; p.n = 1 << 18;
  for (int mode = 0; mode < 2; ++mode)
  for (int which = 0; which < 5; ++which)
    for (unsigned smem : {only_smem}) {
      if (mode != only_mode || which != only_which) continue;
      cudaMemset(g, 0, 1 << 20);
      gate<<<1, 32>>>(p, which, smem, mode, threads, blocks);
      cudaError_t e = cudaDeviceSynchronize();
      float h[5], st;
      for (int k = 0; k < 5; ++k) cudaMemcpy(&h[k], g + 100 * k, 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(&st, g + 1000 + which, 4, cudaMemcpyDeviceToHost);
      printf("mode %d which %d smem %u: sync=%s launch_status=%g  g = %g %g %g %g %g\n", mode, which, smem, cudaGetErrorString(e), st - 100.f,
             h[0], h[1], h[2], h[3], h[4]);
    }
  return 0;
}
