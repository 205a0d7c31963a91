// Device-side launch through a kernel passed as a template argument vs named directly (GPU box).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
struct P { float* g; const float* gout; float fixed; int n; char pad[2500]; };
struct G { int stages; int rows; int stage_bytes; };
template <int X>
__global__ void __launch_bounds__(544, 1) child(const __grid_constant__ P p, const __grid_constant__ G g) {
  extern __shared__ unsigned char smem[];
  smem[g.stage_bytes - 1 - threadIdx.x] = 1;
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += gridDim.x * blockDim.x)
    p.g[i] = *p.gout + smem[g.stage_bytes - 1] + X;
}
struct GP { unsigned grid, block, smem; };
template <auto Kernel, typename... A>
__global__ void gate_auto(const __grid_constant__ P p, GP gp, A... rest) {
  if (threadIdx.x != 0) return;
  Kernel<<<gp.grid, gp.block, gp.smem, cudaStreamTailLaunch>>>(p, rest...);
  p.g[p.n] = 100.f + (float)(int)cudaGetLastError();
}
template <typename... A>
__global__ void gate_named(const __grid_constant__ P p, GP gp, A... rest) {
  if (threadIdx.x != 0) return;
  child<0><<<gp.grid, gp.block, gp.smem, cudaStreamTailLaunch>>>(p, rest...);
  p.g[p.n] = 100.f + (float)(int)cudaGetLastError();
}
__global__ void gate_plain(const __grid_constant__ P p, GP gp, G geo) {
  if (threadIdx.x != 0) return;
  child<0><<<gp.grid, gp.block, gp.smem, cudaStreamTailLaunch>>>(p, geo);
  p.g[p.n] = 100.f + (float)(int)cudaGetLastError();
}
int main(int argc, char** argv) {
  const int which = argc > 1 ? atoi(argv[1]) : 0;
  float *g, *go;
  cudaMalloc(&g, 4 << 20); cudaMalloc(&go, 4);
  cudaMemset(g, 0, 4 << 20);
  float v = 2.f; cudaMemcpy(go, &v, 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(child<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  P p; p.g = g; p.gout = go; p.fixed = 1.f; p.n = (1 << 20) - 1;
  G geo = {2, 8, 131072};
  GP gp{148, 544, 131072};
  if (which == 0) gate_auto<child<0>, G><<<1, 32>>>(p, gp, geo);
  if (which == 1) gate_named<G><<<1, 32>>>(p, gp, geo);
  if (which == 2) gate_plain<<<1, 32>>>(p, gp, geo);
  cudaError_t e = cudaDeviceSynchronize();
  float h, st;
  cudaMemcpy(&h, g + 12345, 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(&st, g + p.n, 4, cudaMemcpyDeviceToHost);
  printf("which %d: sync=%s launch_status=%g g[12345]=%g (expect 3)\n", which, cudaGetErrorString(e), st - 100.f, h);
  return 0;
}
