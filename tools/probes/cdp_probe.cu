// Probe (GPU box): does a device-side tail launch honour cudaFuncAttributeMaxDynamicSharedMemorySize set from the
// host, with a 2.5 KB __grid_constant__ parameter block forwarded from the parent?  And what does a gate launch cost
// when it has nothing to do, compared with a full persistent grid that returns at once?
//   nvcc -gencode arch=compute_100a,code=sm_100a -rdc=true -O3 tools/probes/cdp_probe.cu -o /tmp/cdp_probe -lcudadevrt
#include <cuda_runtime.h>
#include <cstdio>
struct P { float* g; const float* gout; float fixed; int n; char pad[2500]; };
struct G { int stages; int rows; int stage_bytes; };
template <int X>
__global__ void __launch_bounds__(544, 1) child(const __grid_constant__ P p, const __grid_constant__ G g) {
  extern __shared__ unsigned char smem[];
  if (*p.gout == p.fixed && X == 1) return;
  smem[g.stage_bytes - 1 - threadIdx.x] = 1;
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += gridDim.x * blockDim.x)
    p.g[i] = *p.gout + smem[g.stage_bytes - 1] + X;
}
struct GP { unsigned grid, block, smem; };
template <auto Kernel, typename... A>
__global__ void gate(const __grid_constant__ P p, GP gp, A... rest) {
  if (*p.gout == p.fixed) return;
  if (threadIdx.x != 0) return;
  Kernel<<<gp.grid, gp.block, gp.smem, cudaStreamTailLaunch>>>(p, rest...);
  p.g[(1 << 20) - 1] = (float)(int)cudaGetLastError();   // device-side launch status
}
__global__ void after(float* g, float* out) { *out = g[12345]; }
int main() {
  float *g, *go, *out;
  cudaMalloc(&g, 4 << 20); cudaMalloc(&go, 4); cudaMalloc(&out, 4);
  cudaMemset(g, 0, 4 << 20);
  cudaFuncSetAttribute(child<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(child<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  P p; p.g = g; p.gout = go; p.fixed = 1.f; p.n = 1 << 20;
  G geo = {2, 8, 131072};
  for (unsigned smem : {40000u, 49152u, 65536u, 131072u})
  for (float v : {2.f, 1.f}) {
    cudaMemcpy(go, &v, 4, cudaMemcpyHostToDevice);
    cudaMemset(g, 0, 4 << 20);
    p.n = (1 << 20) - 1;
    geo.stage_bytes = (int)smem;
    gate<child<0>, G><<<1, 32>>>(p, GP{148, 544, smem}, geo);
    after<<<1, 1>>>(g, out);
    cudaError_t e = cudaDeviceSynchronize();
    float h, st, direct; cudaMemcpy(&h, out, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&direct, g + 12345, 4, cudaMemcpyDeviceToHost);
    printf("   direct read after sync: %g; ", direct);
    cudaMemcpy(&st, g + (1 << 20) - 1, 4, cudaMemcpyDeviceToHost);
    printf("smem=%u gout=%g: %s, device launch status %g, g[12345]=%g (expect %g)\n", smem, v, cudaGetErrorString(e), st, h,
           v == 1.f ? 0.f : v + 1.f);
  }
  geo.stage_bytes = 131072;
  // cost of an idle gate vs an idle full grid, between two real kernels
  float one = 1.f; cudaMemcpy(go, &one, 4, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 3; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      for (int i = 0; i < 200; ++i) {
        if (mode == 1) gate<child<0>, G><<<1, 32>>>(p, GP{148, 544, 131072}, geo);
        if (mode == 2) child<1><<<148, 544, 131072>>>(p, geo);
        after<<<1, 1>>>(g, out);
      }
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("mode %d (%s): %.2f us per iteration\n", mode, mode == 0 ? "no fix-up" : mode == 1 ? "idle gate" : "idle full grid", ms * 5.f);
  }
  return 0;
}
