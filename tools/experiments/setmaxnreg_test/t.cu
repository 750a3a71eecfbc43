// Minimal liveness test of setmaxnreg on sm_100a: 384 threads, warpgroup 0 shrinks, warpgroups 1 and 2 (optionally) grow.
#include <cstdio>
#include <cuda_runtime.h>
template <int DEC, int INC>
__global__ void __launch_bounds__(384, 1) k(int* out, int iters) {
  __shared__ volatile int flag[12];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 12) flag[threadIdx.x] = 0;
  __syncthreads();
  if (warp < 4) {
    if (DEC) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(DEC ? DEC : 24));
    // consumer of the other warpgroups' flags: waits for every softmax-like warp, `iters` rounds
    for (int it = 1; it <= iters; ++it) {
      if (lane == 0)
        for (int w = 4; w < 12; ++w) while (flag[w] < it) { }
      __syncwarp();
      if (lane == 0) flag[warp] = it;
    }
  } else {
    if (INC) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(INC ? INC : 256));
    float acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) acc[i] = threadIdx.x * 0.001f + i;
    for (int it = 1; it <= iters; ++it) {
#pragma unroll
      for (int i = 0; i < 64; ++i) acc[i] = acc[i] * 1.0001f + acc[(i + 7) & 63];
      __syncwarp();
      if (lane == 0) flag[warp] = it;
      if (lane == 0) while (flag[warp & 3] < it) { }   // wait for a warp of warpgroup 0
      __syncwarp();
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) s += acc[i];
    if (s == 12345.678f) out[1] = 1;
  }
  __syncthreads();
  if (threadIdx.x == 0) out[0] = 1;
}
template <int DEC, int INC>
static void run(const char* name) {
  int* d; cudaMalloc(&d, 8); cudaMemset(d, 0, 8);
  k<DEC, INC><<<148, 384>>>(d, 1000);
  cudaError_t e = cudaDeviceSynchronize();
  int h[2] = {0, 0}; cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
  printf("SETMAXNREG_TEST %s: %s done=%d\n", name, cudaGetErrorString(e), h[0]);
  cudaFree(d);
}
int main() {
  run<0, 0>("none");
  run<88, 0>("dec88");
  run<88, 208>("dec88_inc208");
  run<40, 232>("dec40_inc232");
  return 0;
}
