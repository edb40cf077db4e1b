// grid_barrier.cu -- cost of one grid-wide barrier on B200 for a persistent kernel with one CTA per SM:
// cooperative-groups grid.sync against hand-rolled arrive / spin variants.  Build: nvcc -arch=sm_100a -O3 -rdc=true? (no: plain) 
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

template <int V>
__global__ void k_bar(unsigned long long *counter, unsigned long long epoch0, int iters, double *sink, const double *src) {
  cg::grid_group grid = cg::this_grid();
  unsigned long long target = epoch0;
  double acc = 0;
  for (int it = 0; it < iters; ++it) {
    acc += src[(threadIdx.x + it) & 1023];   // a little work between barriers (L1/L2 hit)
    if (V == 0) { grid.sync(); continue; }
    target += gridDim.x;
    __syncthreads();
    if (threadIdx.x == 0) {
      if (V == 1) { __threadfence(); atomicAdd(counter, 1ull); }
      if (V == 2) { asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(counter), "l"(1ull) : "memory"); }
      if (V == 3) { asm volatile("fence.acq_rel.gpu;" ::: "memory"); asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(counter), "l"(1ull) : "memory"); }
      if (V == 4) { asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(counter), "l"(1ull) : "memory"); }
      unsigned long long v;
      if (V == 4) {
        do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(counter) : "memory"); } while (v < target);
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
      } else {
        do { asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(counter) : "memory"); } while (v < target);
      }
    }
    __syncthreads();
  }
  if (acc == 1.2345) *sink = acc;
}

template <int V>
float run(int threads, int grid, int iters, unsigned long long *counter, unsigned long long &epoch, double *sink, double *src) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    void *args[] = {&counter, &epoch, &iters, &sink, &src};
    cudaEventRecord(e0);
    cudaError_t err = cudaLaunchCooperativeKernel((const void *)k_bar<V>, dim3(grid), dim3(threads), args, 0, 0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    if (err != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(err)); return -1; }
    if (V != 0) epoch += (unsigned long long)iters * grid;
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  unsigned long long *counter; double *sink, *src;
  cudaMalloc(&counter, 8); cudaMemset(counter, 0, 8);
  cudaMalloc(&sink, 8); cudaMalloc(&src, 8192); cudaMemset(src, 0, 8192);
  unsigned long long epoch = 0;
  const int iters = 2000;
  printf("SMs %d, %d barriers per launch; us per barrier\n", sms, iters);
  for (int threads : {128, 256, 512, 1024}) {
    for (int mult : {1, 2}) {
      if (threads * mult > 2048) continue;
      const int grid = sms * mult;
      printf("threads %4d grid %4d : grid.sync %.3f  fence+atomicAdd/ld.acquire %.3f  red.release/ld.acquire %.3f  fence.acq_rel+red.relaxed/ld.acquire %.3f  red.release/ld.relaxed+fence %.3f\n",
             threads, grid, run<0>(threads, grid, iters, counter, epoch, sink, src) * 1e3 / iters, run<1>(threads, grid, iters, counter, epoch, sink, src) * 1e3 / iters,
             run<2>(threads, grid, iters, counter, epoch, sink, src) * 1e3 / iters, run<3>(threads, grid, iters, counter, epoch, sink, src) * 1e3 / iters,
             run<4>(threads, grid, iters, counter, epoch, sink, src) * 1e3 / iters);
    }
  }
  return 0;
}
