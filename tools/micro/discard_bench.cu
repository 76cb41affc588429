// Throughput of discard.global.L2 (drop a dirty 128-byte line from L2 without write-back) per SM, by warps and lanes.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o discard_bench discard_bench.cu ; ./discard_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void fill(uint4* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] = make_uint4((uint32_t)i, 1, 2, 3);
}
// every block drops its contiguous share of the buffer; `lanes` lanes of each warp are active
__global__ void drop(char* p, size_t lines, int lanes) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane >= lanes) return;
  const size_t per_block = lines / gridDim.x;
  const size_t b0 = per_block * blockIdx.x;
  for (size_t i = (size_t)warp * lanes + lane; i < per_block; i += (size_t)nw * lanes)
    asm volatile("discard.global.L2 [%0], 128;" ::"l"(p + (b0 + i) * 128) : "memory");
}
__global__ void touch(char* p, size_t lines, int lanes) {      // same loop with a 16-byte store per line, for scale
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane >= lanes) return;
  const size_t per_block = lines / gridDim.x;
  const size_t b0 = per_block * blockIdx.x;
  for (size_t i = (size_t)warp * lanes + lane; i < per_block; i += (size_t)nw * lanes)
    *reinterpret_cast<uint4*>(p + (b0 + i) * 128) = make_uint4(0, 0, 0, 0);
}

int main() {
  const size_t bytes = 96ull << 20;             // fits in L2 (126 MB): the lines are dirty and resident when dropped
  const size_t lines = bytes / 128;
  char* buf;
  cudaMalloc(&buf, bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; ++mode)
    for (int warps : {1, 2, 4, 8, 16})
      for (int lanes : {14, 32}) {
        float best = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
          fill<<<148 * 8, 256>>>(reinterpret_cast<uint4*>(buf), bytes / 16);
          cudaDeviceSynchronize();
          cudaEventRecord(e0);
          if (mode == 0) drop<<<148, warps * 32>>>(buf, lines, lanes);
          else touch<<<148, warps * 32>>>(buf, lines, lanes);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          best = ms < best ? ms : best;
        }
        const double per_sm = (double)lines / 148;
        printf("%s warps/SM=%2d lanes=%2d: %8.1f us, %6.1f ns per line per SM (%.0f lines per SM)\n", mode ? "store16 " : "discard ",
               warps, lanes, best * 1e3, best * 1e6 / per_sm, per_sm);
      }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
