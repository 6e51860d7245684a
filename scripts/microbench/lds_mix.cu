// Microbenchmark (B200): cost of the per-entry broadcast in mb_k2_stars relative to the 512-byte
// factor-row read, and whether warp shuffles compete with shared-memory loads for the same pipe.
// Every mode does one conflict-free LDS.128 row read (4 wavefronts) per step plus:
//   0: nothing            1: broadcast LDS.128      2: broadcast LDS.64     3: broadcast LDS.64 + LDS.32
//   4: 2 x SHFL           5: 3 x SHFL               6: 4 x SHFL
// Prints cycles per step per SM (24 warps resident).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kRows = 64;
constexpr int kThreads = 768;

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) bench(int iters, unsigned long long* cycles, uint32_t* sink) {
  __shared__ __align__(16) uint32_t table[kRows * 128];
  __shared__ __align__(16) uint32_t ent[64 * 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kRows * 128; i += kThreads) table[i] = i * 2654435761u;
  for (int i = threadIdx.x; i < 256; i += kThreads) ent[i] = (i * 40503u) & 63u;
  __syncthreads();
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(table) + lane * 16u;
  const uint32_t ebase = (uint32_t)__cvta_generic_to_shared(ent);
  uint32_t acc = 0, r = warp * 7 + 1, mine = lane * 2654435761u;
  unsigned long long t0 = clock64();
#pragma unroll 4
  for (int it = 0; it < iters; ++it) {
    uint32_t x = 0, y = 0, z = 0, w = 0;
    const uint32_t ea = ebase + ((it & 63) << 4);
    if (MODE == 1) asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(ea));
    if (MODE == 2) asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(x), "=r"(y) : "r"(ea));
    if (MODE == 3) {
      asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(x), "=r"(y) : "r"(ea));
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(z) : "r"(ea + 8));
    }
    if (MODE >= 4) {
      x = __shfl_sync(0xffffffffu, mine + it, it & 31);
      y = __shfl_sync(0xffffffffu, mine ^ it, it & 31);
      if (MODE >= 5) z = __shfl_sync(0xffffffffu, mine + 3 * it, it & 31);
      if (MODE >= 6) w = __shfl_sync(0xffffffffu, mine + 5 * it, it & 31);
    }
    r = (r + 11 + (x & 1)) & (kRows - 1);
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sbase + r * 512u));
    acc += v.x ^ v.y ^ v.z ^ v.w ^ x ^ y ^ z ^ w;
  }
  unsigned long long t1 = clock64();
  sink[blockIdx.x * kThreads + threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(int sms, int iters, unsigned long long* cyc, uint32_t* sink) {
  bench<MODE><<<sms, kThreads>>>(iters, cyc, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("mode %d: %s\n", MODE, cudaGetErrorString(e)); return; }
  unsigned long long h[256];
  cudaMemcpy(h, cyc, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < sms; ++i) avg += (double)h[i];
  avg /= sms;
  printf("mode %d: %.2f cycles per step per SM (24 warps => %.2f per warp-step)\n", MODE, avg / iters, avg / iters / 24);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned long long* cyc; uint32_t* sink;
  cudaMalloc(&cyc, sizeof(unsigned long long) * sms);
  cudaMalloc(&sink, sizeof(uint32_t) * sms * kThreads);
  const int iters = 8192;
  run<0>(sms, iters, cyc, sink); run<1>(sms, iters, cyc, sink); run<2>(sms, iters, cyc, sink);
  run<3>(sms, iters, cyc, sink); run<4>(sms, iters, cyc, sink); run<5>(sms, iters, cyc, sink);
  run<6>(sms, iters, cyc, sink);
  return 0;
}
