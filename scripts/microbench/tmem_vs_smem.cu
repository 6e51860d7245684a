// Microbenchmark (B200): can tensor memory serve table reads IN ADDITION to shared memory?
// Each warp repeatedly reads a 512-byte table row selected by a warp-uniform index, the access
// pattern of mb_k2_stars' factor lookups, (a) from shared memory with LDS.128, (b) from TMEM with
// tcgen05.ld.32x32b.x4, (c) alternating.  Prints bytes/clk/SM for each.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_vs_smem tmem_vs_smem.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kRows = 90;           // table rows (classes)
constexpr int kThreads = 768;

__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ uint4 tmem_ld_x4(uint32_t taddr) {
  uint4 v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(taddr));
  return v;
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// mode 0: smem only, 1: tmem only, 2: alternate 1:1, 3: 2 smem : 1 tmem
__global__ void __launch_bounds__(kThreads, 1) bench(int mode, int iters, unsigned long long* cycles, uint32_t* sink) {
  __shared__ __align__(16) uint32_t table[kRows * 128];   // 100 rows x 512 B
  __shared__ uint32_t tmem_base_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kRows * 128; i += kThreads) table[i] = i * 2654435761u;
  if (warp == 0) {
    uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tmem_base_s);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(dst) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base_s;
  // fill: warps 0..3 write their lane quadrant: row r -> columns 4r..4r+3 of every lane
  if (warp < 4) {
    for (int r = 0; r < kRows; ++r) {
      uint32_t taddr = tbase + ((uint32_t)(warp * 32) << 16) + 4u * r;
      uint32_t a = table[r * 128 + lane * 4], b = table[r * 128 + lane * 4 + 1];
      uint32_t c = table[r * 128 + lane * 4 + 2], d = table[r * 128 + lane * 4 + 3];
      asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(table) + lane * 16u;
  const uint32_t tlane = tbase + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0, r = warp * 7 + 1;
  unsigned long long t0 = clock64();
  for (int it = 0; it < iters; it += 4) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      r = (r * 13 + 5) % kRows;   // warp-uniform row
      bool use_t = mode == 1 || (mode == 2 && (u & 1)) || (mode == 3 && u == 3);
      if (use_t) v[u] = tmem_ld_x4(tlane + 4u * r);
      else v[u] = lds128(sbase + r * 512u);
    }
    if (mode != 0) tmem_wait_ld();
#pragma unroll
    for (int u = 0; u < 4; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  unsigned long long t1 = clock64();
  sink[blockIdx.x * kThreads + threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned long long* cyc; uint32_t* sink;
  cudaMalloc(&cyc, sizeof(unsigned long long) * sms);
  cudaMalloc(&sink, sizeof(uint32_t) * sms * kThreads);
  const int iters = 4096;
  uint32_t ref = 0;
  for (int mode = 0; mode < 4; ++mode) {
    bench<<<sms, kThreads>>>(mode, iters, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
    unsigned long long h[256]; cudaMemcpy(h, cyc, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost);
    uint32_t s0; cudaMemcpy(&s0, sink, 4, cudaMemcpyDeviceToHost);
    if (mode == 0) ref = s0;
    double avg = 0; for (int i = 0; i < sms; ++i) avg += (double)h[i]; avg /= sms;
    double bytes = (double)iters * 512.0 * (kThreads / 32);
    printf("mode %d: %.0f cycles/SM, %.1f B/clk/SM, checksum %s\n", mode, avg, bytes / avg, s0 == ref ? "same" : "DIFFERENT");
  }
  return 0;
}
