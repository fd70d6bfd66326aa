// Bring-up microbenchmark (not a test): cycles per tcgen05.mma (kind::f16, K=16) for the operand shapes and shared
// memory layouts the engine could use.  One CTA, operands are whatever shared memory holds (timing only).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../pocket_tts_b200/csrc/ptx.cuh"
using namespace ptts;

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo >> 4) << 16;
  d |= static_cast<uint64_t>(sbo >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout) << 61;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc(uint32_t m, uint32_t n) {
  return (1u << 4) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

struct Cfg { int m, n, layout_a, layout_b, kstep_a, kstep_b, sbo_a, sbo_b, kb_a, kb_b, kbmask_b, commit_every, dbl; };

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(Cfg c, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[8];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0h
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 8; ++i) mbar_init(&bar2[i], 1); mbar_fence_init(); }
  if (warp == 1) { tmem_alloc(&slot, 256); tmem_relinquish(); }
  asm volatile("fence.proxy.async;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0) {
    const uint32_t idesc = make_idesc(c.m, c.n);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 128 * 1024);
    for (int it = 0; it < 3; ++it) {
      long long t0 = 0, t1 = 0, t2 = 0;
      if (elect_one()) {
        t0 = clock64();
        for (int r = 0; r < reps; ++r) {
          const int kb = (r >> 2) & 7, k = r & 3;
          const uint64_t da = make_desc(a0 + kb * c.kb_a + k * c.kstep_a, 16, c.sbo_a, c.layout_a);
          const uint64_t db = make_desc(b0 + (kb & c.kbmask_b) * c.kb_b + k * c.kstep_b, 16, c.sbo_b, c.layout_b);
          umma_f16(tm, da, db, idesc, r != 0);
          if (c.commit_every && (r & (c.commit_every - 1)) == c.commit_every - 1) { umma_commit(&bar2[(r >> 2) & 7]); if (c.dbl) umma_commit(&bar2[((r >> 2) + 4) & 7]); }
        }
        t1 = clock64();
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, it & 1);
      if (elect_one()) { t2 = clock64(); out[it * 2] = t1 - t0; out[it * 2 + 1] = t2 - t0; }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tm, 256);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024);
  const int reps = 64;
  struct Named { const char* name; Cfg c; };
  // SW128 K-major: 128-byte rows, k-step +32 B, 8-row group 1024 B, k-block = rows*128 B
  // SW32  K-major: 32-byte rows (one k-step per tile), 8-row group 256 B, k-step tile = rows*32 B
  Named cases[] = {
      {"M128 N64  A sw128 B sw128", {128, 64, 2, 2, 32, 32, 1024, 1024, 16384, 8192, 7, 0, 0}},
      {"M128 N64  same, commit every 4", {128, 64, 2, 2, 32, 32, 1024, 1024, 16384, 8192, 7, 4, 0}},
      {"M128 N64  same, commit every 8", {128, 64, 2, 2, 32, 32, 1024, 1024, 16384, 8192, 7, 8, 0}},
      {"M128 N64  same, commit every 16", {128, 64, 2, 2, 32, 32, 1024, 1024, 16384, 8192, 7, 16, 0}},
      {"M128 N64  same, commit every 32", {128, 64, 2, 2, 32, 32, 1024, 1024, 16384, 8192, 7, 32, 0}},
      {"M128 N64  same, 2 commits every 16", {128, 64, 2, 2, 32, 32, 1024, 1024, 16384, 8192, 7, 16, 1}},
      {"M128 N64  same, commit every 1", {128, 64, 2, 2, 32, 32, 1024, 1024, 16384, 8192, 7, 1, 0}},
      {"M128 N128 A sw128 B sw128", {128, 128, 2, 2, 32, 32, 1024, 1024, 16384, 16384, 3, 0, 0}},
      {"M128 N256 A sw128 B sw128", {128, 256, 2, 2, 32, 32, 1024, 1024, 16384, 32768, 1, 0, 0}},
      {"M64  N64  A sw128 B sw128", {64, 64, 2, 2, 32, 32, 1024, 1024, 8192, 8192, 7, 0, 0}},
      {"M64  N128 A sw128 B sw128", {64, 128, 2, 2, 32, 32, 1024, 1024, 8192, 16384, 3, 0, 0}},
      {"M64  N256 A sw128 B sw128", {64, 256, 2, 2, 32, 32, 1024, 1024, 8192, 32768, 1, 0, 0}},
      {"M128 N64  A sw32  B sw128", {128, 64, 6, 2, 4096, 32, 256, 1024, 16384, 8192, 7, 0, 0}},
      {"M128 N64  A sw32  B sw32 ", {128, 64, 6, 6, 4096, 2048, 256, 256, 16384, 8192, 7, 0, 0}},
      {"M128 N16  A sw128 B sw128", {128, 16, 2, 2, 32, 32, 1024, 1024, 16384, 2048, 7, 0, 0}},
      {"M128 N32  A sw128 B sw128", {128, 32, 2, 2, 32, 32, 1024, 1024, 16384, 4096, 7, 0, 0}},
  };
  for (auto& nc : cases) {
    cudaMemset(d, 0, 64);
    mma_rate_kernel<<<1, 128, 201 * 1024>>>(nc.c, reps, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[6];
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    printf("%s: issue %.1f cyc/MMA, complete %.1f cyc/MMA  (%s)\n", nc.name, (double)h[4] / reps, (double)h[5] / reps,
           cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
  }
  return 0;
}
