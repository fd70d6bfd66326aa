// Bring-up microbenchmark (not a test): bytes per second one SM can stream into shared memory with 128 KB in flight,
// for the weight-tile fetch patterns the GEMM could use.  Every CTA streams its own 16 KB tiles through an 8-slot ring.
//   a  tensor TMA, row-major matrix [F][K] (128 rows x 128 B per box, row stride 2 KB)      -- what gemm.cuh does
//   b  tensor TMA, pre-tiled matrix (each 128 x 64 tile contiguous, 16 KB)
//   c  cp.async.bulk 1-D copy of a contiguous 16 KB tile
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_rate tma_rate.cu && ./tma_rate
#include <cstdint>
#include <cstdio>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../pocket_tts_b200/csrc/ptx.cuh"
#include "../../pocket_tts_b200/csrc/host_util.h"
using namespace ptts;

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// mode 0: map is [F][K=1024] row-major, CTA c owns rows [128 c', ...): tile t -> k-block t % 16, row tile base + t / 16
// mode 1: map is the pre-tiled image, tile index linear; mode 2: 1-D bulk copies of the same image
__global__ void __launch_bounds__(128, 1) tma_rate_kernel(const __grid_constant__ CUtensorMap map, const __grid_constant__ CUtensorMap amap, const __half* base, int mode,
                                                         int tiles_per_cta, int region_tiles, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[8];
  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) mbar_init(full + s, 1);
    mbar_fence_init();
    tma_prefetch_desc(&map);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long first = static_cast<long long>(blockIdx.x) * region_tiles;  // this CTA's private tiles
    auto issue = [&](int t, int s) {
      const long long tile = first + (t % region_tiles);
      mbar_arrive_expect_tx(full + s, mode == 3 ? 16384 + 8192 : 16384);
      if (mode == 3) {  // the GEMM's stage: a private weight tile plus an activation tile every CTA reads
        tma_load_3d(smem + s * 24576, &map, full + s, static_cast<int>(tile % 16) * 64, static_cast<int>(tile / 16) * 128, 0);
        tma_load_3d(smem + s * 24576 + 16384, &amap, full + s, (t % 16) * 64, 0, 0);
      } else if (mode == 0) tma_load_3d(smem + s * 16384, &map, full + s, static_cast<int>(tile % 16) * 64, static_cast<int>(tile / 16) * 128, 0);
      else if (mode == 1) tma_load_3d(smem + s * 16384, &map, full + s, 0, static_cast<int>(tile) * 128, 0);
      else bulk_load_1d(smem + s * 16384, base + tile * 8192, 16384, full + s);
    };
    const long long t0 = clock64();
    for (int t = 0; t < 8 && t < tiles_per_cta; ++t) issue(t, t);
    for (int t = 0; t < tiles_per_cta; ++t) {
      const int s = t & 7;
      mbar_wait(full + s, (t >> 3) & 1);
      if (t + 8 < tiles_per_cta) issue(t + 8, s);
    }
    out[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  const int F = 128 * 16 * 148;          // 148 CTAs x 16 row tiles... sized below per run
  (void)F;
  long long* d_out;
  cudaMalloc(&d_out, 148 * 8);
  cudaFuncSetAttribute(tma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 24576 + 1024);
  TmapCache tmaps;
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  for (int region_tiles : {16}) {       // 256 KB per CTA (L2-resident, re-read) or 16 MB per CTA (streams from HBM)
    for (int grid : {1, 48, 148}) {
      const size_t tiles_total = (size_t)grid * region_tiles;
      __half* buf;
      cudaMalloc(&buf, tiles_total * 16384);
      cudaMemset(buf, 0, tiles_total * 16384);
      const int tiles_per_cta = 2048;
      __half* abuf; cudaMalloc(&abuf, 64 * 1024 * 2); cudaMemset(abuf, 0, 64 * 1024 * 2);
      const CUtensorMap am = tmaps.get(abuf, 1024, 64, 1, 1024, 64 * 1024, 64, 1);
      for (int mode = 0; mode < 4; ++mode) {
        // mode 0: [rows = tiles_total/16*128][K = 1024]; mode 1/2: [tiles_total*128][64]
        const CUtensorMap& m = (mode == 0 || mode == 3) ? tmaps.get(buf, 1024, (long long)tiles_total / 16 * 128, 1, 1024, (long long)tiles_total * 8192, 128, 1)
                                           : tmaps.get(buf, 64, (long long)tiles_total * 128, 1, 64, (long long)tiles_total * 8192, 128, 1);
        for (int rep = 0; rep < 2; ++rep) tma_rate_kernel<<<grid, 128, 8 * 24576 + 1024>>>(m, am, buf, mode, tiles_per_cta, region_tiles, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<long long> h(grid);
        cudaMemcpy(h.data(), d_out, grid * 8, cudaMemcpyDeviceToHost);
        double worst = 0, sum = 0;
        for (long long c : h) { worst = c > worst ? c : worst; sum += c; }
        const double bytes = (double)tiles_per_cta * (mode == 3 ? 24576 : 16384);
        printf("region %5d KB/CTA grid %3d mode %c: %.1f B/clk per SM (mean), aggregate %.2f TB/s at %d MHz  (%s)\n", region_tiles * 16, grid,
               "abcd"[mode], bytes / (sum / grid), bytes * grid / (worst / (clk_khz * 1e3)) / 1e12, clk_khz / 1000, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
      }
      cudaFree(buf); cudaFree(abuf);
    }
  }
  return 0;
}
