// The tail of the SEANet decoder as ONE kernel (reference models/seanet.rs:82-88 the last SEANetResnetBlock,
// :379-392 ELU + the final Conv1d 64 -> 1, modules/conv.rs:90-136 streaming left context):
//   h   = ELU(conv_k3(e8) + b_a)            e8 = ELU(x8), the f16 operand the transposed conv left behind (64 channels)
//   a9  = ELU(conv_k1(h) + b_b + x8)        x8 = the block's f32 input (the skip)
//   pcm = conv_k3(a9) + b_f                 64 -> 1, plus the i16 wire format (audio.rs:129-146)
// As three launches this moved [1920, 64] activations through L2 four times per stream and frame (h9 out and in, a9 out and
// in: 63 MB at 64 streams) for 8 MB of algorithmic output.  Here a tile of 128 rows goes e8 -> TMEM -> shared memory ->
// TMEM -> shared memory -> PCM without touching global memory in between:
//   * a tile computes a9 rows [r0 - 2, r0 + 126) of one stream, r0 = 126 j, and emits pcm rows [r0, r0 + 126): the two rows
//     of left context the final conv needs are recomputed by the tile (its first two rows) rather than handed over from the
//     neighbour, so tiles are independent; for the first tile of a stream they are the rows the previous pass left in the
//     slot state (conv_state_move puts them at the head of the a9 buffer), and the tile that owns the stream's last two
//     rows stores them at the buffer's tail for conv_state_move to save.  Rows outside the buffers are zero-filled by
//     TMA and masked.
//   * MMA 1: A = three tap-shifted TMA boxes of e8 ([128 rows x 64 ch] each, the implicit-GEMM form of the k3 conv),
//     B = W_a [64 x 192]; accumulator D1 (64 TMEM columns).  Epilogue 1 (8 warps, thread = row x half of the channels):
//     bias, ELU, f16, stored as the SWIZZLE_128B K-major operand tile of MMA 2.
//   * MMA 2: A = that tile, B = W_b [64 x 64]; accumulator D2.  Epilogue 2: bias, skip (x8 tile brought by TMA, 128-byte
//     swizzled so that a quarter-warp reads conflict-free), ELU, f16 into a shared a9 tile.
//   * final conv: one thread per output row reads its three a9 rows (192 values) from shared memory, same summation order
//     as the stand-alone kernel (bit-identical PCM).
//   * persistent: a CTA walks tiles with stride gridDim.x; the TMA producer runs one tile ahead (two stages of e8 taps + x8),
//     MMA 1 of the next tile is issued as soon as D1 has been drained, so it overlaps epilogue 2 and the final conv.
#pragma once
#include "flow_head.cuh"
#include "kernels.cuh"

namespace ptts {

static constexpr int ST_THREADS = 320;            // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quarter)
static constexpr int ST_ROWS = 128, ST_STEP = 126;
static constexpr int ST_TILE_B = ST_ROWS * 128;   // one [128 rows][128 B] tile: 16 KB
static constexpr int ST_STAGE_B = 5 * ST_TILE_B;  // three e8 taps + two halves of the f32 x8 tile
static constexpr int ST_W_B = 4 * 64 * 128;       // W_a as three [64 x 64] k-blocks + W_b: 32 KB
static constexpr int ST_SMEM = 2 * ST_STAGE_B + ST_W_B + 2 * ST_TILE_B + 1024 + 1024;   // + barriers and final-conv weights + alignment slack

struct SeanetTailParams {
  const float* b_a; const float* b_b;   // [64] biases of the k3 / k1 conv
  const float* w_f; const float* b_f;   // final conv [3][64], [1]
  __half* a9buf;                         // [n][2 + T][64]: only its first two rows per stream are read (the left context
                                         // conv_state_move brought from the slot state) and rows T, T+1 written (the last
                                         // two a9 rows, which conv_state_move saves as the next pass's left context)
  float* pcm; short* pcm16;              // [n][T] (pcm16 may be null)
  int n, T, tiles_per_stream, n_tiles;
};

__device__ __forceinline__ void st_tmem_ld32_wait(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
      "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

__global__ void __launch_bounds__(ST_THREADS, 1)
seanet_tail_kernel(const __grid_constant__ CUtensorMap map_e, const __grid_constant__ CUtensorMap map_x,
                   const __grid_constant__ CUtensorMap map_wa, const __grid_constant__ CUtensorMap map_wb, const SeanetTailParams p) {
  extern __shared__ __align__(1024) uint8_t st_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(st_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_s = smem;                                   // [2][e8 tap 0 | tap 1 | tap 2 | x8 ch 0-31 | x8 ch 32-63]
  uint8_t* w_s = smem + 2 * ST_STAGE_B;                      // W_a k-blocks 0..2 (8 KB each), W_b
  uint8_t* h_s = w_s + ST_W_B;                               // operand tile of MMA 2
  uint8_t* a9_s = h_s + ST_TILE_B;                           // a9 tile for the final conv (same swizzle)
  uint64_t* full = reinterpret_cast<uint64_t*>(a9_s + ST_TILE_B);   // [2]
  uint64_t* e_empty = full + 2;                              // [2] MMA 1 has read the e8 taps of the stage
  uint64_t* x_empty = e_empty + 2;                           // [2] epilogue 2 has read the x8 tile of the stage
  uint64_t* w_full = x_empty + 2;
  uint64_t* tf1 = w_full + 1;                                // D1 complete
  uint64_t* tf2 = tf1 + 1;                                   // D2 complete
  uint64_t* d1_empty = tf2 + 1;                              // D1 drained by epilogue 1
  uint64_t* d2_empty = d1_empty + 1;                         // D2 drained by epilogue 2
  uint64_t* h_ready = d2_empty + 1;                          // operand tile of MMA 2 written
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h_ready + 1);
  float* wf_s = reinterpret_cast<float*>(tmem_slot + 2);     // [192] final conv weights (lives in the 256 B behind the barriers)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_e); tma_prefetch_desc(&map_x); tma_prefetch_desc(&map_wa); tma_prefetch_desc(&map_wb);
    for (int s = 0; s < 2; ++s) { mbar_init(full + s, 1); mbar_init(e_empty + s, 1); mbar_init(x_empty + s, 256); }
    mbar_init(w_full, 1); mbar_init(tf1, 1); mbar_init(tf2, 1);
    mbar_init(d1_empty, 256); mbar_init(d2_empty, 256); mbar_init(h_ready, 256);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, ST_W_B);   // constants: before the dependency on the previous kernel resolves
      for (int t = 0; t < 3; ++t) tma_load_3d(w_s + t * 8192, &map_wa, w_full, 64 * t, 0, 0);
      tma_load_3d(w_s + 3 * 8192, &map_wb, w_full, 0, 0, 0);
      pdl_wait();
      int it = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
        const int s = it & 1, use = it >> 1;
        const int b = tile / p.tiles_per_stream, r0 = (tile - b * p.tiles_per_stream) * ST_STEP;
        if (use > 0) {
          mbar_wait(e_empty + s, (use - 1) & 1);
          mbar_wait(x_empty + s, (use - 1) & 1);
        }
        uint8_t* st = stage_s + s * ST_STAGE_B;
        mbar_arrive_expect_tx(full + s, ST_STAGE_B);
        // a9 row r0 - 2 + i needs e8 buffer rows (r0 - 2 + i) + tap (two state rows sit in front of the buffer) and x8 row r0 - 2 + i
        for (int t = 0; t < 3; ++t) tma_load_3d(st + t * ST_TILE_B, &map_e, full + s, 0, r0 - 2 + t, b);
        tma_load_3d(st + 3 * ST_TILE_B, &map_x, full + s, 0, r0 - 2, b);
        tma_load_3d(st + 4 * ST_TILE_B, &map_x, full + s, 64, r0 - 2, b);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = make_idesc_f16_m128(64);
    const uint64_t dwa = make_sw128_kmajor_desc(smem_u32(w_s)), dwb = make_sw128_kmajor_desc(smem_u32(w_s + 3 * 8192));
    const uint64_t dh = make_sw128_kmajor_desc(smem_u32(h_s));
    mbar_wait(w_full, 0);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int s = it & 1, use = it >> 1;
      mbar_wait(full + s, use & 1);
      if (it > 0) mbar_wait(d1_empty, (it - 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t de = make_sw128_kmajor_desc(smem_u32(stage_s + s * ST_STAGE_B));
#pragma unroll
        for (int t = 0; t < 3; ++t)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base, de + t * (ST_TILE_B >> 4) + 2 * k, dwa + t * (8192 >> 4) + 2 * k, idesc, (t | k) != 0);
        umma_commit(e_empty + s);
        umma_commit(tf1);
      }
      __syncwarp();
      mbar_wait(h_ready, it & 1);
      if (it > 0) mbar_wait(d2_empty, (it - 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        fence_proxy_async_all();   // the operand tile was written by generic-proxy stores of the epilogue warps
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_base + 64, dh + 2 * k, dwb + 2 * k, idesc, k != 0);
        umma_commit(tf2);
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue: thread = (tile row, half of the channels) =====
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int i = quad * 32 + lane;                       // tile row = TMEM lane
    const int etid = threadIdx.x - 64;
    if (etid < 192) wf_s[etid] = p.w_f[etid];             // constants: before the dependency resolves
    const uint32_t tm = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + half * 32;
    float ba[32], bb[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) { ba[c] = __ldg(p.b_a + half * 32 + c); bb[c] = __ldg(p.b_b + half * 32 + c); }
    const float bf = __ldg(p.b_f);
    pdl_wait();
    const uint32_t sw = static_cast<uint32_t>(i & 7);
    const uint32_t row_off = static_cast<uint32_t>(i) * 128u;
    asm volatile("bar.sync 1, 256;" ::: "memory");        // wf_s visible
    int it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      const int b = tile / p.tiles_per_stream, j = tile - b * p.tiles_per_stream, r0 = j * ST_STEP;
      const int r = r0 - 2 + i;                           // a9 / pcm row of this thread
      uint32_t v[32];
      // ---- epilogue 1: h = ELU(D1 + b_a) as the f16 operand tile of MMA 2
      mbar_wait(tf1, it & 1);
      tc_fence_after();
      st_tmem_ld32_wait(tm, v);
      tc_fence_before();
      mbar_arrive(d1_empty);
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {                    // four 16-byte chunks = this thread's 32 channels
        uint32_t pk[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c = c4 * 8 + q * 2;
          pk[q] = pack_half2(elu1_fast(__uint_as_float(v[c]) + ba[c]), elu1_fast(__uint_as_float(v[c + 1]) + ba[c + 1]));
        }
        *reinterpret_cast<uint4*>(h_s + row_off + (((half * 4 + c4) ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      fence_proxy_async_all();
      mbar_arrive(h_ready);
      // ---- epilogue 2: a9 = ELU(D2 + b_b + x8) as f16 into the shared a9 tile
      mbar_wait(tf2, it & 1);
      tc_fence_after();
      st_tmem_ld32_wait(tm + 64, v);
      tc_fence_before();
      mbar_arrive(d2_empty);
      const uint8_t* xs = stage_s + s * ST_STAGE_B + (3 + half) * ST_TILE_B + row_off;   // [128 rows][32 floats], 128-byte swizzle
      float4 sk[8];
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) sk[c4] = *reinterpret_cast<const float4*>(xs + ((static_cast<uint32_t>(c4) ^ sw) << 4));
      mbar_arrive(x_empty + s);
      const bool from_state = (j == 0 && i < 2);          // left context of the stream's first tile: the previous pass's last rows
      uint4 stv[4];
      if (from_state) {
        const uint4* sp = reinterpret_cast<const uint4*>(p.a9buf + (static_cast<long long>(b) * (2 + p.T) + i) * 64 + half * 32);
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) stv[c4] = sp[c4];
      }
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const float s0[8] = {sk[2 * c4].x, sk[2 * c4].y, sk[2 * c4].z, sk[2 * c4].w, sk[2 * c4 + 1].x, sk[2 * c4 + 1].y, sk[2 * c4 + 1].z, sk[2 * c4 + 1].w};
        uint32_t pk[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c = c4 * 8 + q * 2;
          pk[q] = pack_half2(elu1_fast((__uint_as_float(v[c]) + bb[c]) + s0[q * 2]), elu1_fast((__uint_as_float(v[c + 1]) + bb[c + 1]) + s0[q * 2 + 1]));
        }
        uint4 o = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        if (from_state) o = stv[c4];
        *reinterpret_cast<uint4*>(a9_s + row_off + (((half * 4 + c4) ^ sw) << 4)) = o;
        // the last two rows of the stream are the left context of the next pass
        if (r >= p.T - 2 && r < p.T)
          *reinterpret_cast<uint4*>(p.a9buf + (static_cast<long long>(b) * (2 + p.T) + 2 + r) * 64 + half * 32 + c4 * 8) = o;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");      // the a9 tile is complete
      // ---- final conv 64 -> 1 over rows i-2, i-1, i (same order of additions as seanet_final_conv_kernel)
      if (half == 0 && i >= 2 && r < p.T) {
        float acc = bf;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          const int ri = i - 2 + t;
          const uint8_t* rp = a9_s + static_cast<uint32_t>(ri) * 128u;
          const uint32_t rsw = static_cast<uint32_t>(ri & 7);
#pragma unroll
          for (int c8 = 0; c8 < 8; ++c8) {
            const uint4 u = *reinterpret_cast<const uint4*>(rp + ((static_cast<uint32_t>(c8) ^ rsw) << 4));
            const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 f = __half22float2(h[q]);
              acc += f.x * wf_s[t * 64 + c8 * 8 + 2 * q] + f.y * wf_s[t * 64 + c8 * 8 + 2 * q + 1];
            }
          }
        }
        p.pcm[static_cast<long long>(b) * p.T + r] = acc;
        if (p.pcm16) p.pcm16[static_cast<long long>(b) * p.T + r] = static_cast<short>(fminf(fmaxf(acc, -1.f), 1.f) * 32767.f);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");      // the a9 tile may be overwritten by the next tile
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 128);
}

}  // namespace ptts
