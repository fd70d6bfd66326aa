// LSD flow head (SimpleMLPAdaLN) as ONE kernel: input_proj, six AdaLN residual blocks and the final layer, i.e. fourteen
// dependent 512-wide Linears with their LayerNorm / modulate / gate / SiLU glue (reference modules/mlp.rs:135-171
// ResBlock, :275-383 SimpleMLPAdaLN; flow_lm.rs:7-22 Euler step).  As separate launches every one of the 21 kernels
// paid a ~2 us grid hand-off plus pipeline fill for ~0.5 MB of work; here a cluster of four CTAs keeps the chain on
// chip:
//   * CTA `rank` owns features [128 rank, 128 rank + 128) of every layer.  Weights are the MMA-M operand (swap-AB,
//     64 activation rows = MMA-N), the next layer's 128 KB land (two TMA boxes) while this layer's epilogue runs.
//   * the residual stream x never leaves registers: thread = feature, 64 rows per thread (the TMEM accumulator
//     layout), so bias / gate / residual are register arithmetic straight after tcgen05.ld.
//   * LayerNorm needs row statistics over all 512 features: two transpose-reduces per warp give (sum, M2 about the
//     warp's own mean) of its 32 features, 16 such partials per row are exchanged through distributed shared memory
//     once and merged with Chan's update (equal to the oracle's two-pass variance up to rounding).
//   * the f16 operand of the next Linear (h or g, 64 x 512) is all-gathered through L2: each CTA stores its 128
//     columns, one cluster barrier, then TMA brings the full rows back in the swizzled operand layout.
// Every reduction has a fixed order, so results are bit-reproducible.  One cluster per 64 rows of the batch.
#pragma once
#include "gemm.cuh"

namespace ptts {

static constexpr int FH_DIM = 512, FH_DEPTH = 6, FH_ROWS = 64, FH_CLUSTER = 4, FH_FEATS = 128;
static constexpr int FH_LDIM = 32;                  // latent dim (the final Linear's features)
static constexpr int FH_STAGES = 8;                 // weight k-block tiles resident (16 KB each): one whole layer, in two halves
static constexpr int FH_THREADS = 320;              // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quarter)
static constexpr int FH_EROWS = FH_ROWS / 2;        // rows per epilogue thread
static constexpr int FH_LAYERS = 2 + 2 * FH_DEPTH;  // input_proj, (mlp.0, mlp.2) x 6, final
static constexpr int FH_MOD_LD = FH_DEPTH * 3 * FH_DIM + 2 * FH_DIM;
static constexpr int FH_PACK_ROWS = 2 * FH_DEPTH * FH_DIM + 128;  // packed weights: 12 x [512][512] then final [128][512]
static constexpr int FH_ACT_BYTES = 8 * FH_ROWS * 128;            // 8 k-blocks of [64 rows][64 k] f16
static constexpr int FH_W_BYTES = FH_FEATS * 128;                 // one k-block of [128 features][64 k] f16
static constexpr int FH_SMEM = FH_ACT_BYTES + FH_STAGES * FH_W_BYTES + 8 * 6 + 16 +
                               (2 * 16 * FH_ROWS + 2 * FH_ROWS) * 4 + 1024;

struct FlowHeadParams {
  const float* b_in;      // [512]
  const float* b0[FH_DEPTH];
  const float* b2[FH_DEPTH];
  const float* b_final;   // [32]
  const float* ln_w[FH_DEPTH];
  const float* ln_b[FH_DEPTH];
  const float* ws_in;     // int8 mode: per-feature weight-code scales, else null
  const float* ws0[FH_DEPTH];
  const float* ws2[FH_DEPTH];
  const float* ws_final;
  const float* mod;       // row-indexed buffers hold round_up(rows, 64) rows.  [rows][10240]: per block shift | scale | gate, then the final layer's shift | scale
  float* z32;             // [rows][32]  in/out: z += (W h + b) * alpha
  __half* z16;            // [rows][64]  the same as the f16 operand of the next LSD step / nothing reads cols 32..63
  __half* h16;            // [rows][512] scratch
  __half* g16;            // [rows][512] scratch
  float* x_dbg;           // [rows][512] residual stream after the last block (debug tap) or null
  int n;                  // valid rows
  float alpha;            // 1 / lsd_decode_steps
  unsigned long long* trace;  // bring-up: [14][8] %globaltimer stamps of cluster 0 / rank 0, or null
};

__device__ __forceinline__ void st_dsmem_f1(uint32_t cluster_addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// Per-row sums over the 32 lanes of a warp for 32 per-lane values val(0..31) (one per row): a butterfly that halves
// the number of live rows at every exchange (31 shuffles instead of 160).  Lane l ends with the sum of row l.
template <typename F>
__device__ __forceinline__ float warp_rows_sum32(F val, int lane) {
  constexpr unsigned FULL = 0xffffffffu;
  float a[16], b[8], c[4], d[2];
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2, h1 = lane & 1;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float lo = val(i), hi = val(i + 16);
    a[i] = (h16 ? hi : lo) + __shfl_xor_sync(FULL, h16 ? lo : hi, 16);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = (h8 ? a[i + 8] : a[i]) + __shfl_xor_sync(FULL, h8 ? a[i] : a[i + 8], 8);
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] = (h4 ? b[i + 4] : b[i]) + __shfl_xor_sync(FULL, h4 ? b[i] : b[i + 4], 4);
#pragma unroll
  for (int i = 0; i < 2; ++i) d[i] = (h2 ? c[i + 2] : c[i]) + __shfl_xor_sync(FULL, h2 ? c[i] : c[i + 2], 2);
  return (h1 ? d[1] : d[0]) + __shfl_xor_sync(FULL, h1 ? d[0] : d[1], 1);
}
__device__ __forceinline__ float silu_fast(float v) { return __fdividef(v, 1.f + __expf(-v)); }

__global__ void __launch_bounds__(FH_THREADS, 1)
flow_head_kernel(const __grid_constant__ CUtensorMap map_win, const __grid_constant__ CUtensorMap map_wpack,
                 const __grid_constant__ CUtensorMap map_z, const __grid_constant__ CUtensorMap map_h,
                 const __grid_constant__ CUtensorMap map_g, const FlowHeadParams p) {
  extern __shared__ __align__(1024) uint8_t fh_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fh_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* act_s = smem;                                  // 8 x 8 KB
  uint8_t* w_s = smem + FH_ACT_BYTES;                     // ring
  uint64_t* full_w = reinterpret_cast<uint64_t*>(w_s + FH_STAGES * FH_W_BYTES);
  uint64_t* empty_w = full_w + 2;
  uint64_t* full_act = empty_w + 2;
  uint64_t* tmem_full = full_act + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* stat_s = reinterpret_cast<float*>(tmem_slot + 4);  // [2][16][64]
  float* mean_s = stat_s + 2 * 16 * FH_ROWS;                // [64]
  float* rstd_s = mean_s + FH_ROWS;                         // [64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = blockIdx.z;                  // the cluster spans z
  const int row0 = blockIdx.x * FH_ROWS;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_win);
    tma_prefetch_desc(&map_wpack);
    tma_prefetch_desc(&map_z);
    tma_prefetch_desc(&map_h);
    tma_prefetch_desc(&map_g);
    for (int h = 0; h < 2; ++h) {
      mbar_init(full_w + h, 1);
      mbar_init(empty_w + h, 1);
    }
    mbar_init(full_act, 1);
    mbar_init(tmem_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  cluster_sync_all();  // every CTA of the cluster is running before anyone writes into a peer's shared memory

  // layer L: 0 input_proj, 1 + 2i mlp.0 of block i, 2 + 2i mlp.2 of block i, 13 final (rank 0 only: 32 features)
  auto kblocks = [](int L) { return L == 0 ? 1 : 8; };
  auto has_ln = [](int L) { return L == 0 || (L < FH_LAYERS - 1 && (L & 1) == 0); };
  const int n_layers_mine = (rank == 0) ? FH_LAYERS : FH_LAYERS - 1;

  // Weights of a layer sit in two 64 KB halves (k-blocks 0-3 | 4-7), one TMA box and one full/empty barrier pair each, so
  // the MMA warp waits twice and commits twice per layer instead of once per k-block (each wait + fence + commit round
  // costs ~600 cycles of single-thread latency, which at eight rounds per layer was twice the MMA time itself).
  // Half 0 is used by every layer (its load number for layer L is L), half 1 by layers >= 1 (load number L - 1).
  if (warp == 0) {
    // ===== TMA producer: next layer's weights as soon as this layer's MMAs have read theirs, and the activation operand =====
    auto load_w = [&](int L, int h) {  // lane 0 only
      mbar_arrive_expect_tx(full_w + h, (L == 0 ? 1 : 4) * FH_W_BYTES);
      if (L == 0) tma_load_3d(w_s, &map_win, full_w, 0, rank * FH_FEATS, 0);
      else tma_load_3d(w_s + h * 4 * FH_W_BYTES, &map_wpack, full_w + h, 0, (L - 1) * FH_DIM + rank * FH_FEATS, 4 * h);
    };
    if (lane == 0) {  // constants: requested before the dependency on the previous kernel resolves
      load_w(0, 0);
      load_w(1, 1);
    }
    pdl_wait();
    for (int L = 0; L < FH_LAYERS; ++L) {
      if (lane == 0 && L < n_layers_mine) {
        fence_proxy_async_all();  // the peers' generic-proxy stores of this operand, ordered by the cluster barrier
        const CUtensorMap* am = (L == 0) ? &map_z : ((L & 1) ? &map_h : &map_g);
        // one box per layer: (64 k, 64 rows, all k-blocks) lands as [k-block][row][64], i.e. the eight operand tiles
        mbar_arrive_expect_tx(full_act, kblocks(L) * FH_ROWS * 128);
        tma_load_3d(act_s, am, full_act, 0, row0, 0);
        if (p.trace && rank == 0 && blockIdx.x == 0) p.trace[L * 8 + 4] = gtime();
        if (L + 1 < n_layers_mine) {
          mbar_wait(empty_w, L & 1);  // half 0: commit number L
          load_w(L + 1, 0);
          if (L >= 1) {
            mbar_wait(empty_w + 1, (L - 1) & 1);  // half 1: commit number L - 1
            load_w(L + 1, 1);
          }
        }
        if (L == FH_LAYERS - 2) pdl_launch_dependents();
      }
      __syncwarp();
      if (L == FH_LAYERS - 1) break;
      if (has_ln(L)) cluster_sync_all();  // the LayerNorm statistics exchange
      cluster_sync_all();
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = make_idesc_f16_m128(FH_ROWS);
    const uint64_t da0 = make_sw128_kmajor_desc(smem_u32(w_s));
    const uint64_t db0 = make_sw128_kmajor_desc(smem_u32(act_s));
    for (int L = 0; L < FH_LAYERS; ++L) {
      if (L < n_layers_mine) {
        mbar_wait(full_w, L & 1);
        mbar_wait(full_act, L & 1);
        tc_fence_after();
        if (p.trace && rank == 0 && blockIdx.x == 0 && lane == 0) p.trace[L * 8 + 5] = gtime();
        if (elect_one()) {
          const int nk0 = (L == 0) ? 1 : 4;
          for (int kb = 0; kb < nk0; ++kb) {
#pragma unroll
            for (int k = 0; k < 4; ++k)  // k-block tiles are 16 KB (weights) / 8 KB (rows) apart: +1024 / +512 in the address field
              umma_f16(tmem_base, da0 + kb * (FH_W_BYTES >> 4) + 2 * k, db0 + kb * (FH_ROWS * 128 >> 4) + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(empty_w);
          if (L == 0) umma_commit(tmem_full);
        }
        __syncwarp();
        if (L >= 1) {
          mbar_wait(full_w + 1, (L - 1) & 1);
          tc_fence_after();
          if (elect_one()) {
            for (int kb = 4; kb < 8; ++kb) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16(tmem_base, da0 + kb * (FH_W_BYTES >> 4) + 2 * k, db0 + kb * (FH_ROWS * 128 >> 4) + 2 * k, idesc, 1);
            }
            umma_commit(empty_w + 1);
            umma_commit(tmem_full);
          }
          __syncwarp();
        }
        if (p.trace && rank == 0 && blockIdx.x == 0 && lane == 0) p.trace[L * 8 + 6] = gtime();
      }
      if (L == FH_LAYERS - 1) break;
      if (has_ln(L)) cluster_sync_all();  // the LayerNorm statistics exchange
      cluster_sync_all();
    }
  } else {
    // ===== epilogue: thread = (feature, half of the rows), registers = 32 rows =====
    // Two warps per TMEM lane quarter split the 64 rows, so every scheduler has two epilogue warps to interleave and
    // a thread keeps only x (32) plus one prefetched modulation operand (32) live across a layer: no spills (with
    // ~170 KB of shared memory carved out, the L1 that is left could not hold them).
    pdl_wait();
    const int quad = warp & 3;                 // TMEM lanes 32*quad .. +31
    const int half = (warp - 2) >> 2;          // rows 32*half .. +31
    const int fl = quad * 32 + lane;           // feature within this CTA's slice
    const int f = rank * FH_FEATS + fl;        // feature of the 512-wide layer
    const int etid = threadIdx.x - 64;         // 0..255
    const int r0 = row0 + half * FH_EROWS;     // first row of this thread
    const uint32_t tmem_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + half * FH_EROWS;
    float x[FH_EROWS];
#pragma unroll
    for (int r = 0; r < FH_EROWS; ++r) x[r] = 0.f;
    uint32_t stat_peer[FH_CLUSTER];
#pragma unroll
    for (int k = 0; k < FH_CLUSTER; ++k) stat_peer[k] = map_to_rank(smem_u32(stat_s), k);
    // rows past the batch (r0 + r >= n) live in the padding of every row-indexed buffer: computed, never consumed
    __half* const g_out = p.g16 + static_cast<long long>(r0) * FH_DIM + f;
    __half* const h_out = p.h16 + static_cast<long long>(r0) * FH_DIM + f;
    const float* const mod_rows = p.mod + static_cast<long long>(r0) * FH_MOD_LD + f;
    const float* const mean_h = mean_s + half * FH_EROWS;
    const float* const rstd_h = rstd_s + half * FH_EROWS;
    const uint32_t stat_off = ((rank * 4 + quad) * FH_ROWS + half * FH_EROWS + lane) * 4;

    for (int L = 0; L < FH_LAYERS; ++L) {
      const bool active = L < n_layers_mine;
#define FH_TRACE(slot) do { if (p.trace && etid == 0 && rank == 0 && blockIdx.x == 0) p.trace[L * 8 + (slot)] = gtime(); } while (0)
      FH_TRACE(0);
      const bool is_final = (L == FH_LAYERS - 1), is_mlp0 = (L & 1) && !is_final, is_mlp2 = L > 0 && !(L & 1);
      // the per-row operand of this layer's epilogue (gate, or z for the Euler step) does not depend on the
      // accumulator: it is requested on the read-only path before waiting for the MMAs
      float pre[FH_EROWS];
      if (is_final) {
        if (active && fl < FH_LDIM) {
#pragma unroll
          for (int r = 0; r < FH_EROWS; ++r) pre[r] = p.z32[(r0 + r) * FH_LDIM + fl];
        }
      } else if (is_mlp2) {
        const float* gate = mod_rows + ((L - 2) >> 1) * 3 * FH_DIM + 2 * FH_DIM;
#pragma unroll
        for (int r = 0; r < FH_EROWS; ++r) pre[r] = __ldg(gate + r * FH_MOD_LD);
      }
      float ws = 1.f, bias = 0.f;
      if (L == 0) { ws = p.ws_in ? __ldg(p.ws_in + f) : 1.f; bias = __ldg(p.b_in + f); }
      else if (is_final) { if (fl < FH_LDIM) { ws = p.ws_final ? __ldg(p.ws_final + fl) : 1.f; bias = __ldg(p.b_final + fl); } }
      else if (is_mlp0) { const int i = (L - 1) >> 1; ws = p.ws0[i] ? __ldg(p.ws0[i] + f) : 1.f; bias = __ldg(p.b0[i] + f); }
      else { const int i = (L - 2) >> 1; ws = p.ws2[i] ? __ldg(p.ws2[i] + f) : 1.f; bias = __ldg(p.b2[i] + f); }
      if (active) {
        mbar_wait(tmem_full, L & 1);
        tc_fence_after();
      }
      FH_TRACE(1);
      if (active) {
        uint32_t v[FH_EROWS];
        {
          uint32_t v0[16], v1[16];
          tmem_ld16(tmem_lane, v0);
          tmem_ld16(tmem_lane + 16, v1);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) { v[j] = v0[j]; v[16 + j] = v1[j]; }
        }
        tc_fence_before();
        if (L == 0) {
#pragma unroll
          for (int r = 0; r < FH_EROWS; ++r) x[r] = __uint_as_float(v[r]) * ws + bias;
        } else if (is_mlp2) {  // x += gate * (W g + b)
#pragma unroll
          for (int r = 0; r < FH_EROWS; ++r) x[r] += pre[r] * (__uint_as_float(v[r]) * ws + bias);
        } else if (is_mlp0) {  // g = silu(W h + b)
#pragma unroll
          for (int r = 0; r < FH_EROWS; ++r) g_out[r * FH_DIM] = __float2half_rn(silu_fast(__uint_as_float(v[r]) * ws + bias));
        } else if (fl < FH_LDIM) {  // Euler step: z += (W h + b) / S
#pragma unroll
          for (int r = 0; r < FH_EROWS; ++r) {
            const int row = r0 + r;
            if (row < p.n) {
              const float zn = pre[r] + (__uint_as_float(v[r]) * ws + bias) * p.alpha;
              p.z32[row * FH_LDIM + fl] = zn;
              p.z16[row * 64 + fl] = __float2half_rn(zn);
            }
          }
        }
      }
      if (is_final) break;
      if (L == FH_LAYERS - 2 && p.x_dbg) {
#pragma unroll
        for (int r = 0; r < FH_EROWS; ++r) p.x_dbg[static_cast<long long>(r0 + r) * FH_DIM + f] = x[r];
      }
      if (has_ln(L)) {
        // h = LN(x) [* w + b] * (1 + scale) + shift for the next block (j) or the final layer (no affine)
        const int j = L >> 1;  // 0 after input_proj, i + 1 after block i
        const float* shift = mod_rows + j * 3 * FH_DIM;
        const float* scale = shift + FH_DIM;
        float sc[FH_EROWS], sh[FH_EROWS];
#pragma unroll
        for (int r = 0; r < FH_EROWS; ++r) {  // in flight during the statistics
          sc[r] = __ldg(scale + r * FH_MOD_LD);
          sh[r] = __ldg(shift + r * FH_MOD_LD);
        }
        // Row statistics over all 512 features in ONE exchange: every warp reduces its 32 features of each row to
        // (sum, M2 about its own mean) -- two transposes, the local means handed back by shuffle -- the 16 partials per
        // row meet in every CTA's shared memory after one cluster barrier and are merged with Chan's update, which is
        // the two-pass variance up to rounding.  (Two exchanges, mean then centred squares, cost one more cluster
        // barrier per LayerNorm: ~0.9 us x 7 per step.)
        const float e = warp_rows_sum32([&](int r) { return x[r]; }, lane);
        const float mloc = e * (1.f / 32.f);  // lane l: mean of row l over this warp's features
        const float e2 = warp_rows_sum32([&](int r) { const float d = x[r] - __shfl_sync(0xffffffffu, mloc, r); return d * d; }, lane);
#pragma unroll
        for (int k = 0; k < FH_CLUSTER; ++k) {
          st_dsmem_f1(stat_peer[k] + stat_off, e);
          st_dsmem_f1(stat_peer[k] + 16 * FH_ROWS * 4 + stat_off, e2);
        }
        cluster_sync_all();
        if (etid < FH_ROWS) {
          float tot = 0.f;
#pragma unroll
          for (int k = 0; k < 16; ++k) tot += stat_s[k * FH_ROWS + etid];
          const float mean = tot * (1.f / FH_DIM);
          float m2 = 0.f;
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const float d = stat_s[k * FH_ROWS + etid] * (1.f / 32.f) - mean;
            m2 += stat_s[(16 + k) * FH_ROWS + etid] + 32.f * d * d;
          }
          mean_s[etid] = mean;
          rstd_s[etid] = 1.f / sqrtf(m2 * (1.f / FH_DIM) + 1e-6f);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const bool affine = j < FH_DEPTH;
        const float lw = affine ? __ldg(p.ln_w[affine ? j : 0] + f) : 1.f;
        const float lb = affine ? __ldg(p.ln_b[affine ? j : 0] + f) : 0.f;
#pragma unroll
        for (int r = 0; r < FH_EROWS; ++r) {
          const float y = (x[r] - mean_h[r]) * rstd_h[r] * lw + lb;
          h_out[r * FH_DIM] = __float2half_rn(y * (1.f + sc[r]) + sh[r]);
        }
      }
      // publish this layer's operand columns (generic-proxy stores, read back by the peers' TMA) and meet the cluster
      FH_TRACE(2);
      __threadfence();
      fence_proxy_async_all();
      cluster_sync_all();
      FH_TRACE(3);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 64);
}

}  // namespace ptts
