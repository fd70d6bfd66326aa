// LSD flow head (SimpleMLPAdaLN) as ONE kernel for ALL Euler steps of a frame: per step input_proj, six AdaLN residual
// blocks and the final layer, i.e. fourteen dependent 512-wide Linears with their LayerNorm / modulate / gate / SiLU glue
// (reference modules/mlp.rs:135-171 ResBlock, :275-383 SimpleMLPAdaLN; flow_lm.rs:7-22 lsd_decode loop).  The
// modulations of every step are computed beforehand by one Linear over all steps (mlp.rs:322-368), so the chain inside
// the kernel depends on nothing but itself.
//   * The batch is cut into chunks of 16 rows; a cluster of four CTAs owns a chunk (64 streams = 4 clusters = 16 SMs).
//     CTA `rank` owns features [128 rank, 128 rank + 128) of every layer: weights are the MMA-M operand (swap-AB), the 16
//     rows the MMA-N operand; the next layer's 128 KB of weights land (two TMA boxes) while this layer runs.
//   * the residual stream x never leaves registers: thread = feature, 8 rows per thread (the TMEM accumulator
//     layout), so bias / gate / residual are register arithmetic straight after tcgen05.ld.
//   * LayerNorm needs row statistics over all 512 features: a butterfly per warp gives (sum, M2 about the warp's own
//     mean) of its 32 features, the 16 partials per row are exchanged through distributed shared memory once and
//     merged with Chan's update (equal to the oracle's two-pass variance up to rounding).
//   * the f16 operand of the next Linear (h or g, 16 x 512) never leaves the cluster either: every epilogue thread
//     writes its values straight into the operand image (SWIZZLE_128B K-major) in the shared memory of all four CTAs,
//     one cluster barrier, and the next layer's MMAs start.  The image is double-buffered by layer parity, so a CTA
//     that is still multiplying layer L is never overwritten by a peer that has already finished it.
// Every reduction has a fixed order, so results are bit-reproducible.
#pragma once
#include "gemm.cuh"

namespace ptts {

static constexpr int FH_DIM = 512, FH_DEPTH = 6, FH_ROWS = 16, FH_CLUSTER = 4, FH_FEATS = 128;
static constexpr int FH_LDIM = 32;                  // latent dim (the final Linear's features)
static constexpr int FH_STAGES = 8;                 // weight k-block tiles resident (16 KB each): one whole layer, in two halves
static constexpr int FH_THREADS = 320;              // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quarter)
static constexpr int FH_EROWS = FH_ROWS / 2;        // rows per epilogue thread
static constexpr int FH_LAYERS = 2 + 2 * FH_DEPTH;  // input_proj, (mlp.0, mlp.2) x 6, final
static constexpr int FH_MOD_LD = FH_DEPTH * 3 * FH_DIM + 2 * FH_DIM;
static constexpr int FH_PACK_ROWS = 2 * FH_DEPTH * FH_DIM + 128;  // packed weights: 12 x [512][512] then final [128][512]
static constexpr int FH_ACT_KB = FH_ROWS * 128;                   // one k-block of the operand image: [16 rows][64 k] f16
static constexpr int FH_ACT_BYTES = 8 * FH_ACT_KB;                // 8 k-blocks
static constexpr int FH_W_BYTES = FH_FEATS * 128;                 // one k-block of [128 features][64 k] f16
static constexpr int FH_MAX_STEPS = 64;
static constexpr int FH_SMEM = 2 * FH_ACT_BYTES + FH_STAGES * FH_W_BYTES + 8 * 6 + 16 +
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
  const float* mod;       // step s at mod + s * mod_step_stride: [rows][10240], per block shift | scale | gate, then the final layer's shift | scale.  Row-indexed buffers hold round_up(rows, 64) rows
  long long mod_step_stride;
  float* z32;             // [rows][32]  in/out: z += (W h + b) * alpha, once per step
  __half* z16;            // [rows][64]  the same as f16 (cols 32..63 zero): the operand of input_proj
  float* x_dbg;           // [rows][512] residual stream after the last block of the last step (debug tap) or null
  int n;                  // valid rows
  int steps;              // Euler steps (lsd_decode_steps)
  float alpha;            // 1 / lsd_decode_steps
  unsigned long long* trace;  // bring-up: [14][8] %globaltimer stamps of cluster 0 / rank 0 (last step), or null
};

__device__ __forceinline__ void st_dsmem_f1(uint32_t cluster_addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_dsmem_u32(uint32_t cluster_addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// 32 lanes x 8 consecutive 32-bit columns -> 8 registers per thread, load and wait in one statement (the registers are
// not valid before tcgen05.wait::ld)
__device__ __forceinline__ void tmem_ld8_wait(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
      : "r"(taddr)
      : "memory");
}

// Per-row sums over the 32 lanes of a warp for 8 per-lane values val(0..7) (one per row): a butterfly that halves the
// number of live rows at every exchange (9 shuffles instead of 40).  Every lane l ends with the sum of row l >> 2.
template <typename F>
__device__ __forceinline__ float warp_rows_sum8(F val, int lane) {
  constexpr unsigned FULL = 0xffffffffu;
  float a[4], b[2];
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float lo = val(i), hi = val(i + 4);
    a[i] = (h16 ? hi : lo) + __shfl_xor_sync(FULL, h16 ? lo : hi, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) b[i] = (h8 ? a[i + 2] : a[i]) + __shfl_xor_sync(FULL, h8 ? a[i] : a[i + 2], 8);
  float c = (h4 ? b[1] : b[0]) + __shfl_xor_sync(FULL, h4 ? b[0] : b[1], 4);
  c += __shfl_xor_sync(FULL, c, 2);
  c += __shfl_xor_sync(FULL, c, 1);
  return c;
}
__device__ __forceinline__ float silu_fast(float v) { return __fdividef(v, 1.f + __expf(-v)); }

// byte offset of (row r, feature k) in the operand image [k-block][16 rows][64 k], SWIZZLE_128B K-major
__device__ __forceinline__ uint32_t fh_img_off(int r, int k) {
  return static_cast<uint32_t>(k >> 6) * FH_ACT_KB + static_cast<uint32_t>(r) * 128u +
         (static_cast<uint32_t>(((k & 63) >> 3) ^ (r & 7)) << 4) + static_cast<uint32_t>(k & 7) * 2u;
}

__global__ void __launch_bounds__(FH_THREADS, 1)
flow_head_kernel(const __grid_constant__ CUtensorMap map_win, const __grid_constant__ CUtensorMap map_wpack, const FlowHeadParams p) {
  extern __shared__ __align__(1024) uint8_t fh_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fh_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* act_s = smem;                                  // [2][8 k-blocks][16 rows][128 B]
  uint8_t* w_s = smem + 2 * FH_ACT_BYTES;                 // ring
  uint64_t* full_w = reinterpret_cast<uint64_t*>(w_s + FH_STAGES * FH_W_BYTES);
  uint64_t* empty_w = full_w + 2;
  uint64_t* tmem_full = empty_w + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 2);
  float* stat_s = reinterpret_cast<float*>(tmem_slot + 4);  // [2][16 partials][16 rows]
  float* mean_s = stat_s + 2 * 16 * FH_ROWS;                // [16]
  float* rstd_s = mean_s + FH_ROWS;                         // [16]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = blockIdx.z;                  // the cluster spans z
  const int row0 = blockIdx.x * FH_ROWS;
  const int S = p.steps;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_win);
    tma_prefetch_desc(&map_wpack);
    for (int h = 0; h < 2; ++h) {
      mbar_init(full_w + h, 1);
      mbar_init(empty_w + h, 1);
    }
    mbar_init(tmem_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 32);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // layer L of a step: 0 input_proj, 1 + 2i mlp.0 of block i, 2 + 2i mlp.2 of block i, 13 final (rank 0 only: 32 features)
  auto has_ln = [](int L) { return L == 0 || (L < FH_LAYERS - 1 && (L & 1) == 0); };
  const int n_layers_mine = (rank == 0) ? FH_LAYERS : FH_LAYERS - 1;
  // Global layer counter g = s * 14 + L: the operand image of layer g is act_s[g & 1], written by the epilogue of layer
  // g - 1 in every CTA of the cluster (the z rows of the very first layer by this CTA itself).  Ranks 1-3 skip the final
  // layer but still take part in its barrier, so all CTAs count the same layers.
  // Weights of a layer sit in two 64 KB halves (k-blocks 0-3 | 4-7), one TMA box and one full/empty barrier pair each, so
  // the MMA warp waits twice and commits twice per layer instead of once per k-block.  Half 0 is used by every layer of
  // this CTA, half 1 by the layers with eight k-blocks (all but input_proj): each half keeps its own load counter.

  // ===== TMA producer (warp 0, lane 0): the weights of the layer after next are requested as soon as the MMAs of this
  // layer have read theirs.  (sa, La): next layer whose half 0 is to be loaded; (sb, Lb): the same for half 1.
  uint32_t c0 = 0, c1 = 0;  // loads issued so far into half 0 / half 1
  int sa = 0, La = 0, sb = 0, Lb = 1;
  auto load_half0 = [&]() {  // lane 0 of warp 0 only
    if (sa >= S) return;
    if (c0 > 0) mbar_wait(empty_w, (c0 - 1) & 1);
    mbar_arrive_expect_tx(full_w, (La == 0 ? 1 : 4) * FH_W_BYTES);
    if (La == 0) tma_load_3d(w_s, &map_win, full_w, 0, rank * FH_FEATS, 0);
    else tma_load_3d(w_s, &map_wpack, full_w, 0, (La - 1) * FH_DIM + rank * FH_FEATS, 0);
    ++c0;
    if (++La == n_layers_mine) { La = 0; ++sa; }
  };
  auto load_half1 = [&]() {
    if (sb >= S) return;
    if (c1 > 0) mbar_wait(empty_w + 1, (c1 - 1) & 1);
    mbar_arrive_expect_tx(full_w + 1, 4 * FH_W_BYTES);
    tma_load_3d(w_s + 4 * FH_W_BYTES, &map_wpack, full_w + 1, 0, (Lb - 1) * FH_DIM + rank * FH_FEATS, 4);
    ++c1;
    if (++Lb == n_layers_mine) { Lb = 1; ++sb; }
  };
  if (warp == 0 && lane == 0) {  // constants: requested before the dependency on the previous kernel resolves
    load_half0();
    load_half1();
  }
  if (warp != 0) pdl_wait();  // everything below reads what earlier kernels of the stream produced (z, mod)
  cluster_sync_all();         // every CTA of the cluster is running before anyone writes into a peer's shared memory

  if (warp >= 2) {
    // the operand of the first input_proj: z rows [row0, row0 + 16) x 64 halves (cols 32..63 are zero), 2 KB
    const int etid = threadIdx.x - 64;
    if (etid < FH_ROWS * 8) {
      const int r = etid >> 3, c8 = etid & 7;
      const uint4 v = *reinterpret_cast<const uint4*>(p.z16 + static_cast<long long>(row0 + r) * 64 + c8 * 8);
      *reinterpret_cast<uint4*>(act_s + r * 128 + ((c8 ^ (r & 7)) << 4)) = v;
    }
    fence_proxy_async_all();
  }
  __syncthreads();

  if (warp == 0) {
    for (int s = 0; s < S; ++s)
      for (int L = 0; L < FH_LAYERS; ++L) {
        if (lane == 0 && L < n_layers_mine) {
          load_half0();             // waits for this layer's MMAs on half 0
          if (L >= 1) load_half1();
        }
        __syncwarp();
        if (s == S - 1 && L == FH_LAYERS - 1) break;
        if (has_ln(L)) cluster_sync_all();
        cluster_sync_all();
      }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = make_idesc_f16_m128(FH_ROWS);
    const uint64_t da0 = make_sw128_kmajor_desc(smem_u32(w_s));
    uint32_t c0 = 0, c1 = 0;
    int g = 0;
    for (int s = 0; s < S; ++s)
      for (int L = 0; L < FH_LAYERS; ++L, ++g) {
        if (L < n_layers_mine) {
          const uint64_t db0 = make_sw128_kmajor_desc(smem_u32(act_s + (g & 1) * FH_ACT_BYTES));
          mbar_wait(full_w, c0 & 1);
          ++c0;
          tc_fence_after();
          if (elect_one()) {
            fence_proxy_async_all();  // the operand image was written by generic-proxy stores of the cluster, ordered by its barrier
            const int nk0 = (L == 0) ? 1 : 4;
            for (int kb = 0; kb < nk0; ++kb) {
#pragma unroll
              for (int k = 0; k < 4; ++k)  // k-block tiles are 16 KB (weights) / 2 KB (rows) apart in the 16-byte address field
                umma_f16(tmem_base, da0 + kb * (FH_W_BYTES >> 4) + 2 * k, db0 + kb * (FH_ACT_KB >> 4) + 2 * k, idesc, (kb | k) != 0);
            }
            umma_commit(empty_w);
            if (L == 0) umma_commit(tmem_full);
          }
          __syncwarp();
          if (L >= 1) {
            mbar_wait(full_w + 1, c1 & 1);
            ++c1;
            tc_fence_after();
            if (elect_one()) {
              for (int kb = 4; kb < 8; ++kb) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_f16(tmem_base, da0 + kb * (FH_W_BYTES >> 4) + 2 * k, db0 + kb * (FH_ACT_KB >> 4) + 2 * k, idesc, 1);
              }
              umma_commit(empty_w + 1);
              umma_commit(tmem_full);
            }
            __syncwarp();
          }
        }
        if (s == S - 1 && L == FH_LAYERS - 1) break;
        if (has_ln(L)) cluster_sync_all();  // the LayerNorm statistics exchange
        cluster_sync_all();                 // the next operand image is complete in every CTA
      }
  } else {
    // ===== epilogue: thread = (feature, half of the rows), registers = 8 rows =====
    const int quad = warp & 3;                 // TMEM lanes 32*quad .. +31
    const int half = (warp - 2) >> 2;          // rows 8*half .. +7
    const int fl = quad * 32 + lane;           // feature within this CTA's slice
    const int f = rank * FH_FEATS + fl;        // feature of the 512-wide layer
    const int etid = threadIdx.x - 64;         // 0..255
    const int rl0 = half * FH_EROWS;           // first local row of this thread
    const int r0 = row0 + rl0;                 // first batch row of this thread
    const uint32_t tmem_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + rl0;
    float x[FH_EROWS];
#pragma unroll
    for (int r = 0; r < FH_EROWS; ++r) x[r] = 0.f;
    uint32_t stat_peer[FH_CLUSTER], act_peer[FH_CLUSTER];
#pragma unroll
    for (int k = 0; k < FH_CLUSTER; ++k) {
      stat_peer[k] = map_to_rank(smem_u32(stat_s), k);
      act_peer[k] = map_to_rank(smem_u32(act_s), k);
    }
    // rows past the batch (r0 + r >= n) live in the padding of every row-indexed buffer: computed, never consumed
    const float* const mean_h = mean_s + rl0;
    const float* const rstd_h = rstd_s + rl0;
    const uint32_t stat_off = ((rank * 4 + quad) * FH_ROWS + rl0) * 4;
    // operand stores: lanes pair up (features f, f ^ 1) so that a store carries two halves; the even lane of a pair
    // stores rows 0-3, the odd lane rows 4-7
    auto publish = [&](const float (&y)[FH_EROWS], int gnext) {
      const uint32_t buf = static_cast<uint32_t>(gnext & 1) * FH_ACT_BYTES;
#pragma unroll
      for (int r = 0; r < FH_EROWS; ++r) {
        const float other = __shfl_xor_sync(0xffffffffu, y[r], 1);
        if ((r >> 2) == (lane & 1)) {
          const __half2 h2 = (lane & 1) ? __floats2half2_rn(other, y[r]) : __floats2half2_rn(y[r], other);
          const uint32_t off = buf + fh_img_off(rl0 + r, f & ~1);
#pragma unroll
          for (int k = 0; k < FH_CLUSTER; ++k) st_dsmem_u32(act_peer[k] + off, *reinterpret_cast<const uint32_t*>(&h2));
        }
      }
    };

    int g = 0;
    uint32_t tf = 0;  // tmem_full completions consumed by this thread
    for (int s = 0; s < S; ++s) {
      const float* const mod_rows = p.mod + s * p.mod_step_stride + static_cast<long long>(r0) * FH_MOD_LD + f;
      for (int L = 0; L < FH_LAYERS; ++L, ++g) {
        const bool active = L < n_layers_mine;
        const bool last_of_all = (s == S - 1 && L == FH_LAYERS - 1);
#define FH_TRACE(slot) do { if (p.trace && etid == 0 && rank == 0 && blockIdx.x == 0 && s == S - 1) p.trace[L * 8 + (slot)] = gtime(); } while (0)
        FH_TRACE(0);
        if (s == S - 1 && L == FH_LAYERS - 2 && etid == 0) pdl_launch_dependents();
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
        float sc[FH_EROWS], sh[FH_EROWS];
        if (has_ln(L)) {
          // h = LN(x) [* w + b] * (1 + scale) + shift for the next block (j) or the final layer (no affine): the
          // modulation rows are in flight while the MMAs finish
          const float* shift = mod_rows + (L >> 1) * 3 * FH_DIM;
          const float* scale = shift + FH_DIM;
#pragma unroll
          for (int r = 0; r < FH_EROWS; ++r) {
            sc[r] = __ldg(scale + r * FH_MOD_LD);
            sh[r] = __ldg(shift + r * FH_MOD_LD);
          }
        }
        if (active) {
          mbar_wait(tmem_full, tf & 1);
          ++tf;
          tc_fence_after();
        }
        FH_TRACE(1);
        float y[FH_EROWS];
        if (active) {
          uint32_t v[FH_EROWS];
          tmem_ld8_wait(tmem_lane, v);
          tc_fence_before();
          if (L == 0) {
#pragma unroll
            for (int r = 0; r < FH_EROWS; ++r) x[r] = __uint_as_float(v[r]) * ws + bias;
          } else if (is_mlp2) {  // x += gate * (W g + b)
#pragma unroll
            for (int r = 0; r < FH_EROWS; ++r) x[r] += pre[r] * (__uint_as_float(v[r]) * ws + bias);
          } else if (is_mlp0) {  // g = silu(W h + b)
#pragma unroll
            for (int r = 0; r < FH_EROWS; ++r) y[r] = silu_fast(__uint_as_float(v[r]) * ws + bias);
          } else {  // Euler step: z += (W h + b) / S; rank 0, features 0..31 (the other lanes of the warp idle along)
#pragma unroll
            for (int r = 0; r < FH_EROWS; ++r) {
              const int row = r0 + r;
              const float zn = fl < FH_LDIM ? pre[r] + (__uint_as_float(v[r]) * ws + bias) * p.alpha : 0.f;
              y[r] = zn;
              if (fl < FH_LDIM && row < p.n) {
                p.z32[row * FH_LDIM + fl] = zn;
                p.z16[row * 64 + fl] = __float2half_rn(zn);
              }
            }
          }
        }
        if (last_of_all) break;
        if (L == FH_LAYERS - 2 && s == S - 1 && p.x_dbg) {
#pragma unroll
          for (int r = 0; r < FH_EROWS; ++r) p.x_dbg[static_cast<long long>(r0 + r) * FH_DIM + f] = x[r];
        }
        if (has_ln(L)) {
          // Row statistics over all 512 features in ONE exchange: every warp reduces its 32 features of each row to
          // (sum, M2 about its own mean), the 16 partials per row meet in every CTA's shared memory after one cluster
          // barrier and are merged with Chan's update, which is the two-pass variance up to rounding.
          const float e = warp_rows_sum8([&](int r) { return x[r]; }, lane);
          const float mloc = e * (1.f / 32.f);  // lanes 4r..4r+3: mean of row r over this warp's features
          const float e2 = warp_rows_sum8([&](int r) { const float d = x[r] - __shfl_sync(0xffffffffu, mloc, 4 * r); return d * d; }, lane);
          if ((lane & 3) == 0) {
            const uint32_t o = stat_off + (lane >> 2) * 4;
#pragma unroll
            for (int k = 0; k < FH_CLUSTER; ++k) {
              st_dsmem_f1(stat_peer[k] + o, e);
              st_dsmem_f1(stat_peer[k] + 16 * FH_ROWS * 4 + o, e2);
            }
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
          const int j = L >> 1;  // 0 after input_proj, i + 1 after block i
          const bool affine = j < FH_DEPTH;
          const float lw = affine ? __ldg(p.ln_w[affine ? j : 0] + f) : 1.f;
          const float lb = affine ? __ldg(p.ln_b[affine ? j : 0] + f) : 0.f;
#pragma unroll
          for (int r = 0; r < FH_EROWS; ++r) {
            const float t = (x[r] - mean_h[r]) * rstd_h[r] * lw + lb;
            y[r] = t * (1.f + sc[r]) + sh[r];
          }
        }
        // publish this layer's columns of the next operand image in every CTA of the cluster, then meet the cluster.
        // The final layer of a step hands z (32 features of rank 0, zero beyond) to the next step's input_proj: k-block 0.
        FH_TRACE(2);
        if (!is_final) publish(y, g + 1);
        else if (rank == 0 && fl < 64) publish(y, g + 1);
        fence_proxy_async_all();
        cluster_sync_all();
        FH_TRACE(3);
      }
    }
#undef FH_TRACE
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 32);
  // a peer may still be writing into this CTA's shared memory only before the last cluster barrier it took part in, and
  // every CTA leaves after that barrier: nothing to wait for here
}

}  // namespace ptts
