// One GEMM engine for every matrix product on the hot path.
//
//   D[r, f] = sum_k  A[r, k] * W[f, k]          (f16 operands, f32 accumulate in TMEM)
//
// A is an activation tensor kept channels-last as [stream][row][C]; a logical A row (stream b,
// row t) is the window of `taps` consecutive stored rows t..t+taps-1, so K = taps*C.  With
// taps = 1 this is a Linear layer; with taps = k it is the reference's causal StreamingConv1d
// (modules/conv.rs:90-136: [previous | x] then a valid conv) as an implicit GEMM whose left
// context rows simply sit in front of the body rows; with taps = 2 it is StreamingConvTranspose1d
// with k = 2*stride (modules/conv.rs:219-267): y[t*s+rho] = x[t] W[:,:,rho] + x[t-1] W[:,:,rho+s],
// i.e. N = s*Cout and the carried state is the previous input row instead of the partial sums.
// No im2col buffer exists: the K loop walks (tap, 64-channel block) and offsets the TMA row
// coordinate by the tap.
//
// Two operand placements share the kernel:
//   swap = 0  activations on the 128-row MMA-M side, BN features on MMA-N   (rows >= 128: Mimi, SEANet)
//   swap = 1  weights on the MMA-M side (128 features per tile), up to 256 activation rows on
//             MMA-N, so a small decode batch wastes no MMA rows          (FlowLM decode, flow head)
//
// Split-K (decode batches give only F/128 output tiles) runs as a thread-block cluster along z: every CTA of
// the cluster accumulates its K slice in TMEM, stages the partial tile in its own shared memory, and after a
// cluster barrier each CTA sums one interleaved set of rows over all peers through distributed shared memory
// in rank order (bit-reproducible) and stores only those rows.  No workspace, no atomics, and the epilogue
// work is spread over the whole cluster.
//
// Warp roles: warp 0 TMA producer, warp 1 TMEM owner + tcgen05.mma issuer, warps 2-5 TMEM -> smem staging (warps 2-9 first
// expand int8 weight tiles when the weights are stored as bytes),
// then all 12 warps store the tile.  Launched with programmatic dependent launch: everything before
// griddepcontrol.wait (barrier init, TMEM allocation, descriptor prefetch) overlaps the previous kernel's tail.
#pragma once
#include <type_traits>
#include "ptx.cuh"

namespace ptts {

// Address of logical element (r, f) in an output / residual tensor:
//   off = (r / T) * stream_stride + base + (r % T) * ld + f
// which covers plain [rows, F] (T = INT_MAX) and the padded per-stream conv buffers.
struct RowMap {
  int T;
  int ld;
  long long stream_stride;
  long long base;
};
inline RowMap plain_map(int ld) { return RowMap{0x7fffffff, ld, 0, 0}; }

enum { ACT_NONE = 0, ACT_GELU = 1, ACT_SILU = 2, ACT_ELU = 3 };

struct GemmEpi {
  const float* bias;    // [F] or null
  const float* fscale;  // [F] per-feature multiplier (LayerScale, modules/mlp.rs:71-73) or null
  const float* gate;    // per-(r,f) multiplier (flow head gate, modules/mlp.rs:174-176) or null
  RowMap gate_map;
  const float* res;     // f32 residual input or null
  RowMap res_map;
  float* out32;         // f32 output or null
  RowMap out32_map;
  __half* out16;        // f16 output (next GEMM's operand) or null
  RowMap out16_map;
  int act;              // applied to acc + bias
  int act16;            // applied to the f16 copy only (ELU in front of the next SEANet conv)
  float alpha;          // multiplies after the activation
  int reserved;         // inside GemmParams: GEMM_F_* switches of optional paths (kept here so that no field of the
                        // parameter block moved when they were added: 40 more bytes in its hot part measured ~0.5 us
                        // per launch, more constant-bank lines on the prologue's critical path)
  const float* wscale;  // [F] int8 mode: weight-code scale applied to the accumulator first, else null
};

// GEMM_F_W_INT8 (swap-AB only): map_w views one-byte weight codes; warps 2-9 expand each tile to f16 in shared memory.
enum { GEMM_F_W_INT8 = 1 };

struct GemmParams {
  int swap;
  int F, K;
  int n_streams, T, R, G;   // activation rows = n_streams*T; one tile = G streams x R rows
  int taps, cblocks;        // K blocks = taps * cblocks, cblocks = C/64
  int BN;                   // MMA N
  int kb_per_split;
  int stages;
  int tmem_cols;
  int vec4;                 // every epilogue tensor is 16-byte addressable in groups of 4 features
  int epi_mask;             // epi_mask_of(epi): selects the compiled store loop
  int n_act_tiles;          // activation tiles in total (persistent kernel walks them with stride gridDim.x)
  int pdl_trigger;          // where the CTA lets the next kernel launch: 0 entry, 1 all loads issued, 2 accumulator ready
  int resident;             // swap-AB decode GEMM whose whole K slice fits the stages: one barrier, one activation box
  GemmEpi epi;
  // raw view, used by the SIMT cross-check kernel only
  const __half* act;
  long long act_stream_stride;
  int act_ld;
  const __half* w;
  unsigned long long* trace;  // optional [grid][16] %globaltimer stamps (bring-up only)
};

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define PTTS_TRACE(slot)                                                                                           \
  do {                                                                                                             \
    if (p.trace && lane == 0)                                                                                      \
      p.trace[((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 + (slot)] = gtime();            \
  } while (0)

static constexpr int GEMM_BM = 128;
static constexpr int GEMM_BK = 64;
#ifndef PTTS_GEMM_THREADS
#define PTTS_GEMM_THREADS 384
#endif
#ifndef PTTS_GEMM_MINBLOCKS
#define PTTS_GEMM_MINBLOCKS 1
#endif
static constexpr int GEMM_THREADS = PTTS_GEMM_THREADS;  // 12 warps, one CTA per SM (8 warps x 2 CTAs per SM measured 7% slower: the store loops want warps)
static constexpr int GEMM_MAX_SPLIT = 8;  // portable cluster size

// Row part of a RowMap offset (everything except "+ f"); no integer division when the map is a plain
// [rows, F] matrix.
__device__ __forceinline__ long long row_off(const RowMap& m, int r) {
  if (r < m.T) return m.base + static_cast<long long>(r) * m.ld;
  const int b = r / m.T;
  return static_cast<long long>(b) * m.stream_stride + m.base + static_cast<long long>(r - b * m.T) * m.ld;
}

// Same offset when the row's (stream, row-in-stream) pair is already known: no division for a map that follows
// the GEMM's own stream structure.
__device__ __forceinline__ long long row_off_bt(const RowMap& m, int r, int b, int t, int T) {
  if (m.T == T) return static_cast<long long>(b) * m.stream_stride + m.base + static_cast<long long>(t) * m.ld;
  return row_off(m, r);
}
// ELU for a value that is about to be rounded to f16: exp(x) - 1 with the fast exponential (absolute error ~1e-7).
__device__ __forceinline__ float elu1_fast(float x) { return x > 0.f ? x : __expf(x) - 1.f; }

__device__ __forceinline__ float epi_act(int act, float v) {
  if (act == ACT_GELU) return gelu_tanh(v);
  if (act == ACT_SILU) return silu(v);
  if (act == ACT_ELU) return elu1(v);
  return v;
}

// Epilogue shape as a bit mask.  The kernel switches once (warp-uniformly) to a copy of the store loop compiled
// for exactly that shape, so the loop carries no predicated-off instructions for features it does not use;
// EPI_GENERIC keeps every test at run time and serves shapes outside the list.
enum {
  EPI_BIAS = 1, EPI_FSCALE = 2, EPI_GATE = 4, EPI_RES = 8, EPI_OUT32 = 16, EPI_OUT16 = 32, EPI_ELU16 = 64,
  EPI_ACT_SHIFT = 7 /* 2 bits */, EPI_ALPHA = 512, EPI_GENERIC = -1
};
__host__ __device__ inline int epi_mask_of(const GemmEpi& e) {
  return (e.bias ? EPI_BIAS : 0) | (e.fscale ? EPI_FSCALE : 0) | (e.gate ? EPI_GATE : 0) | (e.res ? EPI_RES : 0) |
         (e.out32 ? EPI_OUT32 : 0) | (e.out16 ? EPI_OUT16 : 0) | ((e.out16 && e.act16 == ACT_ELU) ? EPI_ELU16 : 0) |
         (e.act << EPI_ACT_SHIFT) | (e.alpha != 1.f ? EPI_ALPHA : 0);
}
// every shape the engine issues (engine.cu: FlowLM, flow head, Mimi transformer, SEANet)
#define PTTS_EPI_SHAPES(X)                                                                  \
  X(EPI_OUT32)                                                                              \
  X(EPI_BIAS | EPI_OUT32)                                                                   \
  X(EPI_RES | EPI_OUT32)                                                                    \
  X((ACT_GELU << EPI_ACT_SHIFT) | EPI_OUT16)                                                \
  X(EPI_BIAS | (ACT_SILU << EPI_ACT_SHIFT) | EPI_OUT16)                                     \
  X(EPI_BIAS | EPI_GATE | EPI_RES | EPI_OUT32)                                              \
  X(EPI_BIAS | EPI_ALPHA | EPI_RES | EPI_OUT32 | EPI_OUT16)                                 \
  X(EPI_FSCALE | EPI_RES | EPI_OUT32)                                                       \
  X(EPI_FSCALE | EPI_RES | EPI_OUT32 | EPI_OUT16)                                           \
  X(EPI_BIAS | EPI_OUT16 | EPI_ELU16)                                                       \
  X(EPI_BIAS | EPI_OUT32 | EPI_OUT16 | EPI_ELU16)                                           \
  X(EPI_BIAS | EPI_RES | EPI_OUT16 | EPI_ELU16)

__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t cluster_addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(cluster_addr));
  return v;
}
__device__ __forceinline__ float ld_dsmem_f1(uint32_t cluster_addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(cluster_addr));
  return v;
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t cta_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Execution-only rendezvous (no memory ordering): the release form drains every outstanding global store of the CTA
// (MEMBAR.ALL.GPU, ~0.8 us right after an epilogue), which a barrier that merely keeps shared memory alive for the
// peers' reads does not need.
__device__ __forceinline__ void cluster_sync_relaxed() {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}

// int8 weight storage (reference quantize.rs:65-94: per-tensor symmetric codes in [-127, 127]).  TMA drops the raw
// [128 features][64 k] byte tile (8 KB, dense) into the upper half of the 16 KB operand slot; the 256 threads of warps
// 2-9 (two per feature row) pull their 32 bytes into registers, meet at a named barrier (row r's f16 destination
// overlaps the raw bytes of rows 2r-128 and 2r-127), and write the row back as 64 halves in the SWIZZLE_128B K-major
// layout the MMA descriptor expects (16-byte chunk c of row r at r*128 + ((c ^ (r & 7)) << 4)).  Codes are exact in
// f16, so the accumulator is bit-identical to streaming an f16 copy of the codes; the scale stays in the epilogue.
__device__ __forceinline__ void i8x4_to_f16x4(uint32_t w, uint32_t& lo, uint32_t& hi) {
  // byte b -> half 0x6400 | (b ^ 0x80) = 1024 + (b + 128); minus 1152 gives b exactly
  const uint32_t x = w ^ 0x80808080u;
  const uint32_t a = __byte_perm(x, 0x64646464u, 0x4140);
  const uint32_t b = __byte_perm(x, 0x64646464u, 0x4342);
  const __half2 bias = __halves2half2(__ushort_as_half(0x6480), __ushort_as_half(0x6480));  // 1152.0
  const __half2 ra = __hsub2(*reinterpret_cast<const __half2*>(&a), bias);
  const __half2 rb = __hsub2(*reinterpret_cast<const __half2*>(&b), bias);
  lo = *reinterpret_cast<const uint32_t*>(&ra);
  hi = *reinterpret_cast<const uint32_t*>(&rb);
}
__device__ __forceinline__ void expand_i8_tile(uint8_t* tile, int row, int half) {
  const uint4* src = reinterpret_cast<const uint4*>(tile + GEMM_BM * GEMM_BK + row * GEMM_BK + half * 32);
  uint4 raw[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) raw[j] = src[j];
  asm volatile("bar.sync 2, 256;" ::: "memory");  // every raw row is in registers before any f16 row is written
  uint8_t* dst = tile + row * 128;
  const int sw = row & 7;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    uint4 o0, o1;
    i8x4_to_f16x4(raw[j].x, o0.x, o0.y);
    i8x4_to_f16x4(raw[j].y, o0.z, o0.w);
    i8x4_to_f16x4(raw[j].z, o1.x, o1.y);
    i8x4_to_f16x4(raw[j].w, o1.z, o1.w);
    const int c = half * 4 + 2 * j;
    *reinterpret_cast<uint4*>(dst + ((c ^ sw) << 4)) = o0;
    *reinterpret_cast<uint4*>(dst + (((c + 1) ^ sw) << 4)) = o1;
  }
}

// Second half of the epilogue.  The f32 accumulator tile sits in shared memory as [activation row][feature]
// (pitch LD floats) in every CTA of the split-K cluster; CTA `rank` of `nsplit` owns rows rank*nwarps + warp,
// stepping by nsplit*nwarps, sums them over the peers in rank order, applies the epilogue and writes V
// consecutive features per lane, so each warp access is one contiguous 128-byte (V=1) or 512-byte (V=4) run.
// Must inline into the kernel: `p` then stays in the constant bank.  Out of line, the by-reference parameter block
// lives in local memory, and with a 200 KB shared-memory carve-out the remaining L1 cannot hold 384 threads' copies,
// so every field access became an L2 round trip (measured: 6-11 us per 128-row tile instead of ~1 us).
// Tile-invariant part of the fast path below (lane -> (row slot, feature quad), per-feature vectors).  The persistent
// kernel computes it once per CTA: done per tile it was ~260 instructions and a dependent global load (the bias) at the
// head of every 128-row tile, as much work as the store loop itself (ncu: 2.5 M of 6.9 M warp instructions).
struct EpiHoist {
  int my_sub, q, rows_per_iter, f;
  float bv[4], sv[4], wv[4];
};
__device__ __forceinline__ void epi_hoist_init(const GemmParams& p, int f0, int tid, EpiHoist& h) {
  const GemmEpi& e = p.epi;
  const int fv = (p.swap ? GEMM_BM : p.BN) / 4;
  const int lane = tid & 31;
  const int lanes_per_row = fv <= 32 ? fv : 32;
  h.rows_per_iter = 32 / lanes_per_row;
  h.my_sub = lane / lanes_per_row;
  h.q = lane - h.my_sub * lanes_per_row;
  h.f = f0 + h.q * 4;
#pragma unroll
  for (int c = 0; c < 4; ++c) { h.bv[c] = 0.f; h.sv[c] = 1.f; h.wv[c] = 1.f; }
  if (h.my_sub < h.rows_per_iter && h.f < p.F) {
    if (e.wscale) { const float4 t4 = __ldg(reinterpret_cast<const float4*>(e.wscale + h.f)); h.wv[0] = t4.x; h.wv[1] = t4.y; h.wv[2] = t4.z; h.wv[3] = t4.w; }
    if (e.bias) { const float4 t4 = __ldg(reinterpret_cast<const float4*>(e.bias + h.f)); h.bv[0] = t4.x; h.bv[1] = t4.y; h.bv[2] = t4.z; h.bv[3] = t4.w; }
    if (e.fscale) { const float4 t4 = __ldg(reinterpret_cast<const float4*>(e.fscale + h.f)); h.sv[0] = t4.x; h.sv[1] = t4.y; h.sv[2] = t4.z; h.sv[3] = t4.w; }
  }
}

template <int V, int M>
__device__ __forceinline__ void epi_store_tile(const GemmParams& p, uint32_t stile_addr, int LD, int f0, int t0, int b0,
                                            int tid, int nthreads, int rank, int nsplit, const EpiHoist* hp = nullptr) {
  constexpr bool GEN = (M == EPI_GENERIC);
  const GemmEpi& e = p.epi;
  const bool has_bias = GEN ? e.bias != nullptr : (M & EPI_BIAS) != 0;
  const bool has_fscale = GEN ? e.fscale != nullptr : (M & EPI_FSCALE) != 0;
  const bool has_gate = GEN ? e.gate != nullptr : (M & EPI_GATE) != 0;
  const bool has_res = GEN ? e.res != nullptr : (M & EPI_RES) != 0;
  const bool has_o32 = GEN ? e.out32 != nullptr : (M & EPI_OUT32) != 0;
  const bool has_o16 = GEN ? e.out16 != nullptr : (M & EPI_OUT16) != 0;
  const bool elu16 = GEN ? e.act16 == ACT_ELU : (M & EPI_ELU16) != 0;
  const int act = GEN ? e.act : ((M >> EPI_ACT_SHIFT) & 3);
  const float alpha = (GEN || (M & EPI_ALPHA)) ? e.alpha : 1.f;
  const float* __restrict__ bias = e.bias;
  const float* __restrict__ fscale = e.fscale;
  const float* __restrict__ wscale = e.wscale;  // int8 mode only; a run-time test in every shape (uniform, one FMUL)
  const int swap = p.swap, F = p.F, T = p.T, R = p.R, G = p.G, n_streams = p.n_streams;
  const int tile_rows = swap ? p.BN : GEMM_BM;                 // activation rows covered by the tile
  const int fv = (swap ? GEMM_BM : p.BN) / V;                  // feature groups per activation row
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
  // ---- fast path: every row of the tile in one stream, so each tensor's row
  // offset is affine in the tile row; with <= 32 feature groups per row a lane keeps the same features for every
  // row it visits, so bias / LayerScale are loaded once.  This is the path of the persistent SEANet kernels.
  if (V == 4 && (swap || G == 1) && fv <= 32) {
    auto affine = [&](const RowMap& m, long long& off0) {  // offset of tile row 0; rows advance by m.ld
      if (m.T == T && !swap) { off0 = static_cast<long long>(b0) * m.stream_stride + m.base + static_cast<long long>(t0) * m.ld; return true; }
      if (m.T == 0x7fffffff) { off0 = m.base + static_cast<long long>(swap ? t0 : b0 * T + t0) * m.ld; return true; }
      return false;
    };
    long long g0 = 0, r0 = 0, a0 = 0, h0 = 0;
    const bool ok = (!has_gate || affine(e.gate_map, g0)) && (!has_res || affine(e.res_map, r0)) &&
                    (!has_o32 || affine(e.out32_map, a0)) && (!has_o16 || affine(e.out16_map, h0));
    if (ok) {
      int nrows = min(tile_rows, T - t0);
      if (!swap) nrows = min(nrows, R);
      EpiHoist hl;
      if (!hp) epi_hoist_init(p, f0, tid, hl);  // one tile per CTA: nothing to hoist over
      const EpiHoist& h = hp ? *hp : hl;
      const int rows_per_iter = h.rows_per_iter, my_sub = h.my_sub, q = h.q, f = h.f;
      if (my_sub < rows_per_iter && f < F) {
        const float bv[4] = {h.bv[0], h.bv[1], h.bv[2], h.bv[3]}, sv[4] = {h.sv[0], h.sv[1], h.sv[2], h.sv[3]};
        const float wv[4] = {h.wv[0], h.wv[1], h.wv[2], h.wv[3]};
        const int ld_g = e.gate_map.ld, ld_r = e.res_map.ld, ld_a = e.out32_map.ld, ld_h = e.out16_map.ld;
        const float* gate_p = has_gate ? e.gate + g0 + f : nullptr;
        const float* res_p = has_res ? e.res + r0 + f : nullptr;
        float* o32_p = has_o32 ? e.out32 + a0 + f : nullptr;
        __half* o16_p = has_o16 ? e.out16 + h0 + f : nullptr;
        uint32_t peer[GEMM_MAX_SPLIT];
#pragma unroll
        for (int k = 0; k < GEMM_MAX_SPLIT; ++k) peer[k] = (nsplit > 1 && k < nsplit) ? map_to_rank(stile_addr, k) : stile_addr;
        const int step = nwarps * nsplit * rows_per_iter;
        // Rows are taken U at a time, branch-free: every load of the U rows (residual, gate, staged accumulators) is issued
        // first, then the arithmetic of all rows interleaves, then the stores (predicated on row < nrows).  One row per
        // trip was one full memory round trip per row (a load may not move above the previous row's store to a possibly
        // aliasing tensor) and, with 2-3 warps per scheduler, one long dependent chain with nothing to overlap it.
        // U = 4 for whole tiles (persistent / unsplit kernels); a split-K cluster hands each warp at most a row or two, where
        // unrolling would only add clamped duplicate DSMEM reads; the catch-all shape has no registers to spare.
        auto run = [&](auto uc) {
        constexpr int U = decltype(uc)::value;
        for (int row_b = (rank * nwarps + warp) * rows_per_iter + my_sub; row_b < nrows; row_b += step * U) {
          float4 g4[U], r4[U], a4[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int row = min(row_b + u * step, nrows - 1);  // clamped: rows past the tile are computed and dropped
            if (has_gate) g4[u] = *reinterpret_cast<const float4*>(gate_p + static_cast<long long>(row) * ld_g);
            if (has_res) r4[u] = *reinterpret_cast<const float4*>(res_p + static_cast<long long>(row) * ld_r);
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int row = min(row_b + u * step, nrows - 1);
            const uint32_t toff = static_cast<uint32_t>(row * LD + q * 4) * 4u;
            if (nsplit == 1) {
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a4[u].x), "=f"(a4[u].y), "=f"(a4[u].z), "=f"(a4[u].w) : "r"(stile_addr + toff));
            } else {  // split-K: sum the cluster's partial tiles in rank order (bit-reproducible)
              float4 t[GEMM_MAX_SPLIT];
#pragma unroll
              for (int k = 0; k < GEMM_MAX_SPLIT; ++k)
                if (k < nsplit) t[k] = ld_dsmem_f4(peer[k] + toff);
              a4[u] = t[0];
#pragma unroll
              for (int k = 1; k < GEMM_MAX_SPLIT; ++k)
                if (k < nsplit) { a4[u].x += t[k].x; a4[u].y += t[k].y; a4[u].z += t[k].z; a4[u].w += t[k].w; }
            }
          }
          float v[U][4];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const float av[4] = {a4[u].x, a4[u].y, a4[u].z, a4[u].w};
            const float gv[4] = {g4[u].x, g4[u].y, g4[u].z, g4[u].w};
            const float rv[4] = {r4[u].x, r4[u].y, r4[u].z, r4[u].w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              float x = av[c];
              x *= wv[c];  // int8 weight-code scale, 1 otherwise
              if (has_bias) x += bv[c];
              x = epi_act(act, x) * alpha;
              if (has_fscale) x *= sv[c];
              if (has_gate) x *= gv[c];
              if (has_res) x += rv[c];
              v[u][c] = x;
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int row = row_b + u * step;
            if (row < nrows) {
              if (has_o32) *reinterpret_cast<float4*>(o32_p + static_cast<long long>(row) * ld_a) = make_float4(v[u][0], v[u][1], v[u][2], v[u][3]);
              if (has_o16) {
                float w4[4] = {v[u][0], v[u][1], v[u][2], v[u][3]};
                if (elu16) {
#pragma unroll
                  for (int c = 0; c < 4; ++c) w4[c] = elu1_fast(w4[c]);
                }
                const __half2 h0v = __floats2half2_rn(w4[0], w4[1]), h1v = __floats2half2_rn(w4[2], w4[3]);
                uint2 pk;
                pk.x = *reinterpret_cast<const uint32_t*>(&h0v);
                pk.y = *reinterpret_cast<const uint32_t*>(&h1v);
                *reinterpret_cast<uint2*>(o16_p + static_cast<long long>(row) * ld_h) = pk;
              }
            }
          }
        }
        };
        if (GEN || nsplit > 1) run(std::integral_constant<int, 1>{});
        else run(std::integral_constant<int, 4>{});
      }
      return;
    }
  }
  uint32_t peer[GEMM_MAX_SPLIT];
#pragma unroll
  for (int k = 0; k < GEMM_MAX_SPLIT; ++k) peer[k] = (nsplit > 1 && k < nsplit) ? map_to_rank(stile_addr, k) : stile_addr;
  // a narrow tile (fv < 32 feature groups) puts several rows in one warp iteration so no lane idles
  const int lanes_per_row = fv < 32 ? fv : 32;          // fv is a multiple of 4 and <= 64
  const int rows_per_iter = 32 / lanes_per_row;
  const int my_sub = lane / lanes_per_row, my_q0 = lane - my_sub * lanes_per_row;
  for (int row0 = (rank * nwarps + warp) * rows_per_iter; row0 < tile_rows; row0 += nwarps * nsplit * rows_per_iter) {
    const int row = row0 + my_sub;
    if (row >= tile_rows || my_sub >= rows_per_iter) continue;
    int r, b, t;
    if (swap) {
      r = t0 + row; b = 0; t = r;
      if (r >= T) continue;
    } else {
      int g = 0, tt = row;
      if (G > 1) { g = row / R; tt = row - g * R; }
      b = b0 + g; t = t0 + tt;
      if (g >= G || tt >= R || b >= n_streams || t >= T) continue;
      r = b * T + t;
    }
    const float* gate_r = has_gate ? e.gate + row_off_bt(e.gate_map, r, b, t, T) : nullptr;
    const float* res_r = has_res ? e.res + row_off_bt(e.res_map, r, b, t, T) : nullptr;
    float* o32_r = has_o32 ? e.out32 + row_off_bt(e.out32_map, r, b, t, T) : nullptr;
    __half* o16_r = has_o16 ? e.out16 + row_off_bt(e.out16_map, r, b, t, T) : nullptr;
    for (int q = my_q0; q < fv; q += lanes_per_row) {
      const int f = f0 + q * V;
      if (f >= F) break;
      const uint32_t toff = static_cast<uint32_t>(row * LD + q * V) * 4u;
      float v[V], gv[V], rv[V], bv[V], sv[V];
      if (V == 4) {
        float4 a4;
        if (nsplit == 1) {
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a4.x), "=f"(a4.y), "=f"(a4.z), "=f"(a4.w) : "r"(stile_addr + toff));
        } else {
          float4 t[GEMM_MAX_SPLIT];
#pragma unroll
          for (int k = 0; k < GEMM_MAX_SPLIT; ++k)
            if (k < nsplit) t[k] = ld_dsmem_f4(peer[k] + toff);
          a4 = t[0];
#pragma unroll
          for (int k = 1; k < GEMM_MAX_SPLIT; ++k)
            if (k < nsplit) { a4.x += t[k].x; a4.y += t[k].y; a4.z += t[k].z; a4.w += t[k].w; }
        }
        v[0] = a4.x; v[1] = a4.y; v[2] = a4.z; v[3] = a4.w;
        if (has_gate) { const float4 t4 = *reinterpret_cast<const float4*>(gate_r + f); gv[0] = t4.x; gv[1] = t4.y; gv[2] = t4.z; gv[3] = t4.w; }
        if (has_res) { const float4 t4 = *reinterpret_cast<const float4*>(res_r + f); rv[0] = t4.x; rv[1] = t4.y; rv[2] = t4.z; rv[3] = t4.w; }
        if (has_bias) { const float4 t4 = __ldg(reinterpret_cast<const float4*>(bias + f)); bv[0] = t4.x; bv[1] = t4.y; bv[2] = t4.z; bv[3] = t4.w; }
        if (has_fscale) { const float4 t4 = __ldg(reinterpret_cast<const float4*>(fscale + f)); sv[0] = t4.x; sv[1] = t4.y; sv[2] = t4.z; sv[3] = t4.w; }
      } else {
        float a = 0.f;
        for (int k = 0; k < nsplit; ++k) a += ld_dsmem_f1(peer[k] + toff);
        v[0] = a;
        if (has_gate) gv[0] = gate_r[f];
        if (has_res) rv[0] = res_r[f];
        if (has_bias) bv[0] = __ldg(bias + f);
        if (has_fscale) sv[0] = __ldg(fscale + f);
      }
#pragma unroll
      for (int c = 0; c < V; ++c) {
        float x = v[c];
        if (wscale) x *= __ldg(wscale + f + c);
        if (has_bias) x += bv[c];
        x = epi_act(act, x) * alpha;
        if (has_fscale) x *= sv[c];
        if (has_gate) x *= gv[c];
        if (has_res) x += rv[c];
        v[c] = x;
      }
      if (has_o32) {
        if (V == 4) *reinterpret_cast<float4*>(o32_r + f) = make_float4(v[0], v[1], v[2], v[3]);
        else o32_r[f] = v[0];
      }
      if (has_o16) {
        if (elu16) {
#pragma unroll
          for (int c = 0; c < V; ++c) v[c] = elu1_fast(v[c]);
        }
        if (V == 4) {
          const __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
          uint2 pk;
          pk.x = *reinterpret_cast<const uint32_t*>(&h0);
          pk.y = *reinterpret_cast<const uint32_t*>(&h1);
          *reinterpret_cast<uint2*>(o16_r + f) = pk;
        } else {
          o16_r[f] = __float2half_rn(v[0]);
        }
      }
    }
  }
}

__device__ __forceinline__ void epi_dispatch(const GemmParams& p, uint32_t stile_addr, int LD, int f0, int t0, int b0, int tid,
                                             int nthreads, int rank, int nsplit, const EpiHoist* hp = nullptr) {
  if (!p.vec4) {
    epi_store_tile<1, EPI_GENERIC>(p, stile_addr, LD, f0, t0, b0, tid, nthreads, rank, nsplit);
    return;
  }
  switch (p.epi_mask) {
#define PTTS_EPI_CASE(MASK) \
  case (MASK): epi_store_tile<4, (MASK)>(p, stile_addr, LD, f0, t0, b0, tid, nthreads, rank, nsplit, hp); break;
    PTTS_EPI_SHAPES(PTTS_EPI_CASE)
#undef PTTS_EPI_CASE
    default: epi_store_tile<4, EPI_GENERIC>(p, stile_addr, LD, f0, t0, b0, tid, nthreads, rank, nsplit, hp); break;
  }
}

__global__ void __launch_bounds__(GEMM_THREADS, PTTS_GEMM_MINBLOCKS)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_act, const __grid_constant__ CUtensorMap map_w,
               const GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x (M tile 16 KB | N tile BN*128 B)] then barriers
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int n_tile_bytes = p.BN * GEMM_BK * 2;
  const int m_tile_bytes = GEMM_BM * GEMM_BK * 2;
  const int stage_bytes = m_tile_bytes + n_tile_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full_bar = empty_bar + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  uint64_t* wfull_bar = tmem_full_bar + 3;      // int8 storage: raw weight bytes landed (per stage)
  uint64_t* conv_bar = wfull_bar + p.stages;    // int8 storage: f16 operand written by the eight converter warps

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (p.pdl_trigger == 0) pdl_launch_dependents();
  if (warp == 0) PTTS_TRACE(0);
  const bool w_int8 = (p.epi.reserved & GEMM_F_W_INT8) != 0;

  // tile coordinates
  const int tiles_t = (p.T + p.R - 1) / p.R;
  int act_tile, f0;
  if (p.swap) {
    f0 = blockIdx.x * GEMM_BM;
    act_tile = blockIdx.y;
  } else {
    act_tile = blockIdx.x;
    f0 = blockIdx.y * p.BN;
  }
  const int tb = act_tile / tiles_t;
  const int b0 = tb * p.G;
  const int t0 = (act_tile - tb * tiles_t) * p.R;
  const int total_kb = p.taps * p.cblocks;
  const int nsplit = gridDim.z;      // the cluster spans z: rank == blockIdx.z
  const int rank = blockIdx.z;
  const int kb0 = rank * p.kb_per_split;
  const int kb1 = min(total_kb, kb0 + p.kb_per_split);
  const int nkb = kb1 - kb0;         // >= 1 by construction of the split
  // per-feature epilogue constants (bias, LayerScale, int8 scale): requested now, not behind the accumulator -- as the
  // first thing of the store loop they were one L2 round trip on the critical path of every launch
  EpiHoist hoist;
  const bool hoisted = p.vec4 && (p.epi.bias || p.epi.fscale || p.epi.wscale);
  if (hoisted) epi_hoist_init(p, f0, threadIdx.x, hoist);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_act);
    tma_prefetch_desc(&map_w);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    mbar_init(tmem_full_bar, 1);
    if (w_int8) {
      for (int s = 0; s < p.stages; ++s) {
        mbar_init(wfull_bar + s, 1);
        mbar_init(conv_bar + s, 8);
      }
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) PTTS_TRACE(1);
  constexpr int RAW_OFF = GEMM_BM * GEMM_BK;      // raw byte tile sits in the upper half of its f16 slot
  constexpr uint32_t RAW_BYTES = GEMM_BM * GEMM_BK;

  if (warp == 0) {
    // ===== TMA producer =====
    // Weights are constants, so their tiles for the first stages are requested BEFORE griddepcontrol.wait: the HBM
    // fetch of this GEMM's weights overlaps the tail of the kernel it depends on, and only the (L2-resident)
    // activation tiles remain on the critical path once the dependency resolves.
    if (p.resident) {
      // Decode shape: the CTA's whole K slice is resident, so everything lands on ONE barrier: the weight tiles
      // [k-block][128][64] (requested before the dependency resolves), then a single activation box
      // (64 k, BN rows, all k-blocks) = [k-block][row][64] behind them.  One TMA instead of one per k-block: each
      // issue costs this thread ~0.3 us, which at eight k-blocks was the longest phase of the kernel.
      if (elect_one()) {
        if (w_int8) {
          mbar_arrive_expect_tx(wfull_bar, static_cast<uint32_t>(nkb) * RAW_BYTES);
          for (int i = 0; i < nkb; ++i) tma_load_3d(smem + i * m_tile_bytes + RAW_OFF, &map_w, wfull_bar, (kb0 + i) * GEMM_BK, f0, 0);
          mbar_arrive_expect_tx(full_bar, static_cast<uint32_t>(p.kb_per_split * n_tile_bytes));
        } else {
          mbar_arrive_expect_tx(full_bar, static_cast<uint32_t>(nkb * m_tile_bytes + p.kb_per_split * n_tile_bytes));
          for (int i = 0; i < nkb; ++i) tma_load_3d(smem + i * m_tile_bytes, &map_w, full_bar, (kb0 + i) * GEMM_BK, f0, 0);
        }
        pdl_wait();
        tma_load_3d(smem + p.kb_per_split * m_tile_bytes, &map_act, full_bar, 0, t0, kb0);
        PTTS_TRACE(2);
        if (p.pdl_trigger == 1) pdl_launch_dependents();
        PTTS_TRACE(3);
      }
    } else if (elect_one()) {
      const uint32_t act_bytes = p.swap ? n_tile_bytes : (GEMM_BK * p.R * p.G * 2);
      const uint32_t w_bytes = p.swap ? m_tile_bytes : n_tile_bytes;
      const int npre = min(nkb, p.stages);
      for (int i = 0; i < npre; ++i) {
        uint8_t* m_tile = smem + i * stage_bytes;
        if (w_int8) {  // swap only: weights on their own barrier so the expansion can start before the activations exist
          mbar_arrive_expect_tx(wfull_bar + i, RAW_BYTES);
          tma_load_3d(m_tile + RAW_OFF, &map_w, wfull_bar + i, (kb0 + i) * GEMM_BK, f0, 0);
          mbar_arrive_expect_tx(full_bar + i, act_bytes);
          continue;
        }
        mbar_arrive_expect_tx(full_bar + i, act_bytes + w_bytes);
        tma_load_3d(p.swap ? m_tile : m_tile + m_tile_bytes, &map_w, full_bar + i, (kb0 + i) * GEMM_BK, f0, 0);
      }
      pdl_wait();
      // stage, phase, tap and channel block advance by increments: the div/mod forms of these indices made every
      // iteration a ~1000-cycle dependent chain on this one thread, i.e. one activation tile requested per 0.5 us
      int s = 0, tap = kb0 / p.cblocks, cb = kb0 - tap * p.cblocks;
      uint32_t ph = 0;
      for (int i = 0; i < nkb; ++i) {
        uint8_t* m_tile = smem + s * stage_bytes;
        uint8_t* n_tile = m_tile + m_tile_bytes;
        if (i >= npre) {
          mbar_wait(empty_bar + s, ph ^ 1);
          if (w_int8) {
            mbar_arrive_expect_tx(wfull_bar + s, RAW_BYTES);
            tma_load_3d(m_tile + RAW_OFF, &map_w, wfull_bar + s, (kb0 + i) * GEMM_BK, f0, 0);
            mbar_arrive_expect_tx(full_bar + s, act_bytes);
          } else {
            mbar_arrive_expect_tx(full_bar + s, act_bytes + w_bytes);
            tma_load_3d(p.swap ? m_tile : n_tile, &map_w, full_bar + s, (kb0 + i) * GEMM_BK, f0, 0);
          }
        }
        tma_load_3d(p.swap ? n_tile : m_tile, &map_act, full_bar + s, cb * GEMM_BK, t0 + tap, b0);
        if (i == 0) PTTS_TRACE(2);
        if (++cb == p.cblocks) { cb = 0; ++tap; }
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
      if (p.pdl_trigger == 1) pdl_launch_dependents();
      PTTS_TRACE(3);
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one elected thread runs the whole loop =====
    // Per k-block: one barrier wait, four MMAs, one commit.  (Electing a lane, fencing and re-converging the warp in
    // every iteration cost ~600 cycles of single-thread latency per k-block, more than the MMAs themselves.)
    const uint32_t idesc = make_idesc_f16_m128(p.BN);
    if (p.resident) {
      if (elect_one()) {
        const uint64_t da0 = make_sw128_kmajor_desc(smem_u32(smem));
        const uint64_t db0 = make_sw128_kmajor_desc(smem_u32(smem + p.kb_per_split * m_tile_bytes));
        const uint32_t a_adv = static_cast<uint32_t>(m_tile_bytes) >> 4, b_adv = static_cast<uint32_t>(n_tile_bytes) >> 4;
        if (w_int8) mbar_wait(conv_bar, 0);
        mbar_wait(full_bar, 0);
        tc_fence_after();
        PTTS_TRACE(4);
        for (int i = 0; i < nkb; ++i) {
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k)
            umma_f16(tmem_base, da0 + static_cast<uint64_t>(i) * a_adv + 2 * k, db0 + static_cast<uint64_t>(i) * b_adv + 2 * k, idesc, (i | k) != 0);
        }
        umma_commit(tmem_full_bar);
      }
    } else if (elect_one()) {
      const uint64_t d0 = make_sw128_kmajor_desc(smem_u32(smem));
      const uint32_t stage_adv = static_cast<uint32_t>(stage_bytes) >> 4, b_adv = static_cast<uint32_t>(m_tile_bytes) >> 4;
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < nkb; ++i) {
        if (w_int8) mbar_wait(conv_bar + s, ph);
        mbar_wait(full_bar + s, ph);
        tc_fence_after();
        if (i == 0) PTTS_TRACE(4);
        const uint64_t da = d0 + static_cast<uint64_t>(s) * stage_adv;
        const uint64_t db = da + b_adv;
#pragma unroll
        for (int k = 0; k < GEMM_BK / 16; ++k) {
          // +32 B along K inside the 128-B swizzle row = +2 in the 16-byte address field
          umma_f16(tmem_base, da + 2 * k, db + 2 * k, idesc, (i | k) != 0);
        }
        umma_commit(empty_bar + s);  // frees this smem stage once the MMAs above have read it
        if (i == nkb - 1) umma_commit(tmem_full_bar);
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
    }
    __syncwarp();
    PTTS_TRACE(5);
  } else {
    if (w_int8 && warp < 10) {
      // ===== int8 storage: expand the raw weight tiles to the f16 MMA operand (see expand_i8_tile) =====
      const int idx = (warp - 2) * 32 + lane;
      const int row = idx >> 1, half = idx & 1;
      if (p.resident) {
        mbar_wait(wfull_bar, 0);
        for (int i = 0; i < nkb; ++i) expand_i8_tile(smem + i * m_tile_bytes, row, half);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tcgen05.mma reads
        __syncwarp();
        if (lane == 0) mbar_arrive(conv_bar);
      } else {
        int s = 0;
        uint32_t ph = 0;
        for (int i = 0; i < nkb; ++i) {
          mbar_wait(wfull_bar + s, ph);
          expand_i8_tile(smem + s * stage_bytes, row, half);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(conv_bar + s);
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    }
    if (warp < 6) {
    // ===== epilogue, first half: TMEM -> registers -> smem tile [activation row][feature] (raw f32) =====
    // The pipeline stages are dead once tmem_full has arrived (every MMA has consumed its operands), so the tile
    // is staged over them.  Pitch = features + 4 floats: 16-byte aligned rows, conflict-free in both passes.
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    if (p.pdl_trigger == 2 && warp == 2 && lane == 0) pdl_launch_dependents();
    if (warp == 2) PTTS_TRACE(6);
    const int quad = warp & 3;  // a warp may only touch TMEM lanes 32*(warp%4)..+31
    const int i = quad * 32 + lane;
    float* stile = reinterpret_cast<float*>(smem);
    const int LD = (p.swap ? GEMM_BM : p.BN) + 4;
    for (int c = 0; c < p.BN; c += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + c, v);
      tmem_ld_wait();
      if (p.swap) {
        // TMEM lane = feature, column = activation row: lanes write consecutive floats of row c+j
#pragma unroll
        for (int j = 0; j < 16; ++j) stile[(c + j) * LD + i] = __uint_as_float(v[j]);
      } else {
        // TMEM lane = activation row, columns = features: four 16-byte stores into row i
        float4* dst = reinterpret_cast<float4*>(stile + i * LD + c);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                               __uint_as_float(v[4 * j + 3]));
      }
    }
    if (warp == 2) PTTS_TRACE(7);
    }
  }
  // ===== epilogue, second half: every warp of every CTA of the split-K cluster =====
  tc_fence_before();
  __syncthreads();
  if (nsplit > 1) cluster_sync_all();  // all partial tiles are staged and visible cluster-wide
  pdl_wait();  // residual / gate / output tensors belong to earlier kernels until they have completed
  epi_dispatch(p, smem_u32(smem), (p.swap ? GEMM_BM : p.BN) + 4, f0, t0, b0, threadIdx.x, GEMM_THREADS, rank, nsplit, hoisted ? &hoist : nullptr);
  if (warp == 2) PTTS_TRACE(8);
  if (nsplit > 1) cluster_sync_relaxed();  // peers may still be reading this CTA's tile
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
  if (warp == 1) PTTS_TRACE(9);
}


// Persistent form for streaming GEMMs with hundreds of activation tiles and few features (the wide end of SEANet:
// 1920 rows per stream, 64 channels).  One CTA per SM walks activation tiles with stride gridDim.x; the TMA/MMA
// pipeline runs ahead across tile boundaries and the accumulator is double-buffered in TMEM, so the epilogue of
// tile j (10 warps: 4 TMEM readers stage the tile in shared memory, all 10 store it) overlaps the MMAs of tile j+1,
// and barrier setup / TMEM allocation are paid once per SM instead of once per tile.  Activations on MMA-M only.
__global__ void __launch_bounds__(GEMM_THREADS, PTTS_GEMM_MINBLOCKS)
gemm_tc_persistent_kernel(const __grid_constant__ CUtensorMap map_act, const __grid_constant__ CUtensorMap map_w,
                          const GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int n_tile_bytes = p.BN * GEMM_BK * 2;
  const int m_tile_bytes = GEMM_BM * GEMM_BK * 2;
  const int stage_bytes = m_tile_bytes + n_tile_bytes;
  const int LD = p.BN + 4;
  float* stile = reinterpret_cast<float*>(smem + p.stages * stage_bytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stile) + GEMM_BM * LD * 4);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tfull_bar = empty_bar + p.stages;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (p.pdl_trigger == 0) pdl_launch_dependents();
  const int tiles_t = (p.T + p.R - 1) / p.R;
  const int f0 = blockIdx.y * p.BN;
  const int nkb = p.taps * p.cblocks;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_act);
    tma_prefetch_desc(&map_w);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar + a, 1);
      mbar_init(tempty_bar + a, 4);  // one arrival per TMEM-reader warp
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t bytes = GEMM_BK * p.R * p.G * 2 + n_tile_bytes;
      // the weight tiles of the first stages do not depend on the previous kernel: request them before the wait
      const int total_it = ((p.n_act_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)) * nkb;
      const int npre = min(total_it, p.stages);
      for (int i = 0; i < npre; ++i) {
        mbar_arrive_expect_tx(full_bar + i, bytes);
        tma_load_3d(smem + i * stage_bytes + m_tile_bytes, &map_w, full_bar + i, (i % nkb) * GEMM_BK, f0, 0);
      }
      pdl_wait();
      int it = 0, s = 0;
      uint32_t ph = 0;
      // (stream, first row) of the tile advance by increments as well: no division on this thread's critical path
      int tb = static_cast<int>(blockIdx.x) / tiles_t, tt = static_cast<int>(blockIdx.x) - tb * tiles_t;
      const int adv_b = static_cast<int>(gridDim.x) / tiles_t, adv_t = static_cast<int>(gridDim.x) - adv_b * tiles_t;
      for (int tile = blockIdx.x; tile < p.n_act_tiles; tile += gridDim.x) {
        const int b0 = tb * p.G, t0 = tt * p.R;
        int tap = 0, cb = 0;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          uint8_t* m_tile = smem + s * stage_bytes;
          if (it >= npre) {
            mbar_wait(empty_bar + s, ph ^ 1);
            mbar_arrive_expect_tx(full_bar + s, bytes);
            tma_load_3d(m_tile + m_tile_bytes, &map_w, full_bar + s, kb * GEMM_BK, f0, 0);
          }
          tma_load_3d(m_tile, &map_act, full_bar + s, cb * GEMM_BK, t0 + tap, b0);
          if (++cb == p.cblocks) { cb = 0; ++tap; }
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        tb += adv_b;
        tt += adv_t;
        if (tt >= tiles_t) { tt -= tiles_t; ++tb; }
      }
      if (p.pdl_trigger != 0) pdl_launch_dependents();
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_f16_m128(p.BN);
    if (elect_one()) {  // one thread runs the whole loop (see gemm_tc_kernel)
      const uint64_t d0 = make_sw128_kmajor_desc(smem_u32(smem));
      const uint32_t stage_adv = static_cast<uint32_t>(stage_bytes) >> 4, b_adv = static_cast<uint32_t>(m_tile_bytes) >> 4;
      int s = 0, j = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.n_act_tiles; tile += gridDim.x, ++j) {
        const int as = j & 1;
        mbar_wait(tempty_bar + as, ((j >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(full_bar + s, ph);
          tc_fence_after();
          const uint64_t da = d0 + static_cast<uint64_t>(s) * stage_adv;
          const uint64_t db = da + b_adv;
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) umma_f16(tmem_base + as * p.BN, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          umma_commit(empty_bar + s);
          if (kb == nkb - 1) umma_commit(tfull_bar + as);
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else {
    const int etid = threadIdx.x - 64;
    EpiHoist hoist;
    epi_hoist_init(p, f0, etid, hoist);  // constants (bias, scales): before the dependency resolves
    pdl_wait();  // epilogue tensors (residual, outputs) belong to earlier kernels until they have completed
    int j = 0;
    for (int tile = blockIdx.x; tile < p.n_act_tiles; tile += gridDim.x, ++j) {
      const int as = j & 1;
      const int tb = tile / tiles_t;
      const int b0 = tb * p.G, t0 = (tile - tb * tiles_t) * p.R;
      if (warp == 2 && j == 0) PTTS_TRACE(12);
      if (warp < 6) {
        mbar_wait(tfull_bar + as, (j >> 1) & 1);
        tc_fence_after();
        if (warp == 2 && j < 3) PTTS_TRACE(j * 4 + 0);
        const int quad = warp & 3;
        const int i = quad * 32 + lane;
        for (int c = 0; c < p.BN; c += 16) {
          uint32_t v[16];
          tmem_ld16(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * p.BN + c, v);
          tmem_ld_wait();
          float4* dst = reinterpret_cast<float4*>(stile + i * LD + c);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            dst[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                                 __uint_as_float(v[4 * q + 3]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar + as);  // TMEM buffer free: the MMAs of tile j+2 may start
        if (warp == 2 && j < 3) PTTS_TRACE(j * 4 + 1);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(GEMM_THREADS - 64) : "memory");  // tile staged
      if (warp == 2 && j < 3) PTTS_TRACE(j * 4 + 2);
      epi_dispatch(p, smem_u32(stile), LD, f0, t0, b0, etid, GEMM_THREADS - 64, 0, 1, p.vec4 ? &hoist : nullptr);
      if (warp == 2 && j < 3) PTTS_TRACE(j * 4 + 3);
      asm volatile("bar.sync 1, %0;" ::"n"(GEMM_THREADS - 64) : "memory");  // staging buffer free
    }
  }
  if (warp == 2) PTTS_TRACE(13);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// SIMT cross-check of the same contract (tests only; selected by ptts_engine_cfg.debug_gemm or
// ptts_test_gemm(use_simt=1)).  One thread per output element.
__global__ void gemm_simt_kernel(const GemmParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long rows = static_cast<long long>(p.n_streams) * p.T;
  if (idx >= rows * p.F) return;
  const int r = static_cast<int>(idx / p.F);
  const int f = static_cast<int>(idx - static_cast<long long>(r) * p.F);
  const int b = r / p.T, t = r - b * p.T;
  const int C = p.cblocks * GEMM_BK;
  const __half* a = p.act + b * p.act_stream_stride + static_cast<long long>(t) * p.act_ld;
  const __half* w = p.w + static_cast<long long>(f) * p.K;
  float acc = 0.f;
  for (int j = 0; j < p.taps; ++j)
    for (int c = 0; c < C; ++c) acc += __half2float(a[j * p.act_ld + c]) * __half2float(w[j * C + c]);
  const GemmEpi& e = p.epi;
  float v = acc;
  if (e.wscale) v *= e.wscale[f];
  if (e.bias) v += e.bias[f];
  v = epi_act(e.act, v) * e.alpha;
  if (e.fscale) v *= e.fscale[f];
  if (e.gate) v *= e.gate[row_off(e.gate_map, r) + f];
  if (e.res) v += e.res[row_off(e.res_map, r) + f];
  if (e.out32) e.out32[row_off(e.out32_map, r) + f] = v;
  if (e.out16) e.out16[row_off(e.out16_map, r) + f] = __float2half_rn(e.act16 == ACT_ELU ? elu1(v) : v);
}

}  // namespace ptts
