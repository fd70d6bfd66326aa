// OPT-IN (PTTS_FLOW_SMALL=1; off by default).
// LSD flow head for 1-4 rows (a single utterance, BASELINE configs[0]): the same fourteen dependent Linears per Euler step
// as flow_head.cuh (reference modules/mlp.rs:135-171 ResBlock, :275-383 SimpleMLPAdaLN, flow_lm.rs:7-22 lsd_decode), cut for
// latency instead of tensor-core occupancy.
//
// flow_head_kernel gives a 16-row chunk to a cluster of 4 CTAs, each streaming a 128 KB weight slice per layer through one
// SM's TMA port (1.9 us) into a tcgen05 tile that is 1/16 full at one row: 3.5 us per layer, 50 us per frame whatever the
// batch size.  With so few rows a layer is a 512 x 512 matrix-vector product, so here ONE cluster of 8 CTAs owns the batch:
//   * CTA `rank` owns features [64 rank, 64 rank + 64) of every layer; its 64 KB weight slice of layer g + 2 is fetched by
//     one 1-D bulk copy into a two-stage shared-memory ring while layers g and g + 1 run (the first two before
//     griddepcontrol.wait: weights do not depend on the previous kernel), so no layer ever waits for HBM;
//   * a warp owns 8 features; lanes walk the weight rows in 16-byte chunks from shared memory, f32 accumulate,
//     shuffle reduce-scatter (gemv.cuh), one output element per lane;
//   * the residual stream x (f32 [rows][512]) and the f16 operand of the next Linear are REPLICATED in every CTA: an
//     epilogue lane writes its element straight into the shared memory of all eight CTAs (st.shared::cluster), one
//     cluster barrier per layer, and the LayerNorm + modulation in front of the next block is then computed locally
//     (one warp per row over the full 512-wide row, two-pass like ln_rows_kernel) -- no statistics exchange, no second
//     barrier.
// Same arithmetic as the fused kernel up to summation order (f16 operands, f32 accumulate, f32 residual, the int8 mode's
// f16 code copies with per-feature scales in the epilogue); parity against it and the oracle in tests/test_gemv_gpu.py.
#pragma once
#include "flow_head.cuh"
#include "gemv.cuh"
#include "lm_step.cuh"

namespace ptts {

static constexpr int FS_CL = 8;                          // CTAs in the cluster
static constexpr int FS_THREADS = 256;
static constexpr int FS_WARPS = FS_THREADS / 32;
static constexpr int FS_MAX_ROWS = 4;
static constexpr int FS_FEATS = FH_DIM / FS_CL;          // 64 features per CTA
static constexpr int FS_FPW = FS_FEATS / FS_WARPS;       // 8 features per warp
static constexpr int FS_W_STAGE = FS_FEATS * FH_DIM * 2; // 64 KB
static constexpr int FS_OP_BYTES = FS_MAX_ROWS * FH_DIM * 2;
static constexpr int FS_SMEM = 2 * FS_W_STAGE + 2 * FS_OP_BYTES + FS_MAX_ROWS * FH_DIM * 4 + 64;

struct FlowSmallParams {
  FlowHeadParams fp;
  const __half* w_in;     // [512][64] f16 (K padded 32 -> 64)
  const __half* w_pack;   // [12 x 512 + 128][512] f16: mlp.0 / mlp.2 of the six blocks, then the final Linear
};

__device__ __forceinline__ void st_dsmem_u16(uint32_t cluster_addr, uint16_t v) {
  asm volatile("st.shared::cluster.u16 [%0], %1;" ::"r"(cluster_addr), "h"(v) : "memory");
}

template <int ROWS>
__global__ void __launch_bounds__(FS_THREADS, 1) flow_head_small_kernel(const FlowSmallParams q) {
  extern __shared__ __align__(128) uint8_t fs_smem[];
  const FlowHeadParams& p = q.fp;
  uint8_t* w_s = fs_smem;                                               // [2][64 features][K] f16
  uint8_t* op_s = w_s + 2 * FS_W_STAGE;                                 // [2][4 rows][512] f16
  float* x_s = reinterpret_cast<float*>(op_s + 2 * FS_OP_BYTES);        // [4 rows][512] f32
  uint64_t* full = reinterpret_cast<uint64_t*>(x_s + FS_MAX_ROWS * FH_DIM);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int S = p.steps, n = p.n;
  const int NL = S * FH_LAYERS;

  auto issue = [&](int g) {   // thread 0: the weight slice of global layer g into stage g & 1
    const int L = g % FH_LAYERS;
    const __half* src;
    uint32_t bytes;
    if (L == 0) { src = q.w_in + static_cast<size_t>(rank) * FS_FEATS * 64; bytes = FS_FEATS * 64 * 2; }
    else if (L == FH_LAYERS - 1) { src = q.w_pack + (static_cast<size_t>(2 * FH_DEPTH) * FH_DIM + rank * 4) * FH_DIM; bytes = 4 * FH_DIM * 2; }
    else { src = q.w_pack + (static_cast<size_t>(L - 1) * FH_DIM + rank * FS_FEATS) * FH_DIM; bytes = FS_W_STAGE; }
    mbar_arrive_expect_tx(full + (g & 1), bytes);
    bulk_g2s(w_s + (g & 1) * FS_W_STAGE, src, bytes, full + (g & 1));
  };
  if (tid == 0) {
    mbar_init(full, 1);
    mbar_init(full + 1, 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {   // constants: requested before the dependency on the previous kernel resolves
    issue(0);
    if (NL > 1) issue(1);
  }
  pdl_launch_dependents();
  pdl_wait();
  cluster_sync_all();   // every CTA of the cluster is running before anyone writes into a peer's shared memory

  uint32_t op_peer[FS_CL], x_peer[FS_CL];
#pragma unroll
  for (int k = 0; k < FS_CL; ++k) {
    op_peer[k] = map_to_rank(smem_u32(op_s), k);
    x_peer[k] = map_to_rank(smem_u32(x_s), k);
  }
  // operand of the first input_proj: z rows x 64 halves (cols 32..63 are zero in z16), rows past the batch zero
  for (int i = tid; i < FS_MAX_ROWS * 8; i += FS_THREADS) {
    const int r = i >> 3, c = i & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < n) v = __ldcg(reinterpret_cast<const uint4*>(p.z16 + static_cast<long long>(r) * 64) + c);
    *reinterpret_cast<uint4*>(op_s + r * FH_DIM * 2 + c * 16) = v;
  }
  __syncthreads();

  constexpr int NV = FS_FPW * ROWS;       // partial sums per lane
  constexpr int SHARE = 32 / NV;          // lanes that end up holding the same total
  const int idx = lane / SHARE;
  const int ej = idx / ROWS, er = idx - ej * ROWS;        // the (feature slot, row) this lane finishes
  const bool writer = (lane % SHARE) == 0;
  const int f_mine = static_cast<int>(rank) * FS_FEATS + warp * FS_FPW + ej;   // its feature of a 512-wide layer

  int g = 0;
  for (int s = 0; s < S; ++s) {
    const float* const mod_s = p.mod + s * p.mod_step_stride;
    for (int L = 0; L < FH_LAYERS; ++L, ++g) {
      const int buf = g & 1;
      const bool is_final = (L == FH_LAYERS - 1), is_mlp0 = (L & 1) && !is_final, is_mlp2 = L > 0 && !(L & 1);
      const bool ln_next = (L == 0) || (is_mlp2);            // this layer completes x: LayerNorm + modulation follow
      const bool last_of_all = (s == S - 1 && is_final);
      const int K = (L == 0) ? 64 : FH_DIM;
      const uint8_t* wst = w_s + buf * FS_W_STAGE;
      const uint32_t xop = smem_u32(op_s + buf * FS_OP_BYTES);
      // per-element epilogue constants do not depend on the accumulator: requested before the weights are waited for
      float ws = 1.f, bias = 0.f, pre = 0.f;
      if (!is_final) {
        if (L == 0) { ws = p.ws_in ? __ldg(p.ws_in + f_mine) : 1.f; bias = __ldg(p.b_in + f_mine); }
        else if (is_mlp0) { const int i = (L - 1) >> 1; ws = p.ws0[i] ? __ldg(p.ws0[i] + f_mine) : 1.f; bias = __ldg(p.b0[i] + f_mine); }
        else {
          const int i = (L - 2) >> 1;
          ws = p.ws2[i] ? __ldg(p.ws2[i] + f_mine) : 1.f; bias = __ldg(p.b2[i] + f_mine);
          pre = __ldg(mod_s + static_cast<long long>(er) * FH_MOD_LD + i * 3 * FH_DIM + 2 * FH_DIM + f_mine);   // gate
        }
      }
      mbar_wait(full + buf, (g >> 1) & 1);
      if (!is_final) {
        float acc[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] = 0.f;
        const int nchunk = K / 8;
        for (int c = lane; c < nchunk; c += 32) {
          float xf[ROWS][8];
#pragma unroll
          for (int r = 0; r < ROWS; ++r) h8_to_f8(lds_u4(xop + r * FH_DIM * 2 + c * 16), xf[r]);
#pragma unroll
          for (int j = 0; j < FS_FPW; ++j) {
            float wf[8];
            h8_to_f8(*reinterpret_cast<const uint4*>(wst + static_cast<size_t>(warp * FS_FPW + j) * K * 2 + c * 16), wf);
#pragma unroll
            for (int r = 0; r < ROWS; ++r)
#pragma unroll
              for (int e = 0; e < 8; ++e) acc[j * ROWS + r] = fmaf(wf[e], xf[r][e], acc[j * ROWS + r]);
          }
        }
        warp_reduce_scatter<NV>(acc, lane);
        const float v = acc[0] * ws + bias;
        if (is_mlp0) {   // g = silu(W h + b): f16 operand of mlp.2, two features per store
          const float y = silu_fast(v);
          const float other = __shfl_xor_sync(0xffffffffu, y, 4);   // feature slot ej ^ 1, same row
          if (writer && (ej & 1) == 0) {
            const __half2 h2 = __floats2half2_rn(y, other);
            const uint32_t off = static_cast<uint32_t>((buf ^ 1) * FS_OP_BYTES + er * FH_DIM * 2 + f_mine * 2);
#pragma unroll
            for (int k = 0; k < FS_CL; ++k) st_dsmem_u32(op_peer[k] + off, *reinterpret_cast<const uint32_t*>(&h2));
          }
        } else {         // input_proj: x = W z + b;  mlp.2: x += gate * (W g + b)
          const float xn = (L == 0) ? v : x_s[er * FH_DIM + f_mine] + pre * v;
          if (writer) {
            const uint32_t off = static_cast<uint32_t>((er * FH_DIM + f_mine) * 4);
#pragma unroll
            for (int k = 0; k < FS_CL; ++k) st_dsmem_f1(x_peer[k] + off, xn);
            if (L == FH_LAYERS - 2 && s == S - 1 && p.x_dbg && er < n) p.x_dbg[static_cast<long long>(er) * FH_DIM + f_mine] = xn;
          }
        }
      } else {
        // final Linear + Euler step: z += (W h + b) / S; 32 features, four per CTA, one per warp 0-3
        if (warp < 4) {
          const int fl = static_cast<int>(rank) * 4 + warp;
          float acc[ROWS];
#pragma unroll
          for (int r = 0; r < ROWS; ++r) acc[r] = 0.f;
          for (int c = lane; c < FH_DIM / 8; c += 32) {
            float wf[8];
            h8_to_f8(*reinterpret_cast<const uint4*>(wst + static_cast<size_t>(warp) * FH_DIM * 2 + c * 16), wf);
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
              float xf[8];
              h8_to_f8(lds_u4(xop + r * FH_DIM * 2 + c * 16), xf);
#pragma unroll
              for (int e = 0; e < 8; ++e) acc[r] = fmaf(wf[e], xf[e], acc[r]);
            }
          }
          warp_reduce_scatter<ROWS>(acc, lane);
          constexpr int SH2 = 32 / ROWS;
          const int r = lane / SH2;
          if ((lane % SH2) == 0 && r < n) {
            const float wsf = p.ws_final ? __ldg(p.ws_final + fl) : 1.f;
            const float zn = p.z32[r * FH_LDIM + fl] + (acc[0] * wsf + __ldg(p.b_final + fl)) * p.alpha;
            p.z32[r * FH_LDIM + fl] = zn;
            const __half zh = __float2half_rn(zn);
            p.z16[r * 64 + fl] = zh;
            if (!last_of_all) {   // the next step's input_proj operand, in every CTA
              const uint32_t off = static_cast<uint32_t>((buf ^ 1) * FS_OP_BYTES + r * FH_DIM * 2 + fl * 2);
#pragma unroll
              for (int k = 0; k < FS_CL; ++k) st_dsmem_u16(op_peer[k] + off, __half_as_ushort(zh));
            }
          }
        }
        if (last_of_all) break;
        // cols 32..63 of that operand are zero (K is padded to 64), as are rows past the batch: the buffer last held a
        // 512-wide operand.  Local stores; the peers only write cols 0..31 of rows < n.
        for (int i = tid; i < FS_MAX_ROWS * 32; i += FS_THREADS) {
          const int r = i >> 5, c = i & 31;
          const bool keep = (r < n) && c < 16;   // 16 half2 = cols 0..31 of a live row: written by the peers
          if (!keep) *reinterpret_cast<uint32_t*>(op_s + (buf ^ 1) * FS_OP_BYTES + r * FH_DIM * 2 + c * 4) = 0u;
        }
      }
      // the modulation rows and affine parameters of the LayerNorm that follows do not depend on x: requested in front of
      // the cluster barrier, they arrive while it resolves instead of costing an L2 round trip behind it
      float4 sh4[4], sc4[4], lw4[4], lb4[4];
      if (ln_next && warp < ROWS) {
        const int j = L >> 1;
        const float* shift = mod_s + static_cast<long long>(warp) * FH_MOD_LD + j * 3 * FH_DIM;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = (i * 32 + lane) * 4;
          sh4[i] = __ldg(reinterpret_cast<const float4*>(shift + c));
          sc4[i] = __ldg(reinterpret_cast<const float4*>(shift + FH_DIM + c));
          lw4[i] = make_float4(1.f, 1.f, 1.f, 1.f);
          lb4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (j < FH_DEPTH) { lw4[i] = __ldg(reinterpret_cast<const float4*>(p.ln_w[j] + c)); lb4[i] = __ldg(reinterpret_cast<const float4*>(p.ln_b[j] + c)); }
        }
      }
      cluster_sync_all();   // the next operand (or the completed residual row) is in every CTA; stage `buf` is free
      if (tid == 0 && g + 2 < NL) issue(g + 2);
      if (ln_next) {
        // h = LN(x) [* w + b] * (1 + scale) + shift for block j (or the final layer: no affine), one warp per row over
        // the replicated x; two-pass statistics, biased variance, eps inside the sqrt (modules/mlp.rs:29-58,168-171)
        if (warp < ROWS) {
          const int r = warp;
          const float* xr = x_s + r * FH_DIM;
          float v[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 t4 = *reinterpret_cast<const float4*>(xr + (i * 32 + lane) * 4);
            v[4 * i] = t4.x; v[4 * i + 1] = t4.y; v[4 * i + 2] = t4.z; v[4 * i + 3] = t4.w;
          }
          float sum = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) sum += v[i];
          const float mean = warp_sum(sum) * (1.f / FH_DIM);
          float m2 = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) { const float d = v[i] - mean; m2 += d * d; }
          const float rstd = 1.f / sqrtf(warp_sum(m2) * (1.f / FH_DIM) + 1e-6f);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int c = (i * 32 + lane) * 4;
            const float sh[4] = {sh4[i].x, sh4[i].y, sh4[i].z, sh4[i].w}, sc[4] = {sc4[i].x, sc4[i].y, sc4[i].z, sc4[i].w};
            const float lw[4] = {lw4[i].x, lw4[i].y, lw4[i].z, lw4[i].w}, lb[4] = {lb4[i].x, lb4[i].y, lb4[i].z, lb4[i].w};
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float t = (v[4 * i + e] - mean) * rstd * lw[e] + lb[e];
              o[e] = (r < n) ? t * (1.f + sc[e]) + sh[e] : 0.f;
            }
            const __half2 h0 = __floats2half2_rn(o[0], o[1]), h1 = __floats2half2_rn(o[2], o[3]);
            uint2 pk;
            pk.x = *reinterpret_cast<const uint32_t*>(&h0);
            pk.y = *reinterpret_cast<const uint32_t*>(&h1);
            *reinterpret_cast<uint2*>(op_s + (buf ^ 1) * FS_OP_BYTES + r * FH_DIM * 2 + c * 2) = pk;
          }
        }
        __syncthreads();
      }
    }
  }
}

}  // namespace ptts
