// The FlowLM half of one decode step as ONE persistent kernel (reference models/flow_lm.rs:98-145:
// input_linear -> 6 x StreamingTransformerLayer (models/transformer.rs:66-90, modules/attention.rs:104-231) -> out_norm +
// EOS head, then the flow head's cond_embed and every adaLN modulation of every LSD step, modules/mlp.rs:275,322-368).
//
// As separate launches this was ~48 kernels of 4-10 us each for 2-8 MB of weights apiece: every launch paid a grid
// hand-off, a cold TMA pipeline and a split-K epilogue.  Here G CTAs (one per SM) stay resident for the whole step:
//   * weights are pre-tiled in HBM as the exact shared-memory image of a [128 features x 64 k] SWIZZLE_128B operand
//     tile (16 KB, tiles of one feature tile consecutive along k), so the producer warp streams them with plain 1-D
//     bulk copies (cp.async.bulk, no tensor map) into a 4-stage ring.  The order of tiles a CTA needs is static, so
//     the producer never waits for a phase to begin: it runs ahead of the grid barriers and is throttled by the ring
//     alone -- the HBM stream does not drain at layer boundaries;
//   * activations are handed from phase to phase through L2 in the same tiled image ([k-block][64 rows][64 k] f16,
//     swizzled), written by ordinary stores of the phase that produces them and fetched by ONE bulk copy per GEMM unit;
//   * a GEMM phase is split over all CTAs as (feature tile, K slice) units; a unit accumulates [128 features x 64 rows]
//     in TMEM (tcgen05.mma, weights on MMA-M, batch rows on MMA-N) and stores its f32 partial to a workspace in L2;
//     the consumer phase sums the partials of a row in split order (fixed order: bit-reproducible, no atomics), which
//     is also where bias / residual / LayerNorm / GELU / RoPE happen;
//   * phases are separated by a grid barrier (one monotonic counter in global memory, bounded spin).
// Per layer: in_proj | attention | out_proj | +x, LN2 | linear1 | GELU | linear2 | +x, LN1 of the next layer (or out_norm +
// EOS logit after the last).
// Attention (RoPE + KV append + softmax.V): a CTA owns one head and every (G/16)-th batch row, one warp per (row, head).
// Every segment of a voice restarts from the same voice KV (tts_model.rs:940), so the K / V rows of that prefix are the
// same for all rows of a batch: they are staged ONCE per CTA in shared memory (one bulk copy each) and every warp scores
// them from there, instead of every row re-reading them through L2.  The rows a stream owns are streamed by the idle
// MMA warp through a two-slot ring of 32-key chunks per worker warp (bulk copies, 128 KB in flight per SM), so the
// score / softmax / P.V arithmetic never waits on a register load.
#pragma once
#include "kernels.cuh"

namespace ptts {

static constexpr int LM_ROWS = 64;                    // batch rows a step kernel handles (MMA N)
static constexpr int LM_THREADS = 320;                // warp 0 weight producer, warp 1 MMA issuer / KV producer, warps 2-9 workers
static constexpr int LM_WORKERS = 256;
static constexpr int LM_STAGES = 4;                   // weight ring: one GEMM unit of look-ahead
static constexpr int LM_WTILE = 128 * 64 * 2;         // one weight tile image (bytes)
static constexpr int LM_ATILE = LM_ROWS * 64 * 2;     // one activation k-block image (bytes)
static constexpr int LM_ACT_KB = 8;                   // k-blocks of activations resident per unit (adaLN: K = 512 unsplit)
// attention staging (aliases the activation operand buffer, which is idle in the attention phase, plus 96 KB behind it):
// the shared voice prefix of this CTA's head (K and V rows of up to LM_PFX keys) and one ring of LM_KV_SLOTS chunks of
// LM_KV_CHUNK keys (K rows then V rows) per worker warp
static constexpr int LM_PFX = 128, LM_KV_CHUNK = 32, LM_KV_SLOTS = 2;
static constexpr int LM_PFX_BYTES = 2 * LM_PFX * 128;
static constexpr int LM_KV_SLOT_BYTES = 2 * LM_KV_CHUNK * 128;
static constexpr int LM_ATT_BYTES = LM_PFX_BYTES + 8 * LM_KV_SLOTS * LM_KV_SLOT_BYTES;   // 160 KB
static constexpr int LM_DYN_BYTES = LM_STAGES * LM_WTILE + (LM_ATT_BYTES > LM_ACT_KB * LM_ATILE ? LM_ATT_BYTES : LM_ACT_KB * LM_ATILE);
static constexpr int LM_SMEM = LM_DYN_BYTES + 1024 + 1024;
static constexpr int LM_LAYERS = 6, LM_D = 1024, LM_FFN = 4096, LM_HEADS = 16, LM_FLOW = 512, LM_MOD = 10240;
static constexpr int LM_MAX_LSD = 4;                  // modulation rows of up to 4 LSD steps per launch
static constexpr int LM_PH_LAYER0 = 1, LM_PH_PER_LAYER = 8;
static constexpr int LM_PH_COND = LM_PH_LAYER0 + LM_LAYERS * LM_PH_PER_LAYER;  // 49
static constexpr int LM_PH_ADA = LM_PH_COND + 2;                               // 51

enum { LM_G_INPROJ = 0, LM_G_OUTPROJ, LM_G_LIN1, LM_G_LIN2, LM_G_COND, LM_G_ADA, LM_G_KINDS };

struct LmGemmShape { int Mt, KB, S, kbps; };  // feature tiles, k-blocks, K splits, k-blocks per split

struct LmStepParams {
  int n;                       // batch rows, 1..64
  int lsd_steps;               // 1..LM_MAX_LSD
  int flags;                   // bring-up switches: 1 = never stage the voice prefix in shared memory
  int stop_phase;              // bring-up: run only phases [0, stop_phase) when > 0
  LmGemmShape shape[LM_G_KINDS];
  const uint8_t* w_inproj[LM_LAYERS];   // tiled weight images
  const uint8_t* w_outproj[LM_LAYERS];
  const uint8_t* w_lin1[LM_LAYERS];
  const uint8_t* w_lin2[LM_LAYERS];
  const uint8_t* w_cond;
  const uint8_t* w_ada;
  const __half* w_input;       // [1024][64] f16 row-major (k >= 32 zero)
  const float* ln1_w[LM_LAYERS]; const float* ln1_b[LM_LAYERS];
  const float* ln2_w[LM_LAYERS]; const float* ln2_b[LM_LAYERS];
  const float* outnorm_w; const float* outnorm_b; const float* eos_w; const float* eos_b;
  const float* b_cond; const float* b_ada; const float* time_emb;  // [lsd][512]
  // per-step stream state
  const int* row_seq; const StreamCtl* ctl; const float* feedback; const SeqDesc* seqs; const int* own_len;
  SeqDesc* row_desc;
  // activations and workspaces (all device global, L2 resident)
  float* x32;                  // [64][1024] residual stream
  uint8_t* hA;                 // [16 kb][64][64] f16 image: LayerNorm output (operand of in_proj / linear1 / cond_embed)
  uint8_t* attnA;              // [16 kb] image: attention output (operand of out_proj)
  uint8_t* ffnA;               // [64 kb] image: GELU output (operand of linear2)
  uint8_t* yA;                 // [lsd][8 kb] image: silu(c + te[s]) (operand of the adaLN Linears)
  float* ws;                   // split-K partials [S][64][F] of the GEMM phase in flight
  float* z32; __half* z16; float* eos_logit; float* c32; float* mod32 /*[lsd][64][10240]*/; float* h32dbg;
  unsigned long long* bar;     // [0] arrival counter (monotonic), [1] counter value at the start of the next launch
  unsigned long long* trace;   // optional [phases][2] %globaltimer stamps of CTA 0 (phase start, work done)
};

// ---------------------------------------------------------------- small PTX helpers
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ float4 ldcg_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float2 ldcg_f2(const float* p) {
  float2 v;
  asm volatile("ld.global.cg.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 ldnc_u4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
// 32 lanes x 32 consecutive 32-bit columns -> 32 registers per thread, load and wait in ONE asm statement: the registers
// of a tcgen05.ld are not valid until tcgen05.wait::ld, and a compiler that sees two statements is free to spill them
// in between (it did, once this kernel needed spills: every GEMM phase then stored stale registers).
__device__ __forceinline__ void tmem_ld32_wait(uint32_t taddr, uint32_t (&v)[32]) {
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
// tcgen05.mma with both shared-memory descriptors handed over as (low word, shared high word): K-major SWIZZLE_128B
// tiles only differ in their start address, and keeping the 64-bit values out of the compiler's hands matters -- with the
// descriptors built as uint64_t arithmetic (base + k-block * tile + k-step), one build of this kernel came out with the
// high word (layout, SBO, version) of the B descriptor dropped: UTCHMMA then read the operand unswizzled and only batch
// row 0 was right.
static constexpr uint32_t LM_DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO = 1024 B, version 1, SWIZZLE_128B
__device__ __forceinline__ uint32_t lm_desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ void lm_umma_f16(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(LM_DESC_HI)
      : "memory");
}
// byte offset of element (row r, feature k) in an activation image [k-block][64 rows][64 k], SWIZZLE_128B K-major
__device__ __forceinline__ uint32_t lm_img_off(int r, int k) {
  return static_cast<uint32_t>(k >> 6) * LM_ATILE + static_cast<uint32_t>(r) * 128u +
         (static_cast<uint32_t>(((k & 63) >> 3) ^ (r & 7)) << 4) + static_cast<uint32_t>(k & 7) * 2u;
}
__device__ __forceinline__ void lm_store_img4(uint8_t* img, int r, int k, float a, float b, float c, float d) {
  const __half2 h0 = __floats2half2_rn(a, b), h1 = __floats2half2_rn(c, d);
  uint2 pk;
  pk.x = *reinterpret_cast<const uint32_t*>(&h0);
  pk.y = *reinterpret_cast<const uint32_t*>(&h1);
  *reinterpret_cast<uint2*>(img + lm_img_off(r, k)) = pk;
}

// Grid barrier for the 288 threads of warps 1-9 (the producer warp never takes part).  Arrival counter is monotonic over
// the life of the engine: barrier i of a launch that started at counter value `base` completes at base + (i + 1) * G.
// Everything another CTA produced is read either by bulk copies or by ld.global.cg (both served by L2), so no L1
// invalidation is needed behind the acquire.
__device__ __forceinline__ void lm_grid_barrier(unsigned long long* ctr, unsigned long long target, bool trailing_fence) {
  asm volatile("bar.sync 1, 288;" ::: "memory");
  if (threadIdx.x == 64) {
    __threadfence();  // cumulative: publishes the writes of every thread that arrived at the bar.sync above
    atomicAdd(ctr, 1ULL);
    unsigned long long v;
    uint32_t spins = 0;
    while (true) {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory");
      if (v >= target) break;
      if (++spins > (1u << 25)) {
        printf("ptts: FlowLM step kernel grid barrier timed out (block %d: counter %llu, target %llu)\n", blockIdx.x, v, target);
        __trap();
      }
    }
    if (trailing_fence) __threadfence();
  }
  asm volatile("bar.sync 1, 288;" ::: "memory");
}

// Sum over the 256 worker threads; every thread gets the total.  `slot` selects one of four scratch rows so that
// back-to-back reductions need one barrier each.
__device__ __forceinline__ float lm_block_sum(float v, float* red_s, int slot, int wt) {
  v = warp_sum(v);
  if ((wt & 31) == 0) red_s[slot * 8 + (wt >> 5)] = v;
  asm volatile("bar.sync 2, 256;" ::: "memory");
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += red_s[slot * 8 + i];
  return t;
}

struct LmGemm {
  const uint8_t* w; const uint8_t* act; float* out; const float* bias;
  int Mt, KB, S, kbps, F;
};
// GEMM phase descriptor; false for phases without tensor-core work.
__device__ __forceinline__ bool lm_gemm_of(const LmStepParams& p, int ph, LmGemm& g) {
  int kind;
  const uint8_t* w; const uint8_t* act; float* out = p.ws; const float* bias = nullptr; int F;
  if (ph >= LM_PH_LAYER0 && ph < LM_PH_COND) {
    const int l = (ph - LM_PH_LAYER0) >> 3, j = (ph - LM_PH_LAYER0) & 7;
    if (j == 0) { kind = LM_G_INPROJ; w = p.w_inproj[l]; act = p.hA; F = 3 * LM_D; }
    else if (j == 2) { kind = LM_G_OUTPROJ; w = p.w_outproj[l]; act = p.attnA; F = LM_D; }
    else if (j == 4) { kind = LM_G_LIN1; w = p.w_lin1[l]; act = p.hA; F = LM_FFN; }
    else if (j == 6) { kind = LM_G_LIN2; w = p.w_lin2[l]; act = p.ffnA; F = LM_D; }
    else return false;
  } else if (ph == LM_PH_COND) {
    kind = LM_G_COND; w = p.w_cond; act = p.hA; F = LM_FLOW;
  } else if (ph >= LM_PH_ADA) {
    const int s = ph - LM_PH_ADA;
    kind = LM_G_ADA; w = p.w_ada; act = p.yA + static_cast<size_t>(s) * 8 * LM_ATILE; F = LM_MOD;
    out = p.mod32 + static_cast<size_t>(s) * LM_ROWS * LM_MOD; bias = p.b_ada;
  } else {
    return false;
  }
  const LmGemmShape sh = p.shape[kind];
  g.w = w; g.act = act; g.out = out; g.bias = bias; g.Mt = sh.Mt; g.KB = sh.KB; g.S = sh.S; g.kbps = sh.kbps; g.F = F;
  return true;
}

// x (4 consecutive features per thread, the whole row over the 256 workers) -> LayerNorm (biased variance, eps inside the
// sqrt: modules/mlp.rs:29-58) -> f16 image.  Returns the four normalised values (for the EOS dot product).
__device__ __forceinline__ float4 lm_ln_to_image(float4 v, const float* __restrict__ w, const float* __restrict__ b, uint8_t* img,
                                                 int r, int f, float* red_s, int slot0, int wt) {
  const float mean = lm_block_sum(v.x + v.y + v.z + v.w, red_s, slot0, wt) * (1.f / LM_D);
  const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
  const float var = lm_block_sum(dx * dx + dy * dy + dz * dz + dw * dw, red_s, slot0 + 1, wt) * (1.f / LM_D);
  const float rstd = 1.f / sqrtf(var + 1e-5f);
  const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + f)), b4 = __ldg(reinterpret_cast<const float4*>(b + f));
  float4 y;
  y.x = dx * rstd * w4.x + b4.x; y.y = dy * rstd * w4.y + b4.y; y.z = dz * rstd * w4.z + b4.z; y.w = dw * rstd * w4.w + b4.w;
  lm_store_img4(img, r, f, y.x, y.y, y.z, y.w);
  return y;
}

__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint4 lds_u4(const uint8_t* p) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)));
  return v;
}

// Up to 32 keys whose K rows sit at kb and V rows at vb in shared memory (128 bytes per row): one online-softmax update
// of (m, l, acc).  Eight lanes share a key (16 bytes of its K row and of its V row each), the warp takes four keys per
// row of lanes, eight rows = 32 keys per call.  Scores are in base 2 (q carries log2(e) / sqrt(64)), so every
// exponential is one ex2.approx (relative error 2^-22, far below the f16 rounding of K and V).  MASK = false: all 32
// rows are keys; MASK = true: rows past nk are read (they lie inside the slot) and masked.
template <bool MASK>
__device__ __forceinline__ void lm_attn_accum32(const uint8_t* kb, const uint8_t* vb, int nk, const float (&qv)[8], float& m,
                                                float& l, float (&acc)[8], int sub, int part) {
  uint4 ku[8], vu[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int i = u * 4 + sub;
    ku[u] = lds_u4(kb + i * 128 + part * 16);
    vu[u] = lds_u4(vb + i * 128 + part * 16);
  }
  float sc[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const __half2* kh2 = reinterpret_cast<const __half2*>(&ku[u]);
    sc[u] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(kh2[j]);
      sc[u] += f.x * qv[2 * j] + f.y * qv[2 * j + 1];
    }
  }
#pragma unroll
  for (int x = 1; x <= 4; x <<= 1)
#pragma unroll
    for (int u = 0; u < 8; ++u) sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], x);
  float m_new = m;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    if (MASK && u * 4 + sub >= nk) { sc[u] = -INFINITY; vu[u] = make_uint4(0, 0, 0, 0); }  // stale bytes must not reach 0 * v
    m_new = fmaxf(m_new, sc[u]);
  }
  if (!MASK || m_new != -INFINITY) {
    const float corr = exp2f(m - m_new);  // 2^-inf = 0 on the first batch
    l *= corr;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= corr;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float pw = exp2f(sc[u] - m_new);
      l += pw;
      const __half2* vh2 = reinterpret_cast<const __half2*>(&vu[u]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(vh2[j]);
        acc[2 * j] += pw * f.x;
        acc[2 * j + 1] += pw * f.y;
      }
    }
    m = m_new;
  }
}
__device__ __forceinline__ void lm_attn_accum(const uint8_t* kb, const uint8_t* vb, int nk, const float (&qv)[8], float& m, float& l,
                                              float (&acc)[8], int sub, int part) {
  if (nk >= 32) lm_attn_accum32<false>(kb, vb, 32, qv, m, l, acc, sub, part);
  else lm_attn_accum32<true>(kb, vb, nk, qv, m, l, acc, sub, part);
}

// Attention staging in shared memory (see LM_ATT_BYTES) and its barriers.
struct LmAttnSmem {
  uint8_t* pfx_k; uint8_t* pfx_v; uint8_t* rings;   // rings: [8 warps][LM_KV_SLOTS][K rows | V rows]
  uint64_t* pfx_full; uint64_t* kv_full;   // kv_full: [8][LM_KV_SLOTS]
};
// the key runs of one (row, head) that go through the warp's ring: the prefix unless it is staged, then the own rows
struct LmKvRuns { const __half* k[2]; const __half* v[2]; int len[2]; };
__device__ __forceinline__ bool lm_same_prefix(const SeqDesc& a, const SeqDesc& b) {
  return a.prefix == b.prefix && a.prefix_len == b.prefix_len && a.prefix_cap == b.prefix_cap;
}
__device__ __forceinline__ void lm_kv_runs(const SeqDesc& sd, int layer, int h, bool staged, LmKvRuns& r) {
  r.k[0] = kv_row(sd, layer, 0, h, LM_HEADS, 0);
  r.v[0] = kv_row(sd, layer, 1, h, LM_HEADS, 0);
  r.len[0] = staged ? 0 : sd.prefix_len;
  r.k[1] = sd.own + ((static_cast<long long>(layer * 2 + 0) * LM_HEADS + h) * sd.own_cap) * HD;
  r.v[1] = sd.own + ((static_cast<long long>(layer * 2 + 1) * LM_HEADS + h) * sd.own_cap) * HD;
  r.len[1] = sd.pad;  // rows the stream has appended so far (the new row of this step is taken from registers)
}

// One (row, head) of FlowLM decode attention by one warp (reference modules/attention.rs:104-231 with t = 1): q, k, v are
// the sums of the in_proj partials; RoPE at the absolute position; K, V appended to the stream's cache as f16; causal
// softmax(q K^T / 8) V over prefix + own rows with a running maximum (modules/sdpa.rs:36-82 up to f32 rounding).  The
// new key is taken from registers (rounded to f16 like the row later steps will read).
// The warp feeds its own ring: lane 0 requests the first LM_KV_SLOTS chunks before anything else (they fly while q is
// being built and the staged prefix is scored) and re-requests a slot the moment the warp has finished reading it.
__device__ __forceinline__ void lm_attend_item(const LmStepParams& p, const LmAttnSmem& A, int layer, int r, int h, int w, int lane,
                                               int SQ, const SeqDesc& sd0, bool cta_staged, uint32_t& kv_cnt) {
#define LM_ATRACE(i) do { if (p.trace && layer == 1 && blockIdx.x == 0 && w == 0 && lane == 0) p.trace[104 + (i)] = gtime(); } while (0)
  LM_ATRACE(0);
  const SeqDesc sd = p.row_desc[r];
  const int pos = sd.prefix_len + sd.pad;
  const bool staged = cta_staged && lm_same_prefix(sd, sd0);
  LmKvRuns runs;
  lm_kv_runs(sd, layer, h, staged, runs);
  const int c0 = (runs.len[0] + LM_KV_CHUNK - 1) / LM_KV_CHUNK, c1 = (runs.len[1] + LM_KV_CHUNK - 1) / LM_KV_CHUNK;
  const int nchunks = c0 + c1;
  const uint32_t cnt0 = kv_cnt;
  auto chunk_keys = [&](int c) { return c < c0 ? min(LM_KV_CHUNK, runs.len[0] - c * LM_KV_CHUNK) : min(LM_KV_CHUNK, runs.len[1] - (c - c0) * LM_KV_CHUNK); };
  auto issue = [&](int c) {  // lane 0 only
    const int ri = c < c0 ? 0 : 1, off = (c < c0 ? c : c - c0) * LM_KV_CHUNK;
    const uint32_t bytes = static_cast<uint32_t>(chunk_keys(c)) * 128u;
    const int slot = (cnt0 + c) % LM_KV_SLOTS;
    uint8_t* sb = A.rings + (w * LM_KV_SLOTS + slot) * LM_KV_SLOT_BYTES;
    uint64_t* fb = A.kv_full + w * LM_KV_SLOTS + slot;
    mbar_arrive_expect_tx(fb, 2 * bytes);
    bulk_g2s(sb, (ri ? runs.k[1] : runs.k[0]) + static_cast<long long>(off) * HD, bytes, fb);
    bulk_g2s(sb + LM_KV_CHUNK * 128, (ri ? runs.v[1] : runs.v[0]) + static_cast<long long>(off) * HD, bytes, fb);
  };
  if (lane == 0)
    for (int c = 0; c < min(nchunks, LM_KV_SLOTS); ++c) issue(c);

  LM_ATRACE(1);
  float2 q = make_float2(0.f, 0.f), k = q, v = q;
  const float* base = p.ws + static_cast<size_t>(r) * (3 * LM_D) + h * HD + 2 * lane;
  for (int s0 = 0; s0 < SQ; s0 += 8) {  // up to eight partials (24 loads) in flight before the first add
    // unconditional loads from a clamped split index keep the arrays in registers (a predicated definition sent them
    // to local memory, which with this kernel's shared-memory carve-out means L2 round trips)
    float2 a0[8], a1[8], a2[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const float* b = base + static_cast<size_t>(min(s0 + s, SQ - 1)) * LM_ROWS * (3 * LM_D);
      a0[s] = ldcg_f2(b); a1[s] = ldcg_f2(b + LM_D); a2[s] = ldcg_f2(b + 2 * LM_D);
    }
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const float on = (s0 + s < SQ) ? 1.f : 0.f;
      q.x += on * a0[s].x; q.y += on * a0[s].y; k.x += on * a1[s].x; k.y += on * a1[s].y; v.x += on * a2[s].x; v.y += on * a2[s].y;
    }
  }
  LM_ATRACE(2);
  float qr, qi, kr, ki;
  rope_pair(q.x, q.y, pos, lane, qr, qi);
  rope_pair(k.x, k.y, pos, lane, kr, ki);
  LM_ATRACE(3);
  const __half2 kh = __floats2half2_rn(kr, ki), vh = __floats2half2_rn(v.x, v.y);
  reinterpret_cast<__half2*>(const_cast<__half*>(kv_row(sd, layer, 0, h, LM_HEADS, pos)))[lane] = kh;
  reinterpret_cast<__half2*>(const_cast<__half*>(kv_row(sd, layer, 1, h, LM_HEADS, pos)))[lane] = vh;
  const float2 kf = __half22float2(kh), vf = __half22float2(vh);
  constexpr float kScale = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) and log2(e): softmax in base 2
  const float s_new = warp_sum(qr * kf.x + qi * kf.y) * kScale;

  const int sub = lane >> 3, part = lane & 7;
  float qv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float t0 = __shfl_sync(0xffffffffu, qr, part * 4 + (j >> 1));
    const float t1 = __shfl_sync(0xffffffffu, qi, part * 4 + (j >> 1));
    qv[j] = ((j & 1) ? t1 : t0) * kScale;
  }
  float m = -INFINITY, l = 0.f, acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  LM_ATRACE(4);
  if (staged) {
    mbar_wait(A.pfx_full, layer & 1);
    LM_ATRACE(5);
    for (int c = 0; c < sd.prefix_len; c += 32)
      lm_attn_accum(A.pfx_k + c * 128, A.pfx_v + c * 128, min(32, sd.prefix_len - c), qv, m, l, acc, sub, part);
  }
  LM_ATRACE(6);
  for (int c = 0; c < nchunks; ++c) {
    const uint32_t qn = cnt0 + c;
    const int slot = qn % LM_KV_SLOTS;
    mbar_wait(A.kv_full + w * LM_KV_SLOTS + slot, (qn / LM_KV_SLOTS) & 1);
    if (c < 6) LM_ATRACE(7 + 2 * c);
    const uint8_t* sb = A.rings + (w * LM_KV_SLOTS + slot) * LM_KV_SLOT_BYTES;
    lm_attn_accum(sb, sb + LM_KV_CHUNK * 128, chunk_keys(c), qv, m, l, acc, sub, part);
    __syncwarp();  // every lane has its bytes of the slot in registers: the slot may be refilled
    if (lane == 0 && c + LM_KV_SLOTS < nchunks) issue(c + LM_KV_SLOTS);
    if (c < 6) LM_ATRACE(8 + 2 * c);
  }
  LM_ATRACE(19);
  kv_cnt = cnt0 + nchunks;
  // fold the four key sub-groups (fixed order), then the new key
#pragma unroll
  for (int x = 8; x <= 16; x <<= 1) {
    const float m_o = __shfl_xor_sync(0xffffffffu, m, x);
    const float l_o = __shfl_xor_sync(0xffffffffu, l, x);
    const float m_new = fmaxf(m, m_o);
    const float ca = (m == -INFINITY) ? 0.f : exp2f(m - m_new);
    const float cb = (m_o == -INFINITY) ? 0.f : exp2f(m_o - m_new);
    l = l * ca + l_o * cb;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float a_o = __shfl_xor_sync(0xffffffffu, acc[j], x);
      acc[j] = acc[j] * ca + a_o * cb;
    }
    m = m_new;
  }
  {
    const float m_new = fmaxf(m, s_new);
    const float ca = (m == -INFINITY) ? 0.f : exp2f(m - m_new);
    const float cb = exp2f(s_new - m_new);
    l = l * ca + cb;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float t0 = __shfl_sync(0xffffffffu, vf.x, part * 4 + (j >> 1));
      const float t1 = __shfl_sync(0xffffffffu, vf.y, part * 4 + (j >> 1));
      acc[j] = acc[j] * ca + cb * ((j & 1) ? t1 : t0);
    }
  }
  if (sub == 0) {
    const float inv = 1.f / l;
    uint4 o;
    const __half2 h0 = __floats2half2_rn(acc[0] * inv, acc[1] * inv), h1 = __floats2half2_rn(acc[2] * inv, acc[3] * inv);
    const __half2 h2 = __floats2half2_rn(acc[4] * inv, acc[5] * inv), h3 = __floats2half2_rn(acc[6] * inv, acc[7] * inv);
    o.x = *reinterpret_cast<const uint32_t*>(&h0); o.y = *reinterpret_cast<const uint32_t*>(&h1);
    o.z = *reinterpret_cast<const uint32_t*>(&h2); o.w = *reinterpret_cast<const uint32_t*>(&h3);
    // head h is k-block h of the attention image; this lane's 8 dims are 16-byte chunk `part` of the row
    *reinterpret_cast<uint4*>(p.attnA + lm_img_off(r, h * HD + part * 8)) = o;
  }
  LM_ATRACE(20);
#undef LM_ATRACE
}

__global__ void __launch_bounds__(LM_THREADS, 1) flowlm_step_kernel(const LmStepParams p) {
  extern __shared__ __align__(1024) uint8_t lm_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(lm_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;
  uint8_t* act_s = smem + LM_STAGES * LM_WTILE;   // GEMM phases: the unit's operand image; attention phases: KV staging
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + LM_DYN_BYTES);
  uint64_t* empty = full + LM_STAGES;
  uint64_t* act_full = empty + LM_STAGES;
  uint64_t* acc_full = act_full + 1;
  uint64_t* tmem_empty = acc_full + 1;
  LmAttnSmem A;
  A.pfx_k = act_s; A.pfx_v = act_s + LM_PFX * 128; A.rings = act_s + LM_PFX_BYTES;
  A.pfx_full = tmem_empty + 1; A.kv_full = A.pfx_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(A.kv_full + 8 * LM_KV_SLOTS);
  float* red_s = reinterpret_cast<float*>(tmem_slot + 2);  // [4][8]
  float* lat_s = red_s + 32;                                // [32]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x, G = gridDim.x;
  const int NPH = p.stop_phase > 0 ? min(p.stop_phase, LM_PH_ADA + p.lsd_steps) : LM_PH_ADA + p.lsd_steps;

  if (threadIdx.x == 0) {
    for (int s = 0; s < LM_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    mbar_init(act_full, 1);
    mbar_init(acc_full, 1);
    mbar_init(tmem_empty, 8);
    mbar_init(A.pfx_full, 1);
    for (int i = 0; i < 8 * LM_KV_SLOTS; ++i) mbar_init(A.kv_full + i, 1);
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

  if (warp == 0) {
    // ===== weight producer: the static tile sequence of this CTA, throttled by the ring alone =====
    if (lane == 0) {
      int s = 0;
      uint32_t par = 0;
      for (int ph = 0; ph < NPH; ++ph) {
        LmGemm g;
        if (!lm_gemm_of(p, ph, g)) continue;
        const int U = g.Mt * g.S;
        for (int u = cta; u < U; u += G) {
          const int mt = u / g.S, sp = u - mt * g.S, kb0 = sp * g.kbps, nkb = min(g.kbps, g.KB - kb0);
          const uint8_t* src = g.w + (static_cast<size_t>(mt) * g.KB + kb0) * LM_WTILE;
          for (int i = 0; i < nkb; ++i) {
            mbar_wait(empty + s, par ^ 1);
            mbar_arrive_expect_tx(full + s, LM_WTILE);
            bulk_g2s(ring + s * LM_WTILE, src + static_cast<size_t>(i) * LM_WTILE, LM_WTILE, full + s);
            if (++s == LM_STAGES) { s = 0; par ^= 1; }
          }
        }
      }
    }
  } else {
    // ===== warps 1-9: the phase program =====
    pdl_wait();  // everything below reads what earlier kernels of the stream produced
    const int wt = threadIdx.x - 64;  // worker index 0..255 (negative for warp 1)
    unsigned long long bar_base = 0;
    if (threadIdx.x == 64) bar_base = *reinterpret_cast<volatile unsigned long long*>(p.bar + 1);
    int ring_s = 0;
    uint32_t ring_par = 0, uc = 0;  // uc: GEMM units this CTA has run (parity of act_full / acc_full / tmem_empty)
    uint32_t kv_cnt = 0;            // worker warp: chunks taken from its KV ring so far
    const int NG = G / LM_HEADS;    // attention: CTA = (head, row group); CTAs past 16 * NG sit the phase out
    const uint32_t idesc = make_idesc_f16_m128(LM_ROWS);
    const int n = p.n;

    for (int ph = 0; ph < NPH; ++ph) {
      if (p.trace && cta == 0 && threadIdx.x == 64) p.trace[ph * 2] = gtime();
      if (ph == NPH - 1 && threadIdx.x == 64) pdl_launch_dependents();
      LmGemm g;
      if (lm_gemm_of(p, ph, g)) {
        const int U = g.Mt * g.S;
        for (int u = cta; u < U; u += G) {
          const int mt = u / g.S, sp = u - mt * g.S, kb0 = sp * g.kbps, nkb = min(g.kbps, g.KB - kb0);
          if (warp == 1) {
            if (lane == 0) {
              if (uc > 0) {  // the workers have drained the accumulator of the previous unit (whose MMAs are long done)
                mbar_wait(tmem_empty, (uc - 1) & 1);
                tc_fence_after();
              }
              fence_proxy_async_global();  // the operand image was written by generic stores of other CTAs
              mbar_arrive_expect_tx(act_full, static_cast<uint32_t>(nkb) * LM_ATILE);
              bulk_g2s(act_s, g.act + static_cast<size_t>(kb0) * LM_ATILE, static_cast<uint32_t>(nkb) * LM_ATILE, act_full);
              mbar_wait(act_full, uc & 1);
              const uint32_t db0 = lm_desc_lo(smem_u32(act_s));
              for (int i = 0; i < nkb; ++i) {
                mbar_wait(full + ring_s, ring_par);
                tc_fence_after();
                const uint32_t da = lm_desc_lo(smem_u32(ring + ring_s * LM_WTILE));
                const uint32_t db = db0 + static_cast<uint32_t>(i) * (LM_ATILE >> 4);
#pragma unroll
                for (int k = 0; k < 4; ++k)  // +32 B along K inside the 128-B swizzle row = +2 in the 16-byte address field
                  lm_umma_f16(tmem_base, da + 2 * k, db + 2 * k, idesc, (i | k) != 0);
                umma_commit(empty + ring_s);
                if (++ring_s == LM_STAGES) { ring_s = 0; ring_par ^= 1; }
              }
              umma_commit(acc_full);
            } else {
              for (int i = 0; i < nkb; ++i)
                if (++ring_s == LM_STAGES) { ring_s = 0; ring_par ^= 1; }
            }
            __syncwarp();
          } else {
            // epilogue: TMEM [feature lane][row column] -> f32 partial rows in L2
            const int quad = warp & 3, half = (warp - 2) >> 2;
            const int f = mt * 128 + quad * 32 + lane;
            mbar_wait(acc_full, uc & 1);
            tc_fence_after();
            uint32_t v[32];
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + half * 32;
            tmem_ld32_wait(taddr, v);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty);
            const float bias = g.bias ? __ldg(g.bias + f) : 0.f;
            float* dst = g.out + (static_cast<size_t>(sp) * LM_ROWS + half * 32) * g.F + f;
            const int rmax = n - half * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < rmax) dst[static_cast<size_t>(j) * g.F] = __uint_as_float(v[j]) + bias;
          }
          ++uc;
        }
      } else if (warp == 1) {
        // attention phase: the otherwise idle MMA warp stages the shared voice prefix of this CTA's head (one bulk copy
        // for its K rows, one for its V rows)
        if (ph >= LM_PH_LAYER0 && ph < LM_PH_COND && ((ph - LM_PH_LAYER0) & 7) == 1 && lane == 0 && cta < LM_HEADS * NG &&
            cta / LM_HEADS < n) {
          const int l = (ph - LM_PH_LAYER0) >> 3, h = cta % LM_HEADS;
          const SeqDesc sd0 = p.row_desc[cta / LM_HEADS];
          if (!(p.flags & 1) && sd0.prefix_len > 0 && sd0.prefix_len <= LM_PFX) {
            const uint32_t bytes = static_cast<uint32_t>(sd0.prefix_len) * 128u;
            mbar_arrive_expect_tx(A.pfx_full, 2 * bytes);
            bulk_g2s(A.pfx_k, kv_row(sd0, l, 0, h, LM_HEADS, 0), bytes, A.pfx_full);
            bulk_g2s(A.pfx_v, kv_row(sd0, l, 1, h, LM_HEADS, 0), bytes, A.pfx_full);
          }
        }
        __syncwarp();
      } else {
        if (ph == 0) {
          // ---- AR feedback gather + noise + input_linear + LN1 of layer 0 (tts_model.rs:971,1065; flow_lm.rs:118,148-153)
          for (int r = cta; r < n; r += G) {
            const int slot = p.row_seq[r];
            if (wt == 0) {
              SeqDesc d = p.seqs[slot];
              d.pad = p.own_len[slot];
              p.row_desc[r] = d;
            }
            if (wt < 64) {
              float z = 0.f;
              if (wt < LDIM) {
                const StreamCtl c = p.ctl[slot];
                // a step enqueued ahead of the host may run one frame past the end: no noise row exists there
                if (c.noise) z = c.frame < c.max_gen_len ? c.noise[static_cast<long long>(c.frame) * LDIM + wt] : 0.f;
                else if (c.temp > 0.f) z = sqrtf(c.temp) * counter_normal(c.seed, c.frame, wt);
                p.z32[r * LDIM + wt] = z;
                lat_s[wt] = __half2float(__float2half_rn(p.feedback[slot * LDIM + wt]));
              }
              p.z16[r * 64 + wt] = __float2half_rn(z);
            }
            asm volatile("bar.sync 2, 256;" ::: "memory");
            const int f = wt * 4;
            float xv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4* wr = reinterpret_cast<const uint4*>(p.w_input + static_cast<size_t>(f + j) * 64);
              float a = 0.f;
#pragma unroll
              for (int c4 = 0; c4 < 4; ++c4) {
                const uint4 u4 = __ldg(wr + c4);
                const __half2* h2 = reinterpret_cast<const __half2*>(&u4);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 wf = __half22float2(h2[e]);
                  a += wf.x * lat_s[c4 * 8 + 2 * e] + wf.y * lat_s[c4 * 8 + 2 * e + 1];
                }
              }
              xv[j] = a;
            }
            const float4 v = make_float4(xv[0], xv[1], xv[2], xv[3]);
            *reinterpret_cast<float4*>(p.x32 + static_cast<size_t>(r) * LM_D + f) = v;
            lm_ln_to_image(v, p.ln1_w[0], p.ln1_b[0], p.hA, r, f, red_s, 0, wt);
            asm volatile("bar.sync 2, 256;" ::: "memory");  // lat_s / red_s free for the next row
          }
          __threadfence();
          fence_proxy_async_global();
        } else if (ph < LM_PH_COND) {
          const int l = (ph - LM_PH_LAYER0) >> 3, j = (ph - LM_PH_LAYER0) & 7;
          if (j == 1) {
            // ---- attention: this CTA's head, every NG-th row, one warp per (row, head)
            if (cta < LM_HEADS * NG && cta / LM_HEADS < n) {
              const int h = cta % LM_HEADS, g = cta / LM_HEADS, w = warp - 2;
              const SeqDesc sd0 = p.row_desc[g];
              const bool cta_staged = !(p.flags & 1) && sd0.prefix_len > 0 && sd0.prefix_len <= LM_PFX;
              const int SQ = p.shape[LM_G_INPROJ].S;
              for (int i = w; g + NG * i < n; i += 8) lm_attend_item(p, A, l, g + NG * i, h, w, lane, SQ, sd0, cta_staged, kv_cnt);
            }
            __threadfence();
            fence_proxy_async_global();
          } else if (j == 3 || j == 7) {
            // ---- x += sum of the split-K partials of out_proj / linear2; LayerNorm -> operand image of the next Linear
            const int S = p.shape[j == 3 ? LM_G_OUTPROJ : LM_G_LIN2].S;
            const bool last = (j == 7 && l == LM_LAYERS - 1);
            const float* lw = j == 3 ? p.ln2_w[l] : (last ? p.outnorm_w : p.ln1_w[last ? 0 : l + 1]);
            const float* lb = j == 3 ? p.ln2_b[l] : (last ? p.outnorm_b : p.ln1_b[last ? 0 : l + 1]);
            for (int r = cta; r < n; r += G) {
              const int f = wt * 4;
              float* xr = p.x32 + static_cast<size_t>(r) * LM_D + f;
              float4 v = ldcg_f4(xr);
              const float* w0 = p.ws + static_cast<size_t>(r) * LM_D + f;
              for (int s0 = 0; s0 < S; s0 += 16) {  // every partial of the row in flight before the first add
                float4 t[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) t[q] = ldcg_f4(w0 + static_cast<size_t>(min(s0 + q, S - 1)) * LM_ROWS * LM_D);
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                  const float on = (s0 + q < S) ? 1.f : 0.f;
                  v.x += on * t[q].x; v.y += on * t[q].y; v.z += on * t[q].z; v.w += on * t[q].w;
                }
              }
              *reinterpret_cast<float4*>(xr) = v;
              const float4 y = lm_ln_to_image(v, lw, lb, p.hA, r, f, red_s, 0, wt);
              if (last) {  // out_norm + EOS head (flow_lm.rs:132-145): the logit from the f32 row, not the f16 copy
                const float4 e4 = __ldg(reinterpret_cast<const float4*>(p.eos_w + f));
                const float dot = lm_block_sum(y.x * e4.x + y.y * e4.y + y.z * e4.z + y.w * e4.w, red_s, 2, wt);
                if (wt == 0) p.eos_logit[r] = dot + p.eos_b[0];
                if (p.h32dbg) *reinterpret_cast<float4*>(p.h32dbg + static_cast<size_t>(r) * LM_D + f) = y;
              }
              asm volatile("bar.sync 2, 256;" ::: "memory");  // red_s free for the next row
            }
            __threadfence();
            fence_proxy_async_global();
          } else {  // j == 5
            // ---- ffn = gelu(sum of linear1 partials) (tanh form, models/transformer.rs:85) -> operand image of linear2
            const int S = p.shape[LM_G_LIN1].S;
            for (int r = cta; r < n; r += G) {
              float4 v[4];
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4) v[q4] = make_float4(0.f, 0.f, 0.f, 0.f);
              for (int s0 = 0; s0 < S; s0 += 4) {
                float4 t[4][4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                  for (int q4 = 0; q4 < 4; ++q4)
                    t[q][q4] = ldcg_f4(p.ws + (static_cast<size_t>(min(s0 + q, S - 1)) * LM_ROWS + r) * LM_FFN + (q4 * LM_WORKERS + wt) * 4);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float on = (s0 + q < S) ? 1.f : 0.f;
#pragma unroll
                  for (int q4 = 0; q4 < 4; ++q4) { v[q4].x += on * t[q][q4].x; v[q4].y += on * t[q][q4].y; v[q4].z += on * t[q][q4].z; v[q4].w += on * t[q][q4].w; }
                }
              }
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4)
                lm_store_img4(p.ffnA, r, (q4 * LM_WORKERS + wt) * 4, gelu_tanh(v[q4].x), gelu_tanh(v[q4].y), gelu_tanh(v[q4].z),
                              gelu_tanh(v[q4].w));
            }
            __threadfence();
            fence_proxy_async_global();
          }
        } else if (ph == LM_PH_COND + 1) {
          // ---- c = cond_embed(h) + b (modules/mlp.rs:275); y_s = silu(c + te[s]) for every LSD step (mlp.rs:328-330)
          const int S = p.shape[LM_G_COND].S;
          for (int r = cta; r < n; r += G) {
            if (wt < LM_FLOW / 4) {
              const int f = wt * 4;
              const float* w0 = p.ws + static_cast<size_t>(r) * LM_FLOW + f;
              float4 v = __ldg(reinterpret_cast<const float4*>(p.b_cond + f));
              float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
              for (int s = 0; s < S; ++s) {
                const float4 t = ldcg_f4(w0 + static_cast<size_t>(s) * LM_ROWS * LM_FLOW);
                a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
              }
              v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
              *reinterpret_cast<float4*>(p.c32 + static_cast<size_t>(r) * LM_FLOW + f) = v;
              for (int s = 0; s < p.lsd_steps; ++s) {
                const float4 te = __ldg(reinterpret_cast<const float4*>(p.time_emb + s * LM_FLOW + f));
                lm_store_img4(p.yA + static_cast<size_t>(s) * 8 * LM_ATILE, r, f, silu(v.x + te.x), silu(v.y + te.y), silu(v.z + te.z),
                              silu(v.w + te.w));
              }
            }
          }
          __threadfence();
          fence_proxy_async_global();
        }
      }
      if (p.trace && cta == 0 && threadIdx.x == 64) p.trace[ph * 2 + 1] = gtime();
      if (ph + 1 < NPH) lm_grid_barrier(p.bar, bar_base + static_cast<unsigned long long>(ph + 1) * G, (p.flags & 2) != 0);
    }
    if (cta == 0 && threadIdx.x == 64) p.bar[1] = bar_base + static_cast<unsigned long long>(NPH - 1) * G;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 64);
}

// Weight [Fpad][K] f16 row-major -> tile images: tile (mt, kb) at (mt * K/64 + kb) * 16 KB; inside a tile row r (128 B)
// holds its eight 16-byte chunks at chunk ^ (r & 7) (what TMA SWIZZLE_128B would write).  One thread per chunk.
__global__ void lm_tile_weight_kernel(const __half* __restrict__ src, int Fpad, int K, uint8_t* __restrict__ dst) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(Fpad) * K / 8;
  if (idx >= total) return;
  const int chunks_per_row = K / 8;
  const int f = static_cast<int>(idx / chunks_per_row), c = static_cast<int>(idx - static_cast<long long>(f) * chunks_per_row);
  const int mt = f >> 7, r = f & 127, kb = c >> 3, c8 = c & 7;
  const uint4 v = *reinterpret_cast<const uint4*>(src + static_cast<long long>(f) * K + c * 8);
  *reinterpret_cast<uint4*>(dst + (static_cast<long long>(mt) * (K / 64) + kb) * LM_WTILE + r * 128 + ((c8 ^ (r & 7)) << 4)) = v;
}

}  // namespace ptts
