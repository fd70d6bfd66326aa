// Small-batch Linear: D[r, f] = sum_k A[r, k] * W[f, k] for 1-4 activation rows (a single utterance, BASELINE configs[0]).
//
// With so few rows a Linear layer of the FlowLM decode step (reference modules/attention.rs:129,280,
// models/transformer.rs:85) is a pure weight stream -- 2 to 8 MB read once, a few MFLOP -- and the tensor-core path spends
// its time on everything but the stream: 24-64 CTAs each ingesting 128-256 KB through one SM's TMA port, an accumulator
// round trip through TMEM and a split-K exchange.  This kernel is the bandwidth-bound form:
//   * one CTA of 8 warps per SM (two fit: the second slot is where the NEXT launch's CTAs sit with their weights already
//     requested), so the stream is pulled through all 148 SM ports at once;
//   * a warp owns whole features; its lanes walk a weight row in 16-byte chunks (128-bit loads, consecutive lanes on
//     consecutive chunks: every request is a full 512-byte run);
//   * the first batch of weight rows (up to 16 chunks = 256 bytes per lane, 4.8 MB per grid) is requested BEFORE
//     griddepcontrol.wait: weights do not depend on the previous kernel, so the stream of GEMV n+1 is already in flight
//     while GEMV n (or the LayerNorm / attention launch between them) still runs;
//   * the activation rows (f16, at most 32 KB) are staged in shared memory once per CTA after the wait;
//   * f32 accumulation, warp-shuffle reduce-scatter (31 shuffles for 32 partial sums instead of 160), and the same
//     epilogue contract as the tensor-core GEMM (gemm.cuh GemmEpi: int8 scale, bias, activation, alpha, LayerScale, gate,
//     residual, f32 / f16 outputs), one output element per lane;
//   * int8 storage (reference quantize.rs:65-94): the row is streamed as one byte per code -- half the traffic -- and
//     expanded in registers (codes are exact in f16); the per-tensor scale stays in the epilogue.
// Roofline: HBM.  Algorithmic bytes per launch = F*K*(2 | 1) + rows*K*2 + epilogue tensors.
#pragma once
#include "gemm.cuh"

namespace ptts {

struct GemvParams {
  const void* w;     // f16 [Fpad][K] K-major, or int8 codes [Fpad][K]
  const __half* x;   // [rows][K] f16 (unused when ln_x is set)
  int F, K, rows;
  GemmEpi epi;
  // fused LayerNorm prologue (K = 1024 only): the operand is LN(ln_x) * ln_w + ln_b, computed per CTA exactly as
  // ln_rows_kernel does (one warp per row, two-pass, biased variance, eps inside the sqrt) and rounded to f16 once
  const float* ln_x;   // [rows][K] f32 residual stream, or null
  const float* ln_w;
  const float* ln_b;
  float ln_eps;
};

static constexpr int GEMV_THREADS = 256;
static constexpr int GEMV_WARPS = GEMV_THREADS / 32;
static constexpr int GEMV_MAX_ROWS = 4;
static constexpr int GEMV_PRELOAD = 16;   // 16-byte chunks per lane in flight across the PDL wait

__device__ __forceinline__ uint4 ldg_stream_u4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void h8_to_f8(const uint4& v, float (&f)[8]) {
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
  const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&v.z));
  const float2 d = __half22float2(*reinterpret_cast<const __half2*>(&v.w));
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

// Sum N (a power of two <= 32) per-lane partial values over the warp.  Each exchange halves the values a lane still
// carries, so afterwards v[0] of lane L is the warp total of value L >> (5 - log2 N); lanes sharing that index agree.
template <int N>
__device__ __forceinline__ void warp_reduce_scatter(float (&v)[N], int lane) {
  int n = N;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    if (n > 1) {
      const int half = n >> 1;
      const bool upper = (lane & o) != 0;
#pragma unroll
      for (int k = 0; k < N / 2; ++k) {
        if (k < half) {
          const float send = upper ? v[k] : v[k + half];
          const float keep = upper ? v[k + half] : v[k];
          v[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
      }
      n = half;
    } else {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
    }
  }
}

template <int ROWS, int KCH, bool INT8>
struct GemvGeo {
  static constexpr int EPC = INT8 ? 16 : 8;                  // weight elements per 16-byte chunk
  static constexpr int NF0 = GEMV_PRELOAD / KCH < 16 / ROWS ? GEMV_PRELOAD / KCH : 16 / ROWS;
  static constexpr int NF = NF0 > 8 ? 8 : NF0;               // features per warp per batch
  static constexpr int NV = NF * ROWS;                       // partial sums per lane
};

// K = KCH * 32 * (8 | 16).  Feature (batch b, slot i) of global warp gw is (b * NF + i) * W + gw, W = warps in the grid.
template <int ROWS, int KCH, bool INT8>
__global__ void __launch_bounds__(GEMV_THREADS, 2) gemv_rows_kernel(const GemvParams p) {
  using G = GemvGeo<ROWS, KCH, INT8>;
  constexpr int NF = G::NF, NV = G::NV;
  extern __shared__ __align__(16) uint8_t gemv_smem[];
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int W = gridDim.x * GEMV_WARPS;
  const int gw = warp * gridDim.x + blockIdx.x;   // consecutive features on consecutive SMs: a short F still loads every SM evenly
  const int F = p.F, K = p.K;
  const size_t row_bytes = INT8 ? (size_t)K : (size_t)K * 2;
  const uint8_t* wbase = reinterpret_cast<const uint8_t*>(p.w) + (size_t)lane * 16;

  uint4 wreg[NF][KCH];
  auto load_batch = [&](int batch) {
#pragma unroll
    for (int i = 0; i < NF; ++i) {
      const int f = (batch * NF + i) * W + gw;
      if (f < F) {
        const uint8_t* row = wbase + (size_t)f * row_bytes;
#pragma unroll
        for (int j = 0; j < KCH; ++j) wreg[i][j] = ldg_stream_u4(row + (size_t)j * 512);
      }
    }
  };
  load_batch(0);        // independent of the previous kernel: in flight across the dependency
  // so are the per-feature epilogue constants of the element this lane will write for batch 0 (int8 scale, bias, LayerScale):
  // requested here they cost nothing, behind the dependency each would be one more L2 round trip at the tail of the kernel
  const GemmEpi& e = p.epi;
  float c_ws = 1.f, c_bias = 0.f, c_fs = 1.f;
  {
    const int idx0 = lane / (32 / NV), i0 = idx0 / ROWS;
    const int f0 = i0 * W + gw;
    if (f0 < F) {
      if (e.wscale) c_ws = __ldg(e.wscale + f0);
      if (e.bias) c_bias = __ldg(e.bias + f0);
      if (e.fscale) c_fs = __ldg(e.fscale + f0);
    }
  }
  constexpr bool LN_OK = (INT8 ? KCH * 512 : KCH * 256) == 1024;
  const bool fused_ln = LN_OK && p.ln_x != nullptr;
  float* lnw_s = reinterpret_cast<float*>(gemv_smem + ROWS * 1024 * 2);
  float* lnb_s = lnw_s + 1024;
  if (LN_OK && fused_ln) {   // the affine parameters are constants too
    for (int i = threadIdx.x; i < 1024 / 4; i += GEMV_THREADS) {
      reinterpret_cast<float4*>(lnw_s)[i] = __ldg(reinterpret_cast<const float4*>(p.ln_w) + i);
      reinterpret_cast<float4*>(lnb_s)[i] = __ldg(reinterpret_cast<const float4*>(p.ln_b) + i);
    }
    __syncthreads();
  }
  pdl_wait();

  if (LN_OK && fused_ln) {
    if (warp < ROWS) {
      constexpr int C = 1024, PER = C / 32;
      const int r = warp;
      float v[PER];
      if (r < p.rows) {
        const float4* xr = reinterpret_cast<const float4*>(p.ln_x + static_cast<long long>(r) * C);
#pragma unroll
        for (int i = 0; i < PER / 4; ++i) {
          const float4 t = __ldcg(xr + i * 32 + lane);
          v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < PER; ++i) v[i] = 0.f;
      }
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < PER; ++i) sum += v[i];
      const float mean = warp_sum(sum) * (1.f / C);
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < PER; ++i) { const float d = v[i] - mean; q += d * d; }
      const float rstd = 1.f / sqrtf(warp_sum(q) * (1.f / C) + p.ln_eps);
#pragma unroll
      for (int i = 0; i < PER / 4; ++i) {
        const int c = (i * 32 + lane) * 4;
        const float4 w4 = *reinterpret_cast<const float4*>(lnw_s + c), b4 = *reinterpret_cast<const float4*>(lnb_s + c);
        const float wv[4] = {w4.x, w4.y, w4.z, w4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float y = (v[4 * i + j] - mean) * rstd;
          y = y * wv[j] + bv[j];
          o[j] = r < p.rows ? y : 0.f;
        }
        const __half2 h0 = __floats2half2_rn(o[0], o[1]), h1 = __floats2half2_rn(o[2], o[3]);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&h0);
        pk.y = *reinterpret_cast<const uint32_t*>(&h1);
        const int g = c >> 3;
        const int dst = INT8 ? r * K * 2 + (g & 1) * K + (g >> 1) * 16 + (c & 7) * 2 : r * K * 2 + c * 2;
        *reinterpret_cast<uint2*>(gemv_smem + dst) = pk;
      }
    }
  } else
  // activation rows -> shared memory.  f16 weights: the row as it is.  int8: a lane's 16 codes need 16 halves of x, kept as
  // two 16-byte planes per row (even / odd 8-element groups) so that both reads of a warp are conflict-free runs.
  {
    const int groups = ROWS * (K / 8);
    for (int g = threadIdx.x; g < groups; g += GEMV_THREADS) {
      const int r = g / (K / 8), c = g - r * (K / 8);
      uint4 v = make_uint4(0, 0, 0, 0);
      if (r < p.rows) v = __ldcg(reinterpret_cast<const uint4*>(p.x + (size_t)r * K) + c);
      const int dst = INT8 ? r * K * 2 + (c & 1) * K + (c >> 1) * 16 : r * K * 2 + c * 16;
      *reinterpret_cast<uint4*>(gemv_smem + dst) = v;
    }
  }
  __syncthreads();
  const uint32_t xs = smem_u32(gemv_smem);
  const int nbatch = (F + W * NF - 1) / (W * NF);
  for (int batch = 0; batch < nbatch; ++batch) {
    if (batch > 0) load_batch(batch);
    // features of a warp come in increasing order, so the valid ones are a prefix of the batch (warp-uniform): the slots
    // past the end of F are skipped rather than multiplied (in_proj on 1 184 warps fills 2.6 of 4 slots, of 8 with byte codes)
    int nvalid = 0;
#pragma unroll
    for (int i = 0; i < NF; ++i) nvalid += ((batch * NF + i) * W + gw < F) ? 1 : 0;
    float acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = 0.f;
#pragma unroll
    for (int j = 0; j < KCH; ++j) {
      if (!INT8) {
        float xf[ROWS][8];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) h8_to_f8(lds_u4(xs + r * K * 2 + (j * 32 + lane) * 16), xf[r]);
#pragma unroll
        for (int i = 0; i < NF; ++i) {
          if (i < nvalid) {
            float wf[8];
            h8_to_f8(wreg[i][j], wf);
#pragma unroll
            for (int r = 0; r < ROWS; ++r)
#pragma unroll
              for (int q = 0; q < 8; ++q) acc[i * ROWS + r] = fmaf(wf[q], xf[r][q], acc[i * ROWS + r]);
          }
        }
      } else {
#pragma unroll
        for (int h = 0; h < 2; ++h) {   // codes 0-7 and 8-15 of the chunk
          float xf[ROWS][8];
#pragma unroll
          for (int r = 0; r < ROWS; ++r) h8_to_f8(lds_u4(xs + r * K * 2 + h * K + (j * 32 + lane) * 16), xf[r]);
#pragma unroll
          for (int i = 0; i < NF; ++i) {
            if (i >= nvalid) continue;
            uint4 w16;
            i8x4_to_f16x4(h ? wreg[i][j].z : wreg[i][j].x, w16.x, w16.y);
            i8x4_to_f16x4(h ? wreg[i][j].w : wreg[i][j].y, w16.z, w16.w);
            float wf[8];
            h8_to_f8(w16, wf);
#pragma unroll
            for (int r = 0; r < ROWS; ++r)
#pragma unroll
              for (int q = 0; q < 8; ++q) acc[i * ROWS + r] = fmaf(wf[q], xf[r][q], acc[i * ROWS + r]);
          }
        }
      }
    }
    warp_reduce_scatter<NV>(acc, lane);
    constexpr int SHARE = 32 / NV;            // lanes holding the same total
    const int idx = lane / SHARE;
    const int i = idx / ROWS, r = idx - i * ROWS;
    const int f = (batch * NF + i) * W + gw;
    if ((lane % SHARE) == 0 && f < F && r < p.rows) {
      float v = acc[0];
      if (batch > 0) {
        if (e.wscale) c_ws = __ldg(e.wscale + f);
        if (e.bias) c_bias = __ldg(e.bias + f);
        if (e.fscale) c_fs = __ldg(e.fscale + f);
      }
      if (e.wscale) v *= c_ws;
      if (e.bias) v += c_bias;
      v = epi_act(e.act, v) * e.alpha;
      if (e.fscale) v *= c_fs;
      if (e.gate) v *= e.gate[row_off(e.gate_map, r) + f];
      if (e.res) v += e.res[row_off(e.res_map, r) + f];
      if (e.out32) e.out32[row_off(e.out32_map, r) + f] = v;
      if (e.out16) e.out16[row_off(e.out16_map, r) + f] = __float2half_rn(e.act16 == ACT_ELU ? elu1(v) : v);
    }
  }
}

}  // namespace ptts
