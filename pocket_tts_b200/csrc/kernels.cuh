// HBM-bound kernels of the hot path (everything that is not a matrix product): norms, RoPE +
// KV-cache attention, the Mimi front end, the last SEANet conv, per-stream state movement and
// the EOS bookkeeping.  128-bit loads where rows are contiguous, warp-shuffle reductions,
// f32 math throughout; f16 only as storage for GEMM operands and KV rows.
#pragma once
#include "ptx.cuh"

namespace ptts {

static constexpr int HD = 64;           // head dim of both transformers
static constexpr int LDIM = 32;         // latent dim
static constexpr int MIMI_RING = 272;   // >= context(250) + 16 new rows - 1, see mimi_attn_kernel
static constexpr int MIMI_CTX = 250;    // config/b6369a24.yaml:49
static constexpr int FRAME = 1920;

__constant__ float c_inv_freq[HD / 2];  // max_period^(-2i/64), reference modules/rope.rs:9-16

struct StreamCtl {       // per slot, device resident
  int max_gen_len;
  int frames_after_eos;
  float eos_threshold;
  float temp;
  unsigned long long seed;
  const float* noise;    // device [max_gen_len, 32] or null
  int frame;             // frames generated so far == index of the step being run
  int eos_step;          // -1 until the first logit above threshold
  int finished;
  int pad;
};

// One resident sequence: its own KV rows plus an immutable shared prefix (the voice).
// Layout of either buffer: [layer][k|v][head][cap][64] f16.
struct SeqDesc {
  __half* own;
  const __half* prefix;
  int own_cap;
  int prefix_len;
  int prefix_cap;
  int pad;
};

__device__ __forceinline__ const __half* kv_row(const SeqDesc& s, int layer, int kv, int head, int n_heads, int i) {
  if (i < s.prefix_len)
    return s.prefix + ((static_cast<long long>(layer * 2 + kv) * n_heads + head) * s.prefix_cap + i) * HD;
  return s.own + ((static_cast<long long>(layer * 2 + kv) * n_heads + head) * s.own_cap + (i - s.prefix_len)) * HD;
}

// ---------------------------------------------------------------- LayerNorm family
// One warp per row.  out16 = LN(x) [* w + b] [* (1 + scale) + shift]   (reference modules/mlp.rs:29-58,135-137;
// biased variance, eps inside the sqrt).  shift/scale rows live in the flow head's modulation buffer.
template <int C>
__global__ void ln_rows_kernel(const float* __restrict__ x, int rows, const float* __restrict__ w,
                               const float* __restrict__ b, float eps, const float* __restrict__ shift,
                               const float* __restrict__ scale, int mod_ld, __half* __restrict__ out16, int out_ld) {
  pdl_launch_dependents();
  constexpr int PER = C / 32;
  const int lane = threadIdx.x & 31;
  // the affine parameters are constants: requested before the dependency on the producer of x resolves, so that the
  // kernel is one memory round trip (x) instead of two behind griddepcontrol.wait
  float4 w4[PER / 4], b4[PER / 4];
  if (w) {
#pragma unroll
    for (int i = 0; i < PER / 4; ++i) {
      w4[i] = __ldg(reinterpret_cast<const float4*>(w) + i * 32 + lane);
      b4[i] = __ldg(reinterpret_cast<const float4*>(b) + i * 32 + lane);
    }
  }
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * C);
  float v[PER];
#pragma unroll
  for (int i = 0; i < PER / 4; ++i) {
    float4 t = xr[i * 32 + lane];
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) { float d = v[i] - mean; q += d * d; }
  const float rstd = 1.f / sqrtf(warp_sum(q) * (1.f / C) + eps);
#pragma unroll
  for (int i = 0; i < PER / 4; ++i) {
    const int c = (i * 32 + lane) * 4;
    const float wv[4] = {w4[i].x, w4[i].y, w4[i].z, w4[i].w}, bv[4] = {b4[i].x, b4[i].y, b4[i].z, b4[i].w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float y = (v[4 * i + j] - mean) * rstd;
      if (w) y = y * wv[j] + bv[j];
      if (scale) y = y * (1.f + scale[static_cast<long long>(row) * mod_ld + c + j]) +
                     shift[static_cast<long long>(row) * mod_ld + c + j];
      o[j] = y;
    }
    __half2 h0 = __floats2half2_rn(o[0], o[1]), h1 = __floats2half2_rn(o[2], o[3]);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&h0);
    pk.y = *reinterpret_cast<uint32_t*>(&h1);
    *reinterpret_cast<uint2*>(out16 + static_cast<long long>(row) * out_ld + c) = pk;
  }
}

// out_norm + EOS head (reference models/flow_lm.rs:132-145): h16 = LN(x) for the flow head's cond_embed,
// logit = <LN(x), w_eos> + b_eos in f32 (not from the f16 copy: the threshold compare must not see rounding).
__global__ void ln_eos_kernel(const float* __restrict__ x, int rows, const float* __restrict__ w,
                              const float* __restrict__ b, const float* __restrict__ w_eos, const float* __restrict__ b_eos,
                              __half* __restrict__ h16, float* __restrict__ h32, float* __restrict__ eos_logit) {
  pdl_launch_dependents();
  constexpr int C = 1024, PER = 32;
  const int lane = threadIdx.x & 31;
  float4 w4[PER / 4], b4[PER / 4], e4[PER / 4];   // constants, requested before the dependency resolves
#pragma unroll
  for (int i = 0; i < PER / 4; ++i) {
    w4[i] = __ldg(reinterpret_cast<const float4*>(w) + i * 32 + lane);
    b4[i] = __ldg(reinterpret_cast<const float4*>(b) + i * 32 + lane);
    e4[i] = __ldg(reinterpret_cast<const float4*>(w_eos) + i * 32 + lane);
  }
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * C);
  float v[PER];
#pragma unroll
  for (int i = 0; i < PER / 4; ++i) {
    float4 t = xr[i * 32 + lane];
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) { float d = v[i] - mean; q += d * d; }
  const float rstd = 1.f / sqrtf(warp_sum(q) * (1.f / C) + 1e-5f);
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < PER / 4; ++i) {
    const int c = (i * 32 + lane) * 4;
    const float wv[4] = {w4[i].x, w4[i].y, w4[i].z, w4[i].w}, bv[4] = {b4[i].x, b4[i].y, b4[i].z, b4[i].w};
    const float ev[4] = {e4[i].x, e4[i].y, e4[i].z, e4[i].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float y = (v[4 * i + j] - mean) * rstd * wv[j] + bv[j];
      dot += y * ev[j];
      h16[static_cast<long long>(row) * C + c + j] = __float2half_rn(y);
      if (h32) h32[static_cast<long long>(row) * C + c + j] = y;
    }
  }
  dot = warp_sum(dot);
  if (lane == 0) eos_logit[row] = dot + b_eos[0];
}

// ---------------------------------------------------------------- embedding gather (conditioners/text.rs:289-303)
__global__ void embed_rows_kernel(const int* __restrict__ tokens, int rows, const float* __restrict__ lut,
                                  float* __restrict__ x) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x;
  const float4* src = reinterpret_cast<const float4*>(lut + static_cast<long long>(tokens[row]) * 1024);
  float4* dst = reinterpret_cast<float4*>(x + static_cast<long long>(row) * 1024);
  dst[threadIdx.x] = src[threadIdx.x];  // 256 threads x 16 B
}

// ---------------------------------------------------------------- RoPE helper (modules/rope.rs:18-60)
// pair i of a head: (x[2i], x[2i+1]) rotated by pos * inv_freq[i]
__device__ __forceinline__ void rope_pair(float xr, float xi, int pos, int i, float& o_r, float& o_i) {
  float s, c;
  sincosf(static_cast<float>(pos) * c_inv_freq[i], &s, &c);
  o_r = xr * c - xi * s;
  o_i = xr * s + xi * c;
}

// softmax(q K^T / 8) V over keys [0, n_keys) of one (sequence, layer, head), single pass with a running maximum
// (the naive reference path, modules/sdpa.rs:36-82, up to f32 rounding; causality is the caller's key range).
// Block = ATTN_THREADS (4 warps).  Eight lanes share one key: each loads 16 bytes of the K row and 16 bytes of the
// V row, the partial dots are folded with three shuffles; a lane group has 4 keys (8 loads) in flight before the
// first use, so a warp streams 16 keys x 256 bytes per iteration (one loop trip per memory round trip was the limit).
// q in smem (f32, already rotated).  Result: 64 floats at red_s[0..63].  red_s needs 64 + 4*66 floats.
static constexpr int ATTN_THREADS = 128;
__device__ __forceinline__ void attend_block(const SeqDesc& sd, int layer, int head, int n_heads, const float* q_s,
                                             float* red_s, int n_keys) {
  constexpr int NW = ATTN_THREADS / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane >> 3, part = lane & 7;
  float q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) q[j] = q_s[part * 8 + j] * 0.125f;  // 1/sqrt(64) folded into q
  float m = -INFINITY, l = 0.f, acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  constexpr int U = 4;  // keys per lane group in flight: 8 x 16-byte loads before the first use
  for (int i0 = warp * 4 * U; i0 < n_keys; i0 += NW * 4 * U) {
    uint4 ku[U], vu[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * 4 + sub;
      ku[u] = vu[u] = make_uint4(0, 0, 0, 0);
      if (i < n_keys) {
        ku[u] = reinterpret_cast<const uint4*>(kv_row(sd, layer, 0, head, n_heads, i))[part];
        vu[u] = reinterpret_cast<const uint4*>(kv_row(sd, layer, 1, head, n_heads, i))[part];
      }
    }
    float sc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const __half2* kh = reinterpret_cast<const __half2*>(&ku[u]);
      sc[u] = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(kh[j]);
        sc[u] += f.x * q[2 * j] + f.y * q[2 * j + 1];
      }
    }
#pragma unroll
    for (int x = 1; x <= 4; x <<= 1)
#pragma unroll
      for (int u = 0; u < U; ++u) sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], x);
    // one rescale for the group of U keys (keys past the end score -inf and weigh 0)
    float m_new = m;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i0 + u * 4 + sub >= n_keys) sc[u] = -INFINITY;
      m_new = fmaxf(m_new, sc[u]);
    }
    if (m_new != -INFINITY) {
      const float corr = expf(m - m_new);  // exp(-inf) = 0 on the first group
      l *= corr;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] *= corr;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float p = expf(sc[u] - m_new);
        l += p;
        const __half2* vh = reinterpret_cast<const __half2*>(&vu[u]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(vh[j]);
          acc[2 * j] += p * f.x;
          acc[2 * j + 1] += p * f.y;
        }
      }
      m = m_new;
    }
  }
  // fold the 4 key sub-groups of the warp, then the 4 warps, always in the same order
#pragma unroll
  for (int x = 8; x <= 16; x <<= 1) {
    const float m_o = __shfl_xor_sync(0xffffffffu, m, x);
    const float l_o = __shfl_xor_sync(0xffffffffu, l, x);
    const float m_new = fmaxf(m, m_o);
    const float ca = (m == -INFINITY) ? 0.f : expf(m - m_new);
    const float cb = (m_o == -INFINITY) ? 0.f : expf(m_o - m_new);
    l = l * ca + l_o * cb;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float a_o = __shfl_xor_sync(0xffffffffu, acc[j], x);
      acc[j] = acc[j] * ca + a_o * cb;
    }
    m = m_new;
  }
  float* wsm = red_s + 64 + warp * 66;
  if (sub == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) wsm[part * 8 + j] = acc[j];
    if (part == 0) { wsm[64] = m; wsm[65] = l; }
  }
  __syncthreads();
  if (tid < 64) {
    float gm = -INFINITY;
#pragma unroll
    for (int w = 0; w < NW; ++w) gm = fmaxf(gm, red_s[64 + w * 66 + 64]);
    float o = 0.f, lt = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const float mw = red_s[64 + w * 66 + 64];
      const float c = (mw == -INFINITY) ? 0.f : expf(mw - gm);
      o += red_s[64 + w * 66 + tid] * c;
      lt += red_s[64 + w * 66 + 65] * c;
    }
    red_s[tid] = o / lt;
  }
  __syncthreads();
}

// FlowLM decode attention, one new row per stream (reference modules/attention.rs:104-231 with t = 1):
// RoPE(q,k) at the absolute position, append K,V at the cursor, causal SDPA over prefix + own rows, all fused.
// grid (n, heads), block ATTN_THREADS.  The rows already in the cache do not depend on this step's projections, so every
// warp requests its first 16 keys the moment the row descriptor is known; the q / k / v loads, RoPE and the block barrier
// of warp 0 then run inside that memory round trip instead of in front of it.  The new key never goes through the cache:
// it joins the merge of the warps' partial results as a fifth partial (rounded to f16 like the row later steps will read).
__global__ void __launch_bounds__(ATTN_THREADS, 7) flowlm_attn_decode_kernel(const float* __restrict__ qkv,
                                          const SeqDesc* __restrict__ row_desc /*[n], cursor in .pad (step_begin_kernel)*/,
                                          int layer, int n_heads, __half* __restrict__ out16) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int NW = ATTN_THREADS / 32, U = 4;
  __shared__ float sm[64 + 64 + 64 + 4 * 66 + 2];
  float* q_s = sm;            // rotated q
  float* kn_s = sm + 64;      // the new key (rotated, f16-rounded), later the merged output
  float* vn_s = sm + 128;     // the new value (f16-rounded)
  float* red_s = sm + 192;    // [4][66] per-warp partials
  const int b = blockIdx.x, h = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane >> 3, part = lane & 7;
  const int d_model = n_heads * HD;
  const SeqDesc sd = row_desc[b];
  const int pos = sd.prefix_len + sd.pad;   // keys already in the cache = position of the new row
  uint4 ku[U], vu[U];
  int i0 = warp * 4 * U;
  auto request = [&](int base) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = base + u * 4 + sub;
      ku[u] = vu[u] = make_uint4(0, 0, 0, 0);
      if (i < pos) {
        ku[u] = reinterpret_cast<const uint4*>(kv_row(sd, layer, 0, h, n_heads, i))[part];
        vu[u] = reinterpret_cast<const uint4*>(kv_row(sd, layer, 1, h, n_heads, i))[part];
      }
    }
  };
  request(i0);
  if (tid < 32) {
    const float* row = qkv + static_cast<long long>(b) * 3 * d_model;
    const float2 qx = *reinterpret_cast<const float2*>(row + h * HD + 2 * tid);
    const float2 kx = *reinterpret_cast<const float2*>(row + d_model + h * HD + 2 * tid);
    const float2 vx = *reinterpret_cast<const float2*>(row + 2 * d_model + h * HD + 2 * tid);
    float sn, cs;
    sincosf(static_cast<float>(pos) * c_inv_freq[tid], &sn, &cs);   // modules/rope.rs:18-60, one angle for q and k
    q_s[2 * tid] = qx.x * cs - qx.y * sn;
    q_s[2 * tid + 1] = qx.x * sn + qx.y * cs;
    const __half2 kh = __floats2half2_rn(kx.x * cs - kx.y * sn, kx.x * sn + kx.y * cs), vh = __floats2half2_rn(vx.x, vx.y);
    reinterpret_cast<__half2*>(const_cast<__half*>(kv_row(sd, layer, 0, h, n_heads, pos)))[tid] = kh;
    reinterpret_cast<__half2*>(const_cast<__half*>(kv_row(sd, layer, 1, h, n_heads, pos)))[tid] = vh;
    const float2 kf = __half22float2(kh), vf = __half22float2(vh);
    kn_s[2 * tid] = kf.x; kn_s[2 * tid + 1] = kf.y;
    vn_s[2 * tid] = vf.x; vn_s[2 * tid + 1] = vf.y;
  }
  __syncthreads();
  float q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) q[j] = q_s[part * 8 + j] * 0.125f;  // 1/sqrt(64) folded into q
  float m = -INFINITY, l = 0.f, acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (; i0 < pos; i0 += NW * 4 * U) {
    float sc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const __half2* kh = reinterpret_cast<const __half2*>(&ku[u]);
      sc[u] = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(kh[j]);
        sc[u] += f.x * q[2 * j] + f.y * q[2 * j + 1];
      }
    }
#pragma unroll
    for (int x = 1; x <= 4; x <<= 1)
#pragma unroll
      for (int u = 0; u < U; ++u) sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], x);
    // one rescale for the group of U keys (keys past the end score -inf and weigh 0)
    float m_new = m;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i0 + u * 4 + sub >= pos) sc[u] = -INFINITY;
      m_new = fmaxf(m_new, sc[u]);
    }
    if (m_new != -INFINITY) {
      const float corr = expf(m - m_new);  // exp(-inf) = 0 on the first group
      l *= corr;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] *= corr;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float pw = expf(sc[u] - m_new);
        l += pw;
        const __half2* vh = reinterpret_cast<const __half2*>(&vu[u]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(vh[j]);
          acc[2 * j] += pw * f.x;
          acc[2 * j + 1] += pw * f.y;
        }
      }
      m = m_new;
    }
    if (i0 + NW * 4 * U < pos) request(i0 + NW * 4 * U);
  }
  // fold the 4 key sub-groups of the warp, then the 4 warps and the new key, always in the same order
#pragma unroll
  for (int x = 8; x <= 16; x <<= 1) {
    const float m_o = __shfl_xor_sync(0xffffffffu, m, x);
    const float l_o = __shfl_xor_sync(0xffffffffu, l, x);
    const float m_new = fmaxf(m, m_o);
    const float ca = (m == -INFINITY) ? 0.f : expf(m - m_new);
    const float cb = (m_o == -INFINITY) ? 0.f : expf(m_o - m_new);
    l = l * ca + l_o * cb;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float a_o = __shfl_xor_sync(0xffffffffu, acc[j], x);
      acc[j] = acc[j] * ca + a_o * cb;
    }
    m = m_new;
  }
  float* wsm = red_s + warp * 66;
  if (sub == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) wsm[part * 8 + j] = acc[j];
    if (part == 0) { wsm[64] = m; wsm[65] = l; }
  }
  if (warp == 0) {  // score of the new key: q . k_new / 8
    float sn = 0.f;
    if (sub == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) sn += q[j] * kn_s[part * 8 + j];
    }
#pragma unroll
    for (int x = 1; x <= 4; x <<= 1) sn += __shfl_xor_sync(0xffffffffu, sn, x);
    if (lane == 0) sm[192 + 4 * 66] = sn;
  }
  __syncthreads();
  if (tid < 64) {
    const float s_new = sm[192 + 4 * 66];
    float gm = s_new;
#pragma unroll
    for (int w = 0; w < NW; ++w) gm = fmaxf(gm, red_s[w * 66 + 64]);
    float o = 0.f, lt = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const float mw = red_s[w * 66 + 64];
      const float c = (mw == -INFINITY) ? 0.f : expf(mw - gm);
      o += red_s[w * 66 + tid] * c;
      lt += red_s[w * 66 + 65] * c;
    }
    const float cn = expf(s_new - gm);
    o += vn_s[tid] * cn;
    lt += cn;
    out16[static_cast<long long>(b) * d_model + h * HD + tid] = __float2half_rn(o / lt);
  }
}

// Prefill, step 1: RoPE + KV append for every new row (rows of several sequences at once); rotated q kept in f32.
// grid (rows, heads / 4), block 128: one warp per head (a 32-thread CTA per (row, head) was 40 960 CTAs for a 64 x 40-token
// open: 48 us of CTA scheduling per layer for 10 MB of traffic).
__global__ void flowlm_rope_append_kernel(const float* __restrict__ qkv, const int* __restrict__ row_seq,
                                          const int* __restrict__ row_pos, const SeqDesc* __restrict__ seqs, int layer,
                                          int n_heads, float* __restrict__ q_rot) {
  pdl_launch_dependents();
  pdl_wait();
  const int r = blockIdx.x, h = blockIdx.y * 4 + (threadIdx.x >> 5), tid = threadIdx.x & 31;
  if (h >= n_heads) return;
  const int d_model = n_heads * HD;
  const SeqDesc sd = seqs[row_seq[r]];
  const int pos = row_pos[r];
  const float* row = qkv + static_cast<long long>(r) * 3 * d_model;
  float qr, qi, kr, ki;
  rope_pair(row[h * HD + 2 * tid], row[h * HD + 2 * tid + 1], pos, tid, qr, qi);
  rope_pair(row[d_model + h * HD + 2 * tid], row[d_model + h * HD + 2 * tid + 1], pos, tid, kr, ki);
  q_rot[static_cast<long long>(r) * d_model + h * HD + 2 * tid] = qr;
  q_rot[static_cast<long long>(r) * d_model + h * HD + 2 * tid + 1] = qi;
  __half* kd = const_cast<__half*>(kv_row(sd, layer, 0, h, n_heads, pos));
  __half* vd = const_cast<__half*>(kv_row(sd, layer, 1, h, n_heads, pos));
  reinterpret_cast<__half2*>(kd)[tid] = __floats2half2_rn(kr, ki);
  reinterpret_cast<__half2*>(vd)[tid] =
      __floats2half2_rn(row[2 * d_model + h * HD + 2 * tid], row[2 * d_model + h * HD + 2 * tid + 1]);
}

// Prefill, step 2: causal attention of each new row over keys [0, pos] (reference sdpa.rs:36-82 with the
// causal mask of :129-171; shift = Lk - Lq makes row i see keys <= its absolute position).
__global__ void flowlm_attn_prefill_kernel(const float* __restrict__ q_rot, const int* __restrict__ row_seq,
                                           const int* __restrict__ row_pos, const SeqDesc* __restrict__ seqs, int layer,
                                           int n_heads, __half* __restrict__ out16) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float sm[64 + 64 + 4 * 66];
  float* q_s = sm;
  float* red_s = sm + 64;
  const int r = blockIdx.x, h = blockIdx.y, tid = threadIdx.x;
  const int d_model = n_heads * HD;
  const SeqDesc sd = seqs[row_seq[r]];
  const int pos = row_pos[r];
  if (tid < 64) q_s[tid] = q_rot[static_cast<long long>(r) * d_model + h * HD + tid];
  __syncthreads();
  attend_block(sd, layer, h, n_heads, q_s, red_s, pos + 1);
  if (tid < 64) out16[static_cast<long long>(r) * d_model + h * HD + tid] = __float2half_rn(red_s[tid]);
}

// Prefill attention on tensor cores (reference modules/sdpa.rs:36-171 at Lq > 1: softmax(Q K^T / 8 + causal mask) V).
// grid (tiles, heads), block 128.  A tile is up to PF_QT = 64 consecutive new rows of ONE sequence (host-built list); warp w
// owns query rows [16 w, 16 w + 16) of it.  The row-per-CTA kernel above re-reads the sequence's K / V rows for every query
// row (a 64 x 40-token open: 40 960 CTAs x 32 KB = 1.3 GB through L2 per layer, 222 us) and spends ~2 000 SIMT instructions
// per row; here the CTA stages keys [0, last position] of its (sequence, head) in shared memory once (rows padded to 144
// bytes: conflict-free fragment loads) and each warp runs a flash-style pass over them with mma.sync m16n8k16:
//   S = Q K^T   A = Q rows as f16 fragments (held in registers for the whole pass), B = K rows (32-bit shared loads)
//   online softmax on the accumulator fragments (row g / g + 8 of a quad: two shuffles per reduction), exp2 with
//   log2(e) / 8 folded into the scale, causal mask by absolute position, key blocks above the warp's last row skipped
//   O += P V    A = P rounded to f16 straight from the S fragments, B = V through ldmatrix.trans
// f32 accumulation throughout; q and P are rounded to f16 (the K, V rows and the output already are).
static constexpr int PF_QT = 64;
static constexpr int PF_KP = 72;   // shared-memory pitch of a K / V row in halves (144 bytes)
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__global__ void __launch_bounds__(128) flowlm_attn_prefill_mma_kernel(
    const float* __restrict__ q_rot, const int2* __restrict__ tiles, const int* __restrict__ row_seq,
    const int* __restrict__ row_pos, const SeqDesc* __restrict__ seqs, int layer, int n_heads, __half* __restrict__ out16) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t pf_smem[];
  const int2 tl = tiles[blockIdx.x];
  const int r0 = tl.x, nr = tl.y, h = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int d_model = n_heads * HD;
  const SeqDesc sd = seqs[row_seq[r0]];
  const int pos0 = row_pos[r0];                 // absolute position of the tile's first row; rows are consecutive
  const int kmax = pos0 + nr;                   // keys [0, kmax) cover every row of the tile
  const int kpad = (kmax + 15) & ~15;
  __half* ks = reinterpret_cast<__half*>(pf_smem);
  __half* vs = ks + static_cast<size_t>(kpad) * PF_KP;
  for (int idx = tid; idx < kpad * 8; idx += 128) {
    const int i = idx >> 3, c = idx & 7;
    uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
    if (i < kmax) {
      kv = reinterpret_cast<const uint4*>(kv_row(sd, layer, 0, h, n_heads, i))[c];
      vv = reinterpret_cast<const uint4*>(kv_row(sd, layer, 1, h, n_heads, i))[c];
    }
    *reinterpret_cast<uint4*>(ks + i * PF_KP + c * 8) = kv;
    *reinterpret_cast<uint4*>(vs + i * PF_KP + c * 8) = vv;
  }
  __syncthreads();
  if (warp * 16 >= nr) return;
  // Q fragments: rows g and g + 8 of the warp's 16, four k-steps of 16 dims
  uint32_t qa[4][4];
  {
    const int ra = warp * 16 + g, rb = ra + 8;
    const float* qa_p = q_rot + static_cast<long long>(r0 + (ra < nr ? ra : 0)) * d_model + h * HD;
    const float* qb_p = q_rot + static_cast<long long>(r0 + (rb < nr ? rb : 0)) * d_model + h * HD;
#pragma unroll
    for (int s4 = 0; s4 < 4; ++s4) {
      const float2 a_lo = *reinterpret_cast<const float2*>(qa_p + 16 * s4 + 2 * t), a_hi = *reinterpret_cast<const float2*>(qa_p + 16 * s4 + 8 + 2 * t);
      const float2 b_lo = *reinterpret_cast<const float2*>(qb_p + 16 * s4 + 2 * t), b_hi = *reinterpret_cast<const float2*>(qb_p + 16 * s4 + 8 + 2 * t);
      qa[s4][0] = pack_h2(a_lo.x, a_lo.y);
      qa[s4][1] = pack_h2(b_lo.x, b_lo.y);
      qa[s4][2] = pack_h2(a_hi.x, a_hi.y);
      qa[s4][3] = pack_h2(b_hi.x, b_hi.y);
    }
  }
  const float kScale = 0.125f * 1.4426950408889634f;   // 1/sqrt(64) and log2(e): P = exp2(S * kScale - m)
  const int qp_a = pos0 + warp * 16 + g, qp_b = qp_a + 8;   // absolute positions of the thread's two rows
  const int kb_end = (pos0 + min(warp * 16 + 15, nr - 1)) / 16 + 1;   // key blocks up to the warp's last real row
  float m_a = -INFINITY, m_b = -INFINITY, l_a = 0.f, l_b = 0.f;
  float o[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
  const uint32_t vs_u32 = smem_u32(vs);
  for (int kb = 0; kb < kb_end; ++kb) {
    float sacc[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      sacc[j][0] = sacc[j][1] = sacc[j][2] = sacc[j][3] = 0.f;
      const __half* krow = ks + (kb * 16 + 8 * j + g) * PF_KP + 2 * t;
#pragma unroll
      for (int s4 = 0; s4 < 4; ++s4) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(krow + 16 * s4);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(krow + 16 * s4 + 8);
        mma_16816(sacc[j], qa[s4][0], qa[s4][1], qa[s4][2], qa[s4][3], b0, b1);
      }
    }
    // scale, causal mask, block row maxima
    float mx_a = -INFINITY, mx_b = -INFINITY;
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kp = kb * 16 + 8 * j + 2 * t + (e & 1);
        const int qp = (e & 2) ? qp_b : qp_a;
        const float sv = (kp <= qp) ? sacc[j][e] * kScale : -INFINITY;
        sacc[j][e] = sv;
        if (e & 2) mx_b = fmaxf(mx_b, sv); else mx_a = fmaxf(mx_a, sv);
      }
    mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 1));
    mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 2));
    mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 1));
    mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 2));
    const float mn_a = fmaxf(m_a, mx_a), mn_b = fmaxf(m_b, mx_b);
    // key 0 is visible to every row, so mn is finite from the first block on
    const float ca = exp2f(m_a - mn_a), cb = exp2f(m_b - mn_b);   // exp2(-inf) = 0 on the first block
    m_a = mn_a; m_b = mn_b;
    float p[2][4];
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      p[j][0] = exp2f(sacc[j][0] - mn_a); p[j][1] = exp2f(sacc[j][1] - mn_a);
      p[j][2] = exp2f(sacc[j][2] - mn_b); p[j][3] = exp2f(sacc[j][3] - mn_b);
      sa += p[j][0] + p[j][1];
      sb += p[j][2] + p[j][3];
    }
    l_a = l_a * ca + sa;
    l_b = l_b * cb + sb;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { o[nt][0] *= ca; o[nt][1] *= ca; o[nt][2] *= cb; o[nt][3] *= cb; }
    const uint32_t pa0 = pack_h2(p[0][0], p[0][1]), pa1 = pack_h2(p[0][2], p[0][3]);
    const uint32_t pa2 = pack_h2(p[1][0], p[1][1]), pa3 = pack_h2(p[1][2], p[1][3]);
    // V fragments: ldmatrix.x4.trans over keys kb*16 .. +15, two 8-dim column blocks per instruction
    // (lanes 0-7: keys 0-7 of block nt, 8-15: keys 8-15 of nt, 16-23: keys 0-7 of nt+1, 24-31: keys 8-15 of nt+1)
    const uint32_t vrow = vs_u32 + static_cast<uint32_t>(((kb * 16 + (lane & 15)) * PF_KP + (lane >> 4) * 8) * 2);
#pragma unroll
    for (int nt = 0; nt < 8; nt += 2) {
      uint32_t b0, b1, b2, b3;
      ldmatrix_x4_trans(vrow + nt * 16, b0, b1, b2, b3);
      mma_16816(o[nt], pa0, pa1, pa2, pa3, b0, b1);
      mma_16816(o[nt + 1], pa0, pa1, pa2, pa3, b2, b3);
    }
  }
  l_a += __shfl_xor_sync(0xffffffffu, l_a, 1);
  l_a += __shfl_xor_sync(0xffffffffu, l_a, 2);
  l_b += __shfl_xor_sync(0xffffffffu, l_b, 1);
  l_b += __shfl_xor_sync(0xffffffffu, l_b, 2);
  const float ia = 1.f / l_a, ib = 1.f / l_b;
  const int ra = warp * 16 + g, rb = ra + 8;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    if (ra < nr)
      *reinterpret_cast<uint32_t*>(out16 + static_cast<long long>(r0 + ra) * d_model + h * HD + nt * 8 + 2 * t) = pack_h2(o[nt][0] * ia, o[nt][1] * ia);
    if (rb < nr)
      *reinterpret_cast<uint32_t*>(out16 + static_cast<long long>(r0 + rb) * d_model + h * HD + nt * 8 + 2 * t) = pack_h2(o[nt][2] * ib, o[nt][3] * ib);
  }
}

// ---------------------------------------------------------------- Mimi encoder (voice cloning from PCM)
// First SEANetEncoder layer: Conv1d 1 -> 64, k7, causal with zero left context (reference seanet.rs:160-170 over
// conv.rs:90-136).  One input channel is no GEMM: one thread per (sample, channel).  Writes the f32 activation (the
// ResBlock's skip) and ELU(x) as the f16 operand of the block's k3 conv behind its two left-context rows.
__global__ void enc_conv0_kernel(const float* __restrict__ pcm, int T, const float* __restrict__ w /*[7][64]*/,
                                 const float* __restrict__ bias, float* __restrict__ x32 /*[T][64]*/,
                                 __half* __restrict__ e16 /*[2+T][64]*/) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const int t = static_cast<int>(idx >> 6), c = static_cast<int>(idx & 63);
  if (t >= T) return;
  float acc = bias[c];
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    const int tt = t + j - 6;
    if (tt >= 0) acc += __ldg(w + j * 64 + c) * __ldg(pcm + tt);
  }
  x32[static_cast<long long>(t) * 64 + c] = acc;
  e16[static_cast<long long>(2 + t) * 64 + c] = __float2half_rn(elu1(acc));
}

// Encoder transformer attention over a whole prompt (reference attention.rs:104-283 with the causal + window mask of
// sdpa.rs:129-171: row i sees keys (i - context, i]).  Runs once per voice, so it is plain SIMT.
// Step 1: RoPE on q and k in place at the absolute position.  qkv f32 [P][3*512] (q | k | v thirds, heads inside).
__global__ void enc_rope_kernel(float* __restrict__ qkv, int pos0) {
  const int r = blockIdx.x, h = blockIdx.y, i = threadIdx.x;  // 32 threads = 32 pairs
  float* q = qkv + static_cast<long long>(r) * (3 * 512) + h * HD;
  float* k = q + 512;
  float a, b;
  rope_pair(q[2 * i], q[2 * i + 1], pos0 + r, i, a, b);
  q[2 * i] = a; q[2 * i + 1] = b;
  rope_pair(k[2 * i], k[2 * i + 1], pos0 + r, i, a, b);
  k[2 * i] = a; k[2 * i + 1] = b;
}
// Step 2: one warp per (row, head), lane = two head dims, single pass with a running maximum over the window.
__global__ void enc_attn_kernel(const float* __restrict__ qkv, int P, int context, __half* __restrict__ out16 /*[P][512]*/) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int r = w >> 3, h = w & 7;
  if (r >= P) return;
  const float* base = qkv + h * HD + 2 * lane;
  const float2 q = *reinterpret_cast<const float2*>(base + static_cast<long long>(r) * 1536);
  float m = -INFINITY, l = 0.f, a0 = 0.f, a1 = 0.f;
  for (int j = max(0, r - context + 1); j <= r; ++j) {
    const float2 kk = *reinterpret_cast<const float2*>(base + static_cast<long long>(j) * 1536 + 512);
    const float2 vv = *reinterpret_cast<const float2*>(base + static_cast<long long>(j) * 1536 + 1024);
    const float sc = warp_sum(q.x * kk.x + q.y * kk.y) * 0.125f;
    const float m_new = fmaxf(m, sc);
    const float corr = expf(m - m_new), p = expf(sc - m_new);
    l = l * corr + p;
    a0 = a0 * corr + p * vv.x;
    a1 = a1 * corr + p * vv.y;
    m = m_new;
  }
  *reinterpret_cast<__half2*>(out16 + static_cast<long long>(r) * 512 + h * HD + 2 * lane) = __floats2half2_rn(a0 / l, a1 / l);
}
// ConvDownsample1d operand (reference conv.rs:278-312, `replicate` padding on a first call, conv.rs:114-123): the f16
// rows of the transformer output behind 16 copies of its first row.
__global__ void enc_downsample_prep_kernel(const float* __restrict__ x /*[P][512]*/, __half* __restrict__ d16 /*[16+P][512]*/) {
  const int row = blockIdx.x, src = row < 16 ? 0 : row - 16;
  const float4 v = reinterpret_cast<const float4*>(x + static_cast<long long>(src) * 512)[threadIdx.x];  // 128 threads
  const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
  uint2 pk;
  pk.x = *reinterpret_cast<const uint32_t*>(&h0);
  pk.y = *reinterpret_cast<const uint32_t*>(&h1);
  reinterpret_cast<uint2*>(d16 + static_cast<long long>(row) * 512)[threadIdx.x] = pk;
}

// The same operand for the first frame of every later chunk of a long prompt: the reference encodes prompts of more
// than 120 frames chunk by chunk and calls the downsample with step = 0 each time (tts_model.rs:540), so its replicate
// padding restarts there.  Block (b, row): 32 operand rows per boundary b at frame (b + 1) * chunk_frames.
__global__ void enc_downsample_boundary_prep_kernel(const float* __restrict__ x /*[P][512]*/, int chunk_frames,
                                                    __half* __restrict__ d16 /*[n_b][32][512]*/) {
  const int b = blockIdx.x, row = blockIdx.y;
  const long long first = static_cast<long long>(b + 1) * chunk_frames * 16;
  const long long src = row < 16 ? first : first + row - 16;
  const float4 v = reinterpret_cast<const float4*>(x + src * 512)[threadIdx.x];
  const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
  uint2 pk;
  pk.x = *reinterpret_cast<const uint32_t*>(&h0);
  pk.y = *reinterpret_cast<const uint32_t*>(&h1);
  reinterpret_cast<uint2*>(d16 + (static_cast<long long>(b) * 32 + row) * 512)[threadIdx.x] = pk;
}

// ---------------------------------------------------------------- Mimi front end
// latent de-norm (tts_model.rs:1033-1035) -> Quantizer 1x1 conv 32->512 (mimi.rs:32-36) -> depthwise
// ConvTranspose1d k32 s16 with carried partial (conv.rs:314-346 -> :219-267).  grid (n, 4, f), block 128: one channel per
// thread, weights pre-transposed to [k][512] so every access of a warp is contiguous; the 16 partial-sum reads are
// issued before any store to the same buffer.
// f > 1 (codec group): the latents of f consecutive frames of every row wait in a queue (z + j * z_frame_stride, frame
// index of frame 0 in zpos) and are decoded by one launch.  The partial a frame inherits is a function of the previous
// frame's latent alone, so block (b, ., j > 0) recomputes it from queue entry j - 1 instead of waiting for block j - 1;
// only the last frame writes the slot's carried partial.  Same arithmetic as f launches of one frame (bit-identical).
__global__ void mimi_frontend_kernel(const float* __restrict__ z, long long z_frame_stride, const int* __restrict__ zpos,
                                     const int* __restrict__ row_seq, const StreamCtl* __restrict__ ctl,
                                     const float* __restrict__ emb_std, const float* __restrict__ emb_mean,
                                     const float* __restrict__ wq_t /*[32,512]*/, const float* __restrict__ wup_t /*[32,512]*/,
                                     float* __restrict__ partial /*[slots,16,512]*/, float* __restrict__ x /*[n*f*16,512]*/,
                                     float* __restrict__ dbg_quant, int* __restrict__ mimi_pos) {
  pdl_launch_dependents();
  __shared__ float zd[2][LDIM];
  const int b = blockIdx.x, c = blockIdx.y * 128 + threadIdx.x, j = blockIdx.z, f = gridDim.z;
  // this thread's column of the quantizer projection and of the upsampling kernel, and the de-normalisation constants:
  // requested before the dependency on the step that produced the latent resolves
  float wq[LDIM], w0[16], w1[16];
#pragma unroll
  for (int k = 0; k < LDIM; ++k) wq[k] = __ldg(wq_t + k * 512 + c);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    w0[i] = __ldg(wup_t + i * 512 + c);
    w1[i] = __ldg(wup_t + (16 + i) * 512 + c);
  }
  float es = 0.f, em = 0.f;
  if (threadIdx.x < LDIM) { es = __ldg(emb_std + threadIdx.x); em = __ldg(emb_mean + threadIdx.x); }
  pdl_wait();
  const int slot = row_seq[b];
  // zd[1]: the frame whose second half this block also needs -- the previous frame (j > 0: it supplies the partial this
  // frame inherits) or, in block j = 0, the group's last frame (it supplies the partial the slot carries to the next pass;
  // the block that reads the carried partial is the one that replaces it, so no other block ever touches it)
  const int j2 = j > 0 ? j - 1 : f - 1;
  if (threadIdx.x < LDIM) {
    zd[0][threadIdx.x] = z[j * z_frame_stride + b * LDIM + threadIdx.x] * es + em;
    zd[1][threadIdx.x] = z[j2 * z_frame_stride + b * LDIM + threadIdx.x] * es + em;
  }
  // position of the group's first frame.  Without a queue this launch runs after step_end of the same frame: the frame
  // counter has already advanced by one
  if (c == 0 && j == 0) mimi_pos[b] = zpos ? zpos[b] * 16 : (ctl[slot].frame - 1) * 16;
  __syncthreads();
  float qv = 0.f, q2 = 0.f;
#pragma unroll
  for (int k = 0; k < LDIM; ++k) qv += wq[k] * zd[0][k];
#pragma unroll
  for (int k = 0; k < LDIM; ++k) q2 += wq[k] * zd[1][k];
  if (dbg_quant && j == 0) dbg_quant[b * 512 + c] = q2;   // the last frame of the group
  float* part = partial + static_cast<long long>(slot) * 16 * 512;
  float old[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) old[i] = (j == 0) ? part[i * 512 + c] : __fmul_rn(q2, w1[i]);
  float* xo = x + (static_cast<long long>(b) * f + j) * 16 * 512 + c;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    xo[i * 512] = __fmaf_rn(qv, w0[i], old[i]);
    if (j == 0) part[i * 512 + c] = __fmul_rn(q2, w1[i]);
  }
}

// Mimi decoder-transformer attention: 16 new rows per stream, sliding window of 250 positions
// (reference attention.rs:167-264 ring + sdpa.rs:129-171 mask: query at position p sees keys in (p-250, p]).
// Per slot, layer and head the ring holds K as [272][64] and V transposed as [64][272], both f16, indexed by
// position % 272.  272 = 17 blocks of 16 and every frame's 16 rows fill exactly one block, so the block written first
// is the one whose keys fell out of every window of this frame.
// grid (n, 8 heads), block 192: the 16 queries are exactly one m16n8k16 tile.  Warp w takes ring blocks w, w+6, w+12
// (16 keys each) and runs S = Q K^T and O = P V on mma.sync with every operand fragment loaded straight from HBM in
// 16/32-byte pieces: both products sum over an index whose order is free, so the K rows are read with the head dim
// permuted (lane c of a quad owns dims 16c..16c+15 for all four k-steps) and the S columns are assigned to keys so
// that a lane ends up with four consecutive keys (4c..4c+3), which is one 8-byte load per dim of the transposed V.
// The six warps' partial (max, sum, O) are merged in warp order (bit-reproducible).
static constexpr int MATTN_THREADS = 192;
static constexpr int MATTN_WARPS = MATTN_THREADS / 32;
static constexpr int MATTN_BLOCKS = MIMI_RING / 16;  // 17
static constexpr int MATTN_QP = HD + 8;              // q16 pitch (halves)
static constexpr int MATTN_VP = 24;                  // new-V transpose pitch (halves), 16-byte aligned rows
static constexpr int MATTN_OP = HD + 4;              // partial-O pitch (floats)
static constexpr int MATTN_SMEM = 16 * MATTN_QP * 2 + HD * MATTN_VP * 2 + MATTN_WARPS * 16 * (MATTN_OP + 2) * 4;

__device__ __forceinline__ void mma_m16n8k16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                             uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}

__global__ void __launch_bounds__(MATTN_THREADS, 3)
mimi_attn_kernel(const float* __restrict__ qkv_all /*[n*f*16, 1536]*/, const int* __restrict__ row_seq,
                 const int* __restrict__ mimi_pos, __half* ring /*[slots][layer][k|v][8][272*64]*/, int layer, int n_layers,
                 __half* __restrict__ out16_all /*[n*f*16, 512]*/, int f) {
  pdl_launch_dependents();
  pdl_wait();
  // f > 1 (codec group): the f consecutive frames of the row, one after the other -- frame j + 1 sees the ring rows frame j
  // wrote, exactly as f launches would
  for (int fj = 0; fj < f; ++fj) {
  if (fj) __syncthreads();  // the merge of the previous frame has read o_s / m_s / l_s
  const float* qkv = qkv_all + (static_cast<long long>(blockIdx.x) * (f - 1) + fj) * 16 * 1536;
  __half* out16 = out16_all + (static_cast<long long>(blockIdx.x) * (f - 1) + fj) * 16 * 512;
  constexpr int NH = 8, DM = 512, T = 16, NW = MATTN_WARPS;
  constexpr float kScale = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) and log2(e): softmax in base 2
  extern __shared__ __align__(16) unsigned char msm_raw[];
  __half* q16 = reinterpret_cast<__half*>(msm_raw);                      // [16][72]
  __half* vt_s = q16 + 16 * MATTN_QP;                                    // [64][24]
  float* o_s = reinterpret_cast<float*>(vt_s + HD * MATTN_VP);           // [6][16][68]
  float* m_s = o_s + NW * 16 * MATTN_OP;                                 // [6][16]
  float* l_s = m_s + NW * 16;                                            // [6][16]
  const int b = blockIdx.x, h = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, c = lane & 3;
  const int slot = row_seq[b];
  const int p0 = mimi_pos[b] + 16 * fj;  // absolute position of the first new row (multiple of 16)
  __half* kring = ring + ((static_cast<long long>(slot) * n_layers + layer) * 2 * NH + h) * MIMI_RING * HD;
  __half* vring = kring + static_cast<long long>(NH) * MIMI_RING * HD;  // [64][272]
  const int blk0 = (p0 >> 4) % MATTN_BLOCKS;                              // ring block of the new rows
  // RoPE, K ring rows, q (f16) and the new V rows transposed through smem: 16 rows x 32 pairs
  for (int it = tid; it < T * 32; it += MATTN_THREADS) {
    const int t = it >> 5, i = it & 31;
    const float* row = qkv + (static_cast<long long>(b) * T + t) * 3 * DM;
    const float2 qx = *reinterpret_cast<const float2*>(row + h * HD + 2 * i);
    const float2 kx = *reinterpret_cast<const float2*>(row + DM + h * HD + 2 * i);
    const float2 vx = *reinterpret_cast<const float2*>(row + 2 * DM + h * HD + 2 * i);
    float qr, qi, kr, ki;
    rope_pair(qx.x, qx.y, p0 + t, i, qr, qi);
    rope_pair(kx.x, kx.y, p0 + t, i, kr, ki);
    reinterpret_cast<__half2*>(q16 + t * MATTN_QP)[i] = __floats2half2_rn(qr, qi);
    reinterpret_cast<__half2*>(kring + (blk0 * 16 + t) * HD)[i] = __floats2half2_rn(kr, ki);
    vt_s[(2 * i) * MATTN_VP + t] = __float2half_rn(vx.x);
    vt_s[(2 * i + 1) * MATTN_VP + t] = __float2half_rn(vx.y);
  }
  __syncthreads();
  if (tid < 2 * HD) {
    const int d = tid >> 1, hf = tid & 1;
    *reinterpret_cast<uint4*>(vring + d * MIMI_RING + blk0 * 16 + 8 * hf) = *reinterpret_cast<const uint4*>(vt_s + d * MATTN_VP + 8 * hf);
  }
  // q fragments: rows g and g+8, dims 16c..16c+15 (the permuted k index, see above)
  uint32_t qa[8], qb[8];
  {
    const uint4 a0 = *reinterpret_cast<const uint4*>(q16 + g * MATTN_QP + 16 * c);
    const uint4 a1 = *reinterpret_cast<const uint4*>(q16 + g * MATTN_QP + 16 * c + 8);
    const uint4 b0 = *reinterpret_cast<const uint4*>(q16 + (g + 8) * MATTN_QP + 16 * c);
    const uint4 b1 = *reinterpret_cast<const uint4*>(q16 + (g + 8) * MATTN_QP + 16 * c + 8);
    qa[0] = a0.x; qa[1] = a0.y; qa[2] = a0.z; qa[3] = a0.w; qa[4] = a1.x; qa[5] = a1.y; qa[6] = a1.z; qa[7] = a1.w;
    qb[0] = b0.x; qb[1] = b0.y; qb[2] = b0.z; qb[3] = b0.w; qb[4] = b1.x; qb[5] = b1.y; qb[6] = b1.z; qb[7] = b1.w;
  }
  __syncthreads();  // the new K / V rows are visible to the whole CTA
  // this warp's ring blocks and the position of their first key; a block that was never written has posb < 0
  int blk[3], posb[3];
#pragma unroll
  for (int u = 0; u < 3; ++u) {
    blk[u] = warp + u * NW;
    posb[u] = (blk[u] < MATTN_BLOCKS) ? p0 - 16 * ((blk0 - blk[u] + MATTN_BLOCKS) % MATTN_BLOCKS) : -1;
  }
  // K fragments of all three blocks in flight at once: n-tile t of a block, column g  <->  key 4*(g/2) + 2t + (g%2)
  uint4 kf[3][2][2];
#pragma unroll
  for (int u = 0; u < 3; ++u)
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      if (posb[u] >= 0) {
        const uint4* kr = reinterpret_cast<const uint4*>(kring + (blk[u] * 16 + 4 * (g >> 1) + 2 * t + (g & 1)) * HD + 16 * c);
        kf[u][t][0] = kr[0];
        kf[u][t][1] = kr[1];
      } else {
        kf[u][t][0] = kf[u][t][1] = make_uint4(0, 0, 0, 0);
      }
    }
  float sacc[3][2][4];
#pragma unroll
  for (int u = 0; u < 3; ++u)
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      sacc[u][t][0] = sacc[u][t][1] = sacc[u][t][2] = sacc[u][t][3] = 0.f;
      const uint32_t kk[8] = {kf[u][t][0].x, kf[u][t][0].y, kf[u][t][0].z, kf[u][t][0].w,
                              kf[u][t][1].x, kf[u][t][1].y, kf[u][t][1].z, kf[u][t][1].w};
#pragma unroll
      for (int s4 = 0; s4 < 4; ++s4) mma_m16n8k16(sacc[u][t], qa[2 * s4], qb[2 * s4], qa[2 * s4 + 1], qb[2 * s4 + 1], kk[2 * s4], kk[2 * s4 + 1]);
    }
  // V fragments (issued before the softmax arithmetic): dim nt*8+g, keys 4c..4c+3 of the block
  uint2 vf[3][8];
#pragma unroll
  for (int u = 0; u < 3; ++u)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
      vf[u][nt] = (posb[u] >= 0) ? *reinterpret_cast<const uint2*>(vring + (nt * 8 + g) * MIMI_RING + blk[u] * 16 + 4 * c) : make_uint2(0, 0);
  // mask, row maxima and sums over this warp's 48 keys; rows g (index 0) and g+8 (index 1)
  float mrow[2] = {-INFINITY, -INFINITY};
#pragma unroll
  for (int u = 0; u < 3; ++u)
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kp = posb[u] + 4 * c + 2 * t + (e & 1);
        const int qp = p0 + g + 8 * (e >> 1);
        const bool ok = (posb[u] >= 0) && (kp <= qp) && (kp > qp - MIMI_CTX);
        const float sv = ok ? sacc[u][t][e] * kScale : -INFINITY;
        sacc[u][t][e] = sv;
        mrow[e >> 1] = fmaxf(mrow[e >> 1], sv);
      }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    mrow[r] = fmaxf(mrow[r], __shfl_xor_sync(0xffffffffu, mrow[r], 1));
    mrow[r] = fmaxf(mrow[r], __shfl_xor_sync(0xffffffffu, mrow[r], 2));
  }
  const float muse[2] = {mrow[0] == -INFINITY ? 0.f : mrow[0], mrow[1] == -INFINITY ? 0.f : mrow[1]};
  float lrow[2] = {0.f, 0.f};
  float oacc[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) oacc[nt][0] = oacc[nt][1] = oacc[nt][2] = oacc[nt][3] = 0.f;
#pragma unroll
  for (int u = 0; u < 3; ++u) {
    float pv[2][4];
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        pv[t][e] = exp2f(sacc[u][t][e] - muse[e >> 1]);
        lrow[e >> 1] += pv[t][e];
      }
    const uint32_t a0 = pack_half2(pv[0][0], pv[0][1]), a1 = pack_half2(pv[0][2], pv[0][3]);
    const uint32_t a2 = pack_half2(pv[1][0], pv[1][1]), a3 = pack_half2(pv[1][2], pv[1][3]);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) mma_m16n8k16(oacc[nt], a0, a1, a2, a3, vf[u][nt].x, vf[u][nt].y);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 1);
    lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 2);
  }
  if (c == 0) {
    m_s[warp * 16 + g] = mrow[0];
    m_s[warp * 16 + g + 8] = mrow[1];
    l_s[warp * 16 + g] = lrow[0];
    l_s[warp * 16 + g + 8] = lrow[1];
  }
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    *reinterpret_cast<float2*>(o_s + (warp * 16 + g) * MATTN_OP + nt * 8 + 2 * c) = make_float2(oacc[nt][0], oacc[nt][1]);
    *reinterpret_cast<float2*>(o_s + (warp * 16 + g + 8) * MATTN_OP + nt * 8 + 2 * c) = make_float2(oacc[nt][2], oacc[nt][3]);
  }
  __syncthreads();
  // merge the six partials in warp order; every query sees at least its own key, so the sum is positive
  for (int o = tid; o < T * 32; o += MATTN_THREADS) {
    const int t = o >> 5, i = o & 31;
    float mx = -INFINITY;
#pragma unroll
    for (int w = 0; w < NW; ++w) mx = fmaxf(mx, m_s[w * 16 + t]);
    float sx = 0.f, sy = 0.f, den = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const float mw = m_s[w * 16 + t];
      const float f = (mw == -INFINITY) ? 0.f : exp2f(mw - mx);
      const float2 v = *reinterpret_cast<const float2*>(o_s + (w * 16 + t) * MATTN_OP + 2 * i);
      sx += f * v.x;
      sy += f * v.y;
      den += f * l_s[w * 16 + t];
    }
    const float inv = 1.f / den;
    reinterpret_cast<__half2*>(out16 + (static_cast<long long>(b) * T + t) * DM + h * HD)[i] = __floats2half2_rn(sx * inv, sy * inv);
  }
  }  // frames of the group
}

// ---------------------------------------------------------------- flow head glue
// y16[r, :] = silu(c[r, :] + te[:])   (reference modules/mlp.rs:328-330)
__global__ void silu_add_kernel(const float* __restrict__ c /*[n,C]*/, const float* __restrict__ te /*[steps,C]*/, int n, int steps, int C,
                                __half* __restrict__ y16 /*[steps*n,C]: row s*n + r = silu(c[r] + te[s])*/) {
  pdl_launch_dependents();
  pdl_wait();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long long>(steps) * n * C) return;
  const int col = static_cast<int>(i % C);
  const long long row = i / C;
  const int s = static_cast<int>(row / n), r = static_cast<int>(row - static_cast<long long>(s) * n);
  y16[i] = __float2half_rn(silu(c[static_cast<long long>(r) * C + col] + te[s * C + col]));
}

// ---------------------------------------------------------------- step begin / end
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
// Counter-based N(0,1): hash(seed, frame, lane) -> two uniforms -> Box-Muller.  Only used when the caller
// injects no noise (throughput runs); parity runs inject the oracle's noise.
__device__ __forceinline__ float counter_normal(unsigned long long seed, int frame, int k) {
  uint32_t a = mix32(static_cast<uint32_t>(seed) ^ mix32(static_cast<uint32_t>(frame) * 0x9E3779B9U + k));
  uint32_t b = mix32(static_cast<uint32_t>(seed >> 32) ^ mix32(a + 0x85EBCA6BU));
  const float u1 = (static_cast<float>(a >> 8) + 1.f) * (1.f / 16777217.f);
  const float u2 = static_cast<float>(b >> 8) * (1.f / 16777216.f);
  return sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
}

// Gathers the AR feedback latent (BOS at frame 0, reference tts_model.rs:971,1065) as the f16 operand of
// input_linear (K padded 32 -> 64) and the flow head's starting point x_0 (flow_lm.rs:148-153).
// Also flattens the row -> slot -> (KV descriptor, cursor) chain into one 32-byte record per batch row (the cursor rides in
// SeqDesc::pad): the six attention launches of the step then start from one load instead of two dependent round trips.
__global__ void step_begin_kernel(const int* __restrict__ row_seq, const StreamCtl* __restrict__ ctl,
                                  const float* __restrict__ feedback /*[slots,32]*/, __half* __restrict__ lat16 /*[n,64]*/,
                                  float* __restrict__ z32 /*[n,32]*/, __half* __restrict__ z16 /*[n,64]*/,
                                  const SeqDesc* __restrict__ seqs, const int* __restrict__ own_len,
                                  SeqDesc* __restrict__ row_desc /*[n]*/) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, k = threadIdx.x;  // 64 threads
  const int slot = row_seq[b];
  if (k == 63) {
    SeqDesc d = seqs[slot];
    d.pad = own_len[slot];
    row_desc[b] = d;
  }
  const StreamCtl c = ctl[slot];
  float lat = 0.f, z = 0.f;
  if (k < LDIM) {
    lat = feedback[slot * LDIM + k];
    // a step enqueued ahead of the host (PTTS_STEP_AHEAD) may run one frame past the end: no noise row exists there
    if (c.noise) z = c.frame < c.max_gen_len ? c.noise[static_cast<long long>(c.frame) * LDIM + k] : 0.f;
    else if (c.temp > 0.f) z = sqrtf(c.temp) * counter_normal(c.seed, c.frame, k);
    z32[b * LDIM + k] = z;
  }
  lat16[b * 64 + k] = __float2half_rn(lat);
  z16[b * 64 + k] = __float2half_rn(z);
}

// step_begin + input_linear + the first LayerNorm in ONE launch (reference tts_model.rs:971,1065 feedback / BOS;
// flow_lm.rs:118 input_linear, :148-153 noise draw; modules/mlp.rs:29-58 LayerNorm).  As three launches (gather, a tcgen05
// GEMM with K = 32 padded to 64, LayerNorm) this was 13.6 us of mostly launch latency for 2 MFLOP.  grid n, block 256:
// thread t owns features 4t..4t+3 of the row; its 4 x 32 weights and the affine parameters are constants and are requested
// before the dependency on the previous step resolves.  Same arithmetic as the GEMM up to the order of the 32 additions
// (f16 weights, the latent rounded to f16, f32 accumulation).
__global__ void __launch_bounds__(256) flowlm_input_kernel(const int* __restrict__ row_seq, const StreamCtl* __restrict__ ctl,
                                                           const float* __restrict__ feedback, const SeqDesc* __restrict__ seqs,
                                                           const int* __restrict__ own_len, SeqDesc* __restrict__ row_desc,
                                                           float* __restrict__ z32, __half* __restrict__ z16,
                                                           const __half* __restrict__ w_in /*[1024][64], k >= 32 zero*/,
                                                           const float* __restrict__ w_scale /*[1024] int8 mode, else null*/,
                                                           const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                                           float* __restrict__ x32 /*[n,1024]*/, __half* __restrict__ h16 /*[n,1024]*/) {
  pdl_launch_dependents();
  constexpr int C = 1024;
  __shared__ float lat_s[LDIM];
  __shared__ float red_s[2][8];
  const int b = blockIdx.x, t = threadIdx.x, f = t * 4;
  uint4 wr[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) wr[j][c4] = __ldg(reinterpret_cast<const uint4*>(w_in + static_cast<size_t>(f + j) * 64) + c4);
  const float4 lw = __ldg(reinterpret_cast<const float4*>(ln_w + f)), lb = __ldg(reinterpret_cast<const float4*>(ln_b + f));
  float4 ws = make_float4(1.f, 1.f, 1.f, 1.f);
  if (w_scale) ws = __ldg(reinterpret_cast<const float4*>(w_scale + f));
  pdl_wait();
  const int slot = row_seq[b];
  if (t == 64) {
    SeqDesc d = seqs[slot];
    d.pad = own_len[slot];
    row_desc[b] = d;
  }
  if (t < 64) {
    float z = 0.f;
    if (t < LDIM) {
      const StreamCtl c = ctl[slot];
      // a step enqueued ahead of the host (PTTS_STEP_AHEAD) may run one frame past the end: no noise row exists there
      if (c.noise) z = c.frame < c.max_gen_len ? c.noise[static_cast<long long>(c.frame) * LDIM + t] : 0.f;
      else if (c.temp > 0.f) z = sqrtf(c.temp) * counter_normal(c.seed, c.frame, t);
      z32[b * LDIM + t] = z;
      lat_s[t] = __half2float(__float2half_rn(feedback[slot * LDIM + t]));
    }
    z16[b * 64 + t] = __float2half_rn(z);
  }
  __syncthreads();
  float xv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float a = 0.f;
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
      const __half2* h2 = reinterpret_cast<const __half2*>(&wr[j][c4]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 wf = __half22float2(h2[e]);
        a += wf.x * lat_s[c4 * 8 + 2 * e] + wf.y * lat_s[c4 * 8 + 2 * e + 1];
      }
    }
    xv[j] = a;
  }
  xv[0] *= ws.x; xv[1] *= ws.y; xv[2] *= ws.z; xv[3] *= ws.w;
  *reinterpret_cast<float4*>(x32 + static_cast<size_t>(b) * C + f) = make_float4(xv[0], xv[1], xv[2], xv[3]);
  const int warp = t >> 5, lane = t & 31;
  float s = warp_sum(xv[0] + xv[1] + xv[2] + xv[3]);
  if (lane == 0) red_s[0][warp] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red_s[0][i];
  const float mean = tot * (1.f / C);
  const float d0 = xv[0] - mean, d1 = xv[1] - mean, d2 = xv[2] - mean, d3 = xv[3] - mean;
  float q = warp_sum(d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3);
  if (lane == 0) red_s[1][warp] = q;
  __syncthreads();
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) var += red_s[1][i];
  const float rstd = 1.f / sqrtf(var * (1.f / C) + 1e-5f);
  const __half2 h0 = __floats2half2_rn(d0 * rstd * lw.x + lb.x, d1 * rstd * lw.y + lb.y);
  const __half2 h1 = __floats2half2_rn(d2 * rstd * lw.z + lb.z, d3 * rstd * lw.w + lb.w);
  uint2 pk;
  pk.x = *reinterpret_cast<const uint32_t*>(&h0);
  pk.y = *reinterpret_cast<const uint32_t*>(&h1);
  *reinterpret_cast<uint2*>(h16 + static_cast<size_t>(b) * C + f) = pk;
}

// EOS bookkeeping of the frame loop (reference tts_model.rs:1055-1069, D2: the frame at
// eos_step + frames_after_eos is still emitted) + AR feedback + cursor advance.
__global__ void step_end_kernel(const int* __restrict__ row_seq, int n, StreamCtl* __restrict__ ctl,
                                int* __restrict__ own_len, const float* __restrict__ eos_logit,
                                const float* __restrict__ z32, float* __restrict__ feedback,
                                unsigned char* __restrict__ finished_out, float* __restrict__ latent_out,
                                float* __restrict__ logit_out, float* __restrict__ zq /*[n,32] or null*/,
                                int* __restrict__ zq_pos /*[n]*/) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, k = threadIdx.x;  // 32 threads
  if (b >= n) return;
  const int slot = row_seq[b];
  const float zk = z32[b * LDIM + k];
  feedback[slot * LDIM + k] = zk;
  latent_out[b * LDIM + k] = zk;
  if (zq) zq[b * LDIM + k] = zk;  // codec group: this frame's latent waits here for the Mimi front end
  if (k == 0) {
    StreamCtl c = ctl[slot];
    const int step = c.frame;
    if (zq_pos) zq_pos[b] = step;
    const float logit = eos_logit[b];
    if (c.finished) {  // overrun frame of a step enqueued ahead: the stream's counters stay at its last real frame
      finished_out[b] = 1;
      logit_out[b] = logit;
      return;
    }
    if (logit > c.eos_threshold && c.eos_step < 0) c.eos_step = step;
    int fin = 0;
    if (c.eos_step >= 0 && step >= c.eos_step + c.frames_after_eos) fin = 1;
    if (step + 1 >= c.max_gen_len) fin = 1;
    c.finished = fin;
    c.frame = step + 1;
    ctl[slot] = c;
    own_len[slot] += 1;
    finished_out[b] = static_cast<unsigned char>(fin);
    logit_out[b] = logit;
  }
}

// ---------------------------------------------------------------- SEANet tail and streaming state
// Last layer: ELU (already applied by the producer) -> Conv1d 64->1 k3 (reference seanet.rs:379-392).
// One thread per output sample; the 3x64 window of sample t is 384 contiguous bytes.
// Also packs the frame as the reference's wire format (audio.rs:129-146 pcm_i16_le_bytes: clamp to [-1, 1], * 32767,
// truncating cast), so a host that streams i16 PCM reads back half the bytes and converts nothing.
__global__ void seanet_final_conv_kernel(const __half* __restrict__ a /*[n][2+T][64]*/, const float* __restrict__ w /*[3][64]*/,
                                         const float* __restrict__ bias, int n, int T /*1920 x frames of the group*/,
                                         float* __restrict__ pcm /*[n,T]*/, short* __restrict__ pcm16 /*[n,T] or null*/) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float w_s[192];
  if (threadIdx.x < 192) w_s[threadIdx.x] = w[threadIdx.x];
  __syncthreads();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (t >= T) return;
  const uint4* src = reinterpret_cast<const uint4*>(a + (static_cast<long long>(b) * (T + 2) + t) * 64);
  float acc = bias[0];
#pragma unroll
  for (int c = 0; c < 24; ++c) {
    uint4 u = src[c];
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = __half22float2(h[j]);
      acc += f.x * w_s[c * 8 + 2 * j] + f.y * w_s[c * 8 + 2 * j + 1];
    }
  }
  pcm[static_cast<long long>(b) * T + t] = acc;
  if (pcm16) pcm16[static_cast<long long>(b) * T + t] = static_cast<short>(fminf(fmaxf(acc, -1.f), 1.f) * 32767.f);
}

// Left-context rows of every streaming conv (reference `previous`, conv.rs:125-128; and the previous input row
// that replaces ConvTranspose1d's `partial`, conv.rs:246-262) move between the per-slot state and the head of
// the compact per-batch-row activation buffers.
struct ConvSeg {
  __half* buf;      // [max_batch][pad + T][C]
  __half* state;    // [slots][pad][C]
  int pad, T, C, unused;
};
struct ConvSegs { ConvSeg s[8]; };

__global__ void conv_state_move_kernel(const ConvSegs segs, const int* __restrict__ row_seq, int save) {
  pdl_launch_dependents();
  pdl_wait();
  const ConvSeg g = segs.s[blockIdx.y];
  const int b = blockIdx.x;
  const int slot = row_seq[b];
  const int n16 = g.pad * g.C / 8;  // 16-byte packets
  uint4* st = reinterpret_cast<uint4*>(g.state + static_cast<long long>(slot) * g.pad * g.C);
  uint4* hd = reinterpret_cast<uint4*>(g.buf + static_cast<long long>(b) * (g.pad + g.T) * g.C);
  uint4* tl = reinterpret_cast<uint4*>(g.buf + (static_cast<long long>(b) * (g.pad + g.T) + g.T) * g.C);
  for (int i = threadIdx.x; i < n16; i += blockDim.x) {
    if (save) st[i] = tl[i];
    else hd[i] = st[i];
  }
}

// Stream open: zero the per-slot streaming state (reference init_state zeros: conv.rs:71-88,205-217).
__global__ void slot_reset_kernel(const ConvSegs segs, float* __restrict__ partial, const int* __restrict__ slot_list) {
  pdl_launch_dependents();
  pdl_wait();
  const int slot = slot_list[blockIdx.x];
  if (blockIdx.y < 8) {
    const ConvSeg g = segs.s[blockIdx.y];
    uint4* st = reinterpret_cast<uint4*>(g.state + static_cast<long long>(slot) * g.pad * g.C);
    for (int i = threadIdx.x; i < g.pad * g.C / 8; i += blockDim.x) st[i] = make_uint4(0, 0, 0, 0);
  } else {
    float4* p = reinterpret_cast<float4*>(partial + static_cast<long long>(slot) * 16 * 512);
    for (int i = threadIdx.x; i < 16 * 512 / 4; i += blockDim.x) p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// Stream open, packed: the host stages one record per new stream in pinned memory, ONE copy brings them over, and this
// kernel scatters them into the slot arena (KV descriptor, control block, cursor, BOS as the first AR feedback) and
// zeroes the slot's streaming state (reference init_state zeros: conv.rs:71-88,205-217).  grid (n, 10), block 128.
struct OpenRec {
  SeqDesc sd;
  StreamCtl ctl;
  int slot;
  int own_len;
  int pad[2];
};
__global__ void slot_open_kernel(const OpenRec* __restrict__ recs, const ConvSegs segs, float* __restrict__ partial,
                                 SeqDesc* __restrict__ seqs, StreamCtl* __restrict__ ctl, int* __restrict__ own_len,
                                 float* __restrict__ feedback, const float* __restrict__ bos) {
  pdl_launch_dependents();
  pdl_wait();
  const OpenRec r = recs[blockIdx.x];
  const int slot = r.slot;
  if (blockIdx.y < 8) {
    const ConvSeg g = segs.s[blockIdx.y];
    uint4* st = reinterpret_cast<uint4*>(g.state + static_cast<long long>(slot) * g.pad * g.C);
    for (int i = threadIdx.x; i < g.pad * g.C / 8; i += blockDim.x) st[i] = make_uint4(0, 0, 0, 0);
  } else if (blockIdx.y == 8) {
    float4* p = reinterpret_cast<float4*>(partial + static_cast<long long>(slot) * 16 * 512);
    for (int i = threadIdx.x; i < 16 * 512 / 4; i += blockDim.x) p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    if (threadIdx.x == 0) { seqs[slot] = r.sd; ctl[slot] = r.ctl; own_len[slot] = r.own_len; }
    if (threadIdx.x < LDIM) feedback[slot * LDIM + threadIdx.x] = bos[threadIdx.x];
  }
}

// Bring-up probe (PTTS_DIAG_SKIP=4): occupies CTAs for a fixed time without touching memory.
__global__ void spin_kernel(unsigned long long ns) {
  // SM-local clock (about 1.9 cycles per ns): polling %globaltimer from many warps is itself a shared-resource load
  const long long t0 = clock64(), cycles = static_cast<long long>(ns) * 19 / 10;
  while (clock64() - t0 < cycles) { }
}

// The device noise generator on its own (tests: distribution of counter_normal).
__global__ void noise_probe_kernel(unsigned long long seed, int frames, float* __restrict__ out /*[frames,32]*/) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < frames * LDIM) out[i] = counter_normal(seed, i / LDIM, i % LDIM);
}

// Debug: how many f16 values of a buffer are not finite (an f32 -> f16 store saturates to +-inf above 65504).
__global__ void count_nonfinite_f16_kernel(const __half* __restrict__ p, long long n, unsigned long long* __restrict__ count) {
  long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  unsigned int local = 0;
  for (; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const unsigned short u = __half_as_ushort(p[i]);
    local += ((u & 0x7c00u) == 0x7c00u) ? 1u : 0u;
  }
  local = __reduce_add_sync(0xffffffffu, local);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, static_cast<unsigned long long>(local));
}

__global__ void fill_f32_kernel(float* p, float v, long long n) {
  pdl_launch_dependents();
  pdl_wait();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void f32_to_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, long long n) {
  pdl_launch_dependents();
  pdl_wait();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) dst[i] = __float2half_rn(src[i]);
}

}  // namespace ptts
