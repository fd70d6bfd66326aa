/*
 * ptts_internal.h -- test hooks, parity taps and profiling probes of libptts_cuda.so.
 *
 * NOT part of the product ABI (include/ptts.h): a host that replaces the reference's TTSModel binds ptts.h only.  These
 * symbols exist for tests/ (parity against the CPU oracle), bench.py (per-launch timing, launch counts) and bring-up
 * probes; they may change between builds.
 *
 * ptts_engine_cfg.reserved[] test switches (all zero in production):
 *   [0] GEMM operand placement: 1 = never weights-on-M (swap-AB), 2 = always when legal (either also keeps 1-4 row Linear
 *       layers on the tensor-core path instead of the small-batch GEMV)
 *   [1] 1 = never use the small-batch GEMV (csrc/gemv.cuh; Linear layers of 1-4 rows)
 *   [2] force a split-K factor for every non-persistent GEMM
 *   [3] 1 = never use the persistent GEMM kernel
 *   [4] shared-memory budget of a GEMM CTA in KB (default 200)
 *   [5] 1 = flow head as per-layer launches instead of the fused cluster kernel
 *   [6] 1 = decode GEMMs through the staged pipeline instead of the resident-operand path
 *   [7] 1 = int8 mode streams f16 copies of the codes instead of bytes (must be bit-identical)
 *   [8] persistent FlowLM step kernel (csrc/lm_step.cuh): 0 = library default, 1 = never, 2 = always when legal
 */
#ifndef PTTS_INTERNAL_H
#define PTTS_INTERNAL_H

#include "ptts.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Parity taps: copy a named intermediate of the last step for batch row `row`
 * (e.g. "flowlm.h", "mimi.after_upsample", "mimi.after_decoder_transformer", "seanet.convtr2").
 * Returns the number of floats written or a negative status. */
int64_t ptts_debug_read(ptts_engine* e, const char* name, int32_t row, float* out, int64_t cap);

/* Device-time accounting for bench.py: kernel launches issued by the engine since the last reset,
 * and CUDA-event time (ms) of the stages of the most recent ptts_step_timed call. */
int64_t ptts_launch_count(ptts_engine* e, int32_t reset);
int32_t ptts_step_timed(ptts_engine* e, const int32_t* slots, int32_t n, float* stage_ms /*[8]*/);
void* ptts_cuda_stream(ptts_engine* e);
/* Per-launch CUDA-event profile for bench.py's roofline line.  While enabled every kernel launch is
 * bracketed by events on the engine's stream; the report has one text line per kernel class:
 * "<class> <launches> <total_ms> <algorithmic_bytes> <algorithmic_flops>".  Returns the string length. */
int32_t ptts_profile_enable(ptts_engine* e, int32_t on);
int64_t ptts_profile_report(ptts_engine* e, char* buf, int64_t cap);
/* Event time (ms) of an empty kernel bracketed the same way: the fixed cost inside every per-launch figure. */
int32_t ptts_profile_overhead(ptts_engine* e, float* ms_out);

/* Isolated kernel entry points (tests/test_kernels_gpu.py): D[r,f] = sum_k A[r,k] * W[f,k] on
 * host f32 buffers, run through the production GEMM (operands converted to f16).  mode 0 lets the
 * engine choose the tiling, 1 forces activation-as-M, 2 forces weight-as-M (swap-AB);
 * split_k > 1 exercises the cluster split-K epilogue (partials summed over DSMEM in rank order, no atomics). */
int32_t ptts_test_gemm(int32_t device, const float* a, const float* w, const float* bias, float* d, int32_t rows,
                       int32_t feats, int32_t k, int32_t mode, int32_t split_k, int32_t act, int32_t use_simt);
/* The int8 weight path of the decode (swap-AB) GEMM: w is quantised per tensor with the reference's scheme
 * (crates/pocket-tts/src/quantize.rs:65-94: scale = absmax / 127, codes clamp(round(w / scale), -127, 127)),
 * D = (A . codes^T) * scale.  storage 1 streams one-byte codes from HBM and expands them in shared memory
 * (production), 0 streams an f16 copy of the same codes; both must give bit-identical D.  storage 2 = byte codes with the
 * library's own choice of kernel (1-4 rows: the GEMV of csrc/gemv.cuh expanding the codes in registers).  scale_out gets
 * the scale. */
int32_t ptts_test_gemm_int8(int32_t device, const float* a, const float* w, float* d, int32_t rows, int32_t feats,
                            int32_t k, int32_t split_k, int32_t storage, float* scale_out);
/* Launches of the small-batch GEMV (csrc/gemv.cuh) by this process so far: tests assert the family was actually selected. */
int64_t ptts_test_gemv_launches(void);
/* Bring-up probe: back-to-back launches of one GEMM with per-CTA %globaltimer stamps (10 per CTA, ns). */
int32_t ptts_test_gemm_trace(int32_t device, int32_t rows, int32_t feats, int32_t k, int32_t mode, int32_t split_k,
                             int32_t iters, float* us_per_launch, int64_t* stamps, int32_t max_ctas, int32_t* n_ctas);
/* Implicit-GEMM streaming convs on host buffers, channels-last x [n, t, cin], state [n, k-1 (or 1), cin]. */
int32_t ptts_test_conv1d(int32_t device, const float* x, const float* prev, const float* w /*[cout,cin,k]*/,
                         const float* bias, float* y /*[n,t,cout]*/, int32_t n, int32_t t, int32_t cin, int32_t cout,
                         int32_t k);
int32_t ptts_test_convtr1d(int32_t device, const float* x, const float* prev_row, const float* w /*[cin,cout,2s]*/,
                           const float* bias, float* y /*[n,t*s,cout]*/, int32_t n, int32_t t, int32_t cin,
                           int32_t cout, int32_t stride);

/* Kernel time of the decode GEMMs free of per-launch event cost: in_proj / out_proj / linear1 / linear2 at `rows` batch
 * rows replayed as one graph of iters x 6 layers of back-to-back launches per kind (real weights of every layer in turn, so
 * each launch streams its weights from HBM), one event pair around the graph.  us_out[4], bytes_out[4] (algorithmic). */
int32_t ptts_profile_gemm_replay(ptts_engine* e, int32_t rows, int32_t iters, float* us_out, double* bytes_out);
/* The device noise generator on its own: out[frames * 32] = the N(0, 1) draws a stream with this seed would get
 * (before the sqrt(temp) scale); tests check its distribution. */
int32_t ptts_test_noise(int32_t device, uint64_t seed, int32_t frames, float* out);
/* Debug: number of non-finite values in the f16 operand buffers of the last step (an f32 -> f16 store saturates to
 * +-inf above 65504): the overflow counter for checkpoints whose activations outgrow f16. */
int32_t ptts_debug_f16_overflow(ptts_engine* e, int64_t* count_out);

#ifdef __cplusplus
}
#endif
#endif /* PTTS_INTERNAL_H */
