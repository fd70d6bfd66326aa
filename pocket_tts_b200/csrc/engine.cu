// libptts_cuda.so: engine + C ABI (include/ptts.h).
//
// Host-side mirror of the reference's L3/L4 for the generation hot path
// (crates/pocket-tts/src/tts_model.rs:935-1071, models/flow_lm.rs, models/mimi.rs, models/seanet.rs)
// re-designed for one B200: all per-stream state lives in HBM in a slot arena, every call advances a
// whole batch of independent streams, and the only host<->device traffic per step is the PCM frame and
// the finished flags.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <memory>
#include <mutex>
#include <unordered_map>

#include <atomic>
#include "flow_head.cuh"
#include "flow_small.cuh"
#include "gemm.cuh"
#include "gemv.cuh"
#include "host_util.h"
#include "../../include/ptts_internal.h"
#include "kernels.cuh"
#include "lm_step.cuh"
#include "seanet_tail.cuh"

namespace ptts {

static thread_local std::string g_last_error;

// Every kernel goes through here: programmatic dependent launch (the kernel's prologue overlaps its
// predecessor's tail; see pdl_wait in ptx.cuh) and, for split-K GEMMs, a thread-block cluster along z.
template <typename... KArgs, typename... Args>
static void launch_k(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_z,
                     Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attrs[2];
  int na = 0;
  if (pdl) {
    attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (cluster_z > 1) {
    attrs[na].id = cudaLaunchAttributeClusterDimension;
    attrs[na].val.clusterDim.x = 1; attrs[na].val.clusterDim.y = 1; attrs[na].val.clusterDim.z = cluster_z;
    ++na;
  }
  cfg.attrs = attrs; cfg.numAttrs = na;
  PTTS_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}

// ------------------------------------------------------------------------------------------------
// model constants (crates/pocket-tts/config/b6369a24.yaml)
static constexpr int D_MODEL = 1024, N_HEADS = 16, N_LAYERS = 6, D_FFN = 4096;
static constexpr int FLOW_DIM = 512, FLOW_DEPTH = 6, MOD_LD = FLOW_DEPTH * 3 * FLOW_DIM + 2 * FLOW_DIM;  // 10240
static constexpr int MIMI_DIM = 512, MIMI_HEADS = 8, MIMI_LAYERS = 2, MIMI_FFN = 2048, MIMI_T = 16;
static constexpr int N_BINS = 4000;
static constexpr size_t PF_SMEM_MAX = 200 * 1024;   // K + V rows a prefill-attention tile may stage (704 keys); longer ranges use the row-per-CTA kernel
static constexpr int MAX_LSD = 64;   // ptts_engine_set_lsd_steps cap; time_emb is sized for it once

struct Weight16 {   // GEMM operand [Fpad][K] f16, K-major
  DevBuf<__half> w;
  DevBuf<float> wscale;  // int8 mode: the operand holds the integer codes, wscale[f] the per-tensor scale of row f
  DevBuf<int8_t> q8;     // int8 mode: the same codes as one byte each, [Fpad][K]; what the decode (swap-AB) GEMMs stream from HBM
  int F = 0, Fpad = 0, K = 0;
};

struct ActView {    // f16 activations [cap][Tpad][C], channels-last
  const __half* ptr;
  int C, Tpad, cap;
};

struct HostTensor {
  const ptts_tensor_desc* d;
  float qscale = 0.f;      // int8 mode: per-tensor scale (0 = tensor kept in full precision)
  std::vector<float> f32;  // converted copy (fake-quantised q*scale in int8 mode)
  std::vector<int64_t> shape;
  size_t numel() const { size_t n = 1; for (auto s : shape) n *= (size_t)s; return n; }
};

static float bf16_to_f32(uint16_t v) { uint32_t u = (uint32_t)v << 16; float f; std::memcpy(&f, &u, 4); return f; }

// ------------------------------------------------------------------------------------------------
struct Voice {
  DevBuf<__half> kv;  // [layer][k|v][head][len][64]
  int len = 0;
  std::vector<float> prompt;  // the conditioning rows [len][1024] the KV was built from (`audio_prompt`, for ptts_voice_save)
};

struct SlotHost {
  bool in_use = false;
  bool finished = false;
  int frames = 0;
  int eos_step = -1;
  int own_len = 0;
  int max_gen_len = 0;
  Voice* voice = nullptr;
};

struct Engine {
  ptts_engine_cfg cfg{};
  cudaStream_t stream = nullptr;    // A: language-model path, prefill, host copies of the step flags
  cudaStream_t stream_b = nullptr;  // B: codec path (Mimi transformer + SEANet)
  cudaStream_t ls = nullptr;        // stream the launch helpers currently target
  cudaEvent_t ev_a_done = nullptr, ev_front_done = nullptr, ev_b_done = nullptr;
  TmapCache tmaps;
  long long launches = 0;
  bool use_pdl = true;
  int diag_skip = 0;
  bool diag_times = false;      // PTTS_DIAG_TIMES=1: event stamps around both paths of the last step, printed by sync
  cudaEvent_t ev_t[4] = {};
  int trig_a = 1, trig_b = 1;   // GemmParams::pdl_trigger per step stream (PTTS_TRIG_A / PTTS_TRIG_B)
  int split_cta_cap = 48;   // a split-K decode (swap-AB) GEMM never spans more CTAs than this (PTTS_MAX_CTAS)
  int split_cta_cap_b = 48; // the same for the activation-as-M GEMMs of the codec half (PTTS_MAX_CTAS_B)
  int split_cap_override = 0;  // one-shot cap for the next GEMM
  int bn_override = 0;         // one-shot feature-tile width for the next (activation-as-M, non-persistent) GEMM
  struct Tune { int bn = 0, cap = 0; };   // per-call-site (bn, split cap) of the codec GEMMs: PTTS_TUNE_<SITE>=bn,cap
  Tune tune_conv0, tune_ct2, tune_mlin1, tune_mlin2, tune_minproj, tune_moutproj, tune_ct5;
  void apply(const Tune& t) { if (t.bn > 0) bn_override = t.bn; if (t.cap > 0) split_cap_override = t.cap; }
  int gemm_ref_f = 1;          // > 1 while the GEMMs of a codec group of that many frames are issued (see Engine::gemm)
  int lin2_ctas = 0, outproj_ctas = 0, inproj_ctas = 0;   // the same for the other decode GEMMs (0 = the common cap; PTTS_LIN2_CTAS / PTTS_OUTPROJ_CTAS / PTTS_INPROJ_CTAS)
  int lin1_ctas = 64;          // linear1 (32 feature tiles): 64 lets it split in two and keep its K slice resident (PTTS_LIN1_CTAS)
  int persistent_ctas = 132;  // grid of the persistent (codec) GEMMs; fewer leaves SMs to the other stream
  int num_sms = 148;          // grid of the small-batch GEMV (gemv.cuh); PTTS_GEMV_CTAS
  int lsd_steps = 1;
  // per-launch CUDA-event profiling (bench.py roofline pass; off in the timed region)
  struct ProfRec { const char* tag; const char* fn; cudaEvent_t a, b; double bytes, flops; };
  bool profiling = false;
  std::vector<ProfRec> prof_recs;
  std::vector<cudaEvent_t> prof_pool;
  const char* cur_tag = nullptr;
  unsigned long long* gemm_trace = nullptr;
  // LayerNorm requested behind the next GEMM (its own launch; a grid-barrier fusion measured -3% and was removed)
  struct LnSpec { bool set = false; const float* x; int rows, C; const float* w; const float* b; float eps; const float* shift;
                  const float* scale; int mod_ld; __half* out; int out_ld; const char* tag; };
  LnSpec next_ln;
  void ln_pending(const LnSpec& lnreq);
  // LayerNorm in FRONT of the next Linear: folded into the small-batch GEMV's prologue (gemv.cuh) when that kernel takes
  // the call, else launched on its own first.  Only set where ln_fusable() said yes.
  LnSpec pre_ln;
  bool gemv_ln = true;   // PTTS_GEMV_LN=0: LayerNorm launches kept
  bool ln_fusable(int rows, const Weight16& w) const;
  bool gemv(const __half* x, int rows, const Weight16& w, int F, const GemmEpi& epi, const LnSpec& pre);
  void ln_after_next_gemm(const char* tag, const float* x, int rows, int C, const float* w, const float* b, float eps,
                          const float* shift, const float* scale, int mod_ld, __half* out, int out_ld) {
    next_ln = LnSpec{true, x, rows, C, w, b, eps, shift, scale, mod_ld, out, out_ld, tag};
  }
  // bring-up: passed to the next GEMM launches
  double step_kv_bytes = 0;  // FlowLM KV bytes one layer's decode attention reads in the current step
  void tag(const char* t) { cur_tag = t; }
  const char* take_tag(const char* dflt) { const char* t = cur_tag ? cur_tag : dflt; cur_tag = nullptr; return t; }
  void prof_begin(const char* t, double bytes, double flops, const char* fn);
  void prof_end();
  std::string prof_report();
  int NB = 0, NS = 0, KVCAP = 0, PR = 0;  // max batch, slots, own kv rows, prefill rows

  // ---- weights
  std::unordered_map<std::string, HostTensor> host;
  Weight16 w_input, w_inproj[N_LAYERS], w_outproj[N_LAYERS], w_lin1[N_LAYERS], w_lin2[N_LAYERS];
  DevBuf<float> ln1_w[N_LAYERS], ln1_b[N_LAYERS], ln2_w[N_LAYERS], ln2_b[N_LAYERS];
  DevBuf<float> outnorm_w, outnorm_b, eos_w, eos_b, bos, emb_std, emb_mean, lut;
  Weight16 w_cond, w_finproj, w_ada, w_mlp0[FLOW_DEPTH], w_mlp2[FLOW_DEPTH], w_final;
  DevBuf<__half> w_flowpack;   // mlp.0 / mlp.2 of the six blocks and the final Linear, one [6272][512] operand (flow_head.cuh)
  bool flow_small = false;     // PTTS_FLOW_SMALL=1: 1-4 rows through the 8-CTA cluster GEMV form of the flow head (flow_small.cuh).  Opt-in:
                               // 50 -> 33 us per launch but 236 -> 231 us per frame only, and one unexplained test failure in three runs of the round
  bool fused_flow = true;      // ptts_engine_cfg.reserved[5] = 1 or debug_gemm: the per-layer launches instead
  static constexpr int MOD_STEPS = 4;   // Euler steps whose modulation rows fit the scratch: one Linear + one flow-head launch for all of them
  bool mod_all_steps() const { return lsd_steps <= MOD_STEPS; }
  void flow_head_fused(int n, const float* mod, long long mod_step_stride, int steps);
  // ---- persistent FlowLM step kernel (lm_step.cuh): tiled weight images, operand images, split-K workspace, grid barrier
  bool lm_enabled = false;     // built at init when selected (see lm_build) and the weights are f16
  int lm_ctas = 0;             // grid of the step kernel (PTTS_LM_CTAS; default: every SM)
  int lm_flags = 0;            // LmStepParams::flags (PTTS_LM_FLAGS, bring-up)
  DevBuf<uint8_t> lm_wt, lm_hA, lm_attnA, lm_ffnA, lm_yA;
  DevBuf<float> lm_ws, lm_mod;
  DevBuf<unsigned long long> lm_bar, lm_trace;
  LmStepParams lm_params{};
  double lm_weight_bytes = 0;
  void lm_build();
  bool lm_usable(int n) const { return lm_enabled && n >= 1 && n <= LM_ROWS && lsd_steps <= LM_MAX_LSD; }
  void lm_step(int n);
  DevBuf<unsigned long long> fh_trace;  // PTTS_FH_TRACE=1: stage stamps of the fused flow head, printed by sync
  DevBuf<float> b_cond, b_finproj, b_ada, b_mlp0[FLOW_DEPTH], b_mlp2[FLOW_DEPTH], b_final, inln_w[FLOW_DEPTH], inln_b[FLOW_DEPTH];
  DevBuf<float> time_emb;  // [S,512]
  DevBuf<float> wq, wup;
  Weight16 m_inproj[MIMI_LAYERS], m_outproj[MIMI_LAYERS], m_lin1[MIMI_LAYERS], m_lin2[MIMI_LAYERS];
  DevBuf<float> m_ln1_w[MIMI_LAYERS], m_ln1_b[MIMI_LAYERS], m_ln2_w[MIMI_LAYERS], m_ln2_b[MIMI_LAYERS], m_ls1[MIMI_LAYERS], m_ls2[MIMI_LAYERS];
  // Mimi encoder side (voice cloning from PCM); present only when the checkpoint carries the encoder tensors
  bool has_encoder = false;
  DevBuf<float> en_conv0_w, en_conv0_b;                   // [7][64], [64]
  Weight16 en_r1[3], en_r3[3], en_down[3], en_c11;         // ResBlock k3 / k1 convs and the strided conv of each level; last conv
  DevBuf<float> enb_r1[3], enb_r3[3], enb_down[3], enb_c11;
  Weight16 et_inproj[MIMI_LAYERS], et_outproj[MIMI_LAYERS], et_lin1[MIMI_LAYERS], et_lin2[MIMI_LAYERS], en_ds, w_spk;
  DevBuf<float> et_ln1_w[MIMI_LAYERS], et_ln1_b[MIMI_LAYERS], et_ln2_w[MIMI_LAYERS], et_ln2_b[MIMI_LAYERS], et_ls1[MIMI_LAYERS], et_ls2[MIMI_LAYERS];
  DevBuf<unsigned char> enc_arena;   // encoder scratch, grown on demand, reused by every voice_from_pcm call
  void load_encoder_weights();
  void encode_prompt(const float* pcm_host, int n_samples, std::vector<float>& prompt, int* frames);
  Weight16 s_conv0, s_ct2, s_r3a, s_r3b, s_ct5, s_r6a, s_r6b, s_ct8, s_r9a, s_r9b;
  DevBuf<float> sb_conv0, sb_ct2, sb_r3a, sb_r3b, sb_ct5, sb_r6a, sb_r6b, sb_ct8, sb_r9a, sb_r9b, s_final_w, s_final_b;

  // ---- per-slot state
  DevBuf<__half> kv;            // [NS][layer][2][H][KVCAP][64]
  DevBuf<SeqDesc> seqs;         // NS + 1 (last = scratch sequence used while building a voice)
  DevBuf<int> own_len;          // NS + 1
  DevBuf<StreamCtl> ctl;        // NS
  DevBuf<float> feedback;       // [NS,32]
  DevBuf<float> up_partial;     // [NS,16,512]
  DevBuf<__half> mimi_ring;     // [NS][2][2][8][272][64]
  DevBuf<__half> st_tr, st_a0, st_e2, st_a3, st_e5, st_a6, st_e8, st_a9;  // conv left-context rows per slot
  std::vector<SlotHost> slots;
  std::vector<unsigned char> slot_mark;   // check_slots scratch
  std::vector<std::unique_ptr<Voice>> voices;

  // ---- decode scratch (compact by batch row)
  DevBuf<int> row_seq;
  std::vector<int> row_seq_host;
  DevBuf<float> x32, qkv32, eos_logit, c32, mod32, fx32, z32, h32dbg, quant_dbg;
  DevBuf<__half> h16, attn16, ffn16, lat16, y16, fh16, fg16, z16;
  DevBuf<float> mx32, mqkv32;
  DevBuf<int> mimi_pos;  // per batch row: absolute position of the frame's first Mimi token (snapshot by front)
  DevBuf<__half> mh16, mattn16, mffn16;
  DevBuf<__half> tr16, a0, e2, h3, a3, e5, h6, a6, e8, h9, a9;
  DevBuf<float> x2, x5, x8, pcm;
  DevBuf<short> pcm16;          // the same frame as i16 (audio.rs:129-146), written by the last SEANet conv
  short* pin_pcm16[6] = {};   // [NT]
  // stream open: pinned records -> one copy -> scatter kernel; injected noise in per-slot buffers that only grow (a
  // cudaFree per close would be a device-wide synchronisation)
  DevBuf<OpenRec> open_recs;
  OpenRec* pin_open[2] = {};
  cudaEvent_t ev_open[2] = {};
  int open_parity = 0;
  std::vector<DevBuf<float>> noise_pool;
  // per-step host-visible results in ONE buffer ([NB][32] latents | [NB] EOS logits | [NB] finished bytes) so that a
  // step's flags reach the host with one D2H copy on the language-model stream instead of three
  DevBuf<SeqDesc> row_desc;    // [NB] per batch row: KV descriptor + cursor of its stream, rebuilt by step_begin_kernel
  DevBuf<unsigned char> step_out;
  struct { float* p; } latent_out, logit_out;
  struct { unsigned char* p; } finished_dev;
  size_t step_out_bytes() const { return (size_t)NB * (LDIM * 4 + 4 + 1); }
  ConvSegs segs{};             // one frame per codec pass (also what stream open zeroes)
  ConvSegs segs_by_f[5]{};     // [frames per codec pass]
  const ConvSegs& segs_f(int f) const { return segs_by_f[f]; }
  // ---- codec group (set_codec_group): A queues the latents of cg consecutive frames, the codec half runs once per group
  int cg = 1;                  // frames per codec pass
  bool use_queue = true;       // (PTTS_QUEUE=0: off) cg == 1 also hands its latent over through the (double-buffered) queue: A(n+1) then waits for front(n-1), not front(n)
  bool queued() const { return cg > 1 || use_queue; }
  int pend = 0, pend_n = 0;    // frames waiting in the current queue buffer, their batch rows
  int gbuf = 0;                // queue buffer the current group fills
  std::vector<int> pend_rows;  // batch composition of the waiting frames
  long long pend_ticket[4] = {-1, -1, -1, -1};
  DevBuf<float> zq;            // [2][cg][NB][32]
  DevBuf<int> zq_pos;          // [2][cg][NB] frame index of the queued latent
  float* cur_zq = nullptr; int* cur_zqpos = nullptr;   // queue entry step_part_a writes (null: none)
  cudaEvent_t ev_gfront[2] = {};
  void flush_codec();
  void set_codec_group(int frames);
  bool fused_tail = true;      // the last SEANet block + final conv as one kernel (seanet_tail.cuh); PTTS_SEANET_TAIL=0: three launches
  void seanet_tail(int n, int T);
  // SM partition between the two streams (green contexts); 0 SMs = none
  int sms_a = 0, sms_b = 0;
  void* green_ctx[2] = {nullptr, nullptr};
  void green_split(int b_sms, int prio_a, int prio_b);
  // ---- prefill scratch
  DevBuf<float> px32, pqkv32, pqrot;
  DevBuf<__half> ph16, pattn16, pffn16;
  DevBuf<int> prow_seq, prow_pos, ptokens;
  DevBuf<int2> ptiles;            // prefill attention tiles (first row, rows) of the rows in prow_seq / prow_pos
  int pf_tiles = 0, pf_kmax = 0;  // tiles of the pending prefill and the longest key range any of them stages (0 tiles: row-per-CTA kernel)
  bool pf_tiled = true;           // PTTS_PREFILL_TILE=0: the row-per-CTA SIMT prefill attention instead of the tensor-core tiles
  void set_prefill_tiles(const std::vector<int>& rs, const std::vector<int>& rp);
  // ---- pinned staging
  // a ring of tickets: with PTTS_STEP_AHEAD step n+1 is enqueued while the flags of step n and the PCM of step n-1 are
  // still on their way to the host; with a codec group the PCM of a frame leaves with its group, up to cg - 1 steps later
  static constexpr int NT = 6;
  static_assert(NT == sizeof(pin_pcm16) / sizeof(pin_pcm16[0]), "pin_pcm16 holds one buffer per ticket");
  float* pin_pcm[NT] = {}; unsigned char* pin_fin[NT] = {};
  float* pin_lat[NT] = {}; float* pin_logit[NT] = {};
  cudaEvent_t ev_flags[NT] = {}, ev_pcm[NT] = {};
  struct Ticket { long long id = -1; int n = 0; bool flags_done = true, pcm_done = true, want_pcm = false, want_i16 = false, codec_pending = false; std::vector<int> slot_ids; };
  Ticket tickets[NT];
  // a ticket whose flags have not been fetched still lists its slots: closing / reopening one of them in between would make
  // step_flags_impl book the overrun row of the OLD stream onto the NEW one
  bool slot_in_pending_ticket(int s) const {
    for (const Ticket& t : tickets)
      if (!t.flags_done)
        for (int id : t.slot_ids) if (id == s) return true;
    return false;
  }
  long long next_ticket = 0;
  void sync_all() { flush_codec(); PTTS_CUDA(cudaStreamSynchronize(stream)); PTTS_CUDA(cudaStreamSynchronize(stream_b)); }
  cudaEvent_t ev[10]{};

  ~Engine();
  void init(const ptts_engine_cfg& c, const ptts_tensor_desc* w, int nw);
  void load_weights(const ptts_tensor_desc* w, int nw);
  const HostTensor& T(const std::string& name, std::initializer_list<int64_t> shape);
  void vec(DevBuf<float>& dst, const std::string& name, int64_t n);
  void linear(Weight16& dst, const std::string& name, int F, int K, int Kpad = 0, bool codes = true);
  void compute_time_embeddings(int steps);

  void gemm(const ActView& a, int n_streams, int T, int taps, int R, int G, const Weight16& w, int F, GemmEpi epi,
            bool allow_split = false);
  void gemm_rows(const __half* a, int rows, int K, const Weight16& w, int F, GemmEpi epi, bool allow_split = false) {
    gemm(ActView{a, K, std::max(rows, 1), 1}, 1, rows, 1, 0, 1, w, F, epi, allow_split);
  }
  template <int C>
  void ln(const float* x, int rows, const float* w, const float* b, float eps, const float* shift, const float* scale,
          int mod_ld, __half* out, int out_ld);
  void flowlm_layers(int rows, float* x, __half* h, float* qkv, __half* attn, __half* ffn, bool prefill, float* qrot,
                     const int* rseq, const int* rpos);
  void upload_rows(const int* slot_ids, int n);
  void step_kernels(int n, float* stage_ms);
  void step_part_a(int n, bool marks);
  void step_front(int n, int f, int qbuf);
  void step_part_b(int n, int f, bool marks);
  cudaGraphExec_t capture(cudaStream_t st, int n, int part, int arg, long long* kernels);
  void run_step(int n, long long ticket = -1);
  struct PartGraph { cudaGraphExec_t exec; long long kernels; };
  // (part, batch rows, lsd steps [A only], queue entry [A] | frames per row [B]) -> captured half of a decode step
  std::map<std::tuple<int, int, int, int>, PartGraph> graphs;
  PartGraph& graph_of(int part, int n, int arg);
  void drop_graphs();
  void prefill(int rows);
};

template <typename Fn>
static Fn driver_fn(const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  PTTS_CUDA(cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q));
  PTTS_REQUIRE(p && q == cudaDriverEntryPointSuccess, PTTS_ERR_CUDA, "%s not available in this driver", name);
  return reinterpret_cast<Fn>(p);
}
#define PTTS_CU(expr) do { CUresult r_ = (expr); PTTS_REQUIRE(r_ == CUDA_SUCCESS, PTTS_ERR_CUDA, "%s failed (%d)", #expr, (int)r_); } while (0)

Engine::~Engine() {
  for (int i = 0; i < NT; ++i) {
    if (pin_pcm[i]) cudaFreeHost(pin_pcm[i]);
    if (pin_pcm16[i]) cudaFreeHost(pin_pcm16[i]);
    if (pin_lat[i]) cudaFreeHost(pin_lat[i]);  // pin_logit / pin_fin point into it
    if (ev_flags[i]) cudaEventDestroy(ev_flags[i]);
    if (ev_pcm[i]) cudaEventDestroy(ev_pcm[i]);
  }
  for (int i = 0; i < 2; ++i) { if (pin_open[i]) cudaFreeHost(pin_open[i]); if (ev_open[i]) cudaEventDestroy(ev_open[i]); }
  for (auto& e : ev) if (e) cudaEventDestroy(e);
  for (auto& e : prof_pool) cudaEventDestroy(e);
  for (auto& g : graphs) cudaGraphExecDestroy(g.second.exec);
  for (auto& x : ev_gfront) if (x) cudaEventDestroy(x);
  if (ev_a_done) cudaEventDestroy(ev_a_done);
  if (ev_front_done) cudaEventDestroy(ev_front_done);
  if (ev_b_done) cudaEventDestroy(ev_b_done);
  if (stream_b) cudaStreamDestroy(stream_b);
  for (auto& r : prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  if (stream) cudaStreamDestroy(stream);
  if (green_ctx[0]) {
    try {
      auto destroy = driver_fn<CUresult (*)(CUgreenCtx)>("cuGreenCtxDestroy");
      for (void* g : green_ctx) if (g) destroy(static_cast<CUgreenCtx>(g));
    } catch (...) {}
  }
}

void Engine::prof_begin(const char* t, double bytes, double flops, const char* fn) {
  if (!profiling) return;
  cudaEvent_t ev2[2];
  for (auto& x : ev2) {
    if (prof_pool.empty()) { cudaEvent_t n; PTTS_CUDA(cudaEventCreate(&n)); prof_pool.push_back(n); }
    x = prof_pool.back();
    prof_pool.pop_back();
  }
  PTTS_CUDA(cudaEventRecord(ev2[0], ls));
  prof_recs.push_back(ProfRec{t, fn ? fn : t, ev2[0], ev2[1], bytes, flops});
}
void Engine::prof_end() {
  if (!profiling) return;
  PTTS_CUDA(cudaEventRecord(prof_recs.back().b, ls));
}
struct ProfScope {
  Engine& e;
  ProfScope(Engine& en, const char* t, double bytes = 0, double flops = 0, const char* fn = nullptr) : e(en) {
    e.prof_begin(t, bytes, flops, fn);
    ++e.launches;
  }
  ~ProfScope() { e.prof_end(); }
};
// one line per kernel class: "tag kernel_function launches total_ms bytes flops" (bytes/flops are the algorithmic totals)
std::string Engine::prof_report() {
  PTTS_CUDA(cudaStreamSynchronize(stream));
  struct Agg { int n = 0; double ms = 0, bytes = 0, flops = 0; const char* fn = ""; };
  std::map<std::string, Agg> agg;
  for (auto& r : prof_recs) {
    float ms = 0;
    PTTS_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
    Agg& g = agg[r.tag];
    g.n++; g.ms += ms; g.bytes += r.bytes; g.flops += r.flops; g.fn = r.fn;
    prof_pool.push_back(r.a); prof_pool.push_back(r.b);
  }
  prof_recs.clear();
  std::string out;
  for (auto& kv : agg)
    out += fmt("%s %s %d %.6f %.0f %.0f\n", kv.first.c_str(), kv.second.fn, kv.second.n, kv.second.ms, kv.second.bytes, kv.second.flops);
  return out;
}

// ------------------------------------------------------------------------------------------------ weights
const HostTensor& Engine::T(const std::string& name, std::initializer_list<int64_t> shape) {
  auto it = host.find(name);
  PTTS_REQUIRE(it != host.end(), PTTS_ERR_INVALID, "missing tensor '%s'", name.c_str());
  std::vector<int64_t> want(shape);
  PTTS_REQUIRE(it->second.shape == want, PTTS_ERR_INVALID, "tensor '%s' has the wrong shape", name.c_str());
  return it->second;
}

void Engine::vec(DevBuf<float>& dst, const std::string& name, int64_t n) {
  const HostTensor& t = T(name, {n});
  dst.alloc(n);
  PTTS_CUDA(cudaMemcpy(dst.p, t.f32.data(), n * sizeof(float), cudaMemcpyHostToDevice));
}

// `scales` (int8 mode, one per row or empty): rows holding q*scale are stored as their integer codes q, which f16
// represents exactly, and the scale goes to wscale[] for the epilogue; acc(q) * scale == acc(q*scale) up to f32
// rounding, so the numerics are the reference's (quantize.rs:65-94) while the operand is ready for 8-bit storage.
// scales non-empty = int8 mode.  codes = true: the operand holds the integer codes (exact in f16) and wscale[f] their
// scale, applied to the f32 accumulator; used by the FlowLM / flow-head decode GEMMs, which also get the one-byte copy.
// codes = false: the operand holds f16(code * scale), the same rounding every unquantised weight goes through; used by
// the codec GEMMs (Mimi, SEANet), which stream activations, not weights, and whose epilogues then need no scale.
static void upload_f16(Weight16& dst, const std::vector<float>& rows, int F, int K, const std::vector<float>& scales = {},
                       bool codes = true) {
  dst.F = F;
  dst.K = K;
  dst.Fpad = round_up(F, 128);
  std::vector<__half> h((size_t)dst.Fpad * K, __float2half(0.f));
  double worst = 0;
  for (int f = 0; f < F; ++f) {
    const float sc = scales.empty() ? 0.f : scales[f];
    for (int k = 0; k < K; ++k) {
      const size_t i = (size_t)f * K + k;
      if (sc > 0.f) {
        const float q = std::nearbyint(rows[i] / sc);
        h[i] = __float2half_rn(codes ? q : q * sc);
      } else {
        h[i] = __float2half_rn(rows[i]);
        const double err = std::fabs((double)__half2float(h[i]) - rows[i]);
        if (err > worst) worst = err;
      }
    }
  }
  // bf16-representable weights convert exactly unless |w| < 2^-17 or > 65504; report, never hide
  if (worst > 1e-6) fprintf(stderr, "ptts: f16 weight conversion max abs error %.3g\n", worst);
  dst.w.alloc(h.size());
  PTTS_CUDA(cudaMemcpy(dst.w.p, h.data(), h.size() * sizeof(__half), cudaMemcpyHostToDevice));
  bool any = false;
  for (float v : scales) any = any || v > 0.f;
  if (any && codes) {
    std::vector<float> ws(dst.Fpad, 1.f);
    for (int f = 0; f < F; ++f) if (scales[f] > 0.f) ws[f] = scales[f];
    dst.wscale.alloc(ws.size());
    PTTS_CUDA(cudaMemcpy(dst.wscale.p, ws.data(), ws.size() * sizeof(float), cudaMemcpyHostToDevice));
    bool all = true;
    for (int f = 0; f < F; ++f) all = all && scales[f] > 0.f;
    if (all && K % 64 == 0) {  // every row quantised: one-byte codes for the in-kernel dequant path (gemm.cuh, w_int8)
      std::vector<int8_t> q((size_t)dst.Fpad * K, 0);
      for (size_t i = 0; i < (size_t)F * K; ++i) q[i] = (int8_t)__half2float(h[i]);  // |code| <= 127, exact
      dst.q8.alloc(q.size());
      PTTS_CUDA(cudaMemcpy(dst.q8.p, q.data(), q.size(), cudaMemcpyHostToDevice));
    }
  }
}

void Engine::linear(Weight16& dst, const std::string& name, int F, int K, int Kpad, bool codes) {
  const HostTensor& t = T(name, {F, K});
  const std::vector<float> sc = t.qscale > 0.f ? std::vector<float>(F, t.qscale) : std::vector<float>();
  if (Kpad <= K) return upload_f16(dst, t.f32, F, K, sc, codes);
  std::vector<float> p((size_t)F * Kpad, 0.f);
  for (int f = 0; f < F; ++f) std::memcpy(&p[(size_t)f * Kpad], &t.f32[(size_t)f * K], K * sizeof(float));
  upload_f16(dst, p, F, Kpad, sc, codes);
}

// Conv1d weight [cout, cin, k] -> [cout_pad][tap*cin_pad + c]  (tap-major K so the K loop walks taps)
static void conv_weight(Weight16& dst, DevBuf<float>& bias_dst, const HostTensor& w, const HostTensor& b, int cout, int cin,
                        int k, int cout_pad, int cin_pad) {
  std::vector<float> g((size_t)cout_pad * k * cin_pad, 0.f), bb(cout_pad, 0.f);
  for (int o = 0; o < cout; ++o) {
    bb[o] = b.f32[o];
    for (int c = 0; c < cin; ++c)
      for (int j = 0; j < k; ++j) g[((size_t)o * k + j) * cin_pad + c] = w.f32[((size_t)o * cin + c) * k + j];
  }
  upload_f16(dst, g, cout_pad, k * cin_pad, w.qscale > 0.f ? std::vector<float>(cout_pad, w.qscale) : std::vector<float>(), false);
  bias_dst.alloc(cout_pad);
  PTTS_CUDA(cudaMemcpy(bias_dst.p, bb.data(), bb.size() * sizeof(float), cudaMemcpyHostToDevice));
}

// ConvTranspose1d weight [cin, cout, 2s] -> [rho*cout + o][tap*cin + c]; tap 0 multiplies x[t-1] (kernel index
// rho+s), tap 1 multiplies x[t] (kernel index rho): y[t*s+rho] = x[t] W[:,:,rho] + x[t-1] W[:,:,rho+s].
static void convtr_weight(Weight16& dst, DevBuf<float>& bias_dst, const HostTensor& w, const HostTensor& b, int cin, int cout,
                          int s) {
  const int k = 2 * s;
  std::vector<float> g((size_t)s * cout * 2 * cin, 0.f), bb((size_t)s * cout);
  for (int rho = 0; rho < s; ++rho)
    for (int o = 0; o < cout; ++o) {
      bb[(size_t)rho * cout + o] = b.f32[o];
      float* row = &g[((size_t)rho * cout + o) * 2 * cin];
      for (int c = 0; c < cin; ++c) {
        row[c] = w.f32[((size_t)c * cout + o) * k + rho + s];
        row[cin + c] = w.f32[((size_t)c * cout + o) * k + rho];
      }
    }
  upload_f16(dst, g, s * cout, 2 * cin, w.qscale > 0.f ? std::vector<float>((size_t)s * cout, w.qscale) : std::vector<float>(), false);
  bias_dst.alloc(bb.size());
  PTTS_CUDA(cudaMemcpy(bias_dst.p, bb.data(), bb.size() * sizeof(float), cudaMemcpyHostToDevice));
}

void Engine::load_weights(const ptts_tensor_desc* w, int nw) {
  for (int i = 0; i < nw; ++i) {
    HostTensor t;
    t.d = &w[i];
    PTTS_REQUIRE(w[i].ndim >= 1 && w[i].ndim <= 4 && w[i].data && w[i].name, PTTS_ERR_INVALID, "bad tensor descriptor %d", i);
    t.shape.assign(w[i].shape, w[i].shape + w[i].ndim);
    const size_t n = t.numel();
    t.f32.resize(n);
    if (w[i].dtype == PTTS_F32) std::memcpy(t.f32.data(), w[i].data, n * 4);
    else if (w[i].dtype == PTTS_BF16) { const uint16_t* s = (const uint16_t*)w[i].data; for (size_t j = 0; j < n; ++j) t.f32[j] = bf16_to_f32(s[j]); }
    else if (w[i].dtype == PTTS_F16) { const __half* s = (const __half*)w[i].data; for (size_t j = 0; j < n; ++j) t.f32[j] = __half2float(s[j]); }
    else PTTS_REQUIRE(false, PTTS_ERR_INVALID, "tensor '%s': unknown dtype", w[i].name);
    if (cfg.weight_mode == PTTS_W_INT8) {
      // reference policy (quantize.rs:27-41,117-154): per-tensor symmetric, skip < 1024 elements and names containing
      // embed / lut / out_proj / eos_head; Candle's round() is half away from zero
      const std::string nm(w[i].name);
      bool skip = n < 1024;
      for (const char* pat : {"embed", "lut", "out_proj", "eos_head"}) skip = skip || nm.find(pat) != std::string::npos;
      if (!skip) {
        float amax = 0.f;
        for (float v : t.f32) amax = std::max(amax, std::fabs(v));
        const float scale = amax > 0.f ? amax / 127.f : 1.f;
        for (float& v : t.f32) {
          const float r = v / scale;
          float q = std::copysign(std::floor(std::fabs(r) + 0.5f), r);
          q = std::min(127.f, std::max(-127.f, q));
          v = q * scale;
        }
        t.qscale = scale;
      }
    }
    host.emplace(w[i].name, std::move(t));
  }
  linear(w_input, "flow_lm.input_linear.weight", D_MODEL, LDIM, 64);
  for (int l = 0; l < N_LAYERS; ++l) {
    const std::string p = "flow_lm.transformer.layers." + std::to_string(l) + ".";
    linear(w_inproj[l], p + "self_attn.in_proj.weight", 3 * D_MODEL, D_MODEL);
    linear(w_outproj[l], p + "self_attn.out_proj.weight", D_MODEL, D_MODEL);
    linear(w_lin1[l], p + "linear1.weight", D_FFN, D_MODEL);
    linear(w_lin2[l], p + "linear2.weight", D_MODEL, D_FFN);
    vec(ln1_w[l], p + "norm1.weight", D_MODEL); vec(ln1_b[l], p + "norm1.bias", D_MODEL);
    vec(ln2_w[l], p + "norm2.weight", D_MODEL); vec(ln2_b[l], p + "norm2.bias", D_MODEL);
  }
  vec(outnorm_w, "flow_lm.out_norm.weight", D_MODEL); vec(outnorm_b, "flow_lm.out_norm.bias", D_MODEL);
  { const HostTensor& t = T("flow_lm.out_eos.weight", {1, D_MODEL}); eos_w.alloc(D_MODEL);
    PTTS_CUDA(cudaMemcpy(eos_w.p, t.f32.data(), D_MODEL * 4, cudaMemcpyHostToDevice)); }
  vec(eos_b, "flow_lm.out_eos.bias", 1);
  vec(bos, "flow_lm.bos_emb", LDIM); vec(emb_std, "flow_lm.emb_std", LDIM); vec(emb_mean, "flow_lm.emb_mean", LDIM);
  { const HostTensor& t = T("flow_lm.conditioner.embed.weight", {N_BINS + 1, D_MODEL}); lut.alloc(t.f32.size());
    PTTS_CUDA(cudaMemcpy(lut.p, t.f32.data(), t.f32.size() * 4, cudaMemcpyHostToDevice)); }
  const std::string f = "flow_lm.flow_net.";
  linear(w_cond, f + "cond_embed.weight", FLOW_DIM, D_MODEL); vec(b_cond, f + "cond_embed.bias", FLOW_DIM);
  linear(w_finproj, f + "input_proj.weight", FLOW_DIM, LDIM, 64); vec(b_finproj, f + "input_proj.bias", FLOW_DIM);
  {  // all adaLN modulation Linears share the operand silu(c + te): one [10240, 512] GEMM per LSD step
    std::vector<float> wa((size_t)MOD_LD * FLOW_DIM), ba(MOD_LD), wsc(MOD_LD, 0.f);
    for (int i = 0; i < FLOW_DEPTH; ++i) {
      const std::string q = f + "res_blocks." + std::to_string(i) + ".";
      const HostTensor& tw = T(q + "adaLN_modulation.1.weight", {3 * FLOW_DIM, FLOW_DIM});
      const HostTensor& tb = T(q + "adaLN_modulation.1.bias", {3 * FLOW_DIM});
      std::memcpy(&wa[(size_t)i * 3 * FLOW_DIM * FLOW_DIM], tw.f32.data(), tw.f32.size() * 4);
      std::memcpy(&ba[(size_t)i * 3 * FLOW_DIM], tb.f32.data(), tb.f32.size() * 4);
      std::fill(wsc.begin() + (size_t)i * 3 * FLOW_DIM, wsc.begin() + (size_t)(i + 1) * 3 * FLOW_DIM, tw.qscale);
      linear(w_mlp0[i], q + "mlp.0.weight", FLOW_DIM, FLOW_DIM); vec(b_mlp0[i], q + "mlp.0.bias", FLOW_DIM);
      linear(w_mlp2[i], q + "mlp.2.weight", FLOW_DIM, FLOW_DIM); vec(b_mlp2[i], q + "mlp.2.bias", FLOW_DIM);
      vec(inln_w[i], q + "in_ln.weight", FLOW_DIM); vec(inln_b[i], q + "in_ln.bias", FLOW_DIM);
    }
    const HostTensor& tw = T(f + "final_layer.adaLN_modulation.1.weight", {2 * FLOW_DIM, FLOW_DIM});
    const HostTensor& tb = T(f + "final_layer.adaLN_modulation.1.bias", {2 * FLOW_DIM});
    std::memcpy(&wa[(size_t)FLOW_DEPTH * 3 * FLOW_DIM * FLOW_DIM], tw.f32.data(), tw.f32.size() * 4);
    std::memcpy(&ba[(size_t)FLOW_DEPTH * 3 * FLOW_DIM], tb.f32.data(), tb.f32.size() * 4);
    std::fill(wsc.begin() + (size_t)FLOW_DEPTH * 3 * FLOW_DIM, wsc.end(), tw.qscale);
    upload_f16(w_ada, wa, MOD_LD, FLOW_DIM, wsc);
    b_ada.alloc(MOD_LD);
    PTTS_CUDA(cudaMemcpy(b_ada.p, ba.data(), ba.size() * 4, cudaMemcpyHostToDevice));
  }
  linear(w_final, f + "final_layer.linear.weight", LDIM, FLOW_DIM); vec(b_final, f + "final_layer.linear.bias", LDIM);
  {
    static_assert(FH_DIM == FLOW_DIM && FH_DEPTH == FLOW_DEPTH && FH_MOD_LD == MOD_LD, "flow_head.cuh constants");
    w_flowpack.alloc((size_t)FH_PACK_ROWS * FLOW_DIM);
    const size_t blk = (size_t)FLOW_DIM * FLOW_DIM;
    for (int i = 0; i < FLOW_DEPTH; ++i) {
      PTTS_CUDA(cudaMemcpy(w_flowpack.p + (2 * i) * blk, w_mlp0[i].w.p, blk * sizeof(__half), cudaMemcpyDeviceToDevice));
      PTTS_CUDA(cudaMemcpy(w_flowpack.p + (2 * i + 1) * blk, w_mlp2[i].w.p, blk * sizeof(__half), cudaMemcpyDeviceToDevice));
    }
    PTTS_CUDA(cudaMemcpy(w_flowpack.p + 2 * FLOW_DEPTH * blk, w_final.w.p, (size_t)w_final.Fpad * FLOW_DIM * sizeof(__half), cudaMemcpyDeviceToDevice));
  }
  auto upload_transposed_512x32 = [&](DevBuf<float>& dst, const HostTensor& t) {  // [512][32] -> [32][512]
    std::vector<float> tr(32 * 512);
    for (int c = 0; c < 512; ++c) for (int k = 0; k < 32; ++k) tr[k * 512 + c] = t.f32[c * 32 + k];
    dst.alloc(tr.size());
    PTTS_CUDA(cudaMemcpy(dst.p, tr.data(), tr.size() * 4, cudaMemcpyHostToDevice));
  };
  upload_transposed_512x32(wq, T("mimi.quantizer.output_proj.weight", {MIMI_DIM, LDIM, 1}));
  upload_transposed_512x32(wup, T("mimi.upsample.convtr.convtr.weight", {MIMI_DIM, 1, 32}));
  for (int l = 0; l < MIMI_LAYERS; ++l) {
    const std::string p = "mimi.decoder_transformer.transformer.layers." + std::to_string(l) + ".";
    linear(m_inproj[l], p + "self_attn.in_proj.weight", 3 * MIMI_DIM, MIMI_DIM, 0, false);
    linear(m_outproj[l], p + "self_attn.out_proj.weight", MIMI_DIM, MIMI_DIM, 0, false);
    linear(m_lin1[l], p + "linear1.weight", MIMI_FFN, MIMI_DIM, 0, false);
    linear(m_lin2[l], p + "linear2.weight", MIMI_DIM, MIMI_FFN, 0, false);
    vec(m_ln1_w[l], p + "norm1.weight", MIMI_DIM); vec(m_ln1_b[l], p + "norm1.bias", MIMI_DIM);
    vec(m_ln2_w[l], p + "norm2.weight", MIMI_DIM); vec(m_ln2_b[l], p + "norm2.bias", MIMI_DIM);
    vec(m_ls1[l], p + "layer_scale_1.scale", MIMI_DIM); vec(m_ls2[l], p + "layer_scale_2.scale", MIMI_DIM);
  }
  const std::string d = "mimi.decoder.model.";
  auto cw = [&](Weight16& dst, DevBuf<float>& bd, const std::string& n, int cout, int cin, int k, int cout_pad, int cin_pad) {
    conv_weight(dst, bd, T(d + n + ".conv.weight", {cout, cin, k}), T(d + n + ".conv.bias", {cout}), cout, cin, k, cout_pad, cin_pad);
  };
  auto ctw = [&](Weight16& dst, DevBuf<float>& bd, const std::string& n, int cin, int cout, int s) {
    convtr_weight(dst, bd, T(d + n + ".convtr.weight", {cin, cout, 2 * s}), T(d + n + ".convtr.bias", {cout}), cin, cout, s);
  };
  cw(s_conv0, sb_conv0, "0", 512, 512, 7, 512, 512);
  ctw(s_ct2, sb_ct2, "2", 512, 256, 6);
  cw(s_r3a, sb_r3a, "3.block.1", 128, 256, 3, 128, 256);
  cw(s_r3b, sb_r3b, "3.block.3", 256, 128, 1, 256, 128);
  ctw(s_ct5, sb_ct5, "5", 256, 128, 5);
  cw(s_r6a, sb_r6a, "6.block.1", 64, 128, 3, 64, 128);
  cw(s_r6b, sb_r6b, "6.block.3", 128, 64, 1, 128, 64);
  ctw(s_ct8, sb_ct8, "8", 128, 64, 4);
  // hidden width 32 is padded to 64 channels (zero weights, zero bias -> ELU(0) = 0) so K stays a multiple of 64
  cw(s_r9a, sb_r9a, "9.block.1", 32, 64, 3, 64, 64);
  cw(s_r9b, sb_r9b, "9.block.3", 64, 32, 1, 64, 64);
  {
    const HostTensor& tw = T(d + "11.conv.weight", {1, 64, 3});
    std::vector<float> g(192);
    for (int c = 0; c < 64; ++c) for (int j = 0; j < 3; ++j) g[j * 64 + c] = tw.f32[c * 3 + j];
    s_final_w.alloc(192);
    PTTS_CUDA(cudaMemcpy(s_final_w.p, g.data(), 192 * 4, cudaMemcpyHostToDevice));
    vec(s_final_b, d + "11.conv.bias", 1);
  }
  has_encoder = host.count("mimi.encoder.model.0.conv.weight") != 0;
  if (has_encoder) load_encoder_weights();
}

// Encoder side of Mimi (reference models/seanet.rs:148-247, models/mimi.rs:60-98, tts_model.rs:315-330): SEANetEncoder
// with ratios [4, 5, 6], the encoder transformer, ConvDownsample1d (stride 16) and speaker_proj_weight.
void Engine::load_encoder_weights() {
  const std::string e = "mimi.encoder.model.";
  {
    const HostTensor& tw = T(e + "0.conv.weight", {64, 1, 7});
    std::vector<float> g(7 * 64);
    for (int c = 0; c < 64; ++c) for (int j = 0; j < 7; ++j) g[j * 64 + c] = tw.f32[c * 7 + j];
    en_conv0_w.alloc(g.size());
    PTTS_CUDA(cudaMemcpy(en_conv0_w.p, g.data(), g.size() * 4, cudaMemcpyHostToDevice));
    vec(en_conv0_b, e + "0.conv.bias", 64);
  }
  const int res_idx[3] = {1, 4, 7}, down_idx[3] = {3, 6, 9}, ratio[3] = {4, 5, 6};
  for (int l = 0; l < 3; ++l) {
    const int C = 64 << l, H = C / 2, Hp = std::max(H, 64);
    const std::string r = e + std::to_string(res_idx[l]) + ".block.";
    // a 32-wide hidden layer is padded to 64 channels exactly like the decoder's last ResBlock
    conv_weight(en_r1[l], enb_r1[l], T(r + "1.conv.weight", {H, C, 3}), T(r + "1.conv.bias", {H}), H, C, 3, Hp, C);
    conv_weight(en_r3[l], enb_r3[l], T(r + "3.conv.weight", {C, H, 1}), T(r + "3.conv.bias", {C}), C, H, 1, C, Hp);
    // strided conv, k = 2 * ratio: [cout][j][cin] flattened is at once the two-tap weight over rows of `ratio` frames
    const std::string dn = e + std::to_string(down_idx[l]);
    conv_weight(en_down[l], enb_down[l], T(dn + ".conv.weight", {2 * C, C, 2 * ratio[l]}), T(dn + ".conv.bias", {2 * C}), 2 * C, C,
                2 * ratio[l], 2 * C, C);
  }
  conv_weight(en_c11, enb_c11, T(e + "11.conv.weight", {512, 512, 3}), T(e + "11.conv.bias", {512}), 512, 512, 3, 512, 512);
  for (int l = 0; l < MIMI_LAYERS; ++l) {
    const std::string p = "mimi.encoder_transformer.transformer.layers." + std::to_string(l) + ".";
    linear(et_inproj[l], p + "self_attn.in_proj.weight", 3 * MIMI_DIM, MIMI_DIM, 0, false);
    linear(et_outproj[l], p + "self_attn.out_proj.weight", MIMI_DIM, MIMI_DIM, 0, false);
    linear(et_lin1[l], p + "linear1.weight", MIMI_FFN, MIMI_DIM, 0, false);
    linear(et_lin2[l], p + "linear2.weight", MIMI_DIM, MIMI_FFN, 0, false);
    vec(et_ln1_w[l], p + "norm1.weight", MIMI_DIM); vec(et_ln1_b[l], p + "norm1.bias", MIMI_DIM);
    vec(et_ln2_w[l], p + "norm2.weight", MIMI_DIM); vec(et_ln2_b[l], p + "norm2.bias", MIMI_DIM);
    vec(et_ls1[l], p + "layer_scale_1.scale", MIMI_DIM); vec(et_ls2[l], p + "layer_scale_2.scale", MIMI_DIM);
  }
  {
    const HostTensor& tw = T("mimi.downsample.conv.conv.weight", {512, 512, 32});
    HostTensor nob{};
    nob.f32.assign(512, 0.f);
    DevBuf<float> unused;
    conv_weight(en_ds, unused, tw, nob, 512, 512, 32, 512, 512);
  }
  linear(w_spk, "flow_lm.speaker_proj_weight", D_MODEL, MIMI_DIM, 0, false);
}

// TimestepEmbedder + combine on the host in f32 (reference modules/mlp.rs:84-133,296-319); hoisted out of the
// frame loop exactly like tts_model.rs:994-1001.
void Engine::compute_time_embeddings(int steps) {
  std::vector<float> te((size_t)steps * FLOW_DIM, 0.f);
  for (int idx = 0; idx < 2; ++idx) {
    const std::string p = "flow_lm.flow_net.time_embed." + std::to_string(idx) + ".mlp.";
    const HostTensor& w0 = T(p + "0.weight", {FLOW_DIM, 256});
    const HostTensor& b0 = T(p + "0.bias", {FLOW_DIM});
    const HostTensor& w2 = T(p + "2.weight", {FLOW_DIM, FLOW_DIM});
    const HostTensor& b2 = T(p + "2.bias", {FLOW_DIM});
    const HostTensor& al = T(p + "3.alpha", {FLOW_DIM});
    for (int s = 0; s < steps; ++s) {
      const float tval = idx == 0 ? (float)((double)s / steps) : (float)((double)(s + 1) / steps);
      float emb[256], h1[FLOW_DIM], h2[FLOW_DIM];
      for (int i = 0; i < 128; ++i) {
        const float fr = std::exp((float)i * (-std::log(10000.f) / 128.f));
        emb[i] = std::cos(tval * fr);
        emb[128 + i] = std::sin(tval * fr);
      }
      for (int o = 0; o < FLOW_DIM; ++o) {
        float a = b0.f32[o];
        for (int i = 0; i < 256; ++i) a += w0.f32[(size_t)o * 256 + i] * emb[i];
        h1[o] = a / (1.f + std::exp(-a));
      }
      double mean = 0;
      for (int o = 0; o < FLOW_DIM; ++o) {
        float a = b2.f32[o];
        for (int i = 0; i < FLOW_DIM; ++i) a += w2.f32[(size_t)o * FLOW_DIM + i] * h1[i];
        h2[o] = a;
        mean += a;
      }
      mean /= FLOW_DIM;
      double var = 0;
      for (int o = 0; o < FLOW_DIM; ++o) var += (h2[o] - mean) * (h2[o] - mean);
      var /= (FLOW_DIM - 1);  // unbiased: the reference's "RMSNorm" is x * alpha * rsqrt(var(x) + eps), mlp.rs:18-26
      const float r = 1.f / std::sqrt((float)var + 1e-5f);
      for (int o = 0; o < FLOW_DIM; ++o) te[(size_t)s * FLOW_DIM + o] += 0.5f * h2[o] * al.f32[o] * r;
    }
  }
  // one allocation at the maximum size (lsd_steps <= MAX_LSD), overwritten in place: the captured step graphs have
  // time_emb.p + s * FLOW_DIM baked into their kernel arguments, so the buffer must never move
  PTTS_REQUIRE(steps >= 1 && steps <= MAX_LSD, PTTS_ERR_INVALID, "lsd_steps %d outside [1,%d]", steps, MAX_LSD);
  if (!time_emb.p) time_emb.alloc((size_t)MAX_LSD * FLOW_DIM);
  PTTS_CUDA(cudaMemcpy(time_emb.p, te.data(), te.size() * 4, cudaMemcpyHostToDevice));
  lsd_steps = steps;
}

// ------------------------------------------------------------------------------------------------ init
void Engine::green_split(int b_sms, int prio_a, int prio_b) {
  auto getDev = driver_fn<CUresult (*)(CUdevice*, int)>("cuDeviceGet");
  auto getRes = driver_fn<CUresult (*)(CUdevice, CUdevResource*, CUdevResourceType)>("cuDeviceGetDevResource");
  auto split = driver_fn<CUresult (*)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int)>("cuDevSmResourceSplitByCount");
  auto genDesc = driver_fn<CUresult (*)(CUdevResourceDesc*, CUdevResource*, unsigned int)>("cuDevResourceGenerateDesc");
  auto ctxCreate = driver_fn<CUresult (*)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int)>("cuGreenCtxCreate");
  auto streamCreate = driver_fn<CUresult (*)(CUstream*, CUgreenCtx, unsigned int, int)>("cuGreenCtxStreamCreate");
  CUdevice dev;
  PTTS_CU(getDev(&dev, cfg.device));
  CUdevResource all{}, part_b{}, part_a{};
  PTTS_CU(getRes(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
  unsigned int groups = 1;
  PTTS_CU(split(&part_b, &groups, &all, &part_a, 0, (unsigned int)b_sms));
  PTTS_REQUIRE(groups == 1 && part_a.sm.smCount >= 64, PTTS_ERR_INVALID, "SM partition: %u SMs for the codec leaves %u", part_b.sm.smCount, part_a.sm.smCount);
  CUdevResourceDesc da, db;
  PTTS_CU(genDesc(&da, &part_a, 1));
  PTTS_CU(genDesc(&db, &part_b, 1));
  CUgreenCtx ga, gb;
  PTTS_CU(ctxCreate(&ga, da, dev, CU_GREEN_CTX_DEFAULT_STREAM));
  PTTS_CU(ctxCreate(&gb, db, dev, CU_GREEN_CTX_DEFAULT_STREAM));
  CUstream sa, sb;
  PTTS_CU(streamCreate(&sa, ga, CU_STREAM_NON_BLOCKING, prio_a));
  PTTS_CU(streamCreate(&sb, gb, CU_STREAM_NON_BLOCKING, prio_b));
  stream = sa; stream_b = sb;
  green_ctx[0] = ga; green_ctx[1] = gb;
  sms_a = (int)part_a.sm.smCount; sms_b = (int)part_b.sm.smCount;
  persistent_ctas = std::min(persistent_ctas, sms_b);
  split_cta_cap_b = std::min(split_cta_cap_b, sms_b);
  if (std::getenv("PTTS_VERBOSE")) std::fprintf(stderr, "ptts: SM partition: language model %d SMs, codec %d SMs\n", sms_a, sms_b);
}

void Engine::init(const ptts_engine_cfg& c, const ptts_tensor_desc* w, int nw) {
  cfg = c;
  int ndev = 0;
  PTTS_CUDA(cudaGetDeviceCount(&ndev));
  PTTS_REQUIRE(c.device >= 0 && c.device < ndev, PTTS_ERR_CUDA, "CUDA device %d not present (%d devices)", c.device, ndev);
  PTTS_CUDA(cudaSetDevice(c.device));
  cudaDeviceProp prop;
  PTTS_CUDA(cudaGetDeviceProperties(&prop, c.device));
  PTTS_REQUIRE(prop.major == 10, PTTS_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", c.device,
               prop.major, prop.minor);
  PTTS_REQUIRE(c.weight_mode == PTTS_W_F16 || c.weight_mode == PTTS_W_INT8, PTTS_ERR_INVALID, "unknown weight_mode %d", c.weight_mode);
  NS = c.max_slots > 0 ? c.max_slots : 64;
  NB = c.max_batch > 0 ? std::min(c.max_batch, NS) : NS;
  KVCAP = c.kv_capacity > 0 ? c.kv_capacity : 1024;
  PR = 4096;
  {
    // the language-model stream carries the AR critical path: give its CTAs first pick of free SMs; the codec
    // stream fills what is left
    int lo = 0, hi = 0;
    PTTS_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    const bool prio = !(std::getenv("PTTS_PRIO") && std::atoi(std::getenv("PTTS_PRIO")) == 0);
    int green_b = 0;
    if (const char* v = std::getenv("PTTS_GREEN_B_SMS")) green_b = std::atoi(v);
    if (green_b > 0) {
      // SM partition (green contexts): the codec stream gets `green_b` SMs (rounded by the driver to whole co-scheduling
      // groups), the language-model stream the rest.  The two halves of a step then never share an SM, and the codec's
      // CTAs cannot scatter over the GPCs the language model's clusters need.
      green_split(green_b, prio ? hi : lo, lo);
    } else {
      PTTS_CUDA(cudaStreamCreateWithPriority(&stream, cudaStreamNonBlocking, prio ? hi : lo));
      PTTS_CUDA(cudaStreamCreateWithPriority(&stream_b, cudaStreamNonBlocking, lo));
    }
    if (const char* v = std::getenv("PTTS_B_SMS")) persistent_ctas = std::max(8, std::atoi(v));
    if (const char* v = std::getenv("PTTS_MAX_CTAS")) split_cta_cap = std::max(1, std::atoi(v));
    if (const char* v = std::getenv("PTTS_MAX_CTAS_B")) split_cta_cap_b = std::max(1, std::atoi(v));
    if (const char* v = std::getenv("PTTS_LIN1_CTAS")) lin1_ctas = std::max(1, std::atoi(v));
    if (const char* v = std::getenv("PTTS_LIN2_CTAS")) lin2_ctas = std::max(0, std::atoi(v));
    if (const char* v = std::getenv("PTTS_OUTPROJ_CTAS")) outproj_ctas = std::max(0, std::atoi(v));
    if (const char* v = std::getenv("PTTS_INPROJ_CTAS")) inproj_ctas = std::max(0, std::atoi(v));
    if (const char* v = std::getenv("PTTS_TRIG_A")) trig_a = std::atoi(v);
    if (const char* v = std::getenv("PTTS_TRIG_B")) trig_b = std::atoi(v);
    auto tune = [](const char* name, Tune& t) { if (const char* v = std::getenv(name)) std::sscanf(v, "%d,%d", &t.bn, &t.cap); };
    tune("PTTS_TUNE_CONV0", tune_conv0); tune("PTTS_TUNE_CT2", tune_ct2); tune("PTTS_TUNE_CT5", tune_ct5);
    tune("PTTS_TUNE_MLIN1", tune_mlin1); tune("PTTS_TUNE_MLIN2", tune_mlin2);
    tune("PTTS_TUNE_MINPROJ", tune_minproj); tune("PTTS_TUNE_MOUTPROJ", tune_moutproj);
  }
  {
    cudaDeviceProp prop{};
    PTTS_CUDA(cudaGetDeviceProperties(&prop, cfg.device));
    num_sms = prop.multiProcessorCount;
    if (const char* v = std::getenv("PTTS_GEMV_CTAS")) num_sms = std::max(1, std::atoi(v));
    if (const char* v = std::getenv("PTTS_GEMV_LN")) gemv_ln = std::atoi(v) != 0;
    if (const char* v = std::getenv("PTTS_PREFILL_TILE")) pf_tiled = std::atoi(v) != 0;
    if (const char* v = std::getenv("PTTS_FLOW_SMALL")) flow_small = std::atoi(v) != 0;
    PTTS_CUDA(cudaFuncSetAttribute(flow_head_small_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FS_SMEM));
    PTTS_CUDA(cudaFuncSetAttribute(flow_head_small_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, FS_SMEM));
    PTTS_CUDA(cudaFuncSetAttribute(flow_head_small_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, FS_SMEM));
    PTTS_CUDA(cudaFuncSetAttribute(flowlm_attn_prefill_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PF_SMEM_MAX));
  }
  ls = stream;
  PTTS_CUDA(cudaEventCreateWithFlags(&ev_a_done, cudaEventDisableTiming));
  PTTS_CUDA(cudaEventCreateWithFlags(&ev_front_done, cudaEventDisableTiming));
  PTTS_CUDA(cudaEventCreateWithFlags(&ev_b_done, cudaEventDisableTiming));
  for (auto& e : ev) PTTS_CUDA(cudaEventCreate(&e));
  PTTS_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  PTTS_CUDA(cudaFuncSetAttribute(gemm_tc_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  PTTS_CUDA(cudaFuncSetAttribute(mimi_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MATTN_SMEM));
  PTTS_CUDA(cudaFuncSetAttribute(flow_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FH_SMEM));
  fused_flow = !cfg.debug_gemm && cfg.reserved[5] == 0;
  if (std::getenv("PTTS_FH_TRACE")) fh_trace.alloc(FH_LAYERS * 8);
  {
    float inv[HD / 2];
    const float cst = -std::log(10000.f) * 2.f / (float)HD;  // f32 like modules/rope.rs:12-14
    for (int i = 0; i < HD / 2; ++i) inv[i] = std::exp((float)i * cst);
    PTTS_CUDA(cudaMemcpyToSymbol(c_inv_freq, inv, sizeof inv));
  }
  if (const char* v = std::getenv("PTTS_PDL")) use_pdl = std::atoi(v) != 0;
  if (const char* v = std::getenv("PTTS_DIAG_TIMES")) diag_times = std::atoi(v) != 0;
  if (diag_times) for (auto& e4 : ev_t) PTTS_CUDA(cudaEventCreate(&e4));
  if (const char* v = std::getenv("PTTS_DIAG_SKIP")) diag_skip = std::atoi(v);  // 1: time path A alone, 2: path B alone
  load_weights(w, nw);
  compute_time_embeddings(1);

  kv.alloc((size_t)NS * N_LAYERS * 2 * N_HEADS * KVCAP * HD);
  seqs.alloc(NS + 1); own_len.alloc(NS + 1); ctl.alloc(NS); feedback.alloc((size_t)NS * LDIM);
  up_partial.alloc((size_t)NS * 16 * 512);
  mimi_ring.alloc((size_t)NS * MIMI_LAYERS * 2 * MIMI_HEADS * MIMI_RING * HD);
  st_tr.alloc((size_t)NS * 6 * 512); st_a0.alloc((size_t)NS * 512); st_e2.alloc((size_t)NS * 2 * 256);
  st_a3.alloc((size_t)NS * 256); st_e5.alloc((size_t)NS * 2 * 128); st_a6.alloc((size_t)NS * 128);
  st_e8.alloc((size_t)NS * 2 * 64); st_a9.alloc((size_t)NS * 2 * 64);
  slots.resize(NS);

  row_seq.alloc(NB); x32.alloc((size_t)NB * D_MODEL); qkv32.alloc((size_t)NB * 3 * D_MODEL); eos_logit.alloc(NB);
  h16.alloc((size_t)NB * D_MODEL); attn16.alloc((size_t)NB * D_MODEL); ffn16.alloc((size_t)NB * D_FFN);
  h32dbg.alloc((size_t)NB * D_MODEL); quant_dbg.alloc((size_t)NB * 512);
  // the fused flow head works on whole 64-row chunks: its row-indexed buffers are padded so that rows past the
  // batch are readable and writable (never consumed), which keeps every access at a compile-time offset
  const size_t NBP = (size_t)round_up(NB, FH_ROWS);
  lat16.alloc((size_t)NB * 64); c32.alloc((size_t)NB * FLOW_DIM); mod32.alloc((size_t)MOD_STEPS * NBP * MOD_LD);
  fx32.alloc(NBP * FLOW_DIM); y16.alloc((size_t)MOD_STEPS * NB * FLOW_DIM); fh16.alloc(NBP * FLOW_DIM);
  fg16.alloc(NBP * FLOW_DIM); z32.alloc(NBP * LDIM); z16.alloc(NBP * 64);
  mimi_pos.alloc(NB);
  open_recs.alloc(NS);
  for (int i = 0; i < 2; ++i) {
    PTTS_CUDA(cudaMallocHost(&pin_open[i], (size_t)NS * sizeof(OpenRec)));
    PTTS_CUDA(cudaEventCreateWithFlags(&ev_open[i], cudaEventDisableTiming));
  }
  noise_pool.resize(NS);
  step_out.alloc(step_out_bytes());
  row_desc.alloc(NB);
  latent_out.p = reinterpret_cast<float*>(step_out.p);
  logit_out.p = latent_out.p + (size_t)NB * LDIM;
  finished_dev.p = reinterpret_cast<unsigned char*>(logit_out.p + NB);
  for (auto& x : ev_gfront) PTTS_CUDA(cudaEventCreateWithFlags(&x, cudaEventDisableTiming));
  {
    int frames = 1;
    if (const char* v = std::getenv("PTTS_CODEC_GROUP")) frames = std::atoi(v);
    if (const char* v = std::getenv("PTTS_QUEUE")) use_queue = std::atoi(v) != 0;
    if (const char* v = std::getenv("PTTS_SEANET_TAIL")) fused_tail = std::atoi(v) != 0;
    if (cfg.debug_gemm) fused_tail = false;   // the SIMT cross-check path keeps every layer a separate launch
    PTTS_CUDA(cudaFuncSetAttribute(seanet_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM));
    set_codec_group(frames == 2 || frames == 4 ? frames : 1);
  }

  px32.alloc((size_t)PR * D_MODEL); pqkv32.alloc((size_t)PR * 3 * D_MODEL); pqrot.alloc((size_t)PR * D_MODEL);
  ph16.alloc((size_t)PR * D_MODEL); pattn16.alloc((size_t)PR * D_MODEL); pffn16.alloc((size_t)PR * D_FFN);
  prow_seq.alloc(PR); prow_pos.alloc(PR); ptokens.alloc(PR); ptiles.alloc(PR);
  for (int i = 0; i < NT; ++i) {
    PTTS_CUDA(cudaMallocHost(&pin_pcm[i], (size_t)NB * FRAME * 4));
    PTTS_CUDA(cudaMallocHost(&pin_pcm16[i], (size_t)NB * FRAME * 2));
    PTTS_CUDA(cudaMallocHost(&pin_lat[i], step_out_bytes()));  // same layout as step_out
    pin_logit[i] = pin_lat[i] + (size_t)NB * LDIM;
    pin_fin[i] = reinterpret_cast<unsigned char*>(pin_logit[i] + NB);
    PTTS_CUDA(cudaEventCreateWithFlags(&ev_flags[i], cudaEventDisableTiming));
    PTTS_CUDA(cudaEventCreateWithFlags(&ev_pcm[i], cudaEventDisableTiming));
  }
  lm_build();
  for (auto it = host.begin(); it != host.end();)  // only the time-embedding MLPs are needed again (set_lsd_steps)
    it = (it->first.find("time_embed") == std::string::npos) ? host.erase(it) : std::next(it);
  PTTS_CUDA(cudaDeviceSynchronize());
}

// ------------------------------------------------------------------------------------------------ GEMM dispatch
template <int C>
void Engine::ln(const float* x, int rows, const float* w, const float* b, float eps, const float* shift, const float* scale,
                int mod_ld, __half* out, int out_ld) {
  if (rows <= 0) return;
  ProfScope ps(*this, take_tag("layernorm"), (double)rows * C * (4 + 2 + (scale ? 8 : 0)), 0, "ln_rows_kernel");
  launch_k(use_pdl, ln_rows_kernel<C>, (rows + 3) / 4, 128, 0, ls, 1, x, rows, w, b, eps, shift, scale, mod_ld, out, out_ld);
}

void Engine::gemm(const ActView& a, int n_streams, int T, int taps, int R, int G, const Weight16& w, int F, GemmEpi epi,
                  bool allow_split) {
  const long long rows = (long long)n_streams * T;
  if (rows <= 0) return;
  PTTS_REQUIRE(a.C % 64 == 0 && w.K == taps * a.C, PTTS_ERR_INVALID, "gemm: K mismatch (C %d taps %d, weight K %d)", a.C, taps, w.K);
  PTTS_REQUIRE(F <= w.Fpad, PTTS_ERR_INVALID, "gemm: F %d beyond weight rows %d", F, w.Fpad);
  {
    const LnSpec pre = pre_ln;
    pre_ln.set = false;
    if (taps == 1 && n_streams == 1 && gemv(a.ptr, (int)rows, w, F, epi, pre)) return;
    if (pre.set) {     // not taken by the GEMV: the operand is materialised by the LayerNorm kernel
      const char* t = cur_tag;
      ln_pending(pre);
      cur_tag = t;
    }
  }
  GemmParams p{};
  p.F = F; p.K = w.K; p.taps = taps; p.cblocks = a.C / 64;
  p.epi = epi;
  p.epi.wscale = w.wscale.p;  // null outside int8 mode
  p.trace = gemm_trace;
  p.act = a.ptr; p.act_stream_stride = (long long)a.Tpad * a.C; p.act_ld = a.C; p.w = w.w.p;
  const int total_kb = p.taps * p.cblocks;
  const bool plain = (taps == 1 && n_streams == 1);
  const int force = cfg.reserved[0];  // test hook: 1 = never swap, 2 = always swap when legal
  // tile geometry of a GEMM over `rows_g` rows (n_streams x T_g, tiles of G_g streams x R_g rows): operand placement,
  // persistent or one CTA per tile, feature-tile width, grid
  struct Geo { bool swap, persistent; int R, G, bn, act_tiles; dim3 grid; };
  const int bn_ov = bn_override;
  auto geometry = [&](long long rows_g, int T_g, int R_g, int G_g) {
    Geo g{};
    g.swap = plain && rows_g <= 256 && force != 1;
    if (plain && !g.swap) { R_g = 128; G_g = 1; }
    g.R = R_g; g.G = G_g;
    if (g.swap) {
      g.bn = std::max(16, round_up((int)rows_g, 16));
      g.grid = dim3(w.Fpad / 128, 1, 1);
      return g;
    }
    PTTS_REQUIRE(R_g > 0 && G_g > 0 && R_g * G_g <= 128, PTTS_ERR_INVALID, "gemm: bad tile geometry R %d G %d", R_g, G_g);
    g.act_tiles = ((T_g + R_g - 1) / R_g) * ((n_streams + G_g - 1) / G_g);
    const int fcap = round_up(F, 16);
    static const bool b_all_persistent = std::getenv("PTTS_B_ALL_PERSISTENT") && std::atoi(std::getenv("PTTS_B_ALL_PERSISTENT"));  // bring-up
    if ((g.act_tiles >= 148 || (b_all_persistent && ls == stream_b && (long long)g.act_tiles * ((F + 127) / 128) > persistent_ctas)) &&
        cfg.reserved[3] != 1) {
      // enough activation tiles to give every SM several: persistent kernel, accumulator double-buffered in TMEM
      g.bn = std::min(128, fcap);
      g.persistent = true;
    } else {
      // largest tile that still leaves at least ~half the SMs busy: one wave of fat tiles beats two of thin ones
      int bn = std::min(256, fcap);
      while (bn > 64 && (long long)g.act_tiles * ((F + bn - 1) / bn) < 74) bn >>= 1;
      if (bn_ov > 0) bn = std::min(bn_ov, std::min(256, fcap));
      g.bn = std::min(round_up(bn, 16), fcap);
    }
    g.grid = dim3(g.act_tiles, (F + g.bn - 1) / g.bn, 1);
    return g;
  };
  // Split-K whenever the output tiles alone cannot fill the chip (decode batches: 64 rows x F features is only
  // F/128 tiles): a cluster of `splits` CTAs along z shares one output tile (gemm.cuh), at most 8 (portable size).
  // power-of-two cluster sizes only (they pack into a GPC), and the whole grid must fit one wave with slack for
  // cluster placement: 24 tiles x 4 (96 CTAs) beats 24 x 6 (144 CTAs, measured 8.5 vs 15 us)
  (void)allow_split;
  int cap_override = split_cap_override;
  split_cap_override = 0;
  bn_override = 0;
  auto split_of = [&](const Geo& g) {
    int sp = 1;
    const int tiles = g.grid.x * g.grid.y;
    if (!g.persistent && total_kb >= 4) {
      int cap = g.swap ? split_cta_cap : split_cta_cap_b;
      if (cap_override > 0) cap = cap_override;
      while (sp * 2 <= GEMM_MAX_SPLIT && tiles * sp * 2 <= cap && sp * 2 <= total_kb / 2) sp *= 2;
    }
    return sp;
  };
  Geo geo = geometry(rows, T, R, G);
  int splits = split_of(geo);
  if (gemm_ref_f > 1) {
    // Codec group of f frames: the K split decides the order in which an output element's partial sums are added, so it is
    // taken from the geometry the SAME call has with one frame per row -- the PCM of a group is then bit-identical to
    // that of f single-frame passes.  (A tile's own K loop is the same sequence of k-blocks whatever the tile shape.)
    const int f = gemm_ref_f;
    const bool whole = (T == R);   // tiles of whole streams: one frame per row has 1/f of the rows and f times the streams per tile
    const Geo ref = geometry(rows / f, T / f, whole ? R / f : R, whole ? std::min(G * f, 128 / std::max(1, R / f)) : G);
    splits = split_of(ref);
    if (splits > 1) geo.persistent = false;
    if (!geo.swap && !geo.persistent) geo.grid = dim3(geo.act_tiles, (F + geo.bn - 1) / geo.bn, 1);
  }
  const bool swap = geo.swap;
  bool persistent = geo.persistent;
  dim3 grid = geo.grid;
  if (swap) {
    p.swap = 1;
    p.BN = geo.bn;
    p.n_streams = 1; p.T = (int)rows; p.R = p.BN; p.G = 1;
  } else {
    p.swap = 0;
    R = geo.R; G = geo.G;
    p.n_streams = n_streams; p.T = T; p.R = R; p.G = G;
    p.n_act_tiles = geo.act_tiles;
    p.BN = geo.bn;
    if (persistent) grid.x = std::min(geo.act_tiles, std::max(1, persistent_ctas / (int)grid.y));
  }
  const LnSpec lnreq = next_ln;
  next_ln.set = false;
  if (cfg.reserved[2] > 0 && !persistent) splits = std::min(cfg.reserved[2], std::min(total_kb, GEMM_MAX_SPLIT));  // test hook
  p.kb_per_split = (total_kb + splits - 1) / splits;
  splits = (total_kb + p.kb_per_split - 1) / p.kb_per_split;
  p.epi_mask = epi_mask_of(p.epi);
  p.pdl_trigger = (ls == stream_b) ? trig_b : trig_a;
  grid.z = splits;
  static const int env_smem_kb = std::getenv("PTTS_GEMM_SMEM_KB") ? std::atoi(std::getenv("PTTS_GEMM_SMEM_KB")) : 0;
  const long long kSmemBudget = (cfg.reserved[4] > 0 ? cfg.reserved[4] : env_smem_kb > 0 ? env_smem_kb : 200) * 1024LL;  // one CTA per SM
  const int stage_bytes = GEMM_BM * GEMM_BK * 2 + p.BN * GEMM_BK * 2;
  const size_t tile_bytes = swap ? (size_t)p.BN * (GEMM_BM + 4) * 4 : (size_t)GEMM_BM * (p.BN + 4) * 4;
  size_t smem;
  if (persistent) {
    // stages | staging tile | barriers; the pipeline may run a whole tile ahead of the epilogue
    p.stages = std::max(2, std::min(8, (int)((kSmemBudget - (long long)tile_bytes) / stage_bytes)));
    p.tmem_cols = pow2_at_least(2 * p.BN);
    smem = (size_t)p.stages * stage_bytes + tile_bytes + 8 * (2 * p.stages + 4) + 16 + 1024;
  } else {
    p.stages = std::max(2, std::min(std::min(8, p.kb_per_split + 1), (int)(kSmemBudget / stage_bytes)));
    p.tmem_cols = pow2_at_least(p.BN);
    // the epilogue re-uses the stage buffers for its staged f32 tile
    while ((size_t)p.stages * stage_bytes < tile_bytes) ++p.stages;
    // int8 storage: the weight tile arrives as bytes and warps 2-9 expand it to the f16 operand in shared memory
    // (reserved[7] = 1: test hook, stream the f16 copy of the codes instead -- results must be bit-identical)
    if (swap && w.q8.p && !cfg.debug_gemm && cfg.reserved[7] == 0) p.epi.reserved |= GEMM_F_W_INT8;
    // stages | barriers | alignment slack; stays below the 196 KB shared-memory carve-out (a few KB more would select the
    // 228 KB configuration and halve the L1 of every SM the GEMM touches)
    smem = (size_t)p.stages * stage_bytes + 8 * (2 * p.stages + 1) + 16 + 1024;
    if (p.epi.reserved & GEMM_F_W_INT8) smem += 16 * p.stages;
    p.resident = (swap && taps == 1 && n_streams == 1 && a.cap == 1 && p.kb_per_split <= p.stages && !cfg.debug_gemm &&
                  cfg.reserved[6] == 0) ? 1 : 0;   // reserved[6] = 1: test hook, the staged pipeline instead
  }
  auto map_ok = [](const void* ptr, const RowMap& m) {
    return !ptr || ((reinterpret_cast<uintptr_t>(ptr) % 16 == 0) && m.ld % 4 == 0 && m.base % 4 == 0 && m.stream_stride % 4 == 0);
  };
  p.vec4 = (F % 4 == 0) && map_ok(epi.gate, epi.gate_map) && map_ok(epi.res, epi.res_map) && map_ok(epi.out32, epi.out32_map) &&
           map_ok(epi.out16, epi.out16_map) && map_ok(epi.bias, plain_map(4)) && map_ok(epi.fscale, plain_map(4));

  // algorithmic traffic: weights once, the distinct activation rows once, every epilogue tensor once
  const double act_rows = (double)n_streams * (T + taps - 1);
  const bool w_int8 = (p.epi.reserved & GEMM_F_W_INT8) != 0;
  double bytes = (double)F * w.K * (w_int8 ? 1 : 2) + act_rows * a.C * 2;
  bytes += (double)rows * F * ((epi.out32 ? 4 : 0) + (epi.out16 ? 2 : 0) + (epi.res ? 4 : 0) + (epi.gate ? 4 : 0));
  {
    ProfScope ps(*this, take_tag("gemm"), bytes, 2.0 * rows * F * w.K,
                 cfg.debug_gemm ? "gemm_simt_kernel" : (persistent ? "gemm_tc_persistent_kernel" : "gemm_tc_kernel"));
    if (cfg.debug_gemm) {
      const long long n = rows * F;
      launch_k(use_pdl, gemm_simt_kernel, (unsigned)((n + 255) / 256), 256, 0, ls, 1, p);
    } else {
      // resident decode GEMM: (64 k, rows, k-blocks) with the k-block as the slowest box dimension, see gemm.cuh
      const CUtensorMap& ma = p.resident ? tmaps.get(a.ptr, 64, a.Tpad, a.C / 64, a.C, 64, p.BN, p.kb_per_split)
                              : swap     ? tmaps.get(a.ptr, a.C, a.Tpad, a.cap, a.C, (long long)a.Tpad * a.C, p.BN, 1)
                                         : tmaps.get(a.ptr, a.C, a.Tpad, a.cap, a.C, (long long)a.Tpad * a.C, p.R, p.G);
      const CUtensorMap& mw = w_int8 ? tmaps.get_u8(w.q8.p, w.K, w.Fpad, 128)
                                       : tmaps.get(w.w.p, w.K, w.Fpad, 1, w.K, (long long)w.Fpad * w.K, swap ? 128 : p.BN, 1);
      if (persistent) launch_k(use_pdl, gemm_tc_persistent_kernel, grid, GEMM_THREADS, smem, ls, 1, ma, mw, p);
      else launch_k(use_pdl, gemm_tc_kernel, grid, GEMM_THREADS, smem, ls, (int)grid.z, ma, mw, p);
    }
    PTTS_CUDA(cudaGetLastError());
  }
  ln_pending(lnreq);
}

void Engine::ln_pending(const LnSpec& lnreq) {
  if (!lnreq.set) return;
  tag(lnreq.tag);
  if (lnreq.C == 1024) ln<1024>(lnreq.x, lnreq.rows, lnreq.w, lnreq.b, lnreq.eps, lnreq.shift, lnreq.scale, lnreq.mod_ld, lnreq.out, lnreq.out_ld);
  else ln<512>(lnreq.x, lnreq.rows, lnreq.w, lnreq.b, lnreq.eps, lnreq.shift, lnreq.scale, lnreq.mod_ld, lnreq.out, lnreq.out_ld);
}

// Small-batch Linear (gemv.cuh): 1-4 rows, K a multiple of 256 (512 for one-byte codes).  Returns false when the shape or a
// test switch (ptts_engine_cfg.reserved[0] != 0 forces an operand placement of the tensor-core path, reserved[1] = 1 or
// PTTS_GEMV=0 turns the family off) leaves the call to the tensor-core GEMM.
template <int ROWS, bool INT8>
static void launch_gemv(bool pdl, cudaStream_t st, int kch, int grid, size_t smem, const GemvParams& q) {
  switch (kch) {
    case 1:  if constexpr (INT8) launch_k(pdl, gemv_rows_kernel<ROWS, 1, true>, grid, GEMV_THREADS, smem, st, 1, q); break;
    case 2:  launch_k(pdl, gemv_rows_kernel<ROWS, 2, INT8>, grid, GEMV_THREADS, smem, st, 1, q); break;
    case 4:  if constexpr (!INT8) launch_k(pdl, gemv_rows_kernel<ROWS, 4, false>, grid, GEMV_THREADS, smem, st, 1, q); break;
    case 8:  if constexpr (INT8) launch_k(pdl, gemv_rows_kernel<ROWS, 8, true>, grid, GEMV_THREADS, smem, st, 1, q); break;
    case 16: if constexpr (!INT8) launch_k(pdl, gemv_rows_kernel<ROWS, 16, false>, grid, GEMV_THREADS, smem, st, 1, q); break;
    default: break;
  }
}

static std::atomic<long long> g_gemv_launches{0};
static bool gemv_shape_ok(int K, bool int8) {
  const int per_chunk = int8 ? 512 : 256;   // K covered by one 16-byte chunk per lane
  if (K % per_chunk) return false;
  const int kch = K / per_chunk;
  return int8 ? (kch == 1 || kch == 2 || kch == 8) : (kch == 2 || kch == 4 || kch == 16);
}
static bool gemv_env_off() {
  static const bool off = std::getenv("PTTS_GEMV") && std::atoi(std::getenv("PTTS_GEMV")) == 0;
  return off;
}
bool Engine::ln_fusable(int rows, const Weight16& w) const {
  return gemv_ln && rows <= GEMV_MAX_ROWS && !cfg.debug_gemm && cfg.reserved[0] == 0 && cfg.reserved[1] == 0 && !gemv_env_off() &&
         w.K == 1024 && gemv_shape_ok(w.K, w.q8.p && cfg.reserved[7] == 0);
}

bool Engine::gemv(const __half* x, int rows, const Weight16& w, int F, const GemmEpi& epi, const LnSpec& pre) {
  if (rows > GEMV_MAX_ROWS || cfg.debug_gemm || cfg.reserved[0] != 0 || cfg.reserved[1] == 1 || gemv_env_off()) return false;
  const bool int8 = w.q8.p && cfg.reserved[7] == 0;
  if (!gemv_shape_ok(w.K, int8)) return false;
  const int kch = w.K / (int8 ? 512 : 256);
  if (pre.set && !(w.K == 1024 && pre.C == 1024 && pre.rows == rows && pre.out == x && !pre.scale && !pre.shift && pre.w)) return false;
  PTTS_REQUIRE(F <= w.Fpad, PTTS_ERR_INVALID, "gemv: F %d beyond weight rows %d", F, w.Fpad);
  const LnSpec lnreq = next_ln;
  next_ln.set = false;
  split_cap_override = 0;
  bn_override = 0;
  ++g_gemv_launches;
  GemvParams q{};
  q.w = int8 ? (const void*)w.q8.p : (const void*)w.w.p;
  q.x = x; q.F = F; q.K = w.K; q.rows = rows;
  q.epi = epi;
  q.epi.wscale = w.wscale.p;
  if (pre.set) { q.ln_x = pre.x; q.ln_w = pre.w; q.ln_b = pre.b; q.ln_eps = pre.eps; }
  const int R = rows == 1 ? 1 : rows == 2 ? 2 : 4;
  const size_t smem = (size_t)R * w.K * 2 + (pre.set ? 2 * 1024 * sizeof(float) : 0);
  double bytes = (double)F * w.K * (int8 ? 1 : 2) + (double)rows * w.K * (pre.set ? 4 : 2) + (pre.set ? 8.0 * w.K : 0.0);
  bytes += (double)rows * F * ((epi.out32 ? 4 : 0) + (epi.out16 ? 2 : 0) + (epi.res ? 4 : 0) + (epi.gate ? 4 : 0));
  {
    ProfScope ps(*this, take_tag("gemm"), bytes, 2.0 * rows * F * w.K, "gemv_rows_kernel");
    // one CTA per SM (two fit): the second slot is where the next launch's CTAs wait with their weights already requested
    const int grid = num_sms;
    if (int8) {
      if (R == 1) launch_gemv<1, true>(use_pdl, ls, kch, grid, smem, q);
      else if (R == 2) launch_gemv<2, true>(use_pdl, ls, kch, grid, smem, q);
      else launch_gemv<4, true>(use_pdl, ls, kch, grid, smem, q);
    } else {
      if (R == 1) launch_gemv<1, false>(use_pdl, ls, kch, grid, smem, q);
      else if (R == 2) launch_gemv<2, false>(use_pdl, ls, kch, grid, smem, q);
      else launch_gemv<4, false>(use_pdl, ls, kch, grid, smem, q);
    }
    PTTS_CUDA(cudaGetLastError());
  }
  ln_pending(lnreq);
  return true;
}

static GemmEpi epi_none() {
  GemmEpi e{};
  e.alpha = 1.f;
  return e;
}

// ------------------------------------------------------------------------------------------------ FlowLM transformer
// Reference models/transformer.rs:66-90 per layer.  Residual stream x stays f32 in HBM; out_proj and linear2
// accumulate straight into it (split-K, red.add) so no epilogue buffer exists.
// The LayerNorm in front of each GEMM rides behind the GEMM that produces its input (ln_after_next_gemm): the caller
// provides h = LN1_0(x) on entry.
void Engine::flowlm_layers(int rows, float* x, __half* h, float* qkv, __half* attn, __half* ffn, bool is_prefill,
                           float* qrot, const int* rseq, const int* rpos) {
  for (int l = 0; l < N_LAYERS; ++l) {
    GemmEpi e = epi_none();
    e.out32 = qkv; e.out32_map = plain_map(3 * D_MODEL);
    if (!is_prefill && inproj_ctas) split_cap_override = inproj_ctas;
    tag(is_prefill ? "prefill.in_proj" : "flowlm.in_proj"); gemm_rows(h, rows, D_MODEL, w_inproj[l], 3 * D_MODEL, e);
    if (is_prefill) {
      { ProfScope ps(*this, "prefill.rope_append", (double)rows * D_MODEL * (12 + 4 + 4), 0);
        launch_k(use_pdl, flowlm_rope_append_kernel, dim3(rows, (N_HEADS + 3) / 4), 128, 0, ls, 1, qkv, rseq, rpos, seqs.p, l, N_HEADS, qrot); }
      if (l == N_LAYERS - 1) break;  // the prompt pass keeps only KV (reference discards the output, tts_model.rs:958-964)
      { ProfScope ps(*this, "prefill.attn");
        const size_t pf_smem = (size_t)round_up(pf_kmax, 16) * 2 * PF_KP * sizeof(__half);
        if (pf_tiles > 0 && pf_smem <= PF_SMEM_MAX)
          launch_k(use_pdl, flowlm_attn_prefill_mma_kernel, dim3(pf_tiles, N_HEADS), 128, pf_smem, ls, 1, qrot, (const int2*)ptiles.p, rseq, rpos,
                   seqs.p, l, N_HEADS, attn);
        else
          launch_k(use_pdl, flowlm_attn_prefill_kernel, dim3(rows, N_HEADS), ATTN_THREADS, 0, ls, 1, qrot, rseq, rpos, seqs.p, l, N_HEADS, attn); }
    } else {
      { ProfScope ps(*this, "flowlm.attn_decode", step_kv_bytes + (double)rows * D_MODEL * (12 + 4 + 2), 0, "flowlm_attn_decode_kernel");
        launch_k(use_pdl, flowlm_attn_decode_kernel, dim3(rows, N_HEADS), ATTN_THREADS, 0, ls, 1, qkv, (const SeqDesc*)row_desc.p, l, N_HEADS, attn); }
    }
    // x += attn W_o^T, accumulated straight into the f32 residual stream by the (cluster split-K) epilogue; h = LN2(x)
    e = epi_none();
    e.out32 = x; e.out32_map = plain_map(D_MODEL); e.res = x; e.res_map = plain_map(D_MODEL);
    // 1-4 rows: LN2 / the next layer's LN1 run inside the prologue of the GEMV that consumes them (no launch, h never written)
    const bool fuse2 = ln_fusable(rows, w_lin1[l]);
    if (!fuse2) ln_after_next_gemm("flowlm.layernorm", x, rows, D_MODEL, ln2_w[l].p, ln2_b[l].p, 1e-5f, nullptr, nullptr, 0, h, D_MODEL);
    if (!is_prefill && outproj_ctas) split_cap_override = outproj_ctas;
    tag(is_prefill ? "prefill.out_proj" : "flowlm.out_proj"); gemm_rows(attn, rows, D_MODEL, w_outproj[l], D_MODEL, e, true);
    e = epi_none();
    e.act = ACT_GELU; e.out16 = ffn; e.out16_map = plain_map(D_FFN);
    if (!is_prefill) split_cap_override = lin1_ctas;
    if (fuse2) pre_ln = LnSpec{true, x, rows, D_MODEL, ln2_w[l].p, ln2_b[l].p, 1e-5f, nullptr, nullptr, 0, h, D_MODEL, "flowlm.layernorm"};
    tag(is_prefill ? "prefill.linear1" : "flowlm.linear1"); gemm_rows(h, rows, D_MODEL, w_lin1[l], D_FFN, e);
    e = epi_none();
    e.out32 = x; e.out32_map = plain_map(D_MODEL); e.res = x; e.res_map = plain_map(D_MODEL);
    const bool fuse1 = l + 1 < N_LAYERS && ln_fusable(rows, w_inproj[l + 1]);
    if (l + 1 < N_LAYERS && !fuse1)
      ln_after_next_gemm("flowlm.layernorm", x, rows, D_MODEL, ln1_w[l + 1].p, ln1_b[l + 1].p, 1e-5f, nullptr, nullptr, 0, h, D_MODEL);
    if (!is_prefill && lin2_ctas) split_cap_override = lin2_ctas;
    tag(is_prefill ? "prefill.linear2" : "flowlm.linear2"); gemm_rows(ffn, rows, D_FFN, w_lin2[l], D_MODEL, e, true);
    if (fuse1) pre_ln = LnSpec{true, x, rows, D_MODEL, ln1_w[l + 1].p, ln1_b[l + 1].p, 1e-5f, nullptr, nullptr, 0, h, D_MODEL, "flowlm.layernorm"};
  }
}


// ------------------------------------------------------------------------------------------------ one decode step
void Engine::upload_rows(const int* slot_ids, int n) {
  step_kv_bytes = 0;
  for (int i = 0; i < n; ++i) {
    const SlotHost& sh = slots[slot_ids[i]];
    step_kv_bytes += (double)((sh.voice ? sh.voice->len : 0) + sh.own_len + 1) * N_HEADS * HD * 2 * 2;
  }
  if ((int)row_seq_host.size() == n && std::equal(slot_ids, slot_ids + n, row_seq_host.begin())) return;
  flush_codec();   // frames still queued for the codec belong to the old batch map
  PTTS_CUDA(cudaStreamWaitEvent(stream, ev_b_done, 0));  // the codec stream may still be reading the previous batch map
  row_seq_host.assign(slot_ids, slot_ids + n);
  PTTS_CUDA(cudaMemcpyAsync(row_seq.p, row_seq_host.data(), n * sizeof(int), cudaMemcpyHostToDevice, stream));
}

static RowMap stream_map(int T, int ld, long long stream_stride, long long base) { return RowMap{T, ld, stream_stride, base}; }

// One decode step is three launch sequences:
//   A      step_begin, FlowLM transformer step, out_norm + EOS, LSD flow head, step_end      (the AR critical path)
//   front  latent de-norm + quantizer + upsample; snapshots the frame position for Mimi's attention
//   B      Mimi decoder transformer, SEANet decoder                                          (feeds nothing back)
// Frame n+1's A depends only on frame n's A, so run_step() puts A on one stream and front+B on another: the codec
// of frame n overlaps the language model of frame n+1.
// input_proj + six AdaLN residual blocks + final layer of one LSD step as a single cluster kernel (flow_head.cuh)
// All `lsd_steps` Euler steps of the flow head in one launch; step s reads its modulation rows at mod + s * mod_step_stride.
void Engine::flow_head_fused(int n, const float* mod, long long mod_step_stride, int steps) {
  FlowHeadParams fp{};
  fp.b_in = b_finproj.p; fp.b_final = b_final.p;
  fp.ws_in = w_finproj.wscale.p; fp.ws_final = w_final.wscale.p;
  for (int i = 0; i < FLOW_DEPTH; ++i) {
    fp.b0[i] = b_mlp0[i].p; fp.b2[i] = b_mlp2[i].p;
    fp.ln_w[i] = inln_w[i].p; fp.ln_b[i] = inln_b[i].p;
    fp.ws0[i] = w_mlp0[i].wscale.p; fp.ws2[i] = w_mlp2[i].wscale.p;
  }
  fp.mod = mod; fp.mod_step_stride = mod_step_stride; fp.z32 = z32.p; fp.z16 = z16.p; fp.x_dbg = fx32.p;
  fp.n = n; fp.steps = steps; fp.alpha = 1.f / (float)lsd_steps;
  fp.trace = fh_trace.p;
  if (n <= FS_MAX_ROWS && flow_small && cfg.reserved[1] == 0) {
    FlowSmallParams q{};
    q.fp = fp; q.fp.trace = nullptr;
    q.w_in = w_finproj.w.p; q.w_pack = w_flowpack.p;
    const double bytes = (double)steps * ((double)FH_PACK_ROWS * FLOW_DIM * 2 + 512.0 * 64 * 2 + (double)n * MOD_LD * 4.0) + (double)n * (64 * 2 + 32 * 8);
    const double flops = 2.0 * steps * n * (12.0 * FLOW_DIM * FLOW_DIM + 64.0 * FLOW_DIM + 32.0 * FLOW_DIM);
    ProfScope ps(*this, "flow.head_small", bytes, flops, "flow_head_small_kernel");
    if (n == 1) launch_k(use_pdl, flow_head_small_kernel<1>, dim3(1, 1, FS_CL), FS_THREADS, FS_SMEM, ls, FS_CL, q);
    else if (n == 2) launch_k(use_pdl, flow_head_small_kernel<2>, dim3(1, 1, FS_CL), FS_THREADS, FS_SMEM, ls, FS_CL, q);
    else launch_k(use_pdl, flow_head_small_kernel<4>, dim3(1, 1, FS_CL), FS_THREADS, FS_SMEM, ls, FS_CL, q);
    PTTS_CUDA(cudaGetLastError());
    return;
  }
  const CUtensorMap& m_win = tmaps.get(w_finproj.w.p, 64, w_finproj.Fpad, 1, 64, (long long)w_finproj.Fpad * 64, 128, 1);
  // weights: (64 k, rows, k-blocks); one box = four k-block tiles [k-block][128 features][64]
  const CUtensorMap& m_wp = tmaps.get(w_flowpack.p, 64, FH_PACK_ROWS, 8, FLOW_DIM, 64, 128, 4);
  const int chunks = (n + FH_ROWS - 1) / FH_ROWS;
  // weights once per cluster and step, modulation rows once per step, z in and out
  const double bytes = (double)steps * ((double)chunks * ((double)FH_PACK_ROWS * FLOW_DIM * 2 + 512.0 * 64 * 2) + (double)n * MOD_LD * 4.0) + (double)n * (64 * 2 + 32 * 8);
  const double flops = 2.0 * steps * n * (12.0 * FLOW_DIM * FLOW_DIM + 64.0 * FLOW_DIM + 32.0 * FLOW_DIM);
  ProfScope ps(*this, "flow.head_fused", bytes, flops, "flow_head_kernel");
  launch_k(use_pdl, flow_head_kernel, dim3(chunks, 1, FH_CLUSTER), FH_THREADS, FH_SMEM, ls, FH_CLUSTER, m_win, m_wp, fp);
  PTTS_CUDA(cudaGetLastError());
}

// The last SEANet ResBlock (k3 conv, k1 conv, skip) + the final 64 -> 1 conv of every stream over T samples (seanet_tail.cuh).
void Engine::seanet_tail(int n, int T) {
  SeanetTailParams tp{};
  tp.b_a = sb_r9a.p; tp.b_b = sb_r9b.p; tp.w_f = s_final_w.p; tp.b_f = s_final_b.p;
  tp.a9buf = a9.p; tp.pcm = pcm.p; tp.pcm16 = pcm16.p;
  tp.n = n; tp.T = T; tp.tiles_per_stream = (T + ST_STEP - 1) / ST_STEP; tp.n_tiles = n * tp.tiles_per_stream;
  const CUtensorMap& me = tmaps.get(e8.p, 64, 2 + T, NB, 64, (long long)(2 + T) * 64, ST_ROWS, 1);
  // the f32 skip viewed as rows of 128 halves, so that a 128-byte swizzled box is 32 floats wide
  const CUtensorMap& mx = tmaps.get(reinterpret_cast<const __half*>(x8.p), 128, T, NB, 128, (long long)T * 128, ST_ROWS, 1);
  const CUtensorMap& mwa = tmaps.get(s_r9a.w.p, s_r9a.K, s_r9a.Fpad, 1, s_r9a.K, (long long)s_r9a.Fpad * s_r9a.K, 64, 1);
  const CUtensorMap& mwb = tmaps.get(s_r9b.w.p, s_r9b.K, s_r9b.Fpad, 1, s_r9b.K, (long long)s_r9b.Fpad * s_r9b.K, 64, 1);
  const int grid = std::min(tp.n_tiles, persistent_ctas);
  // algorithmic traffic: e8 and x8 in, PCM (+ i16) out, weights once
  const double bytes = (double)n * ((2.0 + T) * 128 + (double)T * 256 + T * 6.0) + 64.0 * 256 * 2 + 64 * 4 * 2 + 192 * 4;
  const double flops = 2.0 * n * T * (64.0 * 192 + 64.0 * 64 + 192);
  ProfScope ps(*this, "seanet.tail_fused", bytes, flops, "seanet_tail_kernel");
  launch_k(use_pdl, seanet_tail_kernel, grid, ST_THREADS, ST_SMEM, ls, 1, me, mx, mwa, mwb, tp);
  PTTS_CUDA(cudaGetLastError());
}

void Engine::lm_step(int n) {
  LmStepParams& q = lm_params;
  q.n = n;
  q.lsd_steps = lsd_steps;
  q.time_emb = time_emb.p;
  q.trace = lm_trace.p;
  q.flags = lm_flags;
  q.stop_phase = std::getenv("PTTS_LM_STOP") ? std::atoi(std::getenv("PTTS_LM_STOP")) : 0;
  {
    const double kv = 6.0 * step_kv_bytes;
    const double bytes = lm_weight_bytes + (double)lsd_steps * MOD_LD * FLOW_DIM * 2 + kv + (double)n * (D_MODEL * 8 + lsd_steps * MOD_LD * 4.0);
    const double flops = 2.0 * n * ((double)N_LAYERS * 12 * D_MODEL * D_MODEL + (double)D_MODEL * FLOW_DIM + (double)lsd_steps * MOD_LD * FLOW_DIM) + kv;
    ProfScope ps(*this, "flowlm.step_kernel", bytes, flops, "flowlm_step_kernel");
    launch_k(use_pdl, flowlm_step_kernel, lm_ctas, LM_THREADS, LM_SMEM, ls, 1, q);
    PTTS_CUDA(cudaGetLastError());
  }
}

void Engine::lm_build() {
  // Selection (ptts_engine_cfg.reserved[8], else PTTS_LM_STEP_KERNEL): 0 = auto, 1 = never, 2 = always when legal.
  // Auto currently means off: at 64 streams the kernel finishes the language-model half in ~300 us against ~365 us of
  // per-layer launches, but it holds every SM it runs on for that whole time, and the codec half of the previous frame
  // (which the per-layer path overlaps on the SMs its small grids leave free) then has nowhere to run; measured
  // 8.5 k audio-s/s (96 CTAs) against 9.2 k.  DESIGN.md section 10 has the numbers.
  lm_enabled = false;
  int mode = cfg.reserved[8];
  if (mode == 0) if (const char* v = std::getenv("PTTS_LM_STEP_KERNEL")) mode = std::atoi(v) ? 2 : 1;
  if (cfg.weight_mode != PTTS_W_F16 || cfg.debug_gemm || mode != 2) return;
  cudaDeviceProp prop;
  PTTS_CUDA(cudaGetDeviceProperties(&prop, cfg.device));
  lm_ctas = prop.multiProcessorCount;
  if (const char* v = std::getenv("PTTS_LM_CTAS")) lm_ctas = std::max(16, std::min(prop.multiProcessorCount, std::atoi(v)));  // attention needs one CTA per head
  if (const char* v = std::getenv("PTTS_LM_FLAGS")) lm_flags = std::atoi(v);
  PTTS_CUDA(cudaFuncSetAttribute(flowlm_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LM_SMEM));
  // tile images of every weight the step kernel streams
  size_t total = 0;
  auto sz = [](const Weight16& w) { return (size_t)w.Fpad * w.K * 2; };
  for (int l = 0; l < N_LAYERS; ++l) total += sz(w_inproj[l]) + sz(w_outproj[l]) + sz(w_lin1[l]) + sz(w_lin2[l]);
  total += sz(w_cond) + sz(w_ada);
  lm_wt.alloc(total);
  size_t off = 0;
  auto tile = [&](const Weight16& w) {
    PTTS_REQUIRE(w.Fpad % 128 == 0 && w.K % 64 == 0, PTTS_ERR_INVALID, "step kernel: weight %d x %d is not tileable", w.Fpad, w.K);
    const long long chunks = (long long)w.Fpad * w.K / 8;
    lm_tile_weight_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, stream>>>(w.w.p, w.Fpad, w.K, lm_wt.p + off);
    const uint8_t* at = lm_wt.p + off;
    off += sz(w);
    return at;
  };
  LmStepParams& q = lm_params;
  q = LmStepParams{};
  for (int l = 0; l < N_LAYERS; ++l) {
    q.w_inproj[l] = tile(w_inproj[l]); q.w_outproj[l] = tile(w_outproj[l]);
    q.w_lin1[l] = tile(w_lin1[l]); q.w_lin2[l] = tile(w_lin2[l]);
    q.ln1_w[l] = ln1_w[l].p; q.ln1_b[l] = ln1_b[l].p; q.ln2_w[l] = ln2_w[l].p; q.ln2_b[l] = ln2_b[l].p;
  }
  q.w_cond = tile(w_cond); q.w_ada = tile(w_ada);
  PTTS_CUDA(cudaGetLastError());
  lm_weight_bytes = (double)total - (double)sz(w_ada);  // the adaLN weights are counted per LSD step
  q.w_input = w_input.w.p;
  q.outnorm_w = outnorm_w.p; q.outnorm_b = outnorm_b.p; q.eos_w = eos_w.p; q.eos_b = eos_b.p;
  q.b_cond = b_cond.p; q.b_ada = b_ada.p;
  // split every GEMM phase over the grid: units = feature tiles x K splits, at most LM_ACT_KB k-blocks per unit
  auto shape = [&](int F, int K, bool unsplit) {
    LmGemmShape sh{};
    sh.Mt = F / 128; sh.KB = K / 64;
    int S = unsplit ? 1 : std::max(1, std::min(sh.KB, lm_ctas / sh.Mt));
    sh.kbps = std::min(LM_ACT_KB, (sh.KB + S - 1) / S);
    sh.S = (sh.KB + sh.kbps - 1) / sh.kbps;
    return sh;
  };
  q.shape[LM_G_INPROJ] = shape(3 * D_MODEL, D_MODEL, false);
  q.shape[LM_G_OUTPROJ] = shape(D_MODEL, D_MODEL, false);
  q.shape[LM_G_LIN1] = shape(D_FFN, D_MODEL, false);
  q.shape[LM_G_LIN2] = shape(D_MODEL, D_FFN, false);
  q.shape[LM_G_COND] = shape(FLOW_DIM, D_MODEL, false);
  q.shape[LM_G_ADA] = shape(MOD_LD, FLOW_DIM, true);   // written straight to the modulation rows: no split
  PTTS_REQUIRE(q.shape[LM_G_ADA].S == 1 && w_ada.Fpad == MOD_LD && w_cond.Fpad == FLOW_DIM, PTTS_ERR_STATE, "step kernel: adaLN / cond shapes");
  size_t ws_floats = 0;
  const int Fs[LM_G_KINDS] = {3 * D_MODEL, D_MODEL, D_FFN, D_MODEL, FLOW_DIM, 0};
  for (int k = 0; k < LM_G_KINDS; ++k) ws_floats = std::max(ws_floats, (size_t)q.shape[k].S * LM_ROWS * Fs[k]);
  lm_ws.alloc(ws_floats);
  lm_hA.alloc((size_t)(D_MODEL / 64) * LM_ATILE); lm_attnA.alloc((size_t)(D_MODEL / 64) * LM_ATILE);
  lm_ffnA.alloc((size_t)(D_FFN / 64) * LM_ATILE); lm_yA.alloc((size_t)LM_MAX_LSD * (FLOW_DIM / 64) * LM_ATILE);
  lm_mod.alloc((size_t)LM_MAX_LSD * LM_ROWS * MOD_LD);
  lm_bar.alloc(2);
  if (std::getenv("PTTS_LM_TRACE")) lm_trace.alloc(2 * 64);
  q.row_seq = row_seq.p; q.ctl = ctl.p; q.feedback = feedback.p; q.seqs = seqs.p; q.own_len = own_len.p; q.row_desc = row_desc.p;
  q.x32 = x32.p; q.hA = lm_hA.p; q.attnA = lm_attnA.p; q.ffnA = lm_ffnA.p; q.yA = lm_yA.p; q.ws = lm_ws.p;
  q.z32 = z32.p; q.z16 = z16.p; q.eos_logit = eos_logit.p; q.c32 = c32.p; q.mod32 = lm_mod.p; q.h32dbg = h32dbg.p;
  q.bar = lm_bar.p;
  PTTS_CUDA(cudaStreamSynchronize(stream));
  lm_enabled = true;
}

void Engine::step_part_a(int n, bool marks) {
  if (lm_usable(n)) {
    lm_step(n);
    if (marks) PTTS_CUDA(cudaEventRecord(ev[1], ls));
    flow_head_fused(n, lm_mod.p, (long long)LM_ROWS * MOD_LD, lsd_steps);
    { ProfScope ps(*this, "step.end", (double)n * 32 * 12, 0);
      launch_k(use_pdl, step_end_kernel, n, 32, 0, ls, 1, row_seq.p, n, ctl.p, own_len.p, eos_logit.p, z32.p, feedback.p, finished_dev.p,
                                            latent_out.p, logit_out.p, cur_zq, cur_zqpos); }
    return;
  }
  // ---- FlowLM AR step (reference models/flow_lm.rs:98-145)
  GemmEpi e = epi_none();
  if (!cfg.debug_gemm && cfg.reserved[0] != 1) {
    // gather + input_linear + LayerNorm 1 of layer 0 in one launch
    ProfScope ps(*this, "flowlm.input_fused", (double)n * (32 * 12 + D_MODEL * 6) + D_MODEL * 64.0, 2.0 * n * D_MODEL * LDIM, "flowlm_input_kernel");
    launch_k(use_pdl, flowlm_input_kernel, n, 256, 0, ls, 1, row_seq.p, ctl.p, feedback.p, (const SeqDesc*)seqs.p, (const int*)own_len.p, row_desc.p,
             z32.p, z16.p, (const __half*)w_input.w.p, (const float*)w_input.wscale.p, (const float*)ln1_w[0].p, (const float*)ln1_b[0].p, x32.p, h16.p);
  } else {
    { ProfScope ps(*this, "step.begin", (double)n * 32 * 12, 0);
      launch_k(use_pdl, step_begin_kernel, n, 64, 0, ls, 1, row_seq.p, ctl.p, feedback.p, lat16.p, z32.p, z16.p, (const SeqDesc*)seqs.p,
               (const int*)own_len.p, row_desc.p); }
    e.out32 = x32.p; e.out32_map = plain_map(D_MODEL);
    ln_after_next_gemm("flowlm.layernorm", x32.p, n, D_MODEL, ln1_w[0].p, ln1_b[0].p, 1e-5f, nullptr, nullptr, 0, h16.p, D_MODEL);
    tag("flowlm.input_linear"); gemm_rows(lat16.p, n, 64, w_input, D_MODEL, e);
  }
  flowlm_layers(n, x32.p, h16.p, qkv32.p, attn16.p, ffn16.p, false, nullptr, row_seq.p, nullptr);
  { ProfScope ps(*this, "flowlm.out_norm_eos", (double)n * D_MODEL * (4 + 2 + 4), 0);
    launch_k(use_pdl, ln_eos_kernel, (n + 3) / 4, 128, 0, ls, 1, x32.p, n, outnorm_w.p, outnorm_b.p, eos_w.p, eos_b.p, h16.p, h32dbg.p,
                                                   eos_logit.p); }
  if (marks) PTTS_CUDA(cudaEventRecord(ev[1], ls));
  // ---- LSD flow head (reference flow_lm.rs:7-22,156-161; modules/mlp.rs:275,322-383)
  e = epi_none();
  e.bias = b_cond.p; e.out32 = c32.p; e.out32_map = plain_map(FLOW_DIM);
  tag("flow.cond_embed"); gemm_rows(h16.p, n, D_MODEL, w_cond, FLOW_DIM, e);
  if (fused_flow && mod_all_steps()) {
    // the modulations of ALL Euler steps by one Linear over lsd_steps x n rows (modules/mlp.rs:322-368), then one launch
    const int rows = lsd_steps * n;
    { ProfScope ps(*this, "flow.silu_add", (double)rows * FLOW_DIM * 6, 0);
      launch_k(use_pdl, silu_add_kernel, (rows * FLOW_DIM + 255) / 256, 256, 0, ls, 1, c32.p, time_emb.p, n, lsd_steps, FLOW_DIM, y16.p); }
    e = epi_none();
    e.bias = b_ada.p; e.out32 = mod32.p; e.out32_map = plain_map(MOD_LD);
    tag("flow.adaln"); gemm_rows(y16.p, rows, FLOW_DIM, w_ada, MOD_LD, e);
    flow_head_fused(n, mod32.p, (long long)n * MOD_LD, lsd_steps);
  } else
  for (int s = 0; s < lsd_steps; ++s) {
    { ProfScope ps(*this, "flow.silu_add", (double)n * FLOW_DIM * 6, 0);
      launch_k(use_pdl, silu_add_kernel, (n * FLOW_DIM + 255) / 256, 256, 0, ls, 1, c32.p, time_emb.p + (size_t)s * FLOW_DIM, n, 1, FLOW_DIM, y16.p); }
    e = epi_none();
    e.bias = b_ada.p; e.out32 = mod32.p; e.out32_map = plain_map(MOD_LD);
    tag("flow.adaln"); gemm_rows(y16.p, n, FLOW_DIM, w_ada, MOD_LD, e);
    if (fused_flow) {
      // more Euler steps than the modulation scratch holds: one launch per step, each a one-step integration of size alpha
      flow_head_fused(n, mod32.p, 0, 1);
      continue;
    }
    e = epi_none();
    e.bias = b_finproj.p; e.out32 = fx32.p; e.out32_map = plain_map(FLOW_DIM);
    // fh = LN(x) * (1 + scale) + shift of each block rides behind the GEMM that produces x (modules/mlp.rs:168-171)
    ln_after_next_gemm("flow.ln_modulate", fx32.p, n, FLOW_DIM, inln_w[0].p, inln_b[0].p, 1e-6f, mod32.p, mod32.p + FLOW_DIM, MOD_LD,
                       fh16.p, FLOW_DIM);
    tag("flow.input_proj"); gemm_rows(z16.p, n, 64, w_finproj, FLOW_DIM, e);
    for (int i = 0; i < FLOW_DEPTH; ++i) {
      const float* shift = mod32.p + (size_t)i * 3 * FLOW_DIM;
      e = epi_none();
      e.bias = b_mlp0[i].p; e.act = ACT_SILU; e.out16 = fg16.p; e.out16_map = plain_map(FLOW_DIM);
      tag("flow.mlp0"); gemm_rows(fh16.p, n, FLOW_DIM, w_mlp0[i], FLOW_DIM, e);
      e = epi_none();
      e.bias = b_mlp2[i].p; e.gate = shift + 2 * FLOW_DIM; e.gate_map = plain_map(MOD_LD);
      e.res = fx32.p; e.res_map = plain_map(FLOW_DIM); e.out32 = fx32.p; e.out32_map = plain_map(FLOW_DIM);
      const float* nshift = mod32.p + (size_t)(i + 1) * 3 * FLOW_DIM;  // next block, or the final layer (no affine)
      const bool fin = (i + 1 == FLOW_DEPTH);
      ln_after_next_gemm("flow.ln_modulate", fx32.p, n, FLOW_DIM, fin ? nullptr : inln_w[i + 1].p, fin ? nullptr : inln_b[i + 1].p,
                         1e-6f, nshift, nshift + FLOW_DIM, MOD_LD, fh16.p, FLOW_DIM);
      tag("flow.mlp2"); gemm_rows(fg16.p, n, FLOW_DIM, w_mlp2[i], FLOW_DIM, e);
    }
    e = epi_none();  // z += (W h + b) / S   (Euler step, flow_lm.rs:15-19)
    e.bias = b_final.p; e.alpha = 1.f / (float)lsd_steps; e.res = z32.p; e.res_map = plain_map(LDIM);
    e.out32 = z32.p; e.out32_map = plain_map(LDIM); e.out16 = z16.p; e.out16_map = plain_map(64);
    tag("flow.final"); gemm_rows(fh16.p, n, FLOW_DIM, w_final, LDIM, e);
  }
  // ---- EOS bookkeeping, AR feedback, cursors (reference tts_model.rs:1055-1069)
  { ProfScope ps(*this, "step.end", (double)n * 32 * 12, 0);
    launch_k(use_pdl, step_end_kernel, n, 32, 0, ls, 1, row_seq.p, n, ctl.p, own_len.p, eos_logit.p, z32.p, feedback.p, finished_dev.p,
                                          latent_out.p, logit_out.p, cur_zq, cur_zqpos); }
}

void Engine::step_front(int n, int f, int qbuf) {
  // ---- Mimi: de-norm + quantizer + upsample (reference mimi.rs:143-157).  f == 0: the single frame A just produced, read
  // from z32; f >= 1: the f queued frames of codec-group buffer `qbuf`
  const bool queued = f >= 1;
  const int nf = queued ? f : 1;
  const float* zsrc = queued ? zq.p + (size_t)qbuf * cg * NB * LDIM : z32.p;
  const int* zpos = queued ? zq_pos.p + (size_t)qbuf * cg * NB : nullptr;
  { ProfScope ps(*this, "mimi.frontend", (double)n * nf * 16 * 512 * 12, 0);
    launch_k(use_pdl, mimi_frontend_kernel, dim3(n, 4, nf), 128, 0, ls, 1, zsrc, (long long)NB * LDIM, zpos, row_seq.p, ctl.p, emb_std.p, emb_mean.p,
             wq.p, wup.p, up_partial.p, mx32.p, quant_dbg.p, mimi_pos.p); }
}

// Codec half for f consecutive frames of every row (f = 1: one frame).  Every buffer is [row][frames of the group][...], so
// a group is simply a longer chunk of each stream for the streaming convolutions (T = 16 f Mimi positions) -- the same
// arithmetic per output element as f single-frame passes.
void Engine::step_part_b(int n, int f, bool marks) {
  const int T1 = MIMI_T * f, T2 = 96 * f, T3 = 480 * f, T4 = FRAME * f;
  const int MR = n * T1;
  const int R1 = std::min(T1, 128), G1 = std::max(1, 128 / T1);   // rows per stream and streams per 128-row tile at 12.5 Hz x 16
  const ConvSegs& sg = segs_f(f);
  struct RefF { int& v; ~RefF() { v = 1; } } ref_guard{gemm_ref_f};
  gemm_ref_f = f;
  GemmEpi e = epi_none();
  tag("mimi.layernorm"); ln<MIMI_DIM>(mx32.p, MR, m_ln1_w[0].p, m_ln1_b[0].p, 1e-5f, nullptr, nullptr, 0, mh16.p, MIMI_DIM);
  for (int l = 0; l < MIMI_LAYERS; ++l) {
    e = epi_none();
    e.out32 = mqkv32.p; e.out32_map = plain_map(3 * MIMI_DIM);
    apply(tune_minproj);
    tag("mimi.in_proj"); gemm_rows(mh16.p, MR, MIMI_DIM, m_inproj[l], 3 * MIMI_DIM, e);
    { ProfScope ps(*this, "mimi.attn", (double)n * f * (16.0 * 1536 * 4 + 8.0 * 266 * 256 + 8.0 * 16 * 256 + 16.0 * 512 * 2), 0, "mimi_attn_kernel");
      launch_k(use_pdl, mimi_attn_kernel, dim3(n, MIMI_HEADS), MATTN_THREADS, MATTN_SMEM, ls, 1, mqkv32.p, row_seq.p, mimi_pos.p, mimi_ring.p, l, MIMI_LAYERS, mattn16.p, f); }
    e = epi_none();
    e.fscale = m_ls1[l].p; e.res = mx32.p; e.res_map = plain_map(MIMI_DIM); e.out32 = mx32.p; e.out32_map = plain_map(MIMI_DIM);
    ln_after_next_gemm("mimi.layernorm", mx32.p, MR, MIMI_DIM, m_ln2_w[l].p, m_ln2_b[l].p, 1e-5f, nullptr, nullptr, 0, mh16.p, MIMI_DIM);
    apply(tune_moutproj);
    tag("mimi.out_proj"); gemm_rows(mattn16.p, MR, MIMI_DIM, m_outproj[l], MIMI_DIM, e, true);
    e = epi_none();
    e.act = ACT_GELU; e.out16 = mffn16.p; e.out16_map = plain_map(MIMI_FFN);
    apply(tune_mlin1);
    tag("mimi.linear1"); gemm_rows(mh16.p, MR, MIMI_DIM, m_lin1[l], MIMI_FFN, e);
    e = epi_none();
    e.fscale = m_ls2[l].p; e.res = mx32.p; e.res_map = plain_map(MIMI_DIM); e.out32 = mx32.p; e.out32_map = plain_map(MIMI_DIM);
    const bool last = (l == MIMI_LAYERS - 1);
    if (last) {  // also emit the f16 operand of SEANet's first conv behind its 6 left-context rows
      e.out16 = tr16.p; e.out16_map = stream_map(T1, 512, (long long)(6 + T1) * 512, 6 * 512);
    }
    if (!last)
      ln_after_next_gemm("mimi.layernorm", mx32.p, MR, MIMI_DIM, m_ln1_w[l + 1].p, m_ln1_b[l + 1].p, 1e-5f, nullptr, nullptr, 0, mh16.p, MIMI_DIM);
    apply(tune_mlin2);
    tag("mimi.linear2"); gemm_rows(mffn16.p, MR, MIMI_FFN, m_lin2[l], MIMI_DIM, e, !last);
  }
  if (marks) PTTS_CUDA(cudaEventRecord(ev[3], ls));
  static const int diag_b_stop = std::getenv("PTTS_DIAG_B_STOP") ? std::atoi(std::getenv("PTTS_DIAG_B_STOP")) : 0;  // bring-up: cut the codec short
  if (diag_b_stop == 1) return;
  // ---- SEANet decoder (reference seanet.rs:309-402) as implicit GEMMs; ELU fused into the producer's epilogue
  { ProfScope ps(*this, "seanet.state_move", (double)n * 5824 * 4, 0);
    launch_k(use_pdl, conv_state_move_kernel, dim3(n, 8), 128, 0, ls, 1, sg, row_seq.p, 0); }
  e = epi_none(); e.bias = sb_conv0.p; e.out16 = a0.p; e.act16 = ACT_ELU; e.out16_map = stream_map(T1, 512, (long long)(1 + T1) * 512, 512);
  apply(tune_conv0);
  tag("seanet.conv0"); gemm(ActView{tr16.p, 512, 6 + T1, NB}, n, T1, 7, R1, G1, s_conv0, 512, e);
  e = epi_none(); e.bias = sb_ct2.p; e.out32 = x2.p; e.out32_map = stream_map(T1, 1536, (long long)T2 * 256, 0);
  e.out16 = e2.p; e.act16 = ACT_ELU; e.out16_map = stream_map(T1, 1536, (long long)(2 + T2) * 256, 2 * 256);
  apply(tune_ct2);
  tag("seanet.convtr2"); gemm(ActView{a0.p, 512, 1 + T1, NB}, n, T1, 2, R1, G1, s_ct2, 1536, e);
  e = epi_none(); e.bias = sb_r3a.p; e.out16 = h3.p; e.act16 = ACT_ELU; e.out16_map = plain_map(128);
  tag("seanet.res3a"); gemm(ActView{e2.p, 256, 2 + T2, NB}, n, T2, 3, 96, 1, s_r3a, 128, e);
  e = epi_none(); e.bias = sb_r3b.p; e.res = x2.p; e.res_map = plain_map(256);
  e.out16 = a3.p; e.act16 = ACT_ELU; e.out16_map = stream_map(T2, 256, (long long)(1 + T2) * 256, 256);
  // the k1 convs run per stream (T rows each) rather than as one flat [n*T] matrix: the epilogue's affine fast path
  // needs the output map's stream structure to match the GEMM's, and the flat form silently took the row-by-row path
  // with a division per row (res9b: 52 us in the step against 23 us for the same GEMM with plain maps)
  tag("seanet.res3b"); gemm(ActView{h3.p, 128, T2, NB}, n, T2, 1, 96, 1, s_r3b, 256, e);
  e = epi_none(); e.bias = sb_ct5.p; e.out32 = x5.p; e.out32_map = stream_map(T2, 640, (long long)T3 * 128, 0);
  e.out16 = e5.p; e.act16 = ACT_ELU; e.out16_map = stream_map(T2, 640, (long long)(2 + T3) * 128, 2 * 128);
  apply(tune_ct5);
  tag("seanet.convtr5"); gemm(ActView{a3.p, 256, 1 + T2, NB}, n, T2, 2, 96, 1, s_ct5, 640, e);
  e = epi_none(); e.bias = sb_r6a.p; e.out16 = h6.p; e.act16 = ACT_ELU; e.out16_map = plain_map(64);
  tag("seanet.res6a"); gemm(ActView{e5.p, 128, 2 + T3, NB}, n, T3, 3, 120, 1, s_r6a, 64, e);
  e = epi_none(); e.bias = sb_r6b.p; e.res = x5.p; e.res_map = plain_map(128);
  e.out16 = a6.p; e.act16 = ACT_ELU; e.out16_map = stream_map(T3, 128, (long long)(1 + T3) * 128, 128);
  tag("seanet.res6b"); gemm(ActView{h6.p, 64, T3, NB}, n, T3, 1, 120, 1, s_r6b, 128, e);
  if (diag_b_stop == 2) return;
  e = epi_none(); e.bias = sb_ct8.p; e.out32 = x8.p; e.out32_map = stream_map(T3, 256, (long long)T4 * 64, 0);
  e.out16 = e8.p; e.act16 = ACT_ELU; e.out16_map = stream_map(T3, 256, (long long)(2 + T4) * 64, 2 * 64);
  tag("seanet.convtr8"); gemm(ActView{a6.p, 128, 1 + T3, NB}, n, T3, 2, 120, 1, s_ct8, 256, e);
  if (fused_tail) {
    seanet_tail(n, T4);
  } else {
    e = epi_none(); e.bias = sb_r9a.p; e.out16 = h9.p; e.act16 = ACT_ELU; e.out16_map = plain_map(64);
    tag("seanet.res9a"); gemm(ActView{e8.p, 64, 2 + T4, NB}, n, T4, 3, 128, 1, s_r9a, 64, e);
    e = epi_none(); e.bias = sb_r9b.p; e.res = x8.p; e.res_map = plain_map(64);
    e.out16 = a9.p; e.act16 = ACT_ELU; e.out16_map = stream_map(T4, 64, (long long)(2 + T4) * 64, 128);
    tag("seanet.res9b"); gemm(ActView{h9.p, 64, T4, NB}, n, T4, 1, 128, 1, s_r9b, 64, e);
    { ProfScope ps(*this, "seanet.final_conv", (double)n * ((2.0 + T4) * 128 + T4 * 4), 2.0 * n * T4 * 192);
      launch_k(use_pdl, seanet_final_conv_kernel, dim3((T4 + 255) / 256, n), 256, 0, ls, 1, a9.p, s_final_w.p, s_final_b.p, n, T4, pcm.p, pcm16.p); }
  }
  { ProfScope ps(*this, "seanet.state_move", (double)n * 5824 * 4, 0);
    launch_k(use_pdl, conv_state_move_kernel, dim3(n, 8), 128, 0, ls, 1, sg, row_seq.p, 1); }
}

// Sequential form on one stream (stage timing, per-launch profiling).
void Engine::step_kernels(int n, float* stage_ms) {
  ls = stream;
  const bool marks = stage_ms != nullptr;
  if (marks) PTTS_CUDA(cudaEventRecord(ev[0], ls));
  cur_zq = nullptr; cur_zqpos = nullptr;   // the sequential form decodes the frame straight from z32
  step_part_a(n, marks);
  if (marks) PTTS_CUDA(cudaEventRecord(ev[2], ls));
  step_front(n, 0, 0);
  step_part_b(n, 1, marks);
  if (marks) PTTS_CUDA(cudaEventRecord(ev[4], ls));
  PTTS_CUDA(cudaGetLastError());
  if (stage_ms) {
    PTTS_CUDA(cudaStreamSynchronize(stream));
    for (int i = 0; i < 4; ++i) PTTS_CUDA(cudaEventElapsedTime(stage_ms + i, ev[i], ev[i + 1]));
    stage_ms[4] = 0.f;
    PTTS_CUDA(cudaEventElapsedTime(stage_ms + 5, ev[0], ev[4]));
    stage_ms[6] = stage_ms[7] = 0.f;
  }
}

cudaGraphExec_t Engine::capture(cudaStream_t st, int n, int part, int arg, long long* kernels) {
  const long long before = launches;
  cudaGraph_t g = nullptr;
  ls = st;
  PTTS_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  try {
    if (part == 0) {   // arg: queue entry (codec group) the frame's latent goes to, or -1
      cur_zq = arg >= 0 ? zq.p + (size_t)arg * NB * LDIM : nullptr;
      cur_zqpos = arg >= 0 ? zq_pos.p + (size_t)arg * NB : nullptr;
      step_part_a(n, false);
    } else {           // arg: frames per row
      step_part_b(n, arg, false);
    }
  } catch (...) {
    cudaStreamEndCapture(st, &g);
    if (g) cudaGraphDestroy(g);
    throw;
  }
  PTTS_CUDA(cudaStreamEndCapture(st, &g));
  *kernels = launches - before;
  launches = before;
  cudaGraphExec_t exec = nullptr;
  PTTS_CUDA(cudaGraphInstantiate(&exec, g, 0));
  PTTS_CUDA(cudaGraphDestroy(g));
  return exec;
}

Engine::PartGraph& Engine::graph_of(int part, int n, int arg) {
  const auto key = std::make_tuple(part, n, lsd_steps * (part == 0), arg);
  auto it = graphs.find(key);
  if (it == graphs.end()) {
    PartGraph pg{};
    pg.exec = capture(part == 0 ? stream : stream_b, n, part, arg, &pg.kernels);
    it = graphs.emplace(key, pg).first;
  }
  return it->second;
}

void Engine::drop_graphs() {
  for (auto& g : graphs) cudaGraphExecDestroy(g.second.exec);
  graphs.clear();
}

// Codec half of the frames waiting in the current group buffer (cg > 1): front end + Mimi transformer + SEANet over
// `pend` frames per row on the codec stream, then the PCM copies the group's tickets asked for.
void Engine::flush_codec() {
  if (pend == 0) return;
  const int n = pend_n, f = pend;
  PTTS_CUDA(cudaStreamWaitEvent(stream_b, ev_a_done, 0));   // recorded behind the group's last A
  ls = stream_b;
  if (diag_times) PTTS_CUDA(cudaEventRecord(ev_t[2], stream_b));
  if (diag_skip != 1) {
    step_front(n, f, gbuf);
    if (cfg.use_cuda_graph) {
      PartGraph& pg = graph_of(1, n, f);
      PTTS_CUDA(cudaGraphLaunch(pg.exec, stream_b));
      launches += pg.kernels;
    } else {
      step_part_b(n, f, false);
    }
  }
  if (diag_times) PTTS_CUDA(cudaEventRecord(ev_t[3], stream_b));
  // the queue of this buffer is consumed by the front end, but the codec scratch (mx32 ... pcm) is shared by both buffers:
  // the next group's codec is ordered behind this one by the stream itself
  PTTS_CUDA(cudaEventRecord(ev_gfront[gbuf], stream_b));
  for (int j = 0; j < f; ++j) {
    const long long id = pend_ticket[j];
    if (id < 0) continue;
    Ticket& t = tickets[id % NT];
    const int par = (int)(id % NT);
    if (t.want_i16)
      PTTS_CUDA(cudaMemcpy2DAsync(pin_pcm16[par], (size_t)FRAME * 2, pcm16.p + (size_t)j * FRAME, (size_t)f * FRAME * 2, (size_t)FRAME * 2, n,
                                  cudaMemcpyDeviceToHost, stream_b));
    else if (t.want_pcm)
      PTTS_CUDA(cudaMemcpy2DAsync(pin_pcm[par], (size_t)FRAME * 4, pcm.p + (size_t)j * FRAME, (size_t)f * FRAME * 4, (size_t)FRAME * 4, n,
                                  cudaMemcpyDeviceToHost, stream_b));
    PTTS_CUDA(cudaEventRecord(ev_pcm[par], stream_b));
    t.codec_pending = false;
  }
  PTTS_CUDA(cudaEventRecord(ev_b_done, stream_b));
  gbuf ^= 1;
  pend = 0;
  ls = stream;
  PTTS_CUDA(cudaGetLastError());
}

// Enqueue one step: A on `stream`, front + B on `stream_b`, each of A and B replayed as a CUDA graph (the step is
// ~100 dependent launches of a few microseconds; every pointer is a fixed engine buffer and the batch composition is
// data, so one graph per (part, rows, lsd_steps, queue entry | frames) serves every step of that shape).
// cg > 1 (codec group): A leaves the frame's latent in a queue and the codec half runs once per cg frames of the same
// batch composition -- the streaming convolutions and the windowed attention do not care whether a stream's next 16 cg
// positions arrive in one call or in cg, and B feeds nothing back into A -- so the codec's ~35 launches are paid once per
// group.  A change of composition, a PCM fetch, close / open and ptts_sync flush a partial group.
void Engine::run_step(int n, long long ticket) {
  if (profiling) {
    // sequential form on `stream`; the codec stream (PCM copy, ev_pcm) must still order behind it
    flush_codec();
    step_kernels(n, nullptr);
    PTTS_CUDA(cudaEventRecord(ev_a_done, stream));
    PTTS_CUDA(cudaEventRecord(ev_front_done, stream));
    PTTS_CUDA(cudaStreamWaitEvent(stream_b, ev_front_done, 0));
    PTTS_CUDA(cudaEventRecord(ev_b_done, stream_b));
    return;
  }
  if (queued()) {
    if (pend > 0 && (pend_n != n || !std::equal(row_seq_host.begin(), row_seq_host.begin() + n, pend_rows.begin()))) flush_codec();
    if (pend == 0) {
      // the queue buffer is free once the front end of the group that used it last has run
      PTTS_CUDA(cudaStreamWaitEvent(stream, ev_gfront[gbuf], 0));
      pend_n = n;
      pend_rows.assign(row_seq_host.begin(), row_seq_host.begin() + n);
    }
    const int entry = gbuf * cg + pend;
    if (diag_times && pend == 0) PTTS_CUDA(cudaEventRecord(ev_t[0], stream));
    if (diag_skip != 2) {
      if (cfg.use_cuda_graph) {
        PartGraph& pg = graph_of(0, n, entry);
        PTTS_CUDA(cudaGraphLaunch(pg.exec, stream));
        launches += pg.kernels;
      } else {
        ls = stream;
        cur_zq = zq.p + (size_t)entry * NB * LDIM; cur_zqpos = zq_pos.p + (size_t)entry * NB;
        step_part_a(n, false);
      }
    }
    if (diag_times) PTTS_CUDA(cudaEventRecord(ev_t[1], stream));
    PTTS_CUDA(cudaEventRecord(ev_a_done, stream));
    pend_ticket[pend] = ticket;
    ++pend;
    if (pend == cg) flush_codec();
    ls = stream;
    PTTS_CUDA(cudaGetLastError());
    return;
  }
  // A(n) may not overwrite z32 / advance the frame counters before front(n-1) has consumed them
  PTTS_CUDA(cudaStreamWaitEvent(stream, ev_front_done, 0));
  if (cfg.use_cuda_graph) {
    PartGraph& ga = graph_of(0, n, -1);
    PartGraph& gb = graph_of(1, n, 1);
    if (diag_times) PTTS_CUDA(cudaEventRecord(ev_t[0], stream));
    if (diag_skip != 2) PTTS_CUDA(cudaGraphLaunch(ga.exec, stream));
    if (diag_times) PTTS_CUDA(cudaEventRecord(ev_t[1], stream));
    PTTS_CUDA(cudaEventRecord(ev_a_done, stream));
    PTTS_CUDA(cudaStreamWaitEvent(stream_b, ev_a_done, 0));
    ls = stream_b;
    if (diag_times) PTTS_CUDA(cudaEventRecord(ev_t[2], stream_b));
    if (diag_skip != 1) step_front(n, 0, 0);
    PTTS_CUDA(cudaEventRecord(ev_front_done, stream_b));
    if (diag_skip == 4) {   // bring-up: a codec stand-in that only occupies CTAs (PTTS_DIAG_SPIN_CTAS x PTTS_DIAG_SPIN_US, PTTS_DIAG_SPIN_SMEM KB each)
      static const int sc = std::getenv("PTTS_DIAG_SPIN_CTAS") ? std::atoi(std::getenv("PTTS_DIAG_SPIN_CTAS")) : 1;
      static const int su = std::getenv("PTTS_DIAG_SPIN_US") ? std::atoi(std::getenv("PTTS_DIAG_SPIN_US")) : 300;
      static const int sk = std::getenv("PTTS_DIAG_SPIN_SMEM") ? std::atoi(std::getenv("PTTS_DIAG_SPIN_SMEM")) : 0;
      static bool once = false;
      if (!once) { once = true; PTTS_CUDA(cudaFuncSetAttribute(spin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); }
      launch_k(false, spin_kernel, sc, 32, (size_t)sk * 1024, stream_b, 1, (unsigned long long)su * 1000ULL);
    } else if (diag_skip != 1 && diag_skip != 3) PTTS_CUDA(cudaGraphLaunch(gb.exec, stream_b));
    if (diag_times) PTTS_CUDA(cudaEventRecord(ev_t[3], stream_b));
    launches += ga.kernels + gb.kernels;
  } else {
    ls = stream;
    cur_zq = nullptr; cur_zqpos = nullptr;
    step_part_a(n, false);
    PTTS_CUDA(cudaEventRecord(ev_a_done, stream));
    PTTS_CUDA(cudaStreamWaitEvent(stream_b, ev_a_done, 0));
    ls = stream_b;
    step_front(n, 0, 0);
    PTTS_CUDA(cudaEventRecord(ev_front_done, stream_b));
    step_part_b(n, 1, false);
  }
  PTTS_CUDA(cudaEventRecord(ev_b_done, stream_b));
  ls = stream;
  PTTS_CUDA(cudaGetLastError());
}

// Frames per codec pass (1, 2 or 4).  The codec scratch is sized for it, so a change re-allocates (device idle) and drops
// the captured graphs, which bake buffer addresses.
void Engine::set_codec_group(int frames) {
  PTTS_REQUIRE(frames == 1 || frames == 2 || frames == 4, PTTS_ERR_INVALID, "codec group of %d frames (supported: 1, 2, 4)", frames);
  flush_codec();
  sync_all();
  if (frames == cg && mx32.p) return;
  drop_graphs();
  cg = frames;
  const size_t f = (size_t)cg, MR = (size_t)NB * MIMI_T * f;
  mx32.alloc(MR * MIMI_DIM); mqkv32.alloc(MR * 3 * MIMI_DIM); mh16.alloc(MR * MIMI_DIM); mattn16.alloc(MR * MIMI_DIM);
  mffn16.alloc(MR * MIMI_FFN);
  tr16.alloc((size_t)NB * (6 + 16 * f) * 512); a0.alloc((size_t)NB * (1 + 16 * f) * 512);
  x2.alloc((size_t)NB * 96 * f * 256); e2.alloc((size_t)NB * (2 + 96 * f) * 256); h3.alloc((size_t)NB * 96 * f * 128);
  a3.alloc((size_t)NB * (1 + 96 * f) * 256); x5.alloc((size_t)NB * 480 * f * 128); e5.alloc((size_t)NB * (2 + 480 * f) * 128);
  h6.alloc((size_t)NB * 480 * f * 64); a6.alloc((size_t)NB * (1 + 480 * f) * 128); x8.alloc((size_t)NB * FRAME * f * 64);
  e8.alloc((size_t)NB * (2 + FRAME * f) * 64); h9.alloc((size_t)NB * FRAME * f * 64); a9.alloc((size_t)NB * (2 + FRAME * f) * 64);
  pcm.alloc((size_t)NB * FRAME * f);
  pcm16.alloc((size_t)NB * FRAME * f);
  zq.alloc((size_t)2 * f * NB * LDIM);
  zq_pos.alloc((size_t)2 * f * NB);
  for (int g = 1; g <= cg; ++g) {
    ConvSegs& sg = segs_by_f[g];
    sg.s[0] = ConvSeg{tr16.p, st_tr.p, 6, 16 * g, 512, 0};
    sg.s[1] = ConvSeg{a0.p, st_a0.p, 1, 16 * g, 512, 0};
    sg.s[2] = ConvSeg{e2.p, st_e2.p, 2, 96 * g, 256, 0};
    sg.s[3] = ConvSeg{a3.p, st_a3.p, 1, 96 * g, 256, 0};
    sg.s[4] = ConvSeg{e5.p, st_e5.p, 2, 480 * g, 128, 0};
    sg.s[5] = ConvSeg{a6.p, st_a6.p, 1, 480 * g, 128, 0};
    sg.s[6] = ConvSeg{e8.p, st_e8.p, 2, FRAME * g, 64, 0};
    sg.s[7] = ConvSeg{a9.p, st_a9.p, 2, FRAME * g, 64, 0};
  }
  segs = segs_by_f[1];
  pend = 0; gbuf = 0;
  PTTS_CUDA(cudaDeviceSynchronize());
}

// PCM -> audio_prompt rows (reference tts_model.rs:504-556 up to the conditioning; models/mimi.rs:113-141):
// zero-pad to whole frames, SEANetEncoder, encoder transformer, ConvDownsample1d, speaker_proj.  One pass over the whole
// prompt (the reference's chunked encoding of long prompts carries one state, i.e. is the same streaming pass; its one
// deviation, the downsample's restarted padding, is patched per chunk boundary below).  Every conv is the same implicit GEMM the decoder uses: a strided conv with k = 2*stride is
// a two-tap conv over the activation viewed as rows of `stride` frames ([T/s][s*C]), so no new GEMM path exists.
void Engine::encode_prompt(const float* pcm_host, int n_samples, std::vector<float>& prompt, int* frames_out) {
  PTTS_REQUIRE(has_encoder, PTTS_ERR_STATE, "this checkpoint has no Mimi encoder tensors (voice cloning needs mimi.encoder.*, mimi.encoder_transformer.*, mimi.downsample.*)");
  const int F = (n_samples + FRAME - 1) / FRAME;
  PTTS_REQUIRE(n_samples >= 1 && F <= 1024, PTTS_ERR_CAPACITY, "voice prompt of %d samples = %d frames (supported: 1..1024 frames)", n_samples, F);
  // tts_model.rs:562-577: the reference encodes long prompts in chunks with ONE carried state, which is exactly one
  // streaming pass -- except that its downsample restarts its replicate padding at every chunk (step = 0, tts_model.rs:540)
  const int chunk_frames = F <= 120 ? F : F <= 600 ? 120 : F <= 1800 ? 180 : 240;
  const int n_boundaries = (F - 1) / chunk_frames;
  ls = stream;
  const int T0 = F * FRAME, ratio[3] = {4, 5, 6};
  int Tl[4] = {T0, T0 / 4, T0 / 20, T0 / 120};
  const int P = Tl[3];
  // Scratch comes from one arena that only ever grows (about 2.4 MB per prompt frame): allocating and freeing a dozen
  // buffers of tens of MB per call made the same 87-frame prompt take anywhere from 12 to 800 ms.
  const size_t need = (size_t)F * 2400000 + (8u << 20);
  if (enc_arena.n < need) { PTTS_CUDA(cudaStreamSynchronize(ls)); enc_arena.alloc(need); }
  PTTS_CUDA(cudaMemsetAsync(enc_arena.p, 0, need, ls));  // zero left-context rows, zero end padding of the prompt
  size_t arena_off = 0;
  auto carve = [&](size_t bytes) {
    void* ptr = enc_arena.p + arena_off;
    arena_off += (bytes + 1023) & ~(size_t)1023;
    PTTS_REQUIRE(arena_off <= need, PTTS_ERR_STATE, "encoder scratch arena too small (%zu > %zu)", arena_off, need);
    return ptr;
  };
  auto f32buf = [&](size_t n) { return static_cast<float*>(carve(n * 4)); };
  auto f16buf = [&](size_t n) { return static_cast<__half*>(carve(n * 2)); };
  float* pcm_d = f32buf(T0);  // the tail past n_samples stays zero: the reference's end padding (tts_model.rs:514-527)
  PTTS_CUDA(cudaMemcpyAsync(pcm_d, pcm_host, (size_t)n_samples * 4, cudaMemcpyHostToDevice, ls));
  float* xcur = f32buf((size_t)T0 * 64);
  __half* ecur = f16buf((size_t)(2 + T0) * 64);
  launch_k(false, enc_conv0_kernel, (unsigned)(((long long)T0 * 64 + 255) / 256), 256, 0, ls, 1, (const float*)pcm_d, T0, (const float*)en_conv0_w.p,
           (const float*)en_conv0_b.p, xcur, ecur);
  GemmEpi e;
  for (int l = 0; l < 3; ++l) {
    const int C = 64 << l, Hp = std::max(C / 2, 64), T = Tl[l], s = ratio[l], Tn = Tl[l + 1];
    // ResBlock (seanet.rs:82-88): v = conv_k1(ELU(conv_k3(ELU(x)))); x += v; then ELU in front of the strided conv
    __half* hbuf = f16buf((size_t)T * Hp);
    e = epi_none(); e.bias = enb_r1[l].p; e.out16 = hbuf; e.act16 = ACT_ELU; e.out16_map = plain_map(Hp);
    tag("encoder.res_a"); gemm(ActView{ecur, C, 2 + T, 1}, 1, T, 3, 128, 1, en_r1[l], Hp, e);
    __half* esbuf = f16buf((size_t)(s + T) * C);  // `s` zero rows of left context in front (k - stride = stride)
    e = epi_none(); e.bias = enb_r3[l].p; e.res = xcur; e.res_map = plain_map(C);
    e.out16 = esbuf; e.act16 = ACT_ELU; e.out16_map = stream_map(T, C, (long long)(s + T) * C, (long long)s * C);
    tag("encoder.res_b"); gemm_rows(hbuf, T, Hp, en_r3[l], C, e);
    // strided conv C -> 2C, k = 2s: two taps over [ (s+T)/s ][ s*C ]
    float* xnext = f32buf((size_t)Tn * 2 * C);
    __half* enext = f16buf((size_t)(2 + Tn) * 2 * C);
    e = epi_none(); e.bias = enb_down[l].p; e.out32 = xnext; e.out32_map = plain_map(2 * C);
    e.out16 = enext; e.act16 = ACT_ELU; e.out16_map = stream_map(Tn, 2 * C, (long long)(2 + Tn) * 2 * C, (long long)2 * 2 * C);
    tag("encoder.down"); gemm(ActView{esbuf, s * C, (s + T) / s, 1}, 1, Tn, 2, 128, 1, en_down[l], 2 * C, e);
    xcur = xnext;
    ecur = enext;
  }
  // last conv k3 512 -> 512 on ELU(x) (seanet.rs:236-246)
  float* tx = f32buf((size_t)P * MIMI_DIM);
  e = epi_none(); e.bias = enb_c11.p; e.out32 = tx; e.out32_map = plain_map(MIMI_DIM);
  tag("encoder.conv11"); gemm(ActView{ecur, MIMI_DIM, 2 + P, 1}, 1, P, 3, 128, 1, en_c11, MIMI_DIM, e);
  // encoder transformer (mimi.rs:129-131; transformer.rs:227-251): causal, context 250, LayerScale
  float* tqkv = f32buf((size_t)P * 3 * MIMI_DIM);
  __half* th16 = f16buf((size_t)P * MIMI_DIM);
  __half* ta16 = f16buf((size_t)P * MIMI_DIM);
  __half* tffn = f16buf((size_t)P * MIMI_FFN);
  tag("encoder.layernorm"); ln<MIMI_DIM>(tx, P, et_ln1_w[0].p, et_ln1_b[0].p, 1e-5f, nullptr, nullptr, 0, th16, MIMI_DIM);
  for (int l = 0; l < MIMI_LAYERS; ++l) {
    e = epi_none(); e.out32 = tqkv; e.out32_map = plain_map(3 * MIMI_DIM);
    tag("encoder.in_proj"); gemm_rows(th16, P, MIMI_DIM, et_inproj[l], 3 * MIMI_DIM, e);
    launch_k(false, enc_rope_kernel, dim3(P, MIMI_HEADS), 32, 0, ls, 1, tqkv, 0);
    launch_k(false, enc_attn_kernel, (unsigned)((P * MIMI_HEADS * 32 + 127) / 128), 128, 0, ls, 1, (const float*)tqkv, P, 250, ta16);
    e = epi_none(); e.fscale = et_ls1[l].p; e.res = tx; e.res_map = plain_map(MIMI_DIM); e.out32 = tx; e.out32_map = plain_map(MIMI_DIM);
    ln_after_next_gemm("encoder.layernorm", tx, P, MIMI_DIM, et_ln2_w[l].p, et_ln2_b[l].p, 1e-5f, nullptr, nullptr, 0, th16, MIMI_DIM);
    tag("encoder.out_proj"); gemm_rows(ta16, P, MIMI_DIM, et_outproj[l], MIMI_DIM, e, true);
    e = epi_none(); e.act = ACT_GELU; e.out16 = tffn; e.out16_map = plain_map(MIMI_FFN);
    tag("encoder.linear1"); gemm_rows(th16, P, MIMI_DIM, et_lin1[l], MIMI_FFN, e);
    e = epi_none(); e.fscale = et_ls2[l].p; e.res = tx; e.res_map = plain_map(MIMI_DIM); e.out32 = tx; e.out32_map = plain_map(MIMI_DIM);
    if (l + 1 < MIMI_LAYERS)
      ln_after_next_gemm("encoder.layernorm", tx, P, MIMI_DIM, et_ln1_w[l + 1].p, et_ln1_b[l + 1].p, 1e-5f, nullptr, nullptr, 0, th16, MIMI_DIM);
    tag("encoder.linear2"); gemm_rows(tffn, P, MIMI_FFN, et_lin2[l], MIMI_DIM, e, true);
  }
  // ConvDownsample1d: stride 16, k 32, no bias, replicate padding (conv.rs:278-312) -> [F][512]; then speaker_proj
  __half* d16 = f16buf((size_t)(16 + P) * MIMI_DIM);
  launch_k(false, enc_downsample_prep_kernel, 16 + P, 128, 0, ls, 1, (const float*)tx, d16);
  __half* lat16 = f16buf((size_t)std::max(F, 256) * MIMI_DIM);
  e = epi_none(); e.out16 = lat16; e.out16_map = plain_map(MIMI_DIM);
  tag("encoder.downsample"); gemm(ActView{d16, 16 * MIMI_DIM, (16 + P) / 16, 1}, 1, F, 2, 128, 1, en_ds, MIMI_DIM, e, true);
  if (n_boundaries > 0) {  // first frame of every later chunk: replicate padding instead of the previous 16 positions
    __half* d16b = f16buf((size_t)n_boundaries * 32 * MIMI_DIM);
    launch_k(false, enc_downsample_boundary_prep_kernel, dim3(n_boundaries, 32), 128, 0, ls, 1, (const float*)tx, chunk_frames, d16b);
    e = epi_none(); e.out16 = lat16;
    e.out16_map = stream_map(1, MIMI_DIM, (long long)chunk_frames * MIMI_DIM, (long long)chunk_frames * MIMI_DIM);
    tag("encoder.downsample"); gemm(ActView{d16b, 16 * MIMI_DIM, 2, n_boundaries}, n_boundaries, 1, 2, 1, 128, en_ds, MIMI_DIM, e, true);
  }
  float* prompt_d = f32buf((size_t)F * D_MODEL);
  e = epi_none(); e.out32 = prompt_d; e.out32_map = plain_map(D_MODEL);
  tag("encoder.speaker_proj"); gemm_rows(lat16, F, MIMI_DIM, w_spk, D_MODEL, e);
  prompt.resize((size_t)F * D_MODEL);
  PTTS_CUDA(cudaMemcpyAsync(prompt.data(), prompt_d, prompt.size() * 4, cudaMemcpyDeviceToHost, ls));
  PTTS_CUDA(cudaStreamSynchronize(ls));
  PTTS_CUDA(cudaGetLastError());
  *frames_out = F;
}

// Tiles for flowlm_attn_prefill_mma_kernel: runs of up to PF_QT consecutive rows of one sequence at consecutive positions.
void Engine::set_prefill_tiles(const std::vector<int>& rs, const std::vector<int>& rp) {
  pf_tiles = 0; pf_kmax = 0;
  if (!pf_tiled || cfg.debug_gemm) return;
  std::vector<int2> t;
  const int rows = (int)rs.size();
  for (int r = 0; r < rows;) {
    int n = 1;
    while (r + n < rows && n < PF_QT && rs[r + n] == rs[r] && rp[r + n] == rp[r + n - 1] + 1) ++n;
    t.push_back(make_int2(r, n));
    pf_kmax = std::max(pf_kmax, rp[r + n - 1] + 1);
    r += n;
  }
  PTTS_CUDA(cudaMemcpyAsync(ptiles.p, t.data(), t.size() * sizeof(int2), cudaMemcpyHostToDevice, stream));
  pf_tiles = (int)t.size();
}

void Engine::prefill(int rows) {
  ls = stream;
  tag("flowlm.layernorm"); ln<D_MODEL>(px32.p, rows, ln1_w[0].p, ln1_b[0].p, 1e-5f, nullptr, nullptr, 0, ph16.p, D_MODEL);
  flowlm_layers(rows, px32.p, ph16.p, pqkv32.p, pattn16.p, pffn16.p, true, pqrot.p, prow_seq.p, prow_pos.p);
}

static std::mutex g_mu;  // engine handles are not re-entrant; this only serialises create/destroy bookkeeping

}  // namespace ptts

// ================================================================================================ C ABI
using namespace ptts;

struct ptts_engine { Engine e; };
struct ptts_voice { Voice v; };

#define PTTS_TRY try {
#define PTTS_CATCH                                            \
  }                                                           \
  catch (const ptts::Error& ex) {                             \
    g_last_error = ex.what();                                 \
    return ex.code;                                           \
  }                                                           \
  catch (const std::exception& ex) {                          \
    g_last_error = ex.what();                                 \
    return PTTS_ERR_INVALID;                                  \
  }

extern "C" {

const char* ptts_last_error(void) { return g_last_error.c_str(); }
int32_t ptts_abi_version(void) { return PTTS_ABI_VERSION; }

int32_t ptts_engine_create(const ptts_engine_cfg* cfg, const ptts_tensor_desc* weights, int32_t n_weights, ptts_engine** out) {
  PTTS_TRY
  PTTS_REQUIRE(cfg && weights && out && n_weights > 0, PTTS_ERR_INVALID, "ptts_engine_create: null argument");
  std::unique_ptr<ptts_engine> h(new ptts_engine);
  h->e.init(*cfg, weights, n_weights);
  *out = h.release();
  return PTTS_OK;
  PTTS_CATCH
}

void ptts_engine_destroy(ptts_engine* e) {
  if (!e) return;
  cudaSetDevice(e->e.cfg.device);
  cudaDeviceSynchronize();
  delete e;
}

int32_t ptts_engine_set_lsd_steps(ptts_engine* h, int32_t lsd_steps) {
  PTTS_TRY
  PTTS_REQUIRE(h && lsd_steps >= 1 && lsd_steps <= MAX_LSD, PTTS_ERR_INVALID, "lsd_steps must be in [1,64]");
  PTTS_CUDA(cudaSetDevice(h->e.cfg.device));
  h->e.sync_all();  // no step may be reading the time embeddings while they are rewritten
  h->e.compute_time_embeddings(lsd_steps);
  return PTTS_OK;
  PTTS_CATCH
}

int32_t ptts_engine_set_codec_group(ptts_engine* h, int32_t frames) {
  PTTS_TRY
  PTTS_REQUIRE(h, PTTS_ERR_INVALID, "null engine");
  Engine& e = h->e;
  PTTS_CUDA(cudaSetDevice(e.cfg.device));
  for (const Engine::Ticket& t : e.tickets)
    PTTS_REQUIRE(t.flags_done && t.pcm_done, PTTS_ERR_STATE, "set_codec_group: step %lld still has unfetched results", t.id);
  e.set_codec_group(frames);
  return PTTS_OK;
  PTTS_CATCH
}

int32_t ptts_voice_from_prompt(ptts_engine* h, const float* audio_prompt, int32_t n_rows, ptts_voice** out) {
  PTTS_TRY
  PTTS_REQUIRE(h && audio_prompt && out, PTTS_ERR_INVALID, "ptts_voice_from_prompt: null argument");
  Engine& e = h->e;
  PTTS_REQUIRE(n_rows >= 1 && n_rows <= 1024, PTTS_ERR_CAPACITY, "voice prompt of %d rows (supported: 1..1024)", n_rows);
  PTTS_CUDA(cudaSetDevice(e.cfg.device));
  std::unique_ptr<ptts_voice> v(new ptts_voice);
  v->v.len = n_rows;
  v->v.prompt.assign(audio_prompt, audio_prompt + (size_t)n_rows * D_MODEL);
  v->v.kv.alloc((size_t)N_LAYERS * 2 * N_HEADS * n_rows * HD);
  SeqDesc sd{v->v.kv.p, nullptr, n_rows, 0, 0, 0};
  PTTS_CUDA(cudaMemcpyAsync(e.seqs.p + e.NS, &sd, sizeof sd, cudaMemcpyHostToDevice, e.stream));
  std::vector<int> rs(n_rows, e.NS), rp(n_rows);
  for (int i = 0; i < n_rows; ++i) rp[i] = i;
  PTTS_CUDA(cudaMemcpyAsync(e.prow_seq.p, rs.data(), n_rows * 4, cudaMemcpyHostToDevice, e.stream));
  PTTS_CUDA(cudaMemcpyAsync(e.prow_pos.p, rp.data(), n_rows * 4, cudaMemcpyHostToDevice, e.stream));
  e.set_prefill_tiles(rs, rp);
  PTTS_CUDA(cudaMemcpyAsync(e.px32.p, audio_prompt, (size_t)n_rows * D_MODEL * 4, cudaMemcpyHostToDevice, e.stream));
  e.prefill(n_rows);
  PTTS_CUDA(cudaStreamSynchronize(e.stream));
  *out = v.release();
  return PTTS_OK;
  PTTS_CATCH
}

int32_t ptts_audio_prompt_from_pcm(ptts_engine* h, const float* pcm24k, int32_t n_samples, float* audio_prompt_out, int32_t cap_rows,
                                   int32_t* n_rows_out) {
  PTTS_TRY
  PTTS_REQUIRE(h && pcm24k && n_rows_out, PTTS_ERR_INVALID, "ptts_audio_prompt_from_pcm: null argument");
  Engine& e = h->e;
  PTTS_CUDA(cudaSetDevice(e.cfg.device));
  e.sync_all();
  std::vector<float> prompt;
  int frames = 0;
  e.encode_prompt(pcm24k, n_samples, prompt, &frames);
  *n_rows_out = frames;
  if (audio_prompt_out) {
    PTTS_REQUIRE(cap_rows >= frames, PTTS_ERR_INVALID, "audio_prompt_out holds %d rows, the prompt has %d", cap_rows, frames);
    std::memcpy(audio_prompt_out, prompt.data(), prompt.size() * 4);
  }
  return PTTS_OK;
  PTTS_CATCH
}

int32_t ptts_voice_from_pcm(ptts_engine* h, const float* pcm24k, int32_t n_samples, ptts_voice** out) {
  PTTS_TRY
  PTTS_REQUIRE(h && pcm24k && out, PTTS_ERR_INVALID, "ptts_voice_from_pcm: null argument");
  Engine& e = h->e;
  PTTS_CUDA(cudaSetDevice(e.cfg.device));
  e.sync_all();
  std::vector<float> prompt;
  int frames = 0;
  e.encode_prompt(pcm24k, n_samples, prompt, &frames);
  return ptts_voice_from_prompt(h, prompt.data(), frames, out);
  PTTS_CATCH
}

void ptts_voice_destroy(ptts_engine* h, ptts_voice* v) {
  if (!v) return;
  if (h) { cudaSetDevice(h->e.cfg.device); cudaStreamSynchronize(h->e.stream); cudaStreamSynchronize(h->e.stream_b); }
  delete v;
}

int32_t ptts_voice_len(const ptts_voice* v) { return v ? v->v.len : -1; }

// Opens n streams: validates everything, claims free slots, stages one record per stream in pinned memory, ONE copy,
// one scatter kernel (KV descriptor, control block, cursor, BOS feedback, zeroed streaming state), then the batched text
// prefill.  No host synchronisation: a recycled slot's last codec frame may still be running on the codec stream, so the
// language-model stream waits for it on the device (ev_b_done), and every host buffer is either pinned and owned by the
// engine until the next open or pageable (staged by the runtime before cudaMemcpyAsync returns).
static void streams_open_impl(Engine& e, int n, ptts_voice* const* voices, const int32_t* tokens, const int32_t* token_offsets,
                              const ptts_stream_params* params, int32_t* slots_out) {
  std::vector<int> free_slots;
  for (int s = 0; s < e.NS && (int)free_slots.size() < n; ++s) if (!e.slots[s].in_use) free_slots.push_back(s);
  PTTS_REQUIRE((int)free_slots.size() == n, PTTS_ERR_CAPACITY, "%d streams requested, %zu slots free", n, free_slots.size());
  for (int i = 0; i < n; ++i) {
    const int nt = token_offsets[i + 1] - token_offsets[i];
    PTTS_REQUIRE(voices[i], PTTS_ERR_INVALID, "stream %d: null voice", i);
    PTTS_REQUIRE(nt >= 0 && nt <= e.PR, PTTS_ERR_INVALID, "stream %d: %d tokens", i, nt);
    PTTS_REQUIRE(params[i].max_gen_len >= 1, PTTS_ERR_INVALID, "stream %d: max_gen_len %d", i, params[i].max_gen_len);
    PTTS_REQUIRE(nt + params[i].max_gen_len <= e.KVCAP, PTTS_ERR_CAPACITY, "stream %d: %d tokens + %d frames exceed kv_capacity %d",
                 i, nt, params[i].max_gen_len, e.KVCAP);
    for (int j = token_offsets[i]; j < token_offsets[i + 1]; ++j)
      PTTS_REQUIRE(tokens[j] >= 0 && tokens[j] <= N_BINS, PTTS_ERR_INVALID, "stream %d: token id %d out of range", i, tokens[j]);
  }
  // the records staged two opens ago must have left this pinned buffer (long done in practice), and the codec stream must
  // be done with the slots being recycled (ordered on the device)
  e.flush_codec();
  const int ob = e.open_parity;
  e.open_parity ^= 1;
  PTTS_CUDA(cudaEventSynchronize(e.ev_open[ob]));
  PTTS_CUDA(cudaStreamWaitEvent(e.stream, e.ev_b_done, 0));
  const size_t per_slot_kv = (size_t)N_LAYERS * 2 * N_HEADS * e.KVCAP * HD;
  // from here on slots are claimed: a failure part-way releases every slot claimed so far (the caller has no ids yet)
  struct Rollback {
    Engine& e; const std::vector<int>& ids; bool armed = true;
    ~Rollback() { if (armed) { cudaStreamSynchronize(e.stream); for (int s : ids) e.slots[s] = SlotHost{}; e.row_seq_host.clear(); } }
  } rollback{e, free_slots};
  for (int i = 0; i < n; ++i) {
    const int s = free_slots[i];
    SlotHost& sh = e.slots[s];
    const int nt = token_offsets[i + 1] - token_offsets[i];
    sh = SlotHost{};
    sh.in_use = true; sh.voice = &voices[i]->v; sh.own_len = nt; sh.max_gen_len = params[i].max_gen_len;
    const float* noise_dev = nullptr;
    if (params[i].noise) {
      DevBuf<float>& nb = e.noise_pool[s];
      const size_t need = (size_t)params[i].max_gen_len * LDIM;
      if (nb.n < need) { PTTS_CUDA(cudaStreamSynchronize(e.stream_b)); nb.alloc(need); }  // grow only; never freed at close
      PTTS_CUDA(cudaMemcpyAsync(nb.p, params[i].noise, need * 4, cudaMemcpyHostToDevice, e.stream));
      noise_dev = nb.p;
    }
    OpenRec& r = e.pin_open[ob][i];
    r.sd = SeqDesc{e.kv.p + (size_t)s * per_slot_kv, sh.voice->kv.p, e.KVCAP, sh.voice->len, sh.voice->len, 0};
    r.ctl = StreamCtl{params[i].max_gen_len, params[i].frames_after_eos, params[i].eos_threshold, params[i].temp,
                      (unsigned long long)params[i].seed, noise_dev, 0, -1, 0, 0};
    r.slot = s; r.own_len = nt; r.pad[0] = r.pad[1] = 0;
    slots_out[i] = s;
  }
  PTTS_CUDA(cudaMemcpyAsync(e.open_recs.p, e.pin_open[ob], (size_t)n * sizeof(OpenRec), cudaMemcpyHostToDevice, e.stream));
  PTTS_CUDA(cudaEventRecord(e.ev_open[ob], e.stream));
  for (int off = 0; off < n; off += 32768) {
    const int cnt = std::min(32768, n - off);
    launch_k(e.use_pdl, slot_open_kernel, dim3(cnt, 10), 128, 0, e.stream, 1, (const OpenRec*)(e.open_recs.p + off), e.segs, e.up_partial.p,
             e.seqs.p, e.ctl.p, e.own_len.p, e.feedback.p, (const float*)e.bos.p);
  }
  // text prefill in groups of at most PR rows (reference tts_model.rs:944-964)
  int i0 = 0;
  while (i0 < n) {
    int i1 = i0, rows = 0;
    while (i1 < n && rows + (token_offsets[i1 + 1] - token_offsets[i1]) <= e.PR) { rows += token_offsets[i1 + 1] - token_offsets[i1]; ++i1; }
    if (rows > 0) {
      std::vector<int> rs(rows), rp(rows);
      int r = 0;
      for (int i = i0; i < i1; ++i)
        for (int j = 0; j < token_offsets[i + 1] - token_offsets[i]; ++j, ++r) { rs[r] = free_slots[i]; rp[r] = voices[i]->v.len + j; }
      PTTS_CUDA(cudaMemcpyAsync(e.prow_seq.p, rs.data(), rows * 4, cudaMemcpyHostToDevice, e.stream));
      PTTS_CUDA(cudaMemcpyAsync(e.prow_pos.p, rp.data(), rows * 4, cudaMemcpyHostToDevice, e.stream));
      e.set_prefill_tiles(rs, rp);
      PTTS_CUDA(cudaMemcpyAsync(e.ptokens.p, tokens + token_offsets[i0], rows * 4, cudaMemcpyHostToDevice, e.stream));
      { ProfScope ps(e, "prefill.embed", (double)rows * 1024 * 8, 0);
        launch_k(e.use_pdl, embed_rows_kernel, rows, 256, 0, e.stream, 1, e.ptokens.p, rows, e.lut.p, e.px32.p); }
      e.prefill(rows);
    }
    i0 = i1;
  }
  e.row_seq_host.clear();
  rollback.armed = false;
}

int32_t ptts_streams_open(ptts_engine* h, int32_t n, ptts_voice* const* voices, const int32_t* tokens,
                          const int32_t* token_offsets, const ptts_stream_params* params, int32_t* slots_out) {
  PTTS_TRY
  PTTS_REQUIRE(h && voices && tokens && token_offsets && params && slots_out && n >= 1, PTTS_ERR_INVALID,
               "ptts_streams_open: null argument");
  PTTS_CUDA(cudaSetDevice(h->e.cfg.device));
  streams_open_impl(h->e, n, voices, tokens, token_offsets, params, slots_out);
  return PTTS_OK;
  PTTS_CATCH
}

static void check_slots(Engine& e, const int32_t* slot_ids, int n, int steps_in_flight = 0) {
  PTTS_REQUIRE(slot_ids && n >= 1 && n <= e.NB, PTTS_ERR_INVALID, "step: n = %d (max_batch %d)", n, e.NB);
  e.slot_mark.assign(e.NS, 0);
  for (int i = 0; i < n; ++i) {
    const int s = slot_ids[i];
    PTTS_REQUIRE(s >= 0 && s < e.NS && e.slots[s].in_use, PTTS_ERR_STATE, "step: slot %d is not open", s);
    // two rows of one batch on the same slot would race on its KV rows and counters
    PTTS_REQUIRE(!e.slot_mark[s], PTTS_ERR_INVALID, "step: slot %d listed twice", s);
    e.slot_mark[s] = 1;
    PTTS_REQUIRE(!e.slots[s].finished, PTTS_ERR_STATE, "step: slot %d already finished", s);
    // a step enqueued ahead of unfetched flags may run past the stream's last frame: its KV row must still exist
    PTTS_REQUIRE(e.slots[s].own_len + steps_in_flight < e.KVCAP, PTTS_ERR_CAPACITY, "step: slot %d KV full", s);
  }
}

// ---- pipelined step: begin (enqueue) / flags (language-model results) / pcm (codec result)
static long long step_begin_impl(Engine& e, const int32_t* slot_ids, int n, int flags) {
  const bool want_i16 = (flags & PTTS_STEP_PCM_I16) != 0;
  const bool want_pcm = (flags & PTTS_STEP_PCM) != 0 || want_i16, ahead = (flags & PTTS_STEP_AHEAD) != 0;
  const long long id = e.next_ticket;
  Engine::Ticket& t = e.tickets[id % Engine::NT];
  PTTS_REQUIRE(t.flags_done && t.pcm_done, PTTS_ERR_STATE, "step %lld still has unfetched results (%d steps may be in flight)", t.id, Engine::NT);
  const Engine::Ticket& prev = e.tickets[(id + Engine::NT - 1) % Engine::NT];
  const Engine::Ticket& prev2 = e.tickets[(id + Engine::NT - 2) % Engine::NT];
  PTTS_REQUIRE(prev.flags_done || ahead, PTTS_ERR_STATE, "fetch the flags of step %lld before beginning the next step (or pass PTTS_STEP_AHEAD)", prev.id);
  PTTS_REQUIRE(prev2.flags_done, PTTS_ERR_STATE, "only one step may be enqueued ahead of unfetched flags (step %lld)", prev2.id);
  check_slots(e, slot_ids, n, prev.flags_done ? 0 : 1);
  e.upload_rows(slot_ids, n);
  const int par = (int)(id % Engine::NT);
  t.id = id; t.n = n; t.want_pcm = want_pcm; t.want_i16 = want_i16; t.codec_pending = e.queued();
  e.run_step(n, id);   // cg > 1: the PCM copy and ev_pcm are issued when the frame's codec group is flushed
  PTTS_CUDA(cudaMemcpyAsync(e.pin_lat[par], e.step_out.p, e.step_out_bytes(), cudaMemcpyDeviceToHost, e.stream));
  PTTS_CUDA(cudaEventRecord(e.ev_flags[par], e.stream));
  if (!e.queued()) {
    if (want_i16) PTTS_CUDA(cudaMemcpyAsync(e.pin_pcm16[par], e.pcm16.p, (size_t)n * FRAME * 2, cudaMemcpyDeviceToHost, e.stream_b));
    else if (want_pcm) PTTS_CUDA(cudaMemcpyAsync(e.pin_pcm[par], e.pcm.p, (size_t)n * FRAME * 4, cudaMemcpyDeviceToHost, e.stream_b));
    PTTS_CUDA(cudaEventRecord(e.ev_pcm[par], e.stream_b));
  }
  t.flags_done = false; t.pcm_done = false;
  t.slot_ids.assign(slot_ids, slot_ids + n);
  e.next_ticket = id + 1;
  return id;
}

static void step_flags_impl(Engine& e, long long id, uint8_t* finished, float* latent_out, float* eos_logit_out) {
  Engine::Ticket& t = e.tickets[id % Engine::NT];
  PTTS_REQUIRE(id >= 0 && t.id == id && !t.flags_done, PTTS_ERR_STATE, "ticket %lld is not pending", id);
  const Engine::Ticket& prev = e.tickets[(id + Engine::NT - 1) % Engine::NT];
  PTTS_REQUIRE(prev.flags_done, PTTS_ERR_STATE, "flags are fetched in step order: step %lld first", prev.id);
  const int par = (int)(id % Engine::NT);
  PTTS_CUDA(cudaEventSynchronize(e.ev_flags[par]));
  if (latent_out) std::memcpy(latent_out, e.pin_lat[par], (size_t)t.n * LDIM * 4);
  if (eos_logit_out) std::memcpy(eos_logit_out, e.pin_logit[par], (size_t)t.n * 4);
  for (int i = 0; i < t.n; ++i) {
    SlotHost& sh = e.slots[t.slot_ids[i]];
    if (sh.finished) {  // enqueued ahead and the stream ended on the step before: this frame is past the end
      if (finished) finished[i] = PTTS_FRAME_OVERRUN;
      continue;
    }
    sh.frames += 1; sh.own_len += 1;
    sh.finished = e.pin_fin[par][i] != 0;
    if (finished) finished[i] = e.pin_fin[par][i];
  }
  t.flags_done = true;
}

static void step_pcm_impl(Engine& e, long long id, float* pcm_out, int16_t* pcm16_out = nullptr) {
  Engine::Ticket& t = e.tickets[id % Engine::NT];
  PTTS_REQUIRE(id >= 0 && t.id == id && !t.pcm_done, PTTS_ERR_STATE, "ticket %lld has no pending PCM", id);
  const int par = (int)(id % Engine::NT);
  if (t.codec_pending) e.flush_codec();   // the frame's group is not full yet: decode what is queued
  PTTS_CUDA(cudaEventSynchronize(e.ev_pcm[par]));
  if (pcm_out) {
    PTTS_REQUIRE(t.want_pcm && !t.want_i16, PTTS_ERR_STATE, "step %lld was not begun with PTTS_STEP_PCM", id);
    std::memcpy(pcm_out, e.pin_pcm[par], (size_t)t.n * FRAME * 4);
  }
  if (pcm16_out) {
    PTTS_REQUIRE(t.want_i16, PTTS_ERR_STATE, "step %lld was not begun with PTTS_STEP_PCM_I16", id);
    std::memcpy(pcm16_out, e.pin_pcm16[par], (size_t)t.n * FRAME * 2);
  }
  t.pcm_done = true;
}

int64_t ptts_step_begin(ptts_engine* h, const int32_t* slot_ids, int32_t n, int32_t flags) {
  try {
    PTTS_REQUIRE(h, PTTS_ERR_INVALID, "null engine");
    PTTS_CUDA(cudaSetDevice(h->e.cfg.device));
    return step_begin_impl(h->e, slot_ids, n, flags);
  } catch (const ptts::Error& ex) {
    g_last_error = ex.what();
    return ex.code;
  } catch (const std::exception& ex) {  // bad_alloc / length_error from the ticket bookkeeping must not cross the C ABI
    g_last_error = ex.what();
    return PTTS_ERR_INVALID;
  }
}

int32_t ptts_step_flags(ptts_engine* h, int64_t ticket, uint8_t* finished, float* latent_out, float* eos_logit_out) {
  PTTS_TRY
  PTTS_REQUIRE(h, PTTS_ERR_INVALID, "null engine");
  PTTS_CUDA(cudaSetDevice(h->e.cfg.device));
  step_flags_impl(h->e, ticket, finished, latent_out, eos_logit_out);
  return PTTS_OK;
  PTTS_CATCH
}

int32_t ptts_step_pcm(ptts_engine* h, int64_t ticket, float* pcm_out) {
  PTTS_TRY
  PTTS_REQUIRE(h, PTTS_ERR_INVALID, "null engine");
  PTTS_CUDA(cudaSetDevice(h->e.cfg.device));
  step_pcm_impl(h->e, ticket, pcm_out);
  return PTTS_OK;
  PTTS_CATCH
}

int32_t ptts_step_pcm_i16(ptts_engine* h, int64_t ticket, int16_t* pcm_out) {
  PTTS_TRY
  PTTS_REQUIRE(h, PTTS_ERR_INVALID, "null engine");
  PTTS_CUDA(cudaSetDevice(h->e.cfg.device));
  step_pcm_impl(h->e, ticket, nullptr, pcm_out);
  return PTTS_OK;
  PTTS_CATCH
}

int32_t ptts_step(ptts_engine* h, const int32_t* slot_ids, int32_t n, float* pcm_out, uint8_t* finished, float* latent_out,
                  float* eos_logit_out) {
  PTTS_TRY
  PTTS_REQUIRE(h, PTTS_ERR_INVALID, "null engine");
  Engine& e = h->e;
  PTTS_CUDA(cudaSetDevice(e.cfg.device));
  const long long id = step_begin_impl(e, slot_ids, n, pcm_out != nullptr ? PTTS_STEP_PCM : 0);
  step_flags_impl(e, id, finished, latent_out, eos_logit_out);
  step_pcm_impl(e, id, pcm_out);
  return PTTS_OK;
  PTTS_CATCH
}

int32_t ptts_step_device(ptts_engine* h, const int32_t* slot_ids, int32_t n) {
  PTTS_TRY
  PTTS_REQUIRE(h, PTTS_ERR_INVALID, "null engine");
  Engine& e = h->e;
  PTTS_CUDA(cudaSetDevice(e.cfg.device));
  check_slots(e, slot_ids, n);
  e.upload_rows(slot_ids, n);
  e.run_step(n);
  for (int i = 0; i < n; ++i) {
    SlotHost& sh = e.slots[slot_ids[i]];
    sh.frames += 1; sh.own_len += 1;
    if (sh.frames >= sh.max_gen_len) sh.finished = true;  // EOS-based finish needs the flags: use ptts_step for that
  }
  return PTTS_OK;
  PTTS_CATCH
}

int32_t ptts_step_timed(ptts_engine* h, const int32_t* slot_ids, int32_t n, float* stage_ms) {
  PTTS_TRY
  PTTS_REQUIRE(h && stage_ms, PTTS_ERR_INVALID, "null argument");
  Engine& e = h->e;
  PTTS_CUDA(cudaSetDevice(e.cfg.device));
  check_slots(e, slot_ids, n);
  e.upload_rows(slot_ids, n);
  e.sync_all();
  e.step_kernels(n, stage_ms);
  for (int i = 0; i < n; ++i) {
    SlotHost& sh = e.slots[slot_ids[i]];
    sh.frames += 1; sh.own_len += 1;
    if (sh.frames >= sh.max_gen_len) sh.finished = true;
  }
  return PTTS_OK;
  PTTS_CATCH
}

int32_t ptts_sync(ptts_engine* h) {
  PTTS_TRY
  PTTS_REQUIRE(h, PTTS_ERR_INVALID, "null engine");
  PTTS_CUDA(cudaSetDevice(h->e.cfg.device));
  h->e.sync_all();
  if (h->e.fh_trace.p) {
    unsigned long long st[FH_LAYERS * 8];
    if (cudaMemcpy(st, h->e.fh_trace.p, sizeof st, cudaMemcpyDeviceToHost) == cudaSuccess && st[0]) {
      std::fprintf(stderr, "ptts flow head trace (ns since layer-0 entry): layer: start | act TMA issued, first k-block landed, last landed | acc_ready math_done barrier\n");
      for (int L = 0; L < FH_LAYERS; ++L)
        std::fprintf(stderr, "  L%02d: %6lld | %6lld %6lld %6lld | %6lld %6lld %6lld\n", L, (long long)(st[L * 8] - st[0]), (long long)(st[L * 8 + 4] - st[0]),
                     (long long)(st[L * 8 + 5] - st[0]), (long long)(st[L * 8 + 6] - st[0]), (long long)(st[L * 8 + 1] - st[0]),
                     (long long)(st[L * 8 + 2] - st[0]), (long long)(st[L * 8 + 3] - st[0]));
    }
  }
  if (h->e.lm_trace.p) {
    unsigned long long st[128];
    if (cudaMemcpy(st, h->e.lm_trace.p, sizeof st, cudaMemcpyDeviceToHost) == cudaSuccess && st[0]) {
      const int nph = LM_PH_ADA + h->e.lsd_steps;
      std::fprintf(stderr, "ptts step kernel trace (CTA 0, ns): phase: start(since phase 0) work barrier\n");
      for (int ph = 0; ph < nph; ++ph) {
        const long long start = (long long)(st[ph * 2] - st[0]), work = (long long)(st[ph * 2 + 1] - st[ph * 2]);
        const long long bar = ph + 1 < nph ? (long long)(st[ph * 2 + 2] - st[ph * 2 + 1]) : 0;
        std::fprintf(stderr, "  ph%02d: %7lld %6lld %6lld\n", ph, start, work, bar);
      }
      std::fprintf(stderr, "  attention of layer 1, CTA 0 warp 0 (ns since its phase start): ");
      for (int i = 0; i < 21; ++i) std::fprintf(stderr, "%lld ", st[104 + i] ? (long long)(st[104 + i] - st[10 * 2]) : -1LL);
      std::fprintf(stderr, "\n");
    }
  }
  if (h->e.diag_times) {
    float a = 0, b = 0, ab = 0, ae = 0;
    if (cudaEventElapsedTime(&a, h->e.ev_t[0], h->e.ev_t[1]) == cudaSuccess && cudaEventElapsedTime(&b, h->e.ev_t[2], h->e.ev_t[3]) == cudaSuccess &&
        cudaEventElapsedTime(&ab, h->e.ev_t[0], h->e.ev_t[2]) == cudaSuccess && cudaEventElapsedTime(&ae, h->e.ev_t[0], h->e.ev_t[3]) == cudaSuccess)
      std::fprintf(stderr, "ptts diag: last step  A %.1f us | B %.1f us | A start -> B start %.1f us | A start -> B end %.1f us\n", a * 1e3f, b * 1e3f,
                   ab * 1e3f, ae * 1e3f);
    (void)cudaGetLastError();
  }
  return PTTS_OK;
  PTTS_CATCH
}

int32_t ptts_stream_set_feedback(ptts_engine* h, int32_t slot, const float* latent32) {
  PTTS_TRY
  PTTS_REQUIRE(h && latent32, PTTS_ERR_INVALID, "null argument");
  Engine& e = h->e;
  PTTS_REQUIRE(slot >= 0 && slot < e.NS && e.slots[slot].in_use, PTTS_ERR_STATE, "slot %d is not open", slot);
  PTTS_CUDA(cudaSetDevice(e.cfg.device));
  PTTS_CUDA(cudaMemcpyAsync(e.feedback.p + (size_t)slot * LDIM, latent32, LDIM * 4, cudaMemcpyHostToDevice, e.stream));
  PTTS_CUDA(cudaStreamSynchronize(e.stream));
  return PTTS_OK;
  PTTS_CATCH
}

// Closing is host bookkeeping: nothing of a slot is freed (its injected-noise buffer belongs to a pool), kernels already
// enqueued for it run to completion, and the next open of the slot orders itself behind them on the device.
static void stream_close_impl(Engine& e, int slot) {
  PTTS_REQUIRE(slot >= 0 && slot < e.NS && e.slots[slot].in_use, PTTS_ERR_STATE, "slot %d is not open", slot);
  PTTS_REQUIRE(!e.slot_in_pending_ticket(slot), PTTS_ERR_STATE, "slot %d is part of a step whose flags have not been fetched (ptts_step_flags first)", slot);
  e.flush_codec();   // its queued frames are decoded before the slot can be recycled
  e.slots[slot] = SlotHost{};
  e.row_seq_host.clear();
}

int32_t ptts_stream_close(ptts_engine* h, int32_t slot) {
  PTTS_TRY
  PTTS_REQUIRE(h, PTTS_ERR_INVALID, "null engine");
  stream_close_impl(h->e, slot);
  return PTTS_OK;
  PTTS_CATCH
}

int32_t ptts_streams_close(ptts_engine* h, const int32_t* slots, int32_t n) {
  PTTS_TRY
  PTTS_REQUIRE(h && slots && n >= 0, PTTS_ERR_INVALID, "ptts_streams_close: null argument");
  for (int i = 0; i < n; ++i) stream_close_impl(h->e, slots[i]);
  return PTTS_OK;
  PTTS_CATCH
}

int32_t ptts_stream_frames(ptts_engine* h, int32_t slot, int32_t* frames_out, int32_t* eos_step_out) {
  PTTS_TRY
  PTTS_REQUIRE(h, PTTS_ERR_INVALID, "null engine");
  Engine& e = h->e;
  PTTS_REQUIRE(slot >= 0 && slot < e.NS && e.slots[slot].in_use, PTTS_ERR_STATE, "slot %d is not open", slot);
  PTTS_CUDA(cudaSetDevice(e.cfg.device));
  StreamCtl c;
  PTTS_CUDA(cudaStreamSynchronize(e.stream));
  PTTS_CUDA(cudaMemcpy(&c, e.ctl.p + slot, sizeof c, cudaMemcpyDeviceToHost));
  if (frames_out) *frames_out = c.frame;
  if (eos_step_out) *eos_step_out = c.eos_step;
  return PTTS_OK;
  PTTS_CATCH
}

int64_t ptts_debug_read(ptts_engine* h, const char* name, int32_t row, float* out, int64_t cap) {
  try {
    PTTS_REQUIRE(h && name && out, PTTS_ERR_INVALID, "null argument");
    Engine& e = h->e;
    PTTS_REQUIRE(row >= 0 && row < e.NB, PTTS_ERR_INVALID, "row %d out of range", row);
    PTTS_CUDA(cudaSetDevice(e.cfg.device));
    e.sync_all();
    const std::string n(name);
    const float* src = nullptr;
    int64_t cnt = 0;
    if (n == "flowlm.x") { src = e.x32.p + (size_t)row * D_MODEL; cnt = D_MODEL; }
    else if (n == "flowlm.h") { src = e.h32dbg.p + (size_t)row * D_MODEL; cnt = D_MODEL; }
    else if (n == "flowlm.qkv") { src = e.qkv32.p + (size_t)row * 3 * D_MODEL; cnt = 3 * D_MODEL; }
    else if (n == "flow.c") { src = e.c32.p + (size_t)row * FLOW_DIM; cnt = FLOW_DIM; }
    else if (n == "flow.x") { src = e.fx32.p + (size_t)row * FLOW_DIM; cnt = FLOW_DIM; }
    else if (n == "mimi.quantized") { src = e.quant_dbg.p + (size_t)row * 512; cnt = 512; }
    else if (n == "mimi.after_decoder_transformer") { src = e.mx32.p + (size_t)row * 16 * 512; cnt = 16 * 512; }
    else if (n == "seanet.convtr2") { src = e.x2.p + (size_t)row * 96 * 256; cnt = 96 * 256; }
    else if (n == "seanet.convtr5") { src = e.x5.p + (size_t)row * 480 * 128; cnt = 480 * 128; }
    else if (n == "seanet.convtr8") { src = e.x8.p + (size_t)row * 1920 * 64; cnt = 1920 * 64; }
    else if (n == "pcm") { src = e.pcm.p + (size_t)row * FRAME; cnt = FRAME; }
    // raw buffers of the persistent step kernel (bring-up): `row` selects a 1 M-float window
    else if (n == "lm.ws") { src = e.lm_ws.p + (size_t)row * (1 << 20); cnt = std::min<int64_t>(1 << 20, (int64_t)e.lm_ws.n - (int64_t)row * (1 << 20)); }
    else if (n == "lm.hA") { src = reinterpret_cast<const float*>(e.lm_hA.p); cnt = (int64_t)e.lm_hA.n / 4; }
    else if (n == "lm.attnA") { src = reinterpret_cast<const float*>(e.lm_attnA.p); cnt = (int64_t)e.lm_attnA.n / 4; }
    else if (n == "lm.ffnA") { src = reinterpret_cast<const float*>(e.lm_ffnA.p); cnt = (int64_t)e.lm_ffnA.n / 4; }
    else PTTS_REQUIRE(false, PTTS_ERR_INVALID, "unknown tap '%s'", name);
    PTTS_REQUIRE(cnt <= cap, PTTS_ERR_INVALID, "tap '%s' needs %lld floats, buffer holds %lld", name, (long long)cnt, (long long)cap);
    PTTS_CUDA(cudaMemcpy(out, src, cnt * 4, cudaMemcpyDeviceToHost));
    return cnt;
  } catch (const ptts::Error& ex) {
    g_last_error = ex.what();
    return ex.code;
  }
}

int64_t ptts_launch_count(ptts_engine* h, int32_t reset) {
  if (!h) return -1;
  const long long v = h->e.launches;
  if (reset) h->e.launches = 0;
  return v;
}

void* ptts_cuda_stream(ptts_engine* h) { return h ? (void*)h->e.stream : nullptr; }

int32_t ptts_profile_enable(ptts_engine* h, int32_t on) {
  PTTS_TRY
  PTTS_REQUIRE(h, PTTS_ERR_INVALID, "null engine");
  PTTS_CUDA(cudaSetDevice(h->e.cfg.device));
  h->e.sync_all();
  h->e.profiling = on != 0;
  return PTTS_OK;
  PTTS_CATCH
}

// CUDA-event time of an empty kernel bracketed exactly like a profiled launch (ms): the fixed cost the per-launch
// timings of ptts_profile_report include.  bench.py reports it next to the raw numbers.
int32_t ptts_profile_overhead(ptts_engine* h, float* ms_out) {
  PTTS_TRY
  PTTS_REQUIRE(h && ms_out, PTTS_ERR_INVALID, "null argument");
  Engine& e = h->e;
  PTTS_CUDA(cudaSetDevice(e.cfg.device));
  e.sync_all();
  const bool was = e.profiling;
  e.profiling = true;
  e.ls = e.stream;
  const size_t first = e.prof_recs.size();
  for (int i = 0; i < 64; ++i) {
    ProfScope ps(e, "_empty");
    launch_k(e.use_pdl, fill_f32_kernel, 1, 32, 0, e.stream, 1, (float*)nullptr, 0.f, (long long)0);
  }
  PTTS_CUDA(cudaStreamSynchronize(e.stream));
  double tot = 0;
  for (size_t i = first; i < e.prof_recs.size(); ++i) {
    float ms = 0;
    PTTS_CUDA(cudaEventElapsedTime(&ms, e.prof_recs[i].a, e.prof_recs[i].b));
    tot += ms;
    e.prof_pool.push_back(e.prof_recs[i].a); e.prof_pool.push_back(e.prof_recs[i].b);
  }
  e.prof_recs.resize(first);
  e.launches -= 64;
  e.profiling = was;
  *ms_out = (float)(tot / 64);
  return PTTS_OK;
  PTTS_CATCH
}

// Kernel time of the decode GEMMs without any per-launch event cost: the four FlowLM Linears of a step (in_proj,
// out_proj, linear1, linear2 at `rows` batch rows) replayed as ONE captured graph of iters x 6 layers of back-to-back
// launches per kind -- real weights of all six layers in turn, so every launch streams its weights from HBM like in a
// step (151 MB per round trip > L2) -- bracketed by a single event pair.  us_out[4] = mean us per launch per kind,
// bytes_out[4] = algorithmic bytes per launch.  Launches are the production path (Engine::gemm, PDL on).
int32_t ptts_profile_gemm_replay(ptts_engine* h, int32_t rows, int32_t iters, float* us_out, double* bytes_out) {
  PTTS_TRY
  PTTS_REQUIRE(h && us_out && rows >= 1 && iters >= 1, PTTS_ERR_INVALID, "ptts_profile_gemm_replay: bad arguments");
  Engine& e = h->e;
  PTTS_REQUIRE(rows <= e.NB, PTTS_ERR_INVALID, "rows %d beyond max_batch %d", rows, e.NB);
  PTTS_CUDA(cudaSetDevice(e.cfg.device));
  e.sync_all();
  e.ls = e.stream;
  const bool was = e.profiling;
  e.profiling = false;
  cudaEvent_t a, b;
  PTTS_CUDA(cudaEventCreate(&a)); PTTS_CUDA(cudaEventCreate(&b));
  for (int kind = 0; kind < 4; ++kind) {
    auto one = [&](int l) {
      GemmEpi ep = epi_none();
      if (kind == 0) { ep.out32 = e.qkv32.p; ep.out32_map = plain_map(3 * D_MODEL); if (e.inproj_ctas) e.split_cap_override = e.inproj_ctas; e.gemm_rows(e.h16.p, rows, D_MODEL, e.w_inproj[l], 3 * D_MODEL, ep); }
      else if (kind == 1) { ep.out32 = e.x32.p; ep.out32_map = plain_map(D_MODEL); ep.res = e.x32.p; ep.res_map = plain_map(D_MODEL);
                            if (e.outproj_ctas) e.split_cap_override = e.outproj_ctas;
                            e.gemm_rows(e.attn16.p, rows, D_MODEL, e.w_outproj[l], D_MODEL, ep, true); }
      else if (kind == 2) { ep.act = ACT_GELU; ep.out16 = e.ffn16.p; ep.out16_map = plain_map(D_FFN); e.split_cap_override = e.lin1_ctas;
                            e.gemm_rows(e.h16.p, rows, D_MODEL, e.w_lin1[l], D_FFN, ep); }
      else { ep.out32 = e.x32.p; ep.out32_map = plain_map(D_MODEL); ep.res = e.x32.p; ep.res_map = plain_map(D_MODEL);
             if (e.lin2_ctas) e.split_cap_override = e.lin2_ctas;
             e.gemm_rows(e.ffn16.p, rows, D_FFN, e.w_lin2[l], D_MODEL, ep, true); }
    };
    for (int l = 0; l < N_LAYERS; ++l) one(l);  // warm (tensor maps, instruction cache)
    PTTS_CUDA(cudaStreamSynchronize(e.stream));
    cudaGraph_t g = nullptr;
    cudaGraphExec_t ge = nullptr;
    PTTS_CUDA(cudaStreamBeginCapture(e.stream, cudaStreamCaptureModeThreadLocal));
    for (int it = 0; it < iters; ++it) for (int l = 0; l < N_LAYERS; ++l) one(l);
    PTTS_CUDA(cudaStreamEndCapture(e.stream, &g));
    PTTS_CUDA(cudaGraphInstantiate(&ge, g, 0));
    PTTS_CUDA(cudaGraphLaunch(ge, e.stream));  // once untimed
    PTTS_CUDA(cudaEventRecord(a, e.stream));
    PTTS_CUDA(cudaGraphLaunch(ge, e.stream));
    PTTS_CUDA(cudaEventRecord(b, e.stream));
    PTTS_CUDA(cudaStreamSynchronize(e.stream));
    float ms = 0;
    PTTS_CUDA(cudaEventElapsedTime(&ms, a, b));
    us_out[kind] = 1000.f * ms / (float)(iters * N_LAYERS);
    cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
    if (bytes_out) {
      const double F = kind == 0 ? 3.0 * D_MODEL : kind == 2 ? (double)D_FFN : (double)D_MODEL, K = kind == 3 ? (double)D_FFN : (double)D_MODEL;
      const double outb = kind == 0 ? 4 : kind == 2 ? 2 : 8;  // f32 out | f16 out | f32 residual in + out
      bytes_out[kind] = F * K * 2 + (double)rows * K * 2 + (double)rows * F * outb;
    }
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  e.launches -= 0;
  e.profiling = was;
  // the replay scribbled over the residual stream / operand buffers of the decode scratch; they are rebuilt by the next step
  return PTTS_OK;
  PTTS_CATCH
}

int64_t ptts_profile_report(ptts_engine* h, char* buf, int64_t cap) {
  try {
    PTTS_REQUIRE(h && buf && cap > 0, PTTS_ERR_INVALID, "null argument");
    PTTS_CUDA(cudaSetDevice(h->e.cfg.device));
    const std::string r = h->e.prof_report();
    PTTS_REQUIRE((int64_t)r.size() + 1 <= cap, PTTS_ERR_INVALID, "profile report needs %zu bytes", r.size() + 1);
    std::memcpy(buf, r.c_str(), r.size() + 1);
    return (int64_t)r.size();
  } catch (const ptts::Error& ex) {
    g_last_error = ex.what();
    return ex.code;
  }
}


// ------------------------------------------------------------------------------------------------ voice files, config
namespace {

// The few safetensors features a voice file needs: 8-byte header length, JSON header with one object per tensor
// ({"dtype", "shape", "data_offsets"}), raw little-endian data behind it.
struct StTensor { std::string dtype; std::vector<int64_t> shape; size_t begin = 0, end = 0; };

std::map<std::string, StTensor> st_parse_header(const std::string& js) {
  std::map<std::string, StTensor> out;
  size_t i = 0;
  auto skip_ws = [&]() { while (i < js.size() && (js[i] == ' ' || js[i] == '\n' || js[i] == '\t' || js[i] == '\r')) ++i; };
  auto parse_string = [&]() {
    PTTS_REQUIRE(i < js.size() && js[i] == '"', PTTS_ERR_INVALID, "safetensors header: string expected at %zu", i);
    std::string v;
    for (++i; i < js.size() && js[i] != '"'; ++i) { if (js[i] == '\\' && i + 1 < js.size()) ++i; v.push_back(js[i]); }
    ++i;
    return v;
  };
  skip_ws();
  PTTS_REQUIRE(i < js.size() && js[i] == '{', PTTS_ERR_INVALID, "safetensors header is not a JSON object");
  ++i;
  while (true) {
    skip_ws();
    if (i >= js.size() || js[i] == '}') break;
    if (js[i] == ',') { ++i; continue; }
    const std::string name = parse_string();
    skip_ws();
    PTTS_REQUIRE(i < js.size() && js[i] == ':', PTTS_ERR_INVALID, "safetensors header: ':' expected");
    ++i;
    skip_ws();
    PTTS_REQUIRE(i < js.size() && js[i] == '{', PTTS_ERR_INVALID, "safetensors header: object expected for '%s'", name.c_str());
    const size_t obj0 = i;
    int depth = 0;
    for (; i < js.size(); ++i) {  // strings in a tensor / metadata object may hold braces only inside quotes
      if (js[i] == '"') { for (++i; i < js.size() && js[i] != '"'; ++i) if (js[i] == '\\') ++i; continue; }
      if (js[i] == '{') ++depth;
      if (js[i] == '}' && --depth == 0) { ++i; break; }
    }
    if (name == "__metadata__") continue;
    const std::string obj = js.substr(obj0, i - obj0);
    StTensor t;
    auto field = [&](const char* key) {
      const size_t k = obj.find(std::string("\"") + key + "\"");
      PTTS_REQUIRE(k != std::string::npos, PTTS_ERR_INVALID, "safetensors tensor '%s' has no %s", name.c_str(), key);
      return obj.find(':', k) + 1;
    };
    { size_t k = field("dtype"); k = obj.find('"', k); t.dtype = obj.substr(k + 1, obj.find('"', k + 1) - k - 1); }
    auto ints = [&](size_t k) {
      std::vector<int64_t> v;
      k = obj.find('[', k);
      const size_t e = obj.find(']', k);
      std::stringstream ss(obj.substr(k + 1, e - k - 1));
      std::string tok;
      while (std::getline(ss, tok, ',')) if (tok.find_first_of("0123456789") != std::string::npos) v.push_back(std::stoll(tok));
      return v;
    };
    t.shape = ints(field("shape"));
    const std::vector<int64_t> off = ints(field("data_offsets"));
    PTTS_REQUIRE(off.size() == 2 && off[0] <= off[1], PTTS_ERR_INVALID, "safetensors tensor '%s': bad data_offsets", name.c_str());
    t.begin = (size_t)off[0]; t.end = (size_t)off[1];
    out[name] = t;
  }
  return out;
}

std::vector<float> st_to_f32(const StTensor& t, const unsigned char* data) {
  size_t n = 1;
  for (auto d : t.shape) n *= (size_t)d;
  std::vector<float> v(n);
  const unsigned char* p = data + t.begin;
  if (t.dtype == "F32") { PTTS_REQUIRE(t.end - t.begin == n * 4, PTTS_ERR_INVALID, "safetensors: size mismatch"); std::memcpy(v.data(), p, n * 4); }
  else if (t.dtype == "BF16") { PTTS_REQUIRE(t.end - t.begin == n * 2, PTTS_ERR_INVALID, "safetensors: size mismatch"); for (size_t i = 0; i < n; ++i) { uint16_t u; std::memcpy(&u, p + 2 * i, 2); v[i] = bf16_to_f32(u); } }
  else if (t.dtype == "F16") { PTTS_REQUIRE(t.end - t.begin == n * 2, PTTS_ERR_INVALID, "safetensors: size mismatch"); for (size_t i = 0; i < n; ++i) { __half hv; std::memcpy(&hv, p + 2 * i, 2); v[i] = __half2float(hv); } }
  else PTTS_REQUIRE(false, PTTS_ERR_INVALID, "safetensors dtype %s not supported for a voice prompt", t.dtype.c_str());
  return v;
}

// A YAML subset: nested maps by indentation, scalars, block lists ("- 6").  -> {"mimi.seanet.ratios.0": "6", ...}
std::map<std::string, std::string> yaml_flatten(std::istream& in) {
  std::map<std::string, std::string> out;
  std::vector<std::pair<int, std::string>> stack;  // (indent, key)
  std::map<std::string, int> list_len;
  std::string line;
  auto trim = [](std::string v) {
    const size_t a = v.find_first_not_of(" \t\r"), b = v.find_last_not_of(" \t\r");
    return a == std::string::npos ? std::string() : v.substr(a, b - a + 1);
  };
  while (std::getline(in, line)) {
    const size_t hash = line.find('#');
    if (hash != std::string::npos && (hash == 0 || line[hash - 1] == ' ')) line = line.substr(0, hash);
    if (trim(line).empty()) continue;
    const int indent = (int)line.find_first_not_of(' ');
    std::string body = trim(line);
    if (body[0] == '-') {  // list item under the innermost key at a smaller-or-equal indent
      while (!stack.empty() && stack.back().first > indent) stack.pop_back();
      while (stack.size() > 1 && stack.back().first == indent && stack[stack.size() - 2].first == indent) stack.pop_back();
      std::string path;
      for (auto& kv : stack) path += (path.empty() ? "" : ".") + kv.second;
      out[path + "." + std::to_string(list_len[path]++)] = trim(body.substr(1));
      continue;
    }
    const size_t colon = body.find(':');
    if (colon == std::string::npos) continue;
    while (!stack.empty() && stack.back().first >= indent) stack.pop_back();
    stack.push_back({indent, trim(body.substr(0, colon))});
    const std::string val = trim(body.substr(colon + 1));
    if (!val.empty()) {
      std::string path;
      for (auto& kv : stack) path += (path.empty() ? "" : ".") + kv.second;
      out[path] = val;
    }
  }
  return out;
}

}  // namespace

int32_t ptts_voice_save(ptts_engine* h, const ptts_voice* v, const char* path, int32_t include_kv) {
  PTTS_TRY
  PTTS_REQUIRE(h && v && path, PTTS_ERR_INVALID, "ptts_voice_save: null argument");
  Engine& e = h->e;
  const int T = v->v.len;
  PTTS_REQUIRE((int)v->v.prompt.size() == T * D_MODEL, PTTS_ERR_STATE, "voice has no conditioning rows to save");
  const size_t n_prompt = (size_t)T * D_MODEL * 4, n_kv = include_kv ? v->v.kv.n * 2 : 0;
  std::string js = fmt("{\"audio_prompt\":{\"dtype\":\"F32\",\"shape\":[1,%d,%d],\"data_offsets\":[0,%zu]}", T, D_MODEL, n_prompt);
  if (include_kv)
    js += fmt(",\"flow_lm_kv\":{\"dtype\":\"F16\",\"shape\":[%d,2,%d,%d,%d],\"data_offsets\":[%zu,%zu]}", N_LAYERS, N_HEADS, T, HD, n_prompt, n_prompt + n_kv);
  js += "}";
  while (js.size() % 8) js.push_back(' ');
  std::vector<__half> kv;
  if (include_kv) {
    PTTS_CUDA(cudaSetDevice(e.cfg.device));
    e.sync_all();
    kv.resize(v->v.kv.n);
    PTTS_CUDA(cudaMemcpy(kv.data(), v->v.kv.p, n_kv, cudaMemcpyDeviceToHost));
  }
  std::ofstream f(path, std::ios::binary);
  PTTS_REQUIRE(f.good(), PTTS_ERR_INVALID, "cannot open '%s' for writing", path);
  const uint64_t hl = js.size();
  f.write(reinterpret_cast<const char*>(&hl), 8);
  f.write(js.data(), (std::streamsize)js.size());
  f.write(reinterpret_cast<const char*>(v->v.prompt.data()), (std::streamsize)n_prompt);
  if (include_kv) f.write(reinterpret_cast<const char*>(kv.data()), (std::streamsize)n_kv);
  PTTS_REQUIRE(f.good(), PTTS_ERR_INVALID, "short write to '%s'", path);
  return PTTS_OK;
  PTTS_CATCH
}

int32_t ptts_voice_load(ptts_engine* h, const char* path, ptts_voice** out) {
  PTTS_TRY
  PTTS_REQUIRE(h && path && out, PTTS_ERR_INVALID, "ptts_voice_load: null argument");
  std::ifstream f(path, std::ios::binary);
  PTTS_REQUIRE(f.good(), PTTS_ERR_INVALID, "cannot open '%s'", path);
  std::vector<unsigned char> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  PTTS_REQUIRE(buf.size() >= 8, PTTS_ERR_INVALID, "'%s' is not a safetensors file", path);
  uint64_t hl;
  std::memcpy(&hl, buf.data(), 8);
  PTTS_REQUIRE(hl <= buf.size() - 8, PTTS_ERR_INVALID, "'%s': header length %llu beyond the file", path, (unsigned long long)hl);
  const auto tensors = st_parse_header(std::string(reinterpret_cast<const char*>(buf.data()) + 8, (size_t)hl));
  const unsigned char* data = buf.data() + 8 + hl;
  const size_t data_len = buf.size() - 8 - hl;
  auto it = tensors.find("audio_prompt");
  PTTS_REQUIRE(it != tensors.end(), PTTS_ERR_INVALID, "'audio_prompt' not found in safetensors file");  // tts_model.rs:474
  const StTensor& tp = it->second;
  PTTS_REQUIRE(tp.end <= data_len, PTTS_ERR_INVALID, "'%s': audio_prompt data beyond the file", path);
  PTTS_REQUIRE(!tp.shape.empty() && tp.shape.back() == D_MODEL, PTTS_ERR_INVALID, "audio_prompt must be [..., %d]", D_MODEL);
  const std::vector<float> prompt = st_to_f32(tp, data);
  const int T = (int)(prompt.size() / D_MODEL);
  auto kvit = tensors.find("flow_lm_kv");
  const std::vector<int64_t> want{N_LAYERS, 2, N_HEADS, T, HD};
  if (kvit != tensors.end() && kvit->second.dtype == "F16" && kvit->second.shape == want && kvit->second.end <= data_len &&
      kvit->second.end - kvit->second.begin == (size_t)N_LAYERS * 2 * N_HEADS * T * HD * 2) {
    // the prefilled KV rows travel with the file: no prefill
    Engine& e = h->e;
    PTTS_REQUIRE(T >= 1 && T <= 1024, PTTS_ERR_CAPACITY, "voice prompt of %d rows (supported: 1..1024)", T);
    PTTS_CUDA(cudaSetDevice(e.cfg.device));
    std::unique_ptr<ptts_voice> v(new ptts_voice);
    v->v.len = T;
    v->v.prompt = prompt;
    v->v.kv.alloc((size_t)N_LAYERS * 2 * N_HEADS * T * HD);
    PTTS_CUDA(cudaMemcpy(v->v.kv.p, data + kvit->second.begin, v->v.kv.n * 2, cudaMemcpyHostToDevice));
    *out = v.release();
    return PTTS_OK;
  }
  return ptts_voice_from_prompt(h, prompt.data(), T, out);
  PTTS_CATCH
}

int32_t ptts_config_check(const char* yaml_path) {
  PTTS_TRY
  PTTS_REQUIRE(yaml_path, PTTS_ERR_INVALID, "ptts_config_check: null path");
  std::ifstream f(yaml_path);
  PTTS_REQUIRE(f.good(), PTTS_ERR_INVALID, "cannot open config '%s'", yaml_path);
  const auto y = yaml_flatten(f);
  auto num = [&](const char* key) {
    auto it = y.find(key);
    PTTS_REQUIRE(it != y.end(), PTTS_ERR_INVALID, "config '%s' has no key %s", yaml_path, key);
    return std::stod(it->second);
  };
  auto want = [&](const char* key, double v) {
    const double got = num(key);
    PTTS_REQUIRE(got == v, PTTS_ERR_INVALID, "config %s = %g, this library is built for %g", key, got, v);
  };
  want("flow_lm.flow.depth", FLOW_DEPTH); want("flow_lm.flow.dim", FLOW_DIM);
  want("flow_lm.transformer.d_model", D_MODEL); want("flow_lm.transformer.hidden_scale", D_FFN / D_MODEL);
  want("flow_lm.transformer.max_period", 10000); want("flow_lm.transformer.num_heads", N_HEADS);
  want("flow_lm.transformer.num_layers", N_LAYERS);
  want("flow_lm.lookup_table.dim", D_MODEL); want("flow_lm.lookup_table.n_bins", N_BINS);
  want("mimi.sample_rate", 24000); want("mimi.channels", 1); want("mimi.frame_rate", 12.5);
  want("mimi.seanet.dimension", MIMI_DIM); want("mimi.seanet.n_filters", 64); want("mimi.seanet.n_residual_layers", 1);
  want("mimi.seanet.ratios.0", 6); want("mimi.seanet.ratios.1", 5); want("mimi.seanet.ratios.2", 4);
  PTTS_REQUIRE(y.find("mimi.seanet.ratios.3") == y.end(), PTTS_ERR_INVALID, "config mimi.seanet.ratios has more than three entries");
  want("mimi.seanet.kernel_size", 7); want("mimi.seanet.residual_kernel_size", 3); want("mimi.seanet.last_kernel_size", 3);
  want("mimi.seanet.compress", 2);
  { auto it = y.find("mimi.seanet.pad_mode");
    PTTS_REQUIRE(it != y.end() && it->second == "constant", PTTS_ERR_INVALID, "config mimi.seanet.pad_mode must be 'constant'"); }
  want("mimi.transformer.d_model", MIMI_DIM); want("mimi.transformer.num_heads", MIMI_HEADS);
  want("mimi.transformer.num_layers", MIMI_LAYERS); want("mimi.transformer.context", MIMI_CTX);
  want("mimi.transformer.dim_feedforward", MIMI_FFN);
  want("mimi.quantizer.dimension", LDIM); want("mimi.quantizer.output_dimension", MIMI_DIM);
  return PTTS_OK;
  PTTS_CATCH
}

int32_t ptts_test_noise(int32_t device, uint64_t seed, int32_t frames, float* out) {
  PTTS_TRY
  PTTS_REQUIRE(out && frames >= 1, PTTS_ERR_INVALID, "ptts_test_noise: bad arguments");
  int ndev = 0;
  PTTS_CUDA(cudaGetDeviceCount(&ndev));
  PTTS_REQUIRE(device >= 0 && device < ndev, PTTS_ERR_CUDA, "CUDA device %d not present", device);
  PTTS_CUDA(cudaSetDevice(device));
  DevBuf<float> d;
  d.alloc((size_t)frames * LDIM);
  noise_probe_kernel<<<(frames * LDIM + 255) / 256, 256>>>((unsigned long long)seed, frames, d.p);
  PTTS_CUDA(cudaGetLastError());
  PTTS_CUDA(cudaMemcpy(out, d.p, d.n * 4, cudaMemcpyDeviceToHost));
  return PTTS_OK;
  PTTS_CATCH
}

int32_t ptts_debug_f16_overflow(ptts_engine* h, int64_t* count_out) {
  PTTS_TRY
  PTTS_REQUIRE(h && count_out, PTTS_ERR_INVALID, "null argument");
  Engine& e = h->e;
  PTTS_CUDA(cudaSetDevice(e.cfg.device));
  e.sync_all();
  DevBuf<unsigned long long> cnt;
  cnt.alloc(1);
  auto scan = [&](const __half* p, size_t n) {
    if (p && n) count_nonfinite_f16_kernel<<<(unsigned)std::min<size_t>(1184, (n + 255) / 256), 256, 0, e.stream>>>(p, (long long)n, cnt.p);
  };
  for (DevBuf<__half>* b : {&e.h16, &e.attn16, &e.ffn16, &e.lat16, &e.y16, &e.fh16, &e.fg16, &e.z16, &e.mh16, &e.mattn16, &e.mffn16, &e.tr16,
                            &e.a0, &e.e2, &e.h3, &e.a3, &e.e5, &e.h6, &e.a6, &e.e8, &e.h9, &e.a9, &e.kv, &e.mimi_ring})
    scan(b->p, b->n);
  scan(reinterpret_cast<const __half*>(e.lm_hA.p), e.lm_hA.n / 2);
  scan(reinterpret_cast<const __half*>(e.lm_attnA.p), e.lm_attnA.n / 2);
  scan(reinterpret_cast<const __half*>(e.lm_ffnA.p), e.lm_ffnA.n / 2);
  PTTS_CUDA(cudaGetLastError());
  unsigned long long c = 0;
  PTTS_CUDA(cudaStreamSynchronize(e.stream));
  PTTS_CUDA(cudaMemcpy(&c, cnt.p, 8, cudaMemcpyDeviceToHost));
  *count_out = (int64_t)c;
  return PTTS_OK;
  PTTS_CATCH
}

// ------------------------------------------------------------------------------------------------ native scheduler
struct ptts_sched {
  struct Seg { int kind = 0; std::vector<int32_t> tokens; ptts_stream_params params{}; std::vector<float> noise; int pause_ms = 0; };
  struct Req { std::vector<Seg> segs; size_t cursor = 0; std::vector<float> out32; std::vector<int16_t> out16; };
  ptts_engine* eng = nullptr;
  ptts_voice* voice = nullptr;
  int max_batch = 0;
  std::vector<Req> reqs;
  long long steps = 0;
};

int32_t ptts_sched_create(ptts_engine* h, ptts_voice* voice, int32_t max_batch, ptts_sched** out) {
  PTTS_TRY
  PTTS_REQUIRE(h && voice && out, PTTS_ERR_INVALID, "ptts_sched_create: null argument");
  std::unique_ptr<ptts_sched> s(new ptts_sched);
  s->eng = h; s->voice = voice;
  s->max_batch = max_batch > 0 ? std::min(max_batch, h->e.NB) : h->e.NB;
  *out = s.release();
  return PTTS_OK;
  PTTS_CATCH
}

void ptts_sched_destroy(ptts_sched* s) { delete s; }

int64_t ptts_sched_submit(ptts_sched* s, const ptts_segment* segs, int32_t n) {
  try {
    PTTS_REQUIRE(s && segs && n >= 1, PTTS_ERR_INVALID, "ptts_sched_submit: null argument");
    ptts_sched::Req r;
    for (int i = 0; i < n; ++i) {
      ptts_sched::Seg g;
      g.kind = segs[i].kind;
      if (g.kind == PTTS_SEG_TEXT) {
        PTTS_REQUIRE(segs[i].n_tokens >= 0 && (segs[i].tokens || segs[i].n_tokens == 0), PTTS_ERR_INVALID, "segment %d: bad tokens", i);
        PTTS_REQUIRE(segs[i].params.max_gen_len >= 1, PTTS_ERR_INVALID, "segment %d: max_gen_len %d", i, segs[i].params.max_gen_len);
        g.tokens.assign(segs[i].tokens, segs[i].tokens + segs[i].n_tokens);
        g.params = segs[i].params;
        if (segs[i].params.noise) g.noise.assign(segs[i].params.noise, segs[i].params.noise + (size_t)segs[i].params.max_gen_len * LDIM);
      } else {
        PTTS_REQUIRE(g.kind == PTTS_SEG_PAUSE && segs[i].pause_ms >= 0, PTTS_ERR_INVALID, "segment %d: unknown kind / negative pause", i);
        g.pause_ms = segs[i].pause_ms;
      }
      r.segs.push_back(std::move(g));
    }
    s->reqs.push_back(std::move(r));
    return (int64_t)s->reqs.size() - 1;
  } catch (const ptts::Error& ex) { g_last_error = ex.what(); return ex.code; }
  catch (const std::exception& ex) { g_last_error = ex.what(); return PTTS_ERR_INVALID; }
}

// The loop of tts_model.BatchScheduler (ahead form) in C++: one step is kept enqueued ahead of the flags of the current
// one whenever no row can reach its max_gen_len on the current step; a row that ends at EOS instead comes back from the
// step enqueued ahead as an overrun row (dropped) and its slot is closed once that step has been drained.
int32_t ptts_sched_run(ptts_sched* s, int32_t pcm_i16) {
  PTTS_TRY
  PTTS_REQUIRE(s, PTTS_ERR_INVALID, "null scheduler");
  Engine& e = s->eng->e;
  PTTS_CUDA(cudaSetDevice(e.cfg.device));
  const int pcm_flag = pcm_i16 ? PTTS_STEP_PCM_I16 : PTTS_STEP_PCM;
  const size_t NR = s->reqs.size();
  for (auto& r : s->reqs) {
    r.cursor = 0; r.out32.clear(); r.out16.clear();
    size_t cap = 0;  // upper bound of the request's samples: no reallocation (a 60 s request is 2.9 MB of i16) in the loop
    for (auto& g : r.segs) cap += g.kind == PTTS_SEG_PAUSE ? (size_t)g.pause_ms * 24 : (size_t)g.params.max_gen_len * FRAME;
    if (pcm_i16) r.out16.reserve(cap); else r.out32.reserve(cap);
  }
  s->steps = 0;
  std::vector<int> waiting(NR);
  for (size_t i = 0; i < NR; ++i) waiting[i] = (int)i;
  std::vector<int> owner(e.NS, -1), budget(e.NS, 0);   // slot -> request, frames the stream may still begin
  std::vector<int> active;                              // slots, in batch row order
  auto silence = [&](ptts_sched::Req& r, int ms) {
    const size_t n = (size_t)ms * 24;                   // pause.rs:183-185 at 24 kHz
    if (pcm_i16) r.out16.insert(r.out16.end(), n, 0); else r.out32.insert(r.out32.end(), n, 0.f);
  };
  auto admit = [&]() {
    std::vector<int> still, owners;
    std::vector<ptts_voice*> voices;
    std::vector<int32_t> toks, offs{0};
    std::vector<ptts_stream_params> params;
    for (int ri : waiting) {
      ptts_sched::Req& r = s->reqs[ri];
      while (r.cursor < r.segs.size() && r.segs[r.cursor].kind == PTTS_SEG_PAUSE) silence(r, r.segs[r.cursor++].pause_ms);
      if (r.cursor >= r.segs.size()) continue;
      if ((int)(active.size() + owners.size()) < s->max_batch) {
        ptts_sched::Seg& g = r.segs[r.cursor++];
        toks.insert(toks.end(), g.tokens.begin(), g.tokens.end());
        offs.push_back((int32_t)toks.size());
        ptts_stream_params p = g.params;
        p.noise = g.noise.empty() ? nullptr : g.noise.data();
        params.push_back(p);
        voices.push_back(s->voice);
        owners.push_back(ri);
      } else {
        still.push_back(ri);
      }
    }
    waiting.swap(still);
    if (owners.empty()) return;
    std::vector<int32_t> slots(owners.size());
    if (toks.empty()) toks.push_back(0);
    streams_open_impl(e, (int)owners.size(), voices.data(), toks.data(), offs.data(), params.data(), slots.data());
    for (size_t i = 0; i < owners.size(); ++i) { owner[slots[i]] = owners[i]; budget[slots[i]] = params[i].max_gen_len; active.push_back(slots[i]); }
  };
  struct Flight { long long ticket; std::vector<int> slots; };
  std::vector<Flight> inflight;
  std::vector<uint8_t> fin;
  auto begin = [&](const std::vector<int>& slots, bool ahead) {
    const long long t = step_begin_impl(e, slots.data(), (int)slots.size(), pcm_flag | (ahead ? PTTS_STEP_AHEAD : 0));
    for (int sl : slots) --budget[sl];
    ++s->steps;
    return t;
  };
  // fetches flags + PCM of the oldest step in flight, appends its frames, returns the slots that finished on it
  auto drain = [&](std::vector<int>& done) {
    Flight f = std::move(inflight.front());
    inflight.erase(inflight.begin());
    const int n = (int)f.slots.size();
    fin.assign(n, 0);
    step_flags_impl(e, f.ticket, fin.data(), nullptr, nullptr);
    // rows go straight from the pinned landing buffer of the step to their request (no intermediate frame copy)
    step_pcm_impl(e, f.ticket, nullptr);
    const int par = (int)(f.ticket % Engine::NT);
    const int16_t* src16 = e.pin_pcm16[par];
    const float* src32 = e.pin_pcm[par];
    for (int i = 0; i < n; ++i) {
      if (fin[i] == PTTS_FRAME_OVERRUN) continue;
      ptts_sched::Req& r = s->reqs[owner[f.slots[i]]];
      if (pcm_i16) r.out16.insert(r.out16.end(), src16 + (size_t)i * FRAME, src16 + (size_t)(i + 1) * FRAME);
      else r.out32.insert(r.out32.end(), src32 + (size_t)i * FRAME, src32 + (size_t)(i + 1) * FRAME);
      if (fin[i] && std::find(done.begin(), done.end(), f.slots[i]) == done.end()) done.push_back(f.slots[i]);
    }
  };
  try {
    admit();
    while (!active.empty() || !inflight.empty()) {
      if (inflight.empty()) inflight.push_back(Flight{begin(active, false), active});
      if (inflight.size() == 1) {
        bool room = true;
        for (int sl : inflight[0].slots) room = room && budget[sl] > 0 && e.slots[sl].own_len + 2 < e.KVCAP;
        if (room) inflight.push_back(Flight{begin(inflight[0].slots, true), inflight[0].slots});
      }
      std::vector<int> done;
      drain(done);
      if (!done.empty()) {
        while (!inflight.empty()) drain(done);   // the step enqueued ahead still lists the finished slots
        for (int sl : done) {
          stream_close_impl(e, sl);
          waiting.push_back(owner[sl]);
          owner[sl] = -1;
          active.erase(std::find(active.begin(), active.end(), sl));
        }
        admit();
      }
    }
  } catch (...) {
    // leave nothing in flight and no slot open on the shared engine
    for (auto& f : inflight) {
      try { step_flags_impl(e, f.ticket, nullptr, nullptr, nullptr); } catch (...) {}
      try { step_pcm_impl(e, f.ticket, nullptr); } catch (...) {}
    }
    for (int sl : active) { try { stream_close_impl(e, sl); } catch (...) {} }
    throw;
  }
  return PTTS_OK;
  PTTS_CATCH
}

int64_t ptts_sched_result_samples(const ptts_sched* s, int64_t req) {
  if (!s || req < 0 || req >= (int64_t)s->reqs.size()) return PTTS_ERR_INVALID;
  const auto& r = s->reqs[(size_t)req];
  return (int64_t)std::max(r.out32.size(), r.out16.size());
}

int32_t ptts_sched_result(const ptts_sched* s, int64_t req, void* pcm_out, int64_t cap) {
  PTTS_TRY
  PTTS_REQUIRE(s && pcm_out && req >= 0 && req < (int64_t)s->reqs.size(), PTTS_ERR_INVALID, "ptts_sched_result: bad arguments");
  const auto& r = s->reqs[(size_t)req];
  const int64_t n = (int64_t)std::max(r.out32.size(), r.out16.size());
  PTTS_REQUIRE(cap >= n, PTTS_ERR_INVALID, "request %lld has %lld samples, buffer holds %lld", (long long)req, (long long)n, (long long)cap);
  if (!r.out16.empty()) std::memcpy(pcm_out, r.out16.data(), r.out16.size() * 2);
  else if (!r.out32.empty()) std::memcpy(pcm_out, r.out32.data(), r.out32.size() * 4);
  return PTTS_OK;
  PTTS_CATCH
}

int32_t ptts_sched_result_view(const ptts_sched* s, int64_t req, const void** data, int64_t* n_samples) {
  PTTS_TRY
  PTTS_REQUIRE(s && data && n_samples && req >= 0 && req < (int64_t)s->reqs.size(), PTTS_ERR_INVALID, "ptts_sched_result_view: bad arguments");
  const auto& r = s->reqs[(size_t)req];
  if (!r.out16.empty()) { *data = r.out16.data(); *n_samples = (int64_t)r.out16.size(); }
  else { *data = r.out32.empty() ? nullptr : (const void*)r.out32.data(); *n_samples = (int64_t)r.out32.size(); }
  return PTTS_OK;
  PTTS_CATCH
}

int64_t ptts_sched_steps(const ptts_sched* s) { return s ? s->steps : -1; }

// ------------------------------------------------------------------------------------------------ isolated kernel tests
struct TestCtx {
  Engine e;
  explicit TestCtx(int device, int use_simt) {
    int ndev = 0;
    PTTS_CUDA(cudaGetDeviceCount(&ndev));
    PTTS_REQUIRE(device >= 0 && device < ndev, PTTS_ERR_CUDA, "CUDA device %d not present", device);
    PTTS_CUDA(cudaSetDevice(device));
    e.cfg.device = device;
    e.cfg.debug_gemm = use_simt;
    PTTS_CUDA(cudaStreamCreateWithFlags(&e.stream, cudaStreamNonBlocking));
    e.ls = e.stream;
    PTTS_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  PTTS_CUDA(cudaFuncSetAttribute(gemm_tc_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
};

static void to_f16_dev(DevBuf<__half>& dst, const float* src, size_t n) {
  std::vector<__half> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = __float2half_rn(src[i]);
  dst.alloc(n);
  PTTS_CUDA(cudaMemcpy(dst.p, h.data(), n * sizeof(__half), cudaMemcpyHostToDevice));
}

int32_t ptts_test_gemm(int32_t device, const float* a, const float* w, const float* bias, float* d, int32_t rows,
                       int32_t feats, int32_t k, int32_t mode, int32_t split_k, int32_t act, int32_t use_simt) {
  PTTS_TRY
  PTTS_REQUIRE(a && w && d && rows > 0 && feats > 0 && k > 0 && k % 64 == 0, PTTS_ERR_INVALID, "bad test_gemm arguments");
  TestCtx t(device, use_simt);
  Engine& e = t.e;
  e.cfg.reserved[0] = mode;
  DevBuf<__half> a16;
  to_f16_dev(a16, a, (size_t)rows * k);
  Weight16 w16;
  upload_f16(w16, std::vector<float>(w, w + (size_t)feats * k), feats, k);
  DevBuf<float> out, bd;
  out.alloc((size_t)rows * feats);
  if (bias) { bd.alloc(feats); PTTS_CUDA(cudaMemcpy(bd.p, bias, feats * 4, cudaMemcpyHostToDevice)); }
  GemmEpi ep = epi_none();
  ep.bias = bias ? bd.p : nullptr; ep.act = act; ep.out32 = out.p; ep.out32_map = plain_map(feats);
  e.cfg.reserved[2] = split_k;
  e.gemm_rows(a16.p, rows, k, w16, feats, ep, split_k > 1);
  PTTS_CUDA(cudaStreamSynchronize(e.stream));
  PTTS_CUDA(cudaMemcpy(d, out.p, (size_t)rows * feats * 4, cudaMemcpyDeviceToHost));
  return PTTS_OK;
  PTTS_CATCH
}

int64_t ptts_test_gemv_launches(void) { return g_gemv_launches.load(); }

int32_t ptts_test_gemm_int8(int32_t device, const float* a, const float* w, float* d, int32_t rows, int32_t feats,
                            int32_t k, int32_t split_k, int32_t storage, float* scale_out) {
  PTTS_TRY
  PTTS_REQUIRE(a && w && d && rows > 0 && rows <= 256 && feats > 0 && k > 0 && k % 64 == 0, PTTS_ERR_INVALID, "bad test_gemm_int8 arguments");
  TestCtx t(device, 0);
  Engine& e = t.e;
  e.cfg.reserved[0] = storage == 2 ? 0 : 2;   // weights on MMA-M: the only tensor-core placement with in-kernel int8 expansion;
                                              // storage 2 = byte codes with the library's own choice (1-4 rows: the GEMV of gemv.cuh)
  e.cfg.reserved[7] = storage ? 0 : 1;
  DevBuf<__half> a16;
  to_f16_dev(a16, a, (size_t)rows * k);
  float amax = 0.f;
  for (size_t i = 0; i < (size_t)feats * k; ++i) amax = std::max(amax, std::fabs(w[i]));
  const float scale = amax / 127.f;
  PTTS_REQUIRE(scale > 0.f, PTTS_ERR_INVALID, "all-zero weight");
  Weight16 w16;
  upload_f16(w16, std::vector<float>(w, w + (size_t)feats * k), feats, k, std::vector<float>(feats, scale));
  PTTS_REQUIRE(w16.q8.p != nullptr, PTTS_ERR_STATE, "int8 codes were not built");
  DevBuf<float> out;
  out.alloc((size_t)rows * feats);
  GemmEpi ep = epi_none();
  ep.out32 = out.p; ep.out32_map = plain_map(feats);
  e.cfg.reserved[2] = split_k;
  e.gemm_rows(a16.p, rows, k, w16, feats, ep, split_k > 1);
  PTTS_CUDA(cudaStreamSynchronize(e.stream));
  PTTS_CUDA(cudaMemcpy(d, out.p, (size_t)rows * feats * 4, cudaMemcpyDeviceToHost));
  if (scale_out) *scale_out = scale;
  return PTTS_OK;
  PTTS_CATCH
}

// Bring-up probe: `iters` back-to-back launches of one GEMM; returns the CUDA-event time per launch and, for the
// last launch, the %globaltimer stamps of up to `max_ctas` CTAs relative to the first CTA's entry (ns):
// [entry, setup done, first TMA issued, producer done, first tile landed, MMAs issued, accumulator ready,
//  tile staged, epilogue done, TMEM freed].
int32_t ptts_test_gemm_trace(int32_t device, int32_t rows, int32_t feats, int32_t k, int32_t mode, int32_t split_k,
                             int32_t iters, float* us_per_launch, int64_t* stamps, int32_t max_ctas, int32_t* n_ctas) {
  PTTS_TRY
  TestCtx t(device, 0);
  Engine& e = t.e;
  e.cfg.reserved[0] = mode;
  DevBuf<__half> a16; a16.alloc((size_t)rows * k);
  Weight16 w16; w16.F = feats; w16.K = k; w16.Fpad = round_up(feats, 128); w16.w.alloc((size_t)w16.Fpad * k);
  DevBuf<float> out; out.alloc((size_t)rows * feats);
  DevBuf<unsigned long long> tr; tr.alloc(16 * 65536);
  DevBuf<__half> out16; out16.alloc((size_t)rows * feats);
  DevBuf<float> resb; resb.alloc((size_t)rows * feats);
  GemmEpi ep = epi_none();
  ep.out32 = out.p; ep.out32_map = plain_map(feats);
  if (mode >= 10) {  // SEANet-style epilogues: 10 = bias-free f32 + f16(ELU) dual output, 11 = residual in, f16(ELU) out
    if (mode == 11) { ep.out32 = nullptr; ep.res = resb.p; ep.res_map = plain_map(feats); }
    ep.out16 = out16.p; ep.out16_map = plain_map(feats); ep.act16 = ACT_ELU;
    ep.bias = resb.p;  // any f32 vector of >= feats elements
    mode = 1;
  }
  e.cfg.reserved[0] = mode;
  e.cfg.reserved[2] = split_k;
  e.gemm_trace = tr.p;
  for (int i = 0; i < 3; ++i) e.gemm_rows(a16.p, rows, k, w16, feats, ep, split_k > 1);
  cudaEvent_t a, b;
  PTTS_CUDA(cudaEventCreate(&a)); PTTS_CUDA(cudaEventCreate(&b));
  PTTS_CUDA(cudaEventRecord(a, e.stream));
  for (int i = 0; i < iters; ++i) e.gemm_rows(a16.p, rows, k, w16, feats, ep, split_k > 1);
  PTTS_CUDA(cudaEventRecord(b, e.stream));
  PTTS_CUDA(cudaStreamSynchronize(e.stream));
  float ms = 0;
  PTTS_CUDA(cudaEventElapsedTime(&ms, a, b));
  *us_per_launch = 1000.f * ms / iters;
  std::vector<unsigned long long> h(16 * (size_t)max_ctas);
  PTTS_CUDA(cudaMemcpy(h.data(), tr.p, h.size() * 8, cudaMemcpyDeviceToHost));
  unsigned long long t0 = ~0ull;
  int n = 0;
  for (int c = 0; c < max_ctas; ++c) if (h[c * 16] != 0 && h[c * 16 + 12] == 0) { t0 = std::min(t0, h[c * 16]); n = c + 1; }
  // persistent kernels stamp slot 12 first (entry of the epilogue warps)
  if (t0 == ~0ull) { for (int c = 0; c < max_ctas; ++c) if (h[c * 16 + 12] != 0) { t0 = std::min(t0, h[c * 16 + 12]); n = c + 1; } }
  for (int c = 0; c < n; ++c) for (int j = 0; j < 16; ++j) stamps[c * 16 + j] = h[c * 16 + j] ? (int64_t)(h[c * 16 + j] - t0) : -1;
  *n_ctas = n;
  cudaEventDestroy(a); cudaEventDestroy(b);
  return PTTS_OK;
  PTTS_CATCH
}

static void pick_tile(int T, int* R, int* G) {
  if (T <= 128) { *R = T; *G = std::max(1, 128 / T); }
  else if (T % 128 == 0) { *R = 128; *G = 1; }
  else if (T % 120 == 0) { *R = 120; *G = 1; }
  else { *R = 128; *G = 1; }
}

int32_t ptts_test_conv1d(int32_t device, const float* x, const float* prev, const float* w, const float* bias, float* y,
                         int32_t n, int32_t t, int32_t cin, int32_t cout, int32_t k) {
  PTTS_TRY
  PTTS_REQUIRE(x && w && bias && y && n > 0 && t > 0 && cin % 64 == 0 && k >= 1, PTTS_ERR_INVALID, "bad test_conv1d arguments");
  TestCtx tc(device, 0);
  Engine& e = tc.e;
  const int pad = k - 1, tp = pad + t;
  std::vector<float> xp((size_t)n * tp * cin, 0.f);
  for (int b = 0; b < n; ++b) {
    if (prev && pad) std::memcpy(&xp[(size_t)b * tp * cin], prev + (size_t)b * pad * cin, (size_t)pad * cin * 4);
    std::memcpy(&xp[((size_t)b * tp + pad) * cin], x + (size_t)b * t * cin, (size_t)t * cin * 4);
  }
  DevBuf<__half> x16;
  to_f16_dev(x16, xp.data(), xp.size());
  HostTensor hw, hb;
  hw.f32.assign(w, w + (size_t)cout * cin * k);
  hb.f32.assign(bias, bias + cout);
  Weight16 wg; DevBuf<float> bg;
  conv_weight(wg, bg, hw, hb, cout, cin, k, cout, cin);
  DevBuf<float> out;
  out.alloc((size_t)n * t * cout);
  GemmEpi ep = epi_none();
  ep.bias = bg.p; ep.out32 = out.p; ep.out32_map = plain_map(cout);
  int R, G;
  pick_tile(t, &R, &G);
  e.gemm(ActView{x16.p, cin, tp, n}, n, t, k, R, G, wg, cout, ep);
  PTTS_CUDA(cudaStreamSynchronize(e.stream));
  PTTS_CUDA(cudaMemcpy(y, out.p, out.n * 4, cudaMemcpyDeviceToHost));
  return PTTS_OK;
  PTTS_CATCH
}

int32_t ptts_test_convtr1d(int32_t device, const float* x, const float* prev_row, const float* w, const float* bias, float* y,
                           int32_t n, int32_t t, int32_t cin, int32_t cout, int32_t stride) {
  PTTS_TRY
  PTTS_REQUIRE(x && w && bias && y && n > 0 && t > 0 && cin % 64 == 0 && stride >= 1, PTTS_ERR_INVALID, "bad test_convtr1d arguments");
  TestCtx tc(device, 0);
  Engine& e = tc.e;
  const int tp = 1 + t;
  std::vector<float> xp((size_t)n * tp * cin, 0.f);
  for (int b = 0; b < n; ++b) {
    if (prev_row) std::memcpy(&xp[(size_t)b * tp * cin], prev_row + (size_t)b * cin, (size_t)cin * 4);
    std::memcpy(&xp[((size_t)b * tp + 1) * cin], x + (size_t)b * t * cin, (size_t)t * cin * 4);
  }
  DevBuf<__half> x16;
  to_f16_dev(x16, xp.data(), xp.size());
  HostTensor hw, hb;
  hw.f32.assign(w, w + (size_t)cin * cout * 2 * stride);
  hb.f32.assign(bias, bias + cout);
  Weight16 wg; DevBuf<float> bg;
  convtr_weight(wg, bg, hw, hb, cin, cout, stride);
  DevBuf<float> out;
  out.alloc((size_t)n * t * stride * cout);
  GemmEpi ep = epi_none();
  ep.bias = bg.p; ep.out32 = out.p;
  ep.out32_map = RowMap{t, stride * cout, (long long)t * stride * cout, 0};
  int R, G;
  pick_tile(t, &R, &G);
  e.gemm(ActView{x16.p, cin, tp, n}, n, t, 2, R, G, wg, stride * cout, ep);
  PTTS_CUDA(cudaStreamSynchronize(e.stream));
  PTTS_CUDA(cudaMemcpy(y, out.p, out.n * 4, cudaMemcpyDeviceToHost));
  return PTTS_OK;
  PTTS_CATCH
}

}  // extern "C"
