"""GPU bring-up report (not a test): runs the isolated kernels and the full path against the oracle
and prints one line per stage so a single gpurun call localises a fault.  Writes gpurun_out/debug.log."""
import os, sys, time, traceback
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from pocket_tts_b200 import synth
from pocket_tts_b200.engine import Engine, StreamSpec, test_gemm, test_conv1d, test_convtr1d
from oracle import ptts_oracle as O

out = ROOT / "gpurun_out"; out.mkdir(exist_ok=True)
logf = open(out / "debug.log", "w")
def log(*a):
    s = " ".join(str(x) for x in a); print(s, flush=True); logf.write(s + "\n"); logf.flush()

def f16r(a): return np.asarray(a, np.float32).astype(np.float16).astype(np.float32)

def gemm_checks():
    rng = np.random.default_rng(0)
    for rows, feats, k, mode, split in [(64, 128, 64, 2, 1), (64, 128, 64, 1, 1), (128, 128, 128, 1, 1), (64, 3072, 1024, 0, 1),
                                         (1024, 1536, 512, 0, 1), (64, 1024, 4096, 0, 8), (300, 640, 512, 1, 1), (1, 1024, 1024, 0, 1)]:
        a = rng.standard_normal((rows, k), dtype=np.float32); w = rng.standard_normal((feats, k), dtype=np.float32) / np.sqrt(k)
        want = f16r(a).astype(np.float64) @ f16r(w).astype(np.float64).T
        for simt in (1, 0):
            try:
                t0 = time.time(); got = test_gemm(a, w, None, mode=mode, split_k=split, use_simt=simt)
                err = np.abs(got - want).max()
                log(f"gemm rows={rows} F={feats} K={k} mode={mode} split={split} simt={simt}: maxerr {err:.3e} ({time.time()-t0:.2f}s)")
                if err > 1e-2 and not simt:
                    bad = np.argwhere(np.abs(got - want) > 1e-2)
                    log("   bad count", len(bad), "first", bad[:5].tolist(), "got", got[tuple(bad[0])], "want", want[tuple(bad[0])])
                    log("   row-err profile (first 16 rows):", np.abs(got - want).max(axis=1)[:16].round(3).tolist())
                    log("   col-err profile (first 16 cols):", np.abs(got - want).max(axis=0)[:16].round(3).tolist())
            except Exception as e:
                log(f"gemm rows={rows} F={feats} K={k} mode={mode} simt={simt}: EXC {e}")

def conv_checks():
    import torch.nn.functional as F
    rng = np.random.default_rng(1)
    for n, t, cin, cout, k in [(3, 16, 512, 512, 7), (2, 96, 256, 128, 3), (2, 480, 128, 64, 3), (1, 1920, 64, 64, 3)]:
        x = rng.standard_normal((n, t, cin), dtype=np.float32); prev = rng.standard_normal((n, k - 1, cin), dtype=np.float32)
        w = rng.standard_normal((cout, cin, k), dtype=np.float32) / np.sqrt(cin * k); b = rng.standard_normal(cout, dtype=np.float32)
        try:
            got = test_conv1d(x, prev, w, b)
            xp = torch.from_numpy(f16r(np.concatenate([prev, x], 1))).permute(0, 2, 1).double()
            want = F.conv1d(xp, torch.from_numpy(f16r(w)).double(), torch.from_numpy(b).double()).permute(0, 2, 1).numpy()
            log(f"conv1d n={n} t={t} cin={cin} cout={cout} k={k}: maxerr {np.abs(got-want).max():.3e}")
        except Exception as e:
            log(f"conv1d t={t}: EXC {e}")
    for n, t, cin, cout, s in [(3, 16, 512, 256, 6), (2, 96, 256, 128, 5), (2, 480, 128, 64, 4)]:
        x = rng.standard_normal((n, t, cin), dtype=np.float32); prev = rng.standard_normal((n, cin), dtype=np.float32)
        w = rng.standard_normal((cin, cout, 2 * s), dtype=np.float32) / np.sqrt(2 * cin); b = rng.standard_normal(cout, dtype=np.float32)
        try:
            got = test_convtr1d(x, prev, w, b, s)
            xin = torch.from_numpy(f16r(np.concatenate([prev[:, None], x], 1))).permute(0, 2, 1).double()
            full = F.conv_transpose1d(xin, torch.from_numpy(f16r(w)).double(), torch.from_numpy(b).double(), stride=s)
            want = full[:, :, s:s + t * s].permute(0, 2, 1).numpy()
            log(f"convtr1d n={n} t={t} cin={cin} cout={cout} s={s}: maxerr {np.abs(got-want).max():.3e}")
        except Exception as e:
            log(f"convtr1d t={t}: EXC {e}")

def snr(ref, x): return 10 * np.log10((ref ** 2).sum() / max(((ref - x) ** 2).sum(), 1e-30))

def full_path(debug_gemm):
    g = np.load(ROOT / "tests/golden/cfg1_lsd1.npz")
    Wnp = synth.make_weights(int(g["weight_seed"]), layer_scale=float(g["layer_scale"]))
    W = O.to_torch(Wnp)
    prompt = synth.make_voice_prompt(int(g["voice_rows"]), seed=7)
    frames = g["tanh_latents"].shape[0]
    t0 = time.time()
    eng = Engine(Wnp, max_slots=4, kv_capacity=256, debug_gemm=debug_gemm)
    log(f"[debug_gemm={debug_gemm}] engine created in {time.time()-t0:.1f}s")
    voice = eng.voice_from_prompt(prompt)
    # oracle with traces for frame 0
    ov = O.voice_state_from_prompt(W, prompt)
    st = ov.clone(); O.flowlm_prefill(W, O.embed_tokens(W, g["tokens"]), st)
    te = O.compute_time_embeddings(W, 1)
    tr = {}
    lat0, logit0 = O.flowlm_step(W, W["flow_lm.bos_emb"].clone(), st, torch.from_numpy(g["noise"][0]), te, "tanh", tr)
    ms = O.MimiState(); pcm0 = O.mimi_decode_step(W, lat0, ms, "tanh", tr)
    for mode, teacher in (("teacher", True), ("free", False)):
        slots = eng.open_streams([voice], [StreamSpec(tokens=g["tokens"], max_gen_len=frames, eos_threshold=1e30, noise=g["noise"])])
        lat_e, pcm_e, log_e = [], [], []
        for f in range(frames):
            if teacher and f > 0: eng.set_feedback(int(slots[0]), g["tanh_latents"][f - 1])
            pcm, fin, lat, logit = eng.step(slots)
            lat_e.append(lat[0]); pcm_e.append(pcm[0]); log_e.append(logit[0])
            if f == 0 and teacher:
                def cmp(name, got, want):
                    want = np.asarray(want, np.float32); got = np.asarray(got, np.float32).reshape(want.shape)
                    log(f"   tap {name:34s} maxerr {np.abs(got-want).max():.3e}  (ref absmax {np.abs(want).max():.3f})")
                cmp("flowlm.h", eng.debug_read("flowlm.h", 0), tr["flowlm.h"].numpy())
                cmp("flowlm.x(l5.out)", eng.debug_read("flowlm.x", 0), tr["flow_lm.transformer.l5.out"].numpy()[-1])
                cmp("latent", lat[0], lat0.numpy())
                cmp("mimi.quantized", eng.debug_read("mimi.quantized", 0), tr["mimi.quantized"].numpy())
                cmp("mimi.after_decoder_transformer", eng.debug_read("mimi.after_decoder_transformer", 0), tr["mimi.after_decoder_transformer"].numpy().T)
                cmp("seanet.convtr2", eng.debug_read("seanet.convtr2", 0), tr["seanet.convtr2"].numpy().T)
                cmp("seanet.convtr5", eng.debug_read("seanet.convtr5", 0), tr["seanet.convtr5"].numpy().T)
                cmp("seanet.convtr8", eng.debug_read("seanet.convtr8", 0), tr["seanet.convtr8"].numpy().T)
                cmp("pcm", pcm[0], pcm0.numpy())
        lat_e, pcm_e, log_e = np.stack(lat_e), np.stack(pcm_e), np.array(log_e)
        le = np.abs(lat_e - g["tanh_latents"]).max(axis=1)
        log(f"[debug_gemm={debug_gemm}] {mode}: latent maxerr per frame {le.round(4).tolist()}")
        log(f"[debug_gemm={debug_gemm}] {mode}: pcm SNR {snr(g['tanh_pcm'], pcm_e):.1f} dB; per-frame {[round(float(snr(g['tanh_pcm'][i], pcm_e[i])),1) for i in range(frames)]}")
        log(f"[debug_gemm={debug_gemm}] {mode}: eos logit maxerr {np.abs(log_e - g['tanh_eos_logits']).max():.3e}")
        eng.close_stream(int(slots[0]))
    log(f"[debug_gemm={debug_gemm}] launches per step ~ {eng.launch_count()}")
    eng.close()

if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "conv", "simt", "tc"]
    for name, fn in (("gemm", gemm_checks), ("conv", conv_checks), ("simt", lambda: full_path(1)), ("tc", lambda: full_path(0))):
        if name in which:
            try:
                fn()
            except Exception:
                log(f"{name}: EXCEPTION\n{traceback.format_exc()}")
