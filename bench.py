#!/usr/bin/env python
"""Benchmark of the Pocket TTS generation hot path on B200 (contract: see the task brief).

Metric (BASELINE.json): audio-seconds generated per wall-second.
Workload at every N: BASELINE.json configs[1] per GPU -- 64 concurrent ~10 s utterances
(40 text tokens, 87-row voice prompt, 125 frames each, LSD 1, temp 0.7, EOS disabled so every stream
runs its 125 frames), random-init b6369a24 weights, synthetic tokens.  Requests are sharded by GPU with
no data-path collective (weak scaling: 64 streams per GPU).

One "step" = one whole batch job: open 64 streams (embedding + text prefill) and run 125 decode steps
(FlowLM step -> LSD flow head -> Mimi decode) for all of them.
  value : device-resident run (ptts_step_device: PCM stays in HBM, no host sync inside the job)
  e2e   : the same job through the public host calls (ptts_streams_open, ptts_step_begin/flags/pcm) with HOST
          buffers: token upload at open; PCM + finished flags + latents copied back every frame, the flags awaited
          before the next frame is issued.
  --impl reference : the reference's CPU path (oracle port of the Candle implementation; the Rust crate
          cannot be built here) on the box's host cores, one single-thread stream per core.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FRAME_SEC = 0.08
STREAMS, FRAMES, TOKENS, VOICE_ROWS = 64, 125, 40, 87
METRIC = "audio-sec generated per wall-sec"
UNIT = "audio_s/s"
CONFIG = {"workload": "configs[1]: b6369a24 f16-operand batch 64 concurrent 10 s utterances (125 frames, 40 tokens, "
                      "87-row voice), generate_stream, per GPU",
          "streams_per_gpu": STREAMS, "frames_per_stream": FRAMES, "tokens_per_stream": TOKENS,
          "voice_rows": VOICE_ROWS, "lsd_decode_steps": 1, "temp": 0.7, "parallelism": "request-sharded, no collective",
          "l2": "no flush: per-step working set (190 MB weights + ~470 MB KV) exceeds the 126 MB L2"}


def ncu_traffic() -> tuple[dict[str, float], str | None]:
    """dram__bytes_read.sum + dram__bytes_write.sum per launch and kernel function, read at run time from the committed
    summary of this round's `ncu --set full` capture (profiles/ncu_traffic.json, written by tests/ncu_summary.py from the
    raw export) -- not a constant in this file."""
    p = ROOT / "profiles" / "ncu_traffic.json"
    if not p.exists():
        return {}, None
    d = json.loads(p.read_text())
    return {k: float(v) for k, v in d.get("bytes_per_launch", {}).items()}, d.get("source")


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=float(d["hbm_gbs"]), tf=float(d["bf16_tflops_sustained"]), src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf=1590.0, src="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------ CPU reference arm
def _cpu_worker(args):
    """One reference stream on one core: the oracle (CPU restatement of the Candle path), f32, 1 thread."""
    idx, frames, tokens_n = args
    import torch
    torch.set_num_threads(1)
    from oracle import ptts_oracle as O
    from pocket_tts_b200 import synth
    W = O.to_torch(synth.make_weights(1234))
    voice = O.voice_state_from_prompt(W, synth.make_voice_prompt(VOICE_ROWS, seed=7))
    tok = synth.make_tokens(tokens_n, seed=1000 + idx)
    noise = synth.make_noise(frames, seed=2000 + idx)
    t0 = time.perf_counter()
    O.generate_segment(W, voice, tok, noise, frames, 0, float("inf"))  # text prefill + frames
    return time.perf_counter() - t0


class CpuPool:
    def __init__(self, procs):
        import multiprocessing as mp
        self.procs = procs
        self.pool = mp.get_context("spawn").Pool(procs)

    def run(self, frames):
        t0 = time.perf_counter()
        inner = self.pool.map(_cpu_worker, [(i, frames, TOKENS) for i in range(self.procs)])
        wall = time.perf_counter() - t0
        return wall, inner

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_sample(procs, frames, repeats=1):
    """Returns audio-s per wall-s of `procs` single-thread streams generating `frames` frames each.
    Wall time is the slowest worker's generate time (weight synthesis excluded)."""
    pool = CpuPool(procs)
    vals = []
    try:
        for _ in range(repeats):
            _, inner = pool.run(frames)
            vals.append(procs * frames * FRAME_SEC / max(inner))
    finally:
        pool.close()
    return vals


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 32))
    frames = FRAMES  # bounded sample: the full 125-frame utterance, but only one stream per core instead of 64
    pool = CpuPool(procs)
    try:
        for _ in range(args.warmup):
            pool.run(2)
        times = []
        for _ in range(args.steps):
            _, inner = pool.run(frames)
            times.append(max(inner))
    finally:
        pool.close()
    ms = 1000 * statistics.mean(times)
    value = procs * frames * FRAME_SEC / (ms / 1000)
    sample = f"{procs} single-thread oracle streams x {frames} frames (+40-token prefill) per step, f32"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": CONFIG,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from pocket_tts_b200 import synth
    from pocket_tts_b200.engine import Engine, StreamSpec

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE {world}"
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL is only the timing barrier / max-over-ranks reduction here (no data-path collective); NCCL_DEBUG is left as
        # the caller set it, the JSON line is the last thing rank 0 prints
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    wnp = synth.make_weights(1234)
    wnp.update(synth.make_encoder_weights(4321))  # voice-cloning side, only used by the voice_from_pcm timing below
    eng_kw = dict(int8_weights=args.int8, codec_group=args.codec_group if args.codec_group > 1 else None,
                  lm_step_kernel=True if args.lm_step_kernel else None)
    eng = Engine(wnp, device=local, max_slots=STREAMS, kv_capacity=TOKENS + FRAMES + 3, **eng_kw)
    del wnp
    voice = eng.voice_from_prompt(synth.make_voice_prompt(VOICE_ROWS, seed=7))
    base = rank * STREAMS  # request ids of this shard
    specs = [StreamSpec(synth.make_tokens(TOKENS, seed=1000 + base + i), FRAMES, 3, 1e30, temp=0.7, seed=base + i)
             for i in range(STREAMS)]
    stream = torch.cuda.ExternalStream(eng.cuda_stream, device=local)

    def job(host: bool, profile=None):
        slots = eng.open_streams([voice] * STREAMS, specs)
        lag = max(1, args.codec_group)   # the PCM of frame f is fetched once frame f + lag has been enqueued (its codec group is complete)
        tickets = []
        t = eng.step_begin(slots) if host else None
        for f in range(FRAMES):
            if host:
                # public pipelined calls: frame f+1 is enqueued ahead of frame f's flags (PTTS_STEP_AHEAD; every stream
                # runs to max_gen_len here, so the host knows frame f is not the last), the flags of frame f are awaited
                # and read, then the PCM of frame f-lag is fetched while the device is already on frame f+1
                nxt = eng.step_begin(slots, ahead=True) if f + 1 < FRAMES else None
                fin, _, _ = eng.step_flags(t)
                tickets.append(t)
                if len(tickets) > lag:
                    pcm = eng.step_pcm(tickets.pop(0))
                t = nxt
            elif profile is not None and f in profile:
                eng.profile(True); eng.step_device(slots); eng.profile(False)
            else:
                eng.step_device(slots)
        for tk in tickets:
            pcm = eng.step_pcm(tk)
        eng.sync()
        if host:
            assert fin.all() and np.isfinite(pcm).all()
        for s in slots:
            eng.close_stream(int(s))

    def timed(host: bool, steps: int, warmup: int):
        for _ in range(warmup):
            job(host)
        barrier()
        eng.launch_count(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            job(host)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        launches = eng.launch_count()
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)   # the job takes as long as its slowest GPU
            ms = float(t.item())
            c = torch.tensor([launches], device="cuda", dtype=torch.int64)
            dist.all_reduce(c, op=dist.ReduceOp.SUM)
            launches = int(c.item())
        return ms / steps, launches

    sampler = ClockSampler(local)
    sampler.start()
    ms_dev, launches = timed(False, args.steps, args.warmup)
    clocks = sampler.stop()
    ms_e2e, _ = timed(True, max(1, min(args.steps, 3)), 1)

    audio_per_job = world * STREAMS * FRAMES * FRAME_SEC
    line = {"metric": METRIC, "value": audio_per_job / (ms_dev / 1000), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 operands, f32 accumulate/residual", "data": "synthetic", "config": CONFIG, "clocks": clocks,
            "e2e": {"value": audio_per_job / (ms_e2e / 1000), "unit": UNIT,
                    "h2d_bytes_per_step": STREAMS * (TOKENS * 4 + 4 + 24) + FRAMES * STREAMS * 4,
                    "d2h_bytes_per_step": FRAMES * STREAMS * (1920 * 4 + 1 + 32 * 4 + 4), "ms_per_step": ms_e2e},
            "gpu_launches": launches, "realtime_factor_per_gpu": STREAMS * FRAMES * FRAME_SEC / (ms_dev / 1000)}

    # BASELINE configs[4] slice: concurrent 60 s long-form requests (6 chunks x 125 frames, [pause:300ms] between chunks)
    # through the library's own continuous-batching scheduler (ptts_sched_*, C++), i16 PCM packed on the device and
    # copied to the host every frame; request-sharded like everything else (each rank runs its share)
    if args.longform > 0:
        from pocket_tts_b200.engine import NativeScheduler
        lf_chunks, lf_pause = 6, 300
        eng_lf = eng if args.longform <= STREAMS else None
        if eng_lf is None:
            voice.close(); eng.close()
            wnp2 = synth.make_weights(1234)
            eng = Engine(wnp2, device=local, max_slots=args.longform, kv_capacity=TOKENS + FRAMES + 3, **eng_kw)
            del wnp2
            voice = eng.voice_from_prompt(synth.make_voice_prompt(VOICE_ROWS, seed=7))
        reqs = []
        for r in range(args.longform):
            segs = []
            for c in range(lf_chunks):
                if c:
                    segs.append(("pause", lf_pause))
                segs.append(("text", StreamSpec(synth.make_tokens(TOKENS, seed=(base * 8 + r) * 16 + c), FRAMES, 3, 1e30, temp=0.7, seed=r * 16 + c)))
            reqs.append(segs)
        ns = NativeScheduler(eng, voice, args.longform)
        barrier()
        t0 = time.perf_counter()
        out = ns.run(reqs, i16=True, view=True)   # PCM read in place from the scheduler's host buffers (ptts_sched_result_view)
        torch.cuda.synchronize()
        lf_s = time.perf_counter() - t0
        lf_phases = dict(ns.last_times)
        want = lf_chunks * FRAMES * 1920 + (lf_chunks - 1) * lf_pause * 24
        assert all(o.shape == (want,) for o in out)
        assert all(int(np.abs(o[:1920 * 4].astype(np.int32)).sum()) > 0 for o in out[:: max(1, len(out) // 16)])   # real audio, host-readable
        del out
        ns.close()
        if world > 1:
            t = torch.tensor([lf_s], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            lf_s = float(t.item())
        line["longform"] = {"workload": f"configs[4] slice: {args.longform} concurrent requests per GPU x {want / 24000:.1f} s (6 chunks x 125 frames, "
                                        f"[pause:300ms]), native scheduler, i16 PCM to the host", "requests_per_gpu": args.longform,
                            "value": world * args.longform * want / 24000.0 / lf_s, "unit": UNIT, "wall_s": lf_s,
                            "phases_s": lf_phases}
        del reqs
        if eng_lf is None and rank == 0:
            # back to the headline engine for the roofline pass below
            voice.close(); eng.close()
            wnp2 = synth.make_weights(1234)
            wnp2.update(synth.make_encoder_weights(4321))
            eng = Engine(wnp2, device=local, max_slots=STREAMS, kv_capacity=TOKENS + FRAMES + 3, **eng_kw)
            del wnp2
            voice = eng.voice_from_prompt(synth.make_voice_prompt(VOICE_ROWS, seed=7))

    if rank == 0:
        # roofline pass: per-launch CUDA events on the engine's stream for 3 mid-utterance decode steps
        job(False, profile={60, 61, 62})
        rep = eng.profile_report()
        ovh_us = eng.profile_overhead_us()  # an empty kernel timed the same way
        pk = peaks()
        n_prof = 3
        classes = []
        tot = sum(max(v["ms"] * 1000 - ovh_us * v["launches"], 0.3 * v["launches"]) for v in rep.values())
        for tag, v in rep.items():
            us_raw = 1000 * v["ms"] / v["launches"]
            us = max(us_raw - ovh_us, 0.3)           # per-launch device time net of the event-pair cost
            by, fl = v["bytes"] / v["launches"], v["flops"] / v["launches"]
            gbs, tfs = by / us / 1e3, fl / us / 1e6
            bound = "hbm" if by / (pk["hbm"] * 1e9) >= fl / (pk["tf"] * 1e12) else "tensor"
            classes.append({"kernel": tag, "launches_per_step": v["launches"] / n_prof, "us_per_launch": us,
                            "us_per_launch_raw": us_raw, "share": us * v["launches"] / tot, "bound": bound, "GB/s": gbs,
                            "TFLOP/s": tfs, "frac": gbs / pk["hbm"] if bound == "hbm" else tfs / pk["tf"],
                            "bytes_per_launch": by, "flops_per_launch": fl})
        classes.sort(key=lambda c: -c["share"])
        # the dominant kernel = the kernel function with the largest share of the step, all its launches together
        fns = {}
        for tag, v in rep.items():
            f = fns.setdefault(v["fn"], dict(launches=0, us=0.0, us_raw=0.0, bytes=0.0, flops=0.0))
            f["launches"] += v["launches"]
            f["us"] += max(v["ms"] * 1000 - ovh_us * v["launches"], 0.3 * v["launches"])
            f["us_raw"] += v["ms"] * 1000
            f["bytes"] += v["bytes"]
            f["flops"] += v["flops"]
        fn, f = max(fns.items(), key=lambda kv: kv[1]["us"])
        nl = f["launches"]
        traffic, traffic_src = ncu_traffic()
        if (fn == "gemm_tc_kernel" and STREAMS > 4) or fn == "gemv_rows_kernel":
            # the dominant kernel's time WITHOUT any event-pair correction: its four decode shapes replayed as graphs of
            # back-to-back launches over the real weights of all six layers (ptts_profile_gemm_replay), one event pair per
            # graph; pooled over the 24 FlowLM launches of a step (the remaining gemm_tc_kernel launches -- flow head glue
            # and the short codec GEMMs -- are listed per class below with the event-pair method)
            rp = eng.gemm_replay(STREAMS, 20)
            us = sum(v["us"] for v in rp.values()) / 4
            by = sum(v["bytes"] for v in rp.values()) / 4
            gbs = by / us / 1e3
            line["roofline"] = {"kernel": fn, "bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"],
                                "traffic": traffic.get(fn), "traffic_source": traffic_src, "peak_source": pk["src"],
                                "us_per_launch": us, "algorithmic_bytes_per_launch": by, "launches_per_step": nl / n_prof,
                                "share_of_step": f["us"] / tot, "per_shape": rp,
                                "how": "the FlowLM decode Linear layers (in_proj, out_proj, linear1, linear2; 24 of the step's launches "
                                       "of this kernel; 1-4 rows run the small-batch GEMV of csrc/gemv.cuh, more the tcgen05 GEMM) replayed as captured graphs of 20 x 6 back-to-back launches over the "
                                       "real weights of every layer, CUDA events around each graph; achieved = algorithmic "
                                       "bytes (weights + operand rows + epilogue tensors) / mean launch time"}
        else:
            gbs, tfs = f["bytes"] / f["us"] / 1e3, f["flops"] / f["us"] / 1e6
            bound = "hbm" if f["bytes"] / (pk["hbm"] * 1e9) >= f["flops"] / (pk["tf"] * 1e12) else "tensor"
            line["roofline"] = {"kernel": fn, "bound": bound, "achieved": gbs if bound == "hbm" else tfs,
                                "peak": pk["hbm"] if bound == "hbm" else pk["tf"], "unit": "GB/s" if bound == "hbm" else "TFLOP/s",
                                "frac": (gbs / pk["hbm"]) if bound == "hbm" else (tfs / pk["tf"]),
                                "traffic": traffic.get(fn), "traffic_source": traffic_src, "peak_source": pk["src"],
                                "us_per_launch": f["us"] / nl, "us_per_launch_raw": f["us_raw"] / nl,
                                "event_pair_overhead_us": ovh_us, "share_of_step": f["us"] / tot, "launches_per_step": nl / n_prof,
                                "algorithmic_bytes_per_launch": f["bytes"] / nl, "algorithmic_flops_per_launch": f["flops"] / nl,
                                "how": "CUDA events around every launch of 3 mid-utterance decode steps on the engine's stream; the "
                                       "time of an empty kernel bracketed the same way is subtracted"}
        # the whole step against the HBM roofline: SURVEY 8(d) algorithmic bytes (weights once + per stream KV read / write,
        # Mimi KV, conv tails, PCM) over the measured step time
        l_mean = VOICE_ROWS + TOKENS + (FRAMES - 1) / 2
        step_bytes = 189.7e6 + STREAMS * (24576.0 * l_mean + 1.02e6 + 0.09e6 + 7680 + 0.2e6)
        step_us = 1000.0 * ms_dev / FRAMES
        line["roofline"]["step_frac"] = step_bytes / (step_us * 1e-6) / (pk["hbm"] * 1e9)
        line["roofline"]["step_algorithmic_bytes"] = step_bytes
        line["roofline"]["step_us"] = step_us
        line["kernel_functions"] = [{"kernel": k, "launches_per_step": v["launches"] / n_prof, "us_per_step": v["us"] / n_prof,
                                     "share": v["us"] / tot} for k, v in sorted(fns.items(), key=lambda kv: -kv[1]["us"])]
        line["kernel_classes"] = classes
        line["step_device_us_sum_of_kernels"] = tot / n_prof
        # time to first audio: open -> first 1920-sample frame on the host, single stream (configs[0] shape)
        ttfa = []
        one = [StreamSpec(synth.make_tokens(12, seed=5), 52, 5, 1e30, temp=0.7, seed=1)]
        for _ in range(12):
            t0 = time.perf_counter()
            s = eng.open_streams([voice], one)
            eng.step(s)
            ttfa.append(1000 * (time.perf_counter() - t0))
            eng.close_stream(int(s[0]))
        t0 = time.perf_counter()
        s = eng.open_streams([voice] * STREAMS, specs)
        eng.step(s)
        t64 = 1000 * (time.perf_counter() - t0)
        for x in s:
            eng.close_stream(int(x))
        # voice cloning from PCM (configs[2] shape: an 87-frame prompt like assets/ref.wav): Mimi encoder + speaker_proj +
        # FlowLM prefill, host PCM in, voice handle out
        pcm = synth.make_pcm(VOICE_ROWS * 1920, seed=3)
        tv = []
        for _ in range(5):
            t0 = time.perf_counter()
            v = eng.voice_from_pcm(pcm)
            tv.append(1000 * (time.perf_counter() - t0))
            v.close()
        # BASELINE configs[0] shape on the same engine: ONE utterance, 125 frames, device-resident steps -- the small-batch
        # path (Linear layers on the GEMV of csrc/gemv.cuh); non-headline, reported beside the 64-stream value
        one = [StreamSpec(synth.make_tokens(TOKENS, seed=5), FRAMES, 3, 1e30, temp=0.7, seed=1)]
        us1 = []
        stream = torch.cuda.ExternalStream(eng.cuda_stream, device=local)   # the long-form pass may have re-created the engine
        for _ in range(4):
            s = eng.open_streams([voice], one)
            eng.step_device(s)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _f in range(FRAMES - 1):
                eng.step_device(s)
            e1.record(stream)
            eng.sync()
            us1.append(1000.0 * e0.elapsed_time(e1) / (FRAMES - 1))
            eng.close_stream(int(s[0]))
        rp1 = eng.gemm_replay(1, 20)     # the four decode Linear shapes at ONE row: the small-batch GEMV, graph-replayed
        by1, t1 = sum(v["bytes"] for v in rp1.values()), sum(v["us"] for v in rp1.values())
        line["single_stream"] = {"workload": "configs[0] shape: 1 utterance x 125 frames, device-resident decode steps",
                                 "us_per_frame": statistics.median(us1[1:]),
                                 "realtime_factor": FRAME_SEC * 1e6 / statistics.median(us1[1:]),
                                 "linear_layers": {"kernel": "gemv_rows_kernel", "bound": "hbm", "achieved": by1 / t1 / 1e3, "unit": "GB/s",
                                                   "peak": peaks()["hbm"], "frac": by1 / t1 / 1e3 / peaks()["hbm"], "per_shape": rp1,
                                                   "traffic": ncu_traffic()[0].get("gemv_rows_kernel"),
                                                   "how": "in_proj / out_proj / linear1 / linear2 at one row over the real weights of "
                                                          "all six layers, 20 x 6 back-to-back launches per captured graph"}}
        line["ttfa_ms"] = {"p50_single_stream": statistics.median(ttfa[2:]), "batch64_first_frames": t64,
                           "voice_from_pcm_87_frames": statistics.median(tv[1:])}
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            procs = max(1, min(cores, 32))
            vals = cpu_sample(procs, FRAMES)
            line["cpu_baseline"] = {"value": vals[0], "unit": UNIT, "cores": procs, "kind": "port",
                                    "sample": f"{procs} single-thread oracle streams x {FRAMES} frames (+40-token prefill), f32, "
                                              f"host has {cores} cores"}
        print(json.dumps(line), flush=True)
    voice.close()
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=STREAMS,
                    help="concurrent streams per GPU (default 64 = BASELINE configs[1]; 256/512 explore configs[3]/[4] shapes)")
    ap.add_argument("--longform", type=int, default=512,
                    help="requests per GPU of the configs[4] slice run through the native scheduler (0 = skip); reported under `longform`")
    ap.add_argument("--codec-group", type=int, default=1, choices=[1, 2, 4],
                    help="frames per codec pass (ptts_engine_set_codec_group; non-headline when > 1: same PCM, the Mimi decoder + SEANet run once per group)")
    ap.add_argument("--lm-step-kernel", action="store_true",
                    help="force the persistent FlowLM step kernel (csrc/lm_step.cuh) instead of the per-layer launches (non-headline)")
    ap.add_argument("--int8", action="store_true",
                    help="per-tensor int8 weights (reference quantize.rs scheme), one-byte codes expanded in the decode GEMMs: configs[3] with --streams 256")
    args = ap.parse_args()
    if args.int8:
        CONFIG["workload"] = CONFIG["workload"].replace("f16-operand", "int8-weight (non-headline)")
        CONFIG["weights"] = "int8 per-tensor codes in HBM, f16 MMA operands, scale in the epilogue"
    if args.codec_group > 1:
        CONFIG["codec_group"] = args.codec_group
    if args.lm_step_kernel:
        CONFIG["lm_step_kernel"] = True
    if args.streams != STREAMS:
        globals()["STREAMS"] = args.streams
        CONFIG["streams_per_gpu"] = args.streams
        CONFIG["workload"] = CONFIG["workload"].replace("batch 64", f"batch {args.streams} (non-headline size)")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
