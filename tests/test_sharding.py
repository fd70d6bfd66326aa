"""Multi-GPU path is request sharding with no data-path collective.  The host logic is
checked with a world_size-2 gloo group on CPU."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pocket_tts_b200.tts_model import shard_requests


def test_shards_partition():
    for n in (0, 1, 7, 64, 4096):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                seen += list(shard_requests(n, world, r))
            assert seen == list(range(n))
            sizes = [len(shard_requests(n, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_req, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_requests(n_req, world, rank)
    frames = torch.tensor([len(mine) * 125], dtype=torch.int64)  # each request emits 125 frames
    ms = torch.tensor([10.0 + rank], dtype=torch.float64)        # per-rank device time
    dist.all_reduce(frames, op=dist.ReduceOp.SUM)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)                    # bench.py contract: max over ranks
    if rank == 0:
        q.put((int(frames), float(ms)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_aggregate_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 129, q)) for r in range(2)]
    for p in procs:
        p.start()
    frames, ms = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert frames == 129 * 125 and ms == 11.0


class _FakeEngine:
    """Host-only stand-in for the engine: a stream's frame f is the constant (spec id * 1000 + f); a stream finishes
    after min(spec.eos_at, spec.max_gen_len) frames.  Same call protocol as pocket_tts_b200.engine.Engine (open / begin /
    flags / pcm / close, PTTS_STEP_AHEAD with overrun rows), with the protocol's ordering rules asserted."""

    def __init__(self, max_batch):
        import numpy as np
        self.np, self.max_batch = np, max_batch
        self.free = list(range(max_batch))
        self.live = {}       # slot -> [spec, frames done, finished]
        self.tickets = {}    # ticket -> [slots, per-row frame index (-1 = overrun), per-row last-frame flag, flags fetched]
        self.next = 0
        self.batch_sizes = []
        self.ahead_steps = 0
        self.last_overrun = None

    def open_streams(self, voices, specs):
        assert len(voices) == len(specs) <= len(self.free)
        out = []
        for s in specs:
            slot = self.free.pop(0)
            self.live[slot] = [s, 0, False]
            out.append(slot)
        return self.np.asarray(out, self.np.int32)

    def step_begin(self, slots, want_pcm=True, ahead=False):
        assert all(int(s) in self.live for s in slots) and len(set(map(int, slots))) == len(slots)
        unfetched = [t for t, v in self.tickets.items() if not v[3]]
        assert len(unfetched) <= (1 if ahead else 0), "flags of the previous step must be fetched first (one step ahead at most)"
        assert len(self.tickets) < 3
        self.ahead_steps += bool(ahead)
        self.batch_sizes.append(len(slots))
        rows, last = [], []
        for s in slots:
            st = self.live[int(s)]
            if st[2]:
                assert ahead, "a finished stream may only be stepped by a step enqueued ahead"
                rows.append(-1); last.append(True)
                continue
            rows.append(st[1])
            st[1] += 1
            st[2] = st[1] >= min(getattr(st[0], "eos_at", 1 << 30), st[0].max_gen_len)
            last.append(st[2])
        self.tickets[self.next] = [[int(s) for s in slots], rows, last, False]
        self.next += 1
        return self.next - 1

    def step_flags(self, t):
        assert all(v[3] for k, v in self.tickets.items() if k < t), "flags are fetched in step order"
        slots, rows, last, _ = self.tickets[t]
        self.tickets[t][3] = True
        self.last_overrun = self.np.array([r < 0 for r in rows])
        return self.np.array(last), None, None

    def step_pcm(self, t, want=True):
        slots, rows, _, fetched = self.tickets.pop(t)
        assert fetched
        if not want:
            return None
        return self.np.stack([self.np.full(4, self.live[s][0].tokens * 1000 + f, self.np.float32) for s, f in zip(slots, rows)])

    def close_stream(self, slot):
        assert not any(slot in t[0] for t in self.tickets.values()), "a closed slot still has an undrained step"
        self.free.append(slot)
        del self.live[slot]


def test_continuous_batching_host_logic_without_gpu():
    """BASELINE configs[4] host logic (tts_model.BatchScheduler): chunks of one request in order, pauses as exact zero
    runs (pause.rs:183-185), different requests share the batch, never more than max_batch rows, slots recycled."""
    import numpy as np
    from types import SimpleNamespace
    from pocket_tts_b200.tts_model import BatchScheduler
    eng = _FakeEngine(max_batch=3)

    def chunk(cid, frames):
        return ("text", SimpleNamespace(tokens=cid, max_gen_len=frames))

    requests = [
        [chunk(1, 2), ("pause", 500), chunk(2, 3)],
        [chunk(3, 1)],
        [("pause", 100), chunk(4, 4), chunk(5, 1), ("pause", 250)],
        [chunk(6, 2)],
        [],
    ]
    # the fake engine's frame is 4 samples long; silence_samples is in real samples
    out = BatchScheduler(eng, voice=None, max_batch=3).run(requests)

    def frames(cid, n):
        return np.concatenate([np.full(4, cid * 1000 + f, np.float32) for f in range(n)])

    want = [
        np.concatenate([frames(1, 2), np.zeros(12000, np.float32), frames(2, 3)]),
        frames(3, 1),
        np.concatenate([np.zeros(2400, np.float32), frames(4, 4), frames(5, 1), np.zeros(6000, np.float32)]),
        frames(6, 2),
        np.zeros(0, np.float32),
    ]
    for got, w in zip(out, want):
        np.testing.assert_array_equal(got, w)
    assert max(eng.batch_sizes) <= 3 and sorted(eng.free) == [0, 1, 2] and not eng.live and not eng.tickets
    # the same job with the device kept one step ahead of the host, and with streams that end early at "EOS"
    # (the step enqueued ahead then carries an overrun row, which must never reach the output)
    def chunk_eos(cid, frames, eos_at):
        return ("text", SimpleNamespace(tokens=cid, max_gen_len=frames, eos_at=eos_at))

    eos_requests = [
        [chunk_eos(1, 9, 3), ("pause", 100), chunk(2, 3)],
        [chunk_eos(3, 6, 6), chunk_eos(4, 8, 1)],
        [chunk(5, 5)],
        [chunk_eos(6, 7, 2)],
    ]
    for reqs in (requests, eos_requests):
        lock = BatchScheduler(_FakeEngine(max_batch=3), voice=None, max_batch=3).run(reqs)
        eng2 = _FakeEngine(max_batch=3)
        ahead = BatchScheduler(eng2, voice=None, max_batch=3).run(reqs, ahead=True)
        for a, b in zip(ahead, lock):
            np.testing.assert_array_equal(a, b)
        assert sorted(eng2.free) == [0, 1, 2] and not eng2.live and not eng2.tickets
        if reqs is eos_requests:   # (in the first job some row is on its last possible frame at every step)
            assert eng2.ahead_steps >= 3
