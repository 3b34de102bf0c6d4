"""CPU: the batching caller shim (whisperlivekit_b200/batching.py).  Many threads issue single-session calls
(as WhisperLiveKit's worker threads do, reference audio_processor.py:543-551); the shim must coalesce them
into batched engine calls without changing any result, deliver errors to the right callers, and shut down."""
import threading

import numpy as np
import pytest

from golden_util import case_setup
from whisperlivekit_b200.alignatt import AlignAttConfig, StreamingAlignAtt
from whisperlivekit_b200.batching import BatchingEngine


class CountingEngine:
    """Arithmetic-free engine: every result is a pure function of (session, call index), so any mix-up of
    sessions inside a merged batch shows up as a wrong value."""

    def __init__(self, dims):
        from whisperlivekit_b200.dims import SpecialTokens
        self.dims, self.specials = dims, SpecialTokens.for_dims(dims)
        self.align_heads = [(0, 0)]
        self.n, self.audio, self.step, self.calls, self.in_call = 0, {}, {}, [], 0

    def _enter(self, op, sids):
        assert self.in_call == 0, "engine entered concurrently"
        self.in_call += 1
        self.calls.append((op, list(sids)))

    def open_session(self): self.n += 1; self.audio[self.n] = 0; self.step[self.n] = 0; return self.n
    def close_session(self, sid): pass
    def append_audio(self, sid, pcm): self.audio[sid] += len(pcm)
    def drop_audio(self, sid, n): self.audio[sid] -= n
    def clear_audio(self, sid): self.audio[sid] = 0
    def audio_len(self, sid): return self.audio[sid]
    def reset_decoder(self, sid): pass
    def add_logit_bias(self, sid, t, b): pass

    def encode(self, sids):
        self._enter("encode", sids); out = [self.audio[s] // 320 for s in sids]; self.in_call -= 1; return out

    def decode(self, sids, toks, sot_index=0):
        self._enter("decode", sids); assert len(toks) == len(sids); self.in_call -= 1

    def no_speech_prob(self, sids):
        self._enter("no_speech_prob", sids); self.in_call -= 1; return [0.0] * len(sids)

    def suppress(self, sids, toks):
        self._enter("suppress", sids); self.in_call -= 1

    def greedy_and_align(self, sids, window_iters=16):
        self._enter("greedy_and_align", sids)
        out = []
        for s in sids:
            self.step[s] += 1
            k = self.step[s]
            frame = min(40 * k, self.audio[s] // 320 - 1) if k % 6 else self.audio[s] // 320 - 1
            out.append((1000 + 37 * s + k, -0.01 * k, frame))
        self.in_call -= 1
        return out


def _drive(engine, n_sessions, n_chunks, threaded):
    from whisperlivekit_b200.dims import DIMS
    pols = [StreamingAlignAtt(engine, AlignAttConfig(frame_threshold=4)) for _ in range(n_sessions)]
    traces = [[] for _ in pols]

    def run(i):
        for c in range(n_chunks):
            pols[i].insert_audio(np.zeros(8000 + 160 * i, np.float32))
            tr = pols[i].infer()
            traces[i].append((tr.stop, tuple(tr.new_tokens), tuple(tr.step_tokens), tuple(tr.step_frames)))

    if threaded:
        ths = [threading.Thread(target=run, args=(i,)) for i in range(n_sessions)]
        [t.start() for t in ths]
        [t.join(timeout=60) for t in ths]
        assert not any(t.is_alive() for t in ths)
    else:
        for i in range(n_sessions):
            run(i)
    return traces


def test_threads_are_coalesced_and_results_unchanged():
    from whisperlivekit_b200.dims import DIMS
    direct = _drive(CountingEngine(DIMS["micro"]), 12, 5, threaded=False)
    inner = CountingEngine(DIMS["micro"])
    be = BatchingEngine(inner, max_batch=8, max_wait_s=0.05)
    got = _drive(be, 12, 5, threaded=True)
    be.close()
    assert got == direct
    assert be.stats["max_sessions_in_call"] > 1 and be.stats["max_sessions_in_call"] <= 8
    assert be.stats["calls"] < be.stats["requests"]                   # fewer engine calls than caller requests
    assert all(len(set(sids)) == len(sids) for _, sids in inner.calls)   # one call in flight per session


def test_everyone_arrived_fires_without_waiting_for_the_timeout():
    import time
    from whisperlivekit_b200.dims import DIMS
    inner = CountingEngine(DIMS["micro"])
    be = BatchingEngine(inner, max_batch=64, max_wait_s=5.0)          # a timeout would make this test take minutes
    t0 = time.perf_counter()
    got = _drive(be, 4, 2, threaded=True)
    dt = time.perf_counter() - t0
    be.close()
    assert dt < 4.0 and len(got) == 4
    assert be.stats["max_sessions_in_call"] >= 2


def test_errors_reach_the_callers_of_the_failing_batch_only():
    from whisperlivekit_b200.dims import DIMS

    class Failing(CountingEngine):
        def encode(self, sids):
            if 2 in sids:
                raise RuntimeError("boom")
            return super().encode(sids)

    inner = Failing(DIMS["micro"])
    for _ in range(3):
        inner.open_session()
    be = BatchingEngine(inner, max_batch=1, max_wait_s=0.0)           # batches of one: only session 2 fails
    f1, f2, f3 = be.submit("encode", [1]), be.submit("encode", [2]), be.submit("encode", [3])
    assert f1.result(timeout=5) == [0] and f3.result(timeout=5) == [0]
    with pytest.raises(RuntimeError, match="boom"):
        f2.result(timeout=5)
    be.close()
    with pytest.raises(RuntimeError):
        be.submit("encode", [1])


def test_shim_over_the_oracle_matches_direct_calls():
    """Real arithmetic behind the shim (CPU oracle, micro model): per-session token streams are identical
    whether sessions are driven one after the other or concurrently through the batcher."""
    from oracle import whisper_oracle as wo
    g, dims, sd, audio, heads = case_setup("micro")

    def drive(engine, threaded):
        pols = [StreamingAlignAtt(engine, AlignAttConfig()) for _ in range(3)]
        out = [[] for _ in pols]

        def run(i):
            a = audio[4000 * i:]
            for c in range(3):
                pols[i].insert_audio(a[c * 8000:(c + 1) * 8000])
                tr = pols[i].infer()
                out[i].append((tr.stop, tuple(tr.step_tokens), tuple(tr.step_frames)))

        if threaded:
            ths = [threading.Thread(target=run, args=(i,)) for i in range(3)]
            [t.start() for t in ths]
            [t.join(timeout=300) for t in ths]
        else:
            for i in range(3):
                run(i)
        return out

    direct = drive(wo.OracleEngine(dims, sd, heads), False)
    be = BatchingEngine(wo.OracleEngine(dims, sd, heads), max_batch=8, max_wait_s=0.02)
    got = drive(be, True)
    be.close()
    assert got == direct
    assert be.stats["max_sessions_in_call"] >= 2


def test_qwen_tower_calls_are_coalesced_too():
    """The same shim in front of the Qwen3 tower engine API (forward_chunk / append_audio per stream from worker
    threads): outputs equal the sequential run, calls are merged."""
    from oracle.make_golden_qwen import mel_stream
    from oracle.qwen_oracle import QwenTowerOracle
    from whisperlivekit_b200.qwen_dims import QWEN_DIMS, synthetic_tower_state_dict
    dims = QWEN_DIMS["qnano"]
    sd = synthetic_tower_state_dict(dims, seed=2)
    mels = mel_stream(1500, dims.n_mels, seed=6)

    def drive(engine, threaded):
        sids = [engine.open_session() for _ in range(4)]
        out = [[] for _ in sids]

        def run(i):
            pos = 100 * i
            for _ in range(6):
                out[i].append(engine.forward_chunk([sids[i]], [mels[pos: pos + 100 + 7 * i]])[0])
                pos += 100 + 7 * i
            out[i].append(engine.flush_pending([sids[i]])[0])

        if threaded:
            ths = [threading.Thread(target=run, args=(i,)) for i in range(4)]
            [t.start() for t in ths]
            [t.join(timeout=120) for t in ths]
            assert not any(t.is_alive() for t in ths)
        else:
            for i in range(4):
                run(i)
        return out

    direct = drive(QwenTowerOracle(dims, sd), False)
    be = BatchingEngine(QwenTowerOracle(dims, sd), max_batch=4, max_wait_s=0.05)
    got = drive(be, True)
    be.close()
    for a, b in zip(got, direct):
        for x, y in zip(a, b):
            assert x.shape == y.shape and (x.size == 0 or np.abs(x - y).max() < 1e-6)
    assert be.stats["by_op"]["forward_chunk"]["calls"] < 24 and be.stats["max_sessions_in_call"] >= 2


def test_asyncio_callers_need_no_thread_per_stream():
    """submit() hands back a concurrent Future: an event loop awaits many streams' calls at once on ONE thread (the
    reference parks a worker thread per stream, audio_processor.py:543-551) and they are still merged."""
    import asyncio
    from whisperlivekit_b200.dims import DIMS
    inner = CountingEngine(DIMS["micro"])
    sids = [inner.open_session() for _ in range(16)]
    for s in sids:
        inner.append_audio(s, np.zeros(16000 + 320 * s, np.float32))
    be = BatchingEngine(inner, max_batch=16, max_wait_s=0.05)

    async def stream(sid):
        content = (await asyncio.wrap_future(be.submit("encode", [sid])))[0]
        await asyncio.wrap_future(be.submit("decode", [sid], [[1, 2, 3]], sot_index=0))
        tok = (await asyncio.wrap_future(be.submit("greedy_and_align", [sid], window_iters=16)))[0]
        return content, tok[0]

    async def main():
        return await asyncio.gather(*[stream(s) for s in sids])

    res = asyncio.run(main())
    be.close()
    assert [r[0] for r in res] == [(16000 + 320 * s) // 320 for s in sids]
    assert [r[1] for r in res] == [1000 + 37 * s + 1 for s in sids]
    assert be.stats["max_sessions_in_call"] >= 8 and be.stats["calls"] <= 12
