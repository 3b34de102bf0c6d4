"""CPU: host-side logic of the AlignAtt mirror (no engine math): rolling window bookkeeping, DRY
penalties, special-token tables, trace objects -- driven with a scripted fake engine, like the
reference's own policy tests drive fake models (tests/test_backend_deep_bugs.py:158-319)."""
import numpy as np
import pytest

from whisperlivekit_b200.alignatt import AlignAttConfig, StreamingAlignAtt, dry_penalties
from whisperlivekit_b200.dims import DIMS, SpecialTokens


class FakeEngine:
    """Scripted engine: returns (token, logprob, frame) from a list; records the calls it gets."""

    def __init__(self, dims, script, no_speech=0.0):
        self.dims, self.specials = dims, SpecialTokens.for_dims(dims)
        self.script, self.no_speech = list(script), no_speech
        self.audio, self.calls, self.bias = 0, [], []

    def open_session(self): return 0
    def close_session(self, sid): self.calls.append(("close",))
    def append_audio(self, sid, pcm): self.audio += len(pcm)
    def drop_audio(self, sid, n): self.audio -= n; self.calls.append(("drop", n))
    def clear_audio(self, sid): self.audio = 0
    def audio_len(self, sid): return self.audio
    def encode(self, sids): self.calls.append(("encode", self.audio)); return [self.audio // 320]
    def decode(self, sids, toks, sot_index=0): self.calls.append(("decode", list(toks[0])))
    def no_speech_prob(self, sids): return [self.no_speech]
    def suppress(self, sids, toks): self.calls.append(("suppress", len(toks)))
    def add_logit_bias(self, sid, toks, b): self.bias.append((list(toks), list(b)))
    def greedy_and_align(self, sids, window_iters=16): return [self.script.pop(0)]


D = DIMS["micro"]
SP = SpecialTokens.for_dims(D)


def test_stop_when_attention_reaches_the_end_drops_last_token():
    # content = 8000/320 = 25 frames; frame 24 is within frame_threshold (25) of the end -> stop, token dropped
    eng = FakeEngine(D, [(1000, -0.1, 24)])
    p = StreamingAlignAtt(eng, AlignAttConfig())
    p.insert_audio(np.zeros(8000, np.float32))
    tr = p.infer()
    assert tr.stop == "attention_end" and tr.new_tokens == [] and tr.step_tokens == [1000]
    # first decode call feeds the whole prefix
    assert [c for c in eng.calls if c[0] == "decode"][0][1] == list(SP.sot_sequence_including_notimestamps())


def test_tokens_accumulate_and_second_infer_prefills_them():
    eng = FakeEngine(D, [(1000, -0.1, 10), (1001, -0.1, 20), (1002, -0.1, 740), (1003, -0.2, 1497),
                         (1004, -0.1, 1499)])
    p = StreamingAlignAtt(eng, AlignAttConfig(frame_threshold=4))
    p.insert_audio(np.zeros(16000 * 30, np.float32))          # content 1500
    tr = p.infer()
    assert tr.new_tokens == [1000, 1001, 1002] and tr.stop == "attention_end"      # 1003 attends 3 frames from the end
    decodes = [c[1] for c in eng.calls if c[0] == "decode"]
    assert decodes[1:] == [[1000], [1001], [1002]]            # single-token steps after the prefill
    p.insert_audio(np.zeros(8000, np.float32))                # > 30 s with two segments: the first one leaves
    tr2 = p.infer()
    assert [c[1] for c in eng.calls if c[0] == "decode"][-1][:2] == [SP.sot_prev, 1000]   # moved into the context
    assert tr2.step_tokens == [1004]


def test_eot_completes_and_is_not_kept():
    eng = FakeEngine(D, [(1000, -0.1, 10), (SP.eot, -0.1, 12)])
    p = StreamingAlignAtt(eng, AlignAttConfig(frame_threshold=0))
    p.insert_audio(np.zeros(16000 * 10, np.float32))
    tr = p.infer()
    assert tr.stop == "eot" and tr.new_tokens == [1000]


def test_no_speech_short_circuits_before_any_suppression():
    eng = FakeEngine(D, [], no_speech=0.9)
    p = StreamingAlignAtt(eng, AlignAttConfig())
    p.insert_audio(np.zeros(8000, np.float32))
    tr = p.infer()
    assert tr.no_speech and tr.stop == "no_speech" and not any(c[0] == "suppress" for c in eng.calls)


def test_rewind_resets_to_committed_tokens():
    eng = FakeEngine(D, [(1000, -0.1, 900), (1001, -0.1, 100)])
    p = StreamingAlignAtt(eng, AlignAttConfig(frame_threshold=0))
    p.insert_audio(np.zeros(16000 * 30, np.float32))
    tr = p.infer()
    assert tr.stop == "rewind" and tr.new_tokens == [] and p.last_attend_frame == -200


def test_window_rolls_oldest_chunk_into_context():
    eng = FakeEngine(D, [(1000, -0.1, 10), (SP.eot, 0, 10)] + [(SP.eot, 0, 10)] * 200)
    p = StreamingAlignAtt(eng, AlignAttConfig(frame_threshold=0, audio_max_len=1.0))
    p.insert_audio(np.zeros(8000, np.float32)); p.infer()
    p.insert_audio(np.zeros(8000, np.float32)); p.infer()
    assert p.tokens[1] == [1000]
    p.insert_audio(np.zeros(8000, np.float32))                  # 1.5 s > 1.0 s: the first chunk leaves the window
    assert ("drop", 8000) in eng.calls and p.context == [1000] and p.cumulative_time_offset == 0.5
    assert p._current_tokens()[:2] == [SP.sot_prev, 1000]
    assert eng.audio == 16000


def test_dry_penalties_match_reference_rule():
    # "... a b c a b" -> last = b; earlier 'b' at index 1 followed by c with match length 2 ('a b')
    seq = [SP.sot, 10, 11, 12, 10, 11]
    assert dry_penalties(seq, SP.eot) == [(12, 1.0)]
    assert dry_penalties([1, 2, 3], SP.eot) == []
    assert dry_penalties([SP.sot, 10, 11, 12, 10, SP.eot], SP.eot) == []


def test_special_token_layouts():
    en, ml, v3 = (SpecialTokens.for_dims(DIMS[k]) for k in ("tiny.en", "tiny", "large-v3"))
    assert (en.eot, en.sot, en.no_timestamps, en.timestamp_begin) == (50256, 50257, 50362, 50363)
    assert (ml.eot, ml.sot, ml.no_speech, ml.no_timestamps) == (50257, 50258, 50362, 50363)
    assert (v3.num_languages, v3.translate, v3.no_timestamps, v3.timestamp_begin) == (100, 50359, 50364, 50365)
    assert list(ml.sot_sequence_including_notimestamps()) == [50258, 50259, 50359, 50363]
