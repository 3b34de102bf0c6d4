"""GPU: edge cases of the C-ABI path against the oracle on the same inputs (fp32 mode, micro dims) -- the sizes the
reference itself can reach: one hop of audio, exactly / more than the 30 s window, a text context filled to
n_text_ctx, ragged batches with a 1-token row next to a long prefill (multi-token calls only as the first call of an
epoch: the reference's mask slicing, model.py:164-169, admits nothing else); and the error contract (status code + message,
never a crash)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from golden_util import case_setup
from whisperlivekit_b200.weights import synthetic_audio


@pytest.fixture(scope="module")
def pair():
    from oracle import whisper_oracle as wo
    from whisperlivekit_b200.engine import WhisperEngine
    g, dims, sd, audio, heads = case_setup("micro")
    eng = WhisperEngine(dims, sd, heads, precision="fp32", max_sessions=4, max_batch=4)
    yield eng, wo.OracleEngine(dims, sd, heads), g, dims
    eng.close()


def _both(pair, audio, prefix, steps=()):
    eng, orc, g, dims = pair
    out = []
    for E in (eng, orc):
        s = E.open_session()
        E.append_audio(s, audio)
        content = E.encode([s])[0]
        E.decode([s], [list(prefix)], sot_index=0)
        lg = [E.read_logits(s).copy()]
        for t in steps:
            E.decode([s], [[int(t)]])
            lg.append(E.read_logits(s).copy())
        tok, lp, frame = E.greedy_and_align([s])[0]
        out.append((content, lg, tok, frame, E.read_encoder(s).copy()))
        E.close_session(s)
    return out


@pytest.mark.parametrize("seconds", [0.02, 0.05, 0.3, 29.99, 30.0, 30.5])
def test_audio_lengths_at_the_limits(pair, seconds):
    """320 samples (one encoder frame of content) ... more than the 30 s window (content_mel_len > 1500 frames is what the reference
    computes too, simul_whisper.py:349-350; the mel is trimmed to 3000 frames, audio.py:65-88)."""
    g = pair[2]
    audio = synthetic_audio(31.0, seed=5)[: int(round(seconds * 16000))]
    (c0, l0, t0, f0, e0), (c1, l1, t1, f1, e1) = _both(pair, audio, g["forced_prefix"])
    assert c0 == c1
    assert np.abs(e0 - e1).max() < 1e-3
    assert np.abs(l0[0] - l1[0]).max() < 1e-3
    assert (t0, f0) == (t1, f1)


def test_text_context_filled_to_n_text_ctx(pair):
    """Prefill + steps up to exactly n_text_ctx tokens (the reference loop's bound, align_att_base.py:206), one more
    is refused."""
    from whisperlivekit_b200._lib import WlkError
    eng, orc, g, dims = pair
    n = dims.n_text_ctx
    rng = np.random.default_rng(3)
    prefix = list(g["forced_prefix"]) + [int(t) for t in rng.integers(1000, 5000, n - 8 - len(g["forced_prefix"]))]
    steps = [int(t) for t in rng.integers(1000, 5000, 8)]
    audio = synthetic_audio(4.0, seed=9)
    (c0, l0, t0, f0, _), (c1, l1, t1, f1, _) = _both(pair, audio, prefix, steps)
    for a, b in zip(l0, l1):
        assert np.abs(a - b).max() < 1e-3
    assert (t0, f0) == (t1, f1)
    s = eng.open_session()
    eng.append_audio(s, audio)
    eng.encode([s])
    eng.decode([s], [prefix + steps], sot_index=0)                    # exactly n_text_ctx rows
    with pytest.raises(WlkError, match="exceed n_text_ctx"):
        eng.decode([s], [[1234]])
    eng.close_session(s)


def test_ragged_batch_one_token_next_to_a_long_prefill(pair):
    eng, orc, g, dims = pair
    audio = synthetic_audio(6.0, seed=11)
    res = []
    for E in (eng, orc):
        a, b, c = E.open_session(), E.open_session(), E.open_session()
        E.append_audio(a, audio); E.append_audio(b, audio[:20000]); E.append_audio(c, audio[5000:5640])
        E.encode([a, b, c])
        E.decode([a, b, c], [list(g["forced_prefix"]), [int(g["forced_prefix"][0])], list(g["forced_prefix"])[:3]], sot_index=0)
        E.decode([c, a], [[1500], [1501]])                            # a subset of the batch, in a different order
        res.append([E.read_logits(s).copy() for s in (a, b, c)] + [E.greedy_and_align([a, b, c])])
        for s in (a, b, c):
            E.close_session(s)
    for x, y in zip(res[0][:3], res[1][:3]):
        assert np.abs(x - y).max() < 1e-3
    assert [(r[0], r[2]) for r in res[0][3]] == [(r[0], r[2]) for r in res[1][3]]


def test_error_contract(pair):
    from whisperlivekit_b200._lib import WlkError
    eng, orc, g, dims = pair
    s = eng.open_session()
    with pytest.raises(WlkError, match="no audio"):
        eng.encode([s])
    with pytest.raises(WlkError, match="decode before encode"):
        eng.decode([s], [[1, 2, 3]])
    eng.append_audio(s, synthetic_audio(1.0, seed=1))
    eng.encode([s])
    with pytest.raises(WlkError, match="twice"):
        eng.decode([s, s], [[1], [2]])
    with pytest.raises(WlkError, match="out of range"):
        eng.decode([s], [[dims.n_vocab]])
    with pytest.raises(WlkError, match="decode rows 0|empty token list"):
        eng.decode([s], [[]])
    with pytest.raises(WlkError, match="invalid session"):
        eng.decode([99], [[1]])
    with pytest.raises(WlkError, match="cannot drop"):
        eng.drop_audio(s, 10 ** 9)
    with pytest.raises(WlkError, match="overflow"):
        eng.append_audio(s, np.zeros(2 * 480000, np.float32))
    eng.close_session(s)
    with pytest.raises(WlkError, match="invalid session"):
        eng.close_session(s)
    sids = [eng.open_session() for _ in range(4)]
    with pytest.raises(WlkError, match="in use"):
        eng.open_session()
    for x in sids:
        eng.close_session(x)


def test_pcm16_ingest_equals_host_conversion(pair):
    """wlk_session_append_pcm16 == the reference's host conversion (int16 / 32768.0) followed by append_audio: same
    mel, bit for bit."""
    eng, orc, g, dims = pair
    pcm = (np.clip(synthetic_audio(2.0, seed=21), -1, 1) * 32767).astype(np.int16)
    a, b = eng.open_session(), eng.open_session()
    eng.append_pcm16(a, pcm[:12000].tobytes())
    eng.append_pcm16(a, pcm[12000:])
    eng.append_audio(b, pcm.astype(np.float32) / 32768.0)
    assert eng.audio_len(a) == eng.audio_len(b) == len(pcm)
    assert eng.encode([a, b]) == [100, 100]
    assert np.array_equal(eng.read_mel(a), eng.read_mel(b))
    eng.close_session(a); eng.close_session(b)


def test_incremental_log_mel_is_bit_identical_to_full_recompute(monkeypatch):
    """N1, exact part: a rolling window (0.5 s in, oldest 0.5 s out, plus growth and odd-sized steps) encoded
    incrementally -- only the ~2 leading and ~50 trailing frames are transformed, the rest of the raw log-mel is moved --
    gives bit-identical mel and encoder output to a session that recomputes every frame (fresh session, same audio)."""
    from whisperlivekit_b200.dims import DIMS
    from whisperlivekit_b200.engine import WhisperEngine
    from whisperlivekit_b200.weights import synthetic_audio, synthetic_state_dict
    dims = DIMS["micro"]
    sd = synthetic_state_dict(dims, seed=11)
    heads = [(0, 1), (1, 0), (1, 1)]
    eng = WhisperEngine(dims, sd, heads, precision="fp32", max_sessions=2, max_batch=2)
    audio = synthetic_audio(40.0, seed=77)
    s = eng.open_session()
    window = np.zeros(0, np.float32)
    pos = 0
    # (append samples, drop samples): growth, steady 0.5 s slides at the 30 s cap, a slide that is not a whole number of
    # frames (falls back to a full pass), a pure append, a big drop
    plan = [(160000, 0), (8000, 0), (320000 - 8000, 0), (8000, 8000), (8000, 8000), (8000, 8000), (5000, 5001), (3000, 0),
            (8000, 8000), (16000, 100000), (8000, 8000)]
    for k, (app, drop) in enumerate(plan):
        seg = audio[pos: pos + app]; pos += app
        eng.append_audio(s, seg)
        window = np.concatenate([window, seg])
        if drop:
            eng.drop_audio(s, drop)
            window = window[drop:]
        c = eng.encode([s])[0]
        mel_inc, enc_inc = eng.read_mel(s), eng.read_encoder(s)
        f = eng.open_session()                                  # fresh session: full pass over the same window
        eng.append_audio(f, window)
        assert eng.encode([f])[0] == c
        mel_full, enc_full = eng.read_mel(f), eng.read_encoder(f)
        eng.close_session(f)
        assert np.array_equal(mel_inc, mel_full), (k, np.abs(mel_inc - mel_full).max())
        assert np.array_equal(enc_inc, enc_full), k
    eng.close()
