"""GPU: the Silero VAD engine (wlk_vad_*, through the C ABI) against the probabilities recorded from the reference's
scripted model (tests/golden/vad.npz, oracle/make_golden_vad.py) and against the CPU oracle on ragged batches."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _audio():
    from whisperlivekit_b200.weights import synthetic_audio
    return np.concatenate([synthetic_audio(2.0, seed=31), np.zeros(8000, np.float32), 0.3 * synthetic_audio(1.5, seed=32)])


def test_vad_engine_matches_reference_fixture():
    from oracle.vad_oracle import synthetic_vad_state_dict
    from whisperlivekit_b200.vad import B200VadModel, VadEngine
    g = dict(np.load(os.path.join(HERE, "golden", "vad.npz")))
    audio, n = _audio(), int(g["n_windows"])
    eng = VadEngine(synthetic_vad_state_dict(seed=9), max_sessions=4)
    # one call for the whole stream (the kernel walks the windows in order) ...
    s = eng.open_session()
    probs = eng.forward([s], [audio[: n * 512]])[0]
    assert np.abs(probs - g["probs_seeded"]).max() < 2e-5
    # ... equals window-by-window calls through the scripted-model duck type, after a reset
    m = B200VadModel(eng)
    one = np.asarray([float(m(audio[i * 512:(i + 1) * 512], 16000)[0, 0]) for i in range(n)], np.float32)
    assert np.abs(one - probs).max() < 1e-6
    m.reset_states()
    again = float(m(audio[:512], 16000)[0, 0])
    assert abs(again - probs[0]) < 1e-6
    with pytest.raises(ValueError):
        m(audio[:500], 16000)
    eng.close()


def test_vad_engine_ragged_batch_equals_oracle():
    from oracle.vad_oracle import VadOracle, synthetic_vad_state_dict
    from whisperlivekit_b200.vad import VadEngine
    from whisperlivekit_b200.weights import synthetic_audio
    sd = synthetic_vad_state_dict(seed=4)
    eng, orc = VadEngine(sd, max_sessions=48), VadOracle(sd)
    rng = np.random.default_rng(3)
    n = 40
    se = [eng.open_session() for _ in range(n)]
    so = [orc.open_session() for _ in range(n)]
    streams = [synthetic_audio(3.0, seed=100 + i) * float(rng.uniform(0.05, 1.0)) for i in range(n)]
    pos = [0] * n
    for call in range(4):                                           # streams advance by different numbers of windows per call
        k = [int(rng.integers(0, 9)) for _ in range(n)]
        chunks = [streams[i][pos[i]: pos[i] + k[i] * 512] for i in range(n)]
        got = eng.forward(se, chunks)
        for i in range(n):
            want = np.concatenate([orc.forward([so[i]], [chunks[i][j * 512:(j + 1) * 512]]) for j in range(k[i])]) if k[i] else np.zeros(0, np.float32)
            assert got[i].shape == want.shape
            if k[i]:
                assert np.abs(got[i] - want).max() < 2e-5, (call, i)
            pos[i] += k[i] * 512
    eng.close()
