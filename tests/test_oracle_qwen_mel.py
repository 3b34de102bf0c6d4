"""CPU: the incremental log-mel oracle against fixtures recorded from the reference's StreamingMelExtractor over the
real Hugging Face WhisperFeatureExtractor (oracle/make_golden_qwen_mel.py)."""
import os

import numpy as np

from whisperlivekit_b200.weights import mel_filterbank

HERE = os.path.dirname(os.path.abspath(__file__))


def mel_case():
    from oracle.make_golden_qwen_mel import MEL_SCHEDULE, speechlike
    g = dict(np.load(os.path.join(HERE, "golden", "qwen_mel.npz")))
    assert [int(x) for x in g["schedule"]] == MEL_SCHEDULE
    return g, speechlike(sum(MEL_SCHEDULE), seed=17), MEL_SCHEDULE


def check_mel_stream(append, flush, emitted, g, audio, sched, tol):
    a, worst, total = 0, 0.0, 0
    for i, n in enumerate(sched + [-1]):
        m = flush() if n < 0 else append(audio[a: a + n])
        a += max(n, 0)
        m = np.zeros((0, 128), np.float32) if m is None else np.asarray(m)
        assert m.shape[0] == int(g[f"frames{i}"]), f"call {i}: frames"
        assert emitted() == int(g[f"emitted{i}"])
        total += m.shape[0]
        if m.size:
            worst = max(worst, float(np.abs(m.reshape(-1)[g[f"idx{i}"]] - g[f"val{i}"]).max()))
            worst = max(worst, float(np.abs(m.astype(np.float64).sum(axis=1) - g[f"rowsum{i}"]).max()) / 128 ** 0.5)
    assert total == sum(sched) // 160
    assert worst < tol, worst
    return worst


def test_oracle_matches_reference_extractor_fixtures():
    from oracle.qwen_mel_oracle import StreamingMelOracle
    g, audio, sched = mel_case()
    o = StreamingMelOracle(mel_filterbank(128))
    check_mel_stream(o.append, o.flush, lambda: o.emitted_frames, g, audio, sched, 5e-5)
