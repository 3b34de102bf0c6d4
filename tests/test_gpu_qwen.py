"""GPU: the Qwen3-ASR causal audio tower through the C ABI (wlk_qwen_*) against (a) fixtures recorded from the
reference's QwenAudioCausalKVEncoder and (b) the CPU oracle on the same seeded inputs.

Tolerances: fp32 mode (SIMT GEMMs, fp32 activations) within 1e-3 of the reference on outputs of std ~0.55;
bf16 mode (tcgen05 GEMMs where the shape allows, bf16 activations and K/V, fp32 residual stream / LayerNorm /
softmax) within 6e-2."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from test_oracle_qwen import check_stream, qwen_case
from whisperlivekit_b200.qwen_dims import QWEN_DIMS, synthetic_tower_state_dict


@pytest.mark.parametrize("name", ["qnano", "qnano-chunk", "qnano-tail", "qnano-tail-bidir"])
def test_fp32_tower_matches_reference_fixtures(name):
    from whisperlivekit_b200.qwen_engine import QwenTowerEngine
    g, dims, sd, mels, sched = qwen_case(name)
    eng = QwenTowerEngine(dims, sd, precision="fp32", max_sessions=2, max_batch=2)
    worst = check_stream(eng, g, mels, sched, 1e-3)
    eng.close()


@pytest.mark.parametrize("name", ["qnano", "qnano-chunk", "qnano-tail", "qnano-tail-bidir"])
def test_bf16_tower_close_to_reference_fixtures(name):
    from whisperlivekit_b200.qwen_engine import QwenTowerEngine
    g, dims, sd, mels, sched = qwen_case(name)
    eng = QwenTowerEngine(dims, sd, precision="bf16", max_sessions=2, max_batch=2)
    check_stream(eng, g, mels, sched, 6e-2)
    eng.close()


def test_ragged_batch_of_sessions_equals_oracle():
    """Three sessions fed different amounts per call (one gets several blocks at once, one nothing): batched rounds
    on the device == the oracle session by session; a reset session restarts at position 0."""
    from oracle.qwen_oracle import QwenTowerOracle
    from whisperlivekit_b200.qwen_engine import QwenTowerEngine
    g, dims, sd, mels, sched = qwen_case("qnano")
    eng = QwenTowerEngine(dims, sd, precision="fp32", max_sessions=4, max_batch=4)
    orc = QwenTowerOracle(dims, sd)
    plan = [[200, 30, 0], [190, 600, 0], [10, 0, 385], [0, 500, 7]]
    res = {}
    for tag, E in (("cuda", eng), ("oracle", orc)):
        sids = [E.open_session() for _ in range(3)]
        pos = [0, 300, 700]
        outs = []
        for call in plan:
            chunks = [mels[pos[i]: pos[i] + call[i]] for i in range(3)]
            pos = [pos[i] + call[i] for i in range(3)]
            outs.append(E.forward_chunk(sids, chunks))
        E.reset_session(sids[1])
        outs.append(E.forward_chunk([sids[1]], [mels[:400]]))
        res[tag] = (outs, [E.emitted_steps(s) for s in sids], [E.pending_frames(s) for s in sids])
    assert res["cuda"][1:] == res["oracle"][1:]
    n_rows = 0
    for a, b in zip(res["cuda"][0], res["oracle"][0]):
        for x, y in zip(a, b):
            assert x.shape == y.shape
            n_rows += x.shape[0]
            if x.size:
                assert np.abs(x - y).max() < 1e-3
    assert n_rows > 200
    eng.close()


def test_real_geometry_fp32_and_bf16_against_oracle():
    """Qwen3-ASR-0.6B tower geometry (d 896, 18 layers, 14 heads, conv 480): two blocks of a stream plus a sub-block
    remainder; fp32 within 1e-3 of the oracle, bf16 (tcgen05 GEMMs: K = 4320 im2col, 7680 conv_out) within 8e-2."""
    from oracle.make_golden_qwen import mel_stream
    from oracle.qwen_oracle import QwenTowerOracle
    from whisperlivekit_b200.qwen_engine import QwenTowerEngine
    dims = QWEN_DIMS["qwen3-asr-0.6b"]
    sd = synthetic_tower_state_dict(dims, seed=5)
    mels = mel_stream(500, dims.n_mels, seed=8)
    orc = QwenTowerOracle(dims, sd)
    so = orc.open_session()
    ref = [orc.forward_chunk([so], [mels[:250]])[0], orc.forward_chunk([so], [mels[250:]])[0]]
    assert ref[0].shape == (24, dims.out_dim) and ref[1].shape == (24, dims.out_dim)
    for prec, tol in (("fp32", 1e-3), ("bf16", 8e-2)):
        eng = QwenTowerEngine(dims, sd, precision=prec, max_sessions=2, max_batch=2)
        s = eng.open_session()
        got = [eng.forward_chunk([s], [mels[:250]])[0], eng.forward_chunk([s], [mels[250:]])[0]]
        for a, b in zip(got, ref):
            assert a.shape == b.shape
            err = np.abs(a - b).max()
            assert err < tol, (prec, err, float(np.std(b)))
        assert eng.pending_frames(s) == 500 - 384 and eng.emitted_steps(s) == 48
        eng.close()


def test_qwen_error_contract():
    from whisperlivekit_b200._lib import WlkError
    from whisperlivekit_b200.qwen_engine import QwenTowerEngine
    g, dims, sd, mels, sched = qwen_case("qnano")
    eng = QwenTowerEngine(dims, None, precision="fp32", max_sessions=1, max_batch=1)
    s = eng.open_session()
    with pytest.raises(WlkError, match="not finalized"):
        eng.forward_chunk([s], [mels[:200]])
    with pytest.raises(WlkError, match="missing"):
        eng.load_state_dict({k: v for k, v in sd.items() if k != "proj2.bias"})
    with pytest.raises(WlkError, match="wrong shape"):
        eng.load_state_dict({"proj1.weight": np.zeros((3, 3), np.float32)})
    with pytest.raises(WlkError, match="in use"):
        eng.open_session()
    with pytest.raises(WlkError, match="invalid session"):
        eng.forward_chunk([5], [mels[:8]])
    eng.close()


def test_device_mel_front_end_matches_reference_extractor():
    """wlk_qwen_append_audio (StreamingMelExtractor.append / flush on the device) against fixtures recorded from the
    reference extractor over the Hugging Face featurizer: same frames per call, values within 2e-4 (fp32 DFT)."""
    from test_oracle_qwen_mel import check_mel_stream, mel_case
    from whisperlivekit_b200.qwen_engine import QwenTowerEngine
    g, audio, sched = mel_case()
    dims = QWEN_DIMS["qnano"]
    eng = QwenTowerEngine(dims, None, precision="fp32", max_sessions=2, max_batch=2)
    eng.load_mel_filters()
    s = eng.open_session()
    seen = [0]

    def append(a):
        m = eng.mel_append([s], [a])[0]
        seen[0] += m.shape[0]
        return m

    def flush():
        m = eng.mel_flush([s])[0]
        seen[0] += m.shape[0]
        return m

    check_mel_stream(append, flush, lambda: seen[0], g, audio, sched, 2e-4)
    eng.close()


def test_audio_to_tower_end_to_end_batched():
    """Raw audio of three streams, appended in different chunkings, through the device mel front end and the tower,
    against the oracle chain (StreamingMelOracle -> QwenTowerOracle) stream by stream."""
    from oracle.make_golden_qwen_mel import speechlike
    from oracle.qwen_mel_oracle import StreamingMelOracle
    from oracle.qwen_oracle import QwenTowerOracle
    from whisperlivekit_b200.qwen_engine import QwenTowerEngine
    from whisperlivekit_b200.weights import mel_filterbank
    dims = QWEN_DIMS["qnano"]
    sd = synthetic_tower_state_dict(dims, seed=11)
    eng = QwenTowerEngine(dims, sd, precision="fp32", max_sessions=3, max_batch=3)
    eng.load_mel_filters()
    audio = [speechlike(16000 * 7, seed=40 + i) for i in range(3)]
    chunk = [4000, 2560, 9000]
    sids = [eng.open_session() for _ in range(3)]
    orc = QwenTowerOracle(dims, sd)
    osid = [orc.open_session() for _ in range(3)]
    omel = [StreamingMelOracle(mel_filterbank(dims.n_mels)) for _ in range(3)]
    pos = [0, 0, 0]
    worst, rows = 0.0, 0
    for call in range(12):
        parts = []
        for i in range(3):
            parts.append(audio[i][pos[i]: pos[i] + chunk[i]])
            pos[i] += chunk[i]
        mels = eng.mel_append(sids, parts)
        got = eng.forward_chunk(sids, mels)
        for i in range(3):
            m = omel[i].append(parts[i])
            m = np.zeros((0, dims.n_mels), np.float32) if m is None else m
            assert mels[i].shape == m.shape
            if m.size:
                assert np.abs(mels[i] - m).max() < 2e-4
            ref = orc.forward_chunk([osid[i]], [m])[0]
            assert got[i].shape == ref.shape
            if ref.size:
                worst = max(worst, float(np.abs(got[i] - ref).max()))
                rows += ref.shape[0]
    assert rows >= 90 and worst < 2e-3, (rows, worst)
    eng.close()
