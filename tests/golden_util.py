"""Helpers shared by the CPU (oracle) and GPU (engine) parity tests."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    return dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))


def case_setup(name):
    """-> (golden dict, dims, state_dict, audio, align_heads)"""
    from whisperlivekit_b200.dims import ModelDimensions
    from whisperlivekit_b200.weights import synthetic_state_dict, synthetic_audio
    g = load_case(name)
    dims = ModelDimensions(*[int(x) for x in g["dims"]])
    sd = synthetic_state_dict(dims, seed=int(g["weight_seed"]))
    audio = synthetic_audio(float(g["audio_seconds"]), seed=int(g["audio_seed"]))
    heads = [tuple(int(v) for v in r) for r in g["align_heads"]]
    return g, dims, sd, audio, heads


def sampled_diff(g, key, arr):
    """max |golden sample - arr| over the strided sample stored for `key`."""
    a = np.asarray(arr, np.float32)
    assert tuple(a.shape) == tuple(int(x) for x in g[key + "__shape"]), (key, a.shape, g[key + "__shape"])
    got = a.reshape(-1)[g[key + "__idx"]]
    ref = g[key + "__val"]
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), fin), key
    return float(np.abs(got[fin] - ref[fin]).max()), float(g[key + "__absmax"])


def run_policy(engine, audio, nonspeech_prob, chunk=8000):
    """Drive StreamingAlignAtt over 0.5 s chunks exactly like oracle/make_golden.py
    drives the reference's AlignAtt.  -> dict of flat traces."""
    from whisperlivekit_b200.alignatt import AlignAttConfig, StreamingAlignAtt
    pol = StreamingAlignAtt(engine, AlignAttConfig(nonspeech_prob=nonspeech_prob))
    new_tokens, step_tokens, step_frames, offs_t, offs_s = [], [], [], [0], [0]
    n_chunks = int(np.ceil(len(audio) / chunk))
    for c in range(n_chunks):
        pol.insert_audio(audio[c * chunk:(c + 1) * chunk])
        tr = pol.infer(is_last=(c == n_chunks - 1))
        new_tokens += tr.new_tokens
        offs_t.append(len(new_tokens))
        step_tokens += tr.step_tokens
        step_frames += tr.step_frames
        offs_s.append(len(step_tokens))
    pol.close()
    return dict(new_tokens=np.asarray(new_tokens), new_tokens_offsets=np.asarray(offs_t),
                step_tokens=np.asarray(step_tokens), step_frames=np.asarray(step_frames),
                step_offsets=np.asarray(offs_s))
