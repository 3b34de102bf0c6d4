"""CohortRunner (whisperlivekit_b200/cohort.py): many policies advanced in lockstep by one thread, every round one
batched engine call -- the emitted tokens / frames must equal the same policies each driven alone (CPU oracle engine),
and the reference's own traces (golden fixtures)."""
import numpy as np

from golden_util import case_setup
from whisperlivekit_b200.alignatt import AlignAttConfig, StreamingAlignAtt
from whisperlivekit_b200.cohort import CohortRunner


def _trace(tr):
    return (tr.stop, tuple(tr.new_tokens), tuple(tr.step_tokens), tuple(tr.step_frames), tr.content_mel_len, tr.prefix_len)


def test_cohort_equals_policies_driven_alone_and_reference():
    from oracle import whisper_oracle as wo
    g, dims, sd, audio, heads = case_setup("micro")
    eng = wo.OracleEngine(dims, sd, heads)
    eng.max_batch = 3                                           # forces the runner to split a round into two engine calls
    n, n_chunks = 5, int(np.ceil(len(audio) / 8000))
    offs = [0, 1600, 0, 4000, 800]                              # streams see different audio (ragged windows)

    def feed(i, c):
        a = audio[offs[i]:]
        return a[c * 8000:(c + 1) * 8000]

    alone = []
    for i in range(n):
        p = StreamingAlignAtt(eng, AlignAttConfig(nonspeech_prob=1.01))
        tr = []
        for c in range(n_chunks):
            seg = feed(i, c)
            if len(seg):
                p.insert_audio(seg)
            tr.append(_trace(p.infer(is_last=(c == n_chunks - 1))))
        alone.append(tr)
        p.close()
    pols = [StreamingAlignAtt(eng, AlignAttConfig(nonspeech_prob=1.01)) for _ in range(n)]
    runner = CohortRunner(eng)
    together = [[] for _ in range(n)]
    for c in range(n_chunks):
        for i, p in enumerate(pols):
            seg = feed(i, c)
            if len(seg):
                p.insert_audio(seg)
        for i, tr in enumerate(runner.run(pols, is_last=(c == n_chunks - 1))):
            together[i].append(_trace(tr))
    assert together == alone
    assert runner.stats["sessions"] / runner.stats["calls"] > 1.5          # rounds really were batched
    # stream 0 is the fixture's stream: the reference's own token / frame trace
    steps = [t for tr in together[0] for t in tr[2]]
    frames = [f for tr in together[0] for f in tr[3]]
    assert steps == list(g["pol_step_tokens"]) and frames == list(g["pol_step_frames"])


def test_continuous_admission_equals_policies_driven_alone():
    """The open form: streams are admitted between rounds (while others are mid-iteration: encode / prefill of the newcomers
    interleave with the token steps of the running ones and share their rounds afterwards); every stream's trace must still
    equal the stream driven alone."""
    from oracle import whisper_oracle as wo
    g, dims, sd, audio, heads = case_setup("micro")
    eng = wo.OracleEngine(dims, sd, heads)
    n, n_chunks = 4, int(np.ceil(len(audio) / 8000))
    offs = [0, 2400, 800, 0]

    def feed(i, c):
        a = audio[offs[i]:]
        return a[c * 8000:(c + 1) * 8000]

    alone = []
    for i in range(n):
        p = StreamingAlignAtt(eng, AlignAttConfig(nonspeech_prob=1.01))
        tr = []
        for c in range(n_chunks):
            seg = feed(i, c)
            if len(seg):
                p.insert_audio(seg)
            tr.append(_trace(p.infer(is_last=(c == n_chunks - 1))))
        alone.append(tr)
        p.close()
    pols = [StreamingAlignAtt(eng, AlignAttConfig(nonspeech_prob=1.01)) for _ in range(n)]
    runner = CohortRunner(eng)
    got = [[] for _ in range(n)]
    nxt = [0] * n                                               # next chunk of every stream
    delay = [0, 1, 3, 6]                                        # stream i may start a chunk only `delay` rounds after the previous one
    wait = list(delay)
    rounds = 0
    while any(c < n_chunks for c in nxt) or runner.busy():
        for i in range(n):                                      # admit whoever is due and idle
            if nxt[i] < n_chunks and i not in runner._pending and wait[i] <= 0:
                seg = feed(i, nxt[i])
                if len(seg):
                    pols[i].insert_audio(seg)
                tr = runner.admit(i, pols[i], is_last=(nxt[i] == n_chunks - 1))
                if tr is not None:
                    got[i].append(_trace(tr)); nxt[i] += 1; wait[i] = delay[i]
        for i, tr in runner.round():
            got[i].append(_trace(tr)); nxt[i] += 1; wait[i] = delay[i]
        wait = [w - 1 for w in wait]
        rounds += 1
        assert rounds < 5000
    assert got == alone
    import pytest
    with pytest.raises(ValueError):
        runner.admit(0, pols[0]); runner.admit(0, pols[0])
