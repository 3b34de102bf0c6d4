"""Cohort runner: many AlignAtt policies, one thread (SURVEY.md section 8f item 1 -- "replace thread-per-call with an
awaitable submit/complete so 512-1024 sessions do not need 512-1024 OS threads; keep process_iter semantics").

``StreamingAlignAtt.infer_steps`` is the policy iteration (reference align_att_base.py:174-322) as a generator that
yields its engine requests.  ``CohortRunner.run`` advances a set of policies in lockstep: each round, the pending
requests of one kind are served by ONE batched engine call (``wlk_encode`` / ``wlk_decode`` / ``wlk_no_speech_prob`` /
``wlk_select`` over all of them) and every policy is resumed with its own result.  Policies stop at their own pace (the
step count is the policy's decision) and simply drop out of the cohort.  Compared with ``batching.BatchingEngine``
(which keeps WhisperLiveKit's thread-per-session call surface) there is no thread hand-off and no GIL convoy between
rounds: a round costs the engine call plus ~20 us of Python per policy.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

_ORDER = ("encode", "decode", "no_speech", "select")


class CohortRunner:
    def __init__(self, engine, max_batch: int = 0):
        self.engine = engine
        self.max_batch = int(max_batch or getattr(engine, "max_batch", 64))
        self.stats = dict(rounds=0, calls=0, sessions=0, cohorts=0, cohort_sessions=0)

    def run(self, policies: Sequence, is_last: bool = False) -> List:
        """One ``infer()`` for every policy (all on this runner's engine).  -> their InferTrace objects, in order."""
        eng = self.engine
        traces: List = [None] * len(policies)
        gens: Dict[int, object] = {}
        pending: Dict[int, tuple] = {}
        for i, p in enumerate(policies):
            g = p.infer_steps(is_last)
            try:
                pending[i] = next(g)
                gens[i] = g
            except StopIteration as stop:
                traces[i] = stop.value
        self.stats["cohorts"] += 1
        self.stats["cohort_sessions"] += len(pending)

        def resume(i, res):
            try:
                pending[i] = gens[i].send(res)
            except StopIteration as stop:
                traces[i] = stop.value
                del pending[i]

        while pending:
            self.stats["rounds"] += 1
            kinds = {r[0] for r in pending.values()}
            kind = next(k for k in _ORDER if k in kinds)
            idx = [i for i, r in pending.items() if r[0] == kind]
            for lo in range(0, len(idx), self.max_batch):
                grp = idx[lo: lo + self.max_batch]
                sids = [policies[i].sid for i in grp]
                self.stats["calls"] += 1
                self.stats["sessions"] += len(grp)
                if kind == "encode":
                    out = eng.encode(sids)
                elif kind == "decode":
                    eng.decode(sids, [pending[i][1] for i in grp], sot_index=policies[grp[0]].sot_index)
                    out = [None] * len(grp)
                elif kind == "no_speech":
                    out = eng.no_speech_prob(sids)
                else:
                    p0 = policies[grp[0]]
                    out = eng.select(sids, p0.suppress_tokens, [p0.sp.blank, p0.sp.eot], [pending[i][1] for i in grp],
                                     [pending[i][2] for i in grp], window_iters=16)
                for i, res in zip(grp, out):
                    resume(i, res)
        return traces
