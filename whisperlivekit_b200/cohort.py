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
    """``run(policies)`` advances a closed set to completion.  ``admit`` / ``round`` are the open form (continuous batching
    at the granularity of policy requests): a host admits the streams whose chunk has arrived BETWEEN rounds, their encode /
    prefill requests are served next (the order below), and from then on they share the token-step rounds of the streams
    that were already running -- with staggered arrivals that roughly halves the number of latency-bound step rounds."""

    def __init__(self, engine, max_batch: int = 0):
        self.engine = engine
        self.max_batch = int(max_batch or getattr(engine, "max_batch", 64))
        self.stats = dict(rounds=0, calls=0, sessions=0, cohorts=0, cohort_sessions=0)
        self._pol: Dict[object, object] = {}
        self._gens: Dict[object, object] = {}
        self._pending: Dict[object, tuple] = {}

    # -- open form -------------------------------------------------------------------------------------------------
    def admit(self, key, policy, is_last: bool = False):
        """Start ``policy``'s iteration under ``key``.  -> its InferTrace if it ended without a single engine request."""
        if key in self._pending:
            raise ValueError(f"{key!r} is already in flight (one infer per stream at a time)")
        g = policy.infer_steps(is_last)
        try:
            self._pending[key] = next(g)
        except StopIteration as stop:
            return stop.value
        self._gens[key], self._pol[key] = g, policy
        return None

    def admit_many(self, items, is_last: bool = False):
        """items: iterable of (key, policy).  -> [(key, trace)] of the ones that finished immediately."""
        done = []
        n = 0
        for key, pol in items:
            n += 1
            tr = self.admit(key, pol, is_last)
            if tr is not None:
                done.append((key, tr))
        self.stats["cohorts"] += 1
        self.stats["cohort_sessions"] += n
        return done

    def busy(self) -> bool:
        return bool(self._pending)

    def round(self):
        """Serve the pending requests of ONE kind with batched engine calls.  -> [(key, trace)] finished in this round."""
        eng, pending = self.engine, self._pending
        if not pending:
            return []
        self.stats["rounds"] += 1
        kinds = {r[0] for r in pending.values()}
        kind = next(k for k in _ORDER if k in kinds)
        idx = [k for k, r in pending.items() if r[0] == kind]
        finished = []
        for lo in range(0, len(idx), self.max_batch):
            grp = idx[lo: lo + self.max_batch]
            pols = [self._pol[k] for k in grp]
            sids = [p.sid for p in pols]
            self.stats["calls"] += 1
            self.stats["sessions"] += len(grp)
            if kind == "encode":
                out = eng.encode(sids)
            elif kind == "decode":
                eng.decode(sids, [pending[k][1] for k in grp], sot_index=pols[0].sot_index)
                out = [None] * len(grp)
            elif kind == "no_speech":
                out = eng.no_speech_prob(sids)
            else:
                p0 = pols[0]
                out = eng.select(sids, p0.suppress_tokens, [p0.sp.blank, p0.sp.eot], [pending[k][1] for k in grp],
                                 [pending[k][2] for k in grp], window_iters=16)
            for k, res in zip(grp, out):
                try:
                    pending[k] = self._gens[k].send(res)
                except StopIteration as stop:
                    finished.append((k, stop.value))
                    del pending[k], self._gens[k], self._pol[k]
        return finished

    # -- closed form -----------------------------------------------------------------------------------------------
    def run(self, policies: Sequence, is_last: bool = False) -> List:
        """One ``infer()`` for every policy (all on this runner's engine).  -> their InferTrace objects, in order."""
        if self._pending:
            raise RuntimeError("run() on a runner that has streams in flight")
        traces: List = [None] * len(policies)
        for i, tr in self.admit_many(list(enumerate(policies)), is_last):
            traces[i] = tr
        while self._pending:
            for i, tr in self.round():
                traces[i] = tr
        return traces
