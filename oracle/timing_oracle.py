"""CPU ORACLE (test infrastructure only) for the word-timestamp math of the LocalAgreement path.

Restates reference whisperlivekit/whisper/timing.py:
  median_filter  :19-54   (reflect-padded sliding median along the last axis)
  dtw_cpu        :81-105  (cost recursion, float32 cost table, strict-less move preference)
  backtrace      :57-79
Pinned on the reference's own numba/torch implementations by oracle/make_golden.py -> tests/golden/timing.npz.
"""
from __future__ import annotations

import numpy as np


def median_filter(x: np.ndarray, filter_width: int) -> np.ndarray:
    pad = filter_width // 2
    if x.shape[-1] <= pad:
        return x
    xp = np.pad(x, [(0, 0)] * (x.ndim - 1) + [(pad, pad)], mode="reflect")
    win = np.lib.stride_tricks.sliding_window_view(xp, filter_width, axis=-1)
    return np.sort(win, axis=-1)[..., pad]


def dtw(x: np.ndarray):
    """x[N tokens, M frames] (the reference passes -attention as float64) -> (text_idx, time_idx)."""
    x = np.asarray(x, dtype=np.float64)
    N, M = x.shape
    cost = np.full((N + 1, M + 1), np.inf, dtype=np.float32)
    trace = -np.ones((N + 1, M + 1), dtype=np.int8)
    cost[0, 0] = 0
    for j in range(1, M + 1):
        for i in range(1, N + 1):
            c0, c1, c2 = cost[i - 1, j - 1], cost[i - 1, j], cost[i, j - 1]
            if c0 < c1 and c0 < c2:
                c, t = c0, 0
            elif c1 < c0 and c1 < c2:
                c, t = c1, 1
            else:
                c, t = c2, 2
            cost[i, j] = x[i - 1, j - 1] + c
            trace[i, j] = t
    i, j = N, M
    trace[0, :] = 2
    trace[:, 0] = 1
    res = []
    while i > 0 or j > 0:
        res.append((i - 1, j - 1))
        if trace[i, j] == 0:
            i -= 1; j -= 1
        elif trace[i, j] == 1:
            i -= 1
        else:
            j -= 1
    r = np.array(res)[::-1, :].T
    return r[0].astype(np.int64), r[1].astype(np.int64)
