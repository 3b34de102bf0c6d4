"""CPU restatement (numpy) of the diarization post-processing step -- test infrastructure only.

``process_predictions`` follows SortformerDiarizationOnline._process_predictions
(reference whisperlivekit/diarization/sortformer_backend.py:313-363) line by line; ``resolve_max_speakers`` follows
``_resolve_max_speakers`` (:135-148).  Pinned by the reference's own known-answer tests
(/root/reference/tests/test_sortformer_max_speakers.py:78-125, 183-215), restated as cases in
tests/test_diarization.py, and -- in the build container -- by running the reference's method itself side by side.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


def resolve_max_speakers(max_speakers: Optional[int], model_speakers: int) -> int:
    """sortformer_backend.py:135-148"""
    if model_speakers < 1:
        raise ValueError("The Sortformer checkpoint exposes no speaker channels.")
    if max_speakers is None:
        return model_speakers
    if isinstance(max_speakers, bool) or not isinstance(max_speakers, int):
        raise ValueError("max_speakers must be an integer.")
    if not 1 <= max_speakers <= model_speakers:
        raise ValueError(f"max_speakers must be between 1 and {model_speakers} for the loaded Sortformer checkpoint.")
    return max_speakers


def frame_segments(preds: np.ndarray, max_speakers: int, len_prediction: Optional[int]) -> Tuple[List[Tuple[int, int, int]], int]:
    """The integer core: (speaker, first frame, end frame) runs of the last `len_prediction` frames.
    -> (segments, len_prediction actually used)   sortformer_backend.py:316-341"""
    preds = np.asarray(preds, np.float32)
    if preds.shape[1] < max_speakers:
        raise RuntimeError("Sortformer returned fewer speaker channels than configured.")
    active = np.argmax(preds[:, :max_speakers], axis=1)                      # :326-327
    if not len(active):
        return [], (len_prediction or 0)
    if len_prediction is None:
        len_prediction = len(active)                                          # :332-333
    cur = active[-len_prediction:]                                            # :336
    segs, start, spk = [], 0, int(cur[0])
    for idx, s in enumerate(cur):                                             # :344-356
        if int(s) != spk:
            segs.append((spk, start, idx))
            start, spk = idx, int(s)
    segs.append((spk, start, len(cur)))                                       # :357-363
    return segs, len_prediction


def process_predictions(preds: np.ndarray, max_speakers: int, len_prediction: Optional[int], chunk_index: int,
                        chunk_duration_seconds: float, global_time_offset: float = 0.0):
    """-> ([(speaker, start_s, end_s)], len_prediction)   with the reference's rounding (:335, :343-361)"""
    segs, lp = frame_segments(preds, max_speakers, len_prediction)
    if not segs:
        return [], lp
    frame_duration = chunk_duration_seconds / lp
    base_time = chunk_index * chunk_duration_seconds + global_time_offset
    out = []
    for k, (spk, a, b) in enumerate(segs):
        start = round(base_time, 2) if k == 0 else round(base_time + a * frame_duration, 2)
        out.append((spk, start, round(base_time + b * frame_duration, 2)))
    return out, lp
