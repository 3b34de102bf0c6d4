"""Host side of the step after the diarization forward.

``DiarizationSegmenter`` mirrors the state ``SortformerDiarizationOnline`` keeps around ``_process_predictions``
(reference whisperlivekit/diarization/sortformer_backend.py:151-200, 313-363): ``max_speakers`` (resolved like
``_resolve_max_speakers``, :135-148), ``_len_prediction``, ``_chunk_index``, ``chunk_duration_seconds``,
``global_time_offset`` / ``insert_silence``.  The arithmetic -- argmax over the retained speaker channels and the
run-length encoding of the chunk's frames -- runs on the device for all streams of a call (``wlk_diar_segments``);
only (speaker, first frame, end frame) triples are copied back, never ``total_preds``.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _lib as L


@dataclass
class SpeakerSegment:
    """reference whisperlivekit/timed_objects.py:88-93"""
    speaker: int
    start: float
    end: float


def resolve_max_speakers(max_speakers: Optional[int], model_speakers: int) -> int:
    """reference sortformer_backend.py:135-148"""
    if model_speakers < 1:
        raise ValueError("The Sortformer checkpoint exposes no speaker channels.")
    if max_speakers is None:
        return model_speakers
    if isinstance(max_speakers, bool) or not isinstance(max_speakers, int):
        raise ValueError("max_speakers must be an integer.")
    if not 1 <= max_speakers <= model_speakers:
        raise ValueError(f"max_speakers must be between 1 and {model_speakers} for the loaded Sortformer checkpoint.")
    return max_speakers


def diar_segments(pred_ptrs: Sequence[int], n_frames_total: Sequence[int], len_prediction: Sequence[int], n_spk: int,
                  max_speakers: int, device: int = 0, max_seg: Optional[int] = None):
    """Batched device call.  pred_ptrs[i] = device address of stream i's fp32 [n_frames_total[i], n_spk] predictions.
    -> list (per stream) of [(speaker, first_frame, end_frame)]"""
    lib = L.load()
    n = len(pred_ptrs)
    nf = np.ascontiguousarray(n_frames_total, np.int32)
    lp = np.ascontiguousarray(len_prediction, np.int32)
    cap = int(max_seg or max(1, int(np.minimum(nf, lp).max())))
    ptrs = (C.c_void_p * n)(*[C.c_void_p(int(p)) for p in pred_ptrs])
    seg = np.zeros((n, cap, 3), np.int32)
    cnt = np.zeros(n, np.int32)
    L.check(lib.wlk_diar_segments(int(device), ptrs, nf.ctypes.data_as(C.c_void_p), lp.ctypes.data_as(C.c_void_p), n,
                                  int(n_spk), int(max_speakers), seg.ctypes.data_as(C.c_void_p),
                                  cnt.ctypes.data_as(C.c_void_p), cap))
    return [[(int(s), int(a), int(b)) for s, a, b in seg[i, : cnt[i]]] for i in range(n)]


class DiarizationSegmenter:
    """Per-stream state of the post-processing step; ``process`` = ``_process_predictions`` for one stream,
    ``process_batch`` serves many streams with one device call."""

    def __init__(self, model_speakers: int, chunk_duration_seconds: float, max_speakers: Optional[int] = None, device: int = 0):
        self.n_spk = int(model_speakers)
        self.max_speakers = resolve_max_speakers(max_speakers, self.n_spk)
        self.chunk_duration_seconds = float(chunk_duration_seconds)
        self.global_time_offset = 0.0
        self.device = int(device)
        self._chunk_index = 0
        self._len_prediction: Optional[int] = None

    def insert_silence(self, silence_duration: float) -> None:
        """reference sortformer_backend.py:236-245"""
        self.global_time_offset += silence_duration

    def _to_times(self, segs) -> List[SpeakerSegment]:
        if not segs:
            return []
        frame_duration = self.chunk_duration_seconds / self._len_prediction              # :335
        base_time = self._chunk_index * self.chunk_duration_seconds + self.global_time_offset   # :342
        out = []
        for k, (spk, a, b) in enumerate(segs):
            start = round(base_time, 2) if k == 0 else round(base_time + a * frame_duration, 2)
            out.append(SpeakerSegment(speaker=spk, start=start, end=round(base_time + b * frame_duration, 2)))
        return out

    @staticmethod
    def process_batch(segmenters: Sequence["DiarizationSegmenter"], pred_ptrs: Sequence[int], n_frames: Sequence[int]):
        """One device call for many streams (same model: n_spk, max_speakers and device are shared); advances every
        segmenter's chunk index like ``diarize()`` does (sortformer_backend.py:308-311)."""
        first = segmenters[0]
        for s, nfr in zip(segmenters, n_frames):
            if s._len_prediction is None and nfr > 0:
                s._len_prediction = int(nfr)                                              # :332-333
        lp = [s._len_prediction or 0 for s in segmenters]
        raw = diar_segments(pred_ptrs, n_frames, lp, first.n_spk, first.max_speakers, first.device)
        out = []
        for s, segs in zip(segmenters, raw):
            out.append(s._to_times(segs))
            s._chunk_index += 1
        return out

    def process(self, preds_dev_ptr: int, n_frames_total: int) -> List[SpeakerSegment]:
        return self.process_batch([self], [preds_dev_ptr], [n_frames_total])[0]
