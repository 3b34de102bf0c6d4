"""CPU ORACLE (test infrastructure only) for the incremental log-mel front end of the Qwen3 streaming backend.

Restates reference third_party/qwen3-asr-causal/src/qwen3_asr_causal/features.py:32-112 (StreamingMelExtractor:
append / _emit / flush and its sample-window bookkeeping) and the featurizer it calls, the Hugging Face
``WhisperFeatureExtractor`` numpy path (transformers 4.57 / 5.x feature_extraction_whisper.py
``_np_extract_fbank_features`` over ``audio_utils.spectrogram``; not under /root/reference -- its published
algorithm: reflect-pad n_fft/2, periodic Hann 400, hop 160, |rfft|^2, Slaney mel filterbank (128 x 201), log10 with
floor 1e-10, last frame dropped, clamp to (max - 8), (x + 4) / 4).
Pinned by tests/test_qwen_mel_reference.py (build container: the reference extractor over the real HF featurizer) and
tests/golden/qwen_mel.npz recorded from it by oracle/make_golden_qwen_mel.py.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

HOP, N_FFT, STFT_PAD, MARGIN_FRAMES = 160, 400, 200, 8


def window_log_mel(samples: np.ndarray, filters: np.ndarray) -> np.ndarray:
    """[frames, n_mels] features of one raw sample window (frames = len // 160)."""
    x = np.asarray(samples, np.float64)
    xp = np.pad(x, (STFT_PAD, STFT_PAD), mode="reflect")
    n_frames = 1 + len(x) // HOP
    win = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(N_FFT) / N_FFT)
    idx = np.arange(N_FFT)[None, :] + HOP * np.arange(n_frames)[:, None]
    spec = np.fft.rfft(xp[idx] * win[None, :], axis=1).astype(np.complex64)      # HF stores the spectrum as complex64
    power = np.abs(spec, dtype=np.float64) ** 2
    mel = np.maximum(1e-10, power @ np.asarray(filters, np.float64).T)
    log_spec = np.log10(mel)[:-1]                                              # drop the last frame
    if log_spec.shape[0] == 0:
        return np.zeros((0, filters.shape[0]), np.float32)
    log_spec = np.maximum(log_spec, log_spec.max() - 8.0)
    return ((log_spec + 4.0) / 4.0).astype(np.float32)


class StreamingMelOracle:
    """features.py:32-112 with ``window_log_mel`` as the featurizer."""

    def __init__(self, filters: np.ndarray):
        self.filters = np.asarray(filters, np.float32)
        self.reset()

    def reset(self) -> None:
        self._buffer = np.zeros(0, np.float32)
        self._buffer_start_frame = 0
        self.emitted_frames = 0
        self._total_samples = 0

    def _emit(self, upto_frame: int) -> Optional[np.ndarray]:                    # features.py:62-84
        if upto_frame <= self.emitted_frames:
            return None
        local_first = self.emitted_frames - self._buffer_start_frame
        local_last = upto_frame - self._buffer_start_frame
        mel = window_log_mel(self._buffer, self.filters)
        if mel.shape[0] < local_last:
            local_last = int(mel.shape[0])
            upto_frame = self._buffer_start_frame + local_last
            if upto_frame <= self.emitted_frames:
                return None
        frames = mel[local_first:local_last]
        self.emitted_frames = upto_frame
        keep_from = max(self._buffer_start_frame, self.emitted_frames - MARGIN_FRAMES)
        cut = (keep_from - self._buffer_start_frame) * HOP
        if cut > 0:
            self._buffer = self._buffer[cut:]
            self._buffer_start_frame = keep_from
        return frames

    def append(self, audio: np.ndarray) -> Optional[np.ndarray]:                 # features.py:86-99
        audio = np.asarray(audio, np.float32).reshape(-1)
        if audio.size:
            self._buffer = np.concatenate([self._buffer, audio])
            self._total_samples += int(audio.size)
        if self._total_samples < STFT_PAD + 1:
            return None
        ready = (self._total_samples - STFT_PAD) // HOP + 1
        ready = min(ready, self._total_samples // HOP)
        return self._emit(ready)

    def flush(self) -> Optional[np.ndarray]:                                     # features.py:101-110
        total_frames = self._total_samples // HOP
        if total_frames <= self.emitted_frames:
            return None
        min_samples = 2 * STFT_PAD + 1
        if self._buffer.size < min_samples:
            self._buffer = np.pad(self._buffer, (0, min_samples - self._buffer.size))
        return self._emit(total_frames)
