"""CPU ORACLE (test infrastructure only) for the Silero VAD forward of the ingest step.

The reference runs the vendored TorchScript model ``whisperlivekit/silero_vad_models/silero_vad.jit`` one 512-sample
window at a time per stream (whisperlivekit/silero_vad_iterator.py:20-29 ``init_jit_model``, :164-178 ``load_jit_vad``,
:288-331 ``FixedVADIterator``).  Restated here from the scripted module's own code (``model._model.code`` and
submodules, 16 kHz branch):
  context   the last 64 samples of the previous window are prepended (576 samples in)
  stft      reflect-pad 64 on the right only, conv1d with the [258, 1, 256] basis at hop 128 -> 4 frames,
            magnitude sqrt(re^2 + im^2) over the first / second 129 channels
  encoder   four Conv1d(k = 3, pad 1) + ReLU blocks: 129 -> 128 (stride 1), 128 -> 64 (stride 2), 64 -> 64 (stride 2),
            64 -> 128 (stride 1): 4 -> 4 -> 2 -> 1 -> 1 frames
  decoder   LSTMCell(128, 128) with the stream's (h, c), then ReLU -> Conv1d(128, 1, k = 1) -> sigmoid
Status: the CUDA engine for this row is not written yet; this module and tests/golden/vad.npz (recorded from the
reference's scripted model with seeded weights of the same shapes, oracle/make_golden_vad.py) are the parity anchor.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

WINDOW, CONTEXT = 512, 64


def synthetic_vad_state_dict(seed: int = 0) -> Dict[str, np.ndarray]:
    """Seeded weights with the scripted model's parameter names / shapes (16 kHz branch), sized so that the
    probabilities spread over (0, 1) instead of saturating."""
    rng = np.random.default_rng(seed)
    n = np.arange(256)
    win = np.sqrt(0.5 - 0.5 * np.cos(2 * np.pi * n / 256))                      # any fixed analysis basis will do
    k = np.arange(129)[:, None]
    basis = np.concatenate([np.cos(2 * np.pi * k * n / 256), -np.sin(2 * np.pi * k * n / 256)], 0) * win
    sd = {"_model.stft.forward_basis_buffer": basis[:, None, :].astype(np.float32)}

    def conv(i, co, ci):
        sd[f"_model.encoder.{i}.reparam_conv.weight"] = (rng.standard_normal((co, ci, 3)) / np.sqrt(3 * ci) * 1.4).astype(np.float32)
        sd[f"_model.encoder.{i}.reparam_conv.bias"] = (rng.standard_normal(co) * 0.05).astype(np.float32)

    conv(0, 128, 129); conv(1, 64, 128); conv(2, 64, 64); conv(3, 128, 64)
    sd["_model.decoder.rnn.weight_ih"] = (rng.standard_normal((512, 128)) / np.sqrt(128)).astype(np.float32)
    sd["_model.decoder.rnn.weight_hh"] = (rng.standard_normal((512, 128)) / np.sqrt(128)).astype(np.float32)
    sd["_model.decoder.rnn.bias_ih"] = (rng.standard_normal(512) * 0.1).astype(np.float32)
    sd["_model.decoder.rnn.bias_hh"] = (rng.standard_normal(512) * 0.1).astype(np.float32)
    sd["_model.decoder.decoder.2.weight"] = (rng.standard_normal((1, 128, 1)) * 4.0).astype(np.float32)
    sd["_model.decoder.decoder.2.bias"] = np.zeros(1, np.float32)
    return sd


class VadOracle:
    """Per-stream Silero VAD state (context, h, c) and the forward of one 512-sample window per stream."""

    def __init__(self, state_dict: Dict[str, np.ndarray]):
        self.W = {k.replace("_model.", "", 1): torch.from_numpy(np.ascontiguousarray(v, np.float32))
                  for k, v in state_dict.items() if k.startswith("_model.")}
        self._s: Dict[int, dict] = {}
        self._next = 0

    def open_session(self) -> int:
        sid = self._next
        self._next += 1
        self.reset_session(sid)
        return sid

    def reset_session(self, sid: int) -> None:
        self._s[sid] = dict(context=torch.zeros(CONTEXT), h=torch.zeros(128), c=torch.zeros(128))

    def close_session(self, sid: int) -> None:
        self._s.pop(sid)

    @torch.no_grad()
    def forward(self, sids, windows) -> np.ndarray:
        """windows [n, 512] fp32 -> speech probabilities [n]."""
        W = self.W
        out = []
        for sid, w in zip(sids, windows):
            s = self._s[sid]
            x = torch.cat([s["context"], torch.from_numpy(np.ascontiguousarray(w, np.float32).reshape(WINDOW))])
            s["context"] = x[-CONTEXT:].clone()
            xp = F.pad(x[None, None], (0, 64), mode="reflect")
            ft = F.conv1d(xp, W["stft.forward_basis_buffer"], stride=128)           # [1, 258, 4]
            y = torch.sqrt(ft[:, :129] ** 2 + ft[:, 129:] ** 2)
            for i, stride in enumerate((1, 2, 2, 1)):
                y = F.relu(F.conv1d(y, W[f"encoder.{i}.reparam_conv.weight"], W[f"encoder.{i}.reparam_conv.bias"],
                                    stride=stride, padding=1))
            gates = (F.linear(y[0, :, 0], W["decoder.rnn.weight_ih"], W["decoder.rnn.bias_ih"])
                     + F.linear(s["h"], W["decoder.rnn.weight_hh"], W["decoder.rnn.bias_hh"]))
            i_g, f_g, g_g, o_g = gates.chunk(4)                                      # torch LSTMCell gate order
            s["c"] = torch.sigmoid(f_g) * s["c"] + torch.sigmoid(i_g) * torch.tanh(g_g)
            s["h"] = torch.sigmoid(o_g) * torch.tanh(s["c"])
            logit = F.linear(F.relu(s["h"]), W["decoder.decoder.2.weight"][:, :, 0], W["decoder.decoder.2.bias"])
            out.append(float(torch.sigmoid(logit)[0]))
        return np.asarray(out, np.float32)
