"""Geometry of the streaming Sortformer diarizer and seeded weights under NeMo's
parameter names.

The reference wraps NeMo's ``SortformerEncLabelModel`` (whisperlivekit/diarization/sortformer_backend.py:68-128) and sets
the streaming parameters at :120-126 (chunk_len 10, left context 10, spkcache 188, fifo 188, update period 144).  NeMo
(3.0.0 in the reference's uv.lock) and the checkpoint ``nvidia/diar_streaming_sortformer_4spk-v2`` are absent from both
containers: the geometry below is the public model card / config of that checkpoint (FastConformer 17 x 512 with 8x
depthwise-striding subsampling, 18 x 192 post-LN Transformer, 4 sigmoid speaker outputs) and weights are seeded.
"""
from __future__ import annotations

from dataclasses import asdict, dataclass
from typing import Dict

import numpy as np


@dataclass(frozen=True)
class SortformerDims:
    n_mels: int = 128            # AudioToMelSpectrogramPreprocessor(features=128), sortformer_backend.py:181-188
    n_fft: int = 512
    win_length: int = 400        # window_size 0.025 s
    hop: int = 160               # window_stride 0.01 s
    conv_channels: int = 256     # ConvSubsampling(dw_striding, factor 8)
    d_model: int = 512           # FastConformer d_model == sortformer_modules.fc_d_model
    n_head: int = 8
    n_layer: int = 17
    ff_mult: int = 4
    conv_kernel: int = 9
    tf_d_model: int = 192        # sortformer_modules.tf_d_model
    tf_n_head: int = 8
    tf_n_layer: int = 18
    tf_inner: int = 768
    n_spk: int = 4
    # streaming parameters (sortformer_backend.py:120-126) and the module defaults they sit on
    spkcache_len: int = 188
    fifo_len: int = 188
    spkcache_update_period: int = 144
    chunk_len: int = 10
    subsampling_factor: int = 10         # sortformer_modules.subsampling_factor as the reference sets it (chunk duration)
    encoder_subsampling: int = 8         # encoder.subsampling_factor (lc / rc in forward_streaming_step)
    spkcache_sil_frames_per_spk: int = 3
    pred_score_threshold: float = 0.25
    scores_boost_latest: float = 0.05
    sil_threshold: float = 0.2
    strong_boost_rate: float = 0.75
    weak_boost_rate: float = 1.5
    min_pos_scores_rate: float = 0.5
    max_index: int = 99999

    @property
    def d_head(self) -> int:
        return self.d_model // self.n_head

    @property
    def tf_d_head(self) -> int:
        return self.tf_d_model // self.tf_n_head

    @property
    def sub_freq(self) -> int:
        """frequency bins left after three stride-2 convolutions"""
        f = self.n_mels
        for _ in range(3):
            f = (f + 2 - 3) // 2 + 1
        return f

    def as_dict(self):
        return asdict(self)


SORTFORMER_DIMS: Dict[str, SortformerDims] = {
    "diar_streaming_sortformer_4spk-v2": SortformerDims(),
    # small geometries for the parity tests (same structure, every code path taken)
    "micro": SortformerDims(n_mels=32, conv_channels=16, d_model=64, n_head=2, n_layer=2, tf_d_model=48, tf_n_head=2,
                            tf_n_layer=2, tf_inner=96, spkcache_len=24, fifo_len=20, spkcache_update_period=14,
                            spkcache_sil_frames_per_spk=1),
    "small": SortformerDims(n_mels=64, conv_channels=32, d_model=128, n_head=4, n_layer=3, tf_d_model=96, tf_n_head=4,
                            tf_n_layer=3, tf_inner=192, spkcache_len=48, fifo_len=40, spkcache_update_period=30),
}


def subsampled_len(t: int) -> int:
    """ConvSubsampling.calc_length: three (k3, s2, p1) stages, floor mode"""
    for _ in range(3):
        t = (t + 2 - 3) // 2 + 1
    return t


def synthetic_sortformer_state_dict(d: SortformerDims, seed: int = 0) -> Dict[str, np.ndarray]:
    """Seeded fp32 weights under NeMo's state-dict names, scaled so that activations stay O(1) through the stack and the
    sigmoid outputs spread over (0, 1) (a cache compression then has speech, silence and ties to sort out)."""
    rng = np.random.default_rng(seed)
    sd: Dict[str, np.ndarray] = {}

    def lin(name, n_out, n_in, bias=True, gain=1.0):
        sd[name + ".weight"] = (gain * rng.standard_normal((n_out, n_in)) / np.sqrt(n_in)).astype(np.float32)
        if bias:
            sd[name + ".bias"] = (0.05 * rng.standard_normal(n_out)).astype(np.float32)

    def norm(name, n):
        sd[name + ".weight"] = (1.0 + 0.1 * rng.standard_normal(n)).astype(np.float32)
        sd[name + ".bias"] = (0.05 * rng.standard_normal(n)).astype(np.float32)

    C, D, F = d.conv_channels, d.d_model, d.sub_freq
    p = "encoder.pre_encode."
    sd[p + "conv.0.weight"] = (rng.standard_normal((C, 1, 3, 3)) / 3.0).astype(np.float32)
    sd[p + "conv.0.bias"] = (0.05 * rng.standard_normal(C)).astype(np.float32)
    for dw, pw in ((2, 3), (5, 6)):
        sd[p + f"conv.{dw}.weight"] = (rng.standard_normal((C, 1, 3, 3)) / 3.0).astype(np.float32)
        sd[p + f"conv.{dw}.bias"] = (0.05 * rng.standard_normal(C)).astype(np.float32)
        sd[p + f"conv.{pw}.weight"] = (1.4 * rng.standard_normal((C, C, 1, 1)) / np.sqrt(C)).astype(np.float32)
        sd[p + f"conv.{pw}.bias"] = (0.05 * rng.standard_normal(C)).astype(np.float32)
    lin(p + "out", D, C * F, gain=0.1)                      # the log-mel input is O(10); xscaling multiplies by sqrt(D)
    for i in range(d.n_layer):
        q = f"encoder.layers.{i}."
        for ff in ("feed_forward1", "feed_forward2"):
            norm(q + "norm_" + ff, D)
            lin(q + ff + ".linear1", d.ff_mult * D, D)
            lin(q + ff + ".linear2", D, d.ff_mult * D)
        norm(q + "norm_self_att", D)
        for nm in ("linear_q", "linear_k", "linear_v"):
            lin(q + "self_attn." + nm, D, D)
        lin(q + "self_attn.linear_out", D, D, gain=0.4)         # attention averages rows: keep the stream position-specific
        lin(q + "self_attn.linear_pos", D, D, bias=False)
        sd[q + "self_attn.pos_bias_u"] = (0.1 * rng.standard_normal((d.n_head, d.d_head))).astype(np.float32)
        sd[q + "self_attn.pos_bias_v"] = (0.1 * rng.standard_normal((d.n_head, d.d_head))).astype(np.float32)
        norm(q + "norm_conv", D)
        sd[q + "conv.pointwise_conv1.weight"] = (rng.standard_normal((2 * D, D, 1)) / np.sqrt(D)).astype(np.float32)
        sd[q + "conv.pointwise_conv1.bias"] = (0.05 * rng.standard_normal(2 * D)).astype(np.float32)
        sd[q + "conv.depthwise_conv.weight"] = (rng.standard_normal((D, 1, d.conv_kernel)) / np.sqrt(d.conv_kernel)).astype(np.float32)
        sd[q + "conv.depthwise_conv.bias"] = (0.05 * rng.standard_normal(D)).astype(np.float32)
        norm(q + "conv.batch_norm", D)
        sd[q + "conv.batch_norm.running_mean"] = (0.05 * rng.standard_normal(D)).astype(np.float32)
        sd[q + "conv.batch_norm.running_var"] = (0.3 + 0.2 * rng.random(D)).astype(np.float32)
        sd[q + "conv.pointwise_conv2.weight"] = (rng.standard_normal((D, D, 1)) / np.sqrt(D)).astype(np.float32)
        sd[q + "conv.pointwise_conv2.bias"] = (0.05 * rng.standard_normal(D)).astype(np.float32)
        norm(q + "norm_out", D)
    T = d.tf_d_model
    lin("sortformer_modules.encoder_proj", T, D)
    for i in range(d.tf_n_layer):
        q = f"transformer_encoder.layers.{i}."
        for nm in ("query_net", "key_net", "value_net"):
            lin(q + "first_sub_layer." + nm, T, T, gain=1.5)
        lin(q + "first_sub_layer.out_projection", T, T, gain=0.4)
        norm(q + "layer_norm_1", T)
        lin(q + "second_sub_layer.dense_in", d.tf_inner, T)
        lin(q + "second_sub_layer.dense_out", T, d.tf_inner)
        norm(q + "layer_norm_2", T)
    lin("sortformer_modules.first_hidden_to_hidden", T, T, gain=1.5)
    lin("sortformer_modules.single_hidden_to_spks", d.n_spk, T, gain=5.0)
    return sd


def synthetic_two_speaker_audio(seconds: float, seed: int = 0, sr: int = 16000) -> np.ndarray:
    """Speech-like test signal: two harmonic 'voices' (different pitch and formant tilt) taking turns every ~1.7 s with
    short silences between turns, plus low noise; fp32 in [-1, 1]."""
    rng = np.random.default_rng(seed)
    n = int(seconds * sr)
    t = np.arange(n) / sr
    out = np.zeros(n, np.float32)
    pos, spk = 0, 0
    while pos < n:
        dur = int(sr * (1.2 + rng.random()))
        gap = int(sr * (0.1 + 0.4 * rng.random()))
        seg = slice(pos, min(n, pos + dur))
        f0 = (110.0, 205.0)[spk] * (1.0 + 0.03 * np.sin(2 * np.pi * 3.1 * t[seg]))
        ph = 2 * np.pi * np.cumsum(f0) / sr
        tilt = (0.75, 0.55)[spk]
        v = sum((tilt ** k) * np.sin((k + 1) * ph) for k in range(8))
        env = 0.5 + 0.5 * np.sin(2 * np.pi * 4.0 * t[seg] + rng.random())
        out[seg] = (0.2 * v * env).astype(np.float32)
        pos += dur + gap
        spk ^= 1
    out += (0.003 * rng.standard_normal(n)).astype(np.float32)
    return np.clip(out, -1.0, 1.0)
