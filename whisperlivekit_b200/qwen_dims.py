"""Geometry of the Qwen3-ASR audio tower as the causal streaming encoder runs it
(reference third_party/qwen3-asr-causal/src/qwen3_asr_causal/causal.py:60-140, config.py:17-45)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

import numpy as np


@dataclass(frozen=True)
class QwenTowerDims:
    n_mels: int = 128
    conv_channels: int = 480          # conv2d1/2/3 output channels (3x3, stride 2, pad 1; 8 mel frames -> 1 step)
    d_model: int = 896
    n_head: int = 14                  # head_dim must be 64
    n_layer: int = 18
    ffn_dim: int = 3584
    out_dim: int = 1024               # proj2 output (the LLM's embedding width)
    max_positions: int = 1500         # rows of the sinusoid table; beyond it the formula is evaluated (causal.py:204-228)
    chunk_frames: int = 8             # one decoder step = 80 ms = 8 mel frames (config.py:87-89)
    block_frames: int = 192           # fixed attention block (config.py:31-36); 0 = consume per chunk
    left_context_steps: int = 150     # 12 s of left context in steps (causal.py:103-106)
    block_bidirectional: bool = True
    conv_out_bias: bool = True
    mutable_tail_steps: int = 0       # bounded mutable tail (causal.py:101-113, 548-640): the last steps are re-encoded by
                                      # every call until they leave the tail; 0 = strict append-only; exclusive with blocks

    @property
    def freq_out(self) -> int:
        return self.n_mels // 8

    @property
    def conv_features(self) -> int:
        return self.conv_channels * self.freq_out

    def as_tuple(self):
        return (self.n_mels, self.conv_channels, self.d_model, self.n_head, self.n_layer, self.ffn_dim, self.out_dim,
                self.max_positions, self.chunk_frames, self.block_frames, self.left_context_steps,
                int(self.block_bidirectional), int(self.conv_out_bias), self.mutable_tail_steps)


QWEN_DIMS: Dict[str, QwenTowerDims] = {
    # test geometries (head_dim 64 like the real tower)
    "qnano": QwenTowerDims(conv_channels=8, d_model=128, n_head=2, n_layer=2, ffn_dim=256, out_dim=96, max_positions=40,
                           left_context_steps=30),
    "qnano-chunk": QwenTowerDims(conv_channels=8, d_model=128, n_head=2, n_layer=2, ffn_dim=256, out_dim=96,
                                 max_positions=4096, block_frames=0, left_context_steps=25, block_bidirectional=False),
    # bounded mutable tail: the last 6 steps (0.48 s) stay re-computable; causal, and bidirectional within a call
    "qnano-tail": QwenTowerDims(conv_channels=8, d_model=128, n_head=2, n_layer=2, ffn_dim=256, out_dim=96,
                                max_positions=4096, block_frames=0, left_context_steps=25, block_bidirectional=False,
                                mutable_tail_steps=6),
    "qnano-tail-bidir": QwenTowerDims(conv_channels=8, d_model=128, n_head=2, n_layer=2, ffn_dim=256, out_dim=96,
                                      max_positions=4096, block_frames=0, left_context_steps=25, block_bidirectional=True,
                                      mutable_tail_steps=6),
    # Qwen3-ASR-0.6B audio tower (public HF audio_config: d_model 896, 18 layers, 14 heads, ffn 3584,
    # downsample_hidden_size 480, output_dim 1024); to be confirmed against the checkpoint when it is mounted
    "qwen3-asr-0.6b": QwenTowerDims(),
}


def sinusoid_table(max_positions: int, d_model: int) -> np.ndarray:
    """The tower's fixed positional table: [sin | cos] of pos * exp(-ln(1e4)/(half-1) * i)."""
    half = d_model // 2
    inv = np.exp(-np.log(10000.0) / float(max(1, half - 1)) * np.arange(half, dtype=np.float32)).astype(np.float32)
    pos = np.arange(max_positions, dtype=np.float32)
    scaled = pos[:, None] * inv[None, :]
    return np.concatenate([np.sin(scaled), np.cos(scaled)], axis=1).astype(np.float32)


def synthetic_tower_state_dict(dims: QwenTowerDims, seed: int = 0) -> Dict[str, np.ndarray]:
    """Seeded weights with the tower's own parameter names, scaled for unit-variance activations."""
    rng = np.random.default_rng(seed)
    C, d, F = dims.conv_channels, dims.d_model, dims.ffn_dim

    def w(*shape, fan_in):
        return (rng.standard_normal(shape) / np.sqrt(fan_in)).astype(np.float32)

    def b(n, s=0.02):
        return (rng.standard_normal(n) * s).astype(np.float32)

    sd = {
        "conv2d1.weight": w(C, 1, 3, 3, fan_in=9) * 1.5, "conv2d1.bias": b(C),
        "conv2d2.weight": w(C, C, 3, 3, fan_in=9 * C) * 1.5, "conv2d2.bias": b(C),
        "conv2d3.weight": w(C, C, 3, 3, fan_in=9 * C) * 1.5, "conv2d3.bias": b(C),
        "conv_out.weight": w(d, dims.conv_features, fan_in=dims.conv_features) * 2.0,
        "positional_embedding.positional_embedding": sinusoid_table(dims.max_positions, d),
        "ln_post.weight": (1.0 + 0.1 * rng.standard_normal(d)).astype(np.float32), "ln_post.bias": b(d),
        "proj1.weight": w(d, d, fan_in=d), "proj1.bias": b(d),
        "proj2.weight": w(dims.out_dim, d, fan_in=d), "proj2.bias": b(dims.out_dim),
    }
    if dims.conv_out_bias:
        sd["conv_out.bias"] = b(d)
    for i in range(dims.n_layer):
        p = f"layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[p + f"self_attn.{n}.weight"] = w(d, d, fan_in=d) * (1.6 if n in ("q_proj", "k_proj") else 1.0)
            sd[p + f"self_attn.{n}.bias"] = b(d)
        for n in ("self_attn_layer_norm", "final_layer_norm"):
            sd[p + n + ".weight"] = (1.0 + 0.1 * rng.standard_normal(d)).astype(np.float32)
            sd[p + n + ".bias"] = b(d)
        sd[p + "fc1.weight"] = w(F, d, fan_in=d); sd[p + "fc1.bias"] = b(F)
        sd[p + "fc2.weight"] = w(d, F, fan_in=F); sd[p + "fc2.bias"] = b(d)
    return sd
