"""CPU ORACLE (test infrastructure only) for the Qwen3-ASR causal-KV audio tower.

Restates reference third_party/qwen3-asr-causal/src/qwen3_asr_causal/causal.py in plain torch CPU ops:
  forward_chunk        :713-782   buffering to whole blocks, one _encode_ready_mels per block
  flush_pending        :687-711   end of stream: the buffered whole 8-frame chunks, as one piece
  _encode_ready_mels   :642-681   conv blocks -> L layers with per-layer K/V cache -> ln_post/proj1/GELU/proj2
  _conv_one_block      :230-248   three 3x3 stride-2 convs + GELU on one 8-frame chunk, linear, + sinusoid(position)
  _position_embedding  :204-228   table rows, or the closed form beyond the table
  _attention_chunk     :292-376   cached K/V (bounded left window), block-bidirectional / causal mask, fp32 softmax
  _layer_chunk         :378-421   pre-LN attention and MLP with residuals
Only the append-only regime is restated (mutable tail off, as the fixed-block production config requires,
config.py:99-104).  Pinned on the reference itself: oracle/make_golden_qwen.py runs the reference's
QwenAudioCausalKVEncoder over a seeded tower of the same geometry -> tests/golden/qwen_*.npz.
Nothing outside tests/, __graft_entry__.smoke() and bench.py's CPU arms may import this module.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.nn.functional as F


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


class QwenTowerOracle:
    """Session API mirrored by whisperlivekit_b200.qwen_engine.QwenTowerEngine."""

    backend = "oracle-cpu"

    def __init__(self, dims, state_dict: Dict[str, np.ndarray]):
        self.dims = dims
        self.W = {k: _t(v) for k, v in state_dict.items()}
        self._s: Dict[int, dict] = {}
        self._next = 0

    # -- sessions -------------------------------------------------------------------
    def open_session(self) -> int:
        sid = self._next
        self._next += 1
        self._s[sid] = dict(buf=torch.zeros(0, self.dims.n_mels), caches=[None] * self.dims.n_layer, emitted=0, tail=[], mutable=0)
        return sid

    def close_session(self, sid: int) -> None:
        self._s.pop(sid)

    def reset_session(self, sid: int) -> None:
        self._s[sid] = dict(buf=torch.zeros(0, self.dims.n_mels), caches=[None] * self.dims.n_layer, emitted=0, tail=[], mutable=0)

    # incremental log-mel front end behind the same method names as QwenTowerEngine (features.py:32-112)
    def load_mel_filters(self, filters=None) -> None:
        from whisperlivekit_b200.weights import mel_filterbank
        self._filters = np.asarray(mel_filterbank(self.dims.n_mels) if filters is None else filters, np.float32)

    def _mel(self, sid: int):
        from oracle.qwen_mel_oracle import StreamingMelOracle
        s = self._s[sid]
        if "mel" not in s:
            s["mel"] = StreamingMelOracle(self._filters)
        return s["mel"]

    def mel_append(self, sids, audios):
        out = [self._mel(sid).append(a) for sid, a in zip(sids, audios)]
        return [np.zeros((0, self.dims.n_mels), np.float32) if m is None else m for m in out]

    def mel_flush(self, sids):
        out = [self._mel(sid).flush() for sid in sids]
        return [np.zeros((0, self.dims.n_mels), np.float32) if m is None else m for m in out]

    def pending_frames(self, sid: int) -> int:
        return int(self._s[sid]["buf"].shape[0])

    def mutable_steps(self, sid: int) -> int:
        return int(self._s[sid].get("mutable", 0))

    def emitted_steps(self, sid: int) -> int:
        return int(self._s[sid]["emitted"])

    # -- pieces -----------------------------------------------------------------------
    def _pos(self, offset: int, length: int) -> torch.Tensor:                       # causal.py:204-228
        table = self.W["positional_embedding.positional_embedding"]
        if offset + length <= table.shape[0]:
            return table[offset: offset + length]
        half = table.shape[1] // 2
        inv = torch.exp(-math.log(10000.0) / float(max(1, half - 1)) * torch.arange(half, dtype=torch.float32))
        pos = torch.arange(offset, offset + length, dtype=torch.float32)
        scaled = pos[:, None] * inv[None, :]
        return torch.cat([scaled.sin(), scaled.cos()], dim=1)

    def _conv_chunk(self, chunk: torch.Tensor, position: int) -> torch.Tensor:      # causal.py:230-248
        W = self.W
        x = chunk.transpose(0, 1)[None, None]                                       # [1, 1, n_mels, frames]
        x = F.gelu(F.conv2d(x, W["conv2d1.weight"], W["conv2d1.bias"], stride=2, padding=1))
        x = F.gelu(F.conv2d(x, W["conv2d2.weight"], W["conv2d2.bias"], stride=2, padding=1))
        x = F.gelu(F.conv2d(x, W["conv2d3.weight"], W["conv2d3.bias"], stride=2, padding=1))
        b, c, f, t = x.shape
        x = F.linear(x.permute(0, 3, 1, 2).contiguous().view(b, t, c * f), W["conv_out.weight"], W.get("conv_out.bias"))
        return x[0] + self._pos(position, t)

    def _layer(self, i: int, h: torch.Tensor, cache, position: int, full_kv: bool = False):   # causal.py:292-421 (423-546 with full_kv)
        W, D = self.W, self.dims
        p = f"layers.{i}."
        n = h.shape[0]
        H, hd = D.n_head, D.d_model // D.n_head
        x = F.layer_norm(h, (D.d_model,), W[p + "self_attn_layer_norm.weight"], W[p + "self_attn_layer_norm.bias"], 1e-5)
        q = F.linear(x, W[p + "self_attn.q_proj.weight"], W[p + "self_attn.q_proj.bias"]).view(n, H, hd).transpose(0, 1)
        k = F.linear(x, W[p + "self_attn.k_proj.weight"], W[p + "self_attn.k_proj.bias"]).view(n, H, hd).transpose(0, 1)
        v = F.linear(x, W[p + "self_attn.v_proj.weight"], W[p + "self_attn.v_proj.bias"]).view(n, H, hd).transpose(0, 1)
        past = 0 if cache is None else cache[0].shape[1]
        if cache is not None:
            k = torch.cat([cache[0], k], dim=1)
            v = torch.cat([cache[1], v], dim=1)
        total = k.shape[1]
        q_pos = torch.arange(position, position + n)
        k_pos = torch.arange(position - past, position - past + total)
        if D.block_bidirectional:
            allowed = (k_pos[None, :] <= position + n - 1).expand(n, total)
        else:
            allowed = k_pos[None, :] <= q_pos[:, None]
        allowed = allowed & (k_pos[None, :] >= (q_pos[:, None] - D.left_context_steps + 1))
        scores = torch.matmul(q, k.transpose(-2, -1)) * float(hd ** -0.5)
        scores = scores.masked_fill(~allowed[None], torch.finfo(torch.float32).min)
        ctx = torch.matmul(F.softmax(scores, dim=-1), v).transpose(0, 1).contiguous().view(n, D.d_model)
        h = h + F.linear(ctx, W[p + "self_attn.out_proj.weight"], W[p + "self_attn.out_proj.bias"])
        keep = total if full_kv else min(total, D.left_context_steps)      # _attention_tail hands all keys of the call back
        new_cache = (k[:, -keep:].clone(), v[:, -keep:].clone())
        y = F.layer_norm(h, (D.d_model,), W[p + "final_layer_norm.weight"], W[p + "final_layer_norm.bias"], 1e-5)
        y = F.gelu(F.linear(y, W[p + "fc1.weight"], W[p + "fc1.bias"]))
        h = h + F.linear(y, W[p + "fc2.weight"], W[p + "fc2.bias"])
        return h, new_cache

    def _encode_ready(self, s: dict, mels: torch.Tensor) -> torch.Tensor:          # causal.py:642-681
        W, D = self.W, self.dims
        position = s["emitted"]
        hs, step = [], position
        for a in range(0, mels.shape[0], D.chunk_frames):
            c = self._conv_chunk(mels[a: a + D.chunk_frames], step)
            hs.append(c)
            step += c.shape[0]
        h = torch.cat(hs, dim=0)
        for i in range(D.n_layer):
            h, s["caches"][i] = self._layer(i, h, s["caches"][i], position)
        h = F.layer_norm(h, (D.d_model,), W["ln_post.weight"], W["ln_post.bias"], 1e-5)
        h = F.gelu(F.linear(h, W["proj1.weight"], W["proj1.bias"]))
        h = F.linear(h, W["proj2.weight"], W["proj2.bias"])
        s["emitted"] += h.shape[0]
        return h

    def _encode_mutable_tail(self, s: dict, ready: torch.Tensor) -> torch.Tensor:  # causal.py:548-640 (+ :423-546)
        """Recompute the mutable tail plus the new chunks over the frozen K/V prefix WITHOUT touching the caches, return
        the hidden rows of all of them, then freeze leading chunks until the tail fits ``mutable_tail_steps``."""
        W, D = self.W, self.dims
        tail = s.setdefault("tail", [])
        blocks = list(tail) + [ready[a: a + D.chunk_frames] for a in range(0, ready.shape[0], D.chunk_frames)]
        if not blocks:
            return torch.zeros(0, D.out_dim)
        position = s["emitted"]
        hs, step, steps = [], position, []
        for b in blocks:
            c = self._conv_chunk(b, step)
            hs.append(c); steps.append(c.shape[0]); step += c.shape[0]
        h = torch.cat(hs, dim=0)
        all_kv = []
        for i in range(D.n_layer):
            h, kv = self._layer(i, h, s["caches"][i], position, full_kv=True)   # [frozen cache + every key of the call]; cache untouched
            all_kv.append(kv)
        h = F.layer_norm(h, (D.d_model,), W["ln_post.weight"], W["ln_post.bias"], 1e-5)
        h = F.gelu(F.linear(h, W["proj1.weight"], W["proj1.bias"]))
        h = F.linear(h, W["proj2.weight"], W["proj2.bias"])
        total_steps = sum(steps)
        freeze_steps = freeze_blocks = 0
        while freeze_blocks < len(blocks) and total_steps - freeze_steps - steps[freeze_blocks] >= D.mutable_tail_steps:
            freeze_steps += steps[freeze_blocks]; freeze_blocks += 1
        if freeze_steps > 0:
            drop = total_steps - freeze_steps                      # the newest keys stay mutable: not in the cache (:628-637)
            for i in range(D.n_layer):
                k_all, v_all = all_kv[i]
                end = k_all.shape[1] - drop
                keep = min(end, D.left_context_steps)
                s["caches"][i] = (k_all[:, end - keep: end].clone(), v_all[:, end - keep: end].clone())
            s["emitted"] += freeze_steps
        s["tail"] = blocks[freeze_blocks:]
        s["mutable"] = total_steps - freeze_steps
        return h

    # -- the entry point ---------------------------------------------------------------
    @torch.no_grad()
    def forward_chunk(self, sids: Sequence[int], mels: Sequence[np.ndarray]) -> List[np.ndarray]:
        """causal.py:713-782 per session: append mel frames [n_i, n_mels], encode every complete block / chunk,
        return the newly emitted rows [steps_i, out_dim]."""
        D = self.dims
        out = []
        for sid, m in zip(sids, mels):
            s = self._s[sid]
            m = _t(np.asarray(m, np.float32).reshape(-1, D.n_mels))
            if m.shape[0] == 0:                                    # causal.py:731-736: an empty append touches nothing
                out.append(np.zeros((0, D.out_dim), np.float32))
                continue
            buf = torch.cat([s["buf"], m], dim=0)
            consume = D.block_frames if D.block_frames > 0 else D.chunk_frames
            ready = (buf.shape[0] // consume) * consume
            s["buf"] = buf[ready:].clone()
            if getattr(D, "mutable_tail_steps", 0) > 0:
                out.append(self._encode_mutable_tail(s, buf[:ready]).numpy())
                continue
            if ready == 0:
                out.append(np.zeros((0, D.out_dim), np.float32))
                continue
            if D.block_frames > 0:
                rows = [self._encode_ready(s, buf[a: a + D.block_frames]) for a in range(0, ready, D.block_frames)]
                out.append(torch.cat(rows, dim=0).numpy())
            else:
                out.append(self._encode_ready(s, buf[:ready]).numpy())
        return out

    @torch.no_grad()
    def flush_pending(self, sids: Sequence[int]) -> List[np.ndarray]:
        """causal.py:687-711: encode the buffered whole conv chunks (< one block) as one piece, drop the rest."""
        D = self.dims
        out = []
        for sid in sids:
            s = self._s[sid]
            ready = (s["buf"].shape[0] // D.chunk_frames) * D.chunk_frames
            buf, s["buf"] = s["buf"], torch.zeros(0, D.n_mels)
            out.append(self._encode_ready(s, buf[:ready]).numpy() if ready else np.zeros((0, D.out_dim), np.float32))
        return out
