"""CPU restatement (torch fp32) of the streaming Sortformer step the reference drives -- TEST INFRASTRUCTURE ONLY.

**PARITY UNPINNED.**  The arithmetic of this row lives in NeMo 3.0.0
(``nemo.collections.asr``: ``SortformerEncLabelModel.forward_streaming_step`` / ``frontend_encoder`` / ``forward_infer``,
``SortformerModules.streaming_update_async`` / ``_compress_spkcache`` and helpers, ``ConformerEncoder`` with
``ConvSubsampling(dw_striding)`` / ``RelPositionMultiHeadAttention`` / ``ConformerConvolution``, the NLP
``TransformerEncoder`` (post-LN), ``FilterbankFeatures``), pinned in the reference's uv.lock:5029-5030 and absent from
/root/reference and from both containers, as is the checkpoint.  This file restates NeMo's published algorithm; it is
anchored on the reference's own call sites -- whisperlivekit/diarization/sortformer_backend.py:
  :120-126  streaming parameters            :175-196  AudioToMelSpectrogramPreprocessor(window 0.025, n_fft 512, 128 mels,
  :212-234  state layout (async: fixed-size            normalize "NA", pad_to 0), chunk duration 1.0 s
            spkcache / fifo + lengths)      :253-311  diarize(): 16 000 samples -> 101 mel frames (+99 of the previous
  :293-300  forward_streaming_step(processed_signal [1,T,128], length, state, total_preds, left_offset 8|0, right_offset 8)
-- and on the one numeric pin the reference's tests hold that runs without NeMo: ``_process_predictions``
(tests/test_sortformer_max_speakers.py, covered by oracle/diar_oracle.py).  Nothing here was checked against NeMo itself.
The state the reference allocates (fixed [1,188,512] caches with length counters) is NeMo's *async* streaming layout, so
the async update is the one restated.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from whisperlivekit_b200.sortformer_dims import SortformerDims, subsampled_len
from whisperlivekit_b200.weights import mel_filterbank

LOG_GUARD = 2.0 ** -24          # FilterbankFeatures(log_zero_guard_type="add", log_zero_guard_value=2**-24)
PREEMPH = 0.97


# ---------------------------------------------------------------------------------------------------------------
# front end: NeMo FilterbankFeatures in eval mode (no dither), normalize "NA", pad_to 0
# ---------------------------------------------------------------------------------------------------------------
def log_mel(audio: np.ndarray, d: SortformerDims) -> torch.Tensor:
    """audio fp32 [N] -> [n_mels, N // hop + 1]   (sortformer_backend.py:273 ``audio2mel.get_features``)"""
    x = torch.as_tensor(np.asarray(audio, np.float32))
    x = torch.cat([x[:1], x[1:] - PREEMPH * x[:-1]])                                   # pre-emphasis
    win = torch.hann_window(d.win_length, periodic=False, dtype=torch.float32)
    spec = torch.stft(x, d.n_fft, hop_length=d.hop, win_length=d.win_length, window=win, center=True,
                      pad_mode="reflect", return_complex=True)
    power = spec.real ** 2 + spec.imag ** 2                                            # mag_power 2.0
    fb = torch.from_numpy(mel_filterbank(d.n_mels, 16000, d.n_fft))
    return torch.log(fb @ power + LOG_GUARD)


def _topk_indices(x: torch.Tensor, k: int) -> torch.Tensor:
    """Indices of the k largest entries of every column of x [n, c] -> [k, c].  NeMo calls torch.topk(sorted=False), whose
    choice among EQUAL values is implementation-defined (and differs between torch's CPU and CUDA kernels); equal scores do
    occur -- a frame kept for two overlapping speakers sits in the cache twice with identical predictions -- and the order
    of the cache feeds the next step's relative-position attention.  Both this oracle and the CUDA kernel fix the choice:
    among equal values the lower index wins (a stable descending sort)."""
    order = torch.sort(x, dim=0, descending=True, stable=True).indices
    return order[:k]


# ---------------------------------------------------------------------------------------------------------------
# model
# ---------------------------------------------------------------------------------------------------------------
class SortformerOracle:
    def __init__(self, dims: SortformerDims, state_dict: Dict[str, np.ndarray]):
        self.d = dims
        self.w = {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)) for k, v in state_dict.items()}
        self._pe_cache: Dict[int, torch.Tensor] = {}

    # ---- ConvSubsampling(dw_striding, factor 8): x [T, n_mels] -> [T', d_model]
    def pre_encode(self, feats: torch.Tensor) -> torch.Tensor:
        w, p = self.w, "encoder.pre_encode."
        x = feats[None, None]                                                          # [1, 1, T, F]
        x = F.relu(F.conv2d(x, w[p + "conv.0.weight"], w[p + "conv.0.bias"], stride=2, padding=1))
        C = self.d.conv_channels
        for dw, pw in ((2, 3), (5, 6)):
            x = F.conv2d(x, w[p + f"conv.{dw}.weight"], w[p + f"conv.{dw}.bias"], stride=2, padding=1, groups=C)
            x = F.relu(F.conv2d(x, w[p + f"conv.{pw}.weight"], w[p + f"conv.{pw}.bias"]))
        _, c, t, f = x.shape
        x = x.transpose(1, 2).reshape(t, c * f)                                        # channel-major flatten (c, f)
        return F.linear(x, w[p + "out.weight"], w[p + "out.bias"])

    # ---- RelPositionalEncoding: rows for relative positions T-1 ... -(T-1)
    def rel_pos_emb(self, T: int) -> torch.Tensor:
        if T not in self._pe_cache:
            D = self.d.d_model
            pos = torch.arange(T - 1, -T, -1, dtype=torch.float32)[:, None]
            div = torch.exp(torch.arange(0, D, 2, dtype=torch.float32) * -(math.log(10000.0) / D))
            pe = torch.zeros(2 * T - 1, D)
            pe[:, 0::2] = torch.sin(pos * div)
            pe[:, 1::2] = torch.cos(pos * div)
            self._pe_cache[T] = pe
        return self._pe_cache[T]

    def _rel_attention(self, x: torch.Tensor, pe: torch.Tensor, q: str) -> torch.Tensor:
        """RelPositionMultiHeadAttention (Transformer-XL form, untied biases), no mask (one unpadded sequence)"""
        w, H, dk = self.w, self.d.n_head, self.d.d_head
        T = x.shape[0]
        qq = F.linear(x, w[q + "linear_q.weight"], w[q + "linear_q.bias"]).view(T, H, dk)
        kk = F.linear(x, w[q + "linear_k.weight"], w[q + "linear_k.bias"]).view(T, H, dk).transpose(0, 1)
        vv = F.linear(x, w[q + "linear_v.weight"], w[q + "linear_v.bias"]).view(T, H, dk).transpose(0, 1)
        pp = F.linear(pe, w[q + "linear_pos.weight"]).view(2 * T - 1, H, dk).transpose(0, 1)   # [H, 2T-1, dk]
        qu = (qq + w[q + "pos_bias_u"]).transpose(0, 1)                                # [H, T, dk]
        qv = (qq + w[q + "pos_bias_v"]).transpose(0, 1)
        ac = qu @ kk.transpose(1, 2)                                                   # [H, T, T]
        bd = qv @ pp.transpose(1, 2)                                                   # [H, T, 2T-1]
        # rel_shift: bd'[h, i, j] = bd[h, i, T - 1 - i + j]
        idx = (T - 1) - torch.arange(T)[:, None] + torch.arange(T)[None, :]
        bd = torch.gather(bd, 2, idx[None].expand(H, T, T))
        att = torch.softmax((ac + bd) / math.sqrt(dk), dim=-1)
        out = (att @ vv).transpose(0, 1).reshape(T, H * dk)
        return F.linear(out, w[q + "linear_out.weight"], w[q + "linear_out.bias"])

    def _conv_module(self, x: torch.Tensor, q: str) -> torch.Tensor:
        w, D, K = self.w, self.d.d_model, self.d.conv_kernel
        y = F.conv1d(x.t()[None], w[q + "pointwise_conv1.weight"], w[q + "pointwise_conv1.bias"])   # [1, 2D, T]
        y = F.glu(y, dim=1)
        y = F.conv1d(y, w[q + "depthwise_conv.weight"], w[q + "depthwise_conv.bias"], padding=(K - 1) // 2, groups=D)
        y = F.batch_norm(y, w[q + "batch_norm.running_mean"], w[q + "batch_norm.running_var"], w[q + "batch_norm.weight"],
                         w[q + "batch_norm.bias"], training=False, eps=1e-5)
        y = F.silu(y)
        y = F.conv1d(y, w[q + "pointwise_conv2.weight"], w[q + "pointwise_conv2.bias"])
        return y[0].t()

    def _ln(self, x, name):
        return F.layer_norm(x, (x.shape[-1],), self.w[name + ".weight"], self.w[name + ".bias"], 1e-5)

    def _ff(self, x, q):
        w = self.w
        return F.linear(F.silu(F.linear(x, w[q + "linear1.weight"], w[q + "linear1.bias"])), w[q + "linear2.weight"], w[q + "linear2.bias"])

    def conformer(self, emb: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
        """ConformerEncoder.forward(bypass_pre_encode=True) on one unpadded sequence [T, d_model]"""
        T = emb.shape[0]
        x = emb * math.sqrt(self.d.d_model)                                            # xscaling
        pe = self.rel_pos_emb(T)
        for i in range(self.d.n_layer):
            q = f"encoder.layers.{i}."
            x = x + 0.5 * self._ff(self._ln(x, q + "norm_feed_forward1"), q + "feed_forward1.")
            x = x + self._rel_attention(self._ln(x, q + "norm_self_att"), pe, q + "self_attn.")
            x = x + self._conv_module(self._ln(x, q + "norm_conv"), q + "conv.")
            x = x + 0.5 * self._ff(self._ln(x, q + "norm_feed_forward2"), q + "feed_forward2.")
            x = self._ln(x, q + "norm_out")
            if taps is not None:
                taps[f"conformer_{i}"] = x.clone()
        return x

    def transformer_and_head(self, x: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
        """encoder_proj -> 18 post-LN Transformer blocks -> forward_speaker_sigmoids; [T, d_model] -> [T, n_spk]"""
        w, H, dk = self.w, self.d.tf_n_head, self.d.tf_d_head
        x = F.linear(x, w["sortformer_modules.encoder_proj.weight"], w["sortformer_modules.encoder_proj.bias"])
        T = x.shape[0]
        scale = math.sqrt(math.sqrt(dk))
        for i in range(self.d.tf_n_layer):
            q = f"transformer_encoder.layers.{i}."
            a = q + "first_sub_layer."
            qq = (F.linear(x, w[a + "query_net.weight"], w[a + "query_net.bias"]) / scale).view(T, H, dk).transpose(0, 1)
            kk = (F.linear(x, w[a + "key_net.weight"], w[a + "key_net.bias"]) / scale).view(T, H, dk).transpose(0, 1)
            vv = F.linear(x, w[a + "value_net.weight"], w[a + "value_net.bias"]).view(T, H, dk).transpose(0, 1)
            att = torch.softmax(qq @ kk.transpose(1, 2), dim=-1)
            ctx = (att @ vv).transpose(0, 1).reshape(T, H * dk)
            x = self._ln(x + F.linear(ctx, w[a + "out_projection.weight"], w[a + "out_projection.bias"]), q + "layer_norm_1")
            f = q + "second_sub_layer."
            h = F.linear(F.relu(F.linear(x, w[f + "dense_in.weight"], w[f + "dense_in.bias"])), w[f + "dense_out.weight"], w[f + "dense_out.bias"])
            x = self._ln(x + h, q + "layer_norm_2")
            if taps is not None:
                taps[f"transformer_{i}"] = x.clone()
        h = F.relu(x)
        h = F.linear(h, w["sortformer_modules.first_hidden_to_hidden.weight"], w["sortformer_modules.first_hidden_to_hidden.bias"])
        h = F.relu(h)
        h = F.linear(h, w["sortformer_modules.single_hidden_to_spks.weight"], w["sortformer_modules.single_hidden_to_spks.bias"])
        return torch.sigmoid(h)

    # -----------------------------------------------------------------------------------------------------------
    # SortformerModules: speaker-cache compression
    # -----------------------------------------------------------------------------------------------------------
    def _log_pred_scores(self, preds):
        th = self.d.pred_score_threshold
        lp = torch.log(torch.clamp(preds, min=th))
        l1 = torch.log(torch.clamp(1.0 - preds, min=th))
        return lp - l1 + l1.sum(dim=1, keepdim=True) - math.log(0.5)

    @staticmethod
    def _disable_low_scores(preds, scores, min_pos):
        neg_inf = torch.tensor(float("-inf"))
        is_speech = preds > 0.5
        scores = torch.where(is_speech, scores, neg_inf)
        is_pos = scores > 0
        repl = (~is_pos) & is_speech & (is_pos.sum(dim=0, keepdim=True) >= min_pos)
        return torch.where(repl, neg_inf, scores)

    @staticmethod
    def _boost_topk(scores, n_boost, scale):
        n = scores.shape[0]
        k = min(n_boost, n)
        if k <= 0:
            return scores
        idx = _topk_indices(scores, k)
        scores = scores.clone()
        cols = torch.arange(scores.shape[1])[None, :].expand_as(idx)
        scores[idx, cols] -= scale * math.log(0.5)
        return scores

    def compress_spkcache(self, emb, preds, mean_sil_emb):
        """SortformerModules._compress_spkcache(permute_spk=False) for one stream.
        emb [n, D], preds [n, n_spk] -> ([spkcache_len, D], [spkcache_len, n_spk])"""
        d = self.d
        n, S = preds.shape
        per_spk = d.spkcache_len // S - d.spkcache_sil_frames_per_spk
        strong = math.floor(per_spk * d.strong_boost_rate)
        weak = math.floor(per_spk * d.weak_boost_rate)
        min_pos = math.floor(per_spk * d.min_pos_scores_rate)
        scores = self._log_pred_scores(preds)
        scores = self._disable_low_scores(preds, scores, min_pos)
        if d.scores_boost_latest > 0:
            scores = scores.clone()
            scores[d.spkcache_len:, :] += d.scores_boost_latest
        scores = self._boost_topk(scores, strong, 2.0)
        scores = self._boost_topk(scores, weak, 1.0)
        if d.spkcache_sil_frames_per_spk > 0:
            scores = torch.cat([scores, torch.full((d.spkcache_sil_frames_per_spk, S), float("inf"))], dim=0)
        n_frames = scores.shape[0]
        n_no_sil = n_frames - d.spkcache_sil_frames_per_spk
        flat = scores.t().reshape(-1)                                                  # speaker-major
        idx = _topk_indices(flat[:, None], d.spkcache_len)[:, 0]
        vals = flat[idx]
        idx = torch.where(vals != float("-inf"), idx, torch.tensor(d.max_index))
        idx, _ = torch.sort(idx)
        disabled = idx == d.max_index
        idx = torch.remainder(idx, n_frames)
        disabled = disabled | (idx >= n_no_sil)
        idx = torch.where(disabled, torch.zeros_like(idx), idx)
        e = torch.where(disabled[:, None], mean_sil_emb[None, :], emb[idx])
        p = torch.where(disabled[:, None], torch.zeros(()), preds[idx])
        return e, p

    def silence_profile(self, mean_sil_emb, n_sil, emb, preds):
        is_sil = preds.sum(dim=-1) < self.d.sil_threshold
        cnt = int(is_sil.sum())
        if cnt == 0:
            return mean_sil_emb, n_sil
        total = mean_sil_emb * n_sil + emb[is_sil].sum(dim=0)
        n_new = n_sil + cnt
        return total / max(n_new, 1), n_new

    # -----------------------------------------------------------------------------------------------------------
    # streaming state + step
    # -----------------------------------------------------------------------------------------------------------
    def init_state(self) -> dict:
        d = self.d
        return dict(spkcache=torch.zeros(d.spkcache_len, d.d_model), spkcache_preds=torch.zeros(d.spkcache_len, d.n_spk),
                    spkcache_len=0, fifo=torch.zeros(d.fifo_len, d.d_model),
                    fifo_preds=torch.zeros(d.fifo_len, d.n_spk), fifo_len=0, mean_sil_emb=torch.zeros(d.d_model), n_sil=0)

    def streaming_update(self, st: dict, chunk: torch.Tensor, preds: torch.Tensor, lc: int, rc: int) -> torch.Tensor:
        """SortformerModules.streaming_update_async for one stream.  chunk [Tc, D] pre-encode rows of this step, preds
        [spkcache_len_valid + fifo_len_valid + Tc, n_spk].  Mutates ``st``; returns chunk_preds [max_chunk_len, n_spk] (rows
        past the valid chunk length stay zero, as in NeMo)."""
        d = self.d
        max_chunk = chunk.shape[0] - lc - rc
        clen = min(max(chunk.shape[0] - lc, 0), max_chunk)
        sl, fl = st["spkcache_len"], st["fifo_len"]
        st["fifo_preds"] = torch.zeros(d.fifo_len, d.n_spk)
        st["fifo_preds"][:fl] = preds[sl:sl + fl]
        chunk_preds = torch.zeros(max_chunk, d.n_spk)
        chunk_preds[:clen] = preds[sl + fl + lc: sl + fl + lc + clen]
        up_fifo = torch.zeros(d.fifo_len + max_chunk, d.d_model)
        up_fifo_preds = torch.zeros(d.fifo_len + max_chunk, d.n_spk)
        up_fifo[:fl] = st["fifo"][:fl]
        up_fifo_preds[:fl] = st["fifo_preds"][:fl]
        up_fifo[fl:fl + clen] = chunk[lc:lc + clen]
        up_fifo_preds[fl:fl + clen] = chunk_preds[:clen]
        new_fl = fl + clen
        max_pop = min(max(d.spkcache_update_period, max_chunk), max_chunk + d.fifo_len)
        up_cache = torch.zeros(d.spkcache_len + max_pop, d.d_model)
        up_cache_preds = torch.zeros(d.spkcache_len + max_pop, d.n_spk)
        up_cache[:sl] = st["spkcache"][:sl]
        up_cache_preds[:sl] = st["spkcache_preds"][:sl]
        if new_fl > d.fifo_len:
            pop = d.spkcache_update_period
            pop = max(pop, max_chunk - d.fifo_len + fl)
            pop = min(pop, new_fl)
            pop_emb, pop_preds = up_fifo[:pop], up_fifo_preds[:pop]
            st["mean_sil_emb"], st["n_sil"] = self.silence_profile(st["mean_sil_emb"], st["n_sil"], pop_emb, pop_preds)
            up_cache[sl:sl + pop] = pop_emb
            # the reference allocates spkcache_preds up front (sortformer_backend.py:219-222), so NeMo's "cache predictions
            # already exist" branch is the one taken from the first pop-out on: the popped rows keep this step's predictions
            up_cache_preds[sl:sl + pop] = pop_preds
            sl += pop
            new_fl -= pop
            up_fifo[:new_fl] = up_fifo[pop:pop + new_fl].clone()
            up_fifo_preds[:new_fl] = up_fifo_preds[pop:pop + new_fl].clone()
            up_fifo[new_fl:] = 0
            up_fifo_preds[new_fl:] = 0
        st["fifo"], st["fifo_preds"], st["fifo_len"] = up_fifo[:d.fifo_len].clone(), up_fifo_preds[:d.fifo_len].clone(), new_fl
        if sl > d.spkcache_len:
            e, p = self.compress_spkcache(up_cache, up_cache_preds, st["mean_sil_emb"])
            st["spkcache"], st["spkcache_preds"], sl = e, p, d.spkcache_len
        else:
            st["spkcache"], st["spkcache_preds"] = up_cache[:d.spkcache_len].clone(), up_cache_preds[:d.spkcache_len].clone()
        st["spkcache_len"] = sl
        return chunk_preds

    def forward_streaming_step(self, feats: torch.Tensor, st: dict, left_offset: int, right_offset: int, taps: Optional[dict] = None):
        """feats [T, n_mels] (time-major, as the reference passes ``chunk_feat_seq_t``) -> chunk_preds [max_chunk_len, n_spk]"""
        d = self.d
        chunk = self.pre_encode(feats)
        seq = torch.cat([st["spkcache"][:st["spkcache_len"]], st["fifo"][:st["fifo_len"]], chunk], dim=0)
        if taps is not None:
            taps["pre_encode"] = chunk.clone()
            taps["sequence"] = seq.clone()
        enc = self.conformer(seq, taps)
        preds = self.transformer_and_head(enc, taps)
        if taps is not None:
            taps["preds"] = preds.clone()
        lc = round(left_offset / d.encoder_subsampling)
        rc = math.ceil(right_offset / d.encoder_subsampling)
        return self.streaming_update(st, chunk, preds, lc, rc)


class OracleDiarizer:
    """The per-stream loop of ``SortformerDiarizationOnline.diarize`` (sortformer_backend.py:253-311) over the oracle."""

    def __init__(self, model: SortformerOracle):
        self.m = model
        self.st = model.init_state()
        self.prev: Optional[torch.Tensor] = None
        self.chunk_index = 0
        self.total_preds = torch.zeros(0, model.d.n_spk)

    def step(self, audio_1s: np.ndarray, taps: Optional[dict] = None) -> torch.Tensor:
        mel = log_mel(audio_1s, self.m.d)                                              # [n_mels, 101]
        total = mel if self.prev is None else torch.cat([self.prev[:, -99:], mel], dim=1)
        self.prev = mel
        if taps is not None:
            taps["mel"] = mel.clone()
        cp = self.m.forward_streaming_step(total.t().contiguous(), self.st, 8 if self.chunk_index > 0 else 0, 8, taps)
        self.total_preds = torch.cat([self.total_preds, cp], dim=0)
        self.chunk_index += 1
        return cp
