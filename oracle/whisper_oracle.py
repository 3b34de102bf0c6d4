"""CPU ORACLE for the streaming-Whisper hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain fp32 CPU restatement (torch CPU tensor ops, no autograd,
no nn.Module) of the reference's algorithm for the path BASELINE.json names:
log-mel -> Whisper encoder -> incremental decoder with cross-attention export
-> AlignAtt attention post-processing -> greedy token update.  Every function
cites the reference file:line it follows.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it; the product (``whisperlivekit_b200``) never
does, and fails loudly without its CUDA extension.

Pinning: oracle/make_golden.py runs the *real* reference
(/root/reference/whisperlivekit) in the build container on seeded weights and
audio and commits small fixtures under tests/golden/;
tests/test_oracle_golden.py checks this restatement against them (and, when
/root/reference is present, against the live reference).  Parity is therefore
pinned on outputs of the reference itself, not on published vectors -- the
reference's own tests hold no golden logits/tokens.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

N_FFT = 400
HOP = 160
N_SAMPLES = 480000
N_FRAMES = 3000


def _t(x) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x
    return torch.from_numpy(np.ascontiguousarray(x))


class Weights:
    """Float32 torch views over a name->ndarray state dict (reference key names)."""

    def __init__(self, sd: Dict[str, np.ndarray]):
        self.t = {k: _t(v).float() for k, v in sd.items()}

    def __getitem__(self, k: str) -> torch.Tensor:
        return self.t[k]

    def get(self, k: str) -> Optional[torch.Tensor]:
        return self.t.get(k)


# ----------------------------------------------------------------------------
# a1. log-mel front end -- reference whisper/audio.py:110-157
# ----------------------------------------------------------------------------
def log_mel_spectrogram(audio: torch.Tensor, filters: torch.Tensor, padding: int = 0) -> torch.Tensor:
    audio = _t(audio).float()
    if padding > 0:
        audio = F.pad(audio, (0, padding))                                   # audio.py:145-146
    window = torch.hann_window(N_FFT)                                        # audio.py:147
    stft = torch.stft(audio, N_FFT, HOP, window=window, return_complex=True) # audio.py:148
    magnitudes = stft[..., :-1].abs() ** 2                                   # audio.py:149
    mel_spec = _t(filters).float() @ magnitudes                              # audio.py:152
    log_spec = torch.clamp(mel_spec, min=1e-10).log10()                      # audio.py:154
    log_spec = torch.maximum(log_spec, log_spec.max() - 8.0)                 # audio.py:155
    return (log_spec + 4.0) / 4.0                                            # audio.py:156


def pad_or_trim(x: torch.Tensor, length: int = N_FRAMES) -> torch.Tensor:
    """reference whisper/audio.py:65-88 (last axis)."""
    if x.shape[-1] > length:
        x = x[..., :length]
    if x.shape[-1] < length:
        x = F.pad(x, (0, length - x.shape[-1]))
    return x


def encode_features(audio, filters) -> Tuple[torch.Tensor, int]:
    """reference simul_whisper/simul_whisper.py:345-350 -> (mel[1,n_mels,3000], content_mel_len)."""
    mel_padded = log_mel_spectrogram(audio, filters, padding=N_SAMPLES).unsqueeze(0)
    mel = pad_or_trim(mel_padded, N_FRAMES)
    content_mel_len = int((mel_padded.shape[2] - mel.shape[2]) / 2)
    return mel, content_mel_len


# ----------------------------------------------------------------------------
# a2/a3. encoder -- reference whisper/model.py:39-59, 81-173, 176-254
# ----------------------------------------------------------------------------
def layer_norm(x, w, b):
    return F.layer_norm(x.float(), (x.shape[-1],), w, b, 1e-5)               # model.py:39-41


def linear(x, w, b=None):
    return F.linear(x, w, b)                                                 # model.py:44-50


def qkv_attention(q, k, v, n_head: int, mask=None):
    """reference whisper/model.py:148-173 (use_sdpa=False branch); returns (out, qk)."""
    n_batch, n_ctx, n_state = q.shape
    scale = (n_state // n_head) ** -0.25
    q = q.view(*q.shape[:2], n_head, -1).permute(0, 2, 1, 3)
    k = k.view(*k.shape[:2], n_head, -1).permute(0, 2, 1, 3)
    v = v.view(*v.shape[:2], n_head, -1).permute(0, 2, 1, 3)
    qk = (q * scale) @ (k * scale).transpose(-1, -2)
    if mask is not None:
        qk = qk + mask[:n_ctx, :n_ctx]
    qk = qk.float()
    w = F.softmax(qk, dim=-1)
    out = (w @ v).permute(0, 2, 1, 3).flatten(start_dim=2)
    return out, qk


def _mha_self(W: Weights, p: str, x, n_head, mask=None, kv_cache: Optional[dict] = None,
              n_text_ctx: int = 448):
    q = linear(x, W[p + ".query.weight"], W[p + ".query.bias"])
    k = linear(x, W[p + ".key.weight"])
    v = linear(x, W[p + ".value.weight"], W[p + ".value.bias"])
    if kv_cache is not None:                                                 # model.py:130-146
        kk, vk = p + "_key", p + "_value"
        if kk not in kv_cache or k.shape[1] > n_text_ctx:
            kv_cache[kk], kv_cache[vk] = k, v
        else:
            k = torch.cat([kv_cache[kk], k], dim=1)
            v = torch.cat([kv_cache[vk], v], dim=1)
            kv_cache[kk], kv_cache[vk] = k, v
    out, qk = qkv_attention(q, k, v, n_head, mask)
    return linear(out, W[p + ".out.weight"], W[p + ".out.bias"]), qk


def _mha_cross(W: Weights, p: str, x, xa, n_head, kv_cache: Optional[dict]):
    q = linear(x, W[p + ".query.weight"], W[p + ".query.bias"])
    kk, vk = p + "_key", p + "_value"
    if kv_cache is not None and kk in kv_cache:                              # model.py:116-125
        k, v = kv_cache[kk], kv_cache[vk]
    else:
        k = linear(xa, W[p + ".key.weight"])
        v = linear(xa, W[p + ".value.weight"], W[p + ".value.bias"])
        if kv_cache is not None:
            kv_cache[kk], kv_cache[vk] = k, v
    out, qk = qkv_attention(q, k, v, n_head, None)
    return linear(out, W[p + ".out.weight"], W[p + ".out.bias"]), qk


def _mlp(W: Weights, p: str, x):
    h = F.gelu(linear(x, W[p + ".mlp.0.weight"], W[p + ".mlp.0.bias"]))     # exact erf GELU
    return linear(h, W[p + ".mlp.2.weight"], W[p + ".mlp.2.bias"])


def encoder_stem(W: Weights, mel: torch.Tensor) -> torch.Tensor:
    """conv1+GELU, conv2(stride 2)+GELU, + positional -- model.py:243-248. -> [B,1500,d]"""
    x = F.gelu(F.conv1d(mel, W["encoder.conv1.weight"], W["encoder.conv1.bias"], padding=1))
    x = F.gelu(F.conv1d(x, W["encoder.conv2.weight"], W["encoder.conv2.bias"], stride=2, padding=1))
    x = x.permute(0, 2, 1)
    return x + W["encoder.positional_embedding"]


def encoder_forward(W: Weights, dims, mel: torch.Tensor, return_layers: bool = False):
    """reference whisper/model.py:238-254."""
    x = encoder_stem(W, mel)
    layers = [x]
    for i in range(dims.n_audio_layer):
        p = f"encoder.blocks.{i}"
        a, _ = _mha_self(W, p + ".attn", layer_norm(x, W[p + ".attn_ln.weight"], W[p + ".attn_ln.bias"]),
                         dims.n_audio_head)
        x = x + a
        x = x + _mlp(W, p, layer_norm(x, W[p + ".mlp_ln.weight"], W[p + ".mlp_ln.bias"]))
        if return_layers:
            layers.append(x)
    x = layer_norm(x, W["encoder.ln_post.weight"], W["encoder.ln_post.bias"])
    return (x, layers) if return_layers else x


# ----------------------------------------------------------------------------
# a4. decoder with dict KV cache -- reference whisper/model.py:281-332
# ----------------------------------------------------------------------------
def decoder_forward(W: Weights, dims, tokens: torch.Tensor, xa: torch.Tensor,
                    kv_cache: Optional[dict]):
    """-> (logits[B,Tq,V] fp32, [L x qk[B,H,Tq,1500]])"""
    offset = 0
    first_key = "decoder.blocks.0.attn_key"
    if kv_cache and first_key in kv_cache:
        offset = kv_cache[first_key].shape[1]                                # model.py:306-311
    x = F.embedding(tokens, W["decoder.token_embedding.weight"]) \
        + W["decoder.positional_embedding"][offset: offset + tokens.shape[-1]]
    n_ctx = dims.n_text_ctx
    mask = torch.empty(n_ctx, n_ctx).fill_(-np.inf).triu_(1)                 # model.py:278
    cross = []
    for i in range(dims.n_text_layer):
        p = f"decoder.blocks.{i}"
        a, _ = _mha_self(W, p + ".attn", layer_norm(x, W[p + ".attn_ln.weight"], W[p + ".attn_ln.bias"]),
                         dims.n_text_head, mask=mask, kv_cache=kv_cache, n_text_ctx=n_ctx)
        x = x + a
        c, qk = _mha_cross(W, p + ".cross_attn",
                           layer_norm(x, W[p + ".cross_attn_ln.weight"], W[p + ".cross_attn_ln.bias"]),
                           xa, dims.n_text_head, kv_cache)
        x = x + c
        cross.append(qk)
        x = x + _mlp(W, p, layer_norm(x, W[p + ".mlp_ln.weight"], W[p + ".mlp_ln.bias"]))
    x = layer_norm(x, W["decoder.ln.weight"], W["decoder.ln.bias"])
    logits = (x @ W["decoder.token_embedding.weight"].t()).float()          # model.py:325-328
    return logits, cross


# ----------------------------------------------------------------------------
# a7. AlignAtt attention post-processing -- reference simul_whisper.py:390-433,
#     median filter reference whisper/timing.py:19-54
# ----------------------------------------------------------------------------
def median_filter(x: torch.Tensor, filter_width: int) -> torch.Tensor:
    pad = filter_width // 2
    if x.shape[-1] <= pad:
        return x
    x = F.pad(x, (pad, pad, 0, 0), mode="reflect")
    return x.unfold(-1, filter_width, 1).sort()[0][..., pad]


def process_cross_attention(accumulated: List[List[torch.Tensor]], align_heads: Sequence[Tuple[int, int]],
                            n_layer: int, content_mel_len: int) -> torch.Tensor:
    """accumulated: list over retained iterations of list over layers of qk[1,H,Tq,1500].
    -> attn[1, sum Tq, content_mel_len]"""
    by_layer: Dict[int, List[Tuple[int, int]]] = {}
    for rank, (l, h) in enumerate(align_heads):                              # simul_whisper.py:151-159
        by_layer.setdefault(l, []).append((rank, h))
    per_head: List[List[torch.Tensor]] = [[] for _ in align_heads]
    flat = [a for it in accumulated for a in it]
    for idx, attn_mat in enumerate(flat):
        l = idx % n_layer
        if l not in by_layer:
            continue
        sm = F.softmax(attn_mat, dim=-1)                                     # :406
        for rank, h in by_layer[l]:
            per_head[rank].append(sm[0, h, :, :].unsqueeze(0))
    tmp = [torch.cat(m, dim=1) for m in per_head if m]
    if not tmp:
        return torch.zeros(1, 1, content_mel_len)
    a = torch.stack(tmp, dim=1)                                              # [1, n_align, T, 1500]
    std, mean = torch.std_mean(a, dim=-2, keepdim=True, unbiased=False)      # :426-428
    a = (a - mean) / (std + 1e-8)
    a = median_filter(a, 7)
    a = a.mean(dim=1)
    return a[:, :, :content_mel_len]


# ----------------------------------------------------------------------------
# a9. logit post-processing + greedy update
# ----------------------------------------------------------------------------
def no_speech_prob(logits_row: torch.Tensor, no_speech: int) -> float:
    """reference simul_whisper.py:370-377 (softmax over the vocabulary at sot_index)."""
    return float(logits_row.float().softmax(dim=-1)[no_speech])


def greedy_update(logits_row: torch.Tensor) -> Tuple[int, float]:
    """reference whisper/decoding.py:271-287 at temperature 0 -> (token, logprob)."""
    nxt = int(logits_row.argmax(dim=-1))
    lp = F.log_softmax(logits_row.float(), dim=-1)[nxt]
    return nxt, float(lp)


# ----------------------------------------------------------------------------
# Engine-shaped wrapper so the package's host policy code can drive the oracle
# exactly as it drives the CUDA engine (tests only).
# ----------------------------------------------------------------------------
class OracleEngine:
    """Implements whisperlivekit_b200.engine.WhisperEngine's session API on the CPU oracle."""

    backend = "oracle-cpu"

    def __init__(self, dims, state_dict: Dict[str, np.ndarray], align_heads, filters=None):
        from whisperlivekit_b200.dims import SpecialTokens
        from whisperlivekit_b200.weights import mel_filterbank
        self.dims = dims
        self.W = Weights(state_dict)
        self.specials = SpecialTokens.for_dims(dims)
        self.align_heads = [tuple(x) for x in align_heads]
        self.filters = _t(filters if filters is not None else mel_filterbank(dims.n_mels))
        self._s: Dict[int, dict] = {}
        self._next = 0

    # -- sessions ---------------------------------------------------------
    def open_session(self) -> int:
        sid = self._next
        self._next += 1
        self._s[sid] = dict(audio=np.zeros(0, np.float32), kv={}, xa=None, iters=[], logits=None,
                            sot_row=None, mel=None, content=0)
        return sid

    def close_session(self, sid: int) -> None:
        self._s.pop(sid)

    def fork_session(self, parent: int) -> int:
        """A beam row: the reference feeds one encoder output to beam_size decoder rows
        (simul_whisper.py:240-243 repeat_interleave of the tokens, xa broadcast in model.py:148-173)."""
        sid = self.open_session()
        self._s[sid]["parent"] = parent
        return sid

    def _enc(self, sid: int) -> dict:
        s = self._s[sid]
        return self._s[s["parent"]] if "parent" in s else s

    def gather_decoder(self, sids: Sequence[int], source_indices: Sequence[int]) -> None:
        """reference simul_whisper/beam.py:15-19: self-attention K/V rows re-indexed, nothing else."""
        snap = [{k: v for k, v in self._s[sid]["kv"].items() if ".cross_attn_" not in k} for sid in sids]
        for i, sid in enumerate(sids):
            kv = self._s[sid]["kv"]
            for k in [k for k in kv if ".cross_attn_" not in k]:
                del kv[k]
            kv.update({k: v.clone() for k, v in snap[source_indices[i]].items()})

    def append_audio(self, sid: int, pcm) -> None:
        s = self._s[sid]
        s["audio"] = np.concatenate([s["audio"], np.asarray(pcm, np.float32).reshape(-1)])

    def drop_audio(self, sid: int, n: int) -> None:
        s = self._s[sid]
        s["audio"] = s["audio"][n:]

    def clear_audio(self, sid: int) -> None:
        self._s[sid]["audio"] = np.zeros(0, np.float32)

    def audio_len(self, sid: int) -> int:
        return int(self._s[sid]["audio"].shape[0])

    def reset_decoder(self, sid: int) -> None:
        s = self._s[sid]
        s["kv"], s["iters"], s["logits"], s["sot_row"] = {}, [], None, None
        # the reference's clean_cache drops the cross K/V too; they are recomputed from the same xa

    # -- hot path ---------------------------------------------------------
    @torch.no_grad()
    def encode(self, sids: Sequence[int]) -> List[int]:
        out = []
        for sid in sids:
            s = self._s[sid]
            mel, content = encode_features(torch.from_numpy(s["audio"]), self.filters)
            s["mel"], s["content"] = mel, content
            s["xa"] = encoder_forward(self.W, self.dims, mel)
            s["kv"], s["iters"], s["logits"], s["sot_row"] = {}, [], None, None   # new infer epoch
            for f in self._s.values():
                if f.get("parent") == sid:
                    f["kv"], f["iters"], f["logits"], f["sot_row"] = {}, [], None, None
            out.append(content)
        return out

    @torch.no_grad()
    def decode(self, sids: Sequence[int], tokens: Sequence[Sequence[int]], sot_index: int = 0) -> None:
        for sid, toks in zip(sids, tokens):
            s = self._s[sid]
            t = torch.tensor([list(toks)], dtype=torch.long)
            logits, cross = decoder_forward(self.W, self.dims, t, self._enc(sid)["xa"], s["kv"])
            if not s["iters"]:
                s["sot_row"] = logits[0, sot_index].clone()
            s["logits"] = logits[0, -1].clone()
            s["iters"].append(cross)

    # -- LocalAgreement path (whisper.transcribe seam) ------------------------------------
    @torch.no_grad()
    def encode_mel(self, sid: int, mel, content_mel_len: int = 1500) -> None:
        s = self._s[sid]
        m = _t(np.asarray(mel, np.float32))[None]
        s["mel"], s["content"] = m, min(1500, int(content_mel_len))
        s["xa"] = encoder_forward(self.W, self.dims, m)
        s["kv"], s["iters"], s["logits"], s["sot_row"] = {}, [], None, None

    @torch.no_grad()
    def decode_all_logits(self, sid: int, tokens: Sequence[int], sot_index: int = 0) -> np.ndarray:
        s = self._s[sid]
        t = torch.tensor([list(tokens)], dtype=torch.long)
        logits, cross = decoder_forward(self.W, self.dims, t, s["xa"], s["kv"])
        if not s["iters"]:
            s["sot_row"] = logits[0, sot_index].clone()
        s["logits"] = logits[0, -1].clone()
        s["iters"].append(cross)
        return logits[0].numpy().copy()

    def read_align_rows(self, sid: int) -> np.ndarray:
        """softmax(qk) rows of the alignment heads over the whole epoch: [n_align, rows, 1500]."""
        s = self._s[sid]
        out = []
        for (l, h) in self.align_heads:
            out.append(torch.cat([F.softmax(it[l][0, h], dim=-1) for it in s["iters"]], dim=0))
        return torch.stack(out).numpy()

    def no_speech_prob(self, sids: Sequence[int]) -> List[float]:
        return [no_speech_prob(self._s[sid]["sot_row"], self.specials.no_speech) for sid in sids]

    def suppress(self, sids: Sequence[int], token_ids: Sequence[int]) -> None:
        for sid in sids:
            self._s[sid]["logits"][list(token_ids)] = -np.inf

    def add_logit_bias(self, sid: int, token_ids: Sequence[int], biases: Sequence[float]) -> None:
        lg = self._s[sid]["logits"]
        for tkn, b in zip(token_ids, biases):
            lg[tkn] = lg[tkn] + b

    def select(self, sids, suppress, first_ids=(), first_mask=None, biases=None, window_iters: int = 16):
        """The engine's fused "pick" call as the plain sequence of the reference's steps (align_att_base.py:229-243)."""
        for i, sid in enumerate(sids):
            if first_mask is not None and first_mask[i] and len(first_ids):
                self.suppress([sid], first_ids)
        self.suppress(sids, suppress)
        if biases is not None:
            for sid, b in zip(sids, biases):
                if b:
                    self.add_logit_bias(sid, [t for t, _ in b], [v for _, v in b])
        return self.greedy_and_align(sids, window_iters)

    @torch.no_grad()
    def greedy_and_align(self, sids: Sequence[int], window_iters: int = 16):
        """-> list of (next_token, logprob, most_attended_frame) per session."""
        res = []
        for sid in sids:
            s = self._s[sid]
            tok, lp = greedy_update(s["logits"])
            attn = process_cross_attention(s["iters"][-window_iters:], self.align_heads,
                                           self.dims.n_text_layer, self._enc(sid)["content"])
            frame = int(torch.argmax(attn[0, -1, :]))
            s["attn"] = attn
            res.append((tok, lp, frame))
        return res

    # -- debug taps ---------------------------------------------------------
    # word-timestamp math of the LocalAgreement path (same host-array entry points as WhisperEngine)
    def median_filter_host(self, x, width: int = 7):
        from oracle import timing_oracle
        return timing_oracle.median_filter(np.ascontiguousarray(x, np.float32), width)

    def dtw_host(self, x):
        from oracle import timing_oracle
        return timing_oracle.dtw(np.ascontiguousarray(x, np.float32))

    def read_mel(self, sid):        return self._s[sid]["mel"][0].numpy()
    def read_encoder(self, sid):    return self._s[sid]["xa"][0].numpy()
    def read_logits(self, sid):     return self._s[sid]["logits"].numpy()
    def read_sot_logits(self, sid): return self._s[sid]["sot_row"].numpy()
    def read_align_attn(self, sid): return self._s[sid]["attn"][0].numpy()
