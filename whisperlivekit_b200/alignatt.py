"""Host side of the AlignAtt seam (the reference's primary plugin surface).

Two things live here:

* ``AlignAttHooks`` -- the ~20 tensor hooks ``AlignAttBase`` declares abstract
  (reference whisperlivekit/simul_whisper/align_att_base.py:541-649), written
  against the engine session API (``engine.WhisperEngine``).  With
  WhisperLiveKit importable, ``plugin.make_b200_alignatt_class()`` mixes these
  into the reference's own ``AlignAttBase`` so its ``infer()`` and the
  SimulStreaming processor run unchanged on the B200 engine.
* ``StreamingAlignAtt`` -- a self-contained mirror of the control flow of
  ``AlignAttBase.infer`` (align_att_base.py:174-322) and ``AlignAtt.insert_audio``
  (simul_whisper.py:219-237) on token ids only (no tokenizer/text), for hosts
  where WhisperLiveKit is not installed (the GPU test box, bench.py).  Same
  names, same stop / rewind / suppression rules, same defaults.

Neither class contains tensor math: that is all behind the engine's C-ABI.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .dims import SpecialTokens

DEC_PAD = 50257                     # reference align_att_base.py:10
TOKENS_PER_SECOND = 50              # reference whisper/audio.py:22


@dataclass
class AlignAttConfig:
    """Mirror of reference simul_whisper/config.py:5-23 (hot-path fields) with the
    values core.py passes for the default SimulStreaming setup (config.py:104-108)."""
    frame_threshold: int = 25
    rewind_threshold: int = 200
    audio_max_len: float = 30.0
    audio_min_len: float = 0.0
    nonspeech_prob: float = 0.5
    beam_size: int = 1
    decoder_type: str = "greedy"
    language: str = "en"
    task: str = "transcribe"
    max_context_tokens: Optional[int] = None
    dry_penalty: bool = True


def dry_penalties(seq: Sequence[int], eot: int) -> List[Tuple[int, float]]:
    """DRY repetition penalty, host part (reference align_att_base.py:492-537).
    Returns [(token, amount_to_subtract)]."""
    seq = list(seq)
    if len(seq) < 5:
        return []
    last = seq[-1]
    if last >= eot:
        return []
    penalties = {}
    for i in range(len(seq) - 2, -1, -1):
        if seq[i] != last:
            continue
        next_tok = seq[i + 1]
        if next_tok >= eot:
            continue
        length = 1
        while length < 50:
            j, k = i - length, len(seq) - 1 - length
            if j < 0 or k <= i:
                break
            if seq[j] != seq[k] or seq[j] >= eot:
                break
            length += 1
        if next_tok not in penalties or length > penalties[next_tok]:
            penalties[next_tok] = length
    return [(tok, 1.0 * 2.0 ** (length - 2)) for tok, length in penalties.items() if length >= 2]


@dataclass
class InferTrace:
    """What one ``infer`` did -- compared token-for-token with the reference in tests."""
    content_mel_len: int = 0
    prefix_len: int = 0
    no_speech: bool = False
    no_speech_prob: float = 0.0
    step_tokens: List[int] = field(default_factory=list)      # argmax token of every loop iteration
    step_frames: List[int] = field(default_factory=list)      # most attended frame of every iteration
    step_logprobs: List[float] = field(default_factory=list)
    new_tokens: List[int] = field(default_factory=list)       # hypothesis appended to state.tokens
    timestamps: List[float] = field(default_factory=list)
    stop: str = ""


class StreamingAlignAtt:
    """Per-session AlignAtt policy over an engine session (greedy, beam_size 1)."""

    def __init__(self, engine, cfg: Optional[AlignAttConfig] = None, lang_index: int = 0):
        self.engine = engine
        self.cfg = cfg or AlignAttConfig()
        if self.cfg.decoder_type != "greedy" or self.cfg.beam_size != 1:
            raise NotImplementedError("StreamingAlignAtt implements the greedy policy (reference default beams=1)")
        self.sp: SpecialTokens = engine.specials
        self.sid = engine.open_session()
        self.max_text_len = engine.dims.n_text_ctx
        self.max_context_tokens = self.cfg.max_context_tokens or self.max_text_len
        self.initial_tokens = list(self.sp.sot_sequence_including_notimestamps(lang_index, self.cfg.task))
        self.sot_index = 0                                   # tokenizer.sot_sequence.index(sot)
        self.suppress_tokens = self.sp.alignatt_suppress_tokens()
        self.segments: List[int] = []                        # sample counts of buffered chunks
        self.tokens: List[List[int]] = [list(self.initial_tokens)]
        self.context: List[int] = []                         # token ids moved out of the window
        self.last_attend_frame = -self.cfg.rewind_threshold
        self.cumulative_time_offset = 0.0
        self.first_timestamp: Optional[float] = None
        self.closed = False

    # -- audio window ------------------------------------------------------
    def segments_len(self) -> float:
        return sum(self.segments) / 16000

    def insert_audio(self, segment: Optional[np.ndarray] = None) -> float:
        """reference simul_whisper.py:219-237."""
        if segment is not None:
            seg = np.ascontiguousarray(np.asarray(segment, dtype=np.float32).reshape(-1))
            self.segments.append(int(seg.shape[0]))
            self.engine.append_audio(self.sid, seg)
        removed_len = 0.0
        segments_len = self.segments_len()
        while len(self.segments) > 1 and segments_len > self.cfg.audio_max_len:
            removed = self.segments[0]
            removed_len = removed / 16000
            segments_len -= removed_len
            self.last_attend_frame -= int(TOKENS_PER_SECOND * removed_len)
            self.cumulative_time_offset += removed_len
            self.segments = self.segments[1:]
            self.engine.drop_audio(self.sid, removed)
            if len(self.tokens) > 1:
                self.context.extend(self.tokens[1])
                self.tokens = [list(self.initial_tokens)] + self.tokens[2:]
        return removed_len

    def refresh_segment(self, complete: bool = False) -> None:
        """reference align_att_base.py:115-132 (token-id form)."""
        self.tokens = [list(self.initial_tokens)]
        self.last_attend_frame = -self.cfg.rewind_threshold
        self.cumulative_time_offset = 0.0
        self.context = []
        if not complete and len(self.segments) > 2:
            drop = sum(self.segments[:-2])
            self.segments = self.segments[-2:]
            self.engine.drop_audio(self.sid, drop)
        else:
            self.segments = []
            self.engine.clear_audio(self.sid)

    def trim_context(self) -> None:
        """reference align_att_base.py:100-113; the reference trims whole words of the
        context *text*, this id-only mirror trims one token at a time."""
        c = len(self.context)
        l = sum(len(t) for t in self.tokens) + c
        while c > self.max_context_tokens or l > self.max_text_len - 20:
            if not self.context:
                break
            self.context.pop(0)
            c -= 1
            l -= 1

    def _current_tokens(self) -> List[int]:
        """reference simul_whisper.py:239-254."""
        toks: List[int] = []
        if self.context:
            toks += [self.sp.sot_prev] + self.context
        for t in self.tokens:
            toks += t
        return toks

    # -- the template infer() ------------------------------------------------
    def infer(self, is_last: bool = False) -> InferTrace:
        """reference align_att_base.py:174-322 (control flow), one engine call per hook."""
        eng, sid, cfg = self.engine, self.sid, self.cfg
        tr = InferTrace()
        if len(self.segments) == 0:
            tr.stop = "no_segments"
            return tr
        if self.segments_len() < cfg.audio_min_len:
            tr.stop = "minseglen"
            return tr

        content_mel_len = eng.encode([sid])[0]                               # _encode
        tr.content_mel_len = content_mel_len
        self.trim_context()
        current_tokens = self._current_tokens()
        token_len_before = len(current_tokens)
        tr.prefix_len = token_len_before

        completed = False
        new_segment = True
        l_absolute_timestamps: List[float] = []
        audio_duration_s = self.segments_len()
        max_tokens = max(50, int(audio_duration_s * 15 * 1.5))
        tokens_produced = 0
        iters = 0

        while not completed and len(current_tokens) < self.max_text_len:
            tokens_produced += 1
            if tokens_produced > max_tokens:
                current_tokens = current_tokens[:token_len_before]
                tr.stop = "loop_detection"
                break
            feed = current_tokens if new_segment else current_tokens[-1:]
            eng.decode([sid], [feed], sot_index=self.sot_index)              # _get_logits_and_cross_attn
            iters += 1
            if new_segment:
                p = eng.no_speech_prob([sid])[0]                             # _check_no_speech
                tr.no_speech_prob = p
                if p > cfg.nonspeech_prob:
                    tr.no_speech = True
                    tr.stop = "no_speech"
                    break
                eng.suppress([sid], [self.sp.blank, self.sp.eot])            # _suppress_blank_tokens
            new_segment = False
            eng.suppress([sid], self.suppress_tokens)                        # _apply_token_suppression
            if cfg.dry_penalty:
                pen = dry_penalties(current_tokens, self.sp.eot)             # _apply_dry_penalty
                if pen:
                    eng.add_logit_bias(sid, [t for t, _ in pen], [-a for _, a in pen])
            tok, logprob, frame = eng.greedy_and_align([sid], window_iters=16)[0]
            if current_tokens[-1] == self.sp.eot:                            # decoding.py:282
                tok = self.sp.eot
            current_tokens = current_tokens + [tok]
            completed = tok == self.sp.eot
            tr.step_tokens.append(tok)
            tr.step_frames.append(frame)
            tr.step_logprobs.append(logprob)
            l_absolute_timestamps.append(frame * 0.02 + self.cumulative_time_offset)

            if completed:
                current_tokens = current_tokens[:-1]
                tr.stop = "eot"
                break
            if (not is_last) and self.last_attend_frame - frame > cfg.rewind_threshold:
                if len(current_tokens) > 1 and current_tokens[-2] >= DEC_PAD:
                    self.last_attend_frame = frame
                else:
                    self.last_attend_frame = -cfg.rewind_threshold
                    current_tokens = [t for seg in self.tokens for t in seg]  # _rewind_tokens
                    tr.stop = "rewind"
                    break
            else:
                self.last_attend_frame = frame
            if content_mel_len - frame <= (4 if is_last else cfg.frame_threshold):
                current_tokens = current_tokens[:-1]
                tr.stop = "attention_end"
                break
        else:
            tr.stop = tr.stop or "max_text_len"

        new_hypothesis = current_tokens[token_len_before:]                   # always_fire: keep all
        n = len(new_hypothesis)
        ts = [float(x) for x in l_absolute_timestamps[:n]]
        if len(ts) < n:
            ts += [ts[-1] if ts else 0.0] * (n - len(ts))
        tr.new_tokens = list(new_hypothesis)
        tr.timestamps = ts
        self.tokens.append(list(new_hypothesis))
        if len(l_absolute_timestamps) >= 2 and self.first_timestamp is None:
            self.first_timestamp = l_absolute_timestamps[0]
        return tr

    def close(self) -> None:
        if not self.closed:
            self.engine.close_session(self.sid)
            self.closed = True
