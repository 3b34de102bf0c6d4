"""Whisper model dimensions and special-token ids for the streaming hot path.

Mirrors ``ModelDimensions`` (reference whisperlivekit/whisper/model.py:25-36) and
the special-token arithmetic of ``get_encoding`` / ``Tokenizer``
(reference whisperlivekit/whisper/tokenizer.py:335-368, 141-158, 176-239).
Only the ids are reproduced here: the engine never needs text, so tiktoken is
not a dependency of the hot path.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict
from typing import List, Tuple

SAMPLE_RATE = 16000
N_FFT = 400
HOP_LENGTH = 160
N_SAMPLES = 480000      # 30 s window           (reference whisper/audio.py:17)
N_FRAMES = 3000         # mel frames per window (reference whisper/audio.py:18)
N_FREQ = 201
D_HEAD = 64             # every released Whisper size uses 64-wide heads


@dataclass(frozen=True)
class ModelDimensions:
    n_mels: int
    n_audio_ctx: int
    n_audio_state: int
    n_audio_head: int
    n_audio_layer: int
    n_vocab: int
    n_text_ctx: int
    n_text_state: int
    n_text_head: int
    n_text_layer: int

    def as_tuple(self) -> Tuple[int, ...]:
        return tuple(asdict(self).values())

    @property
    def is_multilingual(self) -> bool:          # reference whisper/model.py:398-399
        return self.n_vocab >= 51865

    @property
    def num_languages(self) -> int:             # reference whisper/model.py:402-403
        return self.n_vocab - 51765 - int(self.is_multilingual)


# name -> dims. "micro"/"nano" are test-only geometries (true vocab, so the
# special-token layout is real, but small enough for CPU golden generation).
DIMS = {
    "nano":     ModelDimensions(80, 1500, 64, 1, 1, 51864, 448, 64, 1, 2),
    "micro":    ModelDimensions(80, 1500, 128, 2, 2, 51864, 448, 128, 2, 2),
    "micro-ml": ModelDimensions(128, 1500, 128, 2, 2, 51866, 448, 128, 2, 2),
    "tiny":     ModelDimensions(80, 1500, 384, 6, 4, 51865, 448, 384, 6, 4),
    "tiny.en":  ModelDimensions(80, 1500, 384, 6, 4, 51864, 448, 384, 6, 4),
    "base":     ModelDimensions(80, 1500, 512, 8, 6, 51865, 448, 512, 8, 6),
    "base.en":  ModelDimensions(80, 1500, 512, 8, 6, 51864, 448, 512, 8, 6),
    "small":    ModelDimensions(80, 1500, 768, 12, 12, 51865, 448, 768, 12, 12),
    "medium":   ModelDimensions(80, 1500, 1024, 16, 24, 51865, 448, 1024, 16, 24),
    "large-v2": ModelDimensions(80, 1500, 1280, 20, 32, 51865, 448, 1280, 20, 32),
    "large-v3": ModelDimensions(128, 1500, 1280, 20, 32, 51866, 448, 1280, 20, 32),
    "large-v3-turbo": ModelDimensions(128, 1500, 1280, 20, 32, 51866, 448, 1280, 20, 4),
}


@dataclass(frozen=True)
class SpecialTokens:
    """Token ids derived from the vocabulary size alone."""
    eot: int
    sot: int
    lang_begin: int
    num_languages: int
    translate: int
    transcribe: int
    sot_lm: int
    sot_prev: int
    no_speech: int
    no_timestamps: int
    timestamp_begin: int
    blank: int = 220          # tokenizer.encode(" ") == [220] in both BPE vocabularies
    multilingual: bool = False

    @staticmethod
    def for_dims(dims: ModelDimensions) -> "SpecialTokens":
        ml = dims.is_multilingual
        nl = dims.num_languages
        base = 50257 if ml else 50256           # number of BPE ranks
        eot = base
        sot = base + 1
        lang_begin = base + 2
        translate = lang_begin + nl
        return SpecialTokens(
            eot=eot, sot=sot, lang_begin=lang_begin, num_languages=nl,
            translate=translate, transcribe=translate + 1, sot_lm=translate + 2,
            sot_prev=translate + 3, no_speech=translate + 4,
            no_timestamps=translate + 5, timestamp_begin=translate + 6,
            multilingual=ml,
        )

    @property
    def all_language_tokens(self) -> Tuple[int, ...]:
        return tuple(range(self.lang_begin, self.lang_begin + self.num_languages))

    def sot_sequence(self, lang_index: int = 0, task: str = "transcribe") -> Tuple[int, ...]:
        """reference whisper/tokenizer.py:151-158 (lang_index 0 == "en")."""
        if not self.multilingual:
            return (self.sot,)
        task_tok = self.transcribe if task == "transcribe" else self.translate
        return (self.sot, self.lang_begin + lang_index, task_tok)

    def sot_sequence_including_notimestamps(self, lang_index: int = 0,
                                            task: str = "transcribe") -> Tuple[int, ...]:
        return self.sot_sequence(lang_index, task) + (self.no_timestamps,)

    def alignatt_suppress_tokens(self) -> List[int]:
        """The set AlignAtt masks on every step (reference simul_whisper.py:161-172)."""
        s = {self.transcribe, self.translate, self.sot, self.sot_prev, self.sot_lm,
             self.no_timestamps, self.no_speech, *self.all_language_tokens}
        return sorted(s)


# Alignment heads of the released checkpoints, decoded from the base85+gzip
# boolean dumps the reference keeps (whisper/__init__.py:39-54), as
# (layer, head) pairs in row-major order == reference iteration order
# (simul_whisper.py:151-159).  Generated by oracle/make_golden.py and pinned by
# tests/test_oracle_golden.py::test_alignment_heads_match_reference.
ALIGNMENT_HEADS = {
    "tiny.en": [(1, 0), (2, 0), (2, 5), (3, 0), (3, 1), (3, 2), (3, 3), (3, 4)],
    "tiny": [(2, 2), (3, 0), (3, 2), (3, 3), (3, 4), (3, 5)],
    "base.en": [(3, 3), (4, 7), (5, 1), (5, 5), (5, 7)],
    "base": [(3, 1), (4, 2), (4, 3), (4, 7), (5, 1), (5, 2), (5, 4), (5, 6)],
    "small": [(5, 3), (5, 9), (8, 0), (8, 4), (8, 7), (8, 8), (9, 0), (9, 7), (9, 9), (10, 5)],
    "medium": [(13, 15), (15, 4), (15, 15), (16, 1), (20, 0), (23, 4)],
    "large-v2": [(10, 12), (13, 17), (16, 11), (16, 12), (16, 13), (17, 15), (17, 16), (18, 4),
                 (18, 11), (18, 19), (19, 11), (21, 2), (21, 3), (22, 3), (22, 9), (22, 12),
                 (23, 5), (23, 7), (23, 13), (25, 5), (26, 1), (26, 12), (27, 15)],
    "large-v3": [(7, 0), (10, 17), (12, 18), (13, 12), (16, 1), (17, 14), (19, 11),
                 (21, 4), (24, 1), (25, 6)],
    "large-v3-turbo": [(2, 4), (2, 11), (3, 3), (3, 6), (3, 11), (3, 14)],
}


def default_alignment_heads(dims: ModelDimensions) -> List[Tuple[int, int]]:
    """All heads of the upper half of the decoder (reference whisper/model.py:357-361)."""
    return [(l, h) for l in range(dims.n_text_layer // 2, dims.n_text_layer)
            for h in range(dims.n_text_head)]
