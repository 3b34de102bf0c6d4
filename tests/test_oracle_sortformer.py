"""CPU: the Sortformer oracle (oracle/sortformer_oracle.py, PARITY UNPINNED -- NeMo is absent) checked against what can be
checked without NeMo: an independent numpy statement of the front end, the conv-stem length rule the reference relies on
(12 / 23 prediction rows per chunk, sortformer_backend.py:332 "#12"), and the invariants of the speaker-cache update
(lengths, ordering, silence padding, quota per speaker)."""
import math

import numpy as np
import torch

from oracle.sortformer_oracle import OracleDiarizer, SortformerOracle, log_mel
from whisperlivekit_b200.sortformer_dims import (SORTFORMER_DIMS, subsampled_len, synthetic_sortformer_state_dict,
                                                 synthetic_two_speaker_audio)
from whisperlivekit_b200.weights import mel_filterbank


def test_front_end_against_a_direct_dft():
    d = SORTFORMER_DIMS["diar_streaming_sortformer_4spk-v2"]
    a = synthetic_two_speaker_audio(1.0, seed=5)
    got = log_mel(a, d).numpy()
    assert got.shape == (128, 101)                                  # "16 000 samples -> 101 frames"
    x = a.astype(np.float64)
    y = np.concatenate([x[:1], x[1:] - 0.97 * x[:-1]])
    p = np.pad(y, 256, mode="reflect")
    win = np.zeros(512)
    win[56:456] = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(400) / 399)
    fb = mel_filterbank(128, 16000, 512).astype(np.float64)
    for t in (0, 1, 50, 100):
        spec = np.fft.rfft(p[t * 160: t * 160 + 512] * win)
        want = np.log(fb @ (spec.real ** 2 + spec.imag ** 2) + 2.0 ** -24)
        assert np.abs(got[:, t] - want).max() < 2e-3, t


def test_chunk_row_counts_match_the_reference_comment():
    assert subsampled_len(101) == 13 and subsampled_len(200) == 25
    d = SORTFORMER_DIMS["micro"]
    o = OracleDiarizer(SortformerOracle(d, synthetic_sortformer_state_dict(d, 2)))
    a = synthetic_two_speaker_audio(3.0, seed=1)
    assert o.step(a[:16000]).shape == (12, d.n_spk)                 # sortformer_backend.py:332: `#12`
    assert o.step(a[16000:32000]).shape == (23, d.n_spk)
    assert o.total_preds.shape[0] == 35


def test_streaming_lengths_and_fifo_order():
    d = SORTFORMER_DIMS["micro"]
    m = SortformerOracle(d, synthetic_sortformer_state_dict(d, 3))
    st = m.init_state()
    rng = np.random.default_rng(0)
    seen = []
    for k in range(8):
        tc = 13 if k == 0 else 25
        lc, rc = (0, 1) if k == 0 else (1, 1)
        chunk = torch.from_numpy(rng.standard_normal((tc, d.d_model)).astype(np.float32))
        T = st["spkcache_len"] + st["fifo_len"] + tc
        preds = torch.from_numpy(rng.random((T, d.n_spk)).astype(np.float32))
        fl0, sl0 = st["fifo_len"], st["spkcache_len"]
        cp = m.streaming_update(st, chunk, preds, lc, rc)
        clen = tc - lc - rc
        assert cp.shape == (clen, d.n_spk)
        assert torch.equal(cp, preds[sl0 + fl0 + lc: sl0 + fl0 + lc + clen])
        assert 0 <= st["fifo_len"] <= d.fifo_len and 0 <= st["spkcache_len"] <= d.spkcache_len
        seen.append(chunk[lc: lc + clen])
        # the FIFO always ends with the newest chunk rows, in order
        tail = torch.cat(seen)[-st["fifo_len"]:] if st["fifo_len"] else torch.zeros(0, d.d_model)
        assert torch.equal(st["fifo"][: st["fifo_len"]], tail)
        assert torch.count_nonzero(st["fifo"][st["fifo_len"]:]) == 0
    assert st["spkcache_len"] == d.spkcache_len                     # the cache has filled and been compressed


def test_compression_keeps_speech_per_speaker_and_pads_silence():
    d = SORTFORMER_DIMS["small"]
    m = SortformerOracle(d, synthetic_sortformer_state_dict(d, 4))
    n = d.spkcache_len + 30
    rng = np.random.default_rng(1)
    emb = torch.arange(n, dtype=torch.float32)[:, None].repeat(1, d.d_model)     # row r is filled with the value r
    preds = torch.full((n, d.n_spk), 0.05)
    spk = rng.integers(0, 2, n)                                                  # speakers 0 and 1 alternate, 2 and 3 never speak
    conf = 0.6 + 0.39 * rng.random(n)
    for r in range(n):
        preds[r, spk[r]] = float(conf[r])
    preds[:5] = 0.01                                                             # leading silence
    sil = torch.full((d.d_model,), -7.0)
    e, p = m.compress_spkcache(emb, preds, sil)
    assert e.shape == (d.spkcache_len, d.d_model) and p.shape == (d.spkcache_len, d.n_spk)
    rows = e[:, 0]
    is_sil = rows == -7.0
    assert int(is_sil.sum()) >= d.n_spk * d.spkcache_sil_frames_per_spk          # the +inf pads become silence rows
    assert torch.all(p[is_sil] == 0)
    kept = rows[~is_sil].long()
    assert torch.all(kept >= 5)                                                  # silence frames are never kept as speech
    assert torch.equal(p[~is_sil], preds[kept])
    # speaker-major order: first the rows kept for speaker 0 (ascending), then speaker 1 (ascending)
    owner = torch.from_numpy(spk)[kept]
    change = int((owner[1:] != owner[:-1]).sum())
    assert change == 1 and owner[0] == 0
    for s in (0, 1):
        k = kept[owner == s]
        assert torch.all(k[1:] > k[:-1])
    # the two active speakers share the cache about evenly
    assert abs(int((owner == 0).sum()) - int((owner == 1).sum())) <= d.spkcache_len // 4


def test_silence_profile_running_mean():
    d = SORTFORMER_DIMS["micro"]
    m = SortformerOracle(d, synthetic_sortformer_state_dict(d, 5))
    emb = torch.randn(10, d.d_model)
    preds = torch.full((10, d.n_spk), 0.3)
    preds[[2, 7]] = 0.01
    mean, cnt = m.silence_profile(torch.zeros(d.d_model), 0, emb, preds)
    assert cnt == 2 and torch.allclose(mean, emb[[2, 7]].mean(0), atol=1e-6)
    mean2, cnt2 = m.silence_profile(mean, cnt, emb[:3], preds[:3])
    assert cnt2 == 3 and torch.allclose(mean2, (emb[2] * 2 + emb[7]) / 3, atol=1e-6)
    assert m.silence_profile(mean2, cnt2, emb[:2], preds[:2]) == (mean2, cnt2)


def test_rel_shift_is_the_relative_position_lookup():
    """_rel_attention's gather equals NeMo's pad-and-reshape rel_shift on a random matrix"""
    H, T = 2, 7
    bd = torch.randn(1, H, T, 2 * T - 1)
    # NeMo: pad one zero column on the left, view as (b, h, 2T, T), drop the first row, view back, keep the first T columns
    x = torch.nn.functional.pad(bd, (1, 0))
    x = x.view(1, H, 2 * T, T)[:, :, 1:].reshape(1, H, T, 2 * T - 1)[:, :, :, :T]
    idx = (T - 1) - torch.arange(T)[:, None] + torch.arange(T)[None, :]
    got = torch.gather(bd[0], 2, idx[None].expand(H, T, T))
    assert torch.equal(got, x[0])
