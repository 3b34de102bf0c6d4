#!/usr/bin/env python
"""BASELINE config 5 (secondary): Qwen3-ASR-0.6B causal audio tower, 0.25 s chunks (25 mel frames), the encoder fires per
192-frame block; N streams per GPU with staggered phases.  Reports real-time streams = audio seconds per wall second.
    python tools/bench_qwen.py [streams] [ticks]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from whisperlivekit_b200.qwen_dims import QWEN_DIMS, synthetic_tower_state_dict
from whisperlivekit_b200.qwen_engine import QwenTowerEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 48
dims = QWEN_DIMS["qwen3-asr-0.6b"]
eng = QwenTowerEngine(dims, synthetic_tower_state_dict(dims, seed=0), precision=os.environ.get("PREC", "bf16"), max_sessions=B, max_batch=B)
sids = [eng.open_session() for _ in range(B)]
rng = np.random.default_rng(0)
mel = np.clip(0.3 + rng.standard_normal((4096, dims.n_mels)).astype(np.float32), -1, 1.5)
AUDIO = os.environ.get("AUDIO", "1") == "1"            # 1: raw audio in (device mel front end + tower); 0: mel frames in
if AUDIO:
    from whisperlivekit_b200.weights import synthetic_audio
    eng.load_mel_filters()
    pcm = synthetic_audio(40.0, seed=3)
phase = rng.integers(0, 192, B)
eng.forward_chunk(sids, [mel[: int(p)] for p in phase])             # stagger the block boundaries


def tick(k):
    if AUDIO:                                                       # 0.25 s = 4000 samples per stream
        chunks = [pcm[(4000 * k + 997 * i) % 500000: (4000 * k + 997 * i) % 500000 + 4000] for i in range(B)]
        return eng.forward_chunk(sids, eng.mel_append(sids, chunks))
    return eng.forward_chunk(sids, [mel[(37 * k + i) % 4000: (37 * k + i) % 4000 + 25] for i in range(B)])


for k in range(8):                                                  # warm-up: one full block period
    tick(k)
per_tick, rows = [], 0
t_all = time.perf_counter()
for k in range(ticks):
    t0 = time.perf_counter()
    out = tick(8 + k)
    per_tick.append(time.perf_counter() - t0)
    rows += sum(o.shape[0] for o in out)
wall = time.perf_counter() - t_all
per_tick = np.asarray(per_tick) * 1e3
flop_per_step = 2 * (9 * dims.conv_channels * (64 * 4) + 2304 * 0 + dims.conv_channels * 9 * dims.conv_channels * (64 + 16)
                     + dims.conv_features * dims.d_model + dims.n_layer * (4 * dims.d_model ** 2 + 2 * dims.d_model * dims.ffn_dim)
                     + dims.d_model ** 2 + dims.d_model * dims.out_dim)
print(json.dumps(dict(workload="qwen3-asr-0.6b causal audio tower, 0.25 s chunks, block 192 frames" + (", raw audio in (device log-mel)" if AUDIO else ", mel frames in"), streams=B, ticks=ticks,
                      precision=os.environ.get("PREC", "bf16"), realtime_streams=B * 0.25 * ticks / wall,
                      ms_per_tick_mean=float(per_tick.mean()), ms_per_tick_p95=float(np.percentile(per_tick, 95)),
                      ms_per_tick_max=float(per_tick.max()), encoder_steps=rows, gflop_per_step=flop_per_step / 1e9,
                      achieved_tflops=rows * flop_per_step / wall / 1e12, memory=eng.memory())))
