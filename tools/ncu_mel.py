#!/usr/bin/env python
"""Target for an ncu capture of the log-mel kernels: 64 sessions x 30 s of audio, 128 mel bins, a toy model behind
(the mel kernels do not depend on the model size):
    ncu --set full --clock-control none -k regex:mel_power_kernel -c 1 -o gpurun_out/mel python tools/ncu_mel.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from whisperlivekit_b200.dims import ModelDimensions
from whisperlivekit_b200.engine import WhisperEngine
from whisperlivekit_b200.weights import synthetic_audio, synthetic_state_dict
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dims = ModelDimensions(128, 1500, 128, 2, 1, 51866, 448, 128, 2, 1)
eng = WhisperEngine(dims, synthetic_state_dict(dims, seed=0), [(0, 0)], precision="bf16", max_sessions=B, max_batch=B)
sids = [eng.open_session() for _ in range(B)]
a = synthetic_audio(30.0, seed=1)
for s in sids:
    eng.append_audio(s, a)
for _ in range(3):
    eng.encode(sids)
eng.sync()
eng.profile_reset(); eng.profile_enable(True)
for _ in range(5):
    eng.encode(sids)
eng.sync()
p = eng.profile_read()["mel"]
print(f"mel: {p['ms'] / 5 * 1e3:.1f} us per batch of {B} windows = {p['ms'] / 5 / B * 1e3:.2f} us per stream-window; "
      f"{B * 2.69e6 / (p['ms'] / 5 * 1e-3) / 1e9:.1f} GB/s of algorithmic bytes")
