#!/usr/bin/env python
"""Time the decoder token step alone (large-v3 dims): python tools/bench_decode_step.py [streams] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from whisperlivekit_b200.dims import DIMS, ALIGNMENT_HEADS
from whisperlivekit_b200.engine import WhisperEngine
from whisperlivekit_b200.weights import synthetic_state_dict, synthetic_audio
B = int(sys.argv[1]) if len(sys.argv) > 1 else 96
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dims = DIMS["large-v3"]
eng = WhisperEngine(dims, synthetic_state_dict(dims, seed=0), ALIGNMENT_HEADS["large-v3"], precision="bf16", max_sessions=B, max_batch=B)
sids = [eng.open_session() for _ in range(B)]
a = synthetic_audio(30.0, seed=7)
for s in sids:
    eng.append_audio(s, a)
eng.encode(sids)
sp = eng.specials
prefix = list(sp.sot_sequence_including_notimestamps()) + list(range(1000, 1044))
eng.decode(sids, [prefix] * B)
for _ in range(3):
    eng.decode(sids, [[1234]] * B)
eng.sync()
eng.timer_record(0)
for _ in range(steps):
    eng.decode(sids, [[1234]] * B)
eng.timer_record(1)
eng.sync()
ms = eng.timer_elapsed_ms(0, 1) / steps
print(f"decode token step B={B}: {ms:.3f} ms/step  ({ms*1e3/(32*11+3):.1f} us per launch)  PDL={os.environ.get('WLK_PDL','1')} GRAPHS={os.environ.get('WLK_GRAPHS','1')} VAR={os.environ.get('WLK_GEMM_VARIANT','0')}")
eng.profile_reset(); eng.profile_enable(True)
for _ in range(4):
    eng.decode(sids, [[1234]] * B)
eng.sync()
p = eng.profile_read()
print("  profiled classes (ms/step): " + ", ".join(f"{k}={v['ms']/4:.2f}/{v['launches']//4}" for k, v in p.items() if v["launches"]))
