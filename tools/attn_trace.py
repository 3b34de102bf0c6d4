#!/usr/bin/env python
"""Where does a CTA of the tcgen05 encoder attention spend its time?  One CTA in the middle of the grid stamps clock64() at
its pipeline hand-offs (wlk_op_encoder_attention_trace); prints per key tile, in SM clocks relative to the CTA's first stamp.
    python tools/attn_trace.py [streams]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from whisperlivekit_b200 import _lib as L
from whisperlivekit_b200.dims import ModelDimensions
from whisperlivekit_b200.engine import WhisperEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
d, H = 1280, 20
e = WhisperEngine(ModelDimensions(80, 1500, d, H, 1, 51864, 448, 64, 1, 1), None, [(0, 0)], precision="bf16", max_sessions=1, max_batch=1)
qkv = (torch.randn(B * 1500, 3 * d, device="cuda") * 0.35).bfloat16()
out = torch.empty(B * 1500, d, device="cuda", dtype=torch.bfloat16)
torch.cuda.synchronize()
st = np.zeros((12, 8), np.int64)
for _ in range(3):
    L.check(e.lib.wlk_op_encoder_attention_trace(e.h, C.c_void_p(qkv.data_ptr()), B, C.c_void_p(out.data_ptr()), st.ctypes.data_as(C.c_void_p)))
t0 = st[0, 0]
r = st - t0
print("tile | mma: K/V ready  S issue  PV issue  PV issued | softmax w2: S ready  exps done  arrived | w9 arrived || "
      "S issue->ready  softmax  PV issue cost  wait for K/V(j+1)")
for j in range(12):
    nxt_kv = r[j + 1, 5] if j < 11 else 0
    print(f"{j:4d} | {r[j,5]:8d} {r[j,0]:8d} {r[j,1]:8d} {r[j,7]:8d} | {r[j,2]:8d} {r[j,3]:8d} {r[j,4]:8d} | {r[j,6]:8d} || "
          f"{r[j,2]-r[j,0]:6d} {max(r[j,4],r[j,6])-r[j,2]:6d} {r[j,7]-r[j,1]:6d} {nxt_kv-r[j,7] if j < 11 else 0:6d}")
print("mean tile period:", (r[11, 0] - r[1, 0]) / 10.0, "clk")
