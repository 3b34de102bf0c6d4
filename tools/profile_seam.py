#!/usr/bin/env python
"""Where does the host spend its time in the real-time paced seam run?  cProfile around bench.seam_probe at B streams.
    python tools/profile_seam.py [streams] [ticks]"""
import cProfile
import io
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import bench
from whisperlivekit_b200.dims import ALIGNMENT_HEADS, DIMS
from whisperlivekit_b200.engine import WhisperEngine
from whisperlivekit_b200.weights import synthetic_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dims = DIMS["large-v3"]
eng = WhisperEngine(dims, synthetic_state_dict(dims, seed=0), ALIGNMENT_HEADS["large-v3"], precision="bf16", max_sessions=B, max_batch=B)
rng = np.random.default_rng(1)
mode = sys.argv[3] if len(sys.argv) > 3 else "continuous"
bench.seam_probe(eng, B, 2, 2, rng, mode=mode)            # warm up (graphs, tensor maps)
pr = cProfile.Profile()
pr.enable()
r = bench.seam_probe(eng, B, ticks, 4, rng, mode=mode)
pr.disable()
print(mode, {k: r[k] for k in ("streams", "ok", "p50_latency_s", "p95_latency_s", "wall_s", "engine_calls", "cohorts", "mean_cohort", "mean_decode_iterations")})
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18)
print(s.getvalue()[:4500])
eng.close()
