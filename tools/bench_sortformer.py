"""Config-4 diarization leg alone: `streams` Sortformer streams, one 1.0 s step each per tick (host audio in, segments
out through the device run-length kernel).  Prints one JSON line.  python tools/bench_sortformer.py [streams] [ticks]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from whisperlivekit_b200.sortformer_dims import SORTFORMER_DIMS, synthetic_sortformer_state_dict, synthetic_two_speaker_audio
from whisperlivekit_b200.sortformer_engine import B200SortformerDiarization, B200SortformerDiarizationOnline, diarize_batch


def run(streams=64, ticks=20, precision="bf16", device=0):
    d = SORTFORMER_DIMS["diar_streaming_sortformer_4spk-v2"]
    shared = B200SortformerDiarization(d, synthetic_sortformer_state_dict(d, 0), precision=precision, device=device,
                                       max_sessions=streams, max_batch=streams)
    ons = [B200SortformerDiarizationOnline(shared) for _ in range(streams)]
    audio = synthetic_two_speaker_audio(ticks + 2.0, seed=2)
    per = []
    for k in range(ticks):
        chunks = [np.roll(audio[k * 16000:(k + 1) * 16000], 37 * i) for i in range(streams)]
        t0 = time.perf_counter()
        segs = diarize_batch(ons, chunks)
        per.append(time.perf_counter() - t0)
    steady = np.asarray(per[ticks // 2:])              # caches full: 188 + 188 + 25 rows per stream
    st = shared.engine.read_state(ons[0].sid)
    out = dict(workload=f"streaming Sortformer 4spk-v2 geometry, {streams} streams, 1.0 s step per stream and tick, host audio in / "
                        "speaker segments out", streams=streams, ms_per_tick_first=per[0] * 1e3, ms_per_tick_steady=float(steady.mean() * 1e3),
               ms_per_tick_p95=float(np.percentile(steady, 95) * 1e3), realtime_streams=float(streams * 1.0 / steady.mean()),
               spkcache_len=st["spkcache_len"], fifo_len=st["fifo_len"], segments_last_tick=int(sum(len(s) for s in segs)),
               memory=shared.engine.memory(), precision=precision)
    for o in ons:
        o.close()
    shared.close()
    return out


if __name__ == "__main__":
    s = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    t = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    print(json.dumps(run(s, t)))
