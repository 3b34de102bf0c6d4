"""CPU: the host logic of bench.py that does not need a GPU -- the real-time paced seam probes (closed cohorts and continuous
batching) over the oracle engine, their summary rule (p95 < chunk period, no backlog growth), and the contract pieces the driver
reads (configs, the reference arm's availability switch)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from golden_util import case_setup


def test_seam_probes_run_over_the_oracle_engine():
    import bench
    from oracle import whisper_oracle as wo
    g, dims, sd, audio, heads = case_setup("micro")
    eng = wo.OracleEngine(dims, sd, heads)
    eng.max_batch = 8
    rng = np.random.default_rng(0)
    for mode in ("cohort", "continuous"):
        r = bench.seam_probe(eng, 3, 3, 1, rng, mode=mode, context_tokens=20)
        assert r["mode"] == mode and r["errors"] == []
        assert r["engine_calls"] > 0 and r["mean_sessions_per_call"] >= 1.0
        if not r["aborted"]:                                          # the CPU oracle may fall behind real time on a loaded host:
            assert sum(r["stops"].values()) == 3 * 3                  # then the probe gives up (that is its job); otherwise every
            assert 0.0 < r["p50_latency_s"] <= r["p95_latency_s"] <= r["max_latency_s"]   # stream finished every measured tick
        else:
            assert not r["ok"]


def test_seam_summary_rule():
    import bench
    B, ticks, warm = 4, 6, 2
    lat = np.full((B, warm + ticks), 0.2)
    lag = np.full((B, warm + ticks), 0.05)
    stats = dict(prefix=[10], iters=[3], stops={"x": 1})
    ok = bench._seam_summary(B, "cohort", ticks, warm, lat, lag, False, [], stats, 1.0, {})
    assert ok["ok"] and abs(ok["p95_latency_s"] - 0.2) < 1e-9
    slow = lat.copy(); slow[:, -1] = 0.7                                  # p95 over the measured ticks crosses the chunk period
    assert not bench._seam_summary(B, "cohort", ticks, warm, slow, lag, False, [], stats, 1.0, {})["ok"]
    grow = lag.copy(); grow[:, -2:] = 0.4                                 # the start lag grows: a backlog is building
    assert not bench._seam_summary(B, "cohort", ticks, warm, lat, grow, False, [], stats, 1.0, {})["ok"]
    assert not bench._seam_summary(B, "cohort", ticks, warm, lat, lag, True, [], stats, 1.0, {})["ok"]
    assert not bench._seam_summary(B, "cohort", ticks, warm, lat, lag, False, ["boom"], stats, 1.0, {})["ok"]


def test_contract_pieces():
    import bench
    assert bench.CONFIGS[0] == "alignatt-large-v3" and len(bench.CONFIGS) == 6
    assert bench.CHUNK == 8000 and bench.WINDOW == 480000 and bench.CHUNK_S == 0.5
    peaks = bench.load_peaks()
    assert peaks["bf16_tflops"] > 100 and peaks["hbm_gbs"] > 1000
    assert isinstance(bench.reference_available(), bool)
