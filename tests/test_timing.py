"""Word-timestamp math (LocalAgreement path): oracle vs fixtures recorded from the reference's numba
dtw_cpu / torch median_filter (CPU), and the native CUDA kernels vs the same fixtures (GPU, bit-exact)."""
import os

import numpy as np
import pytest

from oracle import timing_oracle as to

G = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "timing.npz")))


def cases():
    for i in range(int(G["n_cases"])):
        yield G[f"x{i}"], G[f"med{i}"], G[f"text{i}"], G[f"time{i}"]


def test_oracle_median_and_dtw_match_reference():
    for x, med, ti, fi in cases():
        m = to.median_filter(x, 7)
        assert np.array_equal(m, med)
        t, f = to.dtw(-m)
        assert np.array_equal(t, ti) and np.array_equal(f, fi)
    t, f = to.dtw(G["xq"])
    assert np.array_equal(t, G["textq"]) and np.array_equal(f, G["timeq"])


@pytest.mark.gpu
def test_cuda_median_and_dtw_match_reference():
    import torch
    from whisperlivekit_b200.dims import DIMS
    from whisperlivekit_b200.engine import WhisperEngine
    eng = WhisperEngine(DIMS["micro"], None, [(0, 0)], precision="fp32", max_sessions=1, max_batch=1)
    for x, med, ti, fi in cases():
        xd = torch.from_numpy(x).cuda()
        od = torch.empty_like(xd)
        torch.cuda.synchronize()
        eng.op_median_filter(xd.data_ptr(), od.data_ptr(), x.shape[0], x.shape[1], 7)
        eng.sync()
        assert np.array_equal(od.cpu().numpy(), med)              # selection, not arithmetic: bit-exact
        neg = (-od).contiguous()
        torch.cuda.synchronize()
        t, f = eng.op_dtw(neg.data_ptr(), x.shape[0], x.shape[1])
        assert np.array_equal(t, ti) and np.array_equal(f, fi)
        # host-array entry points used by the LocalAgreement shim (localagreement.install_native_timing)
        assert np.array_equal(eng.median_filter_host(x, 7), med)
        t, f = eng.dtw_host(-med)
        assert np.array_equal(t, ti) and np.array_equal(f, fi)
    xq = torch.from_numpy(G["xq"]).cuda()
    torch.cuda.synchronize()
    t, f = eng.op_dtw(xq.data_ptr(), *G["xq"].shape)
    assert np.array_equal(t, G["textq"]) and np.array_equal(f, G["timeq"])
    # full-size problem (448 tokens x 1500 frames): path invariants
    big = torch.rand(448, 1500, device="cuda")
    torch.cuda.synchronize()
    t, f = eng.op_dtw(big.data_ptr(), 448, 1500)
    assert t[0] == 0 and f[0] == 0 and t[-1] == 447 and f[-1] == 1499
    assert np.all(np.diff(t) >= 0) and np.all(np.diff(f) >= 0) and np.all((np.diff(t) + np.diff(f)) >= 1)
    eng.close()
