"""CPU, world_size 2 over gloo: the N>1 host logic (stream sharding, weight-blob broadcast,
max-over-ranks timing).  The data path itself has no collective."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from whisperlivekit_b200.sharding import StickyPlacement, broadcast_blob, max_over_ranks, shard_streams


def test_shard_streams_balanced_and_disjoint():
    for n, w in [(512, 8), (7, 2), (3, 4), (64, 1)]:
        sh = shard_streams(n, w)
        assert sorted(x for s in sh for x in s) == list(range(n))
        assert max(map(len, sh)) - min(map(len, sh)) <= 1


def test_sticky_placement():
    p = StickyPlacement(4)
    ranks = [p.open(i) for i in range(10)]
    assert sorted(p.load) == [2, 2, 3, 3]
    p.close(0)
    assert p.open(100) == ranks[0]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    blob = torch.zeros(1 << 16, dtype=torch.uint8)
    if rank == 0:
        blob.copy_(torch.from_numpy(np.random.default_rng(0).integers(0, 255, 1 << 16, dtype=np.uint8)))
    broadcast_blob(blob, src=0)
    mine = shard_streams(9, world)[rank]
    t = max_over_ranks(1.0 + rank)
    q.put((rank, int(blob.sum()), mine, t))
    dist.destroy_process_group()


def test_two_rank_broadcast_and_sharding():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(30)
    assert res[0][1] == res[1][1] != 0
    assert res[0][2] == [0, 1, 2, 3, 4] and res[1][2] == [5, 6, 7, 8]
    assert res[0][3] == res[1][3] == 2.0


def _qwen_worker(rank, world, port, q):
    """Each rank serves its shard of the streams with its own tower (weights broadcast from rank 0 as a flat blob),
    no exchange on the data path; rank-local results go back to the parent for comparison."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.make_golden_qwen import mel_stream
    from oracle.qwen_oracle import QwenTowerOracle
    from whisperlivekit_b200.qwen_dims import QWEN_DIMS, synthetic_tower_state_dict
    dims = QWEN_DIMS["qnano"]
    ref_sd = synthetic_tower_state_dict(dims, seed=4)
    names = sorted(ref_sd)
    sizes = [ref_sd[k].size for k in names]
    blob = torch.zeros(sum(sizes), dtype=torch.float32)
    if rank == 0:
        blob.copy_(torch.from_numpy(np.concatenate([ref_sd[k].reshape(-1) for k in names])))
    broadcast_blob(blob, src=0)                                   # the only collective: weights at init
    sd, off = {}, 0
    for k, n in zip(names, sizes):
        sd[k] = blob[off: off + n].numpy().reshape(ref_sd[k].shape).copy()
        off += n
    eng = QwenTowerOracle(dims, sd)
    mine = shard_streams(5, world)[rank]
    out = {}
    for stream in mine:
        sid = eng.open_session()
        mels = mel_stream(450, dims.n_mels, seed=100 + stream)
        rows = [eng.forward_chunk([sid], [mels[a: a + 150]])[0] for a in range(0, 450, 150)]
        out[stream] = np.concatenate(rows, axis=0)
    q.put((rank, {k: v.tolist() for k, v in out.items()}))
    dist.destroy_process_group()


def test_two_rank_qwen_streams_are_sharded_without_data_exchange():
    from oracle.make_golden_qwen import mel_stream
    from oracle.qwen_oracle import QwenTowerOracle
    from whisperlivekit_b200.qwen_dims import QWEN_DIMS, synthetic_tower_state_dict
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_qwen_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(x for _, d in sorted(q.get(timeout=300) for _ in procs) for x in d.items())
    for p in procs:
        p.join(60)
    assert sorted(res) == [0, 1, 2, 3, 4]
    dims = QWEN_DIMS["qnano"]
    eng = QwenTowerOracle(dims, synthetic_tower_state_dict(dims, seed=4))
    for stream in range(5):
        sid = eng.open_session()
        mels = mel_stream(450, dims.n_mels, seed=100 + stream)
        want = np.concatenate([eng.forward_chunk([sid], [mels[a: a + 150]])[0] for a in range(0, 450, 150)], axis=0)
        np.testing.assert_allclose(np.asarray(res[stream], np.float32), want, atol=1e-6)


def _diar_worker(rank, world, port, q):
    """Config 4's diarization leg under N > 1: every rank serves its shard of the streams with its own Sortformer (weights
    broadcast once as a flat blob), per-stream speaker caches never leave their rank, no exchange on the data path."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.sortformer_oracle import OracleDiarizer, SortformerOracle
    from whisperlivekit_b200.sortformer_dims import SORTFORMER_DIMS, synthetic_sortformer_state_dict, synthetic_two_speaker_audio
    dims = SORTFORMER_DIMS["micro"]
    ref_sd = synthetic_sortformer_state_dict(dims, seed=6)
    names = sorted(ref_sd)
    sizes = [ref_sd[k].size for k in names]
    blob = torch.zeros(sum(sizes), dtype=torch.float32)
    if rank == 0:
        blob.copy_(torch.from_numpy(np.concatenate([ref_sd[k].reshape(-1) for k in names])))
    broadcast_blob(blob, src=0)
    sd, off = {}, 0
    for k, n in zip(names, sizes):
        sd[k] = blob[off: off + n].numpy().reshape(ref_sd[k].shape).copy()
        off += n
    model = SortformerOracle(dims, sd)
    out = {}
    for stream in shard_streams(3, world)[rank]:
        d = OracleDiarizer(model)
        audio = synthetic_two_speaker_audio(4.0, seed=50 + stream)
        for k in range(4):
            d.step(audio[k * 16000:(k + 1) * 16000])
        out[stream] = (d.total_preds.numpy().tolist(), d.st["spkcache_len"], d.st["fifo_len"])
    q.put((rank, out))
    dist.destroy_process_group()


def test_two_rank_diarization_streams_are_sharded_without_data_exchange():
    from oracle.sortformer_oracle import OracleDiarizer, SortformerOracle
    from whisperlivekit_b200.sortformer_dims import SORTFORMER_DIMS, synthetic_sortformer_state_dict, synthetic_two_speaker_audio
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_diar_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(x for _, d in sorted(q.get(timeout=300) for _ in procs) for x in d.items())
    for p in procs:
        p.join(60)
    assert sorted(res) == [0, 1, 2]
    dims = SORTFORMER_DIMS["micro"]
    model = SortformerOracle(dims, synthetic_sortformer_state_dict(dims, seed=6))
    for stream in range(3):
        d = OracleDiarizer(model)
        audio = synthetic_two_speaker_audio(4.0, seed=50 + stream)
        for k in range(4):
            d.step(audio[k * 16000:(k + 1) * 16000])
        preds, sl, fl = res[stream]
        np.testing.assert_allclose(np.asarray(preds, np.float32), d.total_preds.numpy(), atol=1e-6)
        assert (sl, fl) == (d.st["spkcache_len"], d.st["fifo_len"])
