#!/usr/bin/env python
"""bench.py -- streaming Whisper large-v3, 0.5 s chunks, B concurrent streams per GPU.

One "step" = one 0.5 s tick of the hot path for every stream of the job, AlignAtt-style:
    append 0.5 s of PCM (host -> device) and drop the oldest 0.5 s of the full 30 s window,
    log-mel -> 32-layer encoder over all 1500 positions -> cross-K/V for 32 decoder layers,
    decoder prefill of a PREFIX-token prompt, then STEPS_PER_CHUNK greedy iterations of
    (suppress -> argmax/logprob -> alignment-head reduction -> most attended frame -> 1-token decode).
That is the reference's per-chunk work in its parity (full re-encode) mode.

Metric: realtime streams = audio seconds processed per wall second = B * 0.5 / step_time, i.e.
how many concurrent streams the job sustains at RTF < 1; `rtf` is each stream's real-time factor
when B streams share the GPU (step_time / 0.5 s), the reference's definition
(scripts/run_scatter_benchmark.py:194-205).

    python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--streams B] [--model large-v3]
Multi-GPU: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...  (one rank per
GPU; sessions are sharded, NCCL is used once to broadcast the packed weights; weak scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

CHUNK_S = 0.5
CHUNK = 8000
WINDOW = 480000
PREFIX = 48
STEPS_PER_CHUNK = 8


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16_tflops=d.get("bf16_tflops_sustained") or d.get("bf16_tflops"), hbm_gbs=d.get("hbm_gbs"),
                    source="measured (MEASURED_PEAKS.json, sustained)")
    return dict(bf16_tflops=1400.0, hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, samples=len(self.rows))


# ------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference's CPU path on the host cores
# ------------------------------------------------------------------------------------------
def cpu_stream_chunk_seconds(dims, sd, heads, n_chunks=1, threads=None):
    """Time `n_chunks` stream-chunk steps of the same workload on the CPU oracle (fp32 torch CPU ops,
    the restatement of the reference's `--backend whisper` path).  Returns seconds per stream-chunk."""
    import torch
    from oracle import whisper_oracle as wo
    from whisperlivekit_b200.weights import synthetic_audio
    if threads:
        torch.set_num_threads(threads)
    eng = wo.OracleEngine(dims, sd, heads)
    sid = eng.open_session()
    eng.append_audio(sid, synthetic_audio(30.0, seed=1))
    prefix = list(eng.specials.sot_sequence_including_notimestamps()) + list(range(1000, 1000 + PREFIX - 4))
    sup = eng.specials.alignatt_suppress_tokens()
    t0 = time.perf_counter()
    for c in range(n_chunks):
        eng.drop_audio(sid, CHUNK)
        eng.append_audio(sid, synthetic_audio(CHUNK_S, seed=100 + c))
        eng.encode([sid])
        eng.decode([sid], [prefix])
        for _ in range(STEPS_PER_CHUNK):
            eng.suppress([sid], sup)
            tok, _, _ = eng.greedy_and_align([sid])[0]
            eng.decode([sid], [[tok]])
    return (time.perf_counter() - t0) / n_chunks, torch.get_num_threads()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=int(os.environ.get("WLK_BENCH_STREAMS", "96")), help="streams per GPU")
    ap.add_argument("--model", default="large-v3")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    from whisperlivekit_b200.dims import ALIGNMENT_HEADS, DIMS, default_alignment_heads
    from whisperlivekit_b200.weights import synthetic_audio, synthetic_state_dict
    dims = DIMS[args.model]
    heads = ALIGNMENT_HEADS.get(args.model) or default_alignment_heads(dims)
    workload = (f"whisper {args.model} AlignAtt greedy, {CHUNK_S}s chunks, 30 s rolling window fully re-encoded per chunk, "
                f"{PREFIX}-token prefill + {STEPS_PER_CHUNK} decode steps per chunk, {args.streams} streams/GPU")
    metric = "realtime_streams_large_v3_0.5s_chunks"
    unit = "concurrent real-time streams (audio-s per wall-s)"

    # ---------------------------------------------------------------- reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        sd = synthetic_state_dict(dims, seed=0)
        for _ in range(1):                                   # one untimed warm-up stream-chunk
            cpu_stream_chunk_seconds(dims, sd, heads, 1)
        per = []
        for _ in range(max(1, min(args.steps, 3))):          # bounded: each step = ONE stream-chunk
            s, threads = cpu_stream_chunk_seconds(dims, sd, heads, 1)
            per.append(s)
        sec = float(np.mean(per))
        value = CHUNK_S / sec
        line = dict(metric=metric, value=value, unit=unit, n_gpus=args.gpus, steps=len(per), warmup=1,
                    ms_per_step=sec * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                    data="synthetic", impl="reference",
                    config=dict(workload=workload.replace(f"{args.streams} streams/GPU", "1 stream on the host CPU"),
                                model=args.model, note="oracle port of the reference CPU path (vendored torch Whisper, fp32); "
                                "the reference itself cannot travel to the GPU box"),
                    cpu_baseline=dict(value=value, unit=unit, cores=threads, kind="port",
                                      sample=f"{len(per)} stream-chunk steps of the same workload, all host threads"),
                    e2e=dict(value=value, unit=unit, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return

    # ---------------------------------------------------------------- B200 arm
    import torch
    import torch.distributed as dist
    from whisperlivekit_b200.engine import WhisperEngine
    from whisperlivekit_b200.sharding import broadcast_blob
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B = args.streams
    eng = WhisperEngine(dims, None, heads, precision="bf16", device=local_rank, max_sessions=B, max_batch=B,
                        attn_backend=os.environ.get("WLK_ATTN", "auto"))
    if rank == 0:
        sd = synthetic_state_dict(dims, seed=0)
        eng.load_state_dict(sd)
    if world > 1:                                            # NCCL: weight broadcast at init, nothing else
        ptr, nbytes = eng.weight_blob()

        class _Blob:
            __cuda_array_interface__ = dict(shape=(nbytes,), typestr="|u1", data=(ptr, False), version=2)
        blob = torch.as_tensor(_Blob(), device=f"cuda:{local_rank}")
        broadcast_blob(blob, src=0)
        torch.cuda.synchronize()
        if rank != 0:
            eng.adopt_weights()
    sp = eng.specials
    sids = [eng.open_session() for _ in range(B)]
    rng = np.random.default_rng(1000 + rank)
    base = synthetic_audio(36.0, seed=7)
    for i, s in enumerate(sids):
        off = int(rng.integers(0, 16000 * 5))
        eng.append_audio(s, base[off: off + WINDOW] + 0.001 * rng.standard_normal(WINDOW).astype(np.float32))
    prefix = list(sp.sot_sequence_including_notimestamps()) + list(range(1000, 1000 + PREFIX - 4))
    sup = sp.alignatt_suppress_tokens()
    chunk_host = torch.empty(B, CHUNK, dtype=torch.float32).pin_memory()
    h2d = B * CHUNK * 4
    d2h = B * 16 * (STEPS_PER_CHUNK + 1)                     # StepResult per stream per sync

    host = dict(encode=0.0, prefill=0.0, step=0.0, n=0)      # host-side enqueue time of the engine calls (no sync inside)

    def step(with_io: bool, k: int):
        if with_io:
            chunk_host.copy_(torch.from_numpy(0.05 * rng.standard_normal((B, CHUNK)).astype(np.float32)))
            cn = chunk_host.numpy()
            for i, s in enumerate(sids):
                eng.drop_audio(s, CHUNK)
                eng.append_audio(s, cn[i])
        t0 = time.perf_counter()
        eng.encode(sids)
        t1 = time.perf_counter()
        eng.decode(sids, [prefix] * B)
        t2 = time.perf_counter()
        host["encode"] += t1 - t0; host["prefill"] += t2 - t1; host["n"] += 1
        eng.no_speech_prob(sids)
        for _ in range(STEPS_PER_CHUNK):
            eng.suppress(sids, sup)
            r = eng.greedy_and_align(sids)
            t3 = time.perf_counter()
            eng.decode(sids, [[t[0]] for t in r])
            host["step"] += time.perf_counter() - t3

    def timed(with_io: bool, steps: int, warmup: int, profile: bool):
        for k in range(warmup):
            step(with_io, k)
        eng.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if profile:
            eng.profile_reset(); eng.profile_enable(True)
        eng.timer_record(0)
        for k in range(steps):
            step(with_io, k)
        eng.timer_record(1)
        eng.sync()
        ms = eng.timer_elapsed_ms(0, 1)
        prof = eng.profile_read() if profile else None
        eng.profile_enable(False)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            dist.barrier()
        return ms, prof

    if os.environ.get("WLK_NCU"):
        # profiler capture mode: warm up, then exactly one step between cudaProfilerStart/Stop
        # (ncu --profile-from-start off ...).  Numbers printed under a profiler are never bench values.
        for k in range(args.warmup):
            step(False, k)
        eng.sync()
        torch.cuda.profiler.start()
        step(False, 0)
        eng.sync()
        torch.cuda.profiler.stop()
        print(json.dumps(dict(ncu_capture=True, streams=B)))
        return

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_dev, _ = timed(False, args.steps, args.warmup, False)
    clocks = sampler.summary() if sampler else None
    ms_e2e, _ = timed(True, args.steps, max(1, args.warmup // 3), False)
    # same steps once more with a CUDA-event pair around every kernel class launch (engine stream):
    # per-class device time for the roofline; kept out of `value` because ~10^4 event records per step
    # cost a few percent of host-side launch throughput.
    ms_prof, prof = timed(False, args.steps, 0, True)

    if rank == 0:
        peaks = load_peaks()
        total_streams = B * world
        value = total_streams * CHUNK_S * args.steps / (ms_dev / 1e3)
        e2e_value = total_streams * CHUNK_S * args.steps / (ms_e2e / 1e3)
        g = prof["gemm_enc"]
        traffic = None                                     # DRAM bytes per launch of the dominant kernel, from the committed ncu capture
        tpath = os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")
        if os.path.exists(tpath) and B == 96 and args.model == "large-v3":
            traffic = json.load(open(tpath))["mean_dram_bytes_per_launch"]
        ach = g["flops"] / (g["ms"] / 1e3) / 1e12 if g["ms"] else 0.0
        mult = dict(mel=2, align=3)
        launches = int(sum(v["launches"] * mult.get(k, 1) for k, v in prof.items()))
        classes = {k: dict(ms_per_step=v["ms"] / args.steps, launches_per_step=v["launches"] / args.steps,
                           tflops=(v["flops"] / (v["ms"] / 1e3) / 1e12) if v["ms"] and v["flops"] else None)
                   for k, v in prof.items() if v["launches"]}
        line = dict(
            metric=metric, value=value, unit=unit, n_gpus=world, steps=args.steps, warmup=args.warmup,
            ms_per_step=ms_dev / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
            dtype="bf16", data="synthetic (seeded random weights at true large-v3 dims, synthetic speech-like audio)",
            config=dict(workload=workload, model=args.model, streams_per_gpu=B, parallelism=f"sessions sharded x{world}",
                        chunk_s=CHUNK_S, l2="working set (3.4 GB weights + per-stream KV) exceeds the 126 MB L2",
                        rtf_per_stream=(ms_dev / args.steps / 1e3) / CHUNK_S),
            e2e=dict(value=e2e_value, unit=unit, h2d_bytes_per_step=h2d * world, d2h_bytes_per_step=d2h * world,
                     ms_per_step=ms_e2e / args.steps),
            gpu_launches=launches,
            clocks=clocks,
            roofline=dict(bound="tensor", kernel="gemm_tc2_kernel (cta_group::2 pair GEMM; encoder GEMMs, class gemm_enc)", achieved=ach,
                          peak=peaks["bf16_tflops"], unit="TFLOP/s", frac=ach / peaks["bf16_tflops"], traffic=traffic,
                          traffic_note="dram__bytes_read+write per launch, mean of one encoder layer's 4 GEMMs, ncu --set full at 96 "
                                       "streams (profiles/r01_gemm_traffic.json); algorithmic 2.04 GB",
                          peak_source=peaks["source"],
                          flops_per_launch=g["flops"] / max(1, g["launches"]), ms_per_launch=g["ms"] / max(1, g["launches"])),
            kernel_classes=classes,
            profiled_ms_per_step=ms_prof / args.steps,
            host_enqueue_ms=dict(encode_call=1e3 * host["encode"] / host["n"], prefill_call=1e3 * host["prefill"] / host["n"],
                                 decode_step_call=1e3 * host["step"] / host["n"] / STEPS_PER_CHUNK,
                                 note="host time inside the (asynchronous) engine calls, averaged over all passes"),
        )
        if not args.no_cpu_baseline:
            sd_cpu = sd if world == 1 or rank == 0 else None
            sec, threads = cpu_stream_chunk_seconds(dims, sd_cpu, heads, 2)
            line["cpu_baseline"] = dict(value=CHUNK_S / sec, unit=unit, cores=threads, kind="port",
                                        sample="2 stream-chunk steps of the same workload (1 stream), oracle port of the "
                                               "reference CPU path, all host threads", seconds_per_stream_chunk=sec)
        print(json.dumps(line))
    for s in sids:
        eng.close_session(s)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
