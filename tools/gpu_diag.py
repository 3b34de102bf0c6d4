#!/usr/bin/env python
"""GPU bring-up diagnostics: runs each check in its own subprocess (a trapping kernel poisons
its CUDA context, not the others') and appends JSON lines to gpurun_out/diag.jsonl.

    python tools/gpu_diag.py            # all checks
    python tools/gpu_diag.py gemm_tc    # one check, in-process
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "gpurun_out")


def emit(**kw):
    os.makedirs(OUT, exist_ok=True)
    line = json.dumps(kw, default=lambda o: float(o) if hasattr(o, "__float__") else str(o))
    print(line, flush=True)
    with open(os.path.join(OUT, "diag.jsonl"), "a") as f:
        f.write(line + "\n")


def _mk_engine(name="micro", precision="bf16", **kw):
    from golden_util import case_setup
    from whisperlivekit_b200.engine import WhisperEngine
    g, dims, sd, audio, heads = case_setup(name)
    return WhisperEngine(dims, sd, heads, precision=precision, **kw), g, audio


def check_gemm_simt():
    import torch
    from whisperlivekit_b200.dims import DIMS
    from whisperlivekit_b200.engine import WhisperEngine
    eng = WhisperEngine(DIMS["micro"], None, [(0, 0)], precision="fp32", max_sessions=1, max_batch=1)
    for (M, N, K) in [(128, 128, 64), (1500, 384, 384), (37, 51, 20), (16, 1280, 1280)]:
        A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / K ** 0.5
        b = torch.randn(N, device="cuda"); Cm = torch.empty(M, N, device="cuda")
        torch.cuda.synchronize()
        eng.op_gemm("simt", A.data_ptr(), 0, K, W.data_ptr(), 0, K, b.data_ptr(), Cm.data_ptr(), 0, N, M, N, K, False)
        eng.sync()
        ref = A.double() @ W.double().t() + b.double()
        emit(check="gemm_simt", shape=[M, N, K], max_err=(Cm.double() - ref).abs().max().item())


def check_gemm_tc():
    import torch
    from whisperlivekit_b200.dims import DIMS
    from whisperlivekit_b200.engine import WhisperEngine
    eng = WhisperEngine(DIMS["micro"], None, [(0, 0)], precision="bf16", max_sessions=1, max_batch=1)
    shapes = [(128, 256, 64), (128, 64, 64), (128, 128, 128), (256, 512, 256), (1500, 1280, 1280), (3000, 384, 240),
              (129, 264, 72), (4500, 5120, 1280)]
    for (M, N, K) in shapes:
        g = torch.Generator(device="cuda").manual_seed(M + N + K)
        A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
        W = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
        Cm = torch.full((M, N), float("nan"), device="cuda")
        torch.cuda.synchronize()
        eng.op_gemm("tcgen05", A.data_ptr(), 1, K, W.data_ptr(), 1, K, None, Cm.data_ptr(), 0, N, M, N, K, False)
        eng.sync()
        ref = A.float() @ W.float().t()
        err = (Cm - ref).abs()
        nan = torch.isnan(Cm).sum().item()
        err = torch.nan_to_num(err, nan=1e9)
        # error by 128-row block / 64-col block / to localise descriptor or layout mistakes
        mb = [err[i:i + 128].max().item() for i in range(0, M, 128)][:6]
        nb = [err[:, j:j + 64].max().item() for j in range(0, N, 64)][:8]
        rowpat = [err[r].max().item() for r in range(min(M, 16))]
        emit(check="gemm_tc", shape=[M, N, K], max_err=err.max().item(), ref_absmax=ref.abs().max().item(),
             nan=nan, err_by_mblk=mb, err_by_nblk64=nb, err_first_rows=rowpat,
             sample_out=Cm[0, :4].tolist(), sample_ref=ref[0, :4].tolist())
    # timing: encoder-shaped GEMMs at large-v3 sizes, 16 streams
    for (M, N, K) in [(24000, 1280, 1280), (24000, 3840, 1280), (24000, 5120, 1280), (24000, 1280, 5120)]:
        A = torch.randn(M, K, device="cuda").bfloat16(); W = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
        Cm = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        torch.cuda.synchronize()
        for _ in range(3):
            eng.op_gemm("tcgen05", A.data_ptr(), 1, K, W.data_ptr(), 1, K, None, Cm.data_ptr(), 1, N, M, N, K, False)
        eng.timer_record(0)
        iters = 10
        for _ in range(iters):
            eng.op_gemm("tcgen05", A.data_ptr(), 1, K, W.data_ptr(), 1, K, None, Cm.data_ptr(), 1, N, M, N, K, False)
        eng.timer_record(1)
        ms = eng.timer_elapsed_ms(0, 1) / iters
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(iters):
            torch.matmul(A, W.t())
        t1.record(); torch.cuda.synchronize()
        emit(check="gemm_tc_time", shape=[M, N, K], ms=ms, tflops=2.0 * M * N * K / ms / 1e9,
             cublas_ms=t0.elapsed_time(t1) / iters, cublas_tflops=2.0 * M * N * K / (t0.elapsed_time(t1) / iters) / 1e9)


def _forced(eng, g, audio):
    import numpy as np
    from golden_util import sampled_diff
    sid = eng.open_session()
    eng.append_audio(sid, audio)
    content = eng.encode([sid])[0]
    r = dict(content=[content, int(g["content_mel_len"])])
    r["mel"] = sampled_diff(g, "mel", eng.read_mel(sid))
    r["enc"] = sampled_diff(g, "enc", eng.read_encoder(sid))
    eng.decode([sid], [list(g["forced_prefix"])], sot_index=0)
    r["logits_prefill_last"] = sampled_diff(g, "logits_prefill_last", eng.read_logits(sid))
    r["logits_prefill_sot"] = sampled_diff(g, "logits_prefill_sot", eng.read_sot_logits(sid))
    am = []
    for i, t in enumerate(g["forced_steps"]):
        eng.decode([sid], [[int(t)]])
        lg = eng.read_logits(sid)
        am.append(int(np.argmax(lg)))
        if i in (0, 4):
            r[f"logits_step{i}"] = sampled_diff(g, f"logits_step{i}", lg)
    r["argmax_steps"] = [am, [int(x) for x in g["argmax_steps"]]]
    r["greedy"] = eng.greedy_and_align([sid])[0]
    attn = eng.read_align_attn(sid)
    r["align_attn"] = sampled_diff(g, "align_attn", attn)
    r["align_argmax"] = [[int(x) for x in attn.argmax(-1)][:12], [int(x) for x in g["align_argmax_rows"]][:12]]
    r["no_speech"] = eng.no_speech_prob([sid])[0]
    eng.close_session(sid)
    return r


def check_engine(precision, gemm_backend="auto"):
    for name in ("micro", "microml", "tiny"):
        try:
            eng, g, audio = _mk_engine(name, precision, max_sessions=2, max_batch=2, gemm_backend=gemm_backend)
            emit(check=f"engine_{precision}_{gemm_backend}", case=name, **_forced(eng, g, audio))
            eng.close()
        except Exception as ex:
            emit(check=f"engine_{precision}_{gemm_backend}", case=name, error=repr(ex), tb=traceback.format_exc()[-1500:])


def check_policy_fp32():
    from golden_util import run_policy
    for name in ("micro", "microml"):
        eng, g, audio = _mk_engine(name, "fp32", max_sessions=2, max_batch=2)
        for tag, nsp in (("pol", 1.01), ("poldef", 0.5)):
            tr = run_policy(eng, audio, nsp)
            same = {k: bool(list(tr[k]) == list(g[f"{tag}_{k}"])) for k in ("step_tokens", "step_frames", "new_tokens")}
            nt = len(g[f"{tag}_step_tokens"])
            first_bad = next((i for i in range(min(nt, len(tr["step_tokens"])))
                              if tr["step_tokens"][i] != g[f"{tag}_step_tokens"][i]
                              or tr["step_frames"][i] != g[f"{tag}_step_frames"][i]), None)
            emit(check="policy_fp32", case=name, tag=tag, same=same, n_steps=[len(tr["step_tokens"]), nt],
                 first_bad=first_bad)
        eng.close()


def check_large_timing():
    """large-v3 geometry, bf16 mode: per-kernel-class time of one encode + prefill + steps."""
    import numpy as np
    from whisperlivekit_b200.dims import DIMS, ALIGNMENT_HEADS
    from whisperlivekit_b200.engine import WhisperEngine
    from whisperlivekit_b200.weights import synthetic_state_dict, synthetic_audio
    dims = DIMS["large-v3"]
    t0 = time.time()
    sd = synthetic_state_dict(dims, seed=3)
    t1 = time.time()
    B = 8
    eng = WhisperEngine(dims, sd, ALIGNMENT_HEADS["large-v3"], precision="bf16", max_sessions=B, max_batch=B)
    emit(check="large_setup", gen_s=t1 - t0, load_s=time.time() - t1, memory=eng.memory())
    sids = [eng.open_session() for _ in range(B)]
    for i, s in enumerate(sids):
        eng.append_audio(s, synthetic_audio(29.5, seed=100 + i))
    prefix = list(eng.specials.sot_sequence_including_notimestamps()) + list(range(1000, 1044))
    for rep in range(3):
        eng.profile_reset(); eng.profile_enable(True)
        eng.timer_record(0)
        eng.encode(sids)
        eng.timer_record(1)
        eng.decode(sids, [prefix] * B)
        eng.timer_record(2)
        for k in range(4):
            eng.suppress(sids, eng.specials.alignatt_suppress_tokens())
            r = eng.greedy_and_align(sids)
            eng.decode(sids, [[r[i][0]] for i in range(B)])
        eng.timer_record(3)
        eng.sync()
        prof = {k: v for k, v in eng.profile_read().items() if v["launches"]}
        emit(check="large_timing", rep=rep, batch=B, encode_ms=eng.timer_elapsed_ms(0, 1),
             prefill_ms=eng.timer_elapsed_ms(1, 2), steps4_ms=eng.timer_elapsed_ms(2, 3), profile=prof)
        eng.profile_enable(False)
    eng.close()


def check_attn_tc():
    import torch
    from whisperlivekit_b200.dims import ModelDimensions
    from whisperlivekit_b200.engine import WhisperEngine
    for (d, H, B) in [(64, 1, 1), (128, 2, 1), (1280, 20, 2)]:
        eng = WhisperEngine(ModelDimensions(80, 1500, d, H, 1, 51864, 448, 64, 1, 1), None, [(0, 0)], precision="bf16",
                            max_sessions=1, max_batch=1)
        g = torch.Generator(device="cuda").manual_seed(d)
        qkv = (torch.randn(B * 1500, 3 * d, device="cuda", generator=g) * 0.8).bfloat16()
        out = torch.full((B * 1500, d), float("nan"), device="cuda", dtype=torch.bfloat16)
        out_s = torch.empty_like(out)
        torch.cuda.synchronize()
        eng.op_encoder_attention("tcgen05", qkv.data_ptr(), 1, B, out.data_ptr())
        eng.op_encoder_attention("simt", qkv.data_ptr(), 1, B, out_s.data_ptr())
        eng.sync()
        x = qkv.float().view(B, 1500, 3, H, 64)
        q, k, v = x[:, :, 0].transpose(1, 2), x[:, :, 1].transpose(1, 2), x[:, :, 2].transpose(1, 2)
        ref = (torch.softmax(q @ k.transpose(-1, -2), dim=-1) @ v).transpose(1, 2).reshape(B * 1500, d)
        err = torch.nan_to_num((out.float() - ref).abs(), nan=1e9)
        rows = [err[i:i + 128].max().item() for i in range(0, 1500, 128)]
        cols = [err[:, j:j + 16].max().item() for j in range(0, 64, 16)]
        emit(check="attn_tc", shape=[d, H, B], max_err=err.max().item(), nan=int(torch.isnan(out.float()).sum().item()),
             simt_err=(out_s.float() - ref).abs().max().item(), err_by_qtile=rows, err_by_dh16=cols,
             sample_out=out[0, :4].float().tolist(), sample_ref=ref[0, :4].tolist(),
             sample_out_r200=out[200, :4].float().tolist(), sample_ref_r200=ref[200, :4].tolist())
        if d == 1280:
            for name in ("tcgen05", "simt"):
                for _ in range(2):
                    eng.op_encoder_attention(name, qkv.data_ptr(), 1, B, out.data_ptr())
                eng.timer_record(0)
                for _ in range(5):
                    eng.op_encoder_attention(name, qkv.data_ptr(), 1, B, out.data_ptr())
                eng.timer_record(1)
                ms = eng.timer_elapsed_ms(0, 1) / 5
                emit(check="attn_time", backend=name, shape=[d, H, B], ms=ms,
                     tflops=4.0 * B * H * 1500 * 1500 * 64 / ms / 1e9)
        eng.close()


CHECKS = {
    "gemm_simt": check_gemm_simt,
    "gemm_tc": check_gemm_tc,
    "attn_tc": check_attn_tc,
    "engine_fp32": lambda: check_engine("fp32"),
    "policy_fp32": check_policy_fp32,
    "engine_bf16_simt": lambda: check_engine("bf16", "simt"),
    "engine_bf16_tc": lambda: check_engine("bf16", "tcgen05"),
    "large_timing": check_large_timing,
}


def main():
    if len(sys.argv) == 3 and sys.argv[1] == "--one" and sys.argv[2] in CHECKS:
        CHECKS[sys.argv[2]]()
        return
    names = [a for a in sys.argv[1:] if a in CHECKS] or list(CHECKS)
    for name in names:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", name], timeout=600,
                               capture_output=True, text=True)
            tail = (p.stdout[-3000:] if p.returncode else "") + p.stderr[-3000:]
            emit(check="_run", name=name, rc=p.returncode, seconds=time.time() - t0, tail=tail if p.returncode else "")
        except subprocess.TimeoutExpired:
            emit(check="_run", name=name, rc="timeout", seconds=time.time() - t0)


if __name__ == "__main__":
    main()
