#!/usr/bin/env python
"""Stage the UNMODIFIED reference (QuentinFuxa/WhisperLiveKit, pure Python) under oracle/_ref/ so that its own
CPU backend -- the vendored torch Whisper behind ``AlignAtt`` (``--backend whisper``) -- can be timed on the GPU
box's host cores next to the B200 engine (``bench.py --impl reference`` and the ``cpu_baseline`` leg).

    python oracle/stage_reference.py          # build container only: needs /root/reference

Recipe: copy the package sources to a scratch directory (the reference tree is read-only and setuptools writes
build/ next to pyproject.toml), then the one offline install the task allows,
    pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target oracle/_ref <copy>
``--no-deps``: faster-whisper / torchaudio / librosa are not in the offline wheelhouse and the timed path
(whisper/model.py, whisper/audio.py, simul_whisper/*) needs only torch, numpy and tiktoken, which the image has.
oracle/_ref/ is git-ignored (no reference source enters the history) but travels to the GPU box with the snapshot.
Test / benchmark infrastructure only: nothing under whisperlivekit_b200/ imports it.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference"
TARGET = os.path.join(ROOT, "oracle", "_ref")
STAMP = os.path.join(TARGET, ".staged_from")


def staged() -> bool:
    return os.path.isdir(os.path.join(TARGET, "whisperlivekit", "whisper"))


def stage(force: bool = False) -> str:
    if staged() and not force:
        return TARGET
    if not os.path.isdir(os.path.join(REF_SRC, "whisperlivekit")):
        raise RuntimeError(f"{REF_SRC} not present: the reference can only be staged in the build container")
    tmp = tempfile.mkdtemp(prefix="wlk_ref_src_")
    try:
        for name in ("pyproject.toml", "README.md", "LICENSE", "MANIFEST.in"):
            p = os.path.join(REF_SRC, name)
            if os.path.exists(p):
                shutil.copy(p, os.path.join(tmp, name))
        shutil.copytree(os.path.join(REF_SRC, "whisperlivekit"), os.path.join(tmp, "whisperlivekit"),
                        ignore=shutil.ignore_patterns("__pycache__"))
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--find-links", "/opt/wheelhouse", "--target", TARGET, tmp]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or not staged():
            # dependency resolution / build backend trouble: the package is pure Python, a plain copy is equivalent
            os.makedirs(TARGET, exist_ok=True)
            shutil.copytree(os.path.join(tmp, "whisperlivekit"), os.path.join(TARGET, "whisperlivekit"), dirs_exist_ok=True)
            how = "copytree (pip failed: %s)" % (r.stderr.strip().splitlines()[-1] if r.stderr.strip() else r.returncode)
        else:
            how = "pip install --no-deps --target"
        # `soundfile` is imported at module scope by one backend file that the timed path never executes
        with open(os.path.join(TARGET, "soundfile_stub_note.txt"), "w") as f:
            f.write("bench.py installs a stub `soundfile` module in sys.modules before importing whisperlivekit\n")
        with open(STAMP, "w") as f:
            f.write(f"{REF_SRC} via {how}\n")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return TARGET


def import_staged_reference():
    """Put oracle/_ref first on sys.path (with a stub ``soundfile``) and import the reference package."""
    import importlib.machinery
    import types
    if not staged():
        raise RuntimeError("oracle/_ref is empty: run `python oracle/stage_reference.py` in the build container")
    if "soundfile" not in sys.modules:
        try:
            import soundfile  # noqa: F401
        except Exception:
            m = types.ModuleType("soundfile")
            m.__spec__ = importlib.machinery.ModuleSpec("soundfile", loader=None)
            m.read = m.write = m.info = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("soundfile stub"))
            sys.modules["soundfile"] = m
    if TARGET not in sys.path:
        sys.path.insert(0, TARGET)
    import whisperlivekit
    assert os.path.abspath(whisperlivekit.__file__).startswith(TARGET), whisperlivekit.__file__
    return whisperlivekit


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
    print(open(STAMP).read().strip())
