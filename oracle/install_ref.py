#!/usr/bin/env python
"""Recipe for ``oracle/_ref``: a runnable copy of the UNMODIFIED reference files the Speech2Text path imports.

TEST / MEASUREMENT INFRASTRUCTURE.  ``oracle/_ref`` is build output: git-ignored (the repository never holds reference sources) but
not gpurun-ignored, so it travels to the GPU box like a built ``.so``; there ``bench.py --impl reference`` (and ``cpu_baseline``)
time the reference's own ``espnet2.bin.asr_inference.Speech2Text`` on the host cores.  Nothing under ``espnet_b200/`` reads it.

How: a child process imports the reference from /root/reference (with the third-party shims of tests/golden/refshim.py), builds a
small Conformer + Transformer-decoder ``Speech2Text`` and decodes one utterance (joint CTC/attention and, for the LM scorer row,
with a TransformerLM); every module that came from the reference tree is then copied byte for byte to the same relative path under
``oracle/_ref``.  Run by ``__graft_entry__.build()`` whenever /root/reference is present; a no-op otherwise (the GPU box uses the
prebuilt copy)."""
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DEST = os.path.join(HERE, "_ref")
REFERENCE_ROOT = os.environ.get("ESPNET_REFERENCE_ROOT", "/root/reference")

_TRACE = r"""
import json, os, sys
sys.path.insert(0, os.path.join(%(root)r, "tests", "golden"))
sys.path.insert(0, os.path.join(%(root)r, "tests"))
import refshim, refbuild
refshim.install()
import torch
cfg = dict(d_model=32, heads=2, ff=48, enc_layers=1, dec_layers=1, vocab=12, kernel=7)
for kw in (dict(beam_size=3, ctc_weight=0.3, maxlenratio=-3.0), dict(beam_size=2, ctc_weight=1.0, maxlenratio=-2.0), dict(beam_size=2, ctc_weight=0.0, maxlenratio=-2.0)):
    s2t = refbuild.build_reference(cfg, seed=0, **kw)
    s2t(refbuild.waveform(0, 8000).numpy())
try:   # LM shallow fusion (espnet2/lm) and the streaming classes, for the "next" rows
    import espnet2.lm.transformer_lm, espnet2.lm.seq_rnn_lm, espnet2.tasks.lm  # noqa
    import espnet2.bin.asr_inference_streaming, espnet2.asr.encoder.contextual_block_conformer_encoder  # noqa
    import espnet2.legacy.nets.batch_beam_search_online  # noqa
except Exception as e:
    print("optional imports failed:", e, file=sys.stderr)
root = os.path.realpath(refshim.REFERENCE_ROOT)
files = sorted({os.path.realpath(m.__file__) for m in list(sys.modules.values())
                if getattr(m, "__file__", None) and os.path.realpath(m.__file__).startswith(root + os.sep)})
print("FILES=" + json.dumps(files))
"""


def install(verbose=True) -> bool:
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "espnet2")):
        if verbose:
            print(f"oracle/install_ref: {REFERENCE_ROOT} not present; keeping the existing oracle/_ref ({'found' if available() else 'absent'})")
        return available()
    env = dict(os.environ, ESPNET_REFERENCE_ROOT=REFERENCE_ROOT, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, "-c", _TRACE % dict(root=ROOT)], env=env, capture_output=True, text=True, timeout=900)
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("FILES=")]
    if out.returncode != 0 or not line:
        raise RuntimeError("tracing the reference import set failed:\n" + out.stderr[-2000:])
    files = json.loads(line[0][len("FILES="):])
    root = os.path.realpath(REFERENCE_ROOT)
    tmp = DEST + ".tmp"
    shutil.rmtree(tmp, ignore_errors=True)
    n = 0
    for f in files:
        rel = os.path.relpath(f, root)
        dst = os.path.join(tmp, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(f, dst)
        n += 1
    for extra in ("espnet2/version.txt", "version.txt", "LICENSE"):   # read at import time / licence of the copied files
        src = os.path.join(root, extra)
        if os.path.exists(src):
            os.makedirs(os.path.dirname(os.path.join(tmp, extra)), exist_ok=True)
            shutil.copyfile(src, os.path.join(tmp, extra))
    with open(os.path.join(tmp, "MANIFEST.json"), "w") as fh:
        json.dump({"source": root, "files": [os.path.relpath(f, root) for f in files]}, fh, indent=0)
    shutil.rmtree(DEST, ignore_errors=True)
    os.rename(tmp, DEST)
    if verbose:
        size = sum(os.path.getsize(os.path.join(dp, fn)) for dp, _, fns in os.walk(DEST) for fn in fns)
        print(f"oracle/install_ref: copied {n} reference files ({size / 1e6:.1f} MB) to {DEST}")
    return True


def available() -> bool:
    return os.path.isdir(os.path.join(DEST, "espnet2"))


def activate():
    """Point tests/golden/refshim at oracle/_ref (used on the GPU box, where /root/reference does not exist)."""
    if not available():
        raise RuntimeError("oracle/_ref is absent: run `python oracle/install_ref.py` where /root/reference is mounted")
    for p in (os.path.join(ROOT, "tests", "golden"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["ESPNET_REFERENCE_ROOT"] = DEST
    import refshim

    refshim.REFERENCE_ROOT = DEST
    refshim.install()
    return refshim


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
