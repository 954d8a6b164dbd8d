"""Builds oracle/_ref/: the reference's own hot-path modules, byte-compiled from where they lie under /root/reference.

TEST / BASELINE INFRASTRUCTURE.  The reference is pure Python, so "building" it means compiling the modules the path
imports (found by importing the path once through oracle/ref_shims.py and listing what was loaded from the reference
tree) to sourceless bytecode files (`*.refbin`, CPython .pyc format).  Only compiled outputs are written, only into oracle/_ref/ (git-ignored, shipped to
the GPU box by gpurun like our own .so): no reference source enters the repository.  `bench.py --impl reference` and
bench.py's cpu_baseline leg import it there (kind "reference"); without it they fall back to the oracle port.

  python -m oracle.build_ref        # needs /root/reference; a no-op (exit 0) when it is absent
"""
import os
import py_compile
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SRC_ROOT = "/root/reference"


def build(verbose=False):
    if not os.path.isdir(os.path.join(SRC_ROOT, "jyutvoice")):
        return None
    os.environ["JYUTVOICE_REFERENCE"] = SRC_ROOT
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import ref_shims
    ref_shims.REF_ROOT = SRC_ROOT
    ref_shims.build_reference_cfm()
    ref_shims.build_reference_hift()
    ref_shims.build_reference_tts()  # text encoder, duration predictor, JyutVoiceTTS.synthesise
    ref_shims.build_reference_flow_encoder()  # UpsampleConformerEncoder (prompt_h)
    files = set()
    for name, mod in list(sys.modules.items()):
        f = getattr(mod, "__file__", None)
        if f and os.path.abspath(f).startswith(SRC_ROOT + os.sep) and f.endswith(".py"):
            files.add(os.path.abspath(f))
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    for f in sorted(files):
        rel = os.path.relpath(f, SRC_ROOT)
        # compiled module next to where module.py would be; the suffix is not ".pyc" because snapshot tools (gpurun's
        # among them) drop *.pyc files: oracle/ref_shims.py installs a finder that loads ".refbin" bytecode
        dst = os.path.join(OUT, rel[:-3] + ".refbin")
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        py_compile.compile(f, cfile=dst, dfile=rel, doraise=True)
        if verbose:
            print("compiled", rel)
    # packages on the way that were skipped by the shims (jyutvoice/utils/__init__ pulls hydra / lightning): the shim
    # registers a bare package object for them, which needs the directory to exist
    os.makedirs(os.path.join(OUT, "jyutvoice", "utils"), exist_ok=True)
    with open(os.path.join(OUT, "BUILT_FROM"), "w") as fh:
        fh.write(f"{SRC_ROOT}: {len(files)} modules byte-compiled by oracle/build_ref.py (python {sys.version.split()[0]})\n")
    return OUT


if __name__ == "__main__":
    out = build(verbose=True)
    print(out or "reference tree absent: nothing built")
