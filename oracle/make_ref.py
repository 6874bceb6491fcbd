"""Recipe for ``oracle/_ref/``: an UNMODIFIED copy of the reference's hot-path modules, taken from where they lie
under /root/reference (build container only; the GPU box receives the prebuilt directory with the gpurun snapshot).

    python oracle/make_ref.py            # or __graft_entry__.build()

``oracle/_ref/`` is git-ignored (no reference source enters the repository's history) but NOT gpurun-ignored.  It is
test / baseline infrastructure: ``bench.py --impl reference`` and the ``cpu_baseline`` / ``gpu_eager_baseline`` legs time
the reference's OWN ``MVAE`` / ``elbo_loss`` through ``oracle/ref_harness.py`` (``kind: "reference"``); when the directory is
absent they fall back to the oracle port (``kind: "port"``).
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MVAE_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = [(d, f) for d in ("mnist", "fashionmnist", "celeba", "celeba19") for f in ("model.py", "train.py")]


def make(verbose: bool = True) -> bool:
    """Copy the eight files byte for byte; returns False (and leaves any existing copy alone) if the reference tree is
    not available here."""
    if not os.path.isdir(REF):
        if verbose:
            print(f"make_ref: {REF} not present; keeping {'the existing' if os.path.isdir(DST) else 'no'} oracle/_ref")
        return os.path.isdir(DST)
    manifest = {}
    for d, f in FILES:
        src = os.path.join(REF, d, f)
        os.makedirs(os.path.join(DST, d), exist_ok=True)
        shutil.copyfile(src, os.path.join(DST, d, f))
        manifest[f"{d}/{f}"] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": REF, "sha256": manifest}, fh, indent=1)
    if verbose:
        print(f"make_ref: copied {len(FILES)} files into {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
