"""Recipe for ``oracle/_ref/``: the UNMODIFIED reference modules of the hot path, staged so that
they travel to the GPU box -- TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT.

    python oracle/build_ref.py          # /root/reference/{tools,networks,models}.py + configs.yaml -> oracle/_ref/

``oracle/_ref/`` is git-ignored (reference sources never enter this repository's history) but not
gpurun-ignored, so the snapshot that goes to the B200 box carries it -- exactly like the built
``.so`` files.  The reference is pure Python: "building" it is staging the four files it needs for
``WorldModel._train`` / ``ImagBehavior._train`` (models.py imports networks and tools; the
hyper-parameters live in configs.yaml) and recording their SHA-256 in ``MANIFEST.json`` so a
reader can check that what was timed is byte-identical to /root/reference.  Nothing is edited;
the two CPU shims (MLP's ``device="cuda"`` default, PyYAML's string floats) are applied to the
imported modules in memory by ``ref_harness.py``.

Consumers: ``bench.py --impl reference`` / ``cpu_baseline`` (kind "reference"), the same-GPU eager
comparator, and ``tests/test_gpu_reference.py`` (CUDA path vs the reference itself at the real
configs).  The product package never imports anything from here.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("DV3_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ("tools.py", "networks.py", "models.py", "configs.yaml")


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def staged() -> bool:
    return all(os.path.isfile(os.path.join(DST, f)) for f in FILES)


def build(verbose=True) -> bool:
    """Stage the reference files.  Returns True when oracle/_ref is complete afterwards.  On a box
    without /root/reference (the GPU box) the files staged in the build container are used as
    they are."""
    if not os.path.isfile(os.path.join(SRC, "models.py")):
        if verbose:
            print(f"[build_ref] {SRC} not present; using staged copy: {staged()}", flush=True)
        return staged()
    os.makedirs(DST, exist_ok=True)
    manifest = {"source": SRC, "files": {}}
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
        manifest["files"][f] = _sha(os.path.join(DST, f))
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)
    if verbose:
        print(f"[build_ref] staged {len(FILES)} reference files into {DST}", flush=True)
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
