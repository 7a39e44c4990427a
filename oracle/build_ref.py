"""Recipe for ``oracle/_ref/``: the UNMODIFIED reference hot-path modules, for the reference arm.

TEST / MEASUREMENT INFRASTRUCTURE ONLY — nothing under ``windgnn_b200/`` imports this.

The reference's hot path is two pure-Python files (no build step, no native code):

    /root/reference/src/step5_gcn_layer_model.py          GraphConvLayer   (23 lines)
    /root/reference/src/step6_gcn_gru_combined_model.py   GCN_GRU          (27 lines)

This script copies them, byte for byte, into ``oracle/_ref/`` together with a manifest of their
sha256.  ``oracle/_ref/`` is git-ignored (reference sources never enter this repository's history)
but NOT gpurun-ignored, so the copies travel to the GPU box, where ``/root/reference`` does not
exist, and ``bench.py --impl reference`` / the ``cpu_baseline`` and ``gpu_eager`` legs can run the
reference's own classes (``kind: "reference"``).  Without ``oracle/_ref/`` those legs fall back to the
oracle's torch port and say ``kind: "port"``.

    python oracle/build_ref.py        (also called by __graft_entry__.build() when /root/reference exists)
"""

from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("WINDGNN_REFERENCE", "/root/reference")
DEST = os.path.join(HERE, "_ref")
FILES = ("step5_gcn_layer_model.py", "step6_gcn_gru_combined_model.py")


def build_ref(verbose: bool = False) -> bool:
    """Copy the reference modules into oracle/_ref/.  Returns False when the reference is absent."""
    src_dir = os.path.join(REF_SRC, "src")
    if not all(os.path.exists(os.path.join(src_dir, f)) for f in FILES):
        return False
    os.makedirs(DEST, exist_ok=True)
    manifest = {"source": src_dir, "files": {}}
    for f in FILES:
        shutil.copyfile(os.path.join(src_dir, f), os.path.join(DEST, f))
        with open(os.path.join(DEST, f), "rb") as fh:
            manifest["files"][f] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)
    if verbose:
        print(json.dumps(manifest, indent=1))
    return True


if __name__ == "__main__":
    ok = build_ref(verbose=True)
    print("oracle/_ref built" if ok else f"reference not found under {REF_SRC}: oracle/_ref not built")
    sys.exit(0)
