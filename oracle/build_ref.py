"""Recipe that puts the UNMODIFIED reference's Python sources for this path under oracle/_ref/ (git-ignored, NOT
gpurun-ignored) so that the GPU box - which has no /root/reference - can run the reference itself:

    python -m oracle.build_ref          (also run by __graft_entry__.build() when /root/reference is present)

The reference is plain Python: "building" it is a verbatim file copy of the modules the path imports (listed below),
byte for byte, plus a MANIFEST with their SHA-256.  Nothing under oracle/_ref/ is ever committed (see .gitignore) and
nothing under unirec_b200/ imports it: it is test / baseline infrastructure, driven through oracle/reference_shim.py by
tests/, bench.py's `--impl reference` arm and bench.py's same-GPU eager comparator.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "oracle", "_ref")

# what `models.qformer_utils`, `models.user_sequence_encoder`, `training.user_qformer_training` and `models.mwne` import
FILES = [
    "models/__init__.py",
    "models/qformer.py",                    # BertModel and friends (the backbone)
    "models/qformer_model.py",              # QFormerForItemRepresentation (copy without the data-set helpers)
    "models/qformer_utils.py",              # QFormerForItemRepresentation + cache helpers
    "models/item_encoder_pure_value.py",    # imported by qformer_utils (never instantiated on this path)
    "models/mwne.py",                       # TimestampEncoder / GeoCoordinateEncoder / MWNE
    "models/user_sequence_encoder.py",      # PositionalEncoding, UserSequenceEncoder
    "training/user_qformer_training.py",    # UserQFormer
]


def build(verbose: bool = True) -> bool:
    if not os.path.isdir(SRC):
        if verbose:
            print(f"oracle/_ref: {SRC} not present (GPU box): keeping what the snapshot brought", file=sys.stderr)
        return os.path.exists(os.path.join(DST, "MANIFEST.json"))
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    init = os.path.join(DST, "training", "__init__.py")
    if not os.path.exists(init):
        open(init, "w").close()             # the reference's training/ is a script directory (no package file)
    json.dump({"source": SRC, "files": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    if verbose:
        print(f"oracle/_ref: copied {len(FILES)} unmodified reference files", file=sys.stderr)
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
