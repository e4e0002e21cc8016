"""Compile the three release models from the reference tree into hoic_b200/data/*.hoicmodel (+ JSON dump).
Run here (where /root/reference exists); the blobs travel to the GPU box."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hoic_b200 import assets, blob  # noqa: E402

root = assets.reference_root()
assert root, "reference tree not found"
for obj in ("box", "bottle", "banana"):
    m = assets.compile_from_reference(obj, root)
    p = os.path.join(assets.DATA_DIR, f"{obj}.hoicmodel")
    open(p, "wb").write(m.blob)
    print(obj, len(m.blob), "bytes ->", p)
