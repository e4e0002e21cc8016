"""Compile the three release models from the reference tree into hoic_b200/data/*.hoicmodel (+ JSON dump).
Run here (where /root/reference exists); the blobs travel to the GPU box."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hoic_b200 import assets, blob  # noqa: E402

root = assets.reference_root()
assert root, "reference tree not found"
for cfg_id in assets.RELEASE_CONFIGS:
    obj = cfg_id.split("_")[0]
    m = assets.compile_from_reference(obj, root, cfg_id)
    name = obj if cfg_id == f"{obj}_future5_light_add_geom" else cfg_id
    p = os.path.join(assets.DATA_DIR, f"{name}.hoicmodel")
    open(p, "wb").write(m.blob)
    print(cfg_id, len(m.blob), "bytes ->", p, "flags", [int(x) for x in m.arrays["env_ids"][8:11]])
