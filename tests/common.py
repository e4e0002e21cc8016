"""Shared helpers for the test-suite: model blobs, synthetic motions, oracle construction."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from hoic_b200 import blob as hblob  # noqa: E402
from hoic_b200 import motion as hmotion  # noqa: E402
from hoic_b200.assets import load_model  # noqa: E402

OBJECTS = ("box", "bottle", "banana")
_CACHE = {}


def model(obj, cfg_id=None):
    if (obj, cfg_id) not in _CACHE:
        _CACHE[(obj, cfg_id)] = load_model(obj, cfg_id=cfg_id)
    return _CACHE[(obj, cfg_id)]


def oracle_env(obj, seed=1, start_idx=0, cfg_id=None):
    from oracle import OracleEnv
    mdl = model(obj, cfg_id)
    env = OracleEnv(mdl.blob)
    ex = mdl.motion(seed)
    env.set_motion(hmotion.pack_frames(hmotion.load_seq(ex, start_idx)))
    env.set_reward_weights(mdl.cfg.reward_vector())
    return env
