"""Flat, versioned model blob: the one artefact that crosses the C ABI (`hoic_model_create`).

Layout (little endian):
    char     magic[8] = "HOICMDL1"
    uint32   version, n_entries
    entry[n] { char name[40]; uint32 dtype (0=f64, 1=i32, 2=bytes); uint32 ndim; uint32 shape[4];
               uint64 offset; uint64 nbytes }
    data     (each array 8-byte aligned, C order)
The C side (`hoic_b200/csrc/blob_reader.h`, also used by the oracle) looks arrays up by name, so fields can be
added without breaking older readers.  `dump_json` gives the diffable text form SURVEY.md section 7 asks for.
"""
from __future__ import annotations

import json
import struct

import numpy as np

MAGIC = b"HOICMDL1"
VERSION = 1
_ENTRY = struct.Struct("<40sII4IQQ")


def pack(arrays: dict) -> bytes:
    items = []
    for name, v in arrays.items():
        if isinstance(v, (bytes, bytearray)):
            items.append((name, 2, (len(v),), bytes(v)))
            continue
        a = np.ascontiguousarray(v)
        if a.dtype.kind == "f":
            a, dt = a.astype("<f8"), 0
        elif a.dtype.kind in "iub":
            a, dt = a.astype("<i4"), 1
        else:
            raise TypeError(f"{name}: dtype {a.dtype}")
        if a.ndim > 4:
            raise ValueError(name)
        items.append((name, dt, a.shape if a.ndim else (1,), a.tobytes()))
    head = len(MAGIC) + 8 + _ENTRY.size * len(items)
    off = (head + 7) // 8 * 8
    table, chunks = [], []
    for name, dt, shape, raw in items:
        nm = name.encode()
        if len(nm) > 39:
            raise ValueError(f"name too long: {name}")
        shp = list(shape) + [0] * (4 - len(shape))
        table.append(_ENTRY.pack(nm, dt, len(shape), *shp, off, len(raw)))
        pad = (-len(raw)) % 8
        chunks.append(raw + b"\0" * pad)
        off += len(raw) + pad
    out = MAGIC + struct.pack("<II", VERSION, len(items)) + b"".join(table)
    out += b"\0" * ((head + 7) // 8 * 8 - head)
    return out + b"".join(chunks)


def unpack(blob: bytes) -> dict:
    if blob[:8] != MAGIC:
        raise ValueError("not a HOIC model blob")
    version, n = struct.unpack_from("<II", blob, 8)
    if version != VERSION:
        raise ValueError(f"blob version {version}")
    out = {}
    for i in range(n):
        nm, dt, ndim, s0, s1, s2, s3, off, nb = _ENTRY.unpack_from(blob, 16 + i * _ENTRY.size)
        name = nm.rstrip(b"\0").decode()
        raw = blob[off:off + nb]
        if dt == 2:
            out[name] = bytes(raw)
        else:
            shape = (s0, s1, s2, s3)[:ndim]
            out[name] = np.frombuffer(raw, dtype="<f8" if dt == 0 else "<i4").reshape(shape).copy()
    return out


def pack_model(M: dict, cfg) -> bytes:
    """Compiled model (hoic_b200.mjcf) + env constants from the YAML config -> blob."""
    A = {}
    for k, v in M.items():
        if k.startswith("names_") or k == "env":
            continue
        if isinstance(v, (int, np.integer)):
            A[k] = np.array([v], dtype=np.int32)
        else:
            A[k] = v
    A["names_json"] = json.dumps({k: M[k] for k in M if k.startswith("names_")}).encode()
    e = M["env"]
    A["env_hand_body_idx"] = np.array(e["hand_body_idx"], dtype=np.int32)
    A["env_ids"] = np.array([e["obj_body_idx"], e["palm_body"], e["hand_geom_range"][0], e["hand_geom_range"][1],
                             e["obj_geom_range"][0], e["obj_geom_range"][1], cfg.sim_step, cfg.future_w_size,
                             int(bool(cfg.residual_force)), int(bool(cfg.surface_contact)),
                             int(bool(cfg.explain_force))], dtype=np.int32)
    nu = M["nu"]
    # ho_im4.py:103-107
    upper, lower = M["jnt_range"][:-1, 1], M["jnt_range"][:-1, 0]
    base = (upper + lower) / 2.0
    scale = upper - base
    scale[6:] *= 1.2
    A["env_base_pose"] = base
    A["env_ctrl_scale"] = scale
    A["env_jkp"] = np.asarray(cfg.jkp, dtype=np.float64)
    A["env_jkd"] = np.asarray(cfg.jkd, dtype=np.float64)
    A["env_torque_lim"] = np.asarray(cfg.torque_lim, dtype=np.float64)
    assert A["env_jkp"].shape == (nu,)
    A["env_scalars"] = np.array([e["hand_mass"], cfg.residual_force_scale, cfg.residual_torque_scale,
                                 cfg.pos_diff_thresh, cfg.rot_diff_thresh, cfg.jpos_diff_thresh,
                                 cfg.obj_pos_diff_thresh, cfg.obj_rot_diff_thresh])
    return pack(A)


def dump_json(blob: bytes) -> str:
    d = unpack(blob)
    out = {}
    for k, v in d.items():
        out[k] = json.loads(v.decode()) if isinstance(v, bytes) else v.tolist()
    return json.dumps(out, indent=1)
