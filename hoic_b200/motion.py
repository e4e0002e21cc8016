"""Expert-motion format and derivations, plus the synthetic reference motions used while the
released pkl datasets are unavailable offline.

* `derive_expert` restates `DatasetSingleDepth.preprocess_seq` for the keys the env reads
  (uhc/data_loaders/dataset_singledepth.py:78-142): clip to joint range, finite-difference velocities
  (`compute_vel_from_seq` :152-185), body poses by FK (`compute_body_pos_quat_from_seq` :222-237).
* `load_seq` restates the slicing of `load_seq` (:239-288) as used by the sampler (full_seq=True).
* `pack_frames` flattens an expert dict to the [T,212] float layout that crosses the C ABI (`hoic_set_motions`).
* `synth_motion` is the deterministic approach-grasp-lift script of SURVEY.md section 8d.
"""
from __future__ import annotations

import numpy as np

from . import mjcf

FRAME_DIM = 212


# ----------------------------------------------------------------------------- rotations (numpy, batched)


def quat_to_mat_batch(q):
    q = q / np.linalg.norm(q, axis=-1, keepdims=True)
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], axis=-1)
    return R.reshape(q.shape[:-1] + (3, 3))


def matrix_to_axis_angle(R):
    """pytorch3d-style conversion used by the reference (uhc/utils/transforms.py:99-158, 414-490), numpy float64."""
    R = np.asarray(R, dtype=np.float64)
    m = R.reshape(-1, 9)
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = [m[:, i] for i in range(9)]
    qa = np.stack([1 + m00 + m11 + m22, 1 + m00 - m11 - m22, 1 - m00 + m11 - m22, 1 - m00 - m11 + m22], -1)
    qa = np.sqrt(np.maximum(qa, 0))
    cand = np.stack([
        np.stack([qa[:, 0] ** 2, m21 - m12, m02 - m20, m10 - m01], -1),
        np.stack([m21 - m12, qa[:, 1] ** 2, m10 + m01, m02 + m20], -1),
        np.stack([m02 - m20, m10 + m01, qa[:, 2] ** 2, m12 + m21], -1),
        np.stack([m10 - m01, m20 + m02, m21 + m12, qa[:, 3] ** 2], -1)], -2)
    cand = cand / (2.0 * np.maximum(qa, 0.1))[:, :, None]
    q = cand[np.arange(m.shape[0]), qa.argmax(-1)]
    nrm = np.linalg.norm(q[:, 1:], axis=-1)
    half = np.arctan2(nrm, q[:, 0])
    ang = 2 * half
    small = np.abs(ang) < 1e-6
    s = np.where(small, 0.5 - ang * ang / 48, np.sin(half) / np.where(small, 1, ang))
    return (q[:, 1:] / s[:, None]).reshape(R.shape[:-2] + (3,))


# ----------------------------------------------------------------------------- derivations


def _wrap_pi(x):
    x = x.copy()
    while np.any(x > np.pi):
        x[x > np.pi] -= 2 * np.pi
    while np.any(x < -np.pi):
        x[x < -np.pi] += 2 * np.pi
    return x


def compute_vel_from_seq(hand_dof_seq, obj_pose_seq, motion_freq):
    """dataset_singledepth.py:152-185."""
    hv = np.zeros_like(hand_dof_seq)
    hv[1:] = hand_dof_seq[1:] - hand_dof_seq[:-1]
    hv[0] = hv[1]
    hv[:, 3:6] = _wrap_pi(hv[:, 3:6])
    hv = hv * motion_freq
    ov = np.zeros_like(obj_pose_seq[:, 0:3])
    ov[1:] = obj_pose_seq[1:, 0:3] - obj_pose_seq[:-1, 0:3]
    ov[0] = ov[1]
    ov = ov * motion_freq
    Rm = quat_to_mat_batch(obj_pose_seq[:, 3:])
    rel = np.matmul(Rm[1:], np.transpose(Rm[:-1], (0, 2, 1)))
    oav = np.zeros_like(ov)
    oav[1:] = matrix_to_axis_angle(rel) * motion_freq
    oav[0] = oav[1]
    return hv, ov, oav


def compute_body_pos_quat_from_seq(M, hand_dof_seq):
    """dataset_singledepth.py:222-237: FK of the hand dofs with the object left at qpos0."""
    idx = M["env"]["hand_body_idx"]
    T = hand_dof_seq.shape[0]
    bp, bq = np.zeros((T, len(idx), 3)), np.zeros((T, len(idx), 4))
    q = M["qpos0"].copy()
    nh = hand_dof_seq.shape[1]
    for t in range(T):
        q[:nh] = hand_dof_seq[t]
        kin = mjcf.fk(M, q)
        bp[t], bq[t] = kin["xpos"][idx], kin["xquat"][idx]
    return bp, bq


def derive_expert(M, hand_pose_seq, obj_pose_seq, motion_freq=30):
    """Raw (hand dofs, object pose) sequence -> expert dict with the reference's keys (preprocess_seq :78-142)."""
    lo, hi = M["jnt_range"][:-1, 0], M["jnt_range"][:-1, 1]
    hand = np.clip(np.asarray(hand_pose_seq, dtype=np.float64), lo, hi)
    obj = np.asarray(obj_pose_seq, dtype=np.float64).copy()
    hv, ov, oav = compute_vel_from_seq(hand, obj, motion_freq)
    bp, bq = compute_body_pos_quat_from_seq(M, hand)
    return {"hand_dof_seq": hand, "hand_dof_vel_seq": hv, "obj_pose_seq": obj, "obj_vel_seq": ov,
            "obj_angle_vel_seq": oav, "body_pos_seq": bp, "body_quat_seq": bq, "seq_len": hand.shape[0]}


EXPERT_KEYS = ("hand_dof_seq", "hand_dof_vel_seq", "obj_pose_seq", "obj_vel_seq", "obj_angle_vel_seq",
               "body_pos_seq", "body_quat_seq")


def load_seq(expert, start_idx=0):
    """Slice as the sampler does (agent_handmimic.py:450, load_seq(..., full_seq=True)): frames [start_idx, end)."""
    s = min(start_idx, expert["seq_len"])
    out = {k: expert[k][s:] for k in EXPERT_KEYS}
    out["seq_len"] = expert["seq_len"] - s
    return out


def pack_frames(expert) -> np.ndarray:
    T = expert["hand_dof_seq"].shape[0]
    f = np.concatenate([expert[k].reshape(T, -1) for k in EXPERT_KEYS], axis=1)
    assert f.shape[1] == FRAME_DIM, f.shape
    return np.ascontiguousarray(f, dtype=np.float64)


# ----------------------------------------------------------------------------- synthetic motions

TABLE_TOP = 0.5  # C_Table half-height 0.25 centred at z=0.25 (sphere_mesh_hand_add_geom.xml:108)
_FINGERS = ("FF", "MF", "LF", "RF", "TH")


def _minjerk(s):
    s = np.clip(s, 0, 1)
    return 10 * s ** 3 - 15 * s ** 4 + 6 * s ** 5


def _euler_R(a, b, c):
    ca, sa, cb, sb, cc, sc = np.cos(a), np.sin(a), np.cos(b), np.sin(b), np.cos(c), np.sin(c)
    Rx = np.array([[1, 0, 0], [0, ca, -sa], [0, sa, ca]])
    Ry = np.array([[cb, 0, sb], [0, 1, 0], [-sb, 0, cb]])
    Rz = np.array([[cc, -sc, 0], [sc, cc, 0], [0, 0, 1]])
    return Rx @ Ry @ Rz


class _ObjectShape:
    """Convex parts of the object in its body frame as half-spaces (signed-distance lower bound outside)."""

    def __init__(self, M):
        from scipy.spatial import ConvexHull
        o0, o1 = M["env"]["obj_geom_range"]
        self.parts, self.verts = [], []
        for g in range(o0, o1 + 1):
            R, p = mjcf.quat2mat(M["geom_quat"][g]), M["geom_pos"][g]
            if M["geom_type"][g] == mjcf.GEOM_BOX:
                s = M["geom_size"][g]
                v = np.array([[sx, sy, sz] for sx in (-s[0], s[0]) for sy in (-s[1], s[1]) for sz in (-s[2], s[2])])
            else:
                a, n = M["geom_vertadr"][g], M["geom_vertnum"][g]
                v = M["mesh_vert"][a:a + n]
            v = v @ R.T + p
            eq = ConvexHull(v).equations
            self.parts.append((eq[:, :3], eq[:, 3]))
            self.verts.append(v)
        self.all_verts = np.concatenate(self.verts)

    def dist(self, pts_body):
        d = np.full(pts_body.shape[0], np.inf)
        for n, b in self.parts:
            d = np.minimum(d, (pts_body @ n.T + b).max(1))
        return d


def _finger_capsules(M):
    """geom ids of each finger's capsules, by name prefix robot0:C_<FINGER>[123]."""
    out = {}
    for f in _FINGERS:
        out[f] = [i for i, n in enumerate(M["names_geom"]) if n.startswith(f"robot0:C_{f}")]
    return out


def _capsule_points(M, kin, gids, k=7):
    pts, rad = [], []
    for g in gids:
        b = M["geom_bodyid"][g]
        Rb = mjcf.quat2mat(kin["xquat"][b])
        c = kin["xpos"][b] + Rb @ M["geom_pos"][g]
        ax = Rb @ mjcf.quat2mat(M["geom_quat"][g])[:, 2]
        for s in np.linspace(-1, 1, k):
            pts.append(c + ax * s * M["geom_size"][g][1])
            rad.append(M["geom_size"][g][0])
    return np.array(pts), np.array(rad)


def _close_fingers(M, shape, qpos, obj_R, obj_p, pen=0.002, step=0.02):
    """Flex each finger joint in turn (proximal to distal) until the capsules it moves touch the object with
    penetration `pen`; the flexion direction of a joint is whichever closes the gap."""
    caps = _finger_capsules(M)
    jn = M["names_joint"]
    q = qpos.copy()

    def gap(gids):
        kin = mjcf.fk(M, q)
        pts, rad = _capsule_points(M, kin, gids)
        return float((shape.dist((pts - obj_p) @ obj_R) - rad).min())

    for f in _FINGERS:
        plan = [("J3x", 0), ("J2x", 1), ("J1x", 2)]
        if f == "TH":
            plan = [("J3z", 0)] + plan
        for suffix, lvl in plan:
            j = jn.index(f"robot0:{f}{suffix}")
            lo, hi = M["jnt_range"][j]
            a = M["jnt_qposadr"][j]
            moved = caps[f][lvl:]
            q0 = q[a]
            q[a] = np.clip(q0 + step, lo, hi)
            gp = gap(moved)
            q[a] = np.clip(q0 - step, lo, hi)
            gm = gap(moved)
            q[a] = q0
            if abs(gp - gm) < 1e-6:
                continue
            sgn = 1.0 if gp < gm else -1.0
            while gap(moved) > -pen:
                nxt = q[a] + sgn * step
                if nxt < lo + 0.03 or nxt > hi - 0.03:
                    break
                q[a] = nxt
    return q


def synth_motion(M, T=300, seed=1, motion_freq=30, grasp="auto"):
    """Deterministic approach - grasp - lift - hold script -> expert dict.

    frames 0-59 approach with open fingers, 60-119 close, 120-209 lift 0.10 m (minimum jerk), 210-299 hold with a
    2 Hz / 0.05 rad wrist roll; the object rests on the table until frame 120 and then follows the palm rigidly.
    The object is held like a stick between the thumb and the four fingers: long axis along the palm x axis, about
    5.5 cm in front of the palm.  "side": object upright, fingers horizontal; "top": object lying, palm down.
    `seed` != 1 varies object yaw / placement, lift height and adds smooth low-amplitude offsets.
    """
    rng = np.random.RandomState(seed)
    shape = _ObjectShape(M)
    V = shape.all_verts
    ext = V.max(0) - V.min(0)
    order = np.argsort(-ext)  # long, medium, thin object axes
    if grasp == "auto":
        grasp = "side" if ext[order[0]] < 0.15 else "top"
    if grasp == "side":
        euler = np.array([-np.pi / 2, 0.0, -np.pi / 2])  # palm x -> world z, fingers (y) -> world x, normal (-z) -> -y
        assign = (order[0], order[2], order[1])           # object axes mapped onto palm x, y, z
    else:
        euler = np.array([0.0, 0.0, 0.0])                 # palm down, fingers along world y
        assign = (order[0], order[1], order[2])           # thin axis vertical: the stable way to lie
    R = _euler_R(*euler)
    # object rotation: object axis assign[k] -> palm axis k (sign fixed to keep it right-handed)
    P = np.zeros((3, 3))
    for k in range(3):
        P[assign[k], k] = 1.0
    if np.linalg.det(P) < 0:
        P[assign[2], 2] = -1.0
    obj_R = R @ P.T
    yaw = (rng.rand() - 0.5) * 0.2 if seed != 1 else 0.0
    cy, sy = np.cos(yaw), np.sin(yaw)
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    obj_R, R = Rz @ obj_R, Rz @ R
    euler = np.array(_euler_from_R(R))
    q_obj = mjcf.mat2quat(obj_R)
    Vw = V @ obj_R.T
    obj_p = np.array([0.0, 0.0, TABLE_TOP - Vw[:, 2].min() + 2e-4])
    if seed != 1:
        obj_p[:2] += (rng.rand(2) - 0.5) * 0.04
    Vw = Vw + obj_p
    ex, ey, ez = R[:, 0], R[:, 1], R[:, 2]
    px, py, pz = Vw @ ex, Vw @ ey, Vw @ ez
    # object centre at palm-frame (-0.01, 0.058 + half width, -0.055): clear of the zero-pose thumb, below the knuckles
    root = (0.5 * (px.max() + px.min()) + 0.01) * ex + (py.min() - 0.058) * ey + (0.5 * (pz.max() + pz.min()) + 0.055) * ez
    nh = M["nq"] - 7
    lo, hi = M["jnt_range"][:-1, 0], M["jnt_range"][:-1, 1]
    q_open = np.zeros(M["nq"])
    q_open[:3], q_open[3:6] = root, euler
    q_open[6:nh] = np.clip(0.0, lo[6:] + 0.03, hi[6:] - 0.03)
    jn = M["names_joint"]
    for f in _FINGERS[:4]:
        q_open[M["jnt_qposadr"][jn.index(f"robot0:{f}J3x")]] = 0.2
    q_open[nh:nh + 3], q_open[nh + 3:] = obj_p, q_obj
    q_grasp = _close_fingers(M, shape, q_open, obj_R, obj_p)
    fingers_open = q_open[6:nh]
    t = np.arange(T)
    hand = np.zeros((T, nh))
    obj = np.zeros((T, 7))
    approach = 0.08 * (1 - _minjerk(t / 59.0))
    lift_h = 0.10 if seed == 1 else 0.07 + 0.06 * rng.rand()
    lift = lift_h * _minjerk((t - 120) / 89.0)
    hand[:, :3] = root + approach[:, None] * ez[None, :] + lift[:, None] * np.array([0, 0, 1.0])
    hand[:, 3:6] = euler
    roll = np.where(t >= 210, 0.05 * np.sin(2 * np.pi * 2.0 * (t - 210) / motion_freq), 0.0)
    hand[:, 3] += roll
    blend = 0.5 * (1 - np.cos(np.pi * np.clip((t - 60) / 59.0, 0, 1)))
    hand[:, 6:] = fingers_open[None, :] * (1 - blend[:, None]) + q_grasp[None, 6:nh] * blend[:, None]
    if seed != 1:  # smooth per-seed offsets: ~N(0, 2 mm) on the root, ~N(0, 0.01 rad) on joints, low-pass over ~1 s
        k = np.hanning(31)
        k /= k.sum()
        for c in range(nh):
            amp = 0.002 if c < 3 else 0.01
            hand[:, c] += np.convolve(rng.randn(T + 30) * amp * 3.0, k, mode="valid")
    hand = np.clip(hand, lo, hi)
    # object: at rest until 120, then rigidly attached to the palm
    Rp = np.array([_euler_R(*hand[i, 3:6]) for i in range(T)])
    R120, p120 = Rp[120], hand[120, :3]
    for i in range(T):
        if i <= 120:
            obj[i, :3], obj[i, 3:] = obj_p, q_obj
        else:
            dR = Rp[i] @ R120.T
            obj[i, :3] = hand[i, :3] + dR @ (obj_p - p120)
            obj[i, 3:] = mjcf.mat2quat(dR @ obj_R)
            if np.dot(obj[i, 3:], obj[i - 1, 3:]) < 0:
                obj[i, 3:] *= -1
    return derive_expert(M, hand, obj, motion_freq)


def _euler_from_R(R):
    """Inverse of _euler_R (R = Rx(a) Ry(b) Rz(c)), |b| < pi/2."""
    b = np.arcsin(np.clip(R[0, 2], -1, 1))
    a = np.arctan2(-R[1, 2], R[2, 2])
    c = np.arctan2(-R[0, 1], R[0, 0])
    return a, b, c
