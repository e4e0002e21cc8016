"""Physics metrics of a simulated sequence: the batched mirror of `PhysMetrics` (scripts/metrics.py:17-247).

The reference loops over the frames on the CPU (`sim.forward()` per frame, one `qpsolvers` call per frame); here every
frame is one environment of a `HandMimicVecEnv`, the hand-object contacts of all frames come from one forward pass and the
per-frame stability QPs (`solve_force`, :156-196) are solved by `hoic_solve_wrench_qp`, one warp per frame.
`eval_penetration` evaluates the mesh's signed distance with `mesh_sdf` below (pysdf is not available); `eval_penetration2` (the box-specific
analytic depth) is.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import capi
from .assets import Model


def solve_force(contacts, n_contacts, target, center, mu=1.0, w_t=1.0, weights=None, penalties=None, surface_contact=True,
                want_surface=False, device=0):
    """Batched `PhysMetrics.solve_force` / QP of `solve_rfc`: contacts [B, K, 12] (pos, normal, tangent, bitangent),
    n_contacts [B], target [B, 6] (force, torque), center [B, 3].  Returns dict(rest [B,6], score [B], iters [B],
    surface [B, SURF_STRIDE] or None).  score = |rest_f| + sqrt(w_t) |rest_t| (= the reference's rest_ft for w_t = 1)."""
    if not torch.cuda.is_available():
        raise capi.HoicError("solve_force needs a CUDA device (there is no CPU path)")
    L = capi.lib()
    dev = torch.device("cuda", device)
    f32 = dict(dtype=torch.float32, device=dev)
    contacts = torch.as_tensor(contacts, **f32).contiguous()
    B, K = contacts.shape[0], contacts.shape[1]
    assert contacts.shape == (B, K, 12)
    ncon = torch.as_tensor(n_contacts, dtype=torch.int32, device=dev).contiguous()
    target = torch.as_tensor(target, **f32).contiguous().reshape(B, 6)
    center = torch.as_tensor(center, **f32).contiguous().reshape(B, 3)
    weights = None if weights is None else torch.as_tensor(weights, **f32).contiguous().reshape(B, K)
    ec = 5 if surface_contact else 1
    penalties = None if penalties is None else torch.as_tensor(penalties, **f32).contiguous().reshape(B, K * ec * 4)
    rest, score = torch.empty(B, 6, **f32), torch.empty(B, **f32)
    iters = torch.empty(B, dtype=torch.int32, device=dev)
    surface = torch.empty(B, capi.SURF_STRIDE, **f32) if want_surface else None
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    capi.check(L.hoic_solve_wrench_qp(device, B, K, 1 if surface_contact else 0, capi.ptr(contacts), capi.ptr(ncon), capi.ptr(weights),
                                      capi.ptr(penalties), capi.ptr(target), capi.ptr(center), C.c_float(mu), C.c_float(w_t),
                                      capi.ptr(rest), capi.ptr(score), capi.ptr(surface), capi.ptr(iters), stream), "hoic_solve_wrench_qp")
    return dict(rest=rest, score=score, iters=iters, surface=surface)


def _gradient(x):
    """np.gradient(x, axis=0) on a torch tensor: central differences inside, one-sided at the two ends."""
    g = torch.empty_like(x)
    g[1:-1] = (x[2:] - x[:-2]) * 0.5
    g[0], g[-1] = x[1] - x[0], x[-1] - x[-2]
    return g


def _quat_to_mat(q):
    q = q / q.norm(dim=-1, keepdim=True)
    w, x, y, z = q.unbind(-1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1).reshape(q.shape[:-1] + (3, 3))


def _matrix_to_axis_angle(R):
    """pytorch3d-style conversion the reference uses (uhc/utils/transforms.py:99-158, 414-490), via the quaternion."""
    m00, m11, m22 = R[..., 0, 0], R[..., 1, 1], R[..., 2, 2]
    qa = torch.stack([1 + m00 + m11 + m22, 1 + m00 - m11 - m22, 1 - m00 + m11 - m22, 1 - m00 - m11 + m22], -1)
    qa = torch.sqrt(torch.clamp(qa, min=0))
    cand = torch.stack([
        torch.stack([qa[..., 0] ** 2, R[..., 2, 1] - R[..., 1, 2], R[..., 0, 2] - R[..., 2, 0], R[..., 1, 0] - R[..., 0, 1]], -1),
        torch.stack([R[..., 2, 1] - R[..., 1, 2], qa[..., 1] ** 2, R[..., 1, 0] + R[..., 0, 1], R[..., 0, 2] + R[..., 2, 0]], -1),
        torch.stack([R[..., 0, 2] - R[..., 2, 0], R[..., 1, 0] + R[..., 0, 1], qa[..., 2] ** 2, R[..., 1, 2] + R[..., 2, 1]], -1),
        torch.stack([R[..., 1, 0] - R[..., 0, 1], R[..., 2, 0] + R[..., 0, 2], R[..., 2, 1] + R[..., 1, 2], qa[..., 3] ** 2], -1)], -2)
    best = qa.argmax(-1)
    q = torch.gather(cand, -2, best[..., None, None].expand(best.shape + (1, 4))).squeeze(-2)
    q = q / (2 * torch.clamp(torch.gather(qa, -1, best[..., None]), min=0.1))
    nrm = q[..., 1:].norm(dim=-1)
    half = torch.atan2(nrm, q[..., 0])
    ang = 2 * half
    small = ang.abs() < 1e-6
    s = torch.where(small, 0.5 - ang * ang / 48, torch.sin(half) / torch.where(small, torch.ones_like(ang), ang))
    return q[..., 1:] / s[..., None]


def mesh_sdf(points, tri, chunk=512):
    """Signed distance of `points` [P, 3] to the closed triangle mesh `tri` [F, 3, 3], POSITIVE INSIDE (pysdf's convention,
    scripts/metrics.py:49-50, 68-70): exact point-triangle distance (Ericson's closest point on triangle), sign from the
    generalised winding number (sum of the triangles' signed solid angles / 4 pi > 1/2).  Batched torch ops, any device."""
    tri = tri.to(points.dtype)
    a, b, c = tri[:, 0], tri[:, 1], tri[:, 2]
    out = torch.empty(points.shape[0], dtype=points.dtype, device=points.device)
    for s in range(0, points.shape[0], chunk):
        p = points[s:s + chunk, None, :]                                   # [p, 1, 3]
        ab, ac, ap = b - a, c - a, p - a
        d1, d2 = (ab * ap).sum(-1), (ac * ap).sum(-1)
        bp = p - b
        d3, d4 = (ab * bp).sum(-1), (ac * bp).sum(-1)
        cp = p - c
        d5, d6 = (ab * cp).sum(-1), (ac * cp).sum(-1)
        va, vb, vc = d3 * d6 - d5 * d4, d5 * d2 - d1 * d6, d1 * d4 - d3 * d2
        eps = torch.finfo(points.dtype).tiny
        # interior of the face by default, then the edge / vertex regions override (Ericson, Real-Time Collision Detection 5.1.5)
        denom = 1.0 / (va + vb + vc).clamp_min(eps)
        v, w = vb * denom, vc * denom
        q = a + ab * v[..., None] + ac * w[..., None]
        wbc = ((d4 - d3) / ((d4 - d3) + (d5 - d6)).clamp_min(eps))[..., None]
        q = torch.where(((va <= 0) & (d4 - d3 >= 0) & (d5 - d6 >= 0))[..., None], b + (c - b) * wbc, q)
        wac = (d2 / (d2 - d6).clamp_min(eps))[..., None]
        q = torch.where(((vb <= 0) & (d2 >= 0) & (d6 <= 0))[..., None], a + ac * wac, q)
        wab = (d1 / (d1 - d3).clamp_min(eps))[..., None]
        q = torch.where(((vc <= 0) & (d1 >= 0) & (d3 <= 0))[..., None], a + ab * wab, q)
        q = torch.where(((d6 >= 0) & (d5 <= d6))[..., None], c.expand_as(q), q)
        q = torch.where(((d3 >= 0) & (d4 <= d3))[..., None], b.expand_as(q), q)
        q = torch.where(((d1 <= 0) & (d2 <= 0))[..., None], a.expand_as(q), q)
        dist = (p - q).norm(dim=-1).min(dim=1).values
        # winding number: solid angle of every triangle seen from p (Van Oosterom & Strackee)
        ra, rb, rc = a - p, b - p, c - p
        la, lb, lc = ra.norm(dim=-1), rb.norm(dim=-1), rc.norm(dim=-1)
        num = (ra * torch.cross(rb, rc, dim=-1)).sum(-1)
        den = la * lb * lc + (ra * rb).sum(-1) * lc + (rb * rc).sum(-1) * la + (rc * ra).sum(-1) * lb
        wind = (2.0 * torch.atan2(num, den)).sum(1) / (4.0 * np.pi)
        out[s:s + chunk] = torch.where(wind.abs() > 0.5, dist, -dist)
    return out


class PhysMetrics:
    """`PhysMetrics(model, qpos_seq, obj_mesh_fn)` of the reference with the model given as a compiled `Model`;
    `qpos_seq` [T, 33] at 30 Hz."""

    def __init__(self, model: Model, qpos_seq, device=0, motion_freq=30.0):
        from .vecenv import HandMimicVecEnv
        self.model, self.freq = model, float(motion_freq)
        self.device = torch.device("cuda", device)
        self.qpos_seq = torch.as_tensor(np.asarray(qpos_seq), dtype=torch.float32, device=self.device).contiguous()
        T = self.qpos_seq.shape[0]
        A = model.arrays
        ids = A["env_ids"]
        self.hand_geom_range, self.obj_geom_range = (int(ids[2]), int(ids[3])), (int(ids[4]), int(ids[5]))
        self.hand_body_idx = torch.as_tensor(np.asarray(A["env_hand_body_idx"]).astype(np.int64), device=self.device)
        self.obj_mass = float(A["body_mass"][-1])
        self.obj_inertia = torch.as_tensor(np.asarray(A["body_inertia"][-1]), dtype=torch.float32, device=self.device)
        self.box_size = torch.tensor([0.0165, 0.0265, 0.049], device=self.device)       # scripts/metrics.py:43
        # one environment per frame: a single forward gives the poses and contacts of the whole sequence
        env = HandMimicVecEnv(model, T, [model.motion(1)], device=device, auto_reset=False, random_start=False, balance=False)
        env.set_state(self.qpos_seq, torch.zeros(T, capi.NV, device=self.device))
        self.ncon, pairs, data = env.get_contacts()
        self.body_xpos = env.get_derived()[0]
        env.close()
        g1, g2 = pairs[..., 0], pairs[..., 1]
        slot = torch.arange(capi.MAX_CONTACTS, device=self.device)[None, :]
        self.ho_mask = ((slot < self.ncon[:, None]) & (g1 >= self.hand_geom_range[0]) & (g1 <= self.hand_geom_range[1]) &
                        (g2 >= self.obj_geom_range[0]) & (g2 <= self.obj_geom_range[1]))
        self.contact_pos, self.contact_frame = data[..., 1:4], data[..., 4:13]

    # ---- scripts/metrics.py:141-153
    def eval_contact_point(self):
        return self.ho_mask.sum(1)

    # ---- scripts/metrics.py:52-75: signed distance (positive inside) of the object's mesh at every hand-object contact position,
    # negative values clamped to 0, mean over the frame's contacts, in mm.  `obj_mesh`: path of a binary STL (read raw like
    # trimesh.load_mesh does: no MJCF scale) or a [F, 3, 3] triangle array.
    def eval_penetration(self, obj_mesh):
        from . import mjcf
        tri = mjcf.load_stl(obj_mesh, 1.0) if isinstance(obj_mesh, str) else np.asarray(obj_mesh, dtype=np.float64)
        tri = torch.as_tensor(tri, dtype=torch.float32, device=self.device)
        obj_pos, R = self.qpos_seq[:, -7:-4], _quat_to_mat(self.qpos_seq[:, -4:])
        local = torch.einsum("tji,tcj->tci", R, self.contact_pos - obj_pos[:, None, :])        # R^-1 (p - obj_pos)
        idx = self.ho_mask.nonzero()
        pene = torch.zeros(self.qpos_seq.shape[0], device=self.device)
        if idx.numel():
            sdf = mesh_sdf(local[idx[:, 0], idx[:, 1]], tri).clamp_min(0.0)
            pene.index_add_(0, idx[:, 0], sdf)
        return pene * 1000 / torch.clamp(self.ho_mask.sum(1), min=1)

    # ---- scripts/metrics.py:76-101 (box only: depth of each hand-object contact point inside the box, x2, in mm, / ncon)
    def eval_penetration2(self):
        obj_pos, R = self.qpos_seq[:, -7:-4], _quat_to_mat(self.qpos_seq[:, -4:])
        local = torch.einsum("tji,tcj->tci", R, self.contact_pos - obj_pos[:, None, :])        # R^-1 (p - obj_pos)
        inside = (local.abs() < self.box_size).all(-1)
        depth = (self.box_size - local.abs()).min(-1).values
        pene = torch.where(inside & self.ho_mask, depth * 2, torch.zeros_like(depth)).sum(1)
        return pene * 1000 / torch.clamp(self.ncon, min=1)

    # ---- scripts/metrics.py:103-139
    def eval_jitter(self):
        f = self.freq
        obj_acc = _gradient(_gradient(self.qpos_seq[:, -7:-4]) * f) * f
        ang_acc = _gradient(self._obj_ang_vel()) * f
        joint = self.body_xpos[:, self.hand_body_idx]
        joint_acc = _gradient(_gradient(joint) * f) * f
        return float(joint_acc.norm(dim=-1).mean()), float(obj_acc.norm(dim=-1).mean()), float(ang_acc.norm(dim=-1).mean())

    def _obj_ang_vel(self):
        R = _quat_to_mat(self.qpos_seq[:, -4:])
        w = torch.zeros(R.shape[0], 3, device=self.device)
        w[1:] = _matrix_to_axis_angle(R[1:] @ R[:-1].transpose(1, 2)) * self.freq
        w[0] = w[1]
        return w

    # ---- scripts/metrics.py:198-224
    def obtain_target_ft(self):
        f = self.freq
        acc = _gradient(_gradient(self.qpos_seq[:, -7:-4]) * f) * f
        w = self._obj_ang_vel()
        wa = _gradient(w) * f
        R = _quat_to_mat(self.qpos_seq[:, -4:])
        Is = R @ torch.diag(self.obj_inertia)[None] @ R.transpose(1, 2)
        force = self.obj_mass * (acc + torch.tensor([0.0, 0.0, 9.8], device=self.device))
        torque = (Is @ wa[..., None]).squeeze(-1) + torch.cross(w, (Is @ w[..., None]).squeeze(-1), dim=-1)
        return force, torque

    # ---- scripts/metrics.py:226-247
    def eval_stable(self, return_rest=False):
        force, torque = self.obtain_target_ft()
        T, K = self.qpos_seq.shape[0], capi.NHG
        n = self.ho_mask.sum(1)
        if int(n.max()) > K:
            raise capi.HoicError(f"eval_stable: a frame has {int(n.max())} hand-object contacts, the QP kernel holds {K}")
        # compact the hand-object contacts of every frame to the front, keeping their order
        order = torch.argsort((~self.ho_mask).to(torch.int8), dim=1, stable=True)[:, :K]
        pos = torch.gather(self.contact_pos, 1, order[..., None].expand(T, K, 3))
        frm = torch.gather(self.contact_frame, 1, order[..., None].expand(T, K, 9))
        contacts = torch.cat([pos, frm], -1)
        out = solve_force(contacts, n, torch.cat([force, torque], -1), self.qpos_seq[:, -7:-4], mu=1.0, w_t=1.0, device=self.device.index)
        rest_ft = out["score"] / self.obj_mass
        stable = (rest_ft > 0.01).to(torch.float32)      # 1 = the contacts cannot hold the object (> 0.01 m/s^2 unexplained)
        return (stable, rest_ft) if return_rest else stable
