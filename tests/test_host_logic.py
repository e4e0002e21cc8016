"""CPU tests: host logic (MJCF compiler, blob, config, motion derivations), the C-ABI surface, and the oracle pinned by
known answers, by a numpy transcription of the reference's pure-Python observation / reward code, and by goldens."""
import ctypes
import os
import re

import numpy as np
import pytest

from common import OBJECTS, ROOT, model, oracle_env
from hoic_b200 import blob as hblob
from hoic_b200 import mjcf, motion as hmotion
from hoic_b200.config import Config, REWARD_KEYS
import ref_numpy as R


# ------------------------------------------------------------------------------------------------ compiler / blob
@pytest.mark.parametrize("obj,ngeom,nverts", [("box", 23, 0), ("bottle", 24, 388), ("banana", 25, 1877)])
def test_compile_invariants(obj, ngeom, nverts):
    """Sizes implied by ho_im4.py:72-107, 139-167 and SURVEY.md section 8."""
    m = model(obj)
    A = m.arrays
    assert (int(A["nq"][0]), int(A["nv"][0]), int(A["nu"][0]), int(A["nbody"][0]), int(A["njnt"][0])) == (33, 32, 26, 25, 27)
    assert int(A["ngeom"][0]) == ngeom and A["mesh_vert"].shape[0] == nverts
    ids = A["env_ids"]
    assert (ids[2], ids[3]) == (2, 20) and ids[4] == 21 and ids[5] == ngeom - 2 and ids[0] == 24 and ids[1] == 3
    assert abs(A["env_scalars"][0] - 0.635) < 1e-12            # hand mass = sum of <inertial mass>
    assert len(A["pair_geom1"]) == 19                          # explicit <pair>s (LF1-RF1 listed twice)
    assert A["opt_timestep"][0] == 0.0022222 and A["opt_iterations"][0] == 20
    M = m.compiled()
    # class= overrides childclass: hand geoms keep the default friction / condim 3, conaffinity 0
    g = M["names_geom"].index("robot0:C_FF1")
    assert A["geom_condim"][g] == 3 and np.allclose(A["geom_friction"][g], [1, 0.005, 0.0001]) and A["geom_conaffinity"][g] == 0
    # joint armature overrides the class value
    assert np.allclose(A["dof_armature"][:26], 0.005) and np.allclose(A["dof_armature"][26:], 0.001)
    assert np.allclose(A["dof_frictionloss"], 0.001) and np.allclose(A["dof_damping"][:26], 0.01) and np.allclose(A["dof_damping"][26:], 0)


def test_object_inertia_box():
    A = model("box").arrays
    s = np.array([0.0165, 0.0265, 0.049])
    mass = 8 * s.prod() * 125
    assert abs(A["body_mass"][24] - mass) < 1e-12
    I = mass / 3 * np.array([s[1] ** 2 + s[2] ** 2, s[0] ** 2 + s[2] ** 2, s[0] ** 2 + s[1] ** 2])
    assert np.allclose(A["body_inertia"][24], I, rtol=1e-12)
    assert np.allclose(A["body_iquat"][24], [1, 0, 0, 0])


def test_mesh_props_closed_form():
    """mesh_props_legacy on a box mesh reproduces the primitive's volume / CoM / inertia."""
    s = np.array([0.3, 0.2, 0.1])
    v = np.array([[x, y, z] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)], dtype=float) * s + np.array([0.5, -0.2, 0.1])
    from scipy.spatial import ConvexHull
    h = ConvexHull(v)
    V, com, I = mjcf.mesh_props_legacy(v[h.simplices])
    assert abs(V - 8 * s.prod()) < 1e-12 and np.allclose(com, [0.5, -0.2, 0.1])
    Iref = V / 3 * np.array([s[1] ** 2 + s[2] ** 2, s[0] ** 2 + s[2] ** 2, s[0] ** 2 + s[1] ** 2])
    assert np.allclose(np.diag(I), Iref) and np.allclose(I - np.diag(np.diag(I)), 0, atol=1e-15)


def test_blob_roundtrip_and_json():
    m = model("box")
    d = hblob.unpack(m.blob)
    b2 = hblob.pack(d)
    d2 = hblob.unpack(b2)
    assert set(d) == set(d2)
    for k in d:
        if isinstance(d[k], bytes):
            assert d[k] == d2[k]
        else:
            assert np.array_equal(d[k], d2[k])
    assert '"body_pos"' in hblob.dump_json(m.blob)
    with pytest.raises(ValueError):
        hblob.unpack(b"garbage" * 10)


def test_shipped_blobs_match_reference_compile():
    """hoic_b200/data/*.hoicmodel (what the GPU box loads) equal a fresh compile when the reference tree is reachable."""
    from hoic_b200 import assets
    if assets.reference_root() is None:
        pytest.skip("reference tree not reachable")
    for obj in OBJECTS:
        fresh = assets.compile_from_reference(obj, assets.reference_root())
        shipped = open(os.path.join(assets.DATA_DIR, f"{obj}.hoicmodel"), "rb").read()
        a, b = hblob.unpack(fresh.blob), hblob.unpack(shipped)
        for k in a:
            if not isinstance(a[k], bytes):
                assert np.allclose(a[k], b[k], rtol=1e-12, atol=1e-15), k


def test_config_adaptive_schedule():
    """handmimic_config.py:175-195: w_orfc,k_orfc = {0.05,0.1} e^{8e-4 min(i,3000)}, k_p,k_j,k_wp = {6,12,3} e^{-5e-4 ...}."""
    cfg = model("box").cfg
    cfg.update_adaptive_params(0)
    w = cfg.reward_vector()
    d = dict(zip(REWARD_KEYS, w))
    assert np.isclose(d["w_orfc"], 0.05) and np.isclose(d["k_orfc"], 0.1) and np.isclose(d["k_p"], 6) and np.isclose(d["k_j"], 12)
    cfg.update_adaptive_params(5000)
    d = dict(zip(REWARD_KEYS, cfg.reward_vector()))
    assert np.isclose(d["w_orfc"], 0.05 * np.exp(8e-4 * 3000)) and np.isclose(d["k_wp"], 3 * np.exp(-5e-4 * 3000))
    cfg.update_adaptive_params(0)
    assert cfg.sim_step == 15 and cfg.future_w_size == 5 and cfg.residual_force_scale == 2.5 and cfg.unsupported == []


def test_motion_derivations():
    M = model("box").compiled()
    ex = model("box").motion(1)
    T = ex["seq_len"]
    assert T == 300 and hmotion.pack_frames(ex).shape == (300, 212)
    hv = ex["hand_dof_vel_seq"]
    assert np.allclose(hv[1:], (ex["hand_dof_seq"][1:] - ex["hand_dof_seq"][:-1]) * 30) and np.allclose(hv[0], hv[1])
    lo, hi = M["jnt_range"][:-1, 0], M["jnt_range"][:-1, 1]
    assert (ex["hand_dof_seq"] >= lo - 1e-12).all() and (ex["hand_dof_seq"] <= hi + 1e-12).all()
    # FK of the expert equals the oracle's FK at that pose
    env = oracle_env("box")
    q = np.concatenate([ex["hand_dof_seq"][150], M["qpos0"][26:]])
    env.set_state(q, np.zeros(32))
    idx = M["env"]["hand_body_idx"]
    assert np.abs(env.get("xpos").reshape(-1, 3)[idx] - ex["body_pos_seq"][150]).max() < 1e-12
    assert np.abs(env.get("xquat").reshape(-1, 4)[idx] - ex["body_quat_seq"][150]).max() < 1e-12
    sl = hmotion.load_seq(ex, 37)
    assert sl["seq_len"] == 263 and np.array_equal(sl["obj_pose_seq"][0], ex["obj_pose_seq"][37])
    # axis-angle of a known rotation
    Rz = np.array([[np.cos(0.3), -np.sin(0.3), 0], [np.sin(0.3), np.cos(0.3), 0], [0, 0, 1]])
    assert np.allclose(hmotion.matrix_to_axis_angle(Rz), [0, 0, 0.3])


# ------------------------------------------------------------------------------------------------ C ABI
def test_capi_exports_every_declared_symbol():
    from hoic_b200 import build, capi
    so = build.build()
    lib = ctypes.CDLL(so)
    hdr = open(os.path.join(ROOT, "include", "hoic.h")).read()
    declared = set(re.findall(r"\b(hoic_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    lib.hoic_version.restype = ctypes.c_int
    assert lib.hoic_version() == 1
    # the Python mirror of the header's layout constants and of hoic_counters (values through the C compiler itself)
    import subprocess, tempfile
    names = ["HOIC_OBS_DIM", "HOIC_ACT_DIM", "HOIC_FRAME_DIM", "HOIC_INFO_DIM", "HOIC_NQ", "HOIC_NV", "HOIC_MAX_CONTACTS", "HOIC_NHG",
             "HOIC_REC_MAXCON", "HOIC_REC_CON", "HOIC_REC_STRIDE", "HOIC_SURF_MAXPT", "HOIC_SURF_PT", "HOIC_SURF_STRIDE"]
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "c.c")
        open(src, "w").write('#include <stdio.h>\n#include "hoic.h"\nint main(void){' + "".join('printf("%s %%d\\n", (int)(%s));' % (n, n) for n in names) +
                             'printf("COUNTERS %d\\n", (int)(sizeof(hoic_counters) / sizeof(unsigned long long))); return 0;}')
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(td, "c"), src])
        out = dict(l.split() for l in subprocess.check_output([os.path.join(td, "c")]).decode().splitlines())
    val = lambda n: int(out[n])
    assert val("COUNTERS") == len(capi.Counters._fields_)
    assert (val("HOIC_OBS_DIM"), val("HOIC_ACT_DIM"), val("HOIC_FRAME_DIM"), val("HOIC_INFO_DIM")) == (capi.OBS_DIM, capi.ACT_DIM, capi.FRAME_DIM, capi.INFO_DIM)
    assert (val("HOIC_NQ"), val("HOIC_NV"), val("HOIC_MAX_CONTACTS"), val("HOIC_NHG")) == (capi.NQ, capi.NV, capi.MAX_CONTACTS, capi.NHG)
    assert (val("HOIC_REC_MAXCON"), val("HOIC_REC_CON"), val("HOIC_REC_STRIDE")) == (capi.REC_MAXCON, capi.REC_CON, capi.REC_STRIDE)
    assert (val("HOIC_SURF_MAXPT"), val("HOIC_SURF_PT"), val("HOIC_SURF_STRIDE")) == (capi.SURF_MAXPT, capi.SURF_PT, capi.SURF_STRIDE)
    import oracle
    assert (oracle.REC_STRIDE, oracle.SURF_STRIDE) == (capi.REC_STRIDE, capi.SURF_STRIDE)
    body = re.search(r"typedef struct hoic_counters \{(.*?)\} hoic_counters;", hdr, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = [f.strip() for f in re.sub(r"unsigned long long", "", body).replace(";", ",").split(",") if f.strip()]
    assert fields == [n for n, _ in capi.Counters._fields_], (fields, capi.Counters._fields_)


def test_vecenv_refuses_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from hoic_b200 import capi
    from hoic_b200.vecenv import HandMimicVecEnv
    with pytest.raises(capi.HoicError):
        HandMimicVecEnv(model("box"), 2, [model("box").motion(1)])


# ------------------------------------------------------------------------------------------------ oracle: known answers
def test_oracle_mass_matrix_vs_jacobian_sum():
    env = oracle_env("bottle")
    env.reset()
    rng = np.random.RandomState(0)
    M = model("bottle").compiled()
    for _ in range(3):
        q = env.get("qpos")
        q[:26] += rng.randn(26) * 0.2
        q[29:33] = rng.randn(4)
        q[29:33] /= np.linalg.norm(q[29:33])
        env.set_state(q, np.zeros(32))
        H, _, _, _ = mjcf.mass_matrix(M, env.get("qpos"))
        assert np.abs(env.qM() - H).max() < 1e-13


def test_oracle_bias_is_gravity_gradient_and_coriolis():
    """qfrc_bias(q, 0) = dV/dq for the hand joints, and the velocity part matches d/dt(M) v - 1/2 d(v'Mv)/dq."""
    env = oracle_env("box")
    env.reset()
    M = model("box").compiled()
    rng = np.random.RandomState(1)
    q = env.get("qpos")
    q[3:26] += rng.randn(23) * 0.2
    nb = M["nbody"]

    def potential(qq):
        kin = mjcf.fk(M, qq)
        z = [kin["xpos"][b][2] + (mjcf.quat2mat(kin["xquat"][b]) @ M["body_ipos"][b])[2] for b in range(nb)]
        return 9.81 * float(np.dot(M["body_mass"], z))

    env.set_state(q, np.zeros(32))
    bias0 = env.get("qfrc_bias")
    g = np.zeros(26)
    for i in range(26):
        dq = np.zeros(33); dq[i] = 1e-6
        g[i] = (potential(q + dq) - potential(q - dq)) / 2e-6
    assert np.abs(bias0[:26] - g).max() < 2e-6
    v = np.zeros(32); v[:26] = rng.randn(26)
    env.set_state(q, v)
    cor = env.get("qfrc_bias") - bias0

    def Mq(qq):
        return mjcf.mass_matrix(M, qq)[0]
    e = 1e-6
    qd = np.zeros(33); qd[:26] = v[:26]
    Mdot_v = (Mq(q + e * qd) - Mq(q - e * qd)) @ v / (2 * e)
    grad = np.array([v @ (Mq(q + e * np.eye(33)[i]) - Mq(q - e * np.eye(33)[i])) @ v / (2 * e) for i in range(26)])
    assert np.abs(cor[:26] - (Mdot_v[:26] - 0.5 * grad)).max() < 1e-6


def test_oracle_free_fall_and_rest():
    env = oracle_env("box")
    env.reset()
    A = model("box").arrays
    h, m = A["opt_timestep"][0], A["body_mass"][24]
    q = env.get("qpos"); q[26:29] = [0.5, 0.5, 1.5]; q[29:33] = [1, 0, 0, 0]
    env.set_state(q, np.zeros(32)); env.set_ctrl_applied(np.zeros(26), np.zeros(32))
    for _ in range(100):
        env.mj_step()
    # weight m g on inertia m + armature, opposed by the 0.001 N friction loss of the free joint
    a_eff = (m * 9.81 - 0.001) / (m + 0.001)
    assert abs(env.get("qvel")[28] + a_eff * h * 100) < 1e-9
    q = env.get("qpos"); q[26:29] = [0.4, 0.4, 0.549 - 1e-4]; q[29:33] = [1, 0, 0, 0]; q[0:3] = [0, 0, 1.0]
    env.set_state(q, np.zeros(32))
    for _ in range(600):
        env.mj_step()
    assert np.abs(env.get("qvel")[26:]).max() < 1e-10
    assert abs(env.get("qfrc_constraint")[28] - m * 9.81) < 1e-9          # contact force balances the weight
    c = env.contacts()
    tab = c[c[:, 13] == 1]
    assert len(tab) == 4 and np.allclose(tab[:, 4:7], [0, 0, 1]) and set(tab[:, 15]) == {4.0}
    # mj_contactForce (ho_im4.py:876): the four normal forces carry the weight, no friction is needed at rest
    cf = env.contact_forces()[c[:, 13] == 1]
    assert abs(cf[:, 0].sum() - m * 9.81) < 1e-9 and (cf[:, 0] > 0).all() and np.abs(cf[:, 1:]).max() < 1e-9


def test_oracle_contact_force_decodes_the_pyramid():
    """The decoded contact forces (mj_contactForce), rotated to the world frame and summed over the object's contacts,
    reproduce the constraint force on the object's translational dofs up to its 0.001 N friction-loss rows."""
    env = oracle_env("box")
    env.reset()
    rng = np.random.RandomState(3)
    seen = 0
    for t in range(40):
        env.step(np.clip(rng.randn(32) * 0.1, -1, 1))
        c, cf, f = env.contacts(), env.contact_forces(), env.get("efc_force")
        adr = len(f) - int(sum(1 if d == 1 else 2 * (d - 1) for d in c[:, 15]))
        tot = np.zeros(3)
        for k in range(len(c)):
            dim = int(c[k, 15])
            ne = 1 if dim == 1 else 2 * (dim - 1)
            fe = f[adr:adr + ne]
            adr += ne
            assert abs(cf[k, 0] - fe.sum()) < 1e-12 and not cf[k, dim:].any()
            if c[k, 14] >= 21:                      # geom2 is an object geom: the force acts on the object along +normal
                tot += cf[k, :3] @ c[k, 4:13].reshape(3, 3)
                seen += dim > 1
        assert np.abs(env.get("qfrc_constraint")[26:29] - tot).max() < 0.001 + 1e-9
    assert seen >= 10


def test_oracle_newton_reaches_the_optimum():
    from scipy.optimize import minimize
    env = oracle_env("bottle")
    env.reset()
    rng = np.random.RandomState(0)
    for _ in range(40):
        env.step(rng.randn(32) * 0.1)
    J, Mq, aref, D = env.efc_J(), env.qM(), env.get("efc_aref"), env.get("efc_D")
    qacc, qs, force = env.get("qacc"), env.get("qacc_smooth"), env.get("efc_force")
    Rr, fl = 1 / D, 0.001

    def cost(a):
        jar = J @ a - aref
        c = 0.5 * (a - qs) @ Mq @ (a - qs)
        for r, x in enumerate(jar):
            if r < 32:
                Rf = Rr[r] * fl
                c += 0.5 * D[r] * x * x if abs(x) < Rf else fl * (abs(x) - 0.5 * Rf)
            elif x < 0:
                c += 0.5 * D[r] * x * x
        return c
    res = minimize(cost, qacc, method="BFGS", options=dict(gtol=1e-12, maxiter=4000))
    assert cost(qacc) <= res.fun + 1e-9 * abs(res.fun)
    assert np.linalg.norm(Mq @ (qacc - qs) - J.T @ force) < 1e-9          # KKT stationarity


def test_oracle_collision_closed_forms():
    env = oracle_env("box")
    env.reset()
    q = env.get("qpos")
    # box lying flat 1 mm inside the table, rotated 30 deg about z: 4 corner contacts, dist = -1 mm, pos mid-surface
    q[0:3] = [0, 0, 1.5]
    th = np.pi / 6
    q[26:29] = [0.1, -0.2, 0.5 + 0.049 - 1e-3]; q[29:33] = [np.cos(th / 2), 0, 0, np.sin(th / 2)]
    env.set_state(q, np.zeros(32))
    c = env.contacts(); c = c[(c[:, 13] == 1) & (c[:, 14] == 21)]
    assert len(c) == 4 and np.allclose(c[:, 0], -1e-3, atol=1e-12) and np.allclose(c[:, 3], 0.5 - 0.5e-3, atol=1e-12)
    corners = np.array([[sx * 0.0165, sy * 0.0265] for sx in (-1, 1) for sy in (-1, 1)])
    Rz = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    want = corners @ Rz.T + [0.1, -0.2]
    got = c[:, 1:3]
    assert all(np.min(np.linalg.norm(want - g, axis=1)) < 1e-12 for g in got)
    # object dropped to the floor away from the table: plane-box, lowest corners
    q[26:29] = [2.0, 2.0, 0.049 - 2e-3]; q[29:33] = [1, 0, 0, 0]
    env.set_state(q, np.zeros(32))
    c = env.contacts(); c = c[(c[:, 13] == 0) & (c[:, 14] == 21)]
    assert len(c) == 4 and np.allclose(c[:, 0], -2e-3, atol=1e-12) and np.allclose(c[:, 4:7], [0, 0, 1])


def test_oracle_capsule_box_and_mpr_agree_on_a_face_contact():
    """A finger capsule pressed flat onto a face of the box object: analytic capsule-box gives 2 end contacts with the
    face normal; the same pose against the convex-mesh bottle bottom checks MPR depth against plane geometry."""
    env = oracle_env("box")
    env.reset()
    M = model("box").compiled()
    g = M["names_geom"].index("robot0:C_FF1")
    q = env.get("qpos"); q[0:3] = [0, 0, 1.0]; q[6:26] = 0
    env.set_state(q, np.zeros(32))
    gp, gm = env.get("geom_xpos").reshape(-1, 3)[g], env.get("geom_xmat").reshape(-1, 3, 3)[g]
    r, hl = M["geom_size"][g][0], M["geom_size"][g][1]
    axis = gm[:, 2]
    # orthonormal frame with x along the capsule axis; put the box's +z face 0.5 mm into the capsule from "below" (-n)
    n = np.cross(axis, [0.3, 0.5, 0.8]); n /= np.linalg.norm(n)
    y = np.cross(n, axis)
    Rb = np.stack([axis, y, n], axis=1)            # box axes: x along capsule, z along n
    pen = 5e-4
    centre = gp - n * (r - pen + 0.049)
    q[26:29] = centre; q[29:33] = mjcf.mat2quat(Rb)
    env.set_state(q, np.zeros(32))
    c = env.contacts(); c = c[(c[:, 13] == g) & (c[:, 14] == 21)]
    assert len(c) == 2 and np.allclose(c[:, 0], -pen, atol=1e-9)
    assert np.allclose(c[:, 4:7], -n, atol=1e-9)                      # from capsule (geom1) towards the box
    ends = np.sort((c[:, 1:4] - gp) @ axis)
    assert np.allclose(ends, [-min(hl, 0.0165), min(hl, 0.0165)], atol=1e-9)


def test_oracle_capsule_box_axis_through_the_box():
    """A finger capsule whose axis passes through the box interior (a kinematic expert pose can do that): one contact at
    the midpoint of the inside stretch, normal of the face nearest to it, stable under tiny perturbations (the entry
    point of the axis lies on a face; picking it made the result depend on rounding)."""
    env = oracle_env("box")
    env.reset()
    M = model("box").compiled()
    g = M["names_geom"].index("robot0:C_FF1")
    q = env.get("qpos"); q[0:3] = [0, 0, 1.0]; q[6:26] = 0
    env.set_state(q, np.zeros(32))
    gp, gm = env.get("geom_xpos").reshape(-1, 3)[g], env.get("geom_xmat").reshape(-1, 3, 3)[g]
    r = M["geom_size"][g][0]
    axis = gm[:, 2]
    n = np.cross(axis, [0.3, 0.5, 0.8]); n /= np.linalg.norm(n)
    y = np.cross(n, axis)
    Rb = np.stack([axis, y, n], axis=1)            # box axes: x along the capsule, z along n
    out = []
    for eps in (0.0, 1e-9, -1e-9):
        centre = gp - n * (0.049 - 0.004) + y * (0.003 + eps)     # axis 4 mm below the +z face, 3 mm off the middle in y
        q[26:29] = centre; q[29:33] = mjcf.mat2quat(Rb)
        env.set_state(q, np.zeros(32))
        c = env.contacts(); c = c[(c[:, 13] == g) & (c[:, 14] == 21)]
        assert len(c) == 1
        out.append(c[0])
    c = out[0]
    assert np.allclose(c[4:7], -n, atol=1e-9) and abs(c[0] + (0.004 + r)) < 1e-9      # nearest face is +z: depth 4 mm + radius
    assert abs((c[1:4] - gp) @ axis) < 1e-9                                          # at the middle of the inside stretch
    assert all(np.abs(o - c).max() < 1e-6 for o in out[1:])


def test_oracle_mpr_depth_on_mesh_face():
    env = oracle_env("bottle")
    env.reset()
    M = model("bottle").compiled()
    # bottle upright, pushed 0.3 mm into the table: box-mesh MPR contact with normal +z and that depth
    q = env.get("qpos"); q[0:3] = [0, 0, 1.5]
    ex = model("bottle").motion(1)
    q[26:33] = ex["obj_pose_seq"][0]; q[28] -= 3e-4 + 2e-4
    env.set_state(q, np.zeros(32))
    c = env.contacts(); c = c[c[:, 13] == 1]
    assert len(c) >= 1
    assert np.allclose(c[:, 4:7], [0, 0, 1], atol=1e-6) and np.allclose(c[:, 0], -3e-4, atol=2e-6)


# ------------------------------------------------------------------------------------------------ oracle vs transcription
@pytest.mark.parametrize("obj", OBJECTS)
def test_oracle_obs_reward_vs_numpy_transcription(obj):
    env = oracle_env(obj)
    mdl = model(obj)
    M, ex = mdl.compiled(), mdl.motion(1)
    ws = dict(mdl.cfg.reward_weights)
    rng = np.random.RandomState(5)
    env.reset()
    t = 0
    for step in range(12):
        o, r, done, info = env.step(rng.randn(32) * 0.15)
        t += 1
        view = R.EnvView(env.get("qpos"), env.get("qvel"), env.get("xpos").reshape(-1, 3), env.get("xquat").reshape(-1, 4), ex, t,
                         M["env"]["hand_body_idx"], M["env"]["palm_body"], rfc_score=info["rfc_score"])
        assert np.abs(R.get_full_obs_v5(view) - o).max() < 1e-12
        rr, ri = R.ho_mimic_reward_9(view, ws)
        assert abs(rr - r) < 1e-12 and np.abs(ri - info["reward_info"]).max() < 1e-12
        d = R.calc_ho_diff(view)
        assert np.abs(d - env.ho_diff()).max() < 1e-12
        body_fail = d[0] > 0.1 or d[1] > 1.0 or d[2] > 0.1 or d[3] > 0.1 or d[4] > 1.0
        assert info["fail"] == (body_fail or env.counters()["warn_bad"] == 1)
        assert info["end"] == (t >= 300 - 5 - 1)
        if done:
            break


def test_oracle_rfc_qp_vs_bounded_lbfgs():
    """The active-set QP of the oracle reaches the optimum of the reference's QP (ho_im4.py:1063-1068)."""
    rng = np.random.RandomState(3)
    import ctypes as C
    from oracle import lib
    L = lib()
    for trial in range(4):
        k = 1 + trial
        frames = []
        for i in range(k):
            n = rng.randn(3); n /= np.linalg.norm(n)
            t = np.cross(n, [1, 0, 0]); t /= np.linalg.norm(t)
            frames.append(np.concatenate([rng.randn(3) * 0.03 + [0, 0, 0.6], n, t, np.cross(n, t)]))
        F, T = rng.randn(3) * 0.3, rng.randn(3) * 0.005
        ts = rng.randint(1, 30, size=k).astype(float)
        gx, gv, ga = rng.randn(k, 3) * 0.02 + [0, 0, 0.6], rng.randn(k, 3) * 0.05, rng.randn(k, 3)
        obj_p, obj_v, obj_w = np.array([0, 0, 0.6]) + rng.randn(3) * 0.01, rng.randn(3) * 0.05, rng.randn(3)
        score, fval, Q, p = R.solve_rfc_numpy(F, T, frames, ts, gx, gv, ga, obj_p, obj_v, obj_w)
        n = Q.shape[0]
        x = np.zeros(n)
        L.qp_nonneg_test.restype = C.c_int
        Qc, pc = np.ascontiguousarray(Q), np.ascontiguousarray(p)
        L.qp_nonneg_test(C.c_int(n), Qc.ctypes.data_as(C.c_void_p), pc.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p))
        f_or = 0.5 * x @ Q @ x + p @ x
        assert (x >= 0).all() and f_or <= fval + 1e-9 * (1 + abs(fval))
        g = Q @ x + p                                         # KKT: g >= 0, g_i x_i = 0
        assert g.min() > -1e-8 * (1 + np.abs(p).max()) and np.abs(g * x).max() < 1e-7 * (1 + np.abs(p).max())


# ------------------------------------------------------------------------------------------------ goldens
@pytest.mark.parametrize("obj", OBJECTS)
def test_oracle_reproduces_goldens(obj):
    G = np.load(os.path.join(ROOT, "tests", "golden", f"{obj}.npz"))
    env = oracle_env(obj)
    assert np.abs(env.reset() - G["reset_obs"]).max() < 1e-9
    for t in range(len(G["actions"])):
        o, r, d, i = env.step(G["actions"][t])
        assert np.abs(o - G["obs"][t]).max() < 1e-7 and abs(r - G["reward"][t]) < 1e-8
        assert [int(d), int(i["fail"]), int(i["end"])] == G["flags"][t].tolist()
    for k in range(0, len(G["sub_q0"]), 6):
        env.set_state(G["sub_q0"][k], G["sub_v0"][k])
        env.set_warmstart(G["sub_w0"][k])
        env.set_ctrl_applied(G["sub_ctrl"][k], G["sub_qapp"][k])
        env.clear_warnings()
        env.mj_step()
        assert np.abs(env.get("qvel") - G["sub_v1"][k]).max() < 1e-8 and env.counters()["ncon"] == G["sub_ncon"][k]


def test_oracle_preprocess_seq_vs_numpy_derivation():
    """orc_preprocess_seq (the checker of hoic_preprocess_seq) against the independent numpy restatement of
    preprocess_seq in hoic_b200/motion.py (dataset_singledepth.py:78-237), with out-of-range dofs and a wrap of a rotation dof."""
    import oracle
    mdl = model("bottle")
    ex = mdl.motion(2)
    hand, obj = ex["hand_dof_seq"].copy(), ex["obj_pose_seq"].copy()
    hand[:, 7] += 3.0                      # beyond the joint range: clipped
    hand[100:, 4] += 2 * np.pi - 0.01      # a jump through +-pi on a rotation dof: wrapped finite difference
    M = mdl.M if mdl.M else mdl.compiled
    fr, cm, cp = oracle.preprocess_seq(mdl.blob, hand, obj, 30)
    ref = hmotion.pack_frames(hmotion.derive_expert(M, hand, obj, 30))
    assert np.abs(fr - ref).max() < 1e-12
    lo, hi = mdl.arrays["jnt_range"][:-1, 0], mdl.arrays["jnt_range"][:-1, 1]
    assert (fr[:, :26] >= lo - 1e-15).all() and (fr[:, :26] <= hi + 1e-15).all()
    assert np.abs(fr[:, 26 + 4]).max() < 30 * np.pi          # no 2 pi jump in the rotation velocity
    # contact pre-pass on the unperturbed motion: nothing touches during the approach, several hand geoms in the grasp
    _, cm, cp = oracle.preprocess_seq(mdl.blob, ex["hand_dof_seq"], obj, 30)
    assert cm[:10].sum() == 0 and cm[150:].sum(1).min() >= 5
    assert (np.abs(cp[cm == 0]) == 0).all() and np.abs(cp[cm == 1] - obj[np.nonzero(cm)[0], :3]).max() < 0.25
    raw, _, _ = oracle.preprocess_seq(mdl.blob, hand, obj, 30, clip=False)
    assert np.abs(raw[:, 7] - hand[:, 7]).max() == 0
    with pytest.raises(ValueError):
        oracle.preprocess_seq(mdl.blob, hand[:1], obj[:1], 30)


def test_oracle_wrench_qp_known_answers():
    """The standalone wrench QP (solve_rfc's QP / PhysMetrics.solve_force): closed-form cases."""
    import oracle

    def frame(n):
        n = np.array(n, float) / np.linalg.norm(n)
        t = np.cross(n, [1.0, 0, 0]) if abs(n[0]) < 0.9 else np.cross(n, [0, 1.0, 0])
        t /= np.linalg.norm(t)
        return np.concatenate([n, t, np.cross(n, t)])
    W = [0, 0, 0.5, 0, 0, 0]                      # hold 0.5 N against gravity, no torque, about the origin
    pinch = np.array([np.concatenate([[0, -0.02, 0], frame([0, 1, 0])]), np.concatenate([[0, 0.02, 0], frame([0, -1, 0])])])
    rest, score, pf = oracle.solve_wrench_qp(pinch, W, [0, 0, 0], mu=1.0, w_t=1.0)
    assert score < 1e-7 and np.allclose(pf.sum(0), [0, 0, 0.5], atol=1e-7)           # friction of the two pads carries it
    rest, score, _ = oracle.solve_wrench_qp(pinch, W, [0, 0, 0], mu=0.1, w_t=1.0)
    assert score < 1e-6                                                              # ... at any mu > 0 by squeezing harder
    below = np.array([np.concatenate([[0, 0, -0.02], frame([0, 0, 1])])])
    rest, score, pf = oracle.solve_wrench_qp(below, W, [0, 0, 0], mu=0.75, w_t=1e4)
    assert score < 1e-6 and np.allclose(pf.sum(0), [0, 0, 0.5], atol=1e-6)           # a support from below pushes up
    above = np.array([np.concatenate([[0, 0, 0.02], frame([0, 0, -1])])])
    rest, score, pf = oracle.solve_wrench_qp(above, W, [0, 0, 0], mu=0.75, w_t=1e4)
    assert np.allclose(rest, W, atol=1e-9) and abs(score - 0.5) < 1e-9 and np.abs(pf).max() < 1e-9   # contacts cannot pull
    rest, score, _ = oracle.solve_wrench_qp(np.zeros((0, 12)), [0.3, 0, 0.4, 0, 0, 0.01], [0, 0, 0], mu=0.75, w_t=1e4)
    assert abs(score - (0.5 + 100 * 0.01)) < 1e-12                                   # no contact: |F| + sqrt(w_t) |T|
    # an off-centre support leaves a torque the single pad cannot cancel exactly: the weighted residual trades force for torque
    off = np.array([np.concatenate([[0.03, 0, -0.02], frame([0, 0, 1])])])
    r1, s1, _ = oracle.solve_wrench_qp(off, W, [0, 0, 0], mu=0.75, w_t=1.0, surface_contact=False)
    r2, s2, _ = oracle.solve_wrench_qp(off, W, [0, 0, 0], mu=0.75, w_t=1e4, surface_contact=False)
    assert np.linalg.norm(r2[3:]) < np.linalg.norm(r1[3:]) and np.linalg.norm(r2[:3]) > np.linalg.norm(r1[:3])


def test_tracking_window_layout_matches_the_window_relative_accessors():
    """HandObjMimicTest (ho_im_test.py:26-91) looks the expert up relative to a sliding window, whatever cur_t is.  The
    7-frame layouts hoic_b200/track.py feeds the ordinary kernels must give exactly that: checked on the oracle against
    the numpy transcription of get_full_obs_v5 / calc_ho_diff reading a 6-frame window at cur_t = 0."""
    from hoic_b200.dataset import unpack_frames
    mdl = model("box")
    M, ex = mdl.compiled(), mdl.motion(1)
    frames = hmotion.pack_frames(ex)                       # [300, 212]
    env = oracle_env("box")
    rng = np.random.RandomState(2)
    t0 = 40
    window = frames[t0:t0 + 6].copy()
    view_of = lambda w: unpack_frames(np.asarray(w, dtype=np.float64))
    # reset: layout [w0..w5, w5], cur_t = 0
    env.set_motion(np.concatenate([window, window[-1:]]))
    o = env.reset()
    v = R.EnvView(env.get("qpos"), env.get("qvel"), env.get("xpos").reshape(-1, 3), env.get("xquat").reshape(-1, 4), view_of(window), 0,
                  M["env"]["hand_body_idx"], M["env"]["palm_body"])
    assert np.abs(env.get("qpos")[:26] - window[0, :26]).max() == 0 and np.abs(R.get_full_obs_v5(v) - o).max() < 1e-12
    for step in range(5):
        # step: layout [w0, w0..w5], cur_t = 0; afterwards cur_t = 1 and every lookup is relative to the SAME window
        env.set_motion(np.concatenate([window[:1], window]))
        env.cur_t = 0
        o, r, done, info = env.step(rng.randn(32) * 0.1)
        v = R.EnvView(env.get("qpos"), env.get("qvel"), env.get("xpos").reshape(-1, 3), env.get("xquat").reshape(-1, 4), view_of(window), 0,
                      M["env"]["hand_body_idx"], M["env"]["palm_body"], rfc_score=info["rfc_score"])
        # the transcription reads window[cur_t + 1 + idx] = window[1 + idx] for the targets and window[0] for the errors
        v_obs = R.EnvView(v.qpos, v.qvel, v.body_xpos, v.body_xquat, view_of(window), 0, v.hand_body_idx, v.palm_body)
        assert np.abs(R.get_full_obs_v5(v_obs) - o).max() < 1e-12
        assert np.abs(R.calc_ho_diff(v) - env.ho_diff()).max() < 1e-12
        # the tracker delivers the next frame: insert_new_frame (ho_im_test.py:34-38)
        window = np.concatenate([window[1:], frames[t0 + 6 + step][None]])


def test_oracle_spinning_object_friction_loss_and_quaternion_integration():
    """Object spinning about a principal axis in free fall: the free joint's 0.001 N m friction loss decelerates it at
    fl / (I + armature); semi-implicit Euler with the quaternion exponential map gives the rotation angle in closed form."""
    env = oracle_env("box")
    env.reset()
    A = model("box").arrays
    h, I, arm, fl = A["opt_timestep"][0], A["body_inertia"][24][2], A["dof_armature"][31], A["dof_frictionloss"][31]
    q = env.get("qpos"); q[0:3] = [0, 0, 1.5]; q[26:29] = [0.5, 0.5, 3.0]; q[29:33] = [1, 0, 0, 0]
    v = np.zeros(32); v[31] = 20.0
    env.set_state(q, v); env.set_ctrl_applied(np.zeros(26), np.zeros(32))
    n = 50
    for _ in range(n):
        env.mj_step()
    a = fl / (I + arm)
    w = env.get("qvel")[26:32]
    assert abs(w[5] - (20.0 - a * h * n)) < 1e-9 and np.abs(w[3:5]).max() < 1e-12          # no precession about a principal axis
    theta = h * (n * 20.0 - a * h * n * (n + 1) / 2)                                       # sum of the updated velocities
    quat = env.get("qpos")[29:33]
    assert np.abs(quat - [np.cos(theta / 2), 0, 0, np.sin(theta / 2)]).max() < 1e-9 and abs(np.linalg.norm(quat) - 1) < 1e-12
    # about a non-principal direction the gyroscopic term enters: angular momentum R I R^T w only changes by the friction torque
    v[29:32] = [8.0, -5.0, 12.0]
    env.set_state(q, v)
    Ib = A["body_inertia"][24]
    L0 = Ib * v[29:32]                                                                      # body frame = world frame at t = 0
    env.mj_step()
    w1, q1 = env.get("qvel")[29:32], env.get("qpos")[29:33]
    Rm = mjcf.quat2mat(q1) if hasattr(mjcf, "quat2mat") else R.quat2mat(q1)
    # free-joint angular velocity is expressed in the body frame (MuJoCo convention)
    L1 = Rm @ (Ib * w1)
    assert np.abs(L1 - L0).max() < (fl + arm * np.abs(w1 - v[29:32]).max() / h) * h * 3 + 1e-7


def test_frame_table_roundtrip_and_window_packers():
    """[T, 212] frame table <-> expert dict (hoic_b200/dataset.py, motion.py) and the tracking env's frame packers."""
    from hoic_b200.dataset import unpack_frames
    from hoic_b200.track import _pack
    ex = model("box").motion(1)
    fr = hmotion.pack_frames(ex)
    back = unpack_frames(fr)
    for k in hmotion.EXPERT_KEYS:
        assert back[k].shape == ex[k].shape and np.array_equal(back[k], ex[k])
    assert back["seq_len"] == ex["seq_len"] and np.array_equal(hmotion.pack_frames(back), fr)
    one = {k: ex[k][7] for k in hmotion.EXPERT_KEYS}                 # a single frame as the reference's insert_new_frame gets it
    assert np.allclose(_pack([one, one]), fr[7][None].repeat(2, 0)) and _pack(fr[3:9]).shape == (6, 212)
    assert np.allclose(_pack({k: ex[k][3:9] for k in hmotion.EXPERT_KEYS}), fr[3:9])


def test_oracle_joint_limit_row_follows_mujocos_impedance_formulas():
    """A limited joint inside its margin: the row's D and aref against MuJoCo 2.1's published formulas (mj_makeImpedance:
    K = 1 / (dmax^2 tc^2 dr^2), B = 2 / (dmax tc); getimpedance: power-p sigmoid between d0 and dmax over `width`;
    R = (1 - imp) / imp * dof_invweight0; aref = -B vel - K imp (pos - margin)), on both sigmoid branches and saturated."""
    env = oracle_env("box")
    env.reset()
    A = model("box").arrays
    j = 10
    lo, hi = A["jnt_range"][j]
    margin, (tc, dr), (d0, dmax, width, mid, power) = A["jnt_margin"][j], A["jnt_solref"][j], A["jnt_solimp"][j]
    iw = A["dof_invweight0"][j]
    q0 = env.get("qpos"); q0[0:3] = [0, 0, 1.5]; q0[26:29] = [2, 2, 3]
    for dist, vel in [(-0.004, 0.3), (margin - 0.0003, -0.2), (margin - 0.0007, 0.0), (margin - 0.00099, 0.1)]:
        q = q0.copy(); q[j] = hi - dist                    # upper limit: pos = range_hi - q
        v = np.zeros(32); v[j] = vel
        env.set_state(q, v)
        nefc = env.counters()["nefc"]
        pos, D, aref, J = env.get("efc_pos")[32:nefc], env.get("efc_D")[32:nefc], env.get("efc_aref")[32:nefc], env.efc_J()[32:nefc]
        r = int(np.argmin(np.abs(pos - dist)))
        assert abs(pos[r] - dist) < 1e-12 and J[r, j] == -1.0 and np.count_nonzero(J[r]) == 1
        x = min(abs(dist - margin) / width, 1.0)
        y = x ** power / mid ** (power - 1) if x <= mid else 1 - (1 - x) ** power / (1 - mid) ** (power - 1)
        imp = d0 + y * (dmax - d0)
        K, B = 1.0 / (dmax * dmax * tc * tc * dr * dr), 2.0 / (dmax * tc)
        assert abs(D[r] - imp / ((1 - imp) * iw)) < 1e-12 * D[r]
        assert abs(aref[r] - (-B * (-vel) - K * imp * (dist - margin))) < 1e-10 * max(1.0, abs(aref[r]))


def _seg_seg_closest(p1, d1, h1, p2, d2, h2):
    """Closest points of two segments p_i + t d_i, |t| <= h_i (unit d_i), by brute force over the candidate set of the
    standard clamped solution (independent of the oracle's implementation)."""
    best = None
    A = np.array([[1.0, -(d1 @ d2)], [-(d1 @ d2), 1.0]])
    rhs = np.array([(p2 - p1) @ d1, (p1 - p2) @ d2])
    cands = []
    if abs(np.linalg.det(A)) > 1e-12:
        t = np.linalg.solve(A, rhs)
        cands.append((np.clip(t[0], -h1, h1), np.clip(t[1], -h2, h2)))
    for s in (-h1, h1):
        cands.append((s, np.clip((p1 + s * d1 - p2) @ d2, -h2, h2)))
    for u in (-h2, h2):
        cands.append((np.clip((p2 + u * d2 - p1) @ d1, -h1, h1), u))
    for s, u in cands:
        for _ in range(3):                                  # alternate projections: converges for segments
            u = np.clip((p1 + s * d1 - p2) @ d2, -h2, h2)
            s = np.clip((p2 + u * d2 - p1) @ d1, -h1, h1)
        a, b = p1 + s * d1, p2 + u * d2
        if best is None or np.linalg.norm(a - b) < best[0]:
            best = (np.linalg.norm(a - b), a, b)
    return best


def test_oracle_capsule_capsule_pairs_vs_segment_distance():
    """The explicit finger-finger <pair>s (condim 1, capsule-capsule): in random hand poses every reported contact has
    dist = segment distance - r1 - r2, the normal from geom1 to geom2 and the position midway between the surfaces, and no
    penetrating pair goes unreported (mjc_CapsuleCapsule restated; margin = gap = 0 so contact iff penetrating)."""
    env = oracle_env("box")
    env.reset()
    A = model("box").arrays
    rng = np.random.RandomState(7)
    lo, hi = A["jnt_range"][:26, 0], A["jnt_range"][:26, 1]
    pairs = [(int(a), int(b)) for a, b in zip(A["pair_geom1"], A["pair_geom2"]) if A["geom_type"][int(a)] == 3 and A["geom_type"][int(b)] == 3]
    seen = 0
    for trial in range(40):
        q = env.get("qpos")
        q[0:3] = [0, 0, 1.5]; q[3:6] = rng.randn(3) * 0.3; q[26:29] = [2, 2, 3]
        q[6:26] = lo[6:] + rng.rand(20) * (hi[6:] - lo[6:])
        env.set_state(q, np.zeros(32))
        gp, gm = env.get("geom_xpos").reshape(-1, 3), env.get("geom_xmat").reshape(-1, 3, 3)
        c = env.contacts()
        reported = {(int(r[13]), int(r[14])): r for r in c if int(r[15]) == 1}
        for g1, g2 in set(pairs):
            r1, h1, r2, h2 = A["geom_size"][g1][0], A["geom_size"][g1][1], A["geom_size"][g2][0], A["geom_size"][g2][1]
            d, a, b = _seg_seg_closest(gp[g1], gm[g1][:, 2], h1, gp[g2], gm[g2][:, 2], h2)
            dist = d - r1 - r2
            if (g1, g2) in reported:
                row = reported[(g1, g2)]
                n = (b - a) / d
                assert abs(row[0] - dist) < 1e-9 and dist < 0, (g1, g2, row[0], dist)
                assert np.abs(row[4:7] - n).max() < 1e-7
                assert np.abs(row[1:4] - 0.5 * ((a + n * r1) + (b - n * r2))).max() < 1e-9
                seen += 1
            else:
                assert dist > -1e-9, (g1, g2, dist)
    assert seen > 15


def test_oracle_apply_ft_is_the_transposed_point_jacobian():
    """mj_applyFT (gravity compensation and the residual wrench enter the dynamics through it, ho_im4.py:488-501, 526-536):
    qfrc = J_p^T f + J_r^T tau with the point Jacobian of a body-fixed point, checked against finite differences of the
    forward kinematics for the 26 scalar hand joints and against virtual work for the object's free joint."""
    import ctypes as C
    import oracle
    env = oracle_env("box")
    env.reset()
    L = oracle.lib()
    L.orc_env_data.restype = C.c_void_p
    L.orc_apply_ft.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    A = model("box").arrays
    rng = np.random.RandomState(9)
    q = env.get("qpos")
    q[6:26] += rng.randn(20) * 0.1
    q[3:6] = rng.randn(3) * 0.4
    g = 12                                             # a finger capsule; its centre is a body-fixed point
    body = int(A["geom_bodyid"][g])

    def point(qq):
        env.set_state(qq, np.zeros(32))
        return env.get("geom_xpos").reshape(-1, 3)[g].copy()
    eps = 1e-6
    Jfd = np.zeros((3, 26))
    for i in range(26):
        dq = np.zeros(33); dq[i] = eps
        Jfd[:, i] = (point(q + dq) - point(q - dq)) / (2 * eps)
    p = point(q)
    data = C.c_void_p(L.orc_env_data(env.env))
    P = lambda a: np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(C.c_void_p)
    for k in range(3):
        f = np.zeros(3); f[k] = 1.0
        qfrc = np.zeros(32)
        L.orc_apply_ft(C.c_void_p(env.model), data, P(f), P(np.zeros(3)), P(p), body, qfrc.ctypes.data_as(C.c_void_p))
        assert np.abs(qfrc[:26] - Jfd[k]).max() < 1e-7 and not qfrc[26:].any()      # the object's dofs do not move a finger
    # object (free joint): force at an offset point -> translational dofs get f (world frame), rotational dofs r x f in the BODY frame
    ob = 24
    r_world = np.array([0.01, -0.02, 0.015])
    f = np.array([0.3, -0.2, 0.5])
    env.set_state(q, np.zeros(32))
    op = q[26:29]
    Rm = R.quat2mat(q[29:33] / np.linalg.norm(q[29:33]))
    qfrc = np.zeros(32)
    L.orc_apply_ft(C.c_void_p(env.model), data, P(f), P(np.zeros(3)), P(op + r_world), ob, qfrc.ctypes.data_as(C.c_void_p))
    assert np.abs(qfrc[26:29] - f).max() < 1e-12 and np.abs(qfrc[29:32] - Rm.T @ np.cross(r_world, f)).max() < 1e-12
    assert not qfrc[:26].any()


@pytest.mark.parametrize("obj", ["box", "banana"])
def test_oracle_stable_pd_vs_numpy_transcription(obj):
    """compute_torque / compute_desired_accel (ho_im4.py:393-486) + torque clip (:521): the oracle's stable PD on the STALE
    mass matrix and bias of the last forward against the literal transcription with scipy's Cholesky."""
    env = oracle_env(obj)
    mdl = model(obj)
    A, ex = mdl.arrays, mdl.motion(1)
    rng = np.random.RandomState(12)
    env.reset()
    dt = A["opt_timestep"][0]
    clipped = 0
    for t in range(1, 9):
        env.step(rng.randn(32) * 0.2)
        for scale in (0.3, 3.0):                            # the larger one drives some torques into the clip
            a = np.clip(rng.randn(32) * scale, -1, 1)
            got = env.compute_torque(a)
            want = R.compute_torque_numpy(env.get("qpos")[:26], env.get("qvel")[:26], ex["hand_dof_seq"][t], a, env.qM()[:26, :26],
                                          env.get("qfrc_bias")[:26], A["env_jkp"][:26], A["env_jkd"][:26], A["env_ctrl_scale"][:26],
                                          A["env_torque_lim"][:26], dt)
            assert np.abs(got - want).max() < 1e-9 * max(1.0, np.abs(want).max())
            clipped += int((np.abs(want) == A["env_torque_lim"][:26]).sum())
    assert clipped > 0


def test_oracle_reproduces_extras_goldens():
    """Regression fixtures of the widened rows (tools/make_golden.py extras_fixture): test-mode record, wrench QPs, preprocessing."""
    import oracle
    G = np.load(os.path.join(ROOT, "tests", "golden", "extras.npz"))
    env = oracle_env("box", start_idx=100)
    env.set_mode("test"); env.set_record(True); env.reset()
    for t in range(len(G["rec_actions"])):
        env.step(G["rec_actions"][t])
        fr, su = env.record()
        assert np.abs(fr - G["rec_frames"][t]).max() < 1e-7 and np.abs(su - G["rec_surface"][t]).max() < 1e-6
    assert G["rec_frames"][:, :, 114].max() >= 1 and G["rec_surface"][:, 7].max() == 1      # the fixture does contain contacts
    for k in range(len(G["qp_n"])):
        nc = int(G["qp_n"][k])
        rest, score, _ = oracle.solve_wrench_qp(G["qp_contacts"][k][:nc], G["qp_target"][k], G["qp_center"][k], mu=0.75, w_t=1e4)
        assert np.abs(rest - G["qp_rest"][k]).max() < 1e-9 and abs(score - G["qp_score"][k]) < 1e-9
    mdl = model("bottle")
    ex = mdl.motion(2)
    fr, cm, cp = oracle.preprocess_seq(mdl.blob, ex["hand_dof_seq"][::10], ex["obj_pose_seq"][::10], 3.0)
    assert np.abs(fr - G["pre_frames"]).max() < 1e-10 and np.array_equal(cm, G["pre_cmask"]) and np.abs(cp - G["pre_cpos"]).max() < 1e-9


def test_single_env_facade_attribute_surface_with_a_fake_engine():
    """hoic_b200/env_facade.py: the HandObjMimic4 attribute surface the reference's callers read, exercised on
    the CPU with a stand-in that has HandMimicVecEnv's method signatures (the façade only forwards to reset / step / get_state /
    get_derived / get_rfc / set_motions / set_mode, each covered on the GPU by tests/test_gpu_parity.py)."""
    import torch
    from hoic_b200.env_facade import HandObjMimic4, ho_mimic_reward_9
    mdl = model("box")
    ex = mdl.motion(1)

    class Fake:
        def __init__(self):
            self.mode, self.t, self.motions = None, 0, None
        def set_mode(self, m): self.mode = m
        def set_motions(self, ms): self.motions = ms
        def reset(self, motion_id=None, start_idx=None):
            self.t = 0
            return torch.full((1, 617), 0.5)
        def step(self, a):
            assert a.shape == (1, 32) and a.dtype == torch.float32
            self.t += 1
            info = dict(fail=torch.tensor([0], dtype=torch.uint8), end=torch.tensor([1 if self.t >= 2 else 0], dtype=torch.uint8),
                        percent=torch.tensor([self.t / 299.0]), reward_info=torch.arange(9, dtype=torch.float32)[None] * 0.1)
            return torch.full((1, 617), float(self.t)), torch.tensor([0.25 * self.t]), info["end"].clone(), info
        def get_state(self): return torch.arange(33, dtype=torch.float32)[None], torch.arange(32, dtype=torch.float32)[None] * 2
        def get_derived(self): return torch.arange(75, dtype=torch.float32).reshape(1, 25, 3), torch.ones(1, 25, 4), None, None
        def get_rfc(self): return torch.tensor([[0.75, 0, 0, 0, 0, 0, 0, 1.0]])
        def close(self): pass
    fake = Fake()
    env = HandObjMimic4(mdl.cfg, ex, mdl, mode="test", _vecenv=fake)
    assert fake.mode == "test" and env.observation_space.shape[0] == 617 and env.action_space.shape[0] == 32
    assert (env.ndof, env.vf_dim, env.hand_qpos_dim, env.expert_len, env.start_ind, env.cur_t) == (26, 6, 26, 300, 0, 0)
    o = env.reset()
    assert o.dtype == np.float64 and o.shape == (617,)
    o, r, done, info = env.step(np.zeros(32))
    assert r == 1.0 and done is False and info["fail"] is False and info["end"] is False and abs(info["percent"] - 1 / 299.0) < 1e-7
    rew, ri = ho_mimic_reward_9(env, o, np.zeros(32), info)
    assert rew == 0.25 and np.allclose(ri, np.arange(9) * 0.1) and env.rfc_score == 0.75 and env.cur_t == 1
    _, _, done, info = env.step(np.zeros(26))           # a caller without residual force entries: padded
    assert done is True and info["end"] is True
    assert np.array_equal(env.get_hand_qpos(), np.arange(26.0)) and np.array_equal(env.get_obj_qpos(), np.arange(26.0, 33.0))
    assert np.array_equal(env.get_obj_qvel(), np.arange(26.0, 32.0) * 2) and env.data.qpos.shape == (33,)
    assert env.get_wbody_pos().shape == (63,) and env.get_wbody_quat().shape == (84,)
    assert np.array_equal(env.get_wbody_pos()[:3], np.arange(75.0).reshape(25, 3)[env.hand_body_idx[0]])
    assert np.array_equal(env.get_expert_attr("hand_dof_seq", 10_000), ex["hand_dof_seq"][-1])          # clamp (ho_im4.py:740)
    assert np.array_equal(env.get_expert_hand_qpos(delta_t=1), ex["hand_dof_seq"][3]) and env.get_expert_obj_vel().shape == (6,)
    env.set_expert({k: (v[50:] if hasattr(v, "shape") else v) for k, v in ex.items()})
    assert env.expert_len == 250 and env.cur_t == 0 and fake.motions is not None
    env.set_mode("train")
    assert fake.mode == "train" and env.seed(3) == [3]


def test_oracle_reproduces_the_large_state_sets():
    """tests/golden/<obj>_states.npz (tools/make_golden_states.py): 1280 single-mj_step vectors per object.  Oracle regression
    on every 16th state (the CPU suite stays short): contact pairs, nefc, next qpos / qvel."""
    import os
    for obj in ("box", "bottle", "banana"):
        G = np.load(os.path.join(ROOT, "tests", "golden", f"{obj}_states.npz"))
        assert len(G["sub_q0"]) >= 1280 and int((G["sub_kind"] == 1).sum()) >= 256
        assert int((G["sub_ncon"] >= 12).sum()) >= 100 and int(G["sub_nefc"].max()) >= 100
        env = oracle_env(obj)
        env.reset()
        for i in range(0, len(G["sub_q0"]), 16):
            env.clear_warnings()
            env.set_state(G["sub_q0"][i], G["sub_v0"][i])
            env.set_ctrl_applied(G["sub_ctrl"][i], G["sub_qapp"][i])
            env.set_warmstart(G["sub_w0"][i])
            env.mj_step()
            con = env.contacts()
            assert len(con) == G["sub_ncon"][i] and env.counters()["nefc"] == G["sub_nefc"][i]
            assert np.array_equal(con[:, 13:15].astype(int), G["sub_pairs"][i][:len(con)].astype(int))
            assert np.abs(env.get("qpos") - G["sub_q1"][i]).max() < 1e-12 and np.abs(env.get("qvel") - G["sub_v1"][i]).max() < 1e-10
        env.close()


def test_mesh_sdf_positive_inside_against_closed_forms():
    """hoic_b200.metrics.mesh_sdf (stands in for pysdf in PhysMetrics.eval_penetration, scripts/metrics.py:49-75): signed distance
    to a closed triangle mesh, positive inside.  Checked on a box (closed form) and on a finely tessellated sphere."""
    import torch
    from hoic_b200.metrics import mesh_sdf
    h = np.array([0.3, 0.2, 0.5])
    corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], dtype=np.float64) * h
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    tri = np.array([[corners[q[0]], corners[q[1]], corners[q[2]]] for q in quads] + [[corners[q[0]], corners[q[2]], corners[q[3]]] for q in quads])
    rng = np.random.RandomState(0)
    p = rng.uniform(-0.8, 0.8, (400, 3))
    d = np.abs(p) - h
    exact = -(np.linalg.norm(np.maximum(d, 0), axis=1) + np.minimum(d.max(1), 0))      # box SDF, positive inside
    got = mesh_sdf(torch.as_tensor(p), torch.as_tensor(tri)).numpy()
    assert np.abs(got - exact).max() < 1e-12 and (got > 0).sum() > 20 and (got < 0).sum() > 20
    # icosphere-like tessellation: |sdf - (r - |p|)| bounded by the sagitta of the facets
    n, r = 40, 0.4
    th, ph = np.linspace(0, np.pi, n + 1), np.linspace(0, 2 * np.pi, 2 * n + 1)
    P = lambda i, j: r * np.array([np.sin(th[i]) * np.cos(ph[j]), np.sin(th[i]) * np.sin(ph[j]), np.cos(th[i])])
    tris = []
    for i in range(n):
        for j in range(2 * n):
            a, b, c, e = P(i, j), P(i + 1, j), P(i + 1, j + 1), P(i, j + 1)
            if i > 0: tris.append([a, b, e])
            if i < n - 1: tris.append([b, c, e])
    q = rng.uniform(-0.6, 0.6, (200, 3))
    got = mesh_sdf(torch.as_tensor(q), torch.as_tensor(np.array(tris))).numpy()
    assert np.abs(got - (r - np.linalg.norm(q, axis=1))).max() < 2e-3 and ((got > 0) == (np.linalg.norm(q, axis=1) < r)).mean() > 0.98


@pytest.mark.parametrize("obj", ["box", "banana"])
def test_compiled_constants_vs_an_independent_route(obj):
    """mj_setConst constants of the model blob (`stat_meaninertia`, `dof_invweight0`, `body_invweight0`: they set R of every
    constraint row) recomputed WITHOUT hoic_b200.mjcf's mass_matrix / Jacobians: M at qpos0 from the oracle's C
    composite-rigid-body pass, body-CoM Jacobians by central finite differences of the oracle's forward kinematics."""
    mdl = model(obj)
    A = mdl.arrays
    nv, nb = 32, int(A["nbody"][0])
    env = oracle_env(obj)
    q0 = np.asarray(A["qpos0"], dtype=np.float64).copy()
    env.set_state(q0, np.zeros(nv))
    H = env.qM().reshape(nv, nv)
    Hinv = np.linalg.inv(H)
    assert abs(float(A["stat_meaninertia"][0]) - np.trace(H) / nv) < 1e-12
    dw = np.diag(Hinv).copy()
    dw[26:29] = dw[26:29].mean(); dw[29:32] = dw[29:32].mean()        # the object's free joint: translation / rotation averages
    assert np.abs(np.asarray(A["dof_invweight0"]) - dw).max() < 1e-9 * np.abs(dw).max()
    ipos = np.asarray(A["body_ipos"]).reshape(nb, 3)

    def quat2mat(q):
        w, x, y, z = q
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                         [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                         [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])

    def com_and_rot(q):
        env.set_state(q, np.zeros(nv))
        xp, xq = env.get("xpos").reshape(-1, 3)[:nb], env.get("xquat").reshape(-1, 4)[:nb]
        R = np.array([quat2mat(xq[b]) for b in range(nb)])
        return xp + np.einsum("bij,bj->bi", R, ipos), R

    def perturb(q, d, eps):
        q = q.copy()
        if d < 26:
            q[d] += eps
        elif d < 29:
            q[d] += eps
        else:                                   # rotation about the object's body axis d - 29
            ax = np.zeros(3); ax[d - 29] = 1.0
            dq = np.concatenate([[np.cos(eps / 2)], np.sin(eps / 2) * ax])
            w0, x0, y0, z0 = q[29:33]; w1, x1, y1, z1 = dq
            q[29:33] = [w0 * w1 - x0 * x1 - y0 * y1 - z0 * z1, w0 * x1 + x0 * w1 + y0 * z1 - z0 * y1,
                        w0 * y1 - x0 * z1 + y0 * w1 + z0 * x1, w0 * z1 + x0 * y1 - y0 * x1 + z0 * w1]
        return q

    eps = 1e-6
    Jp, Jr = np.zeros((nb, 3, nv)), np.zeros((nb, 3, nv))
    for d in range(nv):
        cp, Rp = com_and_rot(perturb(q0, d, eps))
        cm, Rm = com_and_rot(perturb(q0, d, -eps))
        Jp[:, :, d] = (cp - cm) / (2 * eps)
        dR = np.einsum("bij,bkj->bik", Rp, Rm)                        # R+ R-^T ~ I + [w]x * 2 eps
        Jr[:, :, d] = np.stack([dR[:, 2, 1] - dR[:, 1, 2], dR[:, 0, 2] - dR[:, 2, 0], dR[:, 1, 0] - dR[:, 0, 1]], 1) / (4 * eps)
    bw = np.asarray(A["body_invweight0"]).reshape(nb, 2)
    weld = np.asarray(A["body_weldid"])
    for b in range(1, nb):
        if weld[b] == 0:
            continue
        t = np.trace(Jp[b] @ Hinv @ Jp[b].T) / 3
        r = np.trace(Jr[b] @ Hinv @ Jr[b].T) / 3
        assert abs(bw[b, 0] - t) < 1e-5 * max(t, 1e-12) and abs(bw[b, 1] - r) < 1e-5 * max(r, 1e-12), (b, bw[b], t, r)
    env.close()
