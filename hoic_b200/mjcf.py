"""MJCF -> flat model arrays (host side, float64, run once).

Replaces, for the HandMimic model class only, what the reference obtains from
`mujoco_py.load_model_from_path` (uhc/khrylib/rl/envs/common/mujoco_env.py:15-34) after the
hand/object merge of `DatasetSingleDepth.add_obj` (uhc/data_loaders/dataset_singledepth.py:144-150,
uhc/data_loaders/mjxml/MujocoXML.py:72-106), plus the index discovery of
`HandObjMimic4.__init__` (uhc/envs/ho_im4.py:72-107).

Supported MJCF subset (everything the two asset families use): nested <default> classes with
`childclass`/`class`, bodies with pos/quat/euler, slide/hinge/free joints, explicit <inertial> or
inertia-from-geoms, plane/box/capsule/mesh geoms (binary STL, `scale`), explicit <contact><pair>,
motors with gear 1.  Anything else raises.

Output: a dict name -> numpy array (names follow mjModel where one exists), consumed by
`hoic_b200.blob.pack_model` and by both the CUDA library and the CPU oracle.
"""
from __future__ import annotations

import os
import struct
import xml.etree.ElementTree as ET

import numpy as np

GEOM_PLANE, GEOM_SPHERE, GEOM_CAPSULE, GEOM_BOX, GEOM_MESH = 0, 2, 3, 6, 7
JNT_FREE, JNT_SLIDE, JNT_HINGE = 0, 2, 3
MINVAL = 1e-15

# ----------------------------------------------------------------------------- quaternion helpers (wxyz)


def quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
    ])


def quat2mat(q):
    w, x, y, z = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
    ])


def mat2quat(R):
    """Rotation matrix -> unit quaternion (w >= 0 branch by largest pivot)."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = np.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s])
    elif R[1, 1] > R[2, 2]:
        s = np.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = np.array([(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s])
    else:
        s = np.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = np.array([(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s])
    return q / np.linalg.norm(q)


def axisangle2quat(axis, angle):
    s = np.sin(0.5 * angle)
    return np.array([np.cos(0.5 * angle), axis[0] * s, axis[1] * s, axis[2] * s])


def euler2quat_xyz(e):
    """MJCF default eulerseq="xyz": intrinsic rotations about x, then y', then z''."""
    q = np.array([1.0, 0, 0, 0])
    for k in range(3):
        ax = np.zeros(3)
        ax[k] = 1
        q = quat_mul(q, axisangle2quat(ax, e[k]))
    return q


def _floats(s, n=None):
    v = np.array([float(x) for x in s.split()], dtype=np.float64)
    if n is not None and v.size != n:
        raise ValueError(f"expected {n} numbers, got {s!r}")
    return v


# ----------------------------------------------------------------------------- XML merge (MujocoXML.merge semantics)

_MERGE_SECTIONS = ("actuator", "sensor", "tendon", "equality", "contact")


def merge_mjcf(hand_path: str, obj_path: str) -> ET.Element:
    """Hand MJCF + object MJCF -> one tree, following MujocoXML.merge (MujocoXML.py:72-106,150-161):
    every child of the object's <worldbody> is appended LAST to the hand's <worldbody> (after <light>),
    assets are appended unless an asset of the same tag+name exists, file paths are made absolute."""
    def load(path):
        root = ET.parse(path).getroot()
        folder = os.path.abspath(os.path.dirname(path))
        asset = root.find("asset")
        if asset is not None:
            for node in asset.findall("./*[@file]"):
                node.set("file", os.path.join(folder, node.get("file")))
        return root

    def section(root, name):
        e = root.find(name)
        if e is None:
            e = ET.SubElement(root, name)
        return e

    hand, obj = load(hand_path), load(obj_path)
    for body in list(section(obj, "worldbody")):
        section(hand, "worldbody").append(body)
    hasset = section(hand, "asset")
    for a in list(section(obj, "asset")):
        if not any(x.tag == a.tag and x.get("name") == a.get("name") for x in hasset):
            hasset.append(a)
    for name in _MERGE_SECTIONS:
        for e in list(section(obj, name)):
            section(hand, name).append(e)
    return hand


# ----------------------------------------------------------------------------- mesh processing


def load_stl(path: str, scale) -> np.ndarray:
    """Binary STL -> triangles [n,3,3] float64, scaled."""
    raw = open(path, "rb").read()
    n = struct.unpack("<I", raw[80:84])[0]
    if len(raw) < 84 + 50 * n:
        raise ValueError(f"{path}: not a binary STL")
    rec = np.frombuffer(raw[84:84 + 50 * n], dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]))
    return rec["v"].astype(np.float64) * np.asarray(scale, dtype=np.float64)


def mesh_props_legacy(tri: np.ndarray):
    """Volume, centre of mass and unit-density inertia about the CoM, by MuJoCo 2.1's ("legacy")
    construction: pyramids from the area-weighted centroid of the face centres, |volume| per face.
    For the closed convex collision hulls this equals the exact signed result."""
    a, b, c = tri[:, 0], tri[:, 1], tri[:, 2]
    nrm = np.cross(b - a, c - a)
    area2 = np.linalg.norm(nrm, axis=1)
    ok = area2 > 1e-30
    a, b, c, nrm, area2 = a[ok], b[ok], c[ok], nrm[ok], area2[ok]
    area = 0.5 * area2
    cen = (a + b + c) / 3.0
    facecen = (cen * area[:, None]).sum(0) / area.sum()
    # tetra (facecen, a, b, c): signed volume
    vol = np.einsum("ij,ij->i", np.cross(a - facecen, b - facecen), c - facecen) / 6.0
    vol = np.abs(vol)
    V = vol.sum()
    tcom = (a + b + c + facecen) / 4.0
    com = (tcom * vol[:, None]).sum(0) / V
    # second moments of each tetra about `com`: cov = V/20 * (sum_i p_i p_i^T + s s^T), p_i rel. to com
    P = np.stack([a - com, b - com, c - com, np.broadcast_to(facecen - com, a.shape)], axis=1)  # [n,4,3]
    s = P.sum(1)
    cov = (np.einsum("nki,nkj->nij", P, P) + np.einsum("ni,nj->nij", s, s)) * (vol / 20.0)[:, None, None]
    C = cov.sum(0)
    inertia = np.trace(C) * np.eye(3) - C
    return V, com, inertia


def principal_axes(I: np.ndarray):
    """Symmetric 3x3 -> (moments in decreasing order, right-handed rotation whose columns are the axes)."""
    w, U = np.linalg.eigh(I)
    order = np.argsort(-w, kind="stable")
    w, U = w[order], U[:, order]
    # keep identity when already diagonal & sorted (explicit boxes), make right-handed otherwise
    for k in range(3):
        j = np.argmax(np.abs(U[:, k]))
        if U[j, k] < 0:
            U[:, k] = -U[:, k]
    if np.linalg.det(U) < 0:
        U[:, 2] = -U[:, 2]
    return w, U


# ----------------------------------------------------------------------------- defaults

_ELEM_TAGS = ("geom", "joint", "site", "general", "motor", "mesh", "pair")


class _Defaults:
    def __init__(self, root: ET.Element):
        self.classes = {"main": {t: {} for t in _ELEM_TAGS}}
        d = root.find("default")
        if d is not None:
            self._walk(d, "main", top=True)

    def _walk(self, node, parent, top=False):
        name = "main" if top else node.get("class")
        if name is None:
            raise ValueError("nested <default> needs class")
        if name not in self.classes:
            self.classes[name] = {t: dict(v) for t, v in self.classes[parent].items()}
        for ch in node:
            if ch.tag == "default":
                self._walk(ch, name)
            elif ch.tag in _ELEM_TAGS:
                self.classes[name][ch.tag].update(ch.attrib)

    def resolve(self, elem: ET.Element, childclass):
        cls = elem.get("class") or childclass or "main"
        out = dict(self.classes[cls].get(elem.tag, {}))
        out.update(elem.attrib)
        return out


# ----------------------------------------------------------------------------- compiler


def _frame(attr):
    pos = _floats(attr.get("pos", "0 0 0"), 3)
    if "quat" in attr:
        q = _floats(attr["quat"], 4)
        q = q / np.linalg.norm(q)
    elif "euler" in attr:
        q = euler2quat_xyz(_floats(attr["euler"], 3))
    else:
        q = np.array([1.0, 0, 0, 0])
    for k in ("axisangle", "xyaxes", "zaxis", "fromto"):
        if k in attr:
            raise NotImplementedError(f"orientation spec {k}")
    return pos, q


def compile_mjcf(root: ET.Element) -> dict:
    comp = root.find("compiler")
    if comp is not None:
        if comp.get("angle", "degree") != "radian" or comp.get("coordinate", "local") != "local":
            raise NotImplementedError("only angle=radian coordinate=local")
    opt = root.find("option")
    optd = dict(opt.attrib) if opt is not None else {}
    timestep = float(optd.get("timestep", 0.002))
    iterations = int(optd.get("iterations", 100))
    size = root.find("size")
    njmax = int(size.get("njmax", -1)) if size is not None else -1
    nconmax = int(size.get("nconmax", -1)) if size is not None else -1
    defaults = _Defaults(root)

    # ---- assets: meshes
    meshes = {}
    asset = root.find("asset")
    if asset is not None:
        for m in asset.findall("mesh"):
            a = defaults.resolve(m, None)
            meshes[a["name"]] = dict(file=a["file"], scale=_floats(a.get("scale", "1 1 1"), 3))

    bodies, joints, geoms = [], [], []

    def add_body(elem, parent, childclass):
        if elem.tag == "worldbody":
            bid = 0
            bodies.append(dict(name="world", parent=0, pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]),
                               inertial=None, joints=[], geoms=[]))
        else:
            pos, quat = _frame(elem.attrib)
            bid = len(bodies)
            bodies.append(dict(name=elem.get("name", f"body{bid}"), parent=parent, pos=pos, quat=quat,
                               inertial=None, joints=[], geoms=[]))
            childclass = elem.get("childclass", childclass)
        b = bodies[bid]
        for ch in elem:
            if ch.tag == "inertial":
                ipos, iquat = _frame(ch.attrib)
                if "fullinertia" in ch.attrib:
                    raise NotImplementedError("fullinertia")
                b["inertial"] = dict(pos=ipos, quat=iquat, mass=float(ch.get("mass")),
                                     inertia=_floats(ch.get("diaginertia"), 3))
            elif ch.tag in ("joint", "freejoint"):
                a = defaults.resolve(ch, childclass) if ch.tag == "joint" else dict(ch.attrib, type="free")
                jt = {"free": JNT_FREE, "slide": JNT_SLIDE, "hinge": JNT_HINGE}[a.get("type", "hinge")]
                limited = a.get("limited", "false") == "true" and jt != JNT_FREE
                joints.append(dict(
                    name=a.get("name", f"joint{len(joints)}"), type=jt, body=bid,
                    pos=_floats(a.get("pos", "0 0 0"), 3), axis=_floats(a.get("axis", "0 0 1"), 3),
                    limited=limited, range=_floats(a.get("range", "0 0"), 2),
                    damping=float(a.get("damping", 0)), armature=float(a.get("armature", 0)),
                    margin=float(a.get("margin", 0)), frictionloss=float(a.get("frictionloss", 0)),
                    stiffness=float(a.get("stiffness", 0)), ref=float(a.get("ref", 0)),
                    solreflimit=_floats(a.get("solreflimit", "0.02 1"), 2),
                    solimplimit=_floats(a.get("solimplimit", "0.9 0.95 0.001 0.5 2"), 5),
                    solreffriction=_floats(a.get("solreffriction", "0.02 1"), 2),
                    solimpfriction=_floats(a.get("solimpfriction", "0.9 0.95 0.001 0.5 2"), 5)))
                if joints[-1]["stiffness"] != 0:
                    raise NotImplementedError("joint springs")
                b["joints"].append(len(joints) - 1)
            elif ch.tag == "geom":
                a = defaults.resolve(ch, childclass)
                gt = {"plane": GEOM_PLANE, "sphere": GEOM_SPHERE, "capsule": GEOM_CAPSULE, "box": GEOM_BOX,
                      "mesh": GEOM_MESH}[a.get("type", "sphere")]
                pos, quat = _frame(a)
                sz = np.zeros(3)
                if "size" in a:
                    s = _floats(a["size"])
                    sz[:s.size] = s
                fr = _floats(a.get("friction", "1 0.005 0.0001"))
                fr3 = np.array([1.0, 0.005, 0.0001])
                fr3[:fr.size] = fr
                geoms.append(dict(
                    name=a.get("name", f"geom{len(geoms)}"), type=gt, body=bid, pos=pos, quat=quat, size=sz,
                    contype=int(a.get("contype", 1)), conaffinity=int(a.get("conaffinity", 1)),
                    condim=int(a.get("condim", 3)), friction=fr3,
                    solref=_floats(a.get("solref", "0.02 1"), 2),
                    solimp=_floats(a.get("solimp", "0.9 0.95 0.001 0.5 2"), 5),
                    solmix=float(a.get("solmix", 1)), margin=float(a.get("margin", 0)), gap=float(a.get("gap", 0)),
                    priority=int(a.get("priority", 0)), density=float(a.get("density", 1000)),
                    mass=float(a["mass"]) if "mass" in a else None, mesh=a.get("mesh")))
                b["geoms"].append(len(geoms) - 1)
            elif ch.tag == "body":
                add_body(ch, bid, childclass)
            elif ch.tag in ("light", "site", "camera"):
                pass
            else:
                raise NotImplementedError(f"<{ch.tag}> in body")

    add_body(root.find("worldbody"), 0, None)
    nbody, njnt, ngeom = len(bodies), len(joints), len(geoms)

    # ---- meshes: hull vertices in the CoM/principal frame, geom pose composed with the mesh frame
    from scipy.spatial import ConvexHull
    mesh_vert, mesh_vertadr, mesh_vertnum = [], [], []
    vadr = 0
    for g in geoms:
        g["vertadr"], g["vertnum"] = -1, 0
        g["volume"], g["inertia_unit"] = 0.0, np.zeros(3)
        if g["type"] == GEOM_MESH:
            ms = meshes[g["mesh"]]
            tri = load_stl(ms["file"], ms["scale"])
            V, com, I = mesh_props_legacy(tri)
            mom, U = principal_axes(I)
            # geom frame := (geom frame) o (mesh frame)
            Rg = quat2mat(g["quat"])
            g["pos"] = g["pos"] + Rg @ com
            g["quat"] = quat_mul(g["quat"], mat2quat(U))
            g["quat"] /= np.linalg.norm(g["quat"])
            g["volume"], g["inertia_unit"] = V, mom
            collides = g["contype"] or g["conaffinity"]
            if collides:
                pts = np.unique(tri.reshape(-1, 3), axis=0)
                hull = ConvexHull(pts)
                hv = (pts[np.sort(hull.vertices)] - com) @ U  # into mesh frame: U^T (p - com)
                g["vertadr"], g["vertnum"] = vadr, hv.shape[0]
                mesh_vert.append(hv)
                vadr += hv.shape[0]
                g["size"] = np.abs(hv).max(0)
                g["rbound"] = float(np.linalg.norm(hv, axis=1).max())
            else:
                ext = (tri.reshape(-1, 3) - com) @ U
                g["size"] = np.abs(ext).max(0)
                g["rbound"] = float(np.linalg.norm(ext, axis=1).max())
        elif g["type"] == GEOM_BOX:
            g["rbound"] = float(np.linalg.norm(g["size"]))
            sx, sy, sz_ = g["size"]
            g["volume"] = 8 * sx * sy * sz_
            g["inertia_unit"] = g["volume"] / 3.0 * np.array([sy * sy + sz_ * sz_, sx * sx + sz_ * sz_, sx * sx + sy * sy])
        elif g["type"] == GEOM_CAPSULE:
            r, h = g["size"][0], g["size"][1]
            g["rbound"] = float(r + h)
            vc, vs = np.pi * r * r * 2 * h, 4.0 / 3.0 * np.pi * r ** 3
            g["volume"] = vc + vs
            ixx = vc * (r * r / 4 + (2 * h) ** 2 / 12) + vs * (2 * r * r / 5 + h * h + 3 * h * r / 4)
            izz = vc * r * r / 2 + vs * 2 * r * r / 5
            g["inertia_unit"] = np.array([ixx, ixx, izz])
        elif g["type"] == GEOM_SPHERE:
            r = g["size"][0]
            g["rbound"] = float(r)
            g["volume"] = 4.0 / 3.0 * np.pi * r ** 3
            g["inertia_unit"] = g["volume"] * 0.4 * r * r * np.ones(3)
        else:
            g["rbound"] = 0.0

    # ---- body inertial frames
    for bid, b in enumerate(bodies):
        if b["inertial"] is not None:
            continue
        mass, com, parts = 0.0, np.zeros(3), []
        for gi in b["geoms"]:
            g = geoms[gi]
            if g["type"] == GEOM_PLANE:
                continue
            m = g["mass"] if g["mass"] is not None else g["density"] * g["volume"]
            if m <= 0:
                continue
            scale = m / g["volume"]
            parts.append((m, g["pos"], quat2mat(g["quat"]), g["inertia_unit"] * scale))
            mass += m
            com += m * g["pos"]
        if mass > 0:
            com /= mass
            I = np.zeros((3, 3))
            for m, p, R, Id in parts:
                d = p - com
                I += R @ np.diag(Id) @ R.T + m * (d @ d * np.eye(3) - np.outer(d, d))
            mom, U = principal_axes(I)
            b["inertial"] = dict(pos=com, quat=mat2quat(U), mass=mass, inertia=mom)
        else:
            if b["joints"]:
                raise ValueError(f"moving body {b['name']} has no mass")
            b["inertial"] = dict(pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]), mass=0.0, inertia=np.zeros(3))

    # ---- dofs
    jnt_qposadr, jnt_dofadr, dof_body, dof_jnt, dof_parent = [], [], [], [], []
    body_dofadr, body_dofnum = [-1] * nbody, [0] * nbody
    body_lastdof = [-1] * nbody
    nq = nv = 0
    for bid, b in enumerate(bodies):
        last = body_lastdof[b["parent"]] if bid else -1
        for ji in b["joints"]:
            j = joints[ji]
            jnt_qposadr.append(nq)
            jnt_dofadr.append(nv)
            nd = 6 if j["type"] == JNT_FREE else 1
            if body_dofadr[bid] < 0:
                body_dofadr[bid] = nv
            for k in range(nd):
                dof_body.append(bid)
                dof_jnt.append(ji)
                dof_parent.append(last)
                last = nv
                nv += 1
            nq += 7 if j["type"] == JNT_FREE else 1
            body_dofnum[bid] += nd
        body_lastdof[bid] = last
    # weld / root ids
    body_weld, body_root = [0] * nbody, [0] * nbody
    for bid in range(1, nbody):
        p = bodies[bid]["parent"]
        body_weld[bid] = bid if bodies[bid]["joints"] else body_weld[p]
        body_root[bid] = bid if p == 0 else body_root[p]

    M = dict()
    M["opt_timestep"] = np.array([timestep])
    M["opt_gravity"] = np.array([0.0, 0.0, -9.81])
    M["opt_tolerance"] = np.array([float(optd.get("tolerance", 1e-8))])
    M["opt_impratio"] = np.array([float(optd.get("impratio", 1.0))])
    M["opt_mpr_tolerance"] = np.array([float(optd.get("mpr_tolerance", 1e-6))])
    M["opt_ls_tolerance"] = np.array([0.01])
    M["opt_iterations"] = np.array([iterations], dtype=np.int32)
    M["opt_ls_iterations"] = np.array([50], dtype=np.int32)
    M["opt_mpr_iterations"] = np.array([int(optd.get("mpr_iterations", 50))], dtype=np.int32)
    M["size_njmax"] = np.array([njmax], dtype=np.int32)
    M["size_nconmax"] = np.array([nconmax], dtype=np.int32)
    for k in ("solver", "cone", "integrator", "jacobian"):
        if k in optd and optd[k] not in ("Newton", "pyramidal", "Euler", "dense", "auto"):
            raise NotImplementedError(f"option {k}={optd[k]}")

    I32 = np.int32
    M["body_parentid"] = np.array([b["parent"] for b in bodies], dtype=I32)
    M["body_rootid"] = np.array(body_root, dtype=I32)
    M["body_weldid"] = np.array(body_weld, dtype=I32)
    M["body_dofadr"] = np.array(body_dofadr, dtype=I32)
    M["body_dofnum"] = np.array(body_dofnum, dtype=I32)
    M["body_jntadr"] = np.array([b["joints"][0] if b["joints"] else -1 for b in bodies], dtype=I32)
    M["body_jntnum"] = np.array([len(b["joints"]) for b in bodies], dtype=I32)
    M["body_geomadr"] = np.array([b["geoms"][0] if b["geoms"] else -1 for b in bodies], dtype=I32)
    M["body_geomnum"] = np.array([len(b["geoms"]) for b in bodies], dtype=I32)
    M["body_pos"] = np.array([b["pos"] for b in bodies])
    M["body_quat"] = np.array([b["quat"] for b in bodies])
    M["body_ipos"] = np.array([b["inertial"]["pos"] for b in bodies])
    M["body_iquat"] = np.array([b["inertial"]["quat"] for b in bodies])
    M["body_mass"] = np.array([b["inertial"]["mass"] for b in bodies])
    M["body_inertia"] = np.array([b["inertial"]["inertia"] for b in bodies])
    sub = M["body_mass"].copy()
    for bid in range(nbody - 1, 0, -1):
        sub[bodies[bid]["parent"]] += sub[bid]
    M["body_subtreemass"] = sub

    M["jnt_type"] = np.array([j["type"] for j in joints], dtype=I32)
    M["jnt_bodyid"] = np.array([j["body"] for j in joints], dtype=I32)
    M["jnt_qposadr"] = np.array(jnt_qposadr, dtype=I32)
    M["jnt_dofadr"] = np.array(jnt_dofadr, dtype=I32)
    M["jnt_pos"] = np.array([j["pos"] for j in joints])
    ax = np.array([j["axis"] for j in joints])
    M["jnt_axis"] = ax / np.linalg.norm(ax, axis=1, keepdims=True)
    M["jnt_limited"] = np.array([j["limited"] for j in joints], dtype=I32)
    M["jnt_range"] = np.array([j["range"] for j in joints])
    M["jnt_margin"] = np.array([j["margin"] for j in joints])
    M["jnt_solref"] = np.array([j["solreflimit"] for j in joints])
    M["jnt_solimp"] = np.array([j["solimplimit"] for j in joints])
    M["dof_bodyid"] = np.array(dof_body, dtype=I32)
    M["dof_jntid"] = np.array(dof_jnt, dtype=I32)
    M["dof_parentid"] = np.array(dof_parent, dtype=I32)
    M["dof_armature"] = np.array([joints[j]["armature"] for j in dof_jnt])
    M["dof_damping"] = np.array([joints[j]["damping"] for j in dof_jnt])
    M["dof_frictionloss"] = np.array([joints[j]["frictionloss"] for j in dof_jnt])
    M["dof_solref"] = np.array([joints[j]["solreffriction"] for j in dof_jnt])
    M["dof_solimp"] = np.array([joints[j]["solimpfriction"] for j in dof_jnt])
    qpos0 = np.zeros(nq)
    for ji, j in enumerate(joints):
        a = jnt_qposadr[ji]
        if j["type"] == JNT_FREE:
            qpos0[a:a + 3] = bodies[j["body"]]["pos"]
            qpos0[a + 3:a + 7] = bodies[j["body"]]["quat"]
        else:
            qpos0[a] = j["ref"]
    M["qpos0"] = qpos0

    M["geom_type"] = np.array([g["type"] for g in geoms], dtype=I32)
    M["geom_bodyid"] = np.array([g["body"] for g in geoms], dtype=I32)
    M["geom_contype"] = np.array([g["contype"] for g in geoms], dtype=I32)
    M["geom_conaffinity"] = np.array([g["conaffinity"] for g in geoms], dtype=I32)
    M["geom_condim"] = np.array([g["condim"] for g in geoms], dtype=I32)
    M["geom_priority"] = np.array([g["priority"] for g in geoms], dtype=I32)
    M["geom_pos"] = np.array([g["pos"] for g in geoms])
    M["geom_quat"] = np.array([g["quat"] for g in geoms])
    M["geom_size"] = np.array([g["size"] for g in geoms])
    M["geom_rbound"] = np.array([g["rbound"] for g in geoms])
    M["geom_friction"] = np.array([g["friction"] for g in geoms])
    M["geom_solref"] = np.array([g["solref"] for g in geoms])
    M["geom_solimp"] = np.array([g["solimp"] for g in geoms])
    M["geom_solmix"] = np.array([g["solmix"] for g in geoms])
    M["geom_margin"] = np.array([g["margin"] for g in geoms])
    M["geom_gap"] = np.array([g["gap"] for g in geoms])
    M["geom_vertadr"] = np.array([g["vertadr"] for g in geoms], dtype=I32)
    M["geom_vertnum"] = np.array([g["vertnum"] for g in geoms], dtype=I32)
    M["mesh_vert"] = np.concatenate(mesh_vert) if mesh_vert else np.zeros((0, 3))

    # ---- actuators: motors on joints, gear 1, no limits (actuators sit outside the childclass subtree)
    act = root.find("actuator")
    trn = []
    if act is not None:
        for a in act:
            at = defaults.resolve(a, None)
            if a.tag != "motor" or float(at.get("gear", "1").split()[0]) != 1.0:
                raise NotImplementedError("only <motor gear=1>")
            if at.get("ctrllimited", "false") == "true" or at.get("forcelimited", "false") == "true":
                raise NotImplementedError("limited actuators")
            ji = [j["name"] for j in joints].index(at["joint"])
            trn.append(jnt_dofadr[ji])
    M["actuator_dofid"] = np.array(trn, dtype=I32)

    # ---- explicit contact pairs (params inherited from the two geoms by the dynamic mixing rules)
    names = [g["name"] for g in geoms]
    pg1, pg2, pdim, pfri, psolref, psolimp, pmargin, pgap, psig = [], [], [], [], [], [], [], [], []
    con = root.find("contact")
    if con is not None:
        for p in con:
            if p.tag != "pair":
                raise NotImplementedError(f"<contact><{p.tag}>")
            a = defaults.resolve(p, None)
            g1, g2 = names.index(a["geom1"]), names.index(a["geom2"])
            mix = mix_contact_params(M, g1, g2)
            b1, b2 = geoms[g1]["body"], geoms[g2]["body"]
            if b1 > b2:  # signature uses ordered body pair; geoms follow
                g1, g2, b1, b2 = g2, g1, b2, b1
            pg1.append(g1)
            pg2.append(g2)
            pdim.append(int(a["condim"]) if "condim" in a else mix["condim"])
            pfri.append(_floats(a["friction"], 5) if "friction" in a else mix["friction"])
            psolref.append(_floats(a["solref"], 2) if "solref" in a else mix["solref"])
            psolimp.append(_floats(a["solimp"], 5) if "solimp" in a else mix["solimp"])
            pmargin.append(float(a["margin"]) if "margin" in a else mix["margin"])
            pgap.append(float(a["gap"]) if "gap" in a else mix["gap"])
            psig.append((b1 << 16) + b2)
    order = np.argsort(np.array(psig, dtype=np.int64), kind="stable") if psig else np.zeros(0, dtype=np.int64)
    M["pair_geom1"] = np.array(pg1, dtype=I32)[order]
    M["pair_geom2"] = np.array(pg2, dtype=I32)[order]
    M["pair_dim"] = np.array(pdim, dtype=I32)[order]
    M["pair_signature"] = np.array(psig, dtype=I32)[order]
    M["pair_friction"] = np.array(pfri).reshape(-1, 5)[order]
    M["pair_solref"] = np.array(psolref).reshape(-1, 2)[order]
    M["pair_solimp"] = np.array(psolimp).reshape(-1, 5)[order]
    M["pair_margin"] = np.array(pmargin)[order]
    M["pair_gap"] = np.array(pgap)[order]

    M["names_body"] = [b["name"] for b in bodies]
    M["names_geom"] = names
    M["names_joint"] = [j["name"] for j in joints]
    M["names_actuator"] = [a.get("name", "") for a in act] if act is not None else []
    M["nq"], M["nv"], M["nbody"], M["ngeom"], M["njnt"], M["nu"] = nq, nv, nbody, ngeom, njnt, len(trn)

    set_const(M)
    return M


def mix_contact_params(M, g1, g2):
    """Dynamic contact-parameter mixing (same priority): condim max, friction max, solref mixed when both
    positive else element-wise min, solimp mixed by solmix, margin/gap max."""
    s1, s2 = M["geom_solmix"][g1], M["geom_solmix"][g2]
    mix = s1 / (s1 + s2) if s1 + s2 > 0 else 0.5
    f = np.maximum(M["geom_friction"][g1], M["geom_friction"][g2])
    r1, r2 = M["geom_solref"][g1], M["geom_solref"][g2]
    solref = mix * r1 + (1 - mix) * r2 if (r1[0] > 0 and r2[0] > 0) else np.minimum(r1, r2)
    return dict(
        condim=int(max(M["geom_condim"][g1], M["geom_condim"][g2])),
        friction=np.array([f[0], f[0], f[1], f[2], f[2]]),
        solref=solref,
        solimp=mix * M["geom_solimp"][g1] + (1 - mix) * M["geom_solimp"][g2],
        margin=float(max(M["geom_margin"][g1], M["geom_margin"][g2])),
        gap=float(max(M["geom_gap"][g1], M["geom_gap"][g2])))


# ----------------------------------------------------------------------------- kinematics / mj_setConst (numpy)


def fk(M, qpos):
    """Forward kinematics.  Returns dict xpos[nbody,3], xquat[nbody,4],
    xanchor[njnt,3], xaxis[njnt,3].  Plain numpy: used for compile-time constants and for expert-motion
    derivations (dataset_singledepth.py:222-237) -- the per-step FK of the hot path is the CUDA kernel."""
    nbody = M["nbody"]
    xpos, xquat = np.zeros((nbody, 3)), np.zeros((nbody, 4))
    xquat[0, 0] = 1
    xanchor, xaxis = np.zeros((M["njnt"], 3)), np.zeros((M["njnt"], 3))
    for b in range(1, nbody):
        p = M["body_parentid"][b]
        ja, jn = M["body_jntadr"][b], M["body_jntnum"][b]
        if jn == 1 and M["jnt_type"][ja] == JNT_FREE:
            a = M["jnt_qposadr"][ja]
            pos = qpos[a:a + 3].copy()
            quat = qpos[a + 3:a + 7] / np.linalg.norm(qpos[a + 3:a + 7])
            xanchor[ja], xaxis[ja] = pos, np.array([0, 0, 1.0])
        else:
            Rp = quat2mat(xquat[p])
            pos = xpos[p] + Rp @ M["body_pos"][b]
            quat = quat_mul(xquat[p], M["body_quat"][b])
            for j in range(ja, ja + jn):
                R = quat2mat(quat)
                xanchor[j] = pos + R @ M["jnt_pos"][j]
                xaxis[j] = R @ M["jnt_axis"][j]
                q = qpos[M["jnt_qposadr"][j]] - M["qpos0"][M["jnt_qposadr"][j]]
                if M["jnt_type"][j] == JNT_SLIDE:
                    pos = pos + xaxis[j] * q
                elif M["jnt_type"][j] == JNT_HINGE:
                    quat = quat_mul(quat, axisangle2quat(M["jnt_axis"][j], q))
                    pos = xanchor[j] - quat2mat(quat) @ M["jnt_pos"][j]
                else:
                    raise NotImplementedError
        quat = quat / np.linalg.norm(quat)
        xpos[b], xquat[b] = pos, quat
    return dict(xpos=xpos, xquat=xquat, xanchor=xanchor, xaxis=xaxis)


def body_jacobians(M, kin, point_of_body):
    """Dense 6 x nv Jacobians (rows 0:3 translational at point_of_body[b], 3:6 rotational) per body."""
    nbody, nv = M["nbody"], M["nv"]
    J = np.zeros((nbody, 6, nv))
    for b in range(1, nbody):
        d = M["body_dofadr"][b] + M["body_dofnum"][b] - 1 if M["body_dofnum"][b] else -1
        bb = b
        while d < 0 and bb > 0:  # static link: walk up
            bb = M["body_parentid"][bb]
            d = M["body_dofadr"][bb] + M["body_dofnum"][bb] - 1 if M["body_dofnum"][bb] else -1
        while d >= 0:
            j = M["dof_jntid"][d]
            t = M["jnt_type"][j]
            if t == JNT_SLIDE:
                J[b, 0:3, d] = kin["xaxis"][j]
            elif t == JNT_HINGE:
                J[b, 3:6, d] = kin["xaxis"][j]
                J[b, 0:3, d] = np.cross(kin["xaxis"][j], point_of_body[b] - kin["xanchor"][j])
            else:
                k = d - M["jnt_dofadr"][j]
                if k < 3:
                    J[b, k, d] = 1.0
                else:
                    R = quat2mat(kin["xquat"][M["jnt_bodyid"][j]])
                    axis = R[:, k - 3]
                    J[b, 3:6, d] = axis
                    J[b, 0:3, d] = np.cross(axis, point_of_body[b] - kin["xpos"][M["jnt_bodyid"][j]])
            d = M["dof_parentid"][d]
    return J


def mass_matrix(M, qpos):
    """Dense joint-space inertia by summing body Jacobian quadratic forms (independent of the CRB
    recursion used by the oracle and the kernel -- serves as their cross-check)."""
    kin = fk(M, qpos)
    nbody = M["nbody"]
    xipos = np.array([kin["xpos"][b] + quat2mat(kin["xquat"][b]) @ M["body_ipos"][b] for b in range(nbody)])
    J = body_jacobians(M, kin, xipos)
    H = np.diag(M["dof_armature"]).astype(np.float64)
    for b in range(1, nbody):
        Ri = quat2mat(quat_mul(kin["xquat"][b], M["body_iquat"][b]))
        Iw = Ri @ np.diag(M["body_inertia"][b]) @ Ri.T
        H += M["body_mass"][b] * J[b, 0:3].T @ J[b, 0:3] + J[b, 3:6].T @ Iw @ J[b, 3:6]
    return H, kin, xipos, J


def set_const(M):
    """mj_setConst subset: meaninertia, dof_invweight0, body_invweight0 at qpos0."""
    nv, nbody = M["nv"], M["nbody"]
    H, kin, xipos, J = mass_matrix(M, M["qpos0"])
    Hinv = np.linalg.inv(H)
    M["stat_meaninertia"] = np.array([np.trace(H) / max(nv, 1)])
    dw = np.diag(Hinv).copy()
    for j in range(M["njnt"]):
        if M["jnt_type"][j] == JNT_FREE:
            a = M["jnt_dofadr"][j]
            dw[a:a + 3] = dw[a:a + 3].mean()
            dw[a + 3:a + 6] = dw[a + 3:a + 6].mean()
    M["dof_invweight0"] = dw
    bw = np.zeros((nbody, 2))
    for b in range(1, nbody):
        if M["body_weldid"][b] == 0:
            continue
        A = J[b] @ Hinv @ J[b].T
        bw[b, 0] = max(MINVAL, (A[0, 0] + A[1, 1] + A[2, 2]) / 3)
        bw[b, 1] = max(MINVAL, (A[3, 3] + A[4, 4] + A[5, 5]) / 3)
    M["body_invweight0"] = bw
    M["qM0"] = H


# ----------------------------------------------------------------------------- env-level index discovery


def env_indices(M):
    """Index discovery of HandObjMimic4.__init__ (ho_im4.py:72-107)."""
    hand_body_idx = [i for i, n in enumerate(M["names_body"]) if n.startswith("link")]
    gnames = M["names_geom"]
    h0 = h1 = -1
    for i, n in enumerate(gnames):
        if n.startswith("robot0:") and h0 == -1:
            h0 = i
        if (not n.startswith("robot0:")) and h0 != -1:
            h1 = i - 1
            break
    o0, o1 = h1 + 1, -1
    for i in range(o0, len(gnames)):
        if not gnames[i].startswith("C_"):
            o1 = i - 1
            break
    return dict(hand_body_idx=hand_body_idx, obj_body_idx=hand_body_idx[-1] + 1,
                hand_geom_range=(h0, h1), obj_geom_range=(o0, o1),
                hand_mass=float(sum(M["body_mass"][i] for i in hand_body_idx)),
                palm_body=M["names_body"].index("link_palm"))


def compile_hand_object(hand_xml: str, obj_xml: str) -> dict:
    M = compile_mjcf(merge_mjcf(hand_xml, obj_xml))
    M["env"] = env_indices(M)
    return M
