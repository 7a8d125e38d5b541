"""Independent restatement of Scene::load_file's flattening, in numpy — TEST INFRASTRUCTURE.

The product's JSON/OBJ/MERL loader (tray_rust_b200/csrc/trb_loader.cpp) and the oracle both consume the SAME flattened
`trb_scene_desc`, so a misreading of the scene format shared by the two would be invisible to every GPU-vs-oracle test.
This module re-derives the description from the JSON with different machinery (Python's json, numpy float32 / float64,
LAPACK's SVD instead of the loader's Jacobi iteration) directly from the reference's loading code:

    load_transform      /root/reference/src/scene.rs:749-820     (translate / scale / rotate_x|y|z / rotate / matrix, left-multiplied)
    Transform::*        src/linalg/transform.rs:32-109
    Keyframe::decompose src/linalg/keyframe.rs:32-58              (f64 SVD polar decomposition, flip on det < 0, Quaternion::from_matrix)
    Quaternion::from_matrix src/linalg/quaternion.rs:25-59
    with_keyframes      src/linalg/animated_transform.rs:22-33    (shortest-path quaternion flips)
    group stacking      src/scene.rs:565-574, animated_transform.rs:78-87  (child levels first, then the group's)
    load_objects        src/scene.rs:513-580;  load_geometry :583-640; load_color :705-724; load_animated_color :728-752
    load_materials      src/scene.rs:405-508;  load_camera :253-294; load_film :185-208; load_integrator :296-315
    Mesh::load_obj      src/geometry/mesh.rs:49-76 over tobj 0.1.6 (one model per o/g, fan triangulation, vertices unified per
                        (v, vt, vn) triple in first-seen order — restated from the crate's documented behaviour)
    Merl::load_file     src/material/merl.rs:51-84

`flatten(path)` returns plain Python / numpy data in the order the C ABI's arrays use, for field-by-field comparison.
"""
import json
import math
import os

import numpy as np

F32 = np.float32


def _m(rows):
    return np.array(rows, dtype=F32).reshape(4, 4)


def translate(v):
    return _m([1, 0, 0, v[0], 0, 1, 0, v[1], 0, 0, 1, v[2], 0, 0, 0, 1])


def scale(v):
    return _m([v[0], 0, 0, 0, 0, v[1], 0, 0, 0, 0, v[2], 0, 0, 0, 0, 1])


def _sc(deg):
    r = F32(math.pi) / F32(180.0) * F32(deg)   # linalg::to_radians
    return F32(np.sin(r)), F32(np.cos(r))


def rotate_x(deg):
    s, c = _sc(deg)
    return _m([1, 0, 0, 0, 0, c, -s, 0, 0, s, c, 0, 0, 0, 0, 1])


def rotate_y(deg):
    s, c = _sc(deg)
    return _m([c, 0, s, 0, 0, 1, 0, 0, -s, 0, c, 0, 0, 0, 0, 1])


def rotate_z(deg):
    s, c = _sc(deg)
    return _m([c, -s, 0, 0, s, c, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1])


def rotate(axis, deg):
    a = np.asarray(axis, F32)
    a = a / F32(np.sqrt(np.sum(a * a)))
    s, c = _sc(deg)
    m = np.eye(4, dtype=F32)
    x, y, z = a
    m[0, 0] = x * x + (1 - x * x) * c; m[0, 1] = x * y * (1 - c) - z * s; m[0, 2] = x * z * (1 - c) + y * s
    m[1, 0] = x * y * (1 - c) + z * s; m[1, 1] = y * y + (1 - y * y) * c; m[1, 2] = y * z * (1 - c) - x * s
    m[2, 0] = x * z * (1 - c) - y * s; m[2, 1] = y * z * (1 - c) + x * s; m[2, 2] = z * z + (1 - z * z) * c
    return m


def vec3(e):
    assert isinstance(e, list) and len(e) == 3
    return [F32(x) for x in e]


def load_transform(steps):
    t = np.eye(4, dtype=F32)
    for st in steps:
        ty = st["type"]
        if ty == "translate":
            m = translate(vec3(st["translation"]))
        elif ty == "scale":
            s = st["scaling"]
            m = scale(vec3(s) if isinstance(s, list) else [F32(s)] * 3)
        elif ty == "rotate_x":
            m = rotate_x(st["rotation"])
        elif ty == "rotate_y":
            m = rotate_y(st["rotation"])
        elif ty == "rotate_z":
            m = rotate_z(st["rotation"])
        elif ty == "rotate":
            m = rotate(vec3(st["axis"]), st["rotation"])
        elif ty == "matrix":
            m = _m([x for row in st["matrix"] for x in row])
        else:
            raise ValueError("Unrecognized transform type " + ty)
        t = (m.astype(F32) @ t).astype(F32)   # transform = T_new * transform
    return t


def quat_from_matrix(m):
    """Quaternion::from_matrix (quaternion.rs:25-59), f32; returns (x, y, z, w)."""
    m = m.astype(F32)
    trace = m[0, 0] + m[1, 1] + m[2, 2]
    if trace > 0:
        s = F32(np.sqrt(trace + F32(1)))
        w = s / F32(2)
        s = F32(0.5) / s
        return np.array([s * (m[2, 1] - m[1, 2]), s * (m[0, 2] - m[2, 0]), s * (m[1, 0] - m[0, 1]), w], F32)
    nxt = [1, 2, 0]
    i = 1 if m[1, 1] > m[0, 0] else (2 if m[2, 2] > m[0, 0] else 0)
    j = nxt[i]; k = nxt[j]
    q = np.zeros(3, F32)
    s = F32(np.sqrt((m[i, i] - (m[j, j] + m[k, k])) + F32(1)))
    q[i] = s * F32(0.5)
    if s != 0:
        s = F32(0.5) / s
    w = (m[k, j] - m[j, k]) * s
    q[j] = (m[j, i] + m[i, j]) * s
    q[k] = (m[k, i] + m[i, k]) * s
    return np.array([q[0], q[1], q[2], w], F32)


def decompose(t):
    """Keyframe::decompose: translation column, polar factors of the 3x3 part through an f64 SVD."""
    a = t[:3, :3].astype(np.float64)
    u, s, vt = np.linalg.svd(a)
    q = u @ vt
    p = vt.T @ np.diag(s) @ vt
    if np.linalg.det(q) < 0:
        q, p = -q, -p
    rot = np.eye(4, dtype=F32)
    rot[:3, :3] = q.astype(F32)
    return t[:3, 3].astype(F32), quat_from_matrix(rot), np.array([p[0, 0], p[1, 1], p[2, 2]]).astype(F32)


def keyframe(t):
    tr, q, s = decompose(t)
    return {"t": tr, "q": q, "s": s}


def load_keyframes(e):
    keys = [keyframe(load_transform(cp["transform"])) for cp in e["control_points"]]
    for i in range(1, len(keys)):   # with_keyframes: shortest path
        if float(np.dot(keys[i - 1]["q"].astype(F32), keys[i]["q"].astype(F32))) < 0:
            keys[i]["q"] = -keys[i]["q"]
    return {"degree": int(e.get("degree", 3)), "keys": keys, "knots": [F32(k) for k in e["knots"]]}


def unanimated(t):
    return {"degree": 0, "keys": [keyframe(t)], "knots": [F32(0), F32(1)]}


def object_levels(o):
    if "keyframes" in o:
        return [load_keyframes(o["keyframes"])]
    return [unanimated(load_transform(o["transform"]))]


def load_color(e):
    v = [F32(x) for x in e]
    assert len(v) in (3, 4)
    c = [v[0], v[1], v[2], F32(1)]            # Colorf::new: a = 1
    if len(v) == 4:
        c = [x * v[3] for x in c]             # Colorf * f32 scales all four (Q15)
    return c


def load_animated_color(e):
    if isinstance(e[0], (int, float)):
        return [(load_color(e), F32(0))]
    return [(load_color(k["color"]), F32(k["time"])) for k in e]


MAT_TYPES = {"matte": 0, "plastic": 1, "metal": 2, "specular_metal": 3, "glass": 4, "rough_glass": 5, "merl": 6}
SHAPES = {"sphere": 1, "disk": 2, "rectangle": 3, "plane": 3, "mesh": 4}


def rgb(e):
    c = load_color(e)
    return c[:3]


def scalar(e):
    return F32(e)


def load_materials(path, elems):
    mats, names, merl_files = [], {}, []
    for m in elems:
        ty = m["type"]
        r = {"type": MAT_TYPES[ty], "c0": [0, 0, 0], "c1": [0, 0, 0], "roughness": 0.0, "eta": None, "merl": None}
        if ty in ("glass", "rough_glass"):
            r["c0"], r["c1"], r["eta"] = rgb(m["reflect"]), rgb(m["transmit"]), scalar(m["eta"])
            if ty == "rough_glass":
                r["roughness"] = scalar(m["roughness"])
        elif ty == "matte":
            r["c0"], r["roughness"] = rgb(m["diffuse"]), scalar(m["roughness"])
        elif ty == "plastic":
            r["c0"], r["c1"], r["roughness"] = rgb(m["diffuse"]), rgb(m["gloss"]), scalar(m["roughness"])
        elif ty in ("metal", "specular_metal"):
            r["c0"], r["c1"] = rgb(m["refractive_index"]), rgb(m["absorption_coefficient"])
            if ty == "metal":
                r["roughness"] = scalar(m["roughness"])
        elif ty == "merl":
            f = m["file"]
            r["merl"] = len(merl_files)
            merl_files.append(f if os.path.isabs(f) else os.path.join(path, f))
        assert m["name"] not in names
        names[m["name"]] = len(mats)
        mats.append(r)
    return mats, names, merl_files


def load_merl(file):
    """Merl::load_file: planar f64 r, g, b scaled by (1/1500, 1/1500, 1.66/1500), cast to f32, clamped at 0, interleaved."""
    raw = open(file, "rb").read()
    dims = np.frombuffer(raw[:12], "<i4")
    assert tuple(dims) == (90, 90, 180)
    n = 90 * 90 * 180
    planes = np.frombuffer(raw[12:12 + 3 * n * 8], "<f8").reshape(3, n)
    out = np.empty((n, 3), F32)
    for c, s in enumerate((1.0 / 1500.0, 1.0 / 1500.0, 1.66 / 1500.0)):
        out[:, c] = np.maximum(F32(0), (planes[c] * s).astype(F32))
    return out.reshape(-1)


def load_obj(file):
    """tobj-style: one model per `o` / `g`; faces fan-triangulated; one output vertex per distinct (v, vt, vn) triple, numbered
    in order of first appearance; models without normals or texcoords are skipped by Mesh::load_obj."""
    V, VT, VN = [], [], []
    models, cur = [], None

    def start(name):
        nonlocal cur
        cur = {"name": name, "remap": {}, "pos": [], "nrm": [], "uv": [], "idx": [], "has_n": True, "has_t": True}
        models.append(cur)

    def corner(tok):
        parts = tok.split("/")
        vi = int(parts[0]); ti = int(parts[1]) if len(parts) > 1 and parts[1] else 0; ni = int(parts[2]) if len(parts) > 2 and parts[2] else 0
        vi = vi - 1 if vi > 0 else len(V) + vi
        ti = (ti - 1 if ti > 0 else len(VT) + ti) if ti else None
        ni = (ni - 1 if ni > 0 else len(VN) + ni) if ni else None
        key = (vi, ti, ni)
        if key not in cur["remap"]:
            cur["remap"][key] = len(cur["pos"])
            cur["pos"].append(V[vi])
            if ti is None:
                cur["has_t"] = False
            else:
                cur["uv"].append(VT[ti])
            if ni is None:
                cur["has_n"] = False
            else:
                cur["nrm"].append(VN[ni])
        return cur["remap"][key]

    for line in open(file):
        t = line.split()
        if not t or t[0].startswith("#"):
            continue
        if t[0] == "v":
            V.append([F32(x) for x in t[1:4]])
        elif t[0] == "vt":
            VT.append([F32(x) for x in t[1:3]])
        elif t[0] == "vn":
            VN.append([F32(x) for x in t[1:4]])
        elif t[0] in ("o", "g"):
            start(" ".join(t[1:]) if len(t) > 1 else "unnamed_object")
        elif t[0] == "f":
            if cur is None:
                start("unnamed_object")
            c = [corner(x) for x in t[1:]]
            for k in range(1, len(c) - 1):
                cur["idx"] += [c[0], c[k], c[k + 1]]
    out = {}
    for m in models:
        if not m["idx"] or not m["has_n"] or not m["has_t"]:
            continue
        out[m["name"]] = (np.array(m["pos"], F32), np.array(m["nrm"], F32), np.array(m["uv"], F32), np.array(m["idx"], np.uint32).reshape(-1, 3))
    return out


def flatten(json_path):
    d = json.load(open(json_path))
    base = os.path.dirname(os.path.abspath(json_path))
    mats, mat_names, merl_files = load_materials(base, d["materials"])
    meshes, mesh_index = [], {}
    instances = []

    def geometry(g):
        ty = g["type"]
        if ty == "sphere":
            return SHAPES[ty], F32(g["radius"]), F32(0), None
        if ty == "disk":
            return SHAPES[ty], F32(g["radius"]), F32(g["inner_radius"]), None
        if ty == "plane":
            return SHAPES[ty], F32(2), F32(2), None
        if ty == "rectangle":
            return SHAPES[ty], F32(g["width"]), F32(g["height"]), None
        f = g["file"] if os.path.isabs(g["file"]) else os.path.join(base, g["file"])
        key = (f, g["model"])
        if key not in mesh_index:
            mesh_index[key] = len(meshes)
            meshes.append(load_obj(f)[g["model"]])
        return SHAPES["mesh"], F32(0), F32(0), mesh_index[key]

    def walk(objs, parents):
        for o in objs:
            levels = object_levels(o)
            ty = o["type"]
            if ty == "group":
                walk(o["objects"], levels + parents)            # gi.set_transform(group * t): the child's levels, then the group's
                continue
            stack = levels + parents
            if ty == "emitter":
                em = load_animated_color(o["emission"])
                if o["emitter"] == "point":
                    instances.append({"kind": 2, "shape": 0, "p0": F32(0), "p1": F32(0), "mesh": None, "material": None, "levels": stack, "emission": em})
                else:
                    sh, p0, p1, me = geometry(o["geometry"])
                    assert sh != 4, "mesh is not sampleable"
                    instances.append({"kind": 1, "shape": sh, "p0": p0, "p1": p1, "mesh": me, "material": mat_names[o["material"]], "levels": stack, "emission": em})
            else:
                sh, p0, p1, me = geometry(o["geometry"])
                instances.append({"kind": 0, "shape": sh, "p0": p0, "p1": p1, "mesh": me, "material": mat_names[o["material"]], "levels": stack, "emission": None})

    walk(d["objects"], [])
    cams = []
    for c in (d["cameras"] if "cameras" in d else [d["camera"]]):
        if "keyframes" in c:
            lv = [load_keyframes(c["keyframes"])]
        else:
            lv = [unanimated(load_transform(c["transform"]))]
        fov = c["fov"]
        cams.append({"levels": lv, "fov": [F32(x) for x in fov] if isinstance(fov, list) else F32(fov), "fov_knots": [F32(x) for x in c.get("fov_knots", [])],
                     "fov_degree": int(c.get("fov_spline_degree", 0)), "shutter_size": F32(c.get("shutter_size", 0.5)), "active_at": int(c.get("active_at", 0))})
    cams.sort(key=lambda c: c["active_at"])
    film = d["film"]
    integ = d["integrator"]
    return {"film": film, "integrator": integ, "cameras": cams, "instances": instances, "materials": mats, "meshes": meshes,
            "merl": [load_merl(f) for f in merl_files]}
