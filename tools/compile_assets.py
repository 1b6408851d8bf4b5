#!/usr/bin/env python
"""Asset compiler: reference URDF + OBJ + PNG  ->  real_robots_b200/assets/scene.rrb

Reads the reference's data directory (read-only)
    /root/reference/real_robots/data/kuka_gripper_description/{urdf,meshes}
and writes ONE binary blob (format: real_robots_b200/blob.py) holding everything
the simulator needs: the Kuka+gripper kinematic tree in Bullet's link order
(SURVEY.md App. A.2), masses / inertias (both the URDF values and the
AABB-box values Bullet recomputes when URDF_USE_INERTIA_FROM_FILE is not set,
App. C-2), joint limits / damping, contact parameters, convex collision hulls,
render triangles (position, normal, uv as float32 in each link's inertial
frame, the layout TinyRenderer consumes, App. B.5) and the textures (RGB8,
vertically flipped as TinyRenderer stores them).

The blob is DERIVED DATA: no reference source code is copied.  The script is
deterministic; re-running it on the same reference tree reproduces the blob
byte for byte (checked by tests/test_assets.py when /root/reference exists).

usage: python tools/compile_assets.py [--ref /root/reference] [--out real_robots_b200/assets/scene.rrb]
"""
import argparse
import json
import os
import sys
import xml.etree.ElementTree as ET

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from real_robots_b200.blob import write_blob  # noqa: E402

COLLISION_MARGIN = 0.001          # pybullet gUrdfDefaultCollisionMargin [BULLET-MEM]
DEFAULT_RGBA = (0.8, 0.8, 0.8, 1.0)   # UrdfMaterialColor() default when <material> has no <color> [BULLET-MEM]
DEFAULT_LATERAL_FRICTION = 0.5    # URDF importer default [BULLET-MEM]

# robot.py:49-50 — load order == pybullet body unique ids
BODY_FILES = [("kuka", "kuka_gripper.urdf"), ("table", "table.urdf"), ("cube", "cube.urdf"),
              ("tomato", "tomato.urdf"), ("mustard", "mustard.urdf")]
# robot.py:19-24 (xyz + Euler rpy)
OBJECT_POSES = {
    "table":   [0.00000,  0.00000, 0.08000, 0.00000, 0.00000, 0.00000],
    "mustard": [-0.10000, 0.30000, 0.45000, 1.57080, 3.14159, 0.00000],
    "cube":    [-0.10000, 0.00000, 0.45000, 0.00000, 0.00000, 0.00000],
    "tomato":  [-0.10000, -0.30000, 0.45000, 0.00000, 0.00000, 0.00000],
}
ROBOT_POSITION = [-0.55, 0.0, -0.04]   # robot.py:46


# ----------------------------------------------------------------------------- math helpers
def rpy_to_mat(rpy):
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def euler_to_quat(rpy):
    """pybullet.getQuaternionFromEuler == btQuaternion::setEulerZYX(yaw, pitch, roll); returns xyzw."""
    roll, pitch, yaw = rpy
    hy, hp, hr = yaw * 0.5, pitch * 0.5, roll * 0.5
    cy, sy, cp, sp, cr, sr = np.cos(hy), np.sin(hy), np.cos(hp), np.sin(hp), np.cos(hr), np.sin(hr)
    return np.array([sr * cp * cy - cr * sp * sy,
                     cr * sp * cy + sr * cp * sy,
                     cr * cp * sy - sr * sp * cy,
                     cr * cp * cy + sr * sp * sy])


class Xform:
    def __init__(self, R=None, t=None):
        self.R = np.eye(3) if R is None else np.asarray(R, float)
        self.t = np.zeros(3) if t is None else np.asarray(t, float)

    def __mul__(self, o):
        return Xform(self.R @ o.R, self.R @ o.t + self.t)

    def inv(self):
        return Xform(self.R.T, -self.R.T @ self.t)

    def apply(self, p):
        return p @ self.R.T + self.t


def parse_origin(el):
    if el is None:
        return Xform()
    xyz = [float(v) for v in el.get("xyz", "0 0 0").split()]
    rpy = [float(v) for v in el.get("rpy", "0 0 0").split()]
    return Xform(rpy_to_mat(rpy), xyz)


# ----------------------------------------------------------------------------- OBJ / MTL
def parse_obj(path):
    """Returns dict(v, vt, vn, faces[list of list of (vi,ti,ni)], mtllib, usemtl)."""
    v, vt, vn, faces = [], [], [], []
    mtllib = None
    for line in open(path):
        s = line.split()
        if not s:
            continue
        if s[0] == "v":
            v.append([float(x) for x in s[1:4]])
        elif s[0] == "vt":
            vt.append([float(x) for x in s[1:3]])
        elif s[0] == "vn":
            vn.append([float(x) for x in s[1:4]])
        elif s[0] == "mtllib":
            mtllib = s[1]
        elif s[0] == "f":
            face = []
            for tok in s[1:]:
                p = tok.split("/")
                vi = int(p[0])
                ti = int(p[1]) if len(p) > 1 and p[1] else 0
                ni = int(p[2]) if len(p) > 2 and p[2] else 0
                assert vi > 0 and ti >= 0 and ni >= 0, "negative OBJ indices not handled"
                face.append((vi - 1, ti - 1, ni - 1))
            faces.append(face)
    return dict(v=np.array(v, float), vt=np.array(vt, float).reshape(-1, 2),
                vn=np.array(vn, float).reshape(-1, 3), faces=faces, mtllib=mtllib)


def mtl_texture(obj_path, mtllib):
    if mtllib is None:
        return None
    p = os.path.join(os.path.dirname(obj_path), mtllib)
    if not os.path.exists(p):
        return None
    for line in open(p):
        s = line.split()
        if s and s[0] == "map_Kd":
            return os.path.join(os.path.dirname(obj_path), s[1])
    return None


def triangulate(obj):
    """Fan triangulation (v0, v[k-1], v[k]) in file order — Bullet's tinyobj loader [BULLET-MEM].
    Returns float32 [T,3,8] = pos(3) normal(3) uv(2)."""
    tris = []
    has_vt = len(obj["vt"]) > 0
    for face in obj["faces"]:
        for k in range(2, len(face)):
            tri = []
            corners = (face[0], face[k - 1], face[k])
            for (vi, ti, ni) in corners:
                p = obj["v"][vi]
                uv = obj["vt"][ti] if (has_vt and ti >= 0) else np.array([0.5, 0.5])
                n = obj["vn"][ni] if ni >= 0 else None
                tri.append((p, n, uv))
            if any(c[1] is None for c in tri):
                p0, p1, p2 = tri[0][0], tri[1][0], tri[2][0]
                nn = np.cross(p1 - p0, p2 - p0)
                l2 = nn @ nn
                nn = nn / np.sqrt(l2) if l2 > 1.19e-7 else np.zeros(3)
                tri = [(c[0], nn, c[2]) for c in tri]
            tris.append([np.concatenate([c[0], c[1], c[2]]) for c in tri])
    return np.array(tris, dtype=np.float32).reshape(-1, 3, 8)


# ----------------------------------------------------------------------------- URDF
class Link:
    pass


def parse_urdf(path):
    root = ET.parse(path).getroot()
    materials = {}
    for m in root.findall("material"):
        c = m.find("color")
        if c is not None:
            materials[m.get("name")] = [float(x) for x in c.get("rgba").split()]
    links, order = {}, []
    for le in root.findall("link"):
        L = Link()
        L.name = le.get("name")
        ine = le.find("inertial")
        L.inertial = parse_origin(ine.find("origin") if ine is not None else None)
        L.mass = float(ine.find("mass").get("value")) if ine is not None else 0.0
        if ine is not None and ine.find("inertia") is not None:
            i = ine.find("inertia")
            L.inertia_urdf = [float(i.get(k)) for k in ("ixx", "iyy", "izz")]
            assert all(float(i.get(k)) == 0 for k in ("ixy", "ixz", "iyz"))
        else:
            L.inertia_urdf = [0, 0, 0]
        ce = le.find("contact")
        L.contact = {}
        if ce is not None:
            for k in ce:
                L.contact[k.tag] = float(k.get("value"))
        vis = le.find("visual")
        L.visual = None
        if vis is not None:
            mesh = vis.find("geometry").find("mesh")
            scale = [float(x) for x in mesh.get("scale", "1 1 1").split()]
            mat = vis.find("material")
            matname = mat.get("name") if mat is not None else None
            # a local <material> with texture/color replaces the global of the same name (UrdfParser::parseVisual)
            if mat is not None and (mat.find("texture") is not None or mat.find("color") is not None):
                c = mat.find("color")
                materials[matname] = ([float(x) for x in c.get("rgba").split()] if c is not None
                                      else list(DEFAULT_RGBA))
            L.visual = dict(origin=parse_origin(vis.find("origin")),
                            file=mesh.get("filename").split("/")[-1], scale=scale, material=matname)
        col = le.find("collision")
        L.collision = None
        if col is not None:
            mesh = col.find("geometry").find("mesh")
            scale = [float(x) for x in mesh.get("scale", "1 1 1").split()]
            L.collision = dict(origin=parse_origin(col.find("origin")),
                               file=mesh.get("filename").split("/")[-1], scale=scale)
        L.children = []
        L.parent = None
        L.joint = None
        links[L.name] = L
        order.append(L.name)
    for je in root.findall("joint"):
        J = dict(name=je.get("name"), type=je.get("type"), parent=je.find("parent").get("link"),
                 child=je.find("child").get("link"), origin=parse_origin(je.find("origin")))
        ax = je.find("axis")
        J["axis"] = [float(x) for x in ax.get("xyz").split()] if ax is not None else [1, 0, 0]
        lim = je.find("limit")
        J["lower"] = float(lim.get("lower", "0")) if lim is not None else 0.0
        J["upper"] = float(lim.get("upper", "0")) if lim is not None else 0.0
        dyn = je.find("dynamics")
        J["damping"] = float(dyn.get("damping", "0")) if dyn is not None else 0.0
        links[J["child"]].parent = J["parent"]
        links[J["child"]].joint = J
        links[J["parent"]].children.append(J["child"])
    roots = [n for n in order if links[n].parent is None]
    assert len(roots) == 1
    # Bullet link index = pre-order DFS, children in URDF joint order (ComputeParentIndices) [BULLET-MEM]
    dfs = []

    def rec(n):
        dfs.append(n)
        for c in links[n].children:
            rec(c)
    rec(roots[0])
    return links, dfs, materials


# ----------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                                                  "real_robots_b200", "assets", "scene.rrb"))
    args = ap.parse_args()
    from PIL import Image
    from scipy.spatial import ConvexHull

    data = os.path.join(args.ref, "real_robots", "data", "kuka_gripper_description")
    mesh_dir = os.path.join(data, "meshes")

    shapes = []      # one per (body, link) that has geometry, in render submission order
    tri_all, hull_all = [], []
    hull_face_all = []          # hull facet planes (n, d) for tests / future SAT
    tex_index, tex_meta, tex_data = {}, [], []
    manifest = dict(bodies=[], shapes=[], robot_links=[], robot_joints=[])
    arrays = {}

    def get_texture(path):
        if path is None:
            return -1
        if path in tex_index:
            return tex_index[path]
        im = np.asarray(Image.open(path).convert("RGB"), dtype=np.uint8)
        im = im[::-1].copy()      # TGAImage::flip_vertically() in Model::setDiffuseTextureFromData [BULLET-MEM]
        uniform = bool((im == im[0, 0]).all())
        h, w = im.shape[:2]
        off = sum(len(t) for t in tex_data)
        if uniform:
            payload = im[:1, :1].reshape(-1)      # keep nominal w,h; 1 texel of data
        else:
            payload = im.reshape(-1)
        tex_data.append(payload)
        tex_meta.append([w, h, off, 1 if uniform else 0])
        tex_index[path] = len(tex_meta) - 1
        return tex_index[path]

    robot = None
    for body_id, (bname, fname) in enumerate(BODY_FILES):
        links, dfs, materials = parse_urdf(os.path.join(data, "urdf", fname))
        manifest["bodies"].append(dict(name=bname, id=body_id, links=dfs))
        if bname == "kuka":
            robot = (links, dfs)
        for li, lname in enumerate(dfs):
            L = links[lname]
            link_index = li - 1       # base = -1
            if L.visual is None:
                continue
            # ---- render triangles (float32 pipeline of b3ImportMeshUtility + convertURDFToVisualShape)
            opath = os.path.join(mesh_dir, L.visual["file"])
            obj = parse_obj(opath)
            tris = triangulate(obj)                                   # float32
            scale = np.array(L.visual["scale"], float)
            pos = (tris[:, :, 0:3].astype(np.float64) * scale).astype(np.float32)   # v.xyzw[i] *= meshScale[i]
            vt = L.inertial.inv() * L.visual["origin"]                # localInertiaFrame.inverse()*childTrans
            pos = vt.apply(pos.astype(np.float64)).astype(np.float32)
            nrm = (tris[:, :, 3:6].astype(np.float64) @ vt.R.T).astype(np.float32)
            tris = np.concatenate([pos, nrm, tris[:, :, 6:8]], axis=2).astype(np.float32)
            tex = get_texture(mtl_texture(opath, obj["mtllib"]))
            rgba = materials.get(L.visual["material"], list(DEFAULT_RGBA))
            # ---- collision hull (double): verts*scale, into the inertial frame
            cpath = os.path.join(mesh_dir, L.collision["file"])
            cobj = parse_obj(cpath)
            cv = cobj["v"].astype(np.float32).astype(np.float64) * np.array(L.collision["scale"], float)
            ct = L.inertial.inv() * L.collision["origin"]
            cv = ct.apply(cv)
            hull = ConvexHull(cv)
            hv = cv[np.sort(hull.vertices)]
            # unique facet planes (merge coplanar simplices), outward normals
            eq = np.round(hull.equations, 9)
            eq = np.unique(eq, axis=0)
            fr = L.contact.get("lateral_friction", DEFAULT_LATERAL_FRICTION)
            shapes.append(dict(body=body_id, link=link_index, name=lname,
                               tri_start=sum(len(t) for t in tri_all), tri_count=len(tris),
                               hull_start=sum(len(h) for h in hull_all), hull_count=len(hv),
                               face_start=sum(len(h) for h in hull_face_all), face_count=len(eq),
                               tex=tex, rgba=rgba, friction=fr,
                               restitution=L.contact.get("restitution", 0.0),
                               rolling=L.contact.get("rolling_friction", 0.0),
                               spinning=L.contact.get("spinning_friction", 0.0),
                               aabb_min=(hv.min(0) - 0).tolist(), aabb_max=hv.max(0).tolist()))
            tri_all.append(tris)
            hull_all.append(hv)
            hull_face_all.append(eq)

    S = len(shapes)
    arrays["shape.body"] = np.array([s["body"] for s in shapes], np.int32)
    arrays["shape.link"] = np.array([s["link"] for s in shapes], np.int32)
    arrays["shape.tri_start"] = np.array([s["tri_start"] for s in shapes], np.int32)
    arrays["shape.tri_count"] = np.array([s["tri_count"] for s in shapes], np.int32)
    arrays["shape.hull_start"] = np.array([s["hull_start"] for s in shapes], np.int32)
    arrays["shape.hull_count"] = np.array([s["hull_count"] for s in shapes], np.int32)
    arrays["shape.face_start"] = np.array([s["face_start"] for s in shapes], np.int32)
    arrays["shape.face_count"] = np.array([s["face_count"] for s in shapes], np.int32)
    arrays["shape.tex"] = np.array([s["tex"] for s in shapes], np.int32)
    arrays["shape.rgba"] = np.array([s["rgba"] for s in shapes], np.float32)
    arrays["shape.friction"] = np.array([s["friction"] for s in shapes], np.float64)
    arrays["shape.restitution"] = np.array([s["restitution"] for s in shapes], np.float64)
    arrays["shape.rolling"] = np.array([s["rolling"] for s in shapes], np.float64)
    arrays["shape.spinning"] = np.array([s["spinning"] for s in shapes], np.float64)
    arrays["shape.aabb_min"] = np.array([s["aabb_min"] for s in shapes], np.float64)
    arrays["shape.aabb_max"] = np.array([s["aabb_max"] for s in shapes], np.float64)
    arrays["tri.data"] = np.concatenate(tri_all).astype(np.float32)          # [T,3,8]
    arrays["hull.verts"] = np.concatenate(hull_all).astype(np.float64)       # [V,3]
    arrays["hull.faces"] = np.concatenate(hull_face_all).astype(np.float64)  # [F,4] n.x+d<=0 inside
    arrays["tex.meta"] = np.array(tex_meta, np.int64)                        # [K,4] w,h,offset,uniform
    arrays["tex.data"] = np.concatenate(tex_data).astype(np.uint8)
    manifest["shapes"] = [dict(name=s["name"], body=s["body"], link=s["link"]) for s in shapes]

    # ---- robot tree (Bullet btMultiBody::setupRevolute / setupFixed arguments, URDF2Bullet.cpp [BULLET-MEM])
    links, dfs = robot
    nl = len(dfs) - 1
    parent = np.full(nl, -1, np.int32)
    jtype = np.zeros(nl, np.int32)
    dof = np.full(nl, -1, np.int32)
    axis = np.zeros((nl, 3))
    rot0 = np.zeros((nl, 9))
    dvec = np.zeros((nl, 3))
    evec = np.zeros((nl, 3))
    mass = np.zeros(nl)
    inertia_urdf = np.zeros((nl, 3))
    inertia_box = np.zeros((nl, 3))
    damping = np.zeros(nl)
    lower = np.zeros(nl)
    upper = np.zeros(nl)
    shape_of_link = {s["link"]: i for i, s in enumerate(shapes) if s["body"] == 0}
    ndof = 0
    for li in range(nl):
        L = links[dfs[li + 1]]
        J = L.joint
        P = links[L.parent]
        parent[li] = dfs.index(L.parent) - 1
        offA = P.inertial.inv() * J["origin"]          # offsetInA
        offB = L.inertial.inv()                        # offsetInB
        # parentRotToThis = offsetInB.rot * offsetInA.inverse().rot : maps parent-frame vectors to child-frame vectors
        rot0[li] = (offB.R @ offA.R.T).reshape(-1)
        dvec[li] = offA.t
        evec[li] = -offB.t
        if J["type"] == "revolute":
            jtype[li] = 1
            dof[li] = ndof
            ndof += 1
            axis[li] = offB.R @ np.array(J["axis"], float)
            damping[li] = J["damping"]
            lower[li], upper[li] = J["lower"], J["upper"]
        else:
            assert J["type"] == "fixed"
        mass[li] = L.mass
        inertia_urdf[li] = L.inertia_urdf
        # btCompoundShape::calculateLocalInertia: box inertia of the AABB (hull verts +- 2*margin) [BULLET-MEM]
        s = shapes[shape_of_link[li]]
        ext = (np.array(s["aabb_max"]) - np.array(s["aabb_min"])) + 4 * COLLISION_MARGIN
        lx, ly, lz = ext
        inertia_box[li] = L.mass / 12.0 * np.array([ly * ly + lz * lz, lx * lx + lz * lz, lx * lx + ly * ly])
        manifest["robot_links"].append(L.name)
        manifest["robot_joints"].append(J["name"])
    arrays["robot.parent"] = parent
    arrays["robot.jtype"] = jtype
    arrays["robot.dof"] = dof
    arrays["robot.axis"] = axis
    arrays["robot.rot0"] = rot0
    arrays["robot.dvec"] = dvec
    arrays["robot.evec"] = evec
    arrays["robot.mass"] = mass
    arrays["robot.inertia_urdf"] = inertia_urdf
    arrays["robot.inertia_box"] = inertia_box
    arrays["robot.damping"] = damping
    arrays["robot.lower"] = lower
    arrays["robot.upper"] = upper
    arrays["robot.base_pos"] = np.array(ROBOT_POSITION, float)
    arrays["robot.shape_of_link"] = np.array([shape_of_link[i] for i in range(-1, nl)], np.int32)  # [nl+1], base first

    # ---- table (static, 2 shapes) and objects (free bodies)
    tl, tdfs, _ = parse_urdf(os.path.join(data, "urdf", "table.urdf"))
    arrays["table.pos"] = np.array(OBJECT_POSES["table"][:3], float)
    arrays["table.quat"] = euler_to_quat(OBJECT_POSES["table"][3:])
    # child link COM offset relative to table base frame (fixed joint origin 0 + inertial origin)
    arrays["table.upper_offset"] = (tl["table_upper"].joint["origin"] * tl["table_upper"].inertial).t
    obj_names = ["cube", "tomato", "mustard"]
    omass, oinertia, opose = [], [], []
    for n in obj_names:
        ol, odfs, _ = parse_urdf(os.path.join(data, "urdf", n + ".urdf"))
        L = ol[odfs[0]]
        assert np.allclose(L.inertial.R, np.eye(3)) and np.allclose(L.inertial.t, 0)
        omass.append(L.mass)
        oinertia.append(L.inertia_urdf)          # URDF_USE_INERTIA_FROM_FILE (robot.py:222)
        opose.append(np.concatenate([OBJECT_POSES[n][:3], euler_to_quat(OBJECT_POSES[n][3:])]))
    arrays["obj.mass"] = np.array(omass, float)
    arrays["obj.inertia"] = np.array(oinertia, float)
    arrays["obj.init_pose"] = np.array(opose, float)         # [3,7] xyz + quat xyzw
    arrays["obj.shape"] = np.array([i for i, s in enumerate(shapes) if s["body"] >= 2], np.int32)
    arrays["table.shape"] = np.array([i for i, s in enumerate(shapes) if s["body"] == 1], np.int32)
    manifest["objects"] = obj_names
    manifest["collision_margin"] = COLLISION_MARGIN
    arrays["manifest.json"] = np.frombuffer(json.dumps(manifest, sort_keys=True).encode(), np.uint8)

    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    write_blob(args.out, arrays)
    print("wrote %s: %d shapes, %d tris, %d hull verts, %d textures (%.1f MB), %.1f MB total" % (
        args.out, S, len(arrays["tri.data"]), len(arrays["hull.verts"]), len(tex_meta),
        arrays["tex.data"].nbytes / 1e6, os.path.getsize(args.out) / 1e6))


if __name__ == "__main__":
    main()
