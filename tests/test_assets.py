"""The asset blob (real_robots_b200/assets/scene.rrb) against SURVEY App. A, and — where /root/reference exists
(this container, not the GPU box) — against a fresh run of tools/compile_assets.py on the reference's data directory.

The blob is the only thing the CUDA library and the oracle read of the reference's URDF / OBJ / PNG files, so its
contents are part of the parity argument: body ids are mask values (robot.py:49-50 load order), link order is
pybullet's, joint limits are the action box (robot.py:58-67).
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from real_robots_b200.blob import read_blob

ROOT = os.path.join(os.path.dirname(__file__), "..")
BLOB = os.path.join(ROOT, "real_robots_b200", "assets", "scene.rrb")


@pytest.fixture(scope="module")
def blob():
    return read_blob(BLOB)


def test_bodies_links_and_triangle_counts_match_survey(blob):
    man = json.loads(bytes(blob["manifest.json"]).decode())
    assert [b["name"] for b in man["bodies"]] == ["kuka", "table", "cube", "tomato", "mustard"]     # ids 0..4 = mask values
    assert man["bodies"][1]["links"] == ["table_base", "table_upper"]
    assert man["robot_links"][7:] == ["base", "finger_00", "finger_01", "skin_01", "skin_00",
                                      "finger_10", "finger_11", "skin_11", "skin_10"]
    # SURVEY App. A.5 (fan-triangulated n-gons): per-shape triangle counts and the scene totals for 1 / 2 / 3 objects
    tc = blob["shape.tri_count"]
    assert tc[:9].tolist() == [3038, 2759, 1449, 1938, 1547, 1358, 1157, 1512, 136]
    assert tc[17:].tolist() == [60, 12, 768, 44, 176]
    assert tc[:17].sum() == 16518 and tc[17:19].sum() == 72
    assert [int(tc[:20 + k].sum()) for k in range(3)] == [17358, 17402, 17578]
    assert blob["tri.data"].shape == (17578, 3, 8)
    assert (np.diff(blob["shape.tri_start"]) == tc[:-1]).all()
    # hull vertex counts (scipy ConvexHull of the collision meshes)
    hc = blob["shape.hull_count"]
    assert hc[:9].tolist() == [188, 574, 296, 362, 315, 253, 346, 327, 68] and hc[17:].tolist() == [14, 8, 152, 24, 85]


def test_hull_planes_enclose_their_vertices_and_normals_are_unit(blob):
    for s in range(22):
        v = blob["hull.verts"][blob["shape.hull_start"][s]:][:blob["shape.hull_count"][s]]
        f = blob["hull.faces"][blob["shape.face_start"][s]:][:blob["shape.face_count"][s]]
        assert np.allclose(np.linalg.norm(f[:, :3], axis=1), 1, atol=1e-12)
        d = v @ f[:, :3].T + f[:, 3]
        assert d.max() < 1e-9                       # every vertex behind or on every facet plane
        assert (np.abs(d).min(0) < 1e-9).all()      # every facet touches a vertex
        assert (v.min(0) >= blob["shape.aabb_min"][s] - 1e-12).all() and (v.max(0) <= blob["shape.aabb_max"][s] + 1e-12).all()


def test_kinematic_tree_limits_and_object_constants(blob):
    parent = blob["robot.parent"]
    assert parent[:8].tolist() == [-1, 0, 1, 2, 3, 4, 5, 6]           # serial arm, then the gripper base on link 7
    assert (parent[1:] < np.arange(1, 16)).all()                       # parents precede children (pybullet link order)
    dof = blob["robot.dof"]
    assert sorted(d for d in dof.tolist() if d >= 0) == list(range(11)) and (dof < 0).sum() == 5   # 11 revolute, 5 fixed
    assert np.allclose(blob["robot.damping"][dof >= 0], 0.5)          # kuka_gripper.urdf <dynamics damping="0.5"/>
    assert np.allclose(blob["robot.base_pos"], [-0.55, 0.0, -0.04])   # robot.py:46
    assert np.allclose(blob["obj.mass"], [1.5, 3.0, 2.0])
    # robot.py:19-24 initial poses (xyz; Euler -> quat done by the compiler)
    assert np.allclose(blob["obj.init_pose"][:, :3], [[-0.10, 0.00, 0.45], [-0.10, -0.30, 0.45], [-0.10, 0.30, 0.45]], atol=1e-12)
    assert np.allclose(blob["obj.init_pose"][:2, 3:], [0, 0, 0, 1])
    # mustard: rpy (1.5708, 3.14159, 0) -> the bottle lies on its side, rotated by pi about its long axis
    x, y, z, w = blob["obj.init_pose"][2, 3:]
    assert abs(abs(y) - np.sqrt(0.5)) < 1e-5 and abs(abs(z) - np.sqrt(0.5)) < 1e-5 and abs(x) < 1e-5 and abs(w) < 1e-5
    assert np.allclose(np.linalg.norm(blob["obj.init_pose"][:, 3:], axis=1), 1)
    # friction tags: table 1.0 (table.urdf:5,31), everything else the importer default 0.5
    fr = blob["shape.friction"]
    assert fr[17] == 1.0 and fr[18] == 1.0 and (np.delete(fr, [17, 18]) == 0.5).all()
    assert blob["shape.rolling"][19:].tolist() == [1e-4, 1e-3, 1e-2] and blob["shape.spinning"][19:].tolist() == [1e-4, 1e-3, 1e-2]


def test_textures(blob):
    meta = blob["tex.meta"]                       # w, h, offset, uniform flag
    # the four single-colour 800x600 robot textures are stored as one texel
    assert (meta[:4, :2] == [800, 600]).all() and (meta[:4, 3] == 1).all()
    texels = {tuple(int(c) for c in blob["tex.data"][int(o):int(o) + 3]) for o in meta[:4, 2]}
    assert texels == {(25, 25, 25), (255, 75, 0), (25, 75, 255), (151, 151, 151)}
    w, h, off, uni = meta[-1]
    assert uni == 0 and off + 3 * w * h == blob["tex.data"].size
    tex = blob["shape.tex"]
    assert tex[11] == -1 and tex[12] == -1 and (tex[[0, 17, 19, 20, 21]] >= 0).all()     # skin_0 has no texture


@pytest.mark.skipif(not os.path.isdir("/root/reference/real_robots/data"), reason="reference tree not present (GPU box)")
def test_blob_is_reproduced_byte_for_byte_from_the_reference_data(tmp_path):
    out = tmp_path / "scene.rrb"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "compile_assets.py"),
                           "--ref", "/root/reference", "--out", str(out)], stdout=subprocess.DEVNULL)
    assert out.read_bytes() == open(BLOB, "rb").read()
