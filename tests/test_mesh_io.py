"""Wavefront OBJ reader / writer as product code: fclgpu_load_obj / fclgpu_save_obj mirror
loadOBJFile / saveOBJFile of the reference (test/test_fcl_utility.h:194-309), quirks included.  The fixtures of the
benchmark (tests/golden/env.npz, rob.npz = the reference's env.obj / rob.obj re-encoded) must survive a save / load
round trip bit for bit, and the product loader must agree with the oracle's own restatement of the loader."""
import numpy as np

import fcl_b200 as F
from fcl_b200 import _capi


def test_save_load_round_trip_is_exact(tmp_path, env_rob_npz):
    for k, (v, t) in enumerate(env_rob_npz):
        p = tmp_path / f"m{k}.obj"
        F.saveOBJFile(p, v, t)
        v2, t2 = F.loadOBJFile(p)
        assert v2.tobytes() == np.ascontiguousarray(v, np.float64).tobytes()
        assert np.array_equal(t2, t)


def test_loader_agrees_with_the_oracle_loader_and_builds_the_same_model(tmp_path, env_rob_npz, oracle):
    (v, t), _ = env_rob_npz
    p = tmp_path / "env.obj"
    # the fixture's layout: a non-standard first line that the loader must skip, CRLF line ends, tabs
    with open(p, "w", newline="") as f:
        f.write("%d %d\r\n# comment\r\n" % (len(v), len(t)))
        for x in v:
            f.write("v\t%.17g %.17g %.17g\r\n" % tuple(x))
        for tri in t:
            f.write("f %d %d %d\r\n" % tuple(tri + 1))
    v2, t2 = F.loadOBJFile(p)
    assert v2.tobytes() == np.ascontiguousarray(v, np.float64).tobytes() and np.array_equal(t2, t)
    om = oracle.Model.__new__(oracle.Model)
    om.h = oracle.lib().orc_model_from_obj(str(p).encode(), 0)
    import ctypes as C

    nv, nt, nn = C.c_int(), C.c_int(), C.c_int()
    oracle.lib().orc_model_counts(om.h, C.byref(nv), C.byref(nt), C.byref(nn))
    om.num_vertices, om.num_tris, om.num_bvs = nv.value, nt.value, nn.value
    assert (nv.value, nt.value) == (len(v), len(t))
    m = F.BVHModel.from_obj(p)
    got, ref = m.node_arrays(), om.arrays()
    assert np.array_equal(got["first_child"], ref["first_child"])
    for k in ("axis", "obb_To", "obb_ext", "rss_To", "rss_l", "rss_r"):
        assert got[k].tobytes() == ref[k].tobytes(), k


def test_reference_quirks(tmp_path):
    # slashes, normals / textures remembered, polygons fanned only once a vn / vt line has been seen
    p = tmp_path / "q.obj"
    p.write_text("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nf 1 2 3 4\nvn 0 0 1\nf 1/1/1 2/2/1 3/3/1 4/4/1\ng grp\nvt 0 0\n")
    v, t = F.loadOBJFile(p)
    assert v.shape == (4, 3)
    # before the vn line: the reference's no-normal branch emits (0, 1, 2) once per fan step; afterwards a real fan
    assert t.tolist() == [[0, 1, 2], [0, 1, 2], [0, 1, 2], [0, 2, 3]]


def test_missing_file(tmp_path, capfd):
    v, t = F.loadOBJFile(tmp_path / "nope.obj")
    assert len(v) == 0 and len(t) == 0
    assert "file not exist" in capfd.readouterr().err
    L = _capi.lib()
    assert L.fclgpu_load_obj(None, None, None, None, None) == _capi.ERR_INVALID_ARGUMENT
