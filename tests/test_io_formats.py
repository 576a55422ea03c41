"""CPU tests of the host-side file formats and pose conversions around the path (no GPU, no compute calls into
libi3d_b200): .tsdf grids (src/sparse_voxel_grid.cpp:484-549), intrinsics text files (src/camera.cpp:202-274), TUM trajectory files
(src/rgbd/sensor.cpp:236-347), pose vector <-> matrix (src/math.cpp:151-178).  The .tsdf byte layout is checked against an independent
numpy writer / parser of the reference's raw-struct dump."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host():
    path = os.path.join(ROOT, "intrinsic3d_b200", "libi3d_host.so")
    if not os.path.exists(path):
        import __graft_entry__ as g
        g.build()
    return C.CDLL(path)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


VOXEL_REC = np.dtype([("xyz", "<i4", 3), ("sdf", "<f4"), ("weight", "<f4"), ("color", "u1", 3), ("pad", "u1")])                  # 12 + 12 bytes
SBR_REC = np.dtype([("xyz", "<i4", 3), ("sdf", "<f8"), ("weight", "<f4"), ("color", "u1", 3), ("pad", "u1"), ("albedo", "<f8"), ("sdf_refined", "<f8")])
HEADER = np.dtype([("voxel_size", "<f4"), ("truncation", "<f4"), ("iws", "<f4"), ("size", "<u8"), ("mlf", "<f4")])             # packed: 24 bytes


def _grid(n=500, seed=0):
    rng = np.random.default_rng(seed)
    xyz = np.unique(rng.integers(-40, 40, (n, 3)), axis=0).astype(np.int32)
    rng.shuffle(xyz)
    n = len(xyz)
    sdf = rng.normal(0, 0.01, n).astype(np.float32)
    weight = rng.uniform(0, 3, n).astype(np.float32)
    weight[rng.choice(n, n // 10, replace=False)] = 0.0
    rgb = rng.integers(0, 256, (n, 3)).astype(np.uint8)
    return xyz, sdf, weight, rgb


def test_tsdf_roundtrip_and_byte_layout(host, tmp_path):
    assert VOXEL_REC.itemsize == 24 and SBR_REC.itemsize == 44 and HEADER.itemsize == 24
    xyz, sdf, weight, rgb = _grid()
    n = len(xyz)
    p1, p2 = str(tmp_path / "a.tsdf").encode(), str(tmp_path / "b.tsdf").encode()
    nvalid = C.c_int64(0)
    rc = host.i3dh_io_tsdf_roundtrip(p1, p2, C.c_int64(n), _p(xyz, C.c_int32), _p(sdf, C.c_float), _p(weight, C.c_float), _p(rgb, C.c_uint8), C.c_float(0.004),
                                     C.byref(nvalid))
    assert rc == 0
    assert nvalid.value == int((weight > 0).sum())                     # convert() drops invalid voxels
    # independent parse of what the C++ writer produced
    raw = open(p1, "rb").read()
    hd = np.frombuffer(raw[:24], HEADER)[0]
    assert hd["voxel_size"] == np.float32(0.004) and hd["truncation"] == np.float32(np.float32(0.004) * np.float32(5.0)) and hd["size"] == n
    rec = np.frombuffer(raw[24:], VOXEL_REC)
    assert len(rec) == n and np.array_equal(rec["xyz"], xyz) and np.array_equal(rec["sdf"], sdf) and np.array_equal(rec["weight"], weight)
    assert np.array_equal(rec["color"], rgb) and np.all(rec["pad"] == 0)
    raw2 = open(p2, "rb").read()
    rec2 = np.frombuffer(raw2[24:], SBR_REC)
    keep = weight > 0
    assert np.array_equal(rec2["xyz"], xyz[keep]) and np.array_equal(rec2["sdf"], sdf[keep].astype(np.float64))
    assert np.array_equal(rec2["sdf_refined"], sdf[keep].astype(np.float64)) and np.all(rec2["albedo"] == 0.6)
    assert np.array_equal(rec2["color"], rgb[keep])


def test_tsdf_load_of_foreign_file(host, tmp_path):
    xyz, sdf, weight, rgb = _grid(seed=1)
    n = len(xyz)
    path = str(tmp_path / "ref.tsdf")
    hd = np.zeros(1, HEADER)
    hd["voxel_size"], hd["truncation"], hd["iws"], hd["size"], hd["mlf"] = 0.008, 0.04, 10.0, n, 0.6
    rec = np.zeros(n, VOXEL_REC)
    rec["xyz"], rec["sdf"], rec["weight"], rec["color"] = xyz, sdf, weight, rgb
    rec["pad"] = 0xAB                                                    # the reference writes uninitialised padding
    with open(path, "wb") as f:
        f.write(hd.tobytes()); f.write(rec.tobytes())
    m = C.c_int64(0)
    hdr = np.zeros(3, np.float32)
    oxyz, osdf, ow, orgb = np.zeros((n, 3), np.int32), np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros((n, 3), np.uint8)
    rc = host.i3dh_io_tsdf_load(path.encode(), C.c_int64(n), C.byref(m), _p(hdr, C.c_float), _p(oxyz, C.c_int32), _p(osdf, C.c_float), _p(ow, C.c_float), _p(orgb, C.c_uint8))
    assert rc == 0 and m.value == n
    assert hdr[0] == np.float32(0.008) and hdr[1] == np.float32(0.04)
    assert np.array_equal(oxyz, xyz) and np.array_equal(osdf, sdf) and np.array_equal(ow, weight) and np.array_equal(orgb, rgb)
    # truncated file is an error, not a partial grid
    with open(path, "wb") as f:
        f.write(hd.tobytes()); f.write(rec.tobytes()[:-7])
    assert host.i3dh_io_tsdf_load(path.encode(), C.c_int64(n), C.byref(m), _p(hdr, C.c_float), _p(oxyz, C.c_int32), _p(osdf, C.c_float), _p(ow, C.c_float), _p(orgb, C.c_uint8)) == 1
    assert host.i3dh_io_tsdf_load(b"/nonexistent/x.tsdf", C.c_int64(n), C.byref(m), _p(hdr, C.c_float), _p(oxyz, C.c_int32), _p(osdf, C.c_float), _p(ow, C.c_float), _p(orgb, C.c_uint8)) == 1


def test_intrinsics_and_pose_files(host, tmp_path):
    from intrinsic3d_b200.scene import aa_to_rotation
    ip, pp = str(tmp_path / "intrinsics.txt"), str(tmp_path / "poses.txt")
    with open(ip, "w") as f:
        f.write("640 480\n577.871 0 319.623\n0 580.258 239.624\n0 0 1\n0.01 -0.002 0.0003 0.0001 -0.0002\n")
    rng = np.random.default_rng(2)
    F = 12
    poses = np.concatenate([rng.normal(0, 0.8, (F, 3)), rng.normal(0, 0.5, (F, 3))], 1)
    poses[0, :3] = 0.0                                                   # identity rotation (angle 0)
    poses[1, :3] = np.array([3.1, 0.0, 0.0])                             # close to pi: the negative-w / small-trace branch
    poses[2, :3] = np.array([0.0, 2.0, 2.2])
    poses = np.ascontiguousarray(poses)
    out = np.zeros(14, np.float64)
    rc = host.i3dh_io_camera_poses(ip.encode(), pp.encode(), C.c_int32(F), _p(poses, C.c_double), _p(out, C.c_double))
    assert rc == 0
    assert np.allclose(out[:4], [577.871, 580.258, 319.623, 239.624], rtol=1e-6) and np.allclose(out[4:9], [0.01, -0.002, 0.0003, 0.0001, -0.0002], rtol=1e-6)
    assert (out[9], out[10]) == (640, 480)
    assert out[11] < 1e-9 and out[12] < 1e-14          # angle-axis round trip (angles < pi), orthonormal rotation
    assert out[13] < 2e-6                              # 6 decimals + float quaternion in the trajectory file
    # the file is TUM format: timestamp tx ty tz qx qy qz qw, camera-to-world
    rows = np.loadtxt(pp)
    assert rows.shape == (F, 8)
    for f in range(F):
        R = aa_to_rotation(poses[f, :3])
        t = poses[f, 3:]
        c2w_t = -R.T @ t
        assert np.allclose(rows[f, 1:4], c2w_t, atol=2e-6)
        qx, qy, qz, qw = rows[f, 4:]
        Rq = np.array([[1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw), 2 * (qx * qz + qy * qw)],
                       [2 * (qx * qy + qz * qw), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw)],
                       [2 * (qx * qz - qy * qw), 2 * (qy * qz + qx * qw), 1 - 2 * (qx * qx + qy * qy)]])
        assert np.allclose(Rq, R.T, atol=5e-6)
    # a missing intrinsics file falls back to the reference's defaults
    assert host.i3dh_io_camera_poses(b"/nonexistent/i.txt", pp.encode(), C.c_int32(F), _p(poses, C.c_double), _p(out, C.c_double)) == 1
