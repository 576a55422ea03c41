"""CPU known-answer tests for the grid-level transitions restated in oracle.cpp (SDFAlgorithms::clearVoxelsOutsideThinShell and
upsample<VoxelSBR>).  The reference has no tests for them; the oracle is pinned by independent restatements
(python sets for the pruning rule, vectorised numpy float32 for the interpolation) and by properties."""
import numpy as np
import pytest

f32 = np.float32


@pytest.fixture(scope="module")
def scene():
    from intrinsic3d_b200.scene import make_scene
    s = make_scene(radius_vox=10.0, frames=2, width=64, height=48, band=4.0, seed=2)
    rng = np.random.default_rng(5)
    s["weight"] = s["weight"].copy()
    s["weight"][rng.choice(len(s["weight"]), 300, replace=False)] = 0.0        # invalid voxels take part in both rules
    return s


def _oracle(s):
    import oracle
    o = oracle.Oracle(threads=2)
    o.load_scene(s)
    return o


def _prune_sets(s, thres):
    idx = {tuple(c): i for i, c in enumerate(s["xyz"])}
    sdf, w = s["sdf_refined"], s["weight"]
    keep = set()
    ring = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1), (2, 0, 0), (0, 2, 0), (0, 0, 2)]
    for i, c in enumerate(s["xyz"]):
        if w[i] > 0 and abs(sdf[i]) <= thres:
            keep.add(i)
            for d in ring:
                j = idx.get((c[0] + d[0], c[1] + d[1], c[2] + d[2]))
                if j is not None:
                    keep.add(j)
    offs = [(dx, dy, dz) for dz in range(-2, 3) for dy in range(-2, 3) for dx in range(-2, 3) if (dx, dy, dz) != (0, 0, 0)]
    out = set(keep)
    for i, c in enumerate(s["xyz"]):
        if i in keep:
            continue
        neg = sdf[i] < 0
        for d in offs:
            j = idx.get((c[0] + d[0], c[1] + d[1], c[2] + d[2]))
            if j is not None and ((sdf[j] >= 0) if neg else (sdf[j] < 0)):
                out.add(i)
                break
    return np.array(sorted(out))


@pytest.mark.parametrize("factor", [0.5, 1.0, 2.0])
def test_kg1_clear_voxels_outside_thin_shell(factor, scene):
    s = scene
    thres = factor * float(s["voxel_size"])
    o = _oracle(s)
    m = o.clear_voxels_outside_thin_shell(thres)
    want = _prune_sets(s, thres)
    g = o.grid()
    assert m == len(want) and 0 < m < len(s["xyz"])
    assert np.array_equal(g["xyz"], s["xyz"][want])                      # survivors keep their order
    for k in ("sdf0", "sdf_refined", "albedo", "weight", "rgb"):
        assert np.array_equal(g[k], s[k][want])
    assert g["voxel_size"] == f32(s["voxel_size"])
    # idempotent
    assert o.clear_voxels_outside_thin_shell(thres) == m


def _upsample_numpy(s):
    n = len(s["xyz"])
    idx = {tuple(c): i for i, c in enumerate(s["xyz"])}
    corners = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (0, 1, 1), (1, 0, 1), (1, 1, 1)]
    nb = np.full((n, 8), -1)
    for i, c in enumerate(s["xyz"]):
        for k, d in enumerate(corners):
            nb[i, k] = idx.get((c[0] + d[0], c[1] + d[1], c[2] + d[2]), -1)
    valid = (nb >= 0) & (s["weight"][np.clip(nb, 0, n - 1)] > 0)
    cnt = valid.sum(1)
    out = dict(xyz=np.zeros((8 * n, 3), np.int32), sdf0=np.zeros(8 * n), sdf_refined=np.zeros(8 * n), albedo=np.zeros(8 * n),
               weight=np.zeros(8 * n, f32), rgb=np.zeros((8 * n, 3), np.uint8))
    src = dict(sdf0=s["sdf0"].astype(f32), sdf_refined=s["sdf_refined"].astype(f32), albedo=s["albedo"].astype(f32), weight=s["weight"].astype(f32))
    col = s["rgb"].astype(f32)
    for z in range(2):
        for y in range(2):
            for x in range(2):
                t = (f32(0.5 * x), f32(0.5 * y), f32(0.5 * z))
                acc = {k: np.zeros(n, f32) for k in src}
                accc = np.zeros((n, 3), f32)
                sw = np.zeros(n, f32)
                for k, d in enumerate(corners):
                    w = f32(f32((t[0] if d[0] else f32(1) - t[0]) * (t[1] if d[1] else f32(1) - t[1])) * (t[2] if d[2] else f32(1) - t[2]))
                    ok = valid[:, k]
                    j = np.clip(nb[:, k], 0, n - 1)
                    for key in src:
                        acc[key] = np.where(ok, (acc[key] + (w * src[key][j]).astype(f32)).astype(f32), acc[key])
                    accc = np.where(ok[:, None], (accc + (w * col[j]).astype(f32)).astype(f32), accc)
                    sw = np.where(ok, (sw + w).astype(f32), sw)
                pos = sw > 0
                with np.errstate(divide="ignore", invalid="ignore"):
                    for key in src:
                        acc[key] = np.where(pos, (acc[key] / sw).astype(f32), acc[key])
                    accc = np.where(pos[:, None], (accc / sw[:, None]).astype(f32), accc)
                acc["weight"] = np.where(cnt <= 4, f32(0), np.maximum(acc["weight"], f32(0)))
                sl = slice(4 * z + 2 * y + x, 8 * n, 8)
                out["xyz"][sl] = 2 * s["xyz"] + np.array([x, y, z])
                for key in ("sdf0", "sdf_refined", "albedo"):
                    out[key][sl] = acc[key].astype(np.float64)
                out["weight"][sl] = acc["weight"]
                out["rgb"][sl] = np.trunc((accc + f32(0.5)).astype(f32)).astype(np.uint8)
    return out


def test_kg2_upsample_numpy_float32(scene):
    s = scene
    o = _oracle(s)
    m = o.upsample_grid()
    g = o.grid()
    want = _upsample_numpy(s)
    assert m == 8 * len(s["xyz"])
    assert g["voxel_size"] == f32(f32(s["voxel_size"]) * f32(0.5))
    for k in ("xyz", "sdf0", "sdf_refined", "albedo", "weight", "rgb"):
        assert np.array_equal(g[k], want[k]), k
    assert len(np.unique(g["xyz"], axis=0)) == m


def test_kg3_upsample_properties(scene):
    s = scene
    o = _oracle(s)
    o.upsample_grid()
    g = o.grid()
    n = len(s["xyz"])
    idx = {tuple(c): i for i, c in enumerate(s["xyz"])}
    full = np.array([all((idx.get((c[0] + dx, c[1] + dy, c[2] + dz), -1) >= 0 and s["weight"][idx[(c[0] + dx, c[1] + dy, c[2] + dz)]] > 0)
                         for dx in (0, 1) for dy in (0, 1) for dz in (0, 1)) for c in s["xyz"]])
    assert full.sum() > 1000
    # the (0,0,0) child of a voxel whose cube is fully valid is the voxel itself, rounded through float
    c0 = np.arange(n)[full] * 8
    assert np.array_equal(g["sdf_refined"][c0], s["sdf_refined"][full].astype(f32).astype(np.float64))
    assert np.array_equal(g["weight"][c0], s["weight"][full])
    assert np.array_equal(g["rgb"][c0], s["rgb"][full])
    # children of voxels with at most 4 valid corners carry weight 0 (they are invalid voxels of the fine grid)
    cnt = np.array([sum((idx.get((c[0] + dx, c[1] + dy, c[2] + dz), -1) >= 0 and s["weight"][idx[(c[0] + dx, c[1] + dy, c[2] + dz)]] > 0)
                        for dx in (0, 1) for dy in (0, 1) for dz in (0, 1)) for c in s["xyz"]])
    few = np.repeat(cnt <= 4, 8)
    assert few.sum() > 0 and np.all(g["weight"][few] == 0)
    assert (g["weight"][~few] > 0).mean() > 0.95         # (a child whose own weighted corners are all invalid still gets 0)
    # the fine grid then goes through the same pruning as every level (prepareGridLevel)
    m = o.clear_voxels_outside_thin_shell(2.0 * float(g["voxel_size"]))
    assert 0 < m < 8 * n


def test_kg4_degenerate_grids():
    """An isolated voxel: its 8 children exist but are invalid (1 valid corner <= 4); pruning a grid without sign change and without
    in-shell voxels removes everything (reported as an error, the grid is then empty)."""
    import oracle
    n = 1
    s = dict(xyz=np.array([[3, -2, 7]], np.int32), sdf0=np.array([0.001]), sdf_refined=np.array([0.001]), albedo=np.array([0.6]),
             weight=np.ones(1, np.float32), rgb=np.array([[10, 20, 30]], np.uint8), voxel_size=np.float32(0.004))
    o = oracle.Oracle(threads=1)
    o.set_grid(s)
    assert o.upsample_grid() == 8
    g = o.grid()
    assert np.all(g["weight"] == 0) and np.array_equal(g["xyz"][0], [6, -4, 14]) and np.array_equal(g["xyz"][7], [7, -3, 15])
    # the parent is the only valid corner and has a non-zero trilinear weight for every child (1, 1/2, 1/4 or 1/8): all children copy it
    assert np.all(g["sdf_refined"] == np.float64(np.float32(0.001))) and np.all(g["albedo"] == np.float64(np.float32(0.6)))
    assert np.all(g["rgb"] == np.array([10, 20, 30]))
    o2 = oracle.Oracle(threads=1)
    s2 = dict(s)
    s2["sdf_refined"] = np.array([0.5])
    o2.set_grid(s2)
    with pytest.raises(RuntimeError):
        o2.clear_voxels_outside_thin_shell(0.008)
