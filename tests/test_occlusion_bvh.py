"""The order-free occlusion structure of the scene blob (csrc/host/ah_bvh.h -> LSEC_AH_NODES / LSEC_AH_PRIMS) and the
per-object node chains the confirming traversal walks (LSEC_OBJ_PATH).  CPU: structure and conservativeness.  The
booleans themselves are compared with the oracle on the GPU (tests/test_trace_parity.py)."""
import numpy as np
import pytest
from conftest import small_scene, SMALL
from lumo_b200 import native

LEAF = 0x80000000
NONE = 0xFFFFFFFF


def _prim_world_boxes(B):
    """f64 world boxes of every (triangle | sphere, object) primitive, straight from the blob's f64 data."""
    prims = B.ah_prims
    obj = prims["obj"] & 0x7FFFFFFF
    inst = B.objects["inst"][obj]
    assert np.array_equal((prims["obj"] & 0x80000000) != 0, inst >= 0)
    lo = np.empty((len(prims), 3)); hi = np.empty((len(prims), 3))
    is_sph = (prims["tri"] & 0x80000000) != 0
    tri = prims["tri"] & 0x7FFFFFFF
    tv = B.tri_verts[np.where(is_sph, 0, tri)]
    V = np.stack([tv["a"], tv["b"], tv["c"]], 1)                       # [n, 3, 3] local vertices
    M = np.zeros((len(prims), 3, 4)); M[:, 0, 0] = M[:, 1, 1] = M[:, 2, 2] = 1.0
    has = inst >= 0
    M[has] = B.instances["m"][inst[has]].reshape(-1, 3, 4)
    W = np.einsum("nij,nkj->nki", M[:, :, :3], V) + M[:, None, :, 3]   # world vertices
    lo[:] = W.min(1); hi[:] = W.max(1)
    if is_sph.any():
        r = B.spheres["radius"][tri[is_sph]]
        c = M[is_sph][:, :, 3]                                            # spheres sit at their frame's origin
        s = np.abs(M[is_sph][:, :, :3]).sum(2) * r[:, None]               # box of the transformed sphere's local cube: a superset
        lo[is_sph] = c - s; hi[is_sph] = c + s
    return lo, hi, is_sph


@pytest.mark.parametrize("name", list(SMALL))
def test_structure_and_containment(name):
    prog, blob, _ = small_scene(name)
    B = native.Blob(blob)
    nodes, prims = B.ah_nodes, B.ah_prims
    n_obj = int(B.params["n_objects"]) + int(B.params["n_lights"])
    # every primitive of every object exactly once
    want = 0
    for o in B.objects:
        k = int(o["kind"])
        want += int(B.kd_trees[int(o["geom"])]["n_tris"]) if k in (0, 1) else 1
    assert len(prims) == want
    key = prims["tri"].astype(np.uint64) << np.uint64(32) | (prims["obj"] & 0x7FFFFFFF).astype(np.uint64)
    assert len(np.unique(key)) == len(prims)
    assert (prims["obj"] & 0x7FFFFFFF).max() < n_obj
    lo, hi, is_sph = _prim_world_boxes(B)
    # walk the tree: child boxes contain everything below them; every primitive is reachable exactly once
    seen = np.zeros(len(prims), np.int32)
    stack = [(0, np.full(3, -np.inf), np.full(3, np.inf))]
    n_visited = 0
    while stack:
        ni, plo, phi = stack.pop()
        nd = nodes[ni]; n_visited += 1
        for k in range(4):
            c = int(nd["child"][k])
            clo = np.array([nd["lo_x"][k], nd["lo_y"][k], nd["lo_z"][k]], np.float64); chi = np.array([nd["hi_x"][k], nd["hi_y"][k], nd["hi_z"][k]], np.float64)
            if c == NONE:
                assert (clo > chi).all(); continue
            assert (clo >= plo).all() and (chi <= phi).all(), "child box sticks out of its parent's"
            if c & LEAF:
                first, cnt = c & 0x07FFFFFF, ((c >> 27) & 0xF) + 1
                assert 1 <= cnt <= 4
                for p in range(first, first + cnt):
                    seen[p] += 1
                    if not is_sph[p]:      # outward rounding: the f32 box strictly contains the f64 vertices
                        assert (clo <= lo[p]).all() and (chi >= hi[p]).all(), "leaf box does not contain its primitive"
                    else:
                        assert (clo <= lo[p] + 1e-6 * np.abs(lo[p])).all() and (chi >= hi[p] - 1e-6 * np.abs(hi[p])).all()
            else:
                assert ni < c < len(nodes)
                stack.append((c, clo, chi))
    assert n_visited == len(nodes) and (seen == 1).all()
    # node chains: root of the right BVH first, a leaf that lists the object last, each step a child of the previous node
    off, path = B.obj_path_off, B.obj_path
    assert len(off) == n_obj + 1 and off[0] == 0 and off[-1] == len(path)
    lr = int(B.params["lights_root"]); n0 = int(B.params["n_objects"])
    for g in range(n_obj):
        ch = path[off[g]:off[g + 1]]
        root = 0 if g < n0 else lr
        assert len(ch) >= 1 and ch[0] == root
        for a, b in zip(ch[:-1], ch[1:]):
            nd = B.tlas_nodes[a]
            assert nd["count"] == 0 and (b == a + 1 or (nd["right"] != NONE and b == root + nd["right"]))
        leaf = B.tlas_nodes[ch[-1]]
        listed = B.tlas_leaf[int(leaf["first"]):int(leaf["first"]) + int(leaf["count"])]
        assert (g - (0 if g < n0 else n0)) in listed.tolist()


def _f32(x): return np.asarray(x, np.float32)


def _fma32(a, b, c):
    """fmaf for f32 inputs: the product of two f32 is exact in f64; one rounding to f32 at the end (up to a double rounding the test's margins dwarf)."""
    return _f32(np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64))


def test_slab_test_never_rejects_a_box_the_ray_enters():
    """The f32 slab test of occlude.cuh (ah_make_ray / ah_slab) restated in numpy: for random rays and random boxes it must
    say `enter` whenever the exact (f64) ray enters the box within [0, t_max] — the other way round is only wasted work."""
    rs = np.random.RandomState(3)
    n = 400000
    scale = 10.0 ** rs.uniform(-1, 3, (n, 1))
    o = rs.uniform(-1, 1, (n, 3)) * scale
    d = rs.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[rs.rand(n) < 0.1, 0] = 0.0; d[rs.rand(n) < 0.05, 1] = 1e-25; d /= np.linalg.norm(d, axis=1, keepdims=True)
    # boxes built around points ON the ray so that many tests are grazing ones
    t_in = rs.uniform(0, 1, (n, 1)) * scale * 2
    p = o + t_in * d
    half = np.abs(rs.normal(size=(n, 3))) * scale * 10.0 ** rs.uniform(-7, 0, (n, 1))
    half[rs.rand(n) < 0.2, rs.randint(0, 3)] = 0.0                        # flat boxes
    shift = rs.uniform(-1, 1, (n, 3)) * half * (rs.rand(n, 1) < 0.5)       # half of them with the point strictly inside
    lo64, hi64 = p + shift - half, p + shift + half
    t_max = t_in[:, 0] * rs.choice([0.5, 1.0, 1.0 + 1e-12, 2.0, np.inf], n)
    # exact verdict in f64 (generous: a box is entered if the slab intervals overlap [0, t_max] with a few ulp of slack)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d
        t1, t2 = (lo64 - o) * inv, (hi64 - o) * inv
    tn = np.nanmax(np.minimum(t1, t2), axis=1); tf = np.nanmin(np.maximum(t1, t2), axis=1)
    inside_axis = (d == 0) & (o >= lo64) & (o <= hi64)
    enters = (np.maximum(tn, 0.0) <= np.minimum(tf, t_max)) & ~np.any((d == 0) & ~inside_axis, axis=1)
    # the kernel's boxes: f32, rounded outwards with the builder's padding (ah_bvh.h: ah_box)
    pad = np.maximum(np.abs(lo64), np.abs(hi64)) / 4194304.0 + 1e-30
    lo = _f32(lo64 - pad); lo = np.where(lo.astype(np.float64) > lo64 - pad, np.nextafter(lo, _f32(-np.inf)), lo)
    hi = _f32(hi64 + pad); hi = np.where(hi.astype(np.float64) < hi64 + pad, np.nextafter(hi, _f32(np.inf)), hi)
    # ah_make_ray
    df = _f32(d); small = np.abs(df) < _f32(1e-18)
    df = np.where(small, np.where(np.signbit(d), _f32(-1e-18), _f32(1e-18)), df)
    invf = _f32(1.0) / df
    pf = _f32(_f32(o).astype(np.float64) * invf.astype(np.float64))
    s = _fma32(np.abs(pf), _f32(2.0 ** -22), _f32(1e-30))
    near_c = -pf - s; far_c = -pf + s
    tm = _f32(t_max); tm = np.where(tm.astype(np.float64) < t_max, np.nextafter(tm, _f32(np.inf)), tm)
    near_plane = np.where(df < 0, hi, lo); far_plane = np.where(df < 0, lo, hi)
    tnf = np.maximum(_fma32(near_plane, invf, near_c).max(axis=1), _f32(0.0))
    tff = np.minimum(_fma32(far_plane, invf, far_c).min(axis=1), tm)
    says = tnf <= _f32(tff.astype(np.float64) * (1.0 + 2.0 ** -20))
    assert enters.sum() > n // 4
    assert not (enters & ~says).any(), "the f32 slab test rejected a box the f64 ray enters: %d cases" % int((enters & ~says).sum())
