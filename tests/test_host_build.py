"""The product's native scene builder (csrc/host: kd-tree SAH build, object BVH, alias table, instance
matrices) against the oracle's independent restatement of the reference builders: node for node."""
import numpy as np
import pytest
import oracle_lib
from conftest import small_scene, SMALL
from lumo_b200 import native


@pytest.mark.parametrize("name", list(SMALL))
def test_blob_structures_equal_oracle(name):
    prog, blob, _ = small_scene(name)
    B = native.Blob(blob)
    O = oracle_lib.OracleScene(prog)
    P = B.params
    assert int(P["n_objects"]) == O.n_objects and int(P["n_lights"]) == O.n_lights and int(P["n_shadow_rays"]) == O.n_shadow_rays
    assert np.array_equal(np.concatenate([P["bounds_lo"], P["bounds_hi"]]), O.bounds())
    for which, root, nobj in ((0, 0, O.n_objects), (1, int(P["lights_root"]), O.n_lights)):
        e = O.export_bvh(which)
        n = len(e["right"])
        nodes = B.tlas_nodes[root:root + n]
        assert np.array_equal(np.concatenate([nodes["lo"], nodes["hi"]], axis=1), e["bounds"])
        assert np.array_equal(np.where(nodes["right"] == 0xFFFFFFFF, -1, nodes["right"].astype(np.int64)), e["right"])
        assert np.array_equal(nodes["count"].astype(np.int64), e["count"])
        leaf = np.concatenate([B.tlas_leaf[int(nd["first"]):int(nd["first"]) + int(nd["count"])] for nd in nodes]) if n else np.zeros(0)
        assert np.array_equal(leaf.astype(np.int64), e["leaf"][:len(leaf)])
    base = {0: 0, 1: O.n_objects}
    checked = 0
    for which, nobj in ((0, O.n_objects), (1, O.n_lights)):
        for i in range(min(nobj, 40)):
            obj = B.objects[base[which] + i]
            inst = O.export_instance(which, i)
            assert (inst is not None) == (int(obj["inst"]) >= 0)
            if inst is not None:
                I = B.instances[int(obj["inst"])]
                assert np.array_equal(I["m"].reshape(3, 4), inst[0][:3]) and np.array_equal(I["inv"].reshape(3, 4), inst[1][:3])
            kd = O.export_kd(which, i)
            if kd is None:
                assert int(obj["kind"]) in (2, 3)
                continue
            assert int(obj["kind"]) in (0, 1)
            T = B.kd_trees[int(obj["geom"])]
            n = len(kd["axis"])
            nodes = B.kd_nodes[int(T["root"]):int(T["root"]) + n]
            leaf = (nodes["b"] & 0x80000000) != 0
            assert np.array_equal(leaf.astype(np.int64), kd["leaf"])
            inner = ~leaf
            assert np.array_equal(nodes["point"][inner], kd["point"][inner])
            assert np.array_equal(nodes["b"][inner].astype(np.int64), kd["axis"][inner])
            assert np.array_equal(nodes["a"][inner].astype(np.int64) - int(T["root"]), kd["right"][inner])
            assert np.array_equal((nodes["b"][leaf] & 0x7FFFFFFF).astype(np.int64), kd["count"][leaf])
            ll = np.concatenate([B.kd_leaf[int(a):int(a) + int(b & 0x7FFFFFFF)] for a, b in zip(nodes["a"][leaf], nodes["b"][leaf])]) if leaf.any() else np.zeros(0)
            assert np.array_equal(ll.astype(np.int64), kd["leaf_list"])
            tv = B.tri_verts[int(T["tri_base"]):int(T["tri_base"]) + int(T["n_tris"])]
            assert np.array_equal(np.concatenate([tv["a"], tv["b"], tv["c"]], axis=1), kd["tri_verts"])
            checked += 1
    assert checked > 0
    prob, alias, pdf = O.export_alias()
    assert np.array_equal(B.lights["alias_prob"], prob) and np.array_equal(B.lights["alias"].astype(np.int64), alias) and np.array_equal(B.lights["pdf"], pdf)


def test_kd_leaf_completeness_on_blob():
    """kdtree_tests.rs:83-131 (`splits`, `contains`) on the flattened tree of the 10-triangle test cube
    (kdtree_tests.rs:158-194) and of the bunny stand-in: every triangle is listed by every leaf whose cell
    its bounding box overlaps, and by at least one leaf."""
    from lumo_b200 import meshes, Scene, Material, TriangleMesh, Rectangle, CameraBuilder, Spectrum
    for verts, faces in (meshes.cube10(), meshes.displaced_sphere(1200, seed=3)):
        s = Scene()
        s.add(TriangleMesh.new(verts, faces, [], [], Material.lambertian(Spectrum.from_rgb(0.5, 0.5, 0.5))))
        s.add_light(Rectangle((-1, 5, -1), (-1, 5, 1), (1, 5, 1), Material.light(Spectrum.from_rgb(1, 1, 1))))
        B = native.Blob(native.build_blob(s._program(CameraBuilder.new().resolution((16, 16)).build())))
        T = B.kd_trees[0]
        tv = B.tri_verts[int(T["tri_base"]):int(T["tri_base"]) + int(T["n_tris"])]
        tmin = np.minimum(np.minimum(tv["a"], tv["b"]), tv["c"]); tmax = np.maximum(np.maximum(tv["a"], tv["b"]), tv["c"])
        seen = np.zeros(len(tv), bool)
        stack = [(int(T["root"]), np.array(T["lo"]), np.array(T["hi"]))]
        while stack:
            ni, lo, hi = stack.pop()
            nd = B.kd_nodes[ni]
            if int(nd["b"]) & 0x80000000:
                idx = B.kd_leaf[int(nd["a"]):int(nd["a"]) + (int(nd["b"]) & 0x7FFFFFFF)]
                seen[idx] = True
                inside = np.all((tmin < hi) & (tmax > lo), axis=1)       # strict overlap with the cell
                assert set(np.nonzero(inside)[0]) <= set(int(i) for i in idx)
            else:
                ax, p = int(nd["b"]), float(nd["point"])
                lhi = hi.copy(); lhi[ax] = p; rlo = lo.copy(); rlo[ax] = p
                stack.append((ni + 1, lo, lhi)); stack.append((int(nd["a"]), rlo, hi))
        assert seen.all()


def test_builder_rejects_what_the_reference_rejects():
    from lumo_b200 import Scene, Material, Rectangle, Sphere, CameraBuilder, Spectrum, program
    cam = CameraBuilder.new().resolution((8, 8)).build()
    s = Scene()
    s.add(Rectangle((0, 0, 0), (1, 0, 0), (1, 1, 0), Material.lambertian(Spectrum.from_rgb(0.5, 0.5, 0.5))))
    with pytest.raises(RuntimeError, match="no lights"):            # renderer.rs:42 assert
        native.build_blob(s._program(cam))
    with pytest.raises(AssertionError):
        Sphere(0.0, Material.Blank)                                  # sphere.rs:17
    with pytest.raises(RuntimeError, match="bad magic"):
        native.build_blob(b"not a program....")
    with pytest.raises(AssertionError):
        CameraBuilder.new().vfov(180.0).build()                      # matrices.rs:5


@pytest.mark.parametrize("name", list(SMALL))
def test_light_bounding_spheres_contain_the_lights(name):
    """LumoLight.bound_*: the sphere k_nee_b uses to skip a light's intersection test must contain every point of the light in
    world space with room to spare — vertices of triangle / rectangle / mesh lights, the whole ball of a sphere light — through
    the light's instance transform."""
    prog, blob, _ = small_scene(name)
    B = native.Blob(blob)
    P = B.params
    n_obj, n_l = int(P["n_objects"]), int(P["n_lights"])
    assert len(B.lights) == n_l
    rs = np.random.RandomState(3)
    u = rs.randn(256, 3); u /= np.linalg.norm(u, axis=1, keepdims=True)
    for li in range(n_l):
        o = B.objects[n_obj + li]; L = B.lights[li]
        kind, geom = int(o["kind"]), int(o["geom"])
        if kind in (0, 1):                              # kd mesh; a rectangle is a two-triangle mesh
            T = B.kd_trees[geom]; tv = B.tri_verts[int(T["tri_base"]):int(T["tri_base"]) + int(T["n_tris"])]
            pts = np.concatenate([tv["a"], tv["b"], tv["c"]])
        elif kind == 2: pts = u * float(B.spheres[geom]["radius"])
        else: tv = B.tri_verts[geom:geom + 1]; pts = np.concatenate([tv["a"], tv["b"], tv["c"]])
        if int(o["inst"]) >= 0:
            m = B.instances[int(o["inst"])]["m"].reshape(3, 4)
            pts = pts @ m[:, :3].T + m[:, 3]
        c, r = np.asarray(L["bound_c"]), float(L["bound_r"])
        d = np.linalg.norm(pts - c, axis=1).max()
        assert np.isfinite(r) and r > 0.0
        assert d <= r * (1.0 - 5e-5), (name, li, kind, d, r)          # padded by 1e-4 of the radius
        assert r <= 1.8 * d + 1e-5 * (np.abs(c).max() + 1.0), (name, li, kind, d, r)   # and not uselessly large (half diagonal of the box <= sqrt(3) x the farthest point)


@pytest.mark.parametrize("name", ["bunny", "bistro", "conference"])
def test_blob_does_not_depend_on_the_number_of_host_threads(name, monkeypatch):
    """kd-trees are built in parallel (one tree per task) and the world-space BVH in parallel (big nodes: their passes over the
    primitives; then independent subtrees): the blob must be byte for byte what one thread builds."""
    from lumo_b200 import scenes
    kw = dict(SMALL[name])
    if name == "bunny": kw["n_tris"] = 90000          # more than 2^16 primitives: the top phase of the BVH build splits its passes over the threads
    s, cam, _ = scenes.CONFIGS[name](**kw)
    prog = s._program(cam)
    blobs = []
    for threads in ("1", "3", "8"):
        monkeypatch.setenv("LUMO_HOST_THREADS", threads)
        blobs.append(bytes(native.build_blob(prog)))
    assert blobs[0] == blobs[1] == blobs[2]
