"""G1-G3: the CUDA traversal through the C ABI against the oracle on identical ray
batches.  Integer outputs and distances are compared bit for bit: the device code keeps the
reference's operation order and is compiled without FMA contraction."""
import numpy as np
import pytest
import oracle_lib
from conftest import small_scene, ray_batches, SMALL

pytestmark = pytest.mark.gpu
N = 20000


def _bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


@pytest.mark.parametrize("name", list(SMALL))
def test_closest_hit_ids_and_distances_bit_exact(name, gpu_ctx):
    from lumo_b200 import native
    prog, blob, _ = small_scene(name)
    O = oracle_lib.OracleScene(prog)
    G = native.GpuScene(gpu_ctx, blob)
    for kind, (o, d) in ray_batches(O, N, seed=23).items():
        eo, et, ett, eb = O.trace_closest(o, d)
        go, gt, gtt, gb = G.trace_closest(o, d)
        assert np.array_equal(eo, go), (name, kind, "object ids", int((eo != go).sum()))
        assert np.array_equal(et, gt), (name, kind, "triangle ids", int((et != gt).sum()))
        assert np.array_equal(_bits(ett), _bits(gtt)), (name, kind, "t (0 ulp)")
        assert np.array_equal(_bits(eb), _bits(gb)), (name, kind, "barycentrics (0 ulp)")
        assert (eo != 0xFFFFFFFF).sum() > 0.2 * len(eo) or kind == "incoherent"
    G.close(); O.close()


@pytest.mark.parametrize("name", list(SMALL))
def test_any_hit_and_first_found(name, gpu_ctx):
    from lumo_b200 import native
    prog, blob, _ = small_scene(name)
    O = oracle_lib.OracleScene(prog)
    G = native.GpuScene(gpu_ctx, blob)
    rs = np.random.RandomState(5)
    for kind, (o, d) in ray_batches(O, N, seed=29).items():
        _, _, t, _ = O.trace_closest(o, d)
        # t_max on both sides of the first hit, like hit_light's t_light - 1e-10
        tm = np.where(np.isfinite(t), t * rs.choice([0.5, 1.0 - 1e-12, 1.5, 3.0], size=len(t)), 10.0) - 1e-10
        assert np.array_equal(O.trace_any(o, d, tm), G.trace_any(o, d, tm)), (name, kind, "occlusion booleans")
        assert np.array_equal(_bits(O.trace_first_found(o, d)), _bits(G.trace_first_found(o, d))), (name, kind, "first-found t")
    G.close(); O.close()


def test_visit_counters_equal_oracle(gpu_ctx):
    """The N_* of the byte formula: the faithful kernel visits exactly the oracle's nodes."""
    from lumo_b200 import native
    prog, blob, _ = small_scene("bunny")
    O = oracle_lib.OracleScene(prog)
    G = native.GpuScene(gpu_ctx, blob)
    o, d = ray_batches(O, 5000, seed=31)["secondary"]
    O.counters(reset=True)
    O.trace_closest(o, d, threads=1)
    oc = O.counters(reset=True)
    gpu_ctx.count_visits(True)
    G.trace_closest(o, d)
    gc, _ = gpu_ctx.visits()
    gpu_ctx.count_visits(False)
    for k in ("tlas_nodes", "inst", "kd_nodes", "leaf_idx", "tri_tests", "sphere_tests"):
        assert oc[k] == gc[k], (k, oc[k], gc[k])


def test_edge_cases(gpu_ctx):
    from lumo_b200 import native
    prog, blob, _ = small_scene("cornell")
    G = native.GpuScene(gpu_ctx, blob)
    O = oracle_lib.OracleScene(prog)
    e = np.zeros((0, 3))
    obj, tri, t, bary = G.trace_closest(e, e)                       # empty batch
    assert len(obj) == 0
    # axis-parallel directions (zero components -> inf reciprocals), rays starting on geometry, rays leaving the box
    o = np.array([[278.0, 273.0, -800.0], [278.0, 273.0, 100.0], [278.0, 0.0, 100.0], [278.0, 273.0, 100.0], [0.0, 0.0, 0.0], [278.0, 600.0, 279.6]])
    d = np.array([[0.0, 0.0, 1.0], [0.0, -1.0, 0.0], [0.0, 1.0, 0.0], [1.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0]])
    eo, et, ett, eb = O.trace_closest(o, d)
    go, gt, gtt, gb = G.trace_closest(o, d)
    assert np.array_equal(eo, go) and np.array_equal(et, gt) and np.array_equal(_bits(ett), _bits(gtt))
    assert go[-1] == 0xFFFFFFFF and np.isinf(gtt[-1])
    with pytest.raises(RuntimeError):
        native.GpuScene(gpu_ctx, blob[:-16])                        # malformed blob -> status code, not a crash


def test_flat_traversal_is_bit_exact_too():
    """The lane-refilled state-machine traversal (trace_flat.cuh, opt-in via LUMO_TRACE_FLAT=1) must produce the
    same bits as the default nested one.  A context reads the switch when it is created."""
    import os
    from lumo_b200 import native
    prog, blob, _ = small_scene("bistro")
    O = oracle_lib.OracleScene(prog)
    os.environ["LUMO_TRACE_FLAT"] = "1"
    try:
        ctx = native.GpuContext(0)
    finally:
        del os.environ["LUMO_TRACE_FLAT"]
    G = native.GpuScene(ctx, blob)
    for kind, (o, d) in ray_batches(O, 8000, seed=37).items():
        eo, et, ett, eb = O.trace_closest(o, d)
        go, gt, gtt, gb = G.trace_closest(o, d)
        assert np.array_equal(eo, go) and np.array_equal(et, gt) and np.array_equal(_bits(ett), _bits(gtt)) and np.array_equal(_bits(eb), _bits(gb)), kind
        tm = np.where(np.isfinite(ett), ett * 0.999, 5.0)
        assert np.array_equal(O.trace_any(o, d, tm), G.trace_any(o, d, tm)), kind
        assert np.array_equal(_bits(O.trace_first_found(o, d)), _bits(G.trace_first_found(o, d))), kind
    G.close(); ctx.close(); O.close()
