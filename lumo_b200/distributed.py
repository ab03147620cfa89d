"""Multi-GPU host logic: the scene is replicated, samples are sharded, film accumulators are
combined with one reduce.  One process per GPU over torch.distributed (NCCL on the GPU box; the same code
runs under gloo on CPU in tests/test_distributed.py).

The reference's unit of scheduling is (16x16 tile) x (256-sample batch) pulled from one queue by worker
threads (src/renderer.rs:183-204).  Across GPUs the unit is a contiguous range of global sample indices:
every path's random stream is keyed by (pixel, global sample index), so the union of the ranks' ranges is
exactly the single-GPU render, up to floating-point summation order in the reduce."""
import torch
import torch.distributed as dist


def sample_range(rank, world, total_spp):
    """[begin, end) of global sample indices for `rank`: contiguous, disjoint, covering [0, total_spp);
    sizes differ by at most one."""
    assert 0 <= rank < world and total_spp >= 0
    base, extra = divmod(total_spp, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def film_buffer(width, height, device):
    """One contiguous f64 buffer holding pixels[H,W,4] (sum r*w, g*w, b*w, w) then splats[H,W,3]
    (src/tracer/film.rs:52-76), so that the whole film is ONE collective."""
    buf = torch.zeros(width * height * 7, dtype=torch.float64, device=device)
    return buf, buf[: width * height * 4].view(height, width, 4), buf[width * height * 4:].view(height, width, 3)


def reduce_film(buf, dst=0):
    """Film::add_tile across ranks (film.rs:155-171): sum of accumulators onto `dst`."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(buf, dst=dst, op=dist.ReduceOp.SUM)
    return buf


def reduce_counters(counters, device):
    """Sums the per-rank ray counters on every rank (camera paths, closest-hit, occlusion, cost)."""
    keys = sorted(counters)
    t = torch.tensor([float(counters[k]) for k in keys], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return {k: int(v) for k, v in zip(keys, t.tolist())}


def render_sharded(scene, total_spp, device, **kw):
    """Each rank renders its sample range of a `total_spp` render into a device film buffer
    (lumo_gpu_render_dev) and the buffers are reduced onto rank 0.  Returns (buf, pixels, splats, counters)."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    b, e = sample_range(rank, world, total_spp)
    buf, px, sp = film_buffer(scene.res_x, scene.res_y, device)
    cnt, ms = scene.render_dev(px.data_ptr(), sp.data_ptr(), spp_begin=b, spp_end=e, total_spp=total_spp, **kw)
    reduce_film(buf, 0)
    return buf, px, sp, reduce_counters({k: cnt[k] for k in ("camera_paths", "closest", "occlusion", "cost")}, device)
