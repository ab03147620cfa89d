"""The N>1 host path on CPU: world_size-2 (and 3) gloo processes run lumo_b200.distributed's sharding and
film reduce with a stand-in film producer (the render itself needs a GPU).  The producer adds, for every
global sample index in the rank's range, a value that depends only on (pixel, sample) — like the device's
Philox-keyed paths — so the reduced film must equal the single-process film exactly."""
import os
import socket
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from lumo_b200 import distributed as D


def test_sample_ranges_partition():
    for total in (0, 1, 7, 64, 1024, 1000):
        for world in (1, 2, 3, 4, 8):
            r = [D.sample_range(k, world, total) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lumo_b200", "liblumo_gpu.so")),
                    reason="liblumo_gpu.so not built (needs nvcc)")
def test_c_abi_sample_range_equals_the_python_rule():
    """lumo_gpu_render_multi splits the sample range with lumo_gpu_sample_range; it must be the rule the one-process-per-GPU
    path uses (distributed.sample_range), also with an offset range, and must reject bad arguments."""
    from lumo_b200 import native
    for total in (0, 1, 7, 64, 1000):
        for world in (1, 2, 3, 8, 16):
            for off in (0, 5):
                got = [native.sample_range(g, world, off, off + total) for g in range(world)]
                want = [tuple(off + v for v in D.sample_range(g, world, total)) for g in range(world)]
                assert got == want
    for bad in ((2, 2, 0, 8), (-1, 2, 0, 8), (0, 0, 0, 8), (0, 2, 9, 8)):
        with pytest.raises(RuntimeError, match="bad arguments"):
            native.sample_range(*bad)


def _fake_render(px, sp, begin, end, W, H):
    pix = torch.arange(W * H, dtype=torch.float64).view(H, W)
    for s in range(begin, end):
        v = torch.sin(pix * 0.37 + s * 1.3) ** 2
        px[..., 0] += v; px[..., 1] += 0.5 * v; px[..., 2] += 0.25 * v; px[..., 3] += 1.0
        sp[..., 0] += 0.125 * v


def _worker(rank, world, port, total, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    W, H = 24, 16
    buf, px, sp = D.film_buffer(W, H, "cpu")
    b, e = D.sample_range(rank, world, total)
    _fake_render(px, sp, b, e, W, H)
    D.reduce_film(buf, 0)
    cnt = D.reduce_counters({"camera_paths": (e - b) * W * H, "closest": 3 * (e - b), "occlusion": e - b, "cost": 2 * (e - b)}, "cpu")
    if rank == 0:
        np.save(out, buf.numpy())
        assert cnt["camera_paths"] == total * W * H and cnt["closest"] == 3 * total
    dist.barrier(); dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_film_equals_single_process(world, tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    total = 7
    out = str(tmp_path / "film.npy")
    mp.spawn(_worker, args=(world, port, total, out), nprocs=world, join=True)
    W, H = 24, 16
    buf, px, sp = D.film_buffer(W, H, "cpu")
    _fake_render(px, sp, 0, total, W, H)
    got = np.load(out)
    assert np.allclose(got, buf.numpy(), rtol=1e-13, atol=0)
    assert np.array_equal(got[3:4 * W * H:4], np.full(W * H, float(total)))      # weight channel of pixels[H,W,4]: one per sample
