"""The BASELINE configs as scene builders (examples/{cornell,bunny,dragon,caustics,conference,bistro}.rs).

Cornell and the empty box are the reference's procedural scenes.  The .obj assets of the other
examples are downloaded by the reference at run time (src/parser.rs:149-165) and do not exist
offline, so their meshes are seeded synthetic stand-ins of matching triangle counts (meshes.py,
SURVEY §8d); cameras, materials, transforms and integrators are the examples'.  `scale` < 1 shrinks
triangle counts for tests."""
import math
import numpy as np
from . import meshes
from .api import (Scene, Camera, Material, Texture, Rectangle, Sphere, TriangleMesh, Mesh, Instance, LooseTriangles, Integrator, ToneMap)
from .spectrum import Spectrum


def _tri_mesh(verts, faces, material):
    return TriangleMesh.new(np.asarray(verts, np.float64), np.asarray(faces, np.int64), [], [], material)


def cornell(resolution=(512, 512)):
    """examples/cornell.rs"""
    cam = Camera.cornell_box()
    cam._resolution = tuple(resolution)
    return Scene.cornell_box(), cam, Integrator.PathTrace


def _box(resolution):
    from .api import CameraBuilder
    return CameraBuilder.new().resolution(resolution).build()


def bunny(n_tris=69632, resolution=(1024, 768)):
    """examples/bunny.rs:7-33 with a displaced-sphere stand-in for bunny.obj (seed 7)."""
    def_color = Spectrum.from_srgb(242, 242, 242)
    s = Scene.empty_box(def_color, Material.diffuse(Spectrum.RED()), Material.diffuse(Spectrum.GREEN()))
    v, f = meshes.displaced_sphere(n_tris, seed=7)
    m = _tri_mesh(v, f, Material.metal(Spectrum.YELLOW(), 0.1, 2.5, 3.0))
    s.add(m.to_unit_size().to_origin().set_y(-0.799).translate(0.0, 0.0, -1.5))
    return s, _box(resolution), Integrator.PathTrace


def dragon(n_tris=870400, resolution=(1024, 768)):
    """examples/dragon.rs:7-34 with a displaced torus-knot stand-in for dragon.obj (seed 11)."""
    def_color = Spectrum.from_srgb(242, 242, 242)
    s = Scene.empty_box(def_color, Material.diffuse(Spectrum.RED()), Material.diffuse(Spectrum.GREEN()))
    v, f = meshes.torus_knot(n_tris, seed=11)
    m = _tri_mesh(v, f, Material.transparent(Spectrum.MAGENTA(), 0.03, 1.5))
    s.add(m.to_unit_size().to_origin().rotate_y(5.0 * math.pi / 8.0).scale_uniform(1.3).set_y(-0.799).translate(0.0, 0.0, -1.4))
    return s, _box(resolution), Integrator.PathTrace


def caustics(n_tris=15488, resolution=(1024, 768)):
    """examples/caustics.rs:7-42: one mesh instanced twice (mirror + dispersive glass), BDPT."""
    from .api import CameraBuilder
    cam = CameraBuilder.new().origin(0.0, 0.0, 2.0).zoom(3.0).resolution(resolution).build()
    def_color = Spectrum.from_srgb(242, 242, 242)
    s = Scene.empty_box(def_color, Material.diffuse(Spectrum.MAGENTA()), Material.diffuse(Spectrum.CYAN()))
    v, f = meshes.displaced_sphere(n_tris, seed=13, amplitude=0.35)
    head = _tri_mesh(v, f, Material.Blank).to_unit_size()
    s.add(head.clone(Material.mirror()).to_origin().rotate_y(-math.pi / 8).rotate_z(math.pi / 8).rotate_x(-math.pi / 8).translate(0.5, -0.3, -1.0))
    s.add(head.clone(Material.glass()).to_origin().rotate_y(math.pi / 8).rotate_z(-math.pi / 8).rotate_x(math.pi / 16).translate(-0.35, 0.25, -1.25))
    return s, cam, Integrator.BDPathTrace


def _chunk_dims(tris_per_chunk):
    """(nu, nv) of a closed blob and of an open patch with exactly tris_per_chunk triangles."""
    n = int(round(math.sqrt(tris_per_chunk / 4.0)))       # blob: 2*nu*(nv-1), nu = 2n, nv-1 = n
    assert 2 * (2 * n) * n == tris_per_chunk, "tris_per_chunk must be 4*n^2"
    return (2 * n, n + 1), (n + 1, 2 * n + 1)              # patch: 2*(nu-1)*(nv-1) = 2*n*2n


def conference(n_chunks=64, tris_per_chunk=5184, resolution=(1024, 768)):
    """examples/conference.rs:7-30 with a synthetic room (seed 17): n_chunks kd-trees (floor patches
    and furniture blobs on a grid), diffuse materials, the example's two sphere lights and camera."""
    from .api import CameraBuilder
    cam = CameraBuilder.new().origin(-50.0, 400.0, -350.0).towards(500.0, 0.0, 250.0).resolution(resolution).build()
    rng = np.random.RandomState(17)
    s = Scene()
    (bu, bv), (pu, pv) = _chunk_dims(tris_per_chunk)
    side = int(math.ceil(math.sqrt(n_chunks / 2.0)))
    cell = 1400.0 / side
    palette = [Material.diffuse(Spectrum.from_srgb(*c)) for c in ((200, 190, 170), (120, 80, 60), (60, 90, 140), (180, 60, 50), (90, 140, 90), (230, 230, 230))]
    made = 0
    for k in range(side * side):
        gx, gz = k % side, k // side
        x0, z0 = -300.0 + gx * cell, -700.0 + gz * cell
        if made < n_chunks:                                # floor patch
            v, f = meshes.height_patch(pu, pv, seed=1000 + k, amplitude=0.01)
            v = v * np.array([cell, cell, cell]) + np.array([x0, 0.0, z0])
            s.add(_tri_mesh(v, f, palette[k % 2])); made += 1
        if made < n_chunks:                                # furniture
            v, f = meshes.displaced_blob(bu, bv, seed=2000 + k, amplitude=0.25, squash=(0.35 * cell, 40.0 + 60.0 * rng.rand(), 0.35 * cell))
            v = v + np.array([x0 + 0.5 * cell, 60.0 + 40.0 * rng.rand(), z0 + 0.5 * cell])
            s.add(_tri_mesh(v, f, palette[2 + k % 4])); made += 1
    white = Material.light(Spectrum.WHITE())
    s.add_light(Sphere.new(10.0, white).translate(-200.0, 40.0, -400.0))
    s.add_light(Sphere.new(10.0, white).translate(900.0, 300.0, -600.0))
    return s, cam, Integrator.PathTrace


def bistro(n_chunks=1024, tris_per_chunk=1024, n_emissive=4096, resolution=(1920, 1080)):
    """examples/bistro.rs:15-53 (night exterior) with a synthetic street (seed 19): n_chunks kd-trees
    (ground patches + facade/prop blobs), n_emissive loose emissive triangles (each one light, like
    parser/obj.rs:92-104), a solid-colour environment sphere in place of the HDR, the example's camera."""
    from .api import CameraBuilder
    cam = CameraBuilder.new().origin(-16.0, 5.0, -1.0).towards(0.0, 0.0, 0.0).resolution(resolution).build()
    rng = np.random.RandomState(19)
    s = Scene()
    (bu, bv), (pu, pv) = _chunk_dims(tris_per_chunk)
    side = int(math.ceil(math.sqrt(n_chunks / 2.0)))
    cell = 60.0 / side
    mats = [Material.diffuse(Spectrum.from_srgb(*c)) for c in ((150, 140, 130), (90, 85, 80), (170, 120, 90), (70, 90, 110), (190, 180, 160))]
    mats.append(Material.metal(Spectrum.from_srgb(220, 200, 160), 0.2, 2.5, 3.0))
    mats.append(Material.transparent(Spectrum.WHITE(), 0.0, 1.5))
    made = 0
    for k in range(side * side):
        gx, gz = k % side, k // side
        x0, z0 = -30.0 + gx * cell, -30.0 + gz * cell
        if made < n_chunks:
            v, f = meshes.height_patch(pu, pv, seed=3000 + k, amplitude=0.02)
            v = v * np.array([cell, cell, cell]) + np.array([x0, 0.0, z0])
            s.add(_tri_mesh(v, f, mats[k % 2])); made += 1
        if made < n_chunks:
            street = abs((z0 + 0.5 * cell) - 0.0) < 4.0      # keep a street free along the view direction
            h = (0.4 + 0.8 * rng.rand()) if street else (2.0 + 6.0 * rng.rand())
            v, f = meshes.displaced_blob(bu, bv, seed=4000 + k, amplitude=0.15, squash=(0.4 * cell, h, 0.4 * cell))
            v = v + np.array([x0 + 0.5 * cell, h, z0 + 0.5 * cell])
            s.add(_tri_mesh(v, f, mats[2 + k % 5])); made += 1
    if n_emissive:
        c = np.stack([-28.0 + 56.0 * rng.rand(n_emissive), 2.5 + 2.0 * rng.rand(n_emissive), -28.0 + 56.0 * rng.rand(n_emissive)], -1)
        a = rng.rand(n_emissive) * 2 * np.pi
        e1 = np.stack([np.cos(a), np.zeros_like(a), np.sin(a)], -1) * 0.08
        e2 = np.stack([-np.sin(a), np.zeros_like(a), np.cos(a)], -1) * 0.08
        verts = np.concatenate([c - e1 - e2, c + e1 - e2, c + e2], 0)
        idx = np.arange(n_emissive)
        faces = np.stack([idx, idx + n_emissive, idx + 2 * n_emissive], -1)    # wound to face down (normal -y)
        lamp = Material.light_scale(Spectrum.from_srgb(255, 214, 170), 40.0)
        s.add_light(LooseTriangles(_tri_mesh(verts, faces, lamp), lamp))
    s.set_environment_map(Spectrum.from_srgb(20, 24, 40), 0.001 * 1000.0)
    return s, cam, Integrator.PathTrace


def textured(n_tris=2048, resolution=(256, 192)):
    """Not a reference example: the empty box with every `Texture` kind (texture.rs:23-38) and a bump map
    (material.rs:14) on its surfaces, a uv-mapped mesh, an image-textured light and an image environment map.
    The parity tests and `bench.py --workload textured` use it to exercise the texture / bump path."""
    from .image import Image
    rs = np.random.RandomState(5)
    palette = np.array([[230, 60, 50], [40, 170, 90], [60, 90, 220], [240, 220, 80], [250, 250, 250], [30, 30, 30]], np.uint8)
    img = Image.from_rgb8(palette[rs.randint(0, len(palette), size=(12, 16))])
    warm = Image.from_rgb8(np.array([[252, 201, 138], [255, 240, 220]], np.uint8)[rs.randint(0, 2, size=(4, 4))])
    sky = Image.from_rgb8(np.array([[20, 24, 40], [60, 80, 140]], np.uint8)[(np.arange(8)[:, None] + np.arange(16)[None, :]) % 2])
    yy, xx = np.mgrid[0:16, 0:16]
    bump_rgb = np.stack([128 + 60 * np.sin(xx * 0.9), 128 + 60 * np.cos(yy * 0.7), np.full(xx.shape, 230.0)], -1).astype(np.uint8)
    bump = Image.bump_from_rgb8(bump_rgb)
    white = Spectrum.from_srgb(242, 242, 242)
    checker = Texture.Checkerboard(Texture.Solid(white), Texture.Marble(7, Spectrum.RED()), 6.0)
    nested = Texture.Checkerboard(Texture.Image(img), Texture.Checkerboard(Spectrum.CYAN(), Texture.Mandelbrot(), 3.0), 2.0)
    LIGHT_EPS = 0.001
    ground = -0.8; ceiling = -ground; right = 1.0; left = -right; front = -2.0; back = 0.0; l_dim = 0.25
    s = Scene()
    s.add_light(Rectangle((-l_dim, ceiling - LIGHT_EPS, 0.6 * front + l_dim), (-l_dim, ceiling - LIGHT_EPS, 0.6 * front - l_dim),
                          (l_dim, ceiling - LIGHT_EPS, 0.6 * front - l_dim), Material.light_scale(Texture.Image(warm), 1.0)))
    s.add(Rectangle((left, ground, back), (left, ground, front), (left, ceiling, front), Material.diffuse(Texture.Image(img))))
    s.add(Rectangle((right, ground, front), (right, ground, back), (right, ceiling, back), Material.diffuse(Texture.Mandelbrot())))
    s.add(Rectangle((left, ground, back), (right, ground, back), (right, ground, front), Material.diffuse(checker)))
    s.add(Rectangle((left, ceiling, front), (right, ceiling, front), (right, ceiling, back), Material.diffuse(white)))
    s.add(Rectangle((left, ground, front), (right, ground, front), (right, ceiling, front),
                    Material.microfacet(0.6, 1.5, 0.0, False, False, nested, Spectrum.WHITE(), Spectrum.BLACK(), bump_map=bump)))
    s.add(Sphere(0.3, Material.metal(Texture.Marble(11, Spectrum.YELLOW()), 0.15, 2.5, 3.0)).translate(-0.45, ground + 0.3, -1.3))
    s.add(Sphere(0.25, Material.microfacet(0.05, 1.5, 0.0, True, True, Spectrum.BLACK(), Spectrum.WHITE(), Texture.Image(img), bump_map=bump)).translate(0.5, ground + 0.25, -1.0))
    v, f = meshes.displaced_sphere(n_tris, seed=3, amplitude=0.2)
    uv = np.stack([0.5 + np.arctan2(v[:, 2], v[:, 0]) / (2 * np.pi), 0.5 + np.arcsin(np.clip(v[:, 1] / np.linalg.norm(v, axis=1), -1, 1)) / np.pi], -1)
    m = TriangleMesh.new(np.asarray(v, np.float64), np.asarray(f, np.int64), [], uv, Material.microfacet(0.4, 1.5, 0.0, False, False, checker, Texture.Image(img), Spectrum.BLACK(), bump_map=bump))
    m.face_uvs = np.asarray(f, np.int64)
    s.add(m.to_unit_size().to_origin().scale_uniform(0.5).translate(0.1, 0.1, -1.6))
    s.set_environment_map(Texture.Image(sky), 0.5)
    return s, _box(resolution), Integrator.PathTrace


CONFIGS = {"cornell": cornell, "bunny": bunny, "dragon": dragon, "caustics": caustics, "conference": conference, "bistro": bistro, "textured": textured}
