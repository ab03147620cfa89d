"""Host-side mirror of the reference's public construction API (src/lib.rs:13-22, src/tracer.rs:1-14):
Scene / Camera / CameraBuilder / Material / Texture / Spectrum / Rectangle / Sphere / TriangleMesh /
Instance transforms / Renderer / Integrator / ToneMap / SamplerType / PixelFilter / Film.

Same names, argument meaning and error behaviour (asserts where the reference asserts).  The
objects only *record* the construction; `Renderer.render()` serialises the record as a scene
program (program.py), has the native host library build lumo's kd-trees / BVHs and flatten them
into the device blob, uploads it and runs the CUDA wavefront pipeline through the C ABI
(include/lumo_gpu.h).  There is no CPU rendering path in this package.
"""
import math
import numpy as np
from . import program as P
from .spectrum import Spectrum
from ._tables import ILLUMINANTS
from . import color as _color


class illuminants:                                   # src/tracer/color.rs:23-37
    A, D50, D65, F2, F7, CORNELL = "A", "D50", "D65", "F2", "F7", "CORNELL"


class Integrator:                                    # src/tracer/integrator.rs:17-27
    PathTrace, DirectLight, BDPathTrace = 0, 1, 2
    NAMES = {0: "path tracing", 1: "direct light integration", 2: "bidirectional path tracing"}


class SamplerType:                                   # src/samplers.rs:8-21
    Uniform, Jittered, MultiJittered, Sobol = 0, 1, 2, 3


class ToneMap:                                       # src/tone_mapping.rs:13-35
    def __init__(self, kind=0, arg=0.0):
        self.kind, self.arg = kind, arg
    NoMap = None
    Reinhard = None

    @staticmethod
    def Clamp(mx):
        return ToneMap(1, float(mx))


ToneMap.NoMap = ToneMap(0)
ToneMap.Reinhard = ToneMap(2)


class PixelFilter:                                   # src/tracer/filter.rs:8-67
    def __init__(self, kind, r, p=0.0):
        assert r > 0.0
        self.kind, self.r, self.p = kind, float(r), float(p)

    @staticmethod
    def square(r): return PixelFilter(0, r)
    @staticmethod
    def triangle(r): return PixelFilter(1, r)

    @staticmethod
    def gaussian(r, sigma):
        assert sigma > 0.0
        return PixelFilter(2, r, sigma)

    @staticmethod
    def mitchell(r, b): return PixelFilter(3, r, b)
    @staticmethod
    def default(): return PixelFilter.gaussian(1.5, 1.5 / 4.0)

    def r_disc(self): return int(math.ceil(self.r - 0.5))

    def integral(self):                              # filter.rs:105-116
        r, p = self.r, self.p
        if self.kind == 0: return 2.0 * r * 2.0 * r
        if self.kind == 1: return r * r * r * r
        if self.kind == 3: return r * r * 0.25
        g = lambda x: math.exp(-x * x / (2 * p * p)) / math.sqrt(2 * math.pi * p * p)
        denom = p * math.sqrt(2.0)
        ig = 0.5 * (math.erf(r / denom) - math.erf(-r / denom))
        return (ig - 2.0 * r * g(r)) ** 2


class ColorSpace:                                    # src/tracer/color/space.rs:80-115
    sRGB, DCI_P3, Rec_2020 = 0, 1, 2


class Texture:                                       # src/tracer/texture.rs:23-38
    """Solid(Spectrum) | Checkerboard(Texture, Texture, scale) | Marble(Perlin seed, Spectrum) | Image(Image<Spectrum>) | Mandelbrot."""
    def __init__(self, kind, spec=None, a=None, b=None, scale=1.0, seed=0, image=None):
        self.kind, self.spec, self.a, self.b, self.scale, self.seed, self.image = kind, spec, a, b, scale, seed, image

    @staticmethod
    def from_spectrum(spec): return Texture(P.TEX_SOLID, spec=spec)
    @staticmethod
    def Solid(spec): return Texture(P.TEX_SOLID, spec=spec)
    @staticmethod
    def Checkerboard(t1, t2, scale): return Texture(P.TEX_CHECKER, a=_as_tex(t1), b=_as_tex(t2), scale=float(scale))
    @staticmethod
    def Marble(seed, spec):
        """`Texture::Marble(Perlin::new(seed), spec)`; the reference's `Perlin::default()` draws its seed from the clock."""
        return Texture(P.TEX_MARBLE, spec=spec, seed=int(seed))
    @staticmethod
    def Image(image):
        assert image.kind == "spectrum"
        return Texture(P.TEX_IMAGE, spec=image.mean, image=image)
    @staticmethod
    def Mandelbrot(): return Texture(P.TEX_MANDELBROT)

    def _emit(self, w, cache):
        """texture id in the program (children first)"""
        if id(self) in cache: return cache[id(self)]
        sp = (self.spec if self.spec is not None else Spectrum.BLACK()).as_tuple()
        if self.kind == P.TEX_CHECKER:
            tid = w.texture(P.TEX_CHECKER, a=self.a._emit(w, cache), b=self.b._emit(w, cache), scale=self.scale)
        elif self.kind == P.TEX_IMAGE:
            tid = w.texture(P.TEX_IMAGE, spec=sp, width=self.image.width, height=self.image.height, data=self.image.data)
        else:
            tid = w.texture(self.kind, spec=sp, seed=self.seed)
        cache[id(self)] = tid
        return tid


def _as_tex(t):
    if isinstance(t, Texture): return t
    if isinstance(t, Spectrum): return Texture.Solid(t)
    if callable(t): return Texture.Solid(t())
    raise TypeError("expected Texture or Spectrum")


def _tex(t):
    """the Solid spectrum of a texture argument (Spectrum::BLACK placeholder for the other kinds: the device reads the texture)"""
    t = _as_tex(t)
    return t.spec if (t.kind == P.TEX_SOLID and t.spec is not None) else Spectrum.BLACK()


def _tex_ref(t):
    t = _as_tex(t)
    return None if t.kind == P.TEX_SOLID else t


class Material:                                      # src/tracer/material.rs:12-194
    def __init__(self, kind, **kw):
        self.kind = kind
        self.kw = kw

    @staticmethod
    def microfacet(roughness, eta, k, is_transparent, fresnel_enabled, kd, ks, tf, bump_map=None):
        assert 0.0 <= roughness <= 1.0                # microfacet.rs:27
        assert bump_map is None or bump_map.kind == "normal"
        eta_kind = P.ETA_CONST
        if is_transparent and eta == 1.5: eta_kind = P.ETA_GLASS        # material.rs:37-45
        if is_transparent and eta == 2.5: eta_kind = P.ETA_DIAMOND
        kind = P.M_MFDIELECTRIC if is_transparent else (P.M_MFCONDUCTOR if fresnel_enabled else P.M_MFDIFFUSE)
        return Material(kind, roughness=roughness, eta_kind=eta_kind, eta=eta, k_kind=P.ETA_CONST, k=k,
                        kd=_tex(kd).as_tuple(), ks=_tex(ks).as_tuple(), tf=_tex(tf).as_tuple(),
                        _textures=dict(kd_tex=_tex_ref(kd), ks_tex=_tex_ref(ks), tf_tex=_tex_ref(tf)), _bump=bump_map)

    @staticmethod
    def metal(ks, roughness, eta, k):
        return Material.microfacet(roughness, eta, k, False, True, Spectrum.WHITE(), ks, Spectrum.BLACK())

    @staticmethod
    def diffuse(kd):
        return Material.microfacet(1.0, 1.5, 0.0, False, False, kd, Spectrum.WHITE(), Spectrum.BLACK())

    @staticmethod
    def lambertian(spec):
        return Material(P.M_LAMBERTIAN, kd=_tex(spec).as_tuple())

    @staticmethod
    def transparent(tf, roughness, eta):
        return Material.microfacet(roughness, eta, 0.0, True, True, Spectrum.BLACK(), Spectrum.WHITE(), tf)

    @staticmethod
    def mirror():                                    # material.rs:124-143
        return Material(P.M_MFCONDUCTOR, roughness=0.0, eta_kind=P.ETA_MIRROR, k_kind=P.K_MIRROR,
                        kd=Spectrum.BLACK().as_tuple(), ks=Spectrum.WHITE().as_tuple(), tf=Spectrum.BLACK().as_tuple())

    @staticmethod
    def glass():                                     # material.rs:145-164
        return Material(P.M_MFDIELECTRIC, roughness=0.0, eta_kind=P.ETA_GLASS, k_kind=P.ETA_CONST, k=0.0,
                        kd=Spectrum.BLACK().as_tuple(), ks=Spectrum.WHITE().as_tuple(), tf=Spectrum.WHITE().as_tuple())

    @staticmethod
    def light(ke): return Material.light_scale(ke, 1.0)

    @staticmethod
    def light_scale(ke, scale):
        return Material.Light(ke, illuminants.D65, scale, False)

    @staticmethod
    def Light(ke, illuminant, scale, two_sided):
        return Material(P.M_LIGHT, ke=_tex(ke).as_tuple(), illuminant=ILLUMINANTS.index(illuminant), scale=scale, two_sided=int(two_sided),
                        _textures=dict(ke_tex=_tex_ref(ke)))

    Blank = None

    def is_light(self): return self.kind == P.M_LIGHT


Material.Blank = Material(P.M_BLANK)


# ---- objects -----------------------------------------------------------------------------------
class _Instanceable:
    """src/tracer/object/instance.rs:201-299: every transform wraps the object in an Instance and
    composes `T * current`."""
    def _inst(self):
        return self if isinstance(self, Instance) else Instance(self)

    def translate(self, x, y, z): return self._inst()._op(P.OP_TRANSLATE, x, y, z)

    def scale(self, x, y, z):
        assert x * y * z != 0.0
        return self._inst()._op(P.OP_SCALE, x, y, z)

    def scale_uniform(self, s): return self.scale(s, s, s)
    def rotate_x(self, r): return self._inst()._op(P.OP_ROTX, r)
    def rotate_y(self, r): return self._inst()._op(P.OP_ROTY, r)
    def rotate_z(self, r): return self._inst()._op(P.OP_ROTZ, r)


class Instance(_Instanceable):
    def __init__(self, obj, ops=None, material=None):
        self.obj, self.ops, self.material = obj, list(ops or []), material

    def _op(self, *op):
        return Instance(self.obj, self.ops + [op], self.material)

    def clone(self, material=None):                  # instance.rs:30-37
        return Instance(self.obj, self.ops, material)

    def to_origin(self): return self._op(P.OP_ORIGIN)
    def set_x(self, v): return self._op(P.OP_SETX, v)
    def set_y(self, v): return self._op(P.OP_SETY, v)
    def set_z(self, v): return self._op(P.OP_SETZ, v)


class Face:                                          # triangle_mesh.rs:4-23
    def __init__(self, vidx, nidx=(), tidx=()):
        self.vidx, self.nidx, self.tidx = list(vidx), list(nidx), list(tidx)


class Mesh(_Instanceable):
    """`KdTree<Triangle>` (kdtree.rs:9): vertices + polygon faces + one material."""
    def __init__(self, vertices, faces, normals, uvs, material, face_range=None, shared=None):
        self.vertices, self.faces, self.normals, self.uvs, self.material = vertices, faces, normals, uvs, material
        self.face_range, self.shared = face_range, shared

    def to_unit_size(self):                          # kdtree.rs:93-99
        return Instance(self)._op(P.OP_UNIT)


class TriangleMesh:
    @staticmethod
    def new(vertices, faces, normals, uvs, material):    # triangle_mesh.rs:40-55
        return Mesh(vertices, faces, normals, uvs, material)


class Rectangle(_Instanceable):                      # rectangle.rs:15-70
    def __init__(self, a, b, c, material):
        self.abc = [np.asarray(v, dtype=np.float64) for v in (a, b, c)]
        self.material = material

    @staticmethod
    def new(abc, material):
        return Rectangle(abc[0], abc[1], abc[2], material)

    @staticmethod
    def unit_xz(material):
        X, Z = np.array([1.0, 0, 0]), np.array([0, 0, 1.0])
        return Rectangle(0.5 * (X - Z), -0.5 * (X + Z), 0.5 * (Z - X), material)

    @staticmethod
    def plane(origin, bx, by, extent, material):
        origin = np.asarray(origin, float)
        bx = np.asarray(bx, float) / np.linalg.norm(bx); by = np.asarray(by, float) / np.linalg.norm(by)
        b = origin + bx * extent[0] - by * extent[1]
        a = origin - 2.0 * by * extent[1]
        c = origin - 2.0 * bx * extent[0]
        return Rectangle(a, b, c, material)


class Sphere(_Instanceable):                         # sphere.rs:11-24
    def __init__(self, radius, material):
        assert radius != 0.0
        self.radius, self.material = float(radius), material

    @staticmethod
    def new(radius, material): return Sphere(radius, material)


class LooseTriangles:
    """Emissive faces of an .obj scene: each triangle is its own light (parser/obj.rs:92-104)."""
    def __init__(self, mesh, material):
        self.mesh, self.material = mesh, material


# ---- camera ------------------------------------------------------------------------------------
class CameraType:
    Perspective, Orthographic = 0, 1


class CameraBuilder:                                 # camera/builder.rs:8-150
    def __init__(self):
        self._origin = (0.0, 0.0, 0.0); self._towards = (0.0, 0.0, -1.0); self._up = (0.0, 1.0, 0.0)
        self._zoom = 1.0; self._lens_radius = 0.0; self._focal_length = 0.0
        self._resolution = (1024, 768); self._camera_type = CameraType.Perspective; self._vfov = 90.0
        self._color_space = ColorSpace.DCI_P3; self._pixel_filter = PixelFilter.default(); self._illuminant = illuminants.D65

    @staticmethod
    def new(): return CameraBuilder()
    def origin(self, x, y, z): self._origin = (x, y, z); return self
    def towards(self, x, y, z): self._towards = (x, y, z); return self
    def up(self, x, y, z): self._up = (x, y, z); return self
    def zoom(self, z): self._zoom = z; return self
    def lens_radius(self, r): self._lens_radius = r; return self
    def focal_length(self, f): self._focal_length = f; return self
    def resolution(self, res): self._resolution = tuple(res); return self
    def camera_type(self, t): self._camera_type = t; return self
    def vfov(self, v): self._vfov = v; return self
    def color_space(self, cs): self._color_space = cs; return self
    def pixel_filter(self, f): self._pixel_filter = f; return self
    def illuminant(self, i): self._illuminant = i; return self

    def build(self):
        assert self._lens_radius >= 0.0                                   # camera.rs:46
        assert 0.0 < self._vfov < 180.0                                   # matrices.rs:5
        assert self._resolution[0] > 0 and self._resolution[1] > 0 and self._zoom > 0.0   # matrices.rs:40-41
        d2 = sum((a - b) ** 2 for a, b in zip(self._towards, self._origin))
        assert d2 > 1e-10                                                 # matrices.rs:25
        c = Camera(); c.__dict__.update({k: v for k, v in self.__dict__.items()})
        return c


class Camera:
    @staticmethod
    def builder(): return CameraBuilder()

    @staticmethod
    def cornell_box():                               # camera.rs:139-148
        return (CameraBuilder.new().origin(278.0, 273.0, -800.0).towards(278.0, 273.0, 0.0).zoom(2.8)
                .focal_length(0.035).resolution((512, 512)).illuminant(illuminants.CORNELL).build())

    def get_resolution(self): return self._resolution


# ---- scene -------------------------------------------------------------------------------------
class Scene:                                         # src/tracer/scene.rs:18-117
    def __init__(self):
        self.objects, self.lights = [], []
        self.environment_map = None

    def add(self, obj): self.objects.append(obj)
    def add_light(self, light): self.lights.append(light)

    def set_environment_map(self, env_map, scale):
        self.environment_map = (_as_tex(env_map), float(scale))

    def num_lights(self):
        n = 0
        for l in self.lights:
            base = l.obj if isinstance(l, Instance) else l
            if isinstance(base, LooseTriangles):
                fr = base.mesh.face_range
                n += (fr[1] - fr[0]) if fr else len(base.mesh.faces)
            else:
                n += 1
        return n + (1 if self.environment_map else 0)

    # -- program emission --
    def _program(self, camera):
        w = P.ProgramWriter()
        mats, meshes = {}, {}

        texs = {}

        def mat_id(m):
            if m is None: return -1
            if id(m) not in mats:
                kw = {k: v for k, v in m.kw.items() if not k.startswith("_")}
                for slot, t in m.kw.get("_textures", {}).items():
                    if t is not None: kw[slot] = t._emit(w, texs)
                bump = m.kw.get("_bump")
                if bump is not None:
                    if id(bump) not in texs: texs[id(bump)] = w.texture(P.TEX_BUMP, width=bump.width, height=bump.height, data=bump.data)
                    kw["bump_tex"] = texs[id(bump)]
                mats[id(m)] = w.material(m.kind, **kw)
            return mats[id(m)]

        def mesh_id(m):
            key = id(m.shared) if m.shared is not None else id(m)
            if key not in meshes:
                src = m.shared if m.shared is not None else m
                faces = src.faces
                if len(faces) and isinstance(faces[0], Face):
                    has_n = any(f.nidx for f in faces); has_t = any(f.tidx for f in faces)
                    pad = lambda idx, f: list(idx) if idx else [-1] * len(f.vidx)       # a face without normals / uvs: index -1 per corner
                    meshes[key] = w.mesh(src.vertices, [f.vidx for f in faces], src.normals if has_n else None, src.uvs if has_t else None,
                                         [pad(f.nidx, f) for f in faces] if has_n else None, [pad(f.tidx, f) for f in faces] if has_t else None)
                else:
                    fn = getattr(src, "face_normals", None); ft = getattr(src, "face_uvs", None)
                    meshes[key] = w.mesh(src.vertices, faces, src.normals if fn is not None else None, src.uvs if ft is not None else None, fn, ft)
            return meshes[key]

        def emit(o, is_light):
            ops, inst_mat = (), -1
            if isinstance(o, Instance):
                ops, inst_mat, o = o.ops, mat_id(o.material), o.obj
            if isinstance(o, Mesh):
                mid, nf = mesh_id(o)
                f0, f1 = o.face_range if o.face_range else (0, nf)
                w.object(P.OBJ_KDMESH, is_light, mat_id(o.material), mid, f0, f1, (), inst_mat, ops)
            elif isinstance(o, LooseTriangles):
                mid, nf = mesh_id(o.mesh)
                f0, f1 = o.mesh.face_range if o.mesh.face_range else (0, nf)
                w.object(P.OBJ_LOOSE_TRIS, is_light, mat_id(o.material), mid, f0, f1, (), inst_mat, ops)
            elif isinstance(o, Rectangle):
                w.object(P.OBJ_RECT, is_light, mat_id(o.material), params=np.concatenate(o.abc), inst_material=inst_mat, ops=ops)
            elif isinstance(o, Sphere):
                w.object(P.OBJ_SPHERE, is_light, mat_id(o.material), params=(o.radius,), inst_material=inst_mat, ops=ops)
            else:
                raise TypeError("unsupported object %r" % (o,))

        for o in self.objects: emit(o, False)
        for o in self.lights: emit(o, True)
        if self.environment_map:
            env = _as_tex(self.environment_map[0])
            w.envmap(_tex(env).as_tuple(), self.environment_map[1], -1 if env.kind == P.TEX_SOLID else env._emit(w, texs))
        c = camera
        w.camera(c._origin, c._towards, c._up, c._zoom, c._lens_radius, c._focal_length, c._vfov, c._resolution, c._camera_type,
                 c._pixel_filter.kind, c._pixel_filter.r, c._pixel_filter.p, c._color_space, ILLUMINANTS.index(c._illuminant))
        return w.tobytes()

    # -- procedural scenes --
    @staticmethod
    def empty_box(def_color, mat_left, mat_right):   # scene/empty_box.rs:15-97
        LIGHT_EPS = 0.001
        ground = -0.8; ceiling = -ground; right = 1.0; left = -right; front = -2.0; back = 0.0; l_dim = 0.1
        s = Scene()
        light_tex = Spectrum.from_srgb(252, 201, 138)
        s.add_light(Rectangle((-l_dim, ceiling - LIGHT_EPS, 0.6 * front + l_dim), (-l_dim, ceiling - LIGHT_EPS, 0.6 * front - l_dim),
                              (l_dim, ceiling - LIGHT_EPS, 0.6 * front - l_dim), Material.light(light_tex)))
        s.add(Rectangle((left, ground, back), (left, ground, front), (left, ceiling, front), mat_left))
        s.add(Rectangle((right, ground, front), (right, ground, back), (right, ceiling, back), mat_right))
        s.add(Rectangle((left, ground, back), (right, ground, back), (right, ground, front), Material.diffuse(def_color)))
        s.add(Rectangle((left, ceiling, front), (right, ceiling, front), (right, ceiling, back), Material.diffuse(def_color)))
        s.add(Rectangle((left, ground, front), (right, ground, front), (right, ceiling, front), Material.diffuse(def_color)))
        return s

    @staticmethod
    def cornell_box():                               # scene/cornell_box.rs:8-193
        from ._cornell_data import BOX_SPEC, GREEN_SPEC, RED_SPEC, LIGHT_SPEC, QUADS, SMALL_BOX, BIG_BOX, LIGHT
        white = Spectrum.from_pts(BOX_SPEC)
        lam = lambda spec: Material.lambertian(spec)
        floor, ceil, back = lam(white), lam(white), lam(white)
        left, right = lam(Spectrum.from_pts(RED_SPEC)), lam(Spectrum.from_pts(GREEN_SPEC))
        big, small = lam(white), lam(white)
        light = Material.Light(Spectrum.from_pts(LIGHT_SPEC), illuminants.CORNELL, 1.0, False)
        s = Scene()
        s.add_light(Rectangle(LIGHT[0], LIGHT[1], LIGHT[2], light))
        quad = [Face([0, 1, 2]), Face([0, 2, 3])]
        for verts, m in zip(QUADS, (floor, ceil, back, right, left)):
            s.add(TriangleMesh.new(np.array(verts), quad, [], [], m))
        box_faces = []
        for i in range(5):
            v0 = 4 * i
            box_faces += [Face([v0, v0 + 1, v0 + 2]), Face([v0, v0 + 2, v0 + 3])]
        s.add(TriangleMesh.new(np.array(SMALL_BOX), box_faces, [], [], small))
        s.add(TriangleMesh.new(np.array(BIG_BOX), box_faces, [], [], big))
        return s


# ---- renderer ----------------------------------------------------------------------------------
class Renderer:
    """src/renderer.rs:24-99,159-244.  Same builder surface (`samples / integrator / seed / sampler /
    threads / tone_map / render`); `render()` hands the whole tile x batch schedule to the GPU through
    the C ABI (include/lumo_gpu.h) instead of a thread pool and returns a `Film`.  `threads` is
    accepted for source compatibility and ignored.  Extra, GPU-only knobs: `device`, `wave_paths`,
    `rr_delta`."""
    def __init__(self, scene, camera):
        assert scene.num_lights() != 0                                    # renderer.rs:42
        self.scene, self.camera = scene, camera
        self.resolution = camera.get_resolution()
        self.num_samples = 1                                              # renderer.rs:20
        self._integrator = Integrator.PathTrace
        self._tone_map = ToneMap.NoMap
        self._sampler = SamplerType.MultiJittered
        self._threads = 4
        import time
        self._seed = int(time.time_ns()) & 0xFFFFFFFFFFFFFFFF              # rng.rs:7-22 (time-derived default)
        self._device, self._wave_paths, self._rr_delta = 0, 0, 0.0
        self._blob = None
        self.quiet = False

    @staticmethod
    def new(scene, camera): return Renderer(scene, camera)
    def tone_map(self, tm): self._tone_map = tm; return self
    def samples(self, n): self.num_samples = int(n); return self
    def integrator(self, ig): self._integrator = ig; return self
    def seed(self, s): self._seed = int(s); return self
    def sampler(self, s): self._sampler = s; return self
    def threads(self, n): self._threads = int(n); return self
    def device(self, d): self._device = int(d); return self
    def wave_paths(self, n): self._wave_paths = int(n); return self
    def rr_delta(self, d): self._rr_delta = float(d); return self

    def blob(self):
        """Scene::build (scene.rs:33-52) + flattening: built once, natively (csrc/host)."""
        if self._blob is None:
            from . import native
            self._blob = native.build_blob(self.scene._program(self.camera))
        return self._blob

    def _banner(self, B):                                                 # renderer.rs:101-138
        p = B.params
        print("Starting to render the scene:\n"
              "\t Resolution: %d x %d\n\t Samples: %d\n\t Shadow rays: %d\n\t Integrator: %s\n\t Primitives: %d\n\t Lights: %d\n"
              "\t Seed: %d\n\t Device: cuda:%d" % (self.resolution[0], self.resolution[1], self.num_samples,
                                                   1 if self._integrator == Integrator.BDPathTrace else int(p["n_shadow_rays"]),
                                                   Integrator.NAMES[self._integrator], int(p["n_tris"]), int(p["n_lights"]), self._seed, self._device))

    def render(self):
        import time
        from . import native
        from .film import Film
        start = time.time()
        blob = self.blob()
        ctx = native.GpuContext(self._device)
        gs = None
        try:
            gs = native.GpuScene(ctx, blob)
            if not self.quiet:
                self._banner(gs.blob)
            px, sp, cnt, deltas, ms = gs.render(integrator=self._integrator, spp=self.num_samples, seed=self._seed, sampler=self._sampler,
                                                tone_map=self._tone_map.kind, tone_map_arg=self._tone_map.arg, rr_delta=self._rr_delta,
                                                wave_paths=self._wave_paths)
        finally:
            if gs is not None:
                gs.close()               # the scene before its context, also when the render raised
            ctx.close()
        if not self.quiet:                                                # renderer.rs:237-241
            print("Finished rendering in %.3f s (%d camera rays, %d total rays)" % (time.time() - start, cnt["camera_paths"], cnt["cost"]))
        return Film(px, sp, self.num_samples, self.camera._pixel_filter, self.camera._color_space, counters=cnt,
                    stats={"device_ms": ms, "tile_deltas": deltas})
