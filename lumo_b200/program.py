"""Scene program: the serialised form of the Scene/Camera construction calls a lumo user makes.

The reference builds a scene by calling Rust constructors (`Scene::add`, `Rectangle::new`,
`TriangleMesh::new`, `.to_unit_size().to_origin().translate(..)`, `Camera::builder()...`;
examples/*.rs).  Here those calls are recorded into a flat little-endian byte string which the
native host library (csrc/host, `lumo_host_build`) executes: it builds the kd-trees and object
BVHs with lumo's algorithms and flattens everything into the device scene blob.

Layout: magic "LUMOPRG1", u32 version, u32 n_records, then records
  u32 tag, u32 0, u64 nbytes, payload (8-byte words: i64 or f64).
Texture records (src/tracer/texture.rs:23-38) precede the materials that use them; a material / environment-map
record may end with texture ids (-1 = the Solid spectrum given inline), older programs simply end earlier.
"""
import struct
import numpy as np

MAGIC = b"LUMOPRG1"
TAG_MATERIAL, TAG_MESH, TAG_OBJECT, TAG_ENVMAP, TAG_CAMERA, TAG_TEXTURE = 1, 2, 3, 4, 5, 6
TEX_SOLID, TEX_CHECKER, TEX_MARBLE, TEX_IMAGE, TEX_MANDELBROT, TEX_BUMP = range(6)
OBJ_KDMESH, OBJ_RECT, OBJ_SPHERE, OBJ_LOOSE_TRIS = 0, 1, 2, 3
(OP_UNIT, OP_ORIGIN, OP_SETX, OP_SETY, OP_SETZ, OP_TRANSLATE, OP_SCALE, OP_ROTX, OP_ROTY, OP_ROTZ) = range(10)
# material kinds
M_BLANK, M_LAMBERTIAN, M_MFDIFFUSE, M_MFCONDUCTOR, M_MFDIELECTRIC, M_LIGHT = range(6)
# eta / k table selectors
ETA_CONST, ETA_GLASS, ETA_DIAMOND, ETA_MIRROR, K_MIRROR = range(5)


def _w(*vals):
    """pack a mixed list of int / float as 8-byte words"""
    out = bytearray()
    for v in vals:
        if isinstance(v, (int, np.integer)) and not isinstance(v, bool):
            out += struct.pack("<q", int(v))
        else:
            out += struct.pack("<d", float(v))
    return bytes(out)


class ProgramWriter:
    def __init__(self):
        self.records = []
        self.n_materials = 0
        self.n_meshes = 0
        self.n_textures = 0

    def _rec(self, tag, payload):
        assert len(payload) % 8 == 0
        self.records.append(struct.pack("<IIQ", tag, 0, len(payload)) + payload)

    def material(self, kind, roughness=1.0, eta_kind=ETA_CONST, eta=1.5, k_kind=ETA_CONST, k=0.0,
                 kd=(0, 0, 0, 0), ks=(0, 0, 0, 0), tf=(0, 0, 0, 0), ke=(0, 0, 0, 0), illuminant=2, scale=1.0, two_sided=0,
                 kd_tex=-1, ks_tex=-1, tf_tex=-1, ke_tex=-1, bump_tex=-1):
        p = _w(int(kind), float(roughness), int(eta_kind), float(eta), int(k_kind), float(k))
        for s in (kd, ks, tf, ke):
            p += _w(*[float(v) for v in s])
        p += _w(int(illuminant), float(scale), int(two_sided))
        p += _w(int(kd_tex), int(ks_tex), int(tf_tex), int(ke_tex), int(bump_tex))
        self._rec(TAG_MATERIAL, p)
        self.n_materials += 1
        return self.n_materials - 1

    def mesh(self, vertices, faces, normals=None, uvs=None, face_normals=None, face_uvs=None):
        """faces: list of polygons (index lists) or an (F,k) int array; face_normals/face_uvs:
        same shape as faces (indices into normals/uvs) or None."""
        v = np.ascontiguousarray(vertices, dtype=np.float64).reshape(-1, 3)
        n = np.zeros((0, 3)) if normals is None else np.ascontiguousarray(normals, dtype=np.float64).reshape(-1, 3)
        t = np.zeros((0, 2)) if uvs is None else np.ascontiguousarray(uvs, dtype=np.float64).reshape(-1, 2)

        def flat(fs):
            if isinstance(fs, np.ndarray) and fs.ndim == 2:
                off = np.arange(fs.shape[0] + 1, dtype=np.int64) * fs.shape[1]
                return off, fs.astype(np.int64).reshape(-1)
            off = np.zeros(len(fs) + 1, dtype=np.int64)
            off[1:] = np.cumsum([len(f) for f in fs])
            return off, np.concatenate([np.asarray(f, dtype=np.int64) for f in fs]) if len(fs) else np.zeros(0, np.int64)

        off, vidx = flat(faces)
        has_n = face_normals is not None
        has_t = face_uvs is not None
        p = _w(v.shape[0], n.shape[0], t.shape[0], len(off) - 1, len(vidx), int(has_n), int(has_t))
        p += v.tobytes() + n.tobytes() + t.tobytes() + off.tobytes() + vidx.tobytes()
        if has_n:
            p += flat(face_normals)[1].tobytes()
        if has_t:
            p += flat(face_uvs)[1].tobytes()
        self._rec(TAG_MESH, p)
        self.n_meshes += 1
        return self.n_meshes - 1, len(off) - 1

    def object(self, kind, is_light, material, mesh=-1, face_begin=0, face_end=0, params=(), inst_material=-1, ops=()):
        prm = list(params) + [0.0] * (9 - len(params))
        p = _w(int(kind), int(bool(is_light)), int(material), int(mesh), int(face_begin), int(face_end))
        p += _w(*[float(x) for x in prm])
        p += _w(int(inst_material), len(ops))
        for op in ops:
            o = list(op) + [0.0] * (4 - len(op))
            p += _w(int(o[0]), float(o[1]), float(o[2]), float(o[3]))
        self._rec(TAG_OBJECT, p)

    def envmap(self, spec, scale, tex=-1):
        self._rec(TAG_ENVMAP, _w(*[float(v) for v in spec], float(scale), int(tex)))

    def texture(self, kind, spec=(0, 0, 0, 0), a=-1, b=-1, scale=1.0, seed=0, width=0, height=0, data=None):
        """kind TEX_*; spec: Solid / Marble colour, Image mean; a, b, scale: Checkerboard; seed: Marble (Perlin::new);
        data: Image -> float array [H, W, 4] of Spectrum coefficients, Bump -> float array [H, W, 3] of unit normals."""
        p = _w(int(kind), *[float(v) for v in spec], int(a), int(b), float(scale), int(seed), int(width), int(height))
        if data is not None:
            p += np.ascontiguousarray(data, dtype=np.float64).tobytes()
        self._rec(TAG_TEXTURE, p)
        self.n_textures += 1
        return self.n_textures - 1

    def camera(self, origin, towards, up, zoom, lens_radius, focal_length, vfov, resolution, camera_type,
               filter_kind, filter_r, filter_p, color_space, illuminant):
        p = _w(*[float(v) for v in origin], *[float(v) for v in towards], *[float(v) for v in up],
               float(zoom), float(lens_radius), float(focal_length), float(vfov),
               int(resolution[0]), int(resolution[1]), int(camera_type), int(filter_kind), float(filter_r), float(filter_p),
               int(color_space), int(illuminant))
        self._rec(TAG_CAMERA, p)

    def tobytes(self):
        return MAGIC + struct.pack("<II", 1, len(self.records)) + b"".join(self.records)
