"""`.obj` / `.mtl` ingest — host-side mirror of the reference's parser (src/parser.rs:125-265, parser/obj.rs:6-198,
parser/mtl.rs:10-146, parser/mtl/task.rs:16-116).  Same grammar subset and the same scene structure: one shared
`TriangleMesh`; every `g` / `o` / `usemtl` group becomes its own kd-tree (obj.rs:49-66,92-107); groups whose
material emits become loose triangle lights, one light per triangle (obj.rs:97-104).

Texture maps (`map_Kd`, `map_Ks`, `map_Ke`, `map_Bump`, mtl/task.rs:30-80) are read through an `image_resolver(name) ->
PNG bytes` callback (the reference pulls them out of the scene's zip archive) and become `Texture::Image` / bump maps on
the device; without a resolver they are ignored with a warning.  With `map_ks=False` a `map_Ks` image is the
occlusion/roughness/metalness map of the reference: its mean green / blue channels become roughness and k
(mtl/task.rs:60-68).

The reference's own entry points exist under their names — `mesh_from_path`, `mesh_from_url`, `texture_from_url`,
`scene_from_url`, `scene_from_file` (parser.rs:125-265) — over zip archives cached in `./scenes/` (`_check_cached`,
`_extract_zip`); a missing archive is fetched with urllib where a network exists and is an `ObjError` otherwise."""
import io
import math
import os
import zipfile
import warnings
import numpy as np
from .api import Scene, Material, Texture, TriangleMesh, Mesh, Face, LooseTriangles
from .image import Image, decode_png
from .spectrum import Spectrum


class ObjError(ValueError):
    pass


def _lines(src):
    if isinstance(src, (bytes, bytearray)):
        src = src.decode("utf-8", "replace")
    if isinstance(src, str) and "\n" not in src and src.endswith((".obj", ".mtl")):
        src = open(src).read()
    for line in io.StringIO(src):
        line = line.strip()
        if not line or line.startswith("#"):
            continue
        yield line.split()


def _floats(tokens, n):
    try:
        return [float(t) for t in tokens[1:1 + n]]
    except ValueError:
        raise ObjError("Could not parse double in file")
    

def _idx(tok, length):                                   # parser.rs:52-67: 1-based, negative = relative to the end
    try:
        i = int(tok)
    except ValueError:
        raise ObjError("Could not parse index in file")
    return i - 1 if i > 0 else length + i


def _parse_face(tokens, nv, nn, nt):                     # obj.rs:150-198: fan triangulation
    v, n, t = [], [], []
    for tok in tokens[1:]:
        a = tok.split("/")
        v.append(_idx(a[0], nv))
        if len(a) > 1 and a[1] != "":
            t.append(_idx(a[1], nt))
        if len(a) > 2:
            n.append(_idx(a[2], nn))
    out = []
    for i in range(1, len(v) - 1):
        c = (0, i, i + 1)
        out.append(Face([v[k] for k in c], [n[k] for k in c] if n else [], [t[k] for k in c] if t else []))
    return out


def _parse_common(tokens, V, N, T, F):                   # obj.rs:113-147
    k = tokens[0]
    if k == "v":
        V.append(_floats(tokens, 3))
    elif k == "vn":
        x, y, z = _floats(tokens, 3)
        l2 = x * x + y * y + z * z
        if l2 == 0.0:
            N.append([0.0, 0.0, 1.0])
        else:
            l = math.sqrt(max(l2, 0.0)); N.append([x / l, y / l, z / l])
    elif k == "vt":
        T.append(_floats(tokens, 2))
    elif k == "f":
        F.extend(_parse_face(tokens, len(V), len(N), len(T)))


def mesh_from_obj(src, material):
    """obj::load_file (obj.rs:6-24): the whole file as one mesh with one material."""
    V, N, T, F = [], [], [], []
    for tokens in _lines(src):
        _parse_common(tokens, V, N, T, F)
    return TriangleMesh.new(np.asarray(V, np.float64).reshape(-1, 3), F, np.asarray(N, np.float64).reshape(-1, 3), np.asarray(T, np.float64).reshape(-1, 2), material)


def load_mtl(src, materials=None, indices=None, image_resolver=None, map_ks=True):
    """mtl::load_file + MtlTaskExecutor::exec + MtlConfig::build_material (mtl.rs:52-146, mtl/task.rs:16-116).
    Returns (materials, name -> index); the first definition of a name wins (mtl.rs:134)."""
    materials = [] if materials is None else materials
    indices = {} if indices is None else indices
    blocks, block = [], []
    for tokens in _lines(src):
        if tokens[0] == "newmtl" and block:
            blocks.append(block); block = []
        block.append(tokens)
    if block:
        blocks.append(block)
    for block in blocks:
        cfg = dict(Kd=Spectrum.BLACK(), Ks=Spectrum.BLACK(), Ke=Spectrum.BLACK(), Tf=Spectrum.BLACK(), eta=1.5, k=0.0, roughness=1.0, fresnel=False, transparent=False,
                   map_Kd=None, map_Ks=None, map_Ke=None, map_Bump=None)
        name = ""
        for tokens in block:
            k = tokens[0]
            if k == "newmtl": name = tokens[1]
            elif k in ("Kd", "Ks", "Ke", "Tf"): cfg[k] = Spectrum.from_rgb(*_floats(tokens, 3))
            elif k == "Ni": cfg["eta"] = _floats(tokens, 1)[0]
            elif k == "Ns": cfg["roughness"] = 1.0 - math.sqrt(min(_floats(tokens, 1)[0], 900.0)) / 30.0      # "blender uses this mapping"
            elif k == "illum":
                il = int(_floats(tokens, 1)[0])
                if il == 5: cfg["fresnel"] = True
                elif il == 6: cfg["transparent"] = True
                elif il == 7: cfg["fresnel"] = cfg["transparent"] = True
            elif k in ("map_Kd", "map_Ks", "map_Ke", "map_Bump"):                                       # mtl/task.rs:30-80
                tex_name = " ".join(tokens[1:]).replace("\\", "/")
                if image_resolver is None:
                    warnings.warn("%s of material %r ignored: no image_resolver given" % (k, name)); continue
                rgb = decode_png(image_resolver(tex_name))
                if k == "map_Bump": cfg[k] = Image.bump_from_rgb8(rgb)
                elif k == "map_Ks" and not map_ks:
                    orm = Image.mean_vec3_from_rgb8(rgb)                                             # occlusion, roughness, metalness
                    cfg["roughness"] = float(orm[1]); cfg["k"] = float(orm[2]); cfg["Ks"] = Spectrum.WHITE()
                else: cfg[k] = Image.from_rgb8(rgb)
        if name in indices:
            continue
        if not cfg["Ke"].is_black() or cfg["map_Ke"] is not None:                                       # mtl.rs:59-91
            m = Material.light(Texture.Image(cfg["map_Ke"]) if cfg["map_Ke"] is not None else cfg["Ke"])
        else:
            kd = Texture.Image(cfg["map_Kd"]) if cfg["map_Kd"] is not None else cfg["Kd"]
            ks = Texture.Image(cfg["map_Ks"]) if cfg["map_Ks"] is not None else cfg["Ks"]
            m = Material.microfacet(cfg["roughness"], cfg["eta"], cfg["k"], cfg["transparent"], cfg["fresnel"], kd, ks, cfg["Tf"], bump_map=cfg["map_Bump"])
        materials.append(m); indices[name] = len(materials) - 1
    return materials, indices


def scene_from_obj(obj_src, mtl_src=None, env_map=None, mtl_resolver=None, image_resolver=None, map_ks=True):
    """parser::scene_from_file + obj::load_scene (parser.rs:206-265, obj.rs:27-110).  `mtl_src`: the text of the material
    library named by the caller (the `mtllib` argument of the reference); `mtl_resolver(name) -> text` serves `mtllib`
    lines inside the .obj.  `env_map`: (Spectrum or Texture, scale) or None;
    `image_resolver(name) -> PNG bytes` serves the materials' texture maps."""
    obj_lines = list(_lines(obj_src))
    materials, indices = [], {}
    if mtl_src is not None:
        load_mtl(mtl_src, materials, indices, image_resolver, map_ks)
    for tokens in obj_lines:
        if tokens[0] == "mtllib":
            if mtl_resolver is None:
                raise ObjError("Could not find %s in the archive" % tokens[1])
            load_mtl(mtl_resolver(tokens[1]), materials, indices, image_resolver, map_ks)
    V, N, T = [], [], []
    faces, groups, midx = [], [], None
    for tokens in obj_lines:
        k = tokens[0]
        if k in ("g", "o"):
            if faces:
                groups.append((faces, midx)); faces = []; midx = None
        elif k == "usemtl":
            if faces:
                groups.append((faces, midx)); faces = []
            if tokens[1] not in indices:
                raise ObjError("Could not find material %s" % tokens[1])
            midx = indices[tokens[1]]
        else:
            _parse_common(tokens, V, N, T, faces)
    groups.append((faces, midx))
    base = Mesh(np.asarray(V, np.float64).reshape(-1, 3), [f for g, _ in groups for f in g], np.asarray(N, np.float64).reshape(-1, 3),
                np.asarray(T, np.float64).reshape(-1, 2), None)
    scene = Scene()
    f0 = 0
    for g, mi in groups:
        f1 = f0 + len(g)
        if mi is None:
            if g:
                raise ObjError("faces without a material (no usemtl before them)")     # the reference indexes materials[usize::MAX] and panics
            continue
        if g:
            chunk = Mesh(base.vertices, base.faces, base.normals, base.uvs, materials[mi], face_range=(f0, f1), shared=base)
            if materials[mi].is_light():
                scene.add_light(LooseTriangles(chunk, materials[mi]))
            else:
                scene.add(chunk)
        f0 = f1
    if env_map is not None:
        scene.set_environment_map(env_map[0], env_map[1])
    return scene


# ---------------------------------------------------------------------------------------------------------------------
# The reference's public entry points (parser.rs:125-265): files, cached URLs, zip archives.

SCENE_DIR = "./scenes/"                                    # parser.rs:17


def _extract_zip(data, end_match):
    """parser.rs:82-113: the one archive member whose lower-cased name ends with `end_match`; several or none is an error."""
    end_match = end_match.lower()
    found = None
    try:
        with zipfile.ZipFile(io.BytesIO(data)) as z:
            for info in z.infolist():
                if info.filename.lower().endswith(end_match):
                    if found is not None:
                        raise ObjError("Found multiple %s in the archive" % end_match)
                    found = z.read(info)
    except zipfile.BadZipFile as e:
        raise ObjError(str(e))
    if not found:
        raise ObjError("Could not find %s in the archive" % end_match)
    return found


def _check_cached(url):
    """parser.rs:149-165: `SCENE_DIR` + last path component of the URL; downloaded once if absent."""
    os.makedirs(SCENE_DIR, exist_ok=True)
    path = SCENE_DIR + url.rsplit("/", 1)[-1]
    if not os.path.exists(path):
        print('"%s" not found, downloading from "%s"' % (path, url))
        try:
            import urllib.request
            with urllib.request.urlopen(url, timeout=60) as r:
                body = r.read()
        except Exception as e:
            raise ObjError("could not download %s: %s" % (url, e))
        with open(path, "wb") as f:
            f.write(body)
    return path


def _read(path):
    with open(path, "rb") as f:
        return f.read()


def mesh_from_path(path, material):
    """parser.rs:125-128"""
    print('Loading .OBJ file "%s"' % path)
    return mesh_from_obj(_read(path), material)


def mesh_from_url(url, material):
    """parser.rs:132-147: a plain .obj or the single .obj inside a .zip, cached under SCENE_DIR."""
    path = _check_cached(url)
    print('Loading .OBJ from "%s"' % path)
    data = _read(path)
    if url.endswith(".zip"):
        data = _extract_zip(data, ".obj")
    elif not url.endswith(".obj"):
        raise ObjError("Bad URL, or at least does not end with .zip or .obj")
    return mesh_from_obj(data, material)


def texture_from_url(url, tex_name):
    """parser.rs:168-180: `tex_name` (.png) out of the cached .zip as an Image<Spectrum>."""
    if not tex_name.endswith(".png"):
        raise ObjError("Can only load .png files")
    if not url.endswith(".zip"):
        raise ObjError("Can only extract textures from zip archives")
    path = _check_cached(url)
    print('Loading texture "%s" from "%s"' % (tex_name, path))
    return Image.from_png(_extract_zip(_read(path), tex_name))


def scene_from_url(url, obj_name, map_ks, mtllib=None, env_map=None):
    """parser.rs:184-200"""
    if not url.endswith(".zip"):
        raise ObjError("Can only load scenes from .zip")
    if not obj_name.endswith(".obj"):
        raise ObjError("Can only parse .obj files")
    return scene_from_file(_check_cached(url), obj_name, map_ks, mtllib, env_map)


def scene_from_file(path, obj_name, map_ks, mtllib=None, env_map=None):
    """parser.rs:204-265: scene `obj_name` from the zip at `path`; `mtllib` names a material library in the archive that is
    read first, `mtllib` lines of the .obj follow; texture maps come from the same archive; `env_map` = (file, scale) is a
    flat RGBE image in the archive."""
    print('Loading scene "%s" from "%s"' % (obj_name, path))
    archive = _read(path)
    obj_bytes = _extract_zip(archive, obj_name)
    mtl_src = _extract_zip(archive, mtllib) if mtllib is not None else None
    scene = scene_from_obj(obj_bytes, mtl_src, None, mtl_resolver=lambda name: _extract_zip(archive, name),
                           image_resolver=lambda name: _extract_zip(archive, name), map_ks=map_ks)
    if env_map is not None:
        scene.set_environment_map(Texture.Image(Image.from_hdri_bytes(_extract_zip(archive, env_map[0]))), env_map[1])
    return scene
