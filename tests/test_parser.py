"""`.obj` / `.mtl` ingest (lumo_b200/parser.py) against the reference's parser semantics (src/parser/obj.rs, parser/mtl.rs,
parser/mtl/task.rs): grouping into kd-trees, emissive groups as loose triangle lights, index conventions, material
mapping.  The parsed scene goes through the same scene program as API-built scenes, so oracle and device see it alike."""
import math
import numpy as np
import pytest
import oracle_lib
from lumo_b200 import parser, native, Material, Spectrum, CameraBuilder, program as P

OBJ = """
# a room: floor quad (fan-triangulated), a pyramid with normals and uvs, a lamp
mtllib room.mtl
v -2 0 -2
v  2 0 -2
v  2 0  2
v -2 0  2
v -0.5 0 -0.5
v  0.5 0 -0.5
v  0.5 0  0.5
v -0.5 0  0.5
v  0 1 0
v -0.3 2.5 -0.3
v  0.3 2.5 -0.3
v  0.3 2.5  0.3
v -0.3 2.5  0.3
vn 0 1 0
vn 0 0 0
vt 0 0
vt 1 0
vt 1 1
o floor
usemtl grey
f 1 2 3 4
o pyramid
usemtl gold
f 5//1 6//1 9//1
f 6/1/1 7/2/1 9/3/1
usemtl glass
f -7 -6 -5
f 8 5 9
g lamp
usemtl lamp
f 10 11 12 13
"""
MTL = """
newmtl grey
Kd 0.5 0.5 0.5
Ns 0
newmtl gold
Kd 0 0 0
Ks 1.0 0.8 0.3
Ns 400
Ni 0.4
illum 5
newmtl glass
Tf 1 1 1
Ks 1 1 1
Ns 900
Ni 1.45
illum 7
newmtl lamp
Ke 10 10 10
newmtl grey
Kd 1 0 0
"""


def _scene():
    return parser.scene_from_obj(OBJ, mtl_resolver=lambda name: MTL if name == "room.mtl" else None)


def test_scene_structure():
    s = _scene()
    # groups: floor | gold (2 faces) | glass (2 faces) | lamp -> 3 kd-tree objects + one light group (obj.rs:49-66)
    assert len(s.objects) == 3 and len(s.lights) == 1
    assert s.num_lights() == 2                                    # the lamp quad is two triangles = two lights (obj.rs:97-104)
    kinds = [o.material.kind for o in s.objects]
    assert kinds == [P.M_MFDIFFUSE, P.M_MFCONDUCTOR, P.M_MFDIELECTRIC]
    grey, gold, glass = (o.material.kw for o in s.objects)
    assert grey["roughness"] == 1.0                               # Ns 0 -> 1 - sqrt(0)/30
    assert abs(gold["roughness"] - (1.0 - math.sqrt(400.0) / 30.0)) < 1e-15 and gold["eta"] == 0.4
    assert glass["roughness"] == 0.0 and glass["eta"] == 1.45     # Ns 900 -> 0; transparent + fresnel (illum 7)
    assert np.allclose(grey["kd"], Spectrum.from_rgb(0.5, 0.5, 0.5).as_tuple())      # the first `newmtl grey` wins (mtl.rs:134)
    base = s.objects[0].shared
    assert len(base.vertices) == 13 and len(base.faces) == 2 + 2 + 2 + 2
    assert np.allclose(base.normals[1], [0, 0, 1])                # degenerate normal -> +Z (obj.rs:127-133)
    f = base.faces
    assert f[2].vidx == [4, 5, 8] and f[2].nidx == [0, 0, 0] and f[2].tidx == []      # "5//1"
    assert f[3].tidx == [0, 1, 2]
    assert f[4].vidx == [6, 7, 8]                                 # negative indices count from the end (parser.rs:60-64)


def test_blob_and_oracle_agree_on_parsed_scene():
    s = _scene()
    cam = CameraBuilder.new().origin(0.0, 1.5, 4.0).towards(0.0, 0.8, 0.0).resolution((32, 24)).build()
    prog = s._program(cam)
    B = native.Blob(native.build_blob(prog))
    assert int(B.params["n_objects"]) == 3 and int(B.params["n_lights"]) == 2 and len(B.kd_trees) == 3
    assert [int(t["n_tris"]) for t in B.kd_trees] == [2, 2, 2]
    sh = B.tri_shade[int(B.kd_trees[1]["tri_base"]):int(B.kd_trees[1]["tri_base"]) + 2]
    assert list(sh["flags"]) == [1, 3]                            # first gold face: normals only; second: normals + uvs
    assert list(B.tri_shade[int(B.kd_trees[2]["tri_base"]):int(B.kd_trees[2]["tri_base"]) + 2]["flags"]) == [0, 0]
    O = oracle_lib.OracleScene(prog)
    assert (O.n_objects, O.n_lights) == (3, 2)
    rs = np.random.RandomState(2)
    o, d = O.camera_rays(rs.rand(2000, 2) * np.array([32, 24]))
    obj, tri, t, _ = O.trace_closest(o, d)
    seen = set(int(v) for v in np.unique(obj[obj != 0xFFFFFFFF]))
    assert 0 in seen and (seen & {1, 2}) and (seen & {3, 4})     # floor, pyramid and lamp are seen
    full = O.trace_closest_full(o, d)
    hit_pyr = (obj == 1) & (tri == 1)
    if hit_pyr.any():
        assert np.all((full[hit_pyr, 10] >= 0) & (full[hit_pyr, 10] <= 1))        # interpolated uv of the textured face


def test_mesh_from_obj_and_errors():
    m = parser.mesh_from_obj("v 0 0 0\nv 1 0 0\nv 0 1 0\nv 1 1 0\nf 1 2 4 3\n", Material.lambertian(Spectrum.from_rgb(0.5, 0.5, 0.5)))
    assert len(m.faces) == 2 and m.faces[1].vidx == [0, 3, 2]
    with pytest.raises(parser.ObjError, match="Could not find material"):
        parser.scene_from_obj("v 0 0 0\nusemtl nope\n", mtl_src="newmtl a\nKd 1 1 1\n")
    with pytest.raises(parser.ObjError, match="Could not parse"):
        parser.mesh_from_obj("v 0 zero 0\n", Material.Blank)
    with pytest.raises(parser.ObjError):
        parser.scene_from_obj("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n", mtl_src="newmtl a\nKd 1 1 1\n")      # faces before any usemtl


@pytest.mark.gpu
def test_parsed_scene_renders_like_the_oracle(gpu_ctx):
    s = _scene()
    cam = CameraBuilder.new().origin(0.0, 1.5, 4.0).towards(0.0, 0.8, 0.0).resolution((32, 24)).build()
    prog = s._program(cam); blob = native.build_blob(prog)
    O = oracle_lib.OracleScene(prog); G = native.GpuScene(gpu_ctx, blob)
    rs = np.random.RandomState(4)
    o, d = O.camera_rays(rs.rand(5000, 2) * np.array([32, 24]))
    e = O.trace_closest(o, d); g = G.trace_closest(o, d)
    assert np.array_equal(e[0], g[0]) and np.array_equal(e[1], g[1]) and np.array_equal(e[2].view(np.uint64), g[2].view(np.uint64))
    epx, _, ec, _ = O.render(integrator=0, spp=4, seed=3, rng_mode=1)
    gpx, _, gc, _, _ = G.render(integrator=0, spp=4, seed=3)
    assert gc["camera_paths"] == ec["camera_paths"] and gc["nonfinite"] == 0
    with np.errstate(invalid="ignore", divide="ignore"):
        ei = np.nan_to_num(epx[..., :3] / epx[..., 3:4]); gi = np.nan_to_num(gpx[..., :3] / gpx[..., 3:4])
    assert abs(gi.mean() - ei.mean()) <= 0.05 * abs(ei.mean()) + 1e-9
    G.close(); O.close()


def test_texture_maps_become_image_textures():
    """map_Kd / map_Ks / map_Ke / map_Bump (mtl/task.rs:30-80, mtl.rs:59-91): images served by name, decoded, and attached
    as Texture::Image / bump map; with map_ks=False the map_Ks image is the occlusion-roughness-metalness map."""
    from lumo_b200.image import encode_png
    rs = np.random.RandomState(2)
    pal = np.array([[200, 40, 40], [40, 200, 40], [250, 250, 250]], np.uint8)
    files = {"tex/floor.png": encode_png(pal[rs.randint(0, 3, size=(4, 4))]), "orm.png": encode_png(np.full((2, 2, 3), [255, 64, 128], np.uint8)),
             "nrm.png": encode_png(np.full((2, 2, 3), [128, 128, 255], np.uint8)), "lamp.png": encode_png(np.full((2, 2, 3), 255, np.uint8))}
    mtl = MTL.replace("newmtl grey\nKd 0.5 0.5 0.5\nNs 0", "newmtl grey\nKd 0.5 0.5 0.5\nNs 0\nmap_Kd tex\\floor.png\nmap_Bump nrm.png\nmap_Ks orm.png") \
             .replace("Ke 10 10 10", "Ke 0 0 0\nmap_Ke lamp.png")
    asked = []
    def images(name): asked.append(name); return files[name]
    s = parser.scene_from_obj(OBJ, mtl_resolver=lambda name: mtl, image_resolver=images, map_ks=False)
    assert asked[:3] == ["tex/floor.png", "nrm.png", "orm.png"]                     # backslashes normalised (task.rs:31)
    grey = s.objects[0].material
    assert grey.kw["_textures"]["kd_tex"] is not None and grey.kw["_textures"]["kd_tex"].kind == P.TEX_IMAGE and grey.kw["_bump"] is not None
    assert grey.kw["_textures"]["ks_tex"] is None and abs(grey.kw["roughness"] - 64 / 256) < 1e-12 and abs(grey.kw["k"] - 128 / 256) < 1e-12
    assert s.lights and s.lights[0].material.kw["_textures"]["ke_tex"].kind == P.TEX_IMAGE     # map_Ke alone makes a light (mtl.rs:60)
    s2 = parser.scene_from_obj(OBJ, mtl_resolver=lambda name: mtl, image_resolver=images, map_ks=True)
    assert s2.objects[0].material.kw["_textures"]["ks_tex"].kind == P.TEX_IMAGE
    cam = CameraBuilder.new().origin(0, 2, 6).towards(0, 1, 0).resolution((32, 24)).build()
    prog = s._program(cam)
    B = native.Blob(native.build_blob(prog))
    assert (B.textures["kind"] == P.TEX_IMAGE).sum() == 2 and (B.textures["kind"] == P.TEX_BUMP).sum() == 1
    O = oracle_lib.OracleScene(prog)
    assert O.texture_count() == 3
    O.close()
    with pytest.warns(UserWarning):
        parser.scene_from_obj(OBJ, mtl_resolver=lambda name: mtl)                    # no resolver: maps ignored with a warning


def _zip_bytes(files):
    import io, zipfile
    buf = io.BytesIO()
    with zipfile.ZipFile(buf, "w", zipfile.ZIP_DEFLATED) as z:
        for name, data in files.items():
            z.writestr(name, data)
    return buf.getvalue()


def test_reference_entry_points_over_zip_archives(tmp_path, monkeypatch):
    """parser::{mesh_from_path, mesh_from_url, texture_from_url, scene_from_url, scene_from_file} (parser.rs:125-265):
    archives cached under ./scenes/, the member chosen by a case-insensitive suffix match, duplicates / misses as errors,
    the `mtllib` argument read before the .obj's own `mtllib` lines, texture maps and the RGBE environment map from the
    same archive."""
    from lumo_b200.image import encode_png
    monkeypatch.chdir(tmp_path)
    png = encode_png(np.full((2, 2, 3), [250, 128, 10], np.uint8))
    hdr = b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y 2 +X 2\n" + bytes([128, 128, 128, 129] * 4)
    mtl = MTL.replace("newmtl grey\nKd 0.5 0.5 0.5\nNs 0", "newmtl grey\nKd 0.5 0.5 0.5\nNs 0\nmap_Kd Tex\\Floor.PNG")
    extra = "newmtl gold\nKd 1 1 1\n"                                 # given as the `mtllib` argument: read first, so this gold wins
    (tmp_path / "scenes").mkdir()
    (tmp_path / "scenes" / "room.zip").write_bytes(_zip_bytes({"Room/Room.OBJ": OBJ, "Room/room.mtl": mtl, "Room/extra.mtl": extra,
                                                               "Room/tex/floor.png": png, "Room/sky.hdr": hdr}))
    url = "https://example.invalid/assets/room.zip"
    s = parser.scene_from_url(url, "room.obj", True, "extra.mtl", ("sky.hdr", 2.0))
    assert len(s.objects) == 3 and s.num_lights() == 3 and s.environment_map is not None and s.environment_map[1] == 2.0   # two lamp triangles + the environment
    assert s.objects[0].material.kw["_textures"]["kd_tex"].kind == P.TEX_IMAGE
    assert s.objects[1].material.kind == P.M_MFDIFFUSE                # the argument's `gold` (diffuse) shadows room.mtl's conductor
    s2 = parser.scene_from_file("./scenes/room.zip", "room.obj", True)
    assert s2.objects[1].material.kind == P.M_MFCONDUCTOR and s2.environment_map is None and s2.num_lights() == 2
    img = parser.texture_from_url(url, "floor.png")
    assert img.data.shape[:2] == (2, 2)
    m = parser.mesh_from_url(url, Material.Blank)
    assert len(m.faces) == 8
    (tmp_path / "quad.obj").write_text("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nf 1 2 3 4\n")
    assert len(parser.mesh_from_path(str(tmp_path / "quad.obj"), Material.Blank).faces) == 2
    # errors of the reference, in its words
    with pytest.raises(parser.ObjError, match="Can only load scenes from .zip"): parser.scene_from_url("http://x/room.obj", "room.obj", True)
    with pytest.raises(parser.ObjError, match="Can only parse .obj files"): parser.scene_from_url(url, "room.mtl", True)
    with pytest.raises(parser.ObjError, match="Can only load .png files"): parser.texture_from_url(url, "sky.hdr")
    with pytest.raises(parser.ObjError, match="Can only extract textures from zip archives"): parser.texture_from_url("http://x/a.obj", "a.png")
    with pytest.raises(parser.ObjError, match="Found multiple .mtl"): parser._extract_zip((tmp_path / "scenes" / "room.zip").read_bytes(), ".mtl")
    with pytest.raises(parser.ObjError, match="Could not find nothing.obj"): parser.scene_from_file("./scenes/room.zip", "nothing.obj", True)
    (tmp_path / "scenes" / "thing.bin").write_bytes(b"x")
    with pytest.raises(parser.ObjError, match="Bad URL"): parser.mesh_from_url("http://x/thing.bin", Material.Blank)
    with pytest.raises(parser.ObjError, match="could not download"): parser.mesh_from_url("http://127.0.0.1:9/none.obj", Material.Blank)   # nothing cached, nothing listening
