"""ctypes bindings: liblumo_host.so (scene builder) and liblumo_gpu.so (the C ABI of
include/lumo_gpu.h).  The GPU library is mandatory for every compute entry point: if it is missing
or no CUDA device is present the call raises — there is no CPU fallback in this package."""
import ctypes as C
import os
import weakref
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
HOST_SO = os.path.join(HERE, "liblumo_host.so")
GPU_SO = os.environ.get("LUMO_GPU_SO", os.path.join(HERE, "liblumo_gpu.so"))   # override: A/B timing of two builds

SECTIONS = ["tlas_nodes", "tlas_leaf", "objects", "instances", "kd_trees", "kd_nodes", "kd_leaf", "tri_verts", "tri_shade",
            "normals", "uvs", "rects", "spheres", "materials", "tables", "lights", "textures", "tex_pixels", "tex_f64",
            "ah_nodes", "ah_prims", "obj_path_off", "obj_path"]

TLAS_NODE = np.dtype([("lo", "<f8", 3), ("hi", "<f8", 3), ("right", "<u4"), ("first", "<u4"), ("count", "<u4"), ("pad", "<u4")])
OBJECT = np.dtype([("kind", "<u4"), ("geom", "<u4"), ("inst", "<i4"), ("material", "<i4"), ("rect", "<u4"), ("pad", "<u4", 3)])
INSTANCE = np.dtype([("inv", "<f8", 12), ("m", "<f8", 12), ("nrm", "<f8", 9), ("pad", "<f8")])
KD_TREE = np.dtype([("lo", "<f8", 3), ("hi", "<f8", 3), ("root", "<u4"), ("tri_base", "<u4"), ("n_tris", "<u4"), ("pad", "<u4")])
KD_NODE = np.dtype([("point", "<f8"), ("a", "<u4"), ("b", "<u4")])
TRI_VERTS = np.dtype([("a", "<f8", 3), ("b", "<f8", 3), ("c", "<f8", 3), ("pad", "<f8")])
TRI_SHADE = np.dtype([("n", "<u4", 3), ("t", "<u4", 3), ("flags", "<u4"), ("pad", "<u4")])
RECT = np.dtype([("origin", "<f8", 3), ("b0", "<f8", 3), ("b1", "<f8", 3), ("pad", "<f8")])
SPHERE = np.dtype([("radius", "<f8"), ("pad", "<f8")])
MATERIAL = np.dtype([("kind", "<u4"), ("flags", "<u4"), ("roughness", "<f8"), ("kd", "<f4", 4), ("ks", "<f4", 4), ("tf", "<f4", 4), ("ke", "<f4", 4),
                     ("eta_table", "<u4"), ("k_table", "<u4"), ("illum_table", "<u4"), ("kd_tex", "<u4"), ("scale", "<f8"),
                     ("ks_tex", "<u4"), ("tf_tex", "<u4"), ("ke_tex", "<u4"), ("bump_tex", "<u4"), ("pad", "<u8")])
TEXTURE = np.dtype([("kind", "<u4"), ("a", "<u4"), ("b", "<u4"), ("width", "<u4"), ("height", "<u4"), ("pad", "<u4"), ("data", "<u8"),
                    ("spec", "<f4", 4), ("scale", "<f8"), ("pad2", "<f8")])
AH_NODE = np.dtype([("lo_x", "<f4", 4), ("lo_y", "<f4", 4), ("lo_z", "<f4", 4), ("hi_x", "<f4", 4), ("hi_y", "<f4", 4), ("hi_z", "<f4", 4),
                    ("child", "<u4", 4), ("pad", "<u4", 4)])
AH_PRIM = np.dtype([("tri", "<u4"), ("obj", "<u4")])
LIGHT = np.dtype([("alias_prob", "<f8"), ("pdf", "<f8"), ("area", "<f8"), ("alias", "<u4"), ("pad", "<u4"), ("bound_c", "<f8", 3), ("bound_r", "<f8")])
CAMERA = np.dtype([("screen_to_raster_m", "<f8", 16), ("screen_to_raster_inv", "<f8", 16), ("camera_to_screen_m", "<f8", 16),
                   ("camera_to_screen_inv", "<f8", 16), ("world_to_camera_m", "<f8", 16), ("world_to_camera_inv", "<f8", 16),
                   ("lens_radius", "<f8"), ("focal_length", "<f8"), ("image_plane_area", "<f8"), ("lens_area", "<f8"),
                   ("res_x", "<u4"), ("res_y", "<u4"), ("ortho", "<u4"), ("pad", "<u4")])
FILM = np.dtype([("xyz_to_rgb", "<f8", 9), ("wb", "<f8", 9), ("filter_r", "<f8"), ("filter_p", "<f8"), ("filter_gr", "<f8"),
                 ("filter_kind", "<u4"), ("r_disc", "<u4"), ("color_space", "<u4"), ("pad", "<u4")])
PARAMS = np.dtype([("n_objects", "<u4"), ("n_lights", "<u4"), ("lights_root", "<u4"), ("n_shadow_rays", "<u4"),
                   ("n_tlas_nodes", "<u4"), ("n_kd_trees", "<u4"), ("n_materials", "<u4"), ("n_tris", "<u4"),
                   ("bounds_lo", "<f8", 3), ("bounds_hi", "<f8", 3), ("camera", CAMERA), ("film", FILM)])
SECREF = np.dtype([("offset", "<u8"), ("bytes", "<u8"), ("count", "<u8"), ("pad", "<u8")])
HEADER = np.dtype([("magic", "<u8"), ("version", "<u4"), ("n_sections", "<u4"), ("total_bytes", "<u8"), ("pad", "<u8"),
                   ("params", PARAMS), ("sec", SECREF, len(SECTIONS))])
_SEC_DTYPES = {"tlas_nodes": TLAS_NODE, "tlas_leaf": np.dtype("<u4"), "objects": OBJECT, "instances": INSTANCE, "kd_trees": KD_TREE,
               "kd_nodes": KD_NODE, "kd_leaf": np.dtype("<u4"), "tri_verts": TRI_VERTS, "tri_shade": TRI_SHADE, "normals": np.dtype(("<f8", 3)),
               "uvs": np.dtype(("<f8", 2)), "rects": RECT, "spheres": SPHERE, "materials": MATERIAL, "tables": np.dtype(("<f8", 96)), "lights": LIGHT,
               "textures": TEXTURE, "tex_pixels": np.dtype(("<f4", 4)), "tex_f64": np.dtype("<f8"),
               "ah_nodes": AH_NODE, "ah_prims": AH_PRIM, "obj_path_off": np.dtype("<u4"), "obj_path": np.dtype("<u4")}


class Blob:
    """Parsed view (numpy, zero-copy) of a device scene blob."""
    def __init__(self, data):
        self.data = data
        self.header = np.frombuffer(data, dtype=HEADER, count=1)[0]
        assert int(self.header["magic"]) == 0x31424F4C424D554C, "not a lumo scene blob"
        self.params = self.header["params"]
        for i, name in enumerate(SECTIONS):
            ref = self.header["sec"][i]
            dt = _SEC_DTYPES[name]
            setattr(self, name, np.frombuffer(data, dtype=dt, count=int(ref["count"]), offset=int(ref["offset"])))

    def __len__(self):
        return len(self.data)


_host = None


def host_lib():
    global _host
    if _host is None:
        if not os.path.exists(HOST_SO):
            from . import build
            build.build_host()
        L = C.CDLL(HOST_SO)
        L.lumo_host_build.restype = C.c_int32
        L.lumo_host_build.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        L.lumo_host_free.argtypes = [C.c_void_p]
        L.lumo_host_last_error.restype = C.c_char_p
        L.lumo_host_math_eval.argtypes = [C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_uint64, C.POINTER(C.c_double)]
        L.lumo_host_math_eval.restype = None
        _host = L
    return _host


def build_blob(program_bytes):
    """scene program -> device blob bytes (kd-trees, BVHs, alias table built natively)."""
    L = host_lib()
    p = C.c_void_p(); n = C.c_uint64()
    rc = L.lumo_host_build(program_bytes, len(program_bytes), C.byref(p), C.byref(n))
    if rc != 0:
        raise RuntimeError("lumo_host_build: " + L.lumo_host_last_error().decode())
    try:
        return C.string_at(p, n.value)
    finally:
        L.lumo_host_free(p)


MATH_FN = {"sin": 0, "cos": 1, "atan2": 2, "acos": 3, "atanh": 4, "cosh": 5, "exp": 6, "log": 7, "pow": 8}


def host_math(fn, x, y=None):
    """csrc/common/lumo_math.h evaluated by the host library (the same source the device kernels compile)."""
    x = np.asarray(x, np.float64)
    flat = np.ascontiguousarray(x).reshape(-1); out = np.empty_like(flat)
    yy = None if y is None else np.ascontiguousarray(np.broadcast_to(np.asarray(y, np.float64), x.shape)).reshape(-1)
    if flat.size:
        host_lib().lumo_host_math_eval(MATH_FN[fn], _dp(flat), None if yy is None else _dp(yy), flat.size, _dp(out))
    return out.reshape(x.shape)


# ---- GPU library ---------------------------------------------------------------------------------
class RenderParams(C.Structure):        # include/lumo_gpu.h lumo_render_params
    _fields_ = [("integrator", C.c_int32), ("sampler", C.c_int32), ("tone_map", C.c_int32), ("flags", C.c_int32),
                ("tone_map_arg", C.c_double), ("rr_delta", C.c_double), ("seed", C.c_uint64),
                ("spp_begin", C.c_uint32), ("spp_end", C.c_uint32), ("total_spp", C.c_uint32), ("wave_paths", C.c_uint32)]


class FilmAccum(C.Structure):           # include/lumo_gpu.h lumo_film_accum
    _fields_ = [("pixels", C.POINTER(C.c_double)), ("splats", C.POINTER(C.c_double)), ("counters", C.c_uint64 * 10),
                ("tile_deltas", C.POINTER(C.c_double)), ("device_ms", C.c_double)]


_gpu = None
COUNTER_NAMES = ("camera_paths", "closest", "occlusion", "cost", "gpu_launches", "max_depth", "iterations", "nonfinite", "bdpt_cut_subpaths", "occlusion_mismatches")


def gpu_lib():
    """Loads liblumo_gpu.so.  Raises if it was not built — never falls back to anything else."""
    global _gpu
    if _gpu is None:
        if not os.path.exists(GPU_SO):
            raise RuntimeError("liblumo_gpu.so is not built (run `python -m lumo_b200.build gpu`); lumo_b200 has no CPU fallback")
        L = C.CDLL(GPU_SO)
        L.lumo_gpu_last_error.restype = C.c_char_p
        vp = C.c_void_p
        L.lumo_gpu_device_count.argtypes = [C.POINTER(C.c_int32)]
        L.lumo_gpu_ctx_create.argtypes = [C.c_int32, C.POINTER(vp)]
        L.lumo_gpu_ctx_destroy.argtypes = [vp]
        L.lumo_gpu_scene_upload.argtypes = [vp, C.c_char_p, C.c_uint64, C.POINTER(vp)]
        L.lumo_gpu_scene_destroy.argtypes = [vp]
        dp, u32p, u8p = C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)
        L.lumo_gpu_trace_closest.argtypes = [vp, dp, dp, dp, C.c_uint64, u32p, u32p, dp, dp]
        L.lumo_gpu_trace_any.argtypes = [vp, dp, dp, dp, C.c_uint64, u8p]
        L.lumo_gpu_trace_first_found.argtypes = [vp, dp, dp, C.c_uint64, dp]
        L.lumo_gpu_render.argtypes = [vp, C.POINTER(RenderParams), C.POINTER(FilmAccum)]
        L.lumo_gpu_render_dev.argtypes = [vp, C.POINTER(RenderParams), vp, vp, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
        L.lumo_gpu_trace_closest_dev.argtypes = [vp, vp, vp, C.c_uint64, vp, vp, vp, vp, C.POINTER(C.c_float)]
        L.lumo_gpu_ctx_count_visits.argtypes = [vp, C.c_int32]
        L.lumo_gpu_ctx_set_stream.argtypes = [vp, vp]
        L.lumo_gpu_ctx_kernel_times.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
        L.lumo_gpu_ctx_kernel_times.restype = C.c_int32
        L.lumo_gpu_ctx_iter_log.argtypes = [vp, C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(C.c_uint32)]
        L.lumo_gpu_ctx_iter_log.restype = C.c_int32
        L.lumo_gpu_ctx_set_stream.restype = C.c_int32
        L.lumo_gpu_ctx_visits.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.lumo_gpu_film_encode_dev.argtypes = [vp, vp, vp, C.c_uint64, C.c_double, C.c_double, C.c_int32, u8p, C.POINTER(C.c_float)]
        L.lumo_gpu_film_encode.argtypes = [vp, dp, dp, C.c_uint64, C.c_double, C.c_double, C.c_int32, u8p]
        L.lumo_gpu_render_multi.argtypes = [C.POINTER(vp), C.c_int32, C.POINTER(RenderParams), C.POINTER(FilmAccum)]
        L.lumo_gpu_ctx_closest_mode.argtypes = [vp, C.c_int32]; L.lumo_gpu_ctx_closest_mode.restype = C.c_int32
        L.lumo_gpu_ctx_closest_stats.argtypes = [vp, C.POINTER(C.c_uint64)]; L.lumo_gpu_ctx_closest_stats.restype = C.c_int32
        L.lumo_gpu_ctx_shade_stats.argtypes = [vp, C.POINTER(C.c_uint64)]; L.lumo_gpu_ctx_shade_stats.restype = C.c_int32
        L.lumo_gpu_ctx_occlusion_mode.argtypes = [vp, C.c_int32]; L.lumo_gpu_ctx_occlusion_mode.restype = C.c_int32
        L.lumo_gpu_ctx_occlusion_stats.argtypes = [vp, C.POINTER(C.c_uint64)]; L.lumo_gpu_ctx_occlusion_stats.restype = C.c_int32
        L.lumo_gpu_fp64_peak.argtypes = [vp, dp, dp]; L.lumo_gpu_fp64_peak.restype = C.c_int32
        L.lumo_gpu_math_eval.argtypes = [vp, C.c_int32, dp, dp, C.c_uint64, dp]
        L.lumo_gpu_math_eval.restype = C.c_int32
        L.lumo_gpu_sample_range.argtypes = [C.c_int32, C.c_int32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        for f in ("lumo_gpu_render_dev", "lumo_gpu_trace_closest_dev", "lumo_gpu_ctx_count_visits", "lumo_gpu_ctx_visits","lumo_gpu_device_count", "lumo_gpu_ctx_create", "lumo_gpu_ctx_destroy", "lumo_gpu_scene_upload", "lumo_gpu_scene_destroy",
                  "lumo_gpu_trace_closest", "lumo_gpu_trace_any", "lumo_gpu_trace_first_found", "lumo_gpu_render", "lumo_gpu_film_encode", "lumo_gpu_film_encode_dev", "lumo_gpu_render_multi", "lumo_gpu_sample_range"):
            getattr(L, f).restype = C.c_int32
        _gpu = L
    return _gpu


def _check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (what, rc, gpu_lib().lumo_gpu_last_error().decode()))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class GpuContext:
    def __init__(self, device=0):
        L = gpu_lib()
        self.h = C.c_void_p()
        self._scenes = weakref.WeakSet()     # scenes uploaded through this context: closed before it
        _check(L.lumo_gpu_ctx_create(device, C.byref(self.h)), "lumo_gpu_ctx_create")

    def close(self):
        if self.h:
            for s in list(self._scenes):     # a scene's destructor touches its context: never leave one open behind a destroyed context
                s.close()
            gpu_lib().lumo_gpu_ctx_destroy(self.h); self.h = None

    def set_stream(self, cuda_stream_ptr):
        """Order the context's work on the caller's stream.  A NULL handle — which is what torch's DEFAULT stream reports as
        `.cuda_stream` — selects the context's own non-blocking stream, which is NOT ordered against the default stream: a caller
        that also touches the film with torch ops or NCCL must run on a side stream and pass that (bench.py does)."""
        _check(gpu_lib().lumo_gpu_ctx_set_stream(self.h, C.c_void_p(cuda_stream_ptr)), "lumo_gpu_ctx_set_stream")

    def count_visits(self, enable):
        _check(gpu_lib().lumo_gpu_ctx_count_visits(self.h, int(enable)), "lumo_gpu_ctx_count_visits")

    def visits(self):
        """(closest-hit kernels, occlusion kernels) visit counters"""
        out = (C.c_uint64 * 12)()
        _check(gpu_lib().lumo_gpu_ctx_visits(self.h, out), "lumo_gpu_ctx_visits")
        names = ("tlas_nodes", "inst", "kd_nodes", "leaf_idx", "tri_tests", "sphere_tests")
        return dict(zip(names, (int(v) for v in out[:6]))), dict(zip(names, (int(v) for v in out[6:])))

    def closest_mode(self, mode):
        """0: world-space BVH + the reference traversal on the winning object, reference traversal where not provably equal (default); 1: reference traversal for every ray."""
        _check(gpu_lib().lumo_gpu_ctx_closest_mode(self.h, int(mode)), "lumo_gpu_ctx_closest_mode")

    def shade_stats(self):
        """of the last render: bounces that ran NEE, NEE terms evaluated, shadow rays queued, bounces shaded"""
        out = (C.c_uint64 * 4)()
        _check(gpu_lib().lumo_gpu_ctx_shade_stats(self.h, out), "lumo_gpu_ctx_shade_stats")
        return dict(zip(("nee_bounces", "nee_terms", "shadow_queued", "bounces"), (int(v) for v in out)))

    def closest_stats(self):
        out = (C.c_uint64 * 14)()
        _check(gpu_lib().lumo_gpu_ctx_closest_stats(self.h, out), "lumo_gpu_ctx_closest_stats")
        names = ("nodes", "prims", "tri_tests", "sphere_tests", "fallback", "rays", "why_stack_overflow", "why_object_tie", "why_object_box", "why_object_full_hit_rejected",
                 "why_object_distance", "why_light_tie", "why_light_box", "why_light_hit")
        return dict(zip(names, (int(v) for v in out)))

    def occlusion_mode(self, mode):
        """0: occlusion BVH + confirmation (default); 1: the reference's traversal for shadow rays; 2: both, disagreements counted."""
        _check(gpu_lib().lumo_gpu_ctx_occlusion_mode(self.h, int(mode)), "lumo_gpu_ctx_occlusion_mode")

    def occlusion_stats(self):
        out = (C.c_uint64 * 9)()
        _check(gpu_lib().lumo_gpu_ctx_occlusion_stats(self.h, out), "lumo_gpu_ctx_occlusion_stats")
        return dict(zip(("nodes", "prims", "tri_tests", "sphere_tests", "candidates", "confirmed", "fallback", "mismatches", "robust"), (int(v) for v in out)))

    def iter_log(self):
        cap = 16384
        out = (C.c_uint32 * (2 * cap))(); n = C.c_uint32(0)
        _check(gpu_lib().lumo_gpu_ctx_iter_log(self.h, out, cap, C.byref(n)), "lumo_gpu_ctx_iter_log")
        k = min(int(n.value), cap)
        return [(int(out[2 * i]), int(out[2 * i + 1])) for i in range(k)]

    def kernel_times(self):
        ms = (C.c_double * 4)(); n = (C.c_uint64 * 4)()
        _check(gpu_lib().lumo_gpu_ctx_kernel_times(self.h, ms, n), "lumo_gpu_ctx_kernel_times")
        return {k: (ms[i], int(n[i])) for i, k in enumerate(("regen", "trace", "shade", "occlude"))}

    def fp64_peak(self):
        """Measured DFMA throughput of this GPU (TFLOP/s) — the FP64 issue ceiling next to the bandwidth roofline."""
        tf = C.c_double(0.0); ms = C.c_double(0.0)
        _check(gpu_lib().lumo_gpu_fp64_peak(self.h, C.byref(tf), C.byref(ms)), "lumo_gpu_fp64_peak")
        return {"tflops": tf.value, "ms": ms.value, "what": "8 independent DFMA chains per thread, 2048 threads per SM, 2 flops per DFMA; best of 3 launches"}

    def math_eval(self, fn, x, y=None):
        """csrc/common/lumo_math.h evaluated on the device (parity hook)."""
        x = np.ascontiguousarray(x, np.float64).ravel(); out = np.empty_like(x)
        yy = None if y is None else np.ascontiguousarray(np.broadcast_to(np.asarray(y, np.float64), x.shape))
        _check(gpu_lib().lumo_gpu_math_eval(self.h, MATH_FN[fn], _dp(x), None if yy is None else _dp(yy), x.size, _dp(out)), "lumo_gpu_math_eval")
        return out

    def film_encode(self, pixels, splats, splat_scale, filter_integral, transfer=0):
        """Film::rgb_image on the device from HOST accumulators [H,W,4] / [H,W,3] f64 -> uint8 [H,W,3]."""
        pixels = np.ascontiguousarray(pixels, dtype=np.float64); splats = np.ascontiguousarray(splats, dtype=np.float64)
        assert pixels.shape[-1] == 4 and splats.shape[-1] == 3 and pixels.shape[:-1] == splats.shape[:-1]
        rgb = np.empty(splats.shape, dtype=np.uint8)
        _check(gpu_lib().lumo_gpu_film_encode(self.h, _dp(pixels), _dp(splats), pixels.size // 4, float(splat_scale), float(filter_integral), int(transfer),
                                              rgb.ctypes.data_as(C.POINTER(C.c_uint8))), "lumo_gpu_film_encode")
        return rgb

    def film_encode_dev(self, pixels_ptr, splats_ptr, shape, splat_scale, filter_integral, transfer=0):
        """Same from DEVICE accumulators (raw pointers, as left by render_dev / the NCCL reduce); shape = (H, W).  Returns (uint8 [H,W,3], kernel ms)."""
        rgb = np.empty((int(shape[0]), int(shape[1]), 3), dtype=np.uint8); ms = C.c_float(0.0)
        _check(gpu_lib().lumo_gpu_film_encode_dev(self.h, C.c_void_p(pixels_ptr), C.c_void_p(splats_ptr), rgb.size // 3, float(splat_scale), float(filter_integral),
                                                  int(transfer), rgb.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(ms)), "lumo_gpu_film_encode_dev")
        return rgb, ms.value

    def __del__(self):
        try: self.close()
        except Exception: pass


class GpuScene:
    """An uploaded scene blob; entry points mirror include/lumo_gpu.h one to one (host pointers)."""
    def __init__(self, ctx, blob_bytes, host_ptr=None):
        """blob_bytes: the scene blob; host_ptr: optional address of a (e.g. pinned) host copy of the same bytes to upload from."""
        self.ctx = ctx
        self.h = C.c_void_p()
        src = C.c_char_p(blob_bytes) if host_ptr is None else C.cast(C.c_void_p(host_ptr), C.c_char_p)
        _check(gpu_lib().lumo_gpu_scene_upload(ctx.h, src, len(blob_bytes), C.byref(self.h)), "lumo_gpu_scene_upload")   # validates the blob
        ctx._scenes.add(self)
        self.blob = Blob(blob_bytes)
        cam = self.blob.params["camera"]
        self.res_x, self.res_y = int(cam["res_x"]), int(cam["res_y"])

    def close(self):
        if self.h:
            gpu_lib().lumo_gpu_scene_destroy(self.h); self.h = None

    def __del__(self):
        try: self.close()
        except Exception: pass

    def trace_closest(self, o, d, t_max=None):
        o = np.ascontiguousarray(o, np.float64); d = np.ascontiguousarray(d, np.float64); n = o.shape[0]
        tm = np.full(n, np.inf) if t_max is None else np.ascontiguousarray(t_max, np.float64)
        obj = np.zeros(n, np.uint32); tri = np.zeros(n, np.uint32); t = np.zeros(n); bary = np.zeros((n, 2))
        _check(gpu_lib().lumo_gpu_trace_closest(self.h, _dp(o), _dp(d), _dp(tm), n, obj.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                tri.ctypes.data_as(C.POINTER(C.c_uint32)), _dp(t), _dp(bary)), "lumo_gpu_trace_closest")
        return obj, tri, t, bary

    def trace_any(self, o, d, t_max):
        o = np.ascontiguousarray(o, np.float64); d = np.ascontiguousarray(d, np.float64); tm = np.ascontiguousarray(t_max, np.float64); n = o.shape[0]
        occ = np.zeros(n, np.uint8)
        _check(gpu_lib().lumo_gpu_trace_any(self.h, _dp(o), _dp(d), _dp(tm), n, occ.ctypes.data_as(C.POINTER(C.c_uint8))), "lumo_gpu_trace_any")
        return occ

    def trace_first_found(self, o, d):
        o = np.ascontiguousarray(o, np.float64); d = np.ascontiguousarray(d, np.float64); n = o.shape[0]
        t = np.zeros(n)
        _check(gpu_lib().lumo_gpu_trace_first_found(self.h, _dp(o), _dp(d), n, _dp(t)), "lumo_gpu_trace_first_found")
        return t

    def render(self, integrator=0, spp=1, seed=1, sampler=2, tone_map=0, tone_map_arg=0.0, rr_delta=0.0, spp_begin=0, spp_end=None,
               total_spp=None, wave_paths=0, flags=0, pixels=None, splats=None):
        """pixels / splats: optional caller-owned host arrays [H,W,4] / [H,W,3] f64 (e.g. pinned, reused across frames); overwritten."""
        total = spp if total_spp is None else total_spp
        end = total if spp_end is None else spp_end
        P = RenderParams(integrator, sampler, tone_map, flags, tone_map_arg, rr_delta, seed, spp_begin, end, total, wave_paths)
        W, H = self.res_x, self.res_y
        if pixels is None: pixels = np.zeros((H, W, 4))
        if splats is None: splats = np.zeros((H, W, 3))
        assert pixels.shape == (H, W, 4) and splats.shape == (H, W, 3) and pixels.dtype == np.float64 and splats.dtype == np.float64
        assert pixels.flags["C_CONTIGUOUS"] and splats.flags["C_CONTIGUOUS"]
        ntiles = ((W + 15) // 16) * ((H + 15) // 16)
        deltas = np.zeros(ntiles)
        out = FilmAccum(_dp(pixels), _dp(splats), (C.c_uint64 * 10)(), _dp(deltas), 0.0)
        _check(gpu_lib().lumo_gpu_render(self.h, C.byref(P), C.byref(out)), "lumo_gpu_render")
        return pixels, splats, dict(zip(COUNTER_NAMES, (int(v) for v in out.counters))), deltas, out.device_ms

    def render_dev(self, pixels_ptr, splats_ptr, integrator=0, spp=1, seed=1, sampler=2, tone_map=0, tone_map_arg=0.0, rr_delta=0.0,
                   spp_begin=0, spp_end=None, total_spp=None, wave_paths=0):
        """Film accumulators stay in the caller's DEVICE buffers (raw pointers, e.g. torch tensors' data_ptr())."""
        total = spp if total_spp is None else total_spp
        end = total if spp_end is None else spp_end
        P = RenderParams(integrator, sampler, tone_map, 0, tone_map_arg, rr_delta, seed, spp_begin, end, total, wave_paths)
        cnt = (C.c_uint64 * 10)(); ms = C.c_double(0.0)
        _check(gpu_lib().lumo_gpu_render_dev(self.h, C.byref(P), C.c_void_p(pixels_ptr), C.c_void_p(splats_ptr), cnt, C.byref(ms)), "lumo_gpu_render_dev")
        return dict(zip(COUNTER_NAMES, (int(v) for v in cnt))), ms.value

    def trace_closest_dev(self, o_ptr, d_ptr, n, obj_ptr, tri_ptr, t_ptr, bary_ptr):
        ms = C.c_float(0.0)
        _check(gpu_lib().lumo_gpu_trace_closest_dev(self.h, C.c_void_p(o_ptr), C.c_void_p(d_ptr), n, C.c_void_p(obj_ptr), C.c_void_p(tri_ptr),
                                                    C.c_void_p(t_ptr), C.c_void_p(bary_ptr), C.byref(ms)), "lumo_gpu_trace_closest_dev")
        return ms.value


def render_multi(scenes, integrator=0, spp=1, seed=1, sampler=2, tone_map=0, tone_map_arg=0.0, rr_delta=0.0, spp_begin=0, spp_end=None,
                 total_spp=None, wave_paths=0):
    """lumo_gpu_render_multi: `scenes` = the same blob uploaded as one GpuScene per context / GPU; the sample range is
    split over them and the films are summed on scenes[0]'s GPU.  Returns what GpuScene.render returns."""
    total = spp if total_spp is None else total_spp
    end = total if spp_end is None else spp_end
    P = RenderParams(integrator, sampler, tone_map, 0, tone_map_arg, rr_delta, seed, spp_begin, end, total, wave_paths)
    W, H = scenes[0].res_x, scenes[0].res_y
    pixels = np.zeros((H, W, 4)); splats = np.zeros((H, W, 3))
    deltas = np.zeros(((W + 15) // 16) * ((H + 15) // 16))
    out = FilmAccum(_dp(pixels), _dp(splats), (C.c_uint64 * 10)(), _dp(deltas), 0.0)
    hs = (C.c_void_p * len(scenes))(*[s.h for s in scenes])
    _check(gpu_lib().lumo_gpu_render_multi(hs, len(scenes), C.byref(P), C.byref(out)), "lumo_gpu_render_multi")
    return pixels, splats, dict(zip(COUNTER_NAMES, (int(v) for v in out.counters))), deltas, out.device_ms


def sample_range(g, n, begin, end):
    """lumo_gpu_sample_range: [begin, end) of GPU g of n inside lumo_gpu_render_multi (host arithmetic, no device needed)."""
    b = C.c_uint32(0); e = C.c_uint32(0)
    _check(gpu_lib().lumo_gpu_sample_range(g, n, begin, end, C.byref(b), C.byref(e)), "lumo_gpu_sample_range")
    return b.value, e.value
