"""ctypes wrapper over oracle/liblumo_oracle.so — TEST INFRASTRUCTURE (the checker, never the product)."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")


def build(native=False):
    target = "liblumo_oracle_native.so" if native else "liblumo_oracle.so"
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, target])
    return os.path.join(ORACLE_DIR, target)


class RenderParams(C.Structure):
    _fields_ = [("integrator", C.c_int32), ("sampler", C.c_int32), ("tone_map", C.c_int32), ("rng_mode", C.c_int32),
                ("tone_map_arg", C.c_double), ("rr_delta", C.c_double), ("seed", C.c_uint64),
                ("spp_begin", C.c_uint32), ("spp_end", C.c_uint32), ("total_spp", C.c_uint32), ("threads", C.c_int32),
                ("tile_step", C.c_int32), ("pad", C.c_int32)]


_libs = {}


def lib(native=False):
    if native not in _libs:
        L = C.CDLL(build(native))
        L.oracle_scene_create.restype = C.c_void_p
        L.oracle_scene_create.argtypes = [C.c_char_p, C.c_uint64]
        L.oracle_scene_error.restype = C.c_char_p
        L.oracle_scene_error.argtypes = [C.c_void_p]
        L.oracle_scene_destroy.argtypes = [C.c_void_p]
        for f in ("oracle_export_bvh", "oracle_export_kd", "oracle_export_kd_leaf_total"):
            getattr(L, f).restype = C.c_uint64
        L.oracle_lambda_sample_one.restype = C.c_double
        L.oracle_lambda_sample_one.argtypes = [C.c_double]
        L.oracle_spectrum_sample.restype = C.c_double
        L.oracle_filter_eval.restype = C.c_double
        L.oracle_filter_eval.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]
        L.oracle_filter_integral.restype = C.c_double
        L.oracle_filter_integral.argtypes = [C.c_int, C.c_double, C.c_double]
        _libs[native] = L
    return _libs[native]


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class OracleScene:
    def __init__(self, program_bytes, native=False):
        self.L = lib(native)
        self.h = C.c_void_p(self.L.oracle_scene_create(program_bytes, len(program_bytes)))
        err = self.L.oracle_scene_error(self.h)
        if err:
            raise RuntimeError("oracle: " + err.decode())
        info = np.zeros(5, dtype=np.uint64)
        self.L.oracle_scene_info(self.h, _p(info, C.c_uint64))
        self.n_objects, self.n_lights, self.n_shadow_rays, self.res_x, self.res_y = (int(v) for v in info)

    def close(self):
        if self.h:
            self.L.oracle_scene_destroy(self.h); self.h = None

    def __del__(self):
        try: self.close()
        except Exception: pass

    def bounds(self):
        b = np.zeros(6); self.L.oracle_scene_bounds(self.h, _p(b, C.c_double)); return b

    def trace_closest(self, o, d, threads=8):
        o = np.ascontiguousarray(o, np.float64); d = np.ascontiguousarray(d, np.float64); n = o.shape[0]
        obj = np.zeros(n, np.uint32); tri = np.zeros(n, np.uint32); t = np.zeros(n); bary = np.zeros((n, 2))
        self.L.oracle_trace_closest(self.h, _p(o, C.c_double), _p(d, C.c_double), C.c_uint64(n), _p(obj, C.c_uint32), _p(tri, C.c_uint32),
                                    _p(t, C.c_double), _p(bary, C.c_double), C.c_int(threads))
        return obj, tri, t, bary

    def trace_closest_full(self, o, d, threads=8):
        o = np.ascontiguousarray(o, np.float64); d = np.ascontiguousarray(d, np.float64); n = o.shape[0]
        out = np.zeros((n, 16))
        self.L.oracle_trace_closest_full(self.h, _p(o, C.c_double), _p(d, C.c_double), C.c_uint64(n), _p(out, C.c_double), C.c_int(threads))
        return out

    def trace_any(self, o, d, t_max, threads=8):
        o = np.ascontiguousarray(o, np.float64); d = np.ascontiguousarray(d, np.float64); tm = np.ascontiguousarray(t_max, np.float64); n = o.shape[0]
        occ = np.zeros(n, np.uint8)
        self.L.oracle_trace_any(self.h, _p(o, C.c_double), _p(d, C.c_double), _p(tm, C.c_double), C.c_uint64(n), _p(occ, C.c_uint8), C.c_int(threads))
        return occ

    def trace_first_found(self, o, d, threads=8):
        o = np.ascontiguousarray(o, np.float64); d = np.ascontiguousarray(d, np.float64); n = o.shape[0]
        t = np.zeros(n)
        self.L.oracle_trace_first_found(self.h, _p(o, C.c_double), _p(d, C.c_double), C.c_uint64(n), _p(t, C.c_double), C.c_int(threads))
        return t

    def camera_rays(self, raster_xy, lens_uv=None):
        r = np.ascontiguousarray(raster_xy, np.float64); n = r.shape[0]
        l = np.zeros((n, 2)) if lens_uv is None else np.ascontiguousarray(lens_uv, np.float64)
        o = np.zeros((n, 3)); d = np.zeros((n, 3))
        self.L.oracle_camera_rays(self.h, _p(r, C.c_double), _p(l, C.c_double), C.c_uint64(n), _p(o, C.c_double), _p(d, C.c_double))
        return o, d

    def texture_count(self):
        self.L.oracle_texture_count.restype = C.c_uint64
        return int(self.L.oracle_texture_count(self.h))

    def texture_kind(self, tex): return int(self.L.oracle_texture_kind(self.h, C.c_uint64(tex)))

    def texture_eval(self, tex, uv, lambda_u=0.3):
        uv = np.ascontiguousarray(uv, np.float64); n = uv.shape[0]
        out = np.zeros((n, 4))
        self.L.oracle_texture_eval(self.h, C.c_uint64(tex), _p(uv, C.c_double), C.c_uint64(n), C.c_double(lambda_u), _p(out, C.c_double))
        return out

    def bump_eval(self, tex, uv):
        uv = np.ascontiguousarray(uv, np.float64); n = uv.shape[0]
        out = np.zeros((n, 3))
        self.L.oracle_bump_eval(self.h, C.c_uint64(tex), _p(uv, C.c_double), C.c_uint64(n), _p(out, C.c_double))
        return out

    def counters(self, reset=False):
        c = np.zeros(8, np.uint64); self.L.oracle_counters(_p(c, C.c_uint64), C.c_int(int(reset)))
        return dict(zip(("tlas_nodes", "inst", "kd_nodes", "leaf_idx", "tri_tests", "sphere_tests", "closest", "occlusion"), (int(v) for v in c)))

    def render(self, integrator=0, spp=1, seed=1, sampler=2, tone_map=0, tone_map_arg=0.0, rng_mode=0, rr_delta=0.0,
               threads=8, spp_begin=0, spp_end=None, total_spp=None, tile_step=1):
        """tile_step > 1 (reference schedule only, rng_mode 0): render every tile_step-th 16x16 tile."""
        total = spp if total_spp is None else total_spp
        end = total if spp_end is None else spp_end
        P = RenderParams(integrator, sampler, tone_map, rng_mode, tone_map_arg, rr_delta, seed, spp_begin, end, total, threads, tile_step, 0)
        W, H = self.res_x, self.res_y
        pixels = np.zeros((H, W, 4)); splats = np.zeros((H, W, 3)); cnt = np.zeros(4, np.uint64)
        ntiles = ((W + 15) // 16) * ((H + 15) // 16)
        deltas = np.zeros(ntiles)
        rc = self.L.oracle_render(self.h, C.byref(P), _p(pixels, C.c_double), _p(splats, C.c_double), _p(cnt, C.c_uint64), _p(deltas, C.c_double))
        assert rc == 0
        return pixels, splats, dict(zip(("camera_paths", "closest", "occlusion", "cost"), (int(v) for v in cnt))), deltas

    def export_bvh(self, which):
        n = int(self.L.oracle_export_bvh(self.h, C.c_int(which), C.c_uint64(0), None, None, None, None, None, C.c_uint64(0)))
        nobj = self.n_objects if which == 0 else self.n_lights
        bounds = np.zeros((n, 6)); right = np.zeros(n, np.int64); first = np.zeros(n, np.int64); count = np.zeros(n, np.int64)
        leaf = np.zeros(nobj, np.int64)
        self.L.oracle_export_bvh(self.h, C.c_int(which), C.c_uint64(n), _p(bounds, C.c_double), _p(right, C.c_int64), _p(first, C.c_int64),
                                 _p(count, C.c_int64), _p(leaf, C.c_int64), C.c_uint64(nobj))
        return dict(bounds=bounds, right=right, first=first, count=count, leaf=leaf)

    def export_kd(self, which, idx):
        nt = C.c_uint64(0)
        n = int(self.L.oracle_export_kd(self.h, C.c_int(which), C.c_uint64(idx), C.c_uint64(0), None, None, None, None, None, None, None,
                                        C.c_uint64(0), C.byref(nt), None, C.c_uint64(0)))
        if n == 0:
            return None
        nl = int(self.L.oracle_export_kd_leaf_total(self.h, C.c_int(which), C.c_uint64(idx)))
        ntri = int(nt.value)
        axis = np.zeros(n, np.int64); point = np.zeros(n); right = np.zeros(n, np.int64); leaf = np.zeros(n, np.int64)
        first = np.zeros(n, np.int64); count = np.zeros(n, np.int64); ll = np.zeros(nl, np.int64); tv = np.zeros((ntri, 9))
        self.L.oracle_export_kd(self.h, C.c_int(which), C.c_uint64(idx), C.c_uint64(n), _p(axis, C.c_int64), _p(point, C.c_double),
                                _p(right, C.c_int64), _p(leaf, C.c_int64), _p(first, C.c_int64), _p(count, C.c_int64), _p(ll, C.c_int64),
                                C.c_uint64(nl), C.byref(nt), _p(tv, C.c_double), C.c_uint64(ntri))
        return dict(axis=axis, point=point, right=right, leaf=leaf, first=first, count=count, leaf_list=ll, tri_verts=tv)

    def export_instance(self, which, idx):
        m = np.zeros((4, 4)); inv = np.zeros((4, 4))
        ok = self.L.oracle_export_instance(self.h, C.c_int(which), C.c_uint64(idx), _p(m, C.c_double), _p(inv, C.c_double))
        return (m, inv) if ok else None

    def export_alias(self):
        n = self.n_lights
        prob = np.zeros(n); alias = np.zeros(n, np.int64); pdf = np.zeros(n)
        self.L.oracle_export_alias(self.h, _p(prob, C.c_double), _p(alias, C.c_int64), _p(pdf, C.c_double))
        return prob, alias, pdf
