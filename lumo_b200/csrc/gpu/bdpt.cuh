// Bidirectional path tracing on the device: bd_path_trace.rs:23-290, bd_path_trace/{path_gen,vertex,mis,
// measure}.rs and the camera importance functions camera.rs:167-388.
//
// Vertex-buffer design: the light and camera subpaths of a batch of camera samples are walked as a wavefront (one
// bounce of every live subpath per iteration) into vertex arrays in HBM: LUMO_BDPT_MAXV vertices per subpath in place, and
// for the rare longer ones (specular chains) further blocks of LUMO_BDPT_MAXV from an overflow pool, up to the reference's
// own cap of 1024 (bd_path_trace.rs:7); a subpath is only cut — and counted, counters[8] — if the pool runs dry; the (s,t)
// terms of every sample
// are then evaluated in parallel, one launch sequence per term class — emission and NEE one thread per term, light
// tracing and connections through a visibility-ray queue — and a finish kernel retires the sample into the film.
// Traversal calls (Scene::hit, hit_t, hit_light) are the same faithful routines the wavefront kernels use; visible()
// needs the reference's first-found distance, i.e. scene_hit_t, not an any-hit boolean.
#pragma once
#include "wavefront.cuh"

namespace lumo_dev {

#define LUMO_BDPT_MAXV 64
#define LUMO_BDPT_MAX_DEPTH 1024u   /* bd_path_trace.rs:7 */
#define LUMO_BDPT_OVF_BLOCKS (LUMO_BDPT_MAX_DEPTH / LUMO_BDPT_MAXV)   /* overflow blocks a subpath can own (vertices 64 .. 1087) */

// `delta`: Vertex::is_delta (vertex.rs:90-97) evaluated once when the vertex is made — it depends on the material and the hero
// wavelength only, and the hero wavelength never changes along a sample (termination zeroes the secondary ones).
struct Vtx { DevHit h; C4 gathered; double pdf_fwd, pdf_bck; D3 wo; int light; int delta; };   // vertex.rs:5-12
// The vertices of one subpath: the first LUMO_BDPT_MAXV in the batch's vertex buffer, the rest in blocks of the overflow pool.
struct VtxArr {
    Vtx* base; const uint32_t* tab; Vtx* pool;
    __device__ __forceinline__ Vtx& operator[](int k) const {
        if (k < LUMO_BDPT_MAXV) return base[k];
        return pool[(size_t)tab[k / LUMO_BDPT_MAXV - 1] * LUMO_BDPT_MAXV + (size_t)(k % LUMO_BDPT_MAXV)];
    }
};
__device__ __forceinline__ VtxArr vtx_none() { VtxArr a; a.base = nullptr; a.tab = nullptr; a.pool = nullptr; return a; }

__device__ const LumoMaterial g_blank_material = {LMAT_BLANK, 0u, 1.0, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, 0u, 0u, 0u, LUMO_NONE, 0.0, LUMO_NONE, LUMO_NONE, LUMO_NONE, LUMO_NONE, 0ull};
__device__ __forceinline__ const Mat& vmat(const DevScene& S, const Vtx& v) { return v.h.material < 0 ? g_blank_material : S.materials[v.h.material]; }
__device__ __forceinline__ bool v_is_surface(const Vtx& v) { return v.h.material >= 0; }                    // vertex.rs:86-88 (Blank = camera)
__device__ __forceinline__ bool v_is_light(const Vtx& v) { return v.light >= 0; }
__device__ __forceinline__ bool v_is_delta(const DevScene& S, const Vtx& v, const Lam& l) { return v.delta != 0; }
__device__ __forceinline__ double sa_to_area(double pdf, D3 xo, D3 xi, D3 wi, D3 ngi) { return pdf * fabs(dot(wi, ngi)) / dist2(xo, xi); }   // measure.rs:9-11
__device__ __forceinline__ double v_shading_cosine(const DevScene& S, const Vtx& v, D3 wi) { return shading_cosine(vmat(S, v), wi, v.h.ns); }
__device__ __forceinline__ double v_shading_correction(const DevScene& S, const Vtx& v, D3 wi) {           // vertex.rs:118-126
    const Mat& m = vmat(S, v);
    return shading_cosine(m, wi, v.h.ng) * shading_cosine(m, v.wo, v.h.ns) / (shading_cosine(m, v.wo, v.h.ng) * shading_cosine(m, wi, v.h.ns));
}
__device__ __noinline__ C4 v_bsdf_f(const DevScene& S, const Vtx& v, D3 wi, const Lam& l, int mode) {
    const Mat& m = vmat(S, v);
    if (!mat_is_standard(m)) return c4(0.0);
    const Onb uvw = shading_onb(S, m, v.h);
    return bsdf_f<-1>(S, m, uvw, v.wo, wi, l, mode, v.h);
}
__device__ __forceinline__ C4 v_f(const DevScene& S, const Vtx& v, const Vtx& next, const Lam& l, int mode) { return v_bsdf_f(S, v, normalize(next.h.p - v.h.p), l, mode); }   // vertex.rs:129-132
__device__ __noinline__ double v_bsdf_pdf(const DevScene& S, const Vtx& v, D3 wi, const Lam& l, bool swap_dir) {
    const Mat& m = vmat(S, v);
    if (!mat_is_standard(m)) return 0.0;
    const Onb uvw = shading_onb(S, m, v.h);
    return bsdf_pdf<-1>(S, m, uvw, v.wo, wi, v.h, l, swap_dir);
}
__device__ __forceinline__ double v_pdf_prev(const DevScene& S, const Vtx& v, const Vtx& prev, D3 wi, const Lam& l) {   // vertex.rs:146-160
    if (v_is_delta(S, v, l) || v_is_delta(S, prev, l)) return 0.0;
    const double pdf_sa = v_bsdf_pdf(S, v, wi, l, true);
    const D3 ngp = !v_is_surface(prev) ? -v.wo : prev.h.ng;
    return sa_to_area(pdf_sa, v.h.p, prev.h.p, -v.wo, ngp);
}
__device__ __forceinline__ void v_camera(Vtx& v, D3 xo, double pdf_fwd, const C4& gathered) {               // vertex.rs:16-36
    v.h.t = 0.0; v.h.material = -1; v.h.backface = false /* (-X).X > 0 */; v.h.p = xo; v.h.fp_error = d3(0, 0, 0);
    v.h.ns = d3(1, 0, 0); v.h.ng = d3(1, 0, 0); v.h.u = 1.0; v.h.v = 0.0; wrap_uv(v.h.u, v.h.v);
    v.gathered = gathered; v.pdf_fwd = pdf_fwd; v.pdf_bck = 0.0; v.wo = d3(0, 0, 0); v.light = -1; v.delta = 0;
}

// ---- camera importance (camera.rs:167-388) -----------------------------------------------------------
__device__ __forceinline__ bool cam_bounds(const LumoCamera& C, double x, double y) { return x >= 0.0 && x < (double)C.res_x && y >= 0.0 && y < (double)C.res_y; }
__device__ __forceinline__ void cam_to_raster(const LumoCamera& C, D3 xl, double& x, double& y) {
    const D3 r = xf4_point(C.screen_to_raster_m, xf4_point(C.camera_to_screen_m, xl)); x = r.x; y = r.y;
}
__device__ __noinline__ bool cam_raster_xy(const LumoCamera& C, const Ray& ri, double& x, double& y) {
    if (C.ortho) { cam_to_raster(C, xf4_point(C.world_to_camera_m, ri.o), x, y); return cam_bounds(C, x, y); }
    const D3 wl = xf4_dir(C.world_to_camera_m, ri.d);
    const double ct = wl.z;
    if (ct <= 0.0) return false;
    const double fl = C.lens_radius == 0.0 ? 1.0 / ct : C.focal_length / ct;
    const D3 xl = xf4_point(C.world_to_camera_m, ri.o);
    cam_to_raster(C, xl + wl * fl, x, y);
    return cam_bounds(C, x, y);
}
__device__ __forceinline__ double cam_lens_area(const LumoCamera& C) { return C.lens_radius == 0.0 ? 1.0 : LUMO_PI * powi(C.lens_radius, 2); }
__device__ __noinline__ bool cam_sample_towards(const LumoCamera& C, D3 xi, double r0, double r1, Ray& ri) {
    double dx, dy; square_to_disk(r0, r1, dx, dy);
    if (C.ortho) {
        const D3 lens = C.lens_radius * d3(dx, dy, 0.0);
        const D3 xil = xf4_point(C.world_to_camera_m, xi);
        const D3 xol = xil * d3(1.0, 1.0, 0.0);
        const D3 xo = xf4_point(C.world_to_camera_inv, xol + lens);
        ri.o = xo; ri.d = normalize(normalize(xi - xo));
    } else {
        const D3 xol = C.lens_radius * d3(dx, dy, 0.0);
        const D3 xil = xf4_point(C.world_to_camera_m, xi);
        const D3 wil = normalize(xil - xol);
        ri.o = xf4_point(C.world_to_camera_inv, xol); ri.d = normalize(xf4_dir(C.world_to_camera_inv, wil));
    }
    double x, y;
    return cam_raster_xy(C, ri, x, y);
}
__device__ __forceinline__ double cam_pdf_xo(const LumoCamera& C, const Ray& ri) {
    double x, y;
    if (C.ortho) return cam_raster_xy(C, ri, x, y) ? 1.0 / C.image_plane_area : 0.0;
    const D3 xl = xf4_point(C.world_to_camera_m, ri.o);
    const double r2 = powi(C.lens_radius + LUMO_EPS, 2);
    return dist2(xl, d3(0, 0, 0)) < r2 ? 1.0 / cam_lens_area(C) : 0.0;
}
__device__ __forceinline__ double cam_pdf_wi(const LumoCamera& C, const Ray& ri) {
    double x, y;
    const D3 wl = xf4_dir(C.world_to_camera_m, ri.d);
    if (C.ortho) return (1.0 - wl.z < LUMO_EPS) ? 1.0 : 0.0;
    if (!cam_raster_xy(C, ri, x, y)) return 0.0;
    return 1.0 / (C.image_plane_area * powi(wl.z, 3));
}
__device__ __forceinline__ double cam_pdf_importance(const LumoCamera& C, const Ray& ri, D3 xi) {
    double x, y;
    if (!cam_raster_xy(C, ri, x, y)) return 0.0;
    const double* m = C.world_to_camera_m;   // to_normal_inv = transpose(m 3x3); times (0,0,1)
    const D3 ng = d3(m[0] * 0.0 + m[4] * 0.0 + m[8] * 1.0, m[1] * 0.0 + m[5] * 0.0 + m[9] * 1.0, m[2] * 0.0 + m[6] * 0.0 + m[10] * 1.0);
    const double pdf = dist2(xi, ri.o) / (fabs(dot(ng, ri.d)) * cam_lens_area(C));
    return fmax(pdf, 0.0);
}
__device__ __forceinline__ bool cam_sample_importance(const LumoCamera& C, const Ray& ri, C4& imp, double& x, double& y) {
    if (!cam_raster_xy(C, ri, x, y)) return false;
    if (C.ortho) { imp = (1.0 / C.image_plane_area) * c4(1.0); return true; }
    const D3 wl = xf4_dir(C.world_to_camera_m, ri.d);
    const double denom = C.image_plane_area * powi(wl.z, 4) * cam_lens_area(C);
    imp = (1.0 / denom) * c4(1.0);
    return true;
}

struct BdptCounters { unsigned long long closest, occlusion, overflow; };

// ---- MIS (mis.rs) ----------------------------------------------------------------------------------------
__device__ __forceinline__ void light_leaving_pdf(const DevScene& S, int light, const Ray& r, D3 ng, double& pdf_origin, double& pdf_dir) {   // object.rs:120-127
    pdf_origin = 1.0 / S.lights[light].area;
    pdf_dir = dot(ng, r.d) / LUMO_PI;
}
__device__ __noinline__ double pdf_light_leaving(const DevScene& S, const Vtx& curr, const Vtx& next, const Lam& l) {      // mis.rs:4-33
    if (v_is_delta(S, next, l)) return 0.0;
    if (curr.light < 0) return 0.0;
    const D3 xo = curr.h.p, xi = next.h.p;
    Ray ri; ri.o = xo; ri.d = normalize(xi - xo);
    const D3 wi = ri.d;
    double po, pdf_dir; light_leaving_pdf(S, curr.light, ri, curr.h.ng, po, pdf_dir);
    const D3 ngi = !v_is_surface(next) ? wi : next.h.ng;
    return sa_to_area(pdf_dir, xo, xi, wi, ngi);
}
__device__ __noinline__ double pdf_camera_leaving(const DevScene& S, const Vtx& curr, const Vtx& next, const Lam& l) {     // mis.rs:36-54
    if (v_is_delta(S, next, l)) return 0.0;
    const D3 xo = curr.h.p, xi = next.h.p;
    const D3 wi = normalize(xi - xo);
    Ray r; r.o = xo; r.d = normalize(wi);
    const double pdf_wi = cam_pdf_wi(S.P.camera, r);
    const D3 ngi = !v_is_surface(next) ? wi : next.h.ng;
    return sa_to_area(pdf_wi, xo, xi, wi, ngi);
}
__device__ __forceinline__ double pdf_light_origin(const DevScene& S, const Vtx& v) {                                       // mis.rs:57-64
    if (v.light < 0) return 0.0;
    return S.lights[v.light].pdf / S.lights[v.light].area;
}
__device__ __noinline__ double pdf_connection(const DevScene& S, const Vtx& curr, const Vtx& next, const Lam& l, const Vtx* prev) {   // mis.rs:69-96
    if (v_is_delta(S, next, l)) return 0.0;
    const D3 xo = curr.h.p, xi = next.h.p;
    double pdf_sa; D3 wi;
    if (prev) { const D3 wo = normalize(prev->h.p - xo); pdf_sa = v_bsdf_pdf(S, curr, wo, l, true); wi = curr.wo; }
    else { wi = normalize(xi - xo); pdf_sa = v_bsdf_pdf(S, curr, wi, l, false); }
    const D3 ngi = !v_is_surface(next) ? wi : next.h.ng;
    return sa_to_area(pdf_sa, xo, xi, wi, ngi);
}
__device__ __forceinline__ double map0(double p) { return p == 0.0 ? 1.0 : p; }
// mis.rs:103-239.  lp: light path (s vertices used), cp: camera path (t vertices used).  For s = 1 (NEE) and
// t = 1 (light tracing) the freshly sampled end vertex is passed as `ls1_override` / `ct1_override`.
__device__ __noinline__ double mis_weight(const DevScene& S, const Lam& l, const VtxArr& lp, int s, const VtxArr& cp, int t, const Vtx* ls1_override, const Vtx* ct1_override) {
    if (s + t == 2) return 1.0;
    const Vtx& ct1 = ct1_override ? *ct1_override : cp[t - 1];
    const Vtx& ls1 = s == 0 ? cp[0] : (ls1_override ? *ls1_override : lp[s - 1]);
    // The reference fills three arrays of s + t entries (pdf towards the light, pdf towards the camera, is-delta; mis.rs:139-141)
    // and sums over them twice.  Entry i is a pure function of i: vertex i of the light subpath below the connection, four
    // entries around it, vertex s + t - 1 - i of the camera subpath above — so the sums run over entry(i) directly, in the
    // reference's order, and no array bounds the path length.
    double pr_a = 0.0, pi_a = 0.0, pr_b = 0.0, pi_b = 0.0, pr_c = 0.0, pi_c = 0.0, pr_d = 0.0, pi_d = 0.0; bool dl_a = false, dl_d = false;
    if (s > 1) {
        const Vtx& ls2 = lp[s - 2];
        pr_a = pdf_connection(S, ls1, ls2, l, &ct1); pi_a = ls2.pdf_fwd; dl_a = v_is_delta(S, ls2, l);
    }
    if (s > 0) { pr_b = t == 1 ? pdf_camera_leaving(S, ct1, ls1, l) : pdf_connection(S, ct1, ls1, l, nullptr); pi_b = ls1.pdf_fwd; }
    if (t > 0) {
        const double pb = s == 0 ? pdf_light_origin(S, ct1) : (s == 1 ? pdf_light_leaving(S, ls1, ct1, l) : pdf_connection(S, ls1, ct1, l, nullptr));
        pr_c = ct1.pdf_fwd; pi_c = pb;
    }
    if (t > 1) {
        const Vtx& ct2 = cp[t - 2];
        const double pb = s == 0 ? pdf_light_leaving(S, ct1, ct2, l) : pdf_connection(S, ct1, ct2, l, &ls1);
        pr_d = ct2.pdf_fwd; pi_d = pb; dl_d = v_is_delta(S, ct2, l);
    }
    auto entry = [&](int i, double& pr, double& pi, bool& dl) {
        if (i < s - 2) { const Vtx& v = lp[i]; pr = v.pdf_bck; pi = v.pdf_fwd; dl = v_is_delta(S, v, l); }
        else if (i == s - 2) { pr = pr_a; pi = pi_a; dl = dl_a; }
        else if (i == s - 1) { pr = pr_b; pi = pi_b; dl = false; }
        else if (i == s) { pr = pr_c; pi = pi_c; dl = false; }
        else if (i == s + 1) { pr = pr_d; pi = pi_d; dl = dl_d; }
        else { const Vtx& v = cp[s + t - 1 - i]; pr = v.pdf_fwd; pi = v.pdf_bck; dl = v_is_delta(S, v, l); }
    };
    double sum_ri = 0.0, ri = 1.0;
    {   // towards the light: i = s - 1 .. 0, each needing the delta flag of entry i - 1 as well
        double pr = 0.0, pi = 0.0; bool dl = false;
        if (s > 0) entry(s - 1, pr, pi, dl);
        for (int i = s; i-- > 0;) {
            double prn = 0.0, pin = 0.0; bool dln = false;
            if (i > 0) entry(i - 1, prn, pin, dln);
            ri *= map0(pr) / map0(pi);
            if (!dl && !(i > 0 && dln)) sum_ri += ri * ri;
            pr = prn; pi = pin; dl = dln;
        }
    }
    ri = 1.0; sum_ri += ri;
    {   // towards the camera: i = s .. s + t - 2, each needing the delta flag of entry i + 1
        double pr = 0.0, pi = 0.0; bool dl = false;
        if (s + 1 < s + t) entry(s, pr, pi, dl);
        for (int i = s; i + 1 < s + t; i++) {
            double prn, pin; bool dln;
            entry(i + 1, prn, pin, dln);
            ri *= map0(pi) / map0(pr);
            if (!dl && !dln) sum_ri += ri * ri;
            pr = prn; pi = pin; dl = dln;
        }
    }
    return 1.0 / sum_ri;
}

// ---- connections (bd_path_trace.rs:77-290) ---------------------------------------------------------------
__device__ __noinline__ C4 add_camera_path(const DevScene& S, const Lam& lam, const VtxArr& cp, int t) {
    const Vtx& ct = cp[t - 1];
    if (!v_is_light(ct)) return c4(0.0);
    const C4 rad = ct.gathered * mat_emit(S, vmat(S, ct), lam, ct.h);
    if (is_black(rad)) return c4(0.0);
    return rad * mis_weight(S, lam, vtx_none(), 0, cp, t, nullptr, nullptr);
}
__device__ __noinline__ C4 connect_camera_path(const DevScene& S, Rng& rng, const Lam& lam, const VtxArr& cp, int t, BdptCounters& bc) {
    const Vtx& cl = cp[t - 1];
    if (v_is_delta(S, cl, lam) || v_is_light(cl)) return c4(0.0);
    const uint32_t li = sample_light(S, rng_float(rng));
    const double pdf_light = S.lights[li].pdf;
    const uint32_t lobj = S.P.n_objects + li;
    const LumoObject lo = S.objects[lobj];
    const DevHit& ho = cl.h;
    const D3 xo = ho.p;
    const double r0 = rng_float(rng), r1 = rng_float(rng);
    D3 wi = light_sample_towards(S, lo, xo, r0, r1);
    const double p_sct = v_bsdf_pdf(S, cl, wi, lam, false);
    if (p_sct == 0.0) return c4(0.0);
    const Ray ri = hit_generate_ray(ho, wi);
    DevHit hi;
    if (!light_hit(S, lobj, ri, hi)) return c4(0.0);
    bc.occlusion++;
    if (scene_occluded<false>(S, ri, hi.t - LUMO_EPS, nullptr)) return c4(0.0);
    const D3 xi = hi.p;
    const D3 ngi = !v_is_surface(cl) ? wi : hi.ng;
    const double p_lig = light_sample_towards_pdf(S, lo, ri, xi, ngi) * pdf_light;
    if (p_lig == 0.0) return c4(0.0);
    wi = ri.d;
    const double pdf_origin = sa_to_area(p_lig, xo, xi, wi, ngi);
    const C4 emittance = mat_emit(S, S.materials[hi.material], lam, hi);
    Vtx ll; ll.h = hi; ll.gathered = emittance; ll.light = (int)li; ll.pdf_fwd = pdf_origin; ll.pdf_bck = 0.0; ll.wo = d3(0, 0, 0); ll.delta = 0;   // Vertex::light
    const C4 bsdf = v_f(S, cl, ll, lam, 0);
    const double cos_wi = v_shading_cosine(S, cl, wi);
    const C4 radiance = cl.gathered * bsdf * emittance * c4(1.0) * cos_wi / p_lig;
    return radiance * mis_weight(S, lam, vtx_none(), 1, cp, t, &ll, nullptr);
}
// ---- the estimator: bd_path_trace::integrate (bd_path_trace.rs:23-75) as three kernels over a batch ------
//   k_bw_setup / k_bw_trace / k_bw_step   the two random walks of every sample as a wavefront (below): camera ray,
//                   wavelengths, light subpath, camera subpath; the two vertex arrays go to the batch's vertex
//                   buffer in HBM, together with the number of connection terms the sample has per class;
//   (exclusive scan of the term counts)
//   per term class, each with its own dense term index — emission (s = 0) and NEE (s = 1): k_bdpt_connect<CLASS>, one
//                   thread per term; light tracing (t = 1, a splat) and subpath connections (s, t >= 2): k_bq_prepare /
//                   k_bq_trace / k_bq_finish through a visibility-ray queue; contributions are added to the sample's
//                   radiance with f64 atomics.  As ONE kernel over all terms (39 k SASS instructions, every lane of a
//                   warp in another class) ncu showed 2.8 of 32 lanes active and 89 % of stall samples in instruction fetch;
//   k_bdpt_finish   one thread per sample: reference-style cost, tone map, film.
// Random numbers: the terms that draw (light tracing: 2 per non-delta light vertex; NEE: 3 per camera vertex that is
// neither delta nor a light) consume the sample's stream in the reference's order; a term finds its position by
// counting the drawing terms before it.
struct BdptBatch {
    uint32_t cap;                    // samples per batch
    Vtx* lp; Vtx* cp;                // [cap][LUMO_BDPT_MAXV]: the first LUMO_BDPT_MAXV vertices of every subpath
    Vtx* pool; uint32_t pool_blocks; uint32_t* pool_next;   // overflow blocks of LUMO_BDPT_MAXV vertices, handed out by an atomic counter
    uint32_t* ovf_l; uint32_t* ovf_c;   // [cap][LUMO_BDPT_OVF_BLOCKS]: the blocks a subpath owns
    int* ns; int* nt;                // subpath lengths
    double* lam;                     // [4][cap]
    double* rx; double* ry;
    double* radiance;                // [4][cap]
    uint32_t *pixel, *sample, *draws, *witem, *valid;
    // walk state of the sample's subpath in flight (k_bw_*): first the light subpath, then the camera subpath
    double* w_ray;                   // [6][cap] current ray
    double* w_cam;                   // [6][cap] the camera ray, drawn before the light walk (path_gen.rs draws in this order)
    double* w_gathered;              // [4][cap]
    double *w_pdf_fwd, *w_delta;     // solid-angle pdf of the last scatter; the tile's Russian-roulette threshold
    uint32_t *w_phase, *w_n, *w_depth;   // 0 light walk, 1 camera walk; vertices written; scatter depth
    double *w_ht, *w_hb0, *w_hb1, *w_hb2; uint32_t *w_hobj, *w_htri, *w_have;   // closest hit of the current ray (k_bw_trace)
    uint32_t* act[2];                // active sample lists (double-buffered)
    uint32_t* n_act;                 // [2]
    // visibility-ray queue of the light-tracing and connection classes (k_bq_*): a chunk of terms at a time
    uint32_t qcap;
    unsigned long long* q_term;      // b | s << 32 | t << 48
    double* q_ray;                   // [6][qcap]
    double *q_ht, *q_hb0, *q_hb1, *q_hb2; uint32_t *q_hobj, *q_htri, *q_have;   // closest hit (light tracing) / q_ht = first-found distance (connections)
    uint32_t* q_n;                   // [0] entries, [1] trace cursor
    unsigned long long* n_terms[3];  // per sample and class {light tracing, NEE, connection} (cap + 1 entries, the last one 0)
    unsigned long long* term_off[3]; // their exclusive scans, cap + 1 entries (the emission class has one term per sample)
};
enum BdptClass { BC_LIGHT_TRACE = 0, BC_NEE = 1, BC_CONNECT = 2, BC_EMISSION = 3 };
__device__ __forceinline__ VtxArr bdpt_lp(const BdptBatch& B, uint32_t b) { VtxArr a; a.base = B.lp + (size_t)b * LUMO_BDPT_MAXV; a.tab = B.ovf_l + (size_t)b * LUMO_BDPT_OVF_BLOCKS; a.pool = B.pool; return a; }
__device__ __forceinline__ VtxArr bdpt_cp(const BdptBatch& B, uint32_t b) { VtxArr a; a.base = B.cp + (size_t)b * LUMO_BDPT_MAXV; a.tab = B.ovf_c + (size_t)b * LUMO_BDPT_OVF_BLOCKS; a.pool = B.pool; return a; }


// ---- the two random walks of every sample as a wavefront ----------------------------------------------------
// One subpath per sample is in flight: the light subpath first, then the camera subpath (they share the sample's
// random stream and its wavelengths, which a dispersive scatter may terminate — path_gen.rs:21-50 then :4-19).
// Per bounce: k_bw_trace (Scene::hit of every live subpath's ray, the same traversal as the wave kernels) and
// k_bw_step (the body of path_gen::walk, path_gen.rs:53-157: vertex, BSDF sample, pdfs, throughput, roulette).
// As ONE kernel per sample doing both walks with their traversals inline, ncu showed 3.6 of 32 lanes active.
__device__ __forceinline__ void bw_append(const BdptBatch& B, uint32_t list, bool alive, uint32_t b) {
    const unsigned m = __ballot_sync(0xFFFFFFFFu, alive);
    if (!m) return;
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t base = 0;
    if (lane == (uint32_t)(__ffs(m) - 1)) base = atomicAdd(&B.n_act[list], (uint32_t)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, __ffs(m) - 1);
    if (alive) B.act[list][base + __popc(m & ((1u << lane) - 1u))] = b;
}
__device__ __forceinline__ void bw_store_ray(const BdptBatch& B, double* dst, uint32_t b, const Ray& r) {
    const size_t c = B.cap;
    dst[b] = r.o.x; dst[c + b] = r.o.y; dst[2 * c + b] = r.o.z; dst[3 * c + b] = r.d.x; dst[4 * c + b] = r.d.y; dst[5 * c + b] = r.d.z;
}
__device__ __forceinline__ Ray bw_load_ray(const BdptBatch& B, const double* src, uint32_t b) {
    const size_t c = B.cap;
    Ray r; r.o = d3(src[b], src[c + b], src[2 * c + b]); r.d = d3(src[3 * c + b], src[4 * c + b], src[5 * c + b]);
    return r;
}
__device__ __forceinline__ void bw_finish_sample(const BdptBatch& B, uint32_t b, int ns, int nt) {
    B.ns[b] = ns; B.nt[b] = nt;
    const unsigned long long L = ns >= 2 ? (unsigned long long)(ns - 1) : 0ull, C = nt >= 2 ? (unsigned long long)(nt - 1) : 0ull;
    B.n_terms[BC_LIGHT_TRACE][b] = L; B.n_terms[BC_NEE][b] = C; B.n_terms[BC_CONNECT][b] = L * C;
}

// per sample: pixel, camera ray, wavelengths, the light subpath's root and first ray (path_gen.rs:21-50)
__global__ void __launch_bounds__(128) k_bw_setup(const __grid_constant__ DevScene S, const __grid_constant__ Wave W, const __grid_constant__ WaveParams P, const __grid_constant__ BdptBatch B,
                                                  unsigned long long w0, uint32_t n) {
    const uint32_t Wd = S.P.camera.res_x, Hd = S.P.camera.res_y;
    const uint32_t n_pad = (n + 31u) & ~31u;
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < n_pad; b += gridDim.x * blockDim.x) {
        bool ok = b < n;
        if (ok) {
            const unsigned long long w = w0 + b;
            uint32_t px, py, sample;
            if (P.mode == WM_PILOT) {
                const uint32_t tile = (uint32_t)(w / LUMO_PILOT_N), k = (uint32_t)(w % LUMO_PILOT_N);
                const uint32_t x0 = (tile % P.tiles_x) * 16u, y0 = (tile / P.tiles_x) * 16u;
                px = min(x0 + 2u * (k % 8u), Wd - 1u); py = min(y0 + 2u * (k / 8u), Hd - 1u);
                sample = 0xFFFFFF00u + P.pilot_round;
            } else {
                const unsigned long long per_s = (unsigned long long)P.tiles_x * P.tiles_y * 256ull;
                const uint32_t si = (uint32_t)(w / per_s); const unsigned long long r = w % per_s;
                const uint32_t tile = (uint32_t)(r / 256ull), q = (uint32_t)(r % 256ull);
                const uint32_t blk = q / 32u, in = q % 32u;
                px = (tile % P.tiles_x) * 16u + (blk % 2u) * 8u + (in % 8u);
                py = (tile / P.tiles_x) * 16u + (blk / 2u) * 4u + (in / 8u);
                sample = P.spp_begin + si;
                ok = px < Wd && py < Hd;
            }
            B.valid[b] = ok ? 1u : 0u;
            B.witem[b] = (uint32_t)w;
            if (!ok) bw_finish_sample(B, b, 0, 0);
            else {
                const uint32_t pixel = px + py * Wd;
                Rng rng = rng_make(P.seed, pixel, sample, 0u, 0u);
                double jx, jy;
                if (P.mode == WM_PILOT) { jx = rng_float(rng); jy = rng_float(rng); } else raster_jitter(P, pixel, sample, rng, jx, jy);
                const double rx = (double)px + jx, ry = (double)py + jy;
                const double l0 = rng_float(rng), l1 = rng_float(rng);
                const Ray r = camera_generate_ray(S.P.camera, rx, ry, l0, l1);
                const Lam lam = lam_sample(rng_float(rng));
                const uint32_t li = sample_light(S, rng_float(rng));
                const double pdf_light = S.lights[li].pdf;
                const LumoObject lo = S.objects[S.P.n_objects + li];
                const double a0 = rng_float(rng), a1 = rng_float(rng), b0 = rng_float(rng), b1 = rng_float(rng);
                const DevHit ho = light_sample_on(S, lo, a0, a1);                                  // Sampleable::sample_leaving, object.rs:107-117
                const Onb uvw = onb_new(ho.ns);
                const Ray ri = hit_generate_ray(ho, to_world(uvw, square_to_cos_hemisphere(b0, b1)));
                double pdf_origin, pdf_dir; light_leaving_pdf(S, (int)li, ri, ho.ng, pdf_origin, pdf_dir);
                const C4 emit = mat_emit(S, S.materials[ho.material], lam, ho);
                Vtx& root = B.lp[(size_t)b * LUMO_BDPT_MAXV];
                root.h = ho; root.gathered = emit; root.light = (int)li; root.pdf_fwd = pdf_origin * pdf_light; root.pdf_bck = 0.0; root.wo = d3(0, 0, 0); root.delta = 0;
                const C4 gathered = emit * fabs(dot(ri.d, ho.ns)) / (pdf_light * pdf_origin * pdf_dir);
                bw_store_ray(B, B.w_ray, b, ri); bw_store_ray(B, B.w_cam, b, r);
                for (int k = 0; k < 4; k++) { B.w_gathered[(size_t)k * B.cap + b] = gathered.s[k]; B.lam[(size_t)k * B.cap + b] = lam.l[k]; B.radiance[(size_t)k * B.cap + b] = 0.0; }
                B.w_pdf_fwd[b] = pdf_dir; B.w_delta[b] = W.tile_delta[tile_of(P, S, pixel)];
                B.w_phase[b] = 0u; B.w_n[b] = 1u; B.w_depth[b] = 0u;
                B.rx[b] = rx; B.ry[b] = ry; B.pixel[b] = pixel; B.sample[b] = sample; B.draws[b] = rng.draws;
            }
        }
        bw_append(B, 0u, ok, b);
    }
}

// Scene::hit for the current ray of every live subpath.  Persistent warps pull 32 rays at a time from a cursor (n_act[2]):
// ray costs differ by orders of magnitude (mirror / glass meshes), a static split leaves long tails.
__global__ void __launch_bounds__(128, 8) k_bw_trace(const __grid_constant__ DevScene S, const __grid_constant__ Wave W, const __grid_constant__ BdptBatch B, uint32_t cur) {
    const uint32_t n = B.n_act[cur];
    if (blockIdx.x == 0 && threadIdx.x == 0) { B.n_act[cur ^ 1u] = 0u; if (n) atomicAdd(&W.run->closest, (unsigned long long)n); }
    const uint32_t lane = threadIdx.x & 31u;
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&B.n_act[2], 32u);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= n) break;
        const uint32_t i = base + lane;
        if (i < n) {
            const uint32_t b = B.act[cur][i];
            const Ray r = bw_load_ray(B, B.w_ray, b);
            HitRec h;
            if (scene_hit<false, LUMO_WAVE_KD_ROUND>(S, r, LUMO_INF, h, nullptr)) {
                B.w_ht[b] = h.t; B.w_hb0[b] = h.bary.x; B.w_hb1[b] = h.bary.y; B.w_hb2[b] = h.bary.z; B.w_hobj[b] = h.obj; B.w_htri[b] = h.tri; B.w_have[b] = 1u;
            } else B.w_have[b] = 0u;
        }
    }
}

// one iteration of path_gen::walk (path_gen.rs:53-157) for the live subpath of sample b, given the closest hit of its
// current ray; a light subpath that ends hands over to the sample's camera subpath (path_gen.rs:4-19).  Returns whether the
// sample still has a subpath in flight.
// the vertex about to be written (index n, a multiple of LUMO_BDPT_MAXV) needs a new overflow block
__device__ __forceinline__ bool bw_grow(const BdptBatch& B, uint32_t phase, uint32_t b, int n) {
    const int slot = n / LUMO_BDPT_MAXV - 1;
    if (slot >= (int)LUMO_BDPT_OVF_BLOCKS) return false;
    const uint32_t blk = atomicAdd(B.pool_next, 1u);
    if (blk >= B.pool_blocks) return false;
    (phase == 0u ? B.ovf_l : B.ovf_c)[(size_t)b * LUMO_BDPT_OVF_BLOCKS + slot] = blk;
    return true;
}
__device__ __forceinline__ bool bw_step_one(const DevScene& S, const WaveParams& P, const BdptBatch& B, uint32_t b, bool have, const HitRec& rec, unsigned long long& overflow) {
    const uint32_t phase = B.w_phase[b];
    const int mode = phase == 0u ? 1 : 0;
    const VtxArr vs = phase == 0u ? bdpt_lp(B, b) : bdpt_cp(B, b);
    int n = (int)B.w_n[b];
    uint32_t depth = B.w_depth[b];
    bool end = false;
    if (!have) end = true;
    else if (n >= LUMO_BDPT_MAXV && n % LUMO_BDPT_MAXV == 0 && !bw_grow(B, phase, b, n)) { overflow++; end = true; }   // the next vertex opens a new block and none is left
    else {
        const Ray ro = bw_load_ray(B, B.w_ray, b);
        Lam lam; C4 gathered;
        for (int k = 0; k < 4; k++) { lam.l[k] = B.lam[(size_t)k * B.cap + b]; gathered.s[k] = B.w_gathered[(size_t)k * B.cap + b]; }
        double pdf_fwd = B.w_pdf_fwd[b];
        Rng rng = rng_make(P.seed, B.pixel[b], B.sample[b], 0u, B.draws[b]);
        const DevHit ho = reconstruct_hit(S, ro, rec);
        const Mat& m = S.materials[ho.material];
        const uint32_t prev = depth;
        const D3 wo = -ro.d;
        Vtx& cv = vs[n];                                                                      // Vertex::surface, vertex.rs:51-84
        const bool is_delta = mat_is_delta(S, m, lam);
        cv.pdf_fwd = is_delta ? 0.0 : sa_to_area(pdf_fwd, vs[prev].h.p, ho.p, -wo, ho.ng);
        cv.h = ho; cv.gathered = gathered; cv.wo = wo; cv.pdf_bck = 0.0; cv.light = -1; cv.delta = is_delta ? 1 : 0;
        n++;
        depth += 1;
        const uint32_t curr = depth;
        const double ru = rng_float(rng), r0 = rng_float(rng), r1 = rng_float(rng);
        D3 wi;
        const Onb uvw = shading_onb(S, m, ho);
        if (!bsdf_sample<-1>(S, m, uvw, wo, ho, lam, ru, r0, r1, wi)) {
            if (mode == 1) n--;                                                               // path_gen.rs:97-99: a light path cannot end on a light
            else {                                                                            // Scene::get_light_at (bvh.rs:97-102)
                Ray rl; rl.o = hit_ray_origin(ho, true); rl.d = normalize(-ho.ng);
                RayCtx w; make_ctx(rl, w);
                const uint32_t li = S.P.n_lights ? tlas_hit<true, false>(S, S.P.lights_root, S.P.n_objects, w, 0.0, LUMO_INF, nullptr) : LUMO_NONE;
                vs[curr].light = li == LUMO_NONE ? -1 : (int)li;
            }
            end = true;
        } else {
            const Ray ri = hit_generate_ray(ho, wi);
            wi = ri.d;
            pdf_fwd = bsdf_pdf<-1>(S, m, uvw, wo, wi, ho, lam, false);
            if (pdf_fwd == 0.0) end = true;
            else {
                const double corr = mode == 0 ? 1.0 : v_shading_correction(S, vs[curr], wi);
                const C4 bsdf = bsdf_f<-1>(S, m, uvw, wo, wi, lam, mode, ho);
                gathered = gathered * (bsdf * v_shading_cosine(S, vs[curr], wi) * corr / pdf_fwd);
                vs[prev].pdf_bck = v_pdf_prev(S, vs[curr], vs[prev], wi, lam);
                if (depth >= LUMO_RR_DEPTH) {
                    const double lum = luminance(S, gathered, lam);
                    const double rr = fmin(lum / B.w_delta[b], 1.0);
                    if (rng_float(rng) > rr) end = true;
                    else if (depth >= LUMO_BDPT_MAX_DEPTH) end = true;
                    else gathered = gathered / rr;
                }
                if (!end) {
                    if (is_delta) pdf_fwd = 0.0;
                    bw_store_ray(B, B.w_ray, b, ri);
                    for (int k = 0; k < 4; k++) B.w_gathered[(size_t)k * B.cap + b] = gathered.s[k];
                    B.w_pdf_fwd[b] = pdf_fwd;
                }
            }
        }
        for (int k = 0; k < 4; k++) B.lam[(size_t)k * B.cap + b] = lam.l[k];      // a dispersive scatter may have terminated wavelengths
        B.draws[b] = rng.draws;
    }
    if (!end) { B.w_n[b] = (uint32_t)n; B.w_depth[b] = depth; return true; }
    if (phase == 0u) {                                                                        // light subpath done: start the camera subpath
        B.ns[b] = n;
        const Ray r = bw_load_ray(B, B.w_cam, b);
        const double pdf_wi = cam_pdf_wi(S.P.camera, r), pdf_xo = cam_pdf_xo(S.P.camera, r);
        v_camera(B.cp[(size_t)b * LUMO_BDPT_MAXV], r.o, pdf_xo, c4(1.0));
        bw_store_ray(B, B.w_ray, b, r);
        for (int k = 0; k < 4; k++) B.w_gathered[(size_t)k * B.cap + b] = 1.0;
        B.w_pdf_fwd[b] = pdf_wi; B.w_phase[b] = 1u; B.w_n[b] = 1u; B.w_depth[b] = 0u;
        return true;
    }
    bw_finish_sample(B, b, B.ns[b], n);
    return false;
}
__global__ void __launch_bounds__(128) k_bw_step(const __grid_constant__ DevScene S, const __grid_constant__ Wave W, const __grid_constant__ WaveParams P, const __grid_constant__ BdptBatch B, uint32_t cur) {
    const uint32_t n_act = B.n_act[cur];
    const uint32_t n_pad = (n_act + 31u) & ~31u;
    if (blockIdx.x == 0 && threadIdx.x == 0) B.n_act[2] = 0u;             // the trace kernel's work cursor, for the next bounce
    unsigned long long overflow = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x) {
        bool alive = false; uint32_t b = 0;
        if (i < n_act) {
            b = B.act[cur][i];
            HitRec rec; rec.t = B.w_ht[b]; rec.bary = d3(B.w_hb0[b], B.w_hb1[b], B.w_hb2[b]); rec.obj = B.w_hobj[b]; rec.tri = B.w_htri[b];
            alive = bw_step_one(S, P, B, b, B.w_have[b] != 0u, rec, overflow);
        }
        bw_append(B, cur ^ 1u, alive, b);
    }
    if (overflow) atomicAdd(&W.run->shadow_dropped, overflow);
}
// The stragglers: once only a few thousand subpaths are left (long specular chains), every further bounce of the wavefront costs
// a launch pair and one traversal's latency for next to no work.  One thread per remaining sample then runs its walks to the end.
__global__ void __launch_bounds__(64) k_bw_tail(const __grid_constant__ DevScene S, const __grid_constant__ Wave W, const __grid_constant__ WaveParams P, const __grid_constant__ BdptBatch B, uint32_t cur) {
    const uint32_t n_act = B.n_act[cur];
    unsigned long long overflow = 0, closest = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_act; i += gridDim.x * blockDim.x) {
        const uint32_t b = B.act[cur][i];
        for (;;) {
            const Ray r = bw_load_ray(B, B.w_ray, b);
            HitRec rec; rec.t = 0.0; rec.bary = d3(0, 0, 0); rec.obj = LUMO_NONE; rec.tri = 0u;
            closest++;
            const bool have = scene_hit<false>(S, r, LUMO_INF, rec, nullptr);
            if (!bw_step_one(S, P, B, b, have, rec, overflow)) break;
        }
    }
    if (closest) atomicAdd(&W.run->closest, closest);
    if (overflow) atomicAdd(&W.run->shadow_dropped, overflow);
}

// emission (s = 0) and next-event estimation (s = 1) terms: one thread per term of the class
template <int CLS>
__global__ void __launch_bounds__(128) k_bdpt_connect(const __grid_constant__ DevScene S, const __grid_constant__ Wave W, const __grid_constant__ WaveParams P, const __grid_constant__ BdptBatch B, uint32_t n) {
    static_assert(CLS == BC_EMISSION || CLS == BC_NEE, "light tracing and connections go through k_bq_*");
    BdptCounters bc = {0, 0, 0};
    const unsigned long long* off = B.term_off[BC_NEE];
    const unsigned long long total = CLS == BC_EMISSION ? (unsigned long long)n : off[n];
    const unsigned long long total_pad = (total + 31ull) & ~31ull;
    for (unsigned long long it = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; it < total_pad; it += (unsigned long long)gridDim.x * blockDim.x) {
        __syncwarp();                                                  // lanes whose term ended early must not run ahead into their next term:
        if (it >= total) continue;                                     // without this, 20 terms per thread left 2.5 of 32 lanes converged
        uint32_t b, j = 0;
        if (CLS == BC_EMISSION) b = (uint32_t)it;
        else {                                                         // sample that owns term `it`: last b with off[b] <= it
            uint32_t lo_ = 0, hi_ = n;
            while (hi_ - lo_ > 1u) { const uint32_t mid = (lo_ + hi_) >> 1; if (off[mid] <= it) lo_ = mid; else hi_ = mid; }
            b = lo_; j = (uint32_t)(it - off[b]);
        }
        const int ns = B.ns[b], nt = B.nt[b];
        if (CLS == BC_EMISSION && nt == 0 && ns == 0) continue;        // an invalid sample has no terms at all
        const VtxArr lp = bdpt_lp(B, b), cp = bdpt_cp(B, b);
        Lam lam; for (int k = 0; k < 4; k++) lam.l[k] = B.lam[(size_t)k * B.cap + b];
        C4 contrib = c4(0.0);
        if (CLS == BC_EMISSION) contrib = add_camera_path(S, lam, cp, nt);          // s = 0
        else {                                                         // NEE, t = j + 2 (bd_path_trace.rs:47-55)
            const int t = (int)j + 2;
            uint32_t drawing_l = 0, drawing_c = 0;                     // draws before this term, in the reference's order: 2 per non-delta light
            for (int q = 2; q <= ns; q++) if (!v_is_delta(S, lp[q - 1], lam)) drawing_l++;      // vertex (light tracing), 3 per earlier NEE term
            for (int q = 2; q < t; q++) if (!(v_is_delta(S, cp[q - 1], lam) || v_is_light(cp[q - 1]))) drawing_c++;
            Rng rs = rng_make(P.seed, B.pixel[b], B.sample[b], 0u, B.draws[b] + 2u * drawing_l + 3u * drawing_c);
            contrib = connect_camera_path(S, rs, lam, cp, t, bc);
        }
        if (!is_black(contrib)) for (int k = 0; k < 4; k++) if (contrib.s[k] != 0.0) atomicAdd(&B.radiance[(size_t)k * B.cap + b], contrib.s[k]);
    }
    if (bc.closest) atomicAdd(&W.run->closest, bc.closest);
    if (bc.occlusion) atomicAdd(&W.run->occlusion, bc.occlusion);
}

// ---- light tracing (t = 1) and (s, t >= 2) connections through a visibility-ray queue ---------------------------
// Both classes spend most of their time in one traversal per term, and most terms are rejected before or by it.  Run
// inside the term's thread that traversal kept 4.5-5.8 of 32 lanes busy.  So, per chunk of terms:
//   k_bq_prepare<CLASS>  the cheap rejections of connect_light_path / connect_paths + visible() up to the ray
//                        (bd_path_trace.rs:77-113, 230-290); survivors are compacted into the queue with their ray;
//   k_bq_trace<CLASS>    Scene::hit (light tracing) or Scene::hit_t (visible(), first-found semantics) for the queue;
//   k_bq_finish<CLASS>   the rest of the term for the queue entries: BSDFs, importance, MIS weight, splat / radiance.
__device__ __forceinline__ void bq_term_of(const BdptBatch& B, int cls, unsigned long long it, uint32_t n, uint32_t& b, int& s, int& t) {
    const unsigned long long* off = B.term_off[cls];
    uint32_t lo_ = 0, hi_ = n;
    while (hi_ - lo_ > 1u) { const uint32_t mid = (lo_ + hi_) >> 1; if (off[mid] <= it) lo_ = mid; else hi_ = mid; }
    b = lo_;
    const uint32_t j = (uint32_t)(it - off[b]);
    if (cls == BC_LIGHT_TRACE) { s = (int)j + 2; t = 1; }
    else { const int ns = B.ns[b]; const uint32_t L = (uint32_t)(ns - 1); t = (int)(j / L) + 2; s = (int)(j % L) + 2; }   // t outer, s inner (bd_path_trace.rs:57-66)
}
__device__ __forceinline__ Rng bq_light_trace_rng(const DevScene& S, const WaveParams& P, const BdptBatch& B, uint32_t b, const VtxArr& lp, int s, const Lam& lam) {
    uint32_t drawing = 0;                                               // 2 draws per non-delta light vertex before this one, in the reference's order
    for (int q = 2; q < s; q++) if (!v_is_delta(S, lp[q - 1], lam)) drawing++;
    return rng_make(P.seed, B.pixel[b], B.sample[b], 0u, B.draws[b] + 2u * drawing);
}
template <int CLS>
__global__ void __launch_bounds__(128) k_bq_prepare(const __grid_constant__ DevScene S, const __grid_constant__ Wave W, const __grid_constant__ WaveParams P, const __grid_constant__ BdptBatch B,
                                                    uint32_t n, unsigned long long t0, uint32_t count) {
    const uint32_t count_pad = (count + 31u) & ~31u, lane = threadIdx.x & 31u;
    unsigned long long traced = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count_pad; i += gridDim.x * blockDim.x) {
        bool keep = false; Ray ri; ri.o = d3(0, 0, 0); ri.d = d3(0, 0, 0); unsigned long long key = 0;
        if (i < count) {
            uint32_t b; int s, t;
            bq_term_of(B, CLS, t0 + i, n, b, s, t);
            const VtxArr lp = bdpt_lp(B, b), cp = bdpt_cp(B, b);
            Lam lam; for (int k = 0; k < 4; k++) lam.l[k] = B.lam[(size_t)k * B.cap + b];
            key = (unsigned long long)b | ((unsigned long long)s << 32) | ((unsigned long long)t << 48);
            const Vtx& ll = lp[s - 1];
            if (CLS == BC_LIGHT_TRACE) {                                // connect_light_path up to its Scene::hit (bd_path_trace.rs:77-113)
                if (!v_is_delta(S, ll, lam)) {
                    Rng rs = bq_light_trace_rng(S, P, B, b, lp, s, lam);
                    const double r0 = rng_float(rs), r1 = rng_float(rs);
                    if (cam_sample_towards(S.P.camera, ll.h.p, r0, r1, ri)) {
                        const double p_sct = v_bsdf_pdf(S, ll, -ri.d, lam, false);
                        const double p_imp = cam_pdf_importance(S.P.camera, ri, ll.h.p);
                        keep = !(p_sct == 0.0 || p_imp == 0.0);
                    }
                }
            } else {                                                    // connect_paths up to visible()'s Scene::hit_t (:230-290)
                const Vtx& cl = cp[t - 1];
                if (!(v_is_delta(S, cl, lam) || v_is_light(cl) || v_is_delta(S, ll, lam))) {
                    ri = hit_generate_ray(ll.h, cl.h.p - ll.h.p);
                    keep = !(dot(ri.d, ll.h.ng) < LUMO_EPS);
                }
            }
        }
        const unsigned m = __ballot_sync(0xFFFFFFFFu, keep);
        if (m) {
            uint32_t base = 0;
            if (lane == (uint32_t)(__ffs(m) - 1)) base = atomicAdd(&B.q_n[0], (uint32_t)__popc(m));
            base = __shfl_sync(0xFFFFFFFFu, base, __ffs(m) - 1);
            if (keep) {
                const uint32_t q = base + __popc(m & ((1u << lane) - 1u));
                B.q_term[q] = key;
                const size_t c = B.qcap;
                B.q_ray[q] = ri.o.x; B.q_ray[c + q] = ri.o.y; B.q_ray[2 * c + q] = ri.o.z; B.q_ray[3 * c + q] = ri.d.x; B.q_ray[4 * c + q] = ri.d.y; B.q_ray[5 * c + q] = ri.d.z;
                traced++;
            }
        }
    }
    if (traced) atomicAdd(CLS == BC_LIGHT_TRACE ? &W.run->closest : &W.run->occlusion, traced);
}
template <int CLS>
__global__ void __launch_bounds__(128, 8) k_bq_trace(const __grid_constant__ DevScene S, const __grid_constant__ BdptBatch B) {
    const uint32_t n = B.q_n[0], lane = threadIdx.x & 31u;
    const size_t c = B.qcap;
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&B.q_n[1], 32u);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= n) break;
        const uint32_t q = base + lane;
        if (q < n) {
            Ray r; r.o = d3(B.q_ray[q], B.q_ray[c + q], B.q_ray[2 * c + q]); r.d = d3(B.q_ray[3 * c + q], B.q_ray[4 * c + q], B.q_ray[5 * c + q]);
            if (CLS == BC_LIGHT_TRACE) {
                HitRec h;
                if (scene_hit<false, LUMO_WAVE_KD_ROUND>(S, r, LUMO_INF, h, nullptr)) {
                    B.q_ht[q] = h.t; B.q_hb0[q] = h.bary.x; B.q_hb1[q] = h.bary.y; B.q_hb2[q] = h.bary.z; B.q_hobj[q] = h.obj; B.q_htri[q] = h.tri; B.q_have[q] = 1u;
                } else B.q_have[q] = 0u;
            } else B.q_ht[q] = scene_hit_t<false>(S, r, nullptr);
        }
    }
}
template <int CLS>
__global__ void __launch_bounds__(128) k_bq_finish(const __grid_constant__ DevScene S, const __grid_constant__ Wave W, const __grid_constant__ WaveParams P, const __grid_constant__ BdptBatch B) {
    const uint32_t n = B.q_n[0];
    const uint32_t n_pad = (n + 31u) & ~31u;
    const size_t c = B.qcap;
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n_pad; q += gridDim.x * blockDim.x) {
        __syncwarp();
        if (q >= n) continue;
        const unsigned long long key = B.q_term[q];
        const uint32_t b = (uint32_t)key; const int s = (int)((key >> 32) & 0xFFFFu), t = (int)(key >> 48);
        const VtxArr lp = bdpt_lp(B, b), cp = bdpt_cp(B, b);
        Lam lam; for (int k = 0; k < 4; k++) lam.l[k] = B.lam[(size_t)k * B.cap + b];
        Ray ri; ri.o = d3(B.q_ray[q], B.q_ray[c + q], B.q_ray[2 * c + q]); ri.d = d3(B.q_ray[3 * c + q], B.q_ray[4 * c + q], B.q_ray[5 * c + q]);
        const Vtx& ll = lp[s - 1];
        if (CLS == BC_LIGHT_TRACE) {                                    // the rest of connect_light_path (bd_path_trace.rs:96-113)
            if (!B.q_have[q]) continue;
            HitRec rec; rec.t = B.q_ht[q]; rec.bary = d3(B.q_hb0[q], B.q_hb1[q], B.q_hb2[q]); rec.obj = B.q_hobj[q]; rec.tri = B.q_htri[q];
            const D3 xi = ll.h.p, xo = ri.o, wi = ri.d;
            const DevHit hh = reconstruct_hit(S, ri, rec);
            if (max_element(vabs(hh.p - xi)) > sqrt(LUMO_EPS)) continue;
            C4 color; double sx, sy;
            if (!cam_sample_importance(S.P.camera, ri, color, sx, sy)) continue;
            if (is_black(color)) continue;
            const double p_imp = cam_pdf_importance(S.P.camera, ri, xi);
            color = color / p_imp;
            Vtx cl; v_camera(cl, xo, cam_pdf_xo(S.P.camera, ri), color / p_imp);
            color = color * (ll.gathered * c4(1.0) * v_shading_cosine(S, ll, -wi) * v_shading_correction(S, ll, -wi)
                             * v_f(S, ll, cl, lam, 1) * mis_weight(S, lam, lp, s, vtx_none(), 1, nullptr, &cl));
            if (P.mode == WM_MAIN) film_add_sample(S, W.pixels, W.splats, tone_map(S, (int)P.tone_map, P.tone_map_arg, color, lam), lam, sx, sy, true);
        } else {                                                        // visible()'s comparison and the rest of connect_paths (:243-255, :286-289)
            const Vtx& cl = cp[t - 1];
            const D3 xc = cl.h.p, xl = ll.h.p;
            if (!(fabs(sqrt(fmax(dist2(xl, xc), 0.0)) - B.q_ht[q]) < LUMO_EPS)) continue;
            const D3 wi = normalize(xl - xc);
            const double p_sct = v_bsdf_pdf(S, cl, wi, lam, false) * v_bsdf_pdf(S, ll, -wi, lam, false);
            if (p_sct == 0.0) continue;
            const C4 lb = v_f(S, ll, cl, lam, 1), cb = v_f(S, cl, ll, lam, 0);
            const C4 radiance = ll.gathered * lb * v_shading_cosine(S, ll, -wi) * cl.gathered * cb * v_shading_cosine(S, cl, wi) * c4(1.0) / dist2(xc, xl);
            if (is_black(radiance)) continue;
            const C4 contrib = radiance * mis_weight(S, lam, lp, s, cp, t, nullptr, nullptr);
            if (!is_black(contrib)) for (int k = 0; k < 4; k++) if (contrib.s[k] != 0.0) atomicAdd(&B.radiance[(size_t)k * B.cap + b], contrib.s[k]);
        }
    }
}

__global__ void __launch_bounds__(256) k_bdpt_finish(const __grid_constant__ DevScene S, const __grid_constant__ Wave W, const __grid_constant__ WaveParams P, const __grid_constant__ BdptBatch B, uint32_t n) {
    unsigned long long cost_sum = 0, paths = 0; uint32_t depth_max = 0;     // run counters: per thread here, one atomic per warp at the end
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < n; b += gridDim.x * blockDim.x) {
        if (!B.valid[b]) continue;
        const int ns = B.ns[b], nt = B.nt[b];
        const VtxArr lp = bdpt_lp(B, b), cp = bdpt_cp(B, b);
        Lam lam; C4 radiance;
        for (int k = 0; k < 4; k++) { lam.l[k] = B.lam[(size_t)k * B.cap + b]; radiance.s[k] = B.radiance[(size_t)k * B.cap + b]; }
        unsigned long long cost = (unsigned long long)(ns + nt);                                 // bd_path_trace.rs:31-66
        for (int s = 2; s <= ns; s++) if (!v_is_delta(S, lp[s - 1], lam)) cost += 1;
        for (int t = 2; t <= nt; t++) if (!v_is_delta(S, cp[t - 1], lam) && v_is_light(cp[t - 1])) cost += 1;
        if (nt >= 2 && ns >= 2) cost += (unsigned long long)(nt - 1) * (unsigned long long)(ns - 1);
        if (P.mode == WM_PILOT) { const uint32_t w = B.witem[b]; W.pilot_lum[w] = luminance(S, radiance, lam); W.pilot_cost[w] = (uint32_t)cost; }
        else {
            bool finite = true;
            for (int k = 0; k < 4; k++) finite = finite && isfinite(radiance.s[k]);
            if (!finite) atomicAdd(&W.run->nonfinite, 1ull);
            film_add_sample(S, W.pixels, W.splats, tone_map(S, (int)P.tone_map, P.tone_map_arg, radiance, lam), lam, B.rx[b], B.ry[b], false);
            paths += 1ull; cost_sum += cost; depth_max = max(depth_max, (uint32_t)(ns + nt));
        }
    }
    __syncwarp();
    for (int o = 16; o > 0; o >>= 1) {
        cost_sum += __shfl_down_sync(0xFFFFFFFFu, cost_sum, o); paths += __shfl_down_sync(0xFFFFFFFFu, paths, o);
        depth_max = max(depth_max, __shfl_down_sync(0xFFFFFFFFu, depth_max, o));
    }
    if ((threadIdx.x & 31u) == 0u && paths) { atomicAdd(&W.run->cost, cost_sum); atomicAdd(&W.run->camera_paths, paths); atomicMax(&W.run->max_depth, depth_max); }
}

}  // namespace lumo_dev
