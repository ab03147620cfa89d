//! `extern "C"` declarations of include/lumo_gpu.h, one for one, plus the host builder's three entry points.
//! What each call replaces in lumo is documented in the header; INTEGRATION.md shows `Renderer::render` on top of it.
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_void};

#[repr(C)] pub struct lumo_ctx { _private: [u8; 0] }
#[repr(C)] pub struct lumo_scene { _private: [u8; 0] }

pub const LUMO_OK: i32 = 0;
pub const LUMO_ERR_INVALID: i32 = -1;
pub const LUMO_ERR_CUDA: i32 = -2;
pub const LUMO_ERR_OOM: i32 = -3;
pub const LUMO_ERR_UNSUPPORTED: i32 = -4;

/// Integrator::{PathTrace, DirectLight, BDPathTrace} (src/tracer/integrator.rs:17-27)
pub const LUMO_PATH_TRACE: i32 = 0;
pub const LUMO_DIRECT_LIGHT: i32 = 1;
pub const LUMO_BD_PATH_TRACE: i32 = 2;
/// SamplerType (src/samplers.rs:8-21)
pub const LUMO_SAMPLER_UNIFORM: i32 = 0;
pub const LUMO_SAMPLER_JITTERED: i32 = 1;
pub const LUMO_SAMPLER_MULTI_JITTERED: i32 = 2;
pub const LUMO_SAMPLER_SOBOL: i32 = 3;
/// ToneMap (src/tone_mapping.rs:13-20)
pub const LUMO_TONE_NONE: i32 = 0;
pub const LUMO_TONE_CLAMP: i32 = 1;
pub const LUMO_TONE_REINHARD: i32 = 2;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct lumo_render_params {
    pub integrator: i32,
    pub sampler: i32,
    pub tone_map: i32,
    pub flags: i32,
    pub tone_map_arg: f64,
    pub rr_delta: f64,
    pub seed: u64,
    pub spp_begin: u32,
    pub spp_end: u32,
    pub total_spp: u32,
    pub wave_paths: u32,
}

#[repr(C)]
pub struct lumo_film_accum {
    pub pixels: *mut f64,
    pub splats: *mut f64,
    pub counters: [u64; 10],
    pub tile_deltas: *mut f64,
    pub device_ms: f64,
}

extern "C" {
    pub fn lumo_gpu_device_count(n: *mut i32) -> i32;
    pub fn lumo_gpu_ctx_create(device: i32, ctx: *mut *mut lumo_ctx) -> i32;
    pub fn lumo_gpu_ctx_destroy(ctx: *mut lumo_ctx) -> i32;
    pub fn lumo_gpu_ctx_set_stream(ctx: *mut lumo_ctx, cuda_stream: *mut c_void) -> i32;
    pub fn lumo_gpu_scene_upload(ctx: *mut lumo_ctx, blob: *const c_void, len: u64, scene: *mut *mut lumo_scene) -> i32;
    pub fn lumo_gpu_scene_destroy(scene: *mut lumo_scene) -> i32;
    pub fn lumo_gpu_trace_closest(scene: *mut lumo_scene, origin_xyz: *const f64, dir_xyz: *const f64, t_max: *const f64, n: u64,
                                  obj_id: *mut u32, tri_id: *mut u32, t: *mut f64, bary_uv: *mut f64) -> i32;
    pub fn lumo_gpu_trace_any(scene: *mut lumo_scene, origin_xyz: *const f64, dir_xyz: *const f64, t_max: *const f64, n: u64, occluded: *mut u8) -> i32;
    pub fn lumo_gpu_trace_first_found(scene: *mut lumo_scene, origin_xyz: *const f64, dir_xyz: *const f64, n: u64, t: *mut f64) -> i32;
    pub fn lumo_gpu_render(scene: *mut lumo_scene, params: *const lumo_render_params, out: *mut lumo_film_accum) -> i32;
    pub fn lumo_gpu_render_dev(scene: *mut lumo_scene, params: *const lumo_render_params, pixels_dev: *mut f64, splats_dev: *mut f64,
                               counters10: *mut u64, device_ms: *mut f64) -> i32;
    pub fn lumo_gpu_render_multi(scenes: *mut *mut lumo_scene, n: i32, params: *const lumo_render_params, out: *mut lumo_film_accum) -> i32;
    pub fn lumo_gpu_sample_range(g: i32, n: i32, begin: u32, end: u32, g_begin: *mut u32, g_end: *mut u32) -> i32;
    pub fn lumo_gpu_film_encode_dev(ctx: *mut lumo_ctx, pixels_dev: *const f64, splats_dev: *const f64, n_pixels: u64, splat_scale: f64,
                                    filter_integral: f64, transfer: i32, rgb8: *mut u8, kernel_ms: *mut f32) -> i32;
    pub fn lumo_gpu_film_encode(ctx: *mut lumo_ctx, pixels: *const f64, splats: *const f64, n_pixels: u64, splat_scale: f64,
                                filter_integral: f64, transfer: i32, rgb8: *mut u8) -> i32;
    pub fn lumo_gpu_ctx_count_visits(ctx: *mut lumo_ctx, enable: i32) -> i32;
    pub fn lumo_gpu_ctx_visits(ctx: *mut lumo_ctx, out12: *mut u64) -> i32;
    pub fn lumo_gpu_ctx_closest_mode(ctx: *mut lumo_ctx, mode: i32) -> i32;
    pub fn lumo_gpu_ctx_closest_stats(ctx: *mut lumo_ctx, stats14: *mut u64) -> i32;
    pub fn lumo_gpu_ctx_shade_stats(ctx: *mut lumo_ctx, stats4: *mut u64) -> i32;
    pub fn lumo_gpu_ctx_occlusion_mode(ctx: *mut lumo_ctx, mode: i32) -> i32;
    pub fn lumo_gpu_ctx_occlusion_stats(ctx: *mut lumo_ctx, stats9: *mut u64) -> i32;
    pub fn lumo_gpu_ctx_kernel_times(ctx: *mut lumo_ctx, ms4: *mut f64, launches4: *mut u64) -> i32;
    pub fn lumo_gpu_ctx_iter_log(ctx: *mut lumo_ctx, out: *mut u32, cap: u32, n: *mut u32) -> i32;
    pub fn lumo_gpu_trace_closest_dev(scene: *mut lumo_scene, origin_dev: *const f64, dir_dev: *const f64, n: u64,
                                      obj_dev: *mut u32, tri_dev: *mut u32, t_dev: *mut f64, bary_dev: *mut f64, kernel_ms: *mut f32) -> i32;
    pub fn lumo_gpu_fp64_peak(ctx: *mut lumo_ctx, tflops: *mut f64, ms: *mut f64) -> i32;
    pub fn lumo_gpu_math_eval(ctx: *mut lumo_ctx, fn_: i32, x: *const f64, y: *const f64, n: u64, out: *mut f64) -> i32;
    pub fn lumo_gpu_last_error() -> *const c_char;

    // liblumo_host: scene program (the serialised Scene / Camera construction calls) -> device blob
    pub fn lumo_host_build(program: *const c_void, len: u64, blob: *mut *mut c_void, blob_len: *mut u64) -> i32;
    pub fn lumo_host_free(p: *mut c_void);
    pub fn lumo_host_last_error() -> *const c_char;
}

/// Panics with the library's message, like the reference's own `unwrap`s (errors are panics in lumo).
pub fn check(rc: i32) {
    if rc != LUMO_OK {
        let msg = unsafe { core::ffi::CStr::from_ptr(lumo_gpu_last_error()) }.to_string_lossy().into_owned();
        panic!("liblumo_gpu: {msg} ({rc})");
    }
}
