// vpt_internal.h -- structures shared by the host side (vpt_api.cpp) and the device side (*.cu) of libvpt_b200.
// Not part of the public boundary (that is include/vpt.h).
#pragma once
#include <stdint.h>
#include "../../include/vpt.h"

namespace vpt {

constexpr int kMaxSpheres = VPT_MAX_SPHERES;
constexpr int kMaxEmitters = VPT_MAX_EMITTERS;
constexpr int kThreadsPerBlock = 128; // one pixel tile = kTile consecutive storage-order pixels
constexpr int kTile = 128;

// ---- stages of a path vertex (the queues of the wavefront kernels) -----------------------------------------------
enum : int { SQ_PRIMARY = 0, SQ_MED_POINT, SQ_MED_AREA, SQ_SURF_P, SQ_SURF_L, SQ_SURF_F, SQ_COUNT };
constexpr int kDestFree = SQ_COUNT; // the path ended: its record is free

// ---- fp32 scene ----------------------------------------------------------------------------------------------
// Ray/sphere record for the scan loop (lives in kernel-parameter constant memory; uniform index -> broadcast).
// Re-anchored form (DESIGN.md "fp32 geometry"): for a sphere with centre p and radius r the
// host picks, in double, an anchor q and m = q - p so that for a ray origin o
//     c = |o-p|^2 - r^2 = |o-q|^2 + 2 (o-q).m + c0 ,  c0 = |m|^2 - r^2
// has no 1e10-sized cancellation: huge spheres (the r = 1e5 walls) use q = surface point nearest the scene (c0 ~ 0),
// ordinary spheres use q = p (m = 0, c0 = -r^2).
struct GeomF {
    float qx, qy, qz;
    float mx, my, mz;
    float c0;
    float r2;     // ordinary spheres: det = r^2 - |op - (op.d) d|^2
    int32_t id;   // index into the caller's sphere array
    int32_t big;  // 1: det = b^2 - c
};
// Shading record, indexed by the caller's sphere index (divergent index -> staged in shared memory).
struct MatF {
    float px, py, pz, r;
    float cr, cg, cb;       // albedo
    float lr, lg, lb;       // radiance
    float eta[3], kappa[3];
    float alpha;
    int32_t material;
    int32_t emits;          // any radiance channel > 0 (vptShadeMethods.h:1296)
    int32_t pad;
};
struct SceneF {
    int32_t n_spheres, n_geom, n_emitters, n_area;
    int32_t n_big, pad0, pad1, pad2; // geom[0 .. n_big) are the re-anchored huge spheres, the rest ordinary ones
    int32_t emitters[kMaxEmitters]; // spheres with any radiance channel > 0, in index order
    int32_t area[kMaxEmitters];     // spheres with r > 0 && radiance.x > 0 (misSamplingFunctions.h:106), in index order
    GeomF geom[kMaxSpheres];
    MatF mat[kMaxSpheres];
};

// ---- fp64 scene (REF mode: the reference's own representation and operation order) ----------------------------------
struct SphereD {
    double r, px, py, pz;
    double cr, cg, cb;
    double lr, lg, lb;
    double eta[3], kappa[3];
    double alpha;
    int32_t material;
    int32_t emits;
};
struct SceneD {
    int32_t n_spheres, n_emitters;
    int32_t emitters[kMaxEmitters];
    SphereD s[kMaxSpheres];
};

// ---- per-launch parameters ----------------------------------------------------------------------------------------
struct LaunchParams {
    int32_t width, height, n_pixels;
    int32_t sample_begin, sample_end;
    int32_t tile_rank, tile_count, n_tiles_total;
    int32_t method, max_depth;
    uint32_t key0, key1;
    uint32_t quirks;
    double out_scale; // 1 (SUM) or 1/spp (MEAN)
    // medium / roulette, both precisions
    double sigma_a, sigma_s, continue_prob;
    // camera (rt.cpp:755-759), prepared on the host in double
    double cam_o[3], cam_d[3], cam_cx[3], cam_cy[3];
    // VPT_METHOD_RAYMARCH
    double march_step;
    int32_t march_source, pad_;
};

// fp32 constants derived from LaunchParams on the host (double arithmetic), so that no kernel converts doubles in its hot loop
struct ConstsF {
    float sigma_t, inv_sigma_t, sigma_s, albedo_over_cp, inv_cp, q;
    float n_emitters; // 1 / probSource
    int32_t method, max_depth;
    float cam_o[3], cam_d[3], cam_cx[3], cam_cy[3], inv_w, inv_h;
    float march_step; int32_t march_source;
};

constexpr int kDebugCounters = 32;
struct Counters { // device-side, accumulated with atomics at thread exit
    unsigned long long events, scans, nonfinite, paths;
    unsigned long long dbg[kDebugCounters]; // only written by builds with -DVPT_SMWAVE_PROFILE (tools/smwave_timing.py); see vpt_smwave.cuh
};

// entry points implemented in the .cu files, called from vpt_api.cpp
// multi-kernel HBM wavefront (vpt_kernels_hbm.cu); synchronises the stream internally (it polls a device flag between batches of rounds)
int launch_hbmwave_f32(const SceneF &scene, const LaunchParams &lp, const ConstsF &cf, float *hdr_dev, Counters *counters_dev, void *stream, int n_owned_tiles, uint64_t *launches);
// stream-ordered allocation from the library's per-device scratch pool (vpt_api.cpp); free with cudaFreeAsync
int scratch_alloc_(int device, void **ptr, size_t bytes, void *stream);
void scratch_trim_(int device, size_t keep_bytes); // give the pool's unused memory above keep_bytes back to the driver
int launch_march_f32(const SceneF &scene, const LaunchParams &lp, const ConstsF &cf, float *hdr_dev, Counters *counters_dev, void *stream, int n_blocks);
int launch_render_f32(const SceneF &scene, const LaunchParams &lp, const ConstsF &cf, float *hdr_dev, Counters *counters_dev, void *stream, int n_blocks, int kernel);
int launch_render_f64(const SceneD &scene, const LaunchParams &lp, float *hdr_dev, Counters *counters_dev, void *stream, int n_blocks, int kernel);
int launch_unit_f32(int fn, const SceneF &scene, const LaunchParams &lp, const ConstsF &cf, int n, const double *in_dev, int in_stride, double *out_dev, int out_stride, void *stream);
int launch_unit_f64(int fn, const SceneD &scene, const LaunchParams &lp, int n, const double *in_dev, int in_stride, double *out_dev, int out_stride, void *stream);
int launch_philox(int n, const uint32_t *ctr_dev, const uint32_t *key_dev, uint32_t *out_dev, void *stream);
int launch_fma_peak(float *sink_dev, int n_blocks, int n_threads, int iters, void *stream);
constexpr int kFmaPeakFlopsPerThreadIter = 2 * 16 * 8; // see vpt_kernels_f32.cu fma_peak_kernel

} // namespace vpt
