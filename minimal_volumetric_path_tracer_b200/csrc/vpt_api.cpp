// vpt_api.cpp -- host side of libvpt_b200: the extern "C" boundary declared in include/vpt.h.
// Validates arguments, prepares the device scene (in double: re-anchoring of huge spheres for the fp32 scan, camera basis
// in the reference's operation order rt.cpp:755-759), launches the kernels and moves results.  No CPU compute path exists:
// if CUDA is unavailable the entry points fail.
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "vpt_internal.h"

using namespace vpt;

namespace {

thread_local std::string g_last_cuda_error;
thread_local unsigned long long g_last_debug[kDebugCounters];

int cuda_fail(cudaError_t e, const char *what) {
    g_last_cuda_error = std::string(what) + ": " + cudaGetErrorString(e);
    return e == cudaErrorNoDevice || e == cudaErrorInvalidDevice || e == cudaErrorInsufficientDriver ? VPT_ERR_NO_DEVICE : VPT_ERR_CUDA;
}
#define CUDA_TRY(expr)                                   \
    do {                                                 \
        cudaError_t e_ = (expr);                         \
        if (e_ != cudaSuccess) return cuda_fail(e_, #expr); \
    } while (0)

struct V3 { double x, y, z; };
V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
V3 mul(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
V3 unit(V3 a) { return mul(a, 1.0 / std::sqrt(dot(a, a))); }
V3 v3(const double *p) { return {p[0], p[1], p[2]}; }

bool finite3(const double *p) { return std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]); }
bool emits(const vpt_sphere &s) { return s.radiance[0] > 0 || s.radiance[1] > 0 || s.radiance[2] > 0; } // vptShadeMethods.h:1296

int validate_scene(const vpt_sphere *s, int n, int method = VPT_METHOD_FREE_FLIGHT) {
    if (!s) return VPT_ERR_INVALID_ARGUMENT;
    if (n <= 0 || n > kMaxSpheres) return VPT_ERR_SCENE;
    int n_emit = 0;
    for (int i = 0; i < n; ++i) {
        if (!(s[i].r >= 0) || !std::isfinite(s[i].r) || !finite3(s[i].p) || !finite3(s[i].c) || !finite3(s[i].radiance)) return VPT_ERR_SCENE;
        if (s[i].material == 3 && method != VPT_METHOD_VOLUME_SPHERES) return VPT_ERR_UNSUPPORTED; // volumetric spheres: bdsf leaves pdf and direction unset for them in the active methods
        if (s[i].material < 0 || s[i].material > 3) return VPT_ERR_SCENE;
        if (s[i].material == 1 && (!(s[i].alpha > 0) || !finite3(s[i].eta) || !finite3(s[i].kappa))) return VPT_ERR_SCENE;
        if (emits(s[i])) ++n_emit;
    }
    if (n_emit > kMaxEmitters) return VPT_ERR_SCENE;
    return VPT_OK;
}

int validate_params(const vpt_params *p, bool need_image) {
    if (!p) return VPT_ERR_INVALID_ARGUMENT;
    if (need_image) {
        if (p->width <= 0 || p->height <= 0 || p->spp <= 0) return VPT_ERR_INVALID_ARGUMENT;
        if ((long long)p->width * p->height > 0x7fffffffLL / 4) return VPT_ERR_INVALID_ARGUMENT;
        const bool whole = p->sample_begin == 0 && p->sample_end == 0;
        if (!whole && (p->sample_begin < 0 || p->sample_end <= p->sample_begin || p->sample_end > p->spp)) return VPT_ERR_INVALID_ARGUMENT;
        if ((whole ? p->spp : p->sample_end - p->sample_begin) >= VPT_MAX_SAMPLES_PER_CALL) return VPT_ERR_INVALID_ARGUMENT; // 32-bit sample counters per work item
        const bool all_tiles = p->tile_rank == 0 && p->tile_count == 0;
        if (!all_tiles && (p->tile_count <= 0 || p->tile_rank < 0 || p->tile_rank >= p->tile_count)) return VPT_ERR_INVALID_ARGUMENT;
        if (p->output != VPT_OUTPUT_SUM && p->output != VPT_OUTPUT_MEAN) return VPT_ERR_INVALID_ARGUMENT;
    }
    if (p->method < 0 || p->method > VPT_METHOD_VOLUME_SPHERES) return VPT_ERR_INVALID_ARGUMENT;
    if (p->method == VPT_METHOD_VOLUME_SPHERES && p->precision != VPT_PRECISION_FP64_REF) return VPT_ERR_UNSUPPORTED; // the legacy estimator exists in reference precision only
    if (p->method == VPT_METHOD_RAYMARCH && (!(p->march_step > 0) || !std::isfinite(p->march_step) || p->march_source < 0 || p->march_source >= kMaxSpheres)) return VPT_ERR_INVALID_ARGUMENT;
    if (p->precision != VPT_PRECISION_FP32 && p->precision != VPT_PRECISION_FP64_REF) return VPT_ERR_INVALID_ARGUMENT;
    if (!(p->sigma_a >= 0) || !(p->sigma_s >= 0) || !(p->sigma_a + p->sigma_s > 0) || !std::isfinite(p->sigma_a + p->sigma_s)) return VPT_ERR_INVALID_ARGUMENT;
    if (!(p->continue_prob > 0) || !(p->continue_prob <= 1)) return VPT_ERR_INVALID_ARGUMENT;
    // Path length: the reference bounds it by roulette alone (max_depth <= 0).  A roulette that (almost) never fires would let a path in a scene
    // without emitter geometry run forever -- a GPU hang -- so unlimited depth needs continue_prob <= 0.99 (the kernels' internal cap of
    // VPT_MAX_DEPTH bounces is then reached with probability < 1e-17), anything above needs an explicit 0 < max_depth <= VPT_MAX_DEPTH.
    if (p->max_depth > VPT_MAX_DEPTH) return VPT_ERR_INVALID_ARGUMENT;
    if (p->max_depth <= 0 && p->continue_prob > 0.99) return VPT_ERR_INVALID_ARGUMENT;
    if (!finite3(p->cam_o) || !finite3(p->cam_dir) || !(p->fov > 0) || dot(v3(p->cam_dir), v3(p->cam_dir)) == 0) return VPT_ERR_INVALID_ARGUMENT;
    if (p->quirks & ~(uint32_t)VPT_QUIRKS_REFERENCE) return VPT_ERR_INVALID_ARGUMENT;
    if (p->precision == VPT_PRECISION_FP32 && p->quirks != 0) return VPT_ERR_UNSUPPORTED; // rounding-decided behaviours exist in FP64 only
    if (p->kernel < VPT_KERNEL_AUTO || p->kernel > VPT_KERNEL_WAVEFRONT_HBM) return VPT_ERR_INVALID_ARGUMENT;
    if (p->kernel == VPT_KERNEL_WAVEFRONT || p->kernel == VPT_KERNEL_MEGA_SCAN) return VPT_ERR_UNSUPPORTED; // superseded variants, no longer built (profiles/r1_summary.md has their measurements)
    if (p->kernel == VPT_KERNEL_WAVEFRONT_HBM && p->precision != VPT_PRECISION_FP32) return VPT_ERR_UNSUPPORTED; // the multi-kernel wavefront is FP32 only
    return VPT_OK;
}

// fp32 scene: scan records for r > 0 spheres (r == 0 spheres are never ray-intersected in fp32 semantics), shading records for all.
void build_scene_f32(const vpt_sphere *s, int n, SceneF &out) {
    std::memset(&out, 0, sizeof(out));
    out.n_spheres = n;
    // reference point for anchoring: centroid of the ordinary (non-huge) spheres
    const double kHuge = 64.0; // = the product scan's class boundary (vpt_smwave.cuh kSimpleRootMaxR2): every general-form sphere gets its own anchor
    V3 ref{0, 0, 0};
    int n_small = 0;
    for (int i = 0; i < n; ++i)
        if (s[i].r < kHuge) { ref = add(ref, v3(s[i].p)); ++n_small; }
    if (n_small) ref = mul(ref, 1.0 / n_small);
    for (int i = 0; i < n; ++i) {
        MatF &m = out.mat[i];
        m.px = (float)s[i].p[0]; m.py = (float)s[i].p[1]; m.pz = (float)s[i].p[2]; m.r = (float)s[i].r;
        m.cr = (float)s[i].c[0]; m.cg = (float)s[i].c[1]; m.cb = (float)s[i].c[2];
        m.lr = (float)s[i].radiance[0]; m.lg = (float)s[i].radiance[1]; m.lb = (float)s[i].radiance[2];
        for (int k = 0; k < 3; ++k) { m.eta[k] = (float)s[i].eta[k]; m.kappa[k] = (float)s[i].kappa[k]; }
        m.alpha = (float)s[i].alpha;
        m.material = s[i].material;
        m.emits = emits(s[i]);
        if (m.emits) out.emitters[out.n_emitters++] = i;
        if (s[i].r > 0 && s[i].radiance[0] > 0) out.area[out.n_area++] = i; // misSamplingFunctions.h:106
    }
    for (int pass = 0; pass < 2; ++pass) { // scan records: huge (re-anchored) spheres first, then ordinary ones
        for (int i = 0; i < n; ++i) {
            if (!(s[i].r > 0) || (s[i].r >= kHuge) != (pass == 0)) continue;
            GeomF &g = out.geom[out.n_geom++];
            g.id = i;
            const V3 p = v3(s[i].p);
            if (pass == 0) {
                V3 dir = sub(ref, p);
                dir = dot(dir, dir) > 0 ? unit(dir) : V3{1, 0, 0};
                const V3 q = add(p, mul(dir, s[i].r)); // surface point nearest the scene
                g.qx = (float)q.x; g.qy = (float)q.y; g.qz = (float)q.z;
                const V3 qf{g.qx, g.qy, g.qz};
                const V3 mm = sub(qf, p);
                g.mx = (float)mm.x; g.my = (float)mm.y; g.mz = (float)mm.z;
                g.c0 = (float)(dot(mm, mm) - s[i].r * s[i].r); // true |q_f - p|^2 - r^2 (tiny)
                g.r2 = (float)(s[i].r * s[i].r);
                g.big = 1;
                ++out.n_big;
            } else {
                g.qx = (float)p.x; g.qy = (float)p.y; g.qz = (float)p.z;
                g.mx = g.my = g.mz = 0.0f;
                g.r2 = (float)(s[i].r * s[i].r);
                g.c0 = -g.r2;
                g.big = 0;
            }
        }
    }
}

void build_scene_f64(const vpt_sphere *s, int n, SceneD &out) {
    std::memset(&out, 0, sizeof(out));
    out.n_spheres = n;
    for (int i = 0; i < n; ++i) {
        SphereD &d = out.s[i];
        d.r = s[i].r; d.px = s[i].p[0]; d.py = s[i].p[1]; d.pz = s[i].p[2];
        d.cr = s[i].c[0]; d.cg = s[i].c[1]; d.cb = s[i].c[2];
        d.lr = s[i].radiance[0]; d.lg = s[i].radiance[1]; d.lb = s[i].radiance[2];
        for (int k = 0; k < 3; ++k) { d.eta[k] = s[i].eta[k]; d.kappa[k] = s[i].kappa[k]; }
        d.alpha = s[i].alpha;
        d.material = s[i].material;
        d.emits = emits(s[i]);
        if (d.emits) out.emitters[out.n_emitters++] = i;
    }
}

void build_launch(const vpt_params *p, LaunchParams &lp) {
    std::memset(&lp, 0, sizeof(lp));
    lp.width = p->width; lp.height = p->height; lp.n_pixels = p->width * p->height;
    const bool whole = p->sample_begin == 0 && p->sample_end == 0;
    lp.sample_begin = whole ? 0 : p->sample_begin;
    lp.sample_end = whole ? p->spp : p->sample_end;
    const bool all_tiles = p->tile_count == 0;
    lp.tile_rank = all_tiles ? 0 : p->tile_rank;
    lp.tile_count = all_tiles ? 1 : p->tile_count;
    lp.n_tiles_total = (lp.n_pixels + kTile - 1) / kTile;
    lp.method = p->method; lp.max_depth = p->max_depth;
    lp.key0 = (uint32_t)p->seed; lp.key1 = (uint32_t)(p->seed >> 32);
    lp.quirks = p->quirks;
    lp.out_scale = p->output == VPT_OUTPUT_MEAN ? 1.0 / (double)p->spp : 1.0;
    lp.sigma_a = p->sigma_a; lp.sigma_s = p->sigma_s; lp.continue_prob = p->continue_prob;
    lp.march_step = p->march_step; lp.march_source = p->march_source;
    // camera basis, rt.cpp:755-759
    const V3 d = unit(v3(p->cam_dir));
    const V3 cx{p->width * p->fov / p->height, 0., 0.};
    const V3 cy = mul(unit(cross(cx, d)), p->fov);
    const V3 o = v3(p->cam_o);
    const V3 src[4] = {o, d, cx, cy};
    double *dst[4] = {lp.cam_o, lp.cam_d, lp.cam_cx, lp.cam_cy};
    for (int k = 0; k < 4; ++k) { dst[k][0] = src[k].x; dst[k][1] = src[k].y; dst[k][2] = src[k].z; }
}

void build_consts_f32(const LaunchParams &lp, int n_emitters, ConstsF &k) {
    std::memset(&k, 0, sizeof(k));
    const double st = lp.sigma_a + lp.sigma_s;
    k.sigma_t = (float)st;
    k.inv_sigma_t = (float)(1.0 / st);
    k.sigma_s = (float)lp.sigma_s;
    k.albedo_over_cp = (float)(lp.sigma_s / st / lp.continue_prob);
    k.inv_cp = (float)(1.0 / lp.continue_prob);
    k.q = (float)(1.0 - lp.continue_prob);
    k.n_emitters = (float)n_emitters;
    k.method = lp.method; k.max_depth = lp.max_depth;
    for (int i = 0; i < 3; ++i) { k.cam_o[i] = (float)lp.cam_o[i]; k.cam_d[i] = (float)lp.cam_d[i]; k.cam_cx[i] = (float)lp.cam_cx[i]; k.cam_cy[i] = (float)lp.cam_cy[i]; }
    k.inv_w = (float)(1.0 / lp.width); k.inv_h = (float)(1.0 / lp.height);
    k.march_step = (float)lp.march_step; k.march_source = lp.march_source;
}

int owned_tiles(const LaunchParams &lp) { return (lp.n_tiles_total - lp.tile_rank + lp.tile_count - 1) / lp.tile_count; }

// Makes `device` current for the lifetime of the guard and restores the caller's device afterwards: a library embedded in a process that
// has its own current device (PyTorch, one rank per GPU) must not move it.
struct DeviceGuard {
    int prev = -1, rc = VPT_OK;
    explicit DeviceGuard(int device) {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess) { rc = cuda_fail(e, "cudaGetDeviceCount"); return; }
        if (n <= 0 || device < 0 || device >= n) { g_last_cuda_error = "no such CUDA device"; rc = VPT_ERR_NO_DEVICE; return; }
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        if (prev == device) { prev = -1; return; }
        e = cudaSetDevice(device);
        if (e != cudaSuccess) { rc = cuda_fail(e, "cudaSetDevice"); prev = -1; }
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};

// core: enqueue one render into a device buffer. counters_dev may be null.
int enqueue_render(const vpt_params *p, const vpt_sphere *spheres, int n_spheres, float *hdr_dev, cudaStream_t stream, Counters *counters_dev,
                   const LaunchParams &lp, uint64_t *launches) {
    const size_t bytes = (size_t)lp.n_pixels * 3 * sizeof(float);
    const int blocks = owned_tiles(lp);
    int n_emit = 0;
    for (int i = 0; i < n_spheres; ++i) n_emit += emits(spheres[i]);
    if (lp.tile_count > 1 || (n_emit == 0 && lp.method != VPT_METHOD_RAYMARCH) || blocks == 0) CUDA_TRY(cudaMemsetAsync(hdr_dev, 0, bytes, stream));
    if (lp.method == VPT_METHOD_RAYMARCH && lp.march_source >= n_spheres) return VPT_ERR_INVALID_ARGUMENT;
    if ((n_emit == 0 && lp.method != VPT_METHOD_RAYMARCH) || blocks == 0) return VPT_OK; // no emitter: every path returns black (vptShadeMethods.h:1301)
    int rc;
    if (p->precision == VPT_PRECISION_FP32) {
        SceneF sc;
        build_scene_f32(spheres, n_spheres, sc);
        ConstsF cf;
        build_consts_f32(lp, sc.n_emitters, cf);
        if (lp.method == VPT_METHOD_RAYMARCH) rc = launch_march_f32(sc, lp, cf, hdr_dev, counters_dev, stream, blocks);
        else if (p->kernel == VPT_KERNEL_WAVEFRONT_HBM) { rc = launch_hbmwave_f32(sc, lp, cf, hdr_dev, counters_dev, stream, blocks, launches); if (launches && rc == 0) --*launches; }
        else rc = launch_render_f32(sc, lp, cf, hdr_dev, counters_dev, stream, blocks, p->kernel == VPT_KERNEL_AUTO ? VPT_KERNEL_WAVEFRONT_SM : p->kernel);
    } else {
        SceneD sc;
        build_scene_f64(spheres, n_spheres, sc);
        rc = launch_render_f64(sc, lp, hdr_dev, counters_dev, stream, blocks, p->kernel);
    }
    if (rc != 0) return cuda_fail((cudaError_t)rc, "render kernel launch");
    if (launches) ++*launches;
    return VPT_OK;
}

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// Per-device scratch pool for the frame buffer / counters of the host-buffer entry points: a private stream-ordered memory pool
// whose release threshold keeps the memory between calls (cudaMalloc + cudaFree, or the default pool that trims at every
// synchronisation, cost 5-10 ms per frame -- as much as a 64-spp render).  The only state the library keeps; vpt_trim() releases it.
constexpr int kMaxDevices = 64;
std::mutex g_pool_mutex;
cudaMemPool_t g_pools[kMaxDevices] = {};
cudaError_t scratch_pool(int device, cudaMemPool_t *out) {
    if (device < 0 || device >= kMaxDevices) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (!g_pools[device]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        cudaError_t e = cudaMemPoolCreate(&g_pools[device], &props);
        if (e != cudaSuccess) { g_pools[device] = nullptr; return e; }
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(g_pools[device], cudaMemPoolAttrReleaseThreshold, &keep);
    }
    *out = g_pools[device];
    return cudaSuccess;
}
cudaError_t scratch_alloc(int device, void **ptr, size_t bytes, cudaStream_t stream) {
    cudaMemPool_t pool;
    cudaError_t e = scratch_pool(device, &pool);
    if (e != cudaSuccess) return e;
    return cudaMallocFromPoolAsync(ptr, bytes, pool, stream);
}

// trim a device's scratch pool down to `keep` bytes (the HBM wavefront parks 3.6 GB of queues there for one render)
void scratch_trim(int device, size_t keep) {
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (device >= 0 && device < kMaxDevices && g_pools[device]) cudaMemPoolTrimTo(g_pools[device], keep);
}

// One non-blocking stream per (host thread, device) for the host-buffer entry points, created on first use and kept: creating and
// destroying a stream per call cost as much as the copy of a small frame.
struct ThreadStreams {
    cudaStream_t s[kMaxDevices] = {};
    ~ThreadStreams() {
        for (int d = 0; d < kMaxDevices; ++d)
            if (s[d]) { int prev = -1; if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(d) == cudaSuccess) { cudaStreamDestroy(s[d]); cudaSetDevice(prev); } }
    }
};
thread_local ThreadStreams g_streams;
cudaError_t thread_stream(int device, cudaStream_t *out) { // `device` is current
    if (device < 0 || device >= kMaxDevices) return cudaErrorInvalidDevice;
    if (!g_streams.s[device]) {
        cudaError_t e = cudaStreamCreateWithFlags(&g_streams.s[device], cudaStreamNonBlocking);
        if (e != cudaSuccess) { g_streams.s[device] = nullptr; return e; }
    }
    *out = g_streams.s[device];
    return cudaSuccess;
}

// is `p` page-locked host memory the device can address?  (then the render kernel writes the frame straight into it)
float *mapped_device_pointer(float *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (a.type != cudaMemoryTypeHost || !a.devicePointer) return nullptr;
    return (float *)a.devicePointer;
}

} // namespace

void vpt::scratch_trim_(int device, size_t keep) { scratch_trim(device, keep); }
int vpt::scratch_alloc_(int device, void **ptr, size_t bytes, void *stream) { return (int)scratch_alloc(device, ptr, bytes, (cudaStream_t)stream); }

#pragma GCC visibility push(default)
extern "C" {

void vpt_default_params(vpt_params *p) {
    if (!p) return;
    std::memset(p, 0, sizeof(*p));
    p->width = 1024; p->height = 768; // rt.cpp:752
    p->spp = 64;
    p->method = VPT_METHOD_FREE_FLIGHT; // rt.cpp:794
    p->max_depth = 0;
    p->sigma_a = 0.001; p->sigma_s = 0.009; // rt.cpp:794
    p->continue_prob = 0.6;                 // vptShadeMethods.h:1275
    p->cam_o[0] = 0; p->cam_o[1] = 11.2; p->cam_o[2] = 214; // rt.cpp:755
    p->cam_dir[0] = 0; p->cam_dir[1] = -0.042612; p->cam_dir[2] = -1;
    p->fov = 0.5095;                        // rt.cpp:758
    p->seed = 1;
    p->quirks = VPT_QUIRKS_NONE;
    p->precision = VPT_PRECISION_FP32;
    p->output = VPT_OUTPUT_MEAN;
    p->kernel = VPT_KERNEL_AUTO;
    p->device = 0;
    p->march_source = 7; p->march_step = 0.1; // rt.cpp:791
}

int vpt_default_scene(vpt_sphere *out, int32_t cap) {
    // Sphere.cpp:11-22 restated as data
    static const double rows[10][18] = {
        {1e5, -1e5 - 49, 0, 0, .5, .5, .5, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
        {1e5, 1e5 + 49, 0, 0, .0, .0, .5, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
        {1e5, 0, 0, -1e5 - 81.6, .5, .5, .5, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
        {1e5, 0, -1e5 - 40.8, 0, .5, .5, .5, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
        {1e5, 0, 1e5 + 40.8, 0, .5, .5, .5, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
        {16.5, -23, -24.3, -34.6, 0, 0, 0, 0, 0, 0, 1, 1.66058, 0.88143, 0.521467, 9.2282, 6.27077, 4.83803, 0.09},
        {16.5, 23, -24.3, -3.6, .0, .0, .9, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
        {2, 0, 24.3, -35, 0, 0, 0, 100, 100, 0, 0, 0, 0, 0, 0, 0, 0, 0},
        {0, -23, 24.3, 0, 0, 0, 0, 6000, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
        {2, 23, 24.3, 35, 0, 0, 0, 75, 75, 60, 0, 0, 0, 0, 0, 0, 0, 0},
    };
    if (!out || cap < 10) return VPT_ERR_INVALID_ARGUMENT;
    for (int i = 0; i < 10; ++i) {
        const double *d = rows[i];
        vpt_sphere &s = out[i];
        std::memset(&s, 0, sizeof(s));
        s.r = d[0];
        for (int k = 0; k < 3; ++k) { s.p[k] = d[1 + k]; s.c[k] = d[4 + k]; s.radiance[k] = d[7 + k]; s.eta[k] = d[11 + k]; s.kappa[k] = d[14 + k]; }
        s.material = (int32_t)d[10];
        s.alpha = d[17];
    }
    return 10;
}

int vpt_load_scene(const char *path, vpt_sphere *out, int32_t cap) {
    if (!path || !out || cap <= 0) return VPT_ERR_INVALID_ARGUMENT;
    std::FILE *f = std::fopen(path, "r");
    if (!f) return VPT_ERR_IO;
    int n = 0, rc = VPT_OK;
    char line[2048];
    while (rc == VPT_OK && std::fgets(line, sizeof(line), f)) {
        if (char *hash = std::strchr(line, '#')) *hash = 0;
        double v[18];
        int k = 0;
        char *s = line;
        for (;;) {
            while (*s == ' ' || *s == '\t' || *s == ',' || *s == '\r' || *s == '\n') ++s;
            if (!*s) break;
            char *end = nullptr;
            const double x = std::strtod(s, &end);
            if (end == s || k >= 18) { k = -1; break; }
            v[k++] = x; s = end;
        }
        if (k == 0) continue; // blank or comment line
        if (k != 18 || n >= cap || v[10] != std::floor(v[10])) { rc = VPT_ERR_SCENE; break; }
        vpt_sphere &d = out[n++];
        std::memset(&d, 0, sizeof(d));
        d.r = v[0];
        for (int c = 0; c < 3; ++c) { d.p[c] = v[1 + c]; d.c[c] = v[4 + c]; d.radiance[c] = v[7 + c]; d.eta[c] = v[11 + c]; d.kappa[c] = v[14 + c]; }
        d.material = (int32_t)v[10];
        d.alpha = v[17];
    }
    std::fclose(f);
    if (rc != VPT_OK) return rc;
    return n > 0 ? n : VPT_ERR_SCENE;
}

int vpt_render_device(const vpt_params *p, const vpt_sphere *spheres, int32_t n_spheres, float *hdr_dev, void *cuda_stream, vpt_stats *stats) {
    const double t0 = now_ms();
    int rc = validate_params(p, true);
    if (rc) return rc;
    rc = validate_scene(spheres, n_spheres, p->method);
    if (rc) return rc;
    if (!hdr_dev) return VPT_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(p->device);
    if (guard.rc) return guard.rc;
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    LaunchParams lp;
    build_launch(p, lp);
    if (!stats) return enqueue_render(p, spheres, n_spheres, hdr_dev, stream, nullptr, lp, nullptr);

    std::memset(stats, 0, sizeof(*stats));
    Counters *counters_dev = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    CUDA_TRY(scratch_alloc(p->device, (void **)&counters_dev, sizeof(Counters), stream));
    rc = VPT_OK;
    cudaError_t ce;
    do {
        if ((ce = cudaMemsetAsync(counters_dev, 0, sizeof(Counters), stream)) != cudaSuccess) break;
        if ((ce = cudaEventCreate(&e0)) != cudaSuccess) break;
        if ((ce = cudaEventCreate(&e1)) != cudaSuccess) break;
        if ((ce = cudaEventRecord(e0, stream)) != cudaSuccess) break;
        rc = enqueue_render(p, spheres, n_spheres, hdr_dev, stream, counters_dev, lp, &stats->launches);
        if (rc) break;
        if ((ce = cudaEventRecord(e1, stream)) != cudaSuccess) break;
        if ((ce = cudaStreamSynchronize(stream)) != cudaSuccess) break;
        float ms = 0;
        if ((ce = cudaEventElapsedTime(&ms, e0, e1)) != cudaSuccess) break;
        Counters c;
        if ((ce = cudaMemcpy(&c, counters_dev, sizeof(c), cudaMemcpyDeviceToHost)) != cudaSuccess) break;
        stats->kernel_ms = ms;
        stats->events = c.events; stats->scene_scans = c.scans; stats->nonfinite = c.nonfinite;
        stats->paths = c.paths;
        std::memcpy(g_last_debug, c.dbg, sizeof(g_last_debug));
    } while (0);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFreeAsync(counters_dev, stream);
    if (rc) return rc;
    if (ce != cudaSuccess) return cuda_fail(ce, "vpt_render_device");
    stats->total_ms = now_ms() - t0;
    return VPT_OK;
}

int vpt_render(const vpt_params *p, const vpt_sphere *spheres, int32_t n_spheres, float *hdr_rgb, vpt_stats *stats) {
    const double t0 = now_ms();
    int rc = validate_params(p, true);
    if (rc) return rc;
    rc = validate_scene(spheres, n_spheres, p->method);
    if (rc) return rc;
    if (!hdr_rgb) return VPT_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(p->device);
    if (guard.rc) return guard.rc;
    const size_t bytes = (size_t)p->width * p->height * 3 * sizeof(float);
    cudaStream_t stream = nullptr;
    CUDA_TRY(thread_stream(p->device, &stream));
    cudaError_t ce;
    if (float *mapped = mapped_device_pointer(hdr_rgb)) {
        // page-locked, mapped host frame: the kernel stores every pixel once, straight into the caller's buffer, while it renders
        rc = vpt_render_device(p, spheres, n_spheres, mapped, stream, stats);
        ce = cudaStreamSynchronize(stream);
        if (ce != cudaSuccess && rc == VPT_OK) rc = cuda_fail(ce, "vpt_render");
    } else {
        float *dev = nullptr;
        ce = scratch_alloc(p->device, (void **)&dev, bytes, stream); // stream-ordered, from the library's per-device pool
        if (ce != cudaSuccess) return cuda_fail(ce, "cudaMallocAsync");
        rc = vpt_render_device(p, spheres, n_spheres, dev, stream, stats); // with stats == NULL nothing synchronises before the copy below
        if (rc == VPT_OK) {
            ce = cudaMemcpyAsync(hdr_rgb, dev, bytes, cudaMemcpyDeviceToHost, stream);
            if (ce != cudaSuccess) rc = cuda_fail(ce, "copy HDR to host");
        }
        cudaFreeAsync(dev, stream);
        ce = cudaStreamSynchronize(stream);
        if (ce != cudaSuccess && rc == VPT_OK) rc = cuda_fail(ce, "vpt_render");
    }
    if (rc == VPT_OK && stats) stats->total_ms = now_ms() - t0;
    return rc;
}

int vpt_render_multi(const vpt_params *p, const vpt_sphere *spheres, int32_t n_spheres, const int32_t *devices, int32_t n_devices, float *hdr_rgb, vpt_stats *stats) {
    const double t0 = now_ms();
    int rc = validate_params(p, true);
    if (rc) return rc;
    rc = validate_scene(spheres, n_spheres, p->method);
    if (rc) return rc;
    if (!hdr_rgb || !devices || n_devices <= 0) return VPT_ERR_INVALID_ARGUMENT;
    if (p->tile_count > 1) return VPT_ERR_INVALID_ARGUMENT; // the tile split is chosen here
    const size_t bytes = (size_t)p->width * p->height * 3 * sizeof(float);
    const int n_tiles = (p->width * p->height + kTile - 1) / kTile;
    const size_t tile_bytes = (size_t)kTile * 3 * sizeof(float);
    std::vector<int> rcs(n_devices, VPT_OK);
    std::vector<vpt_stats> sts(n_devices);
    std::vector<std::string> errs(n_devices);
    std::memset(hdr_rgb, 0, bytes);
    auto worker = [&](int k) {
        vpt_params q = *p;
        q.device = devices[k];
        q.tile_rank = k; q.tile_count = n_devices;
        DeviceGuard guard(q.device);
        int r = guard.rc;
        float *dev = nullptr;
        cudaStream_t stream = nullptr;
        cudaError_t ae;
        if (!r && (ae = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking)) != cudaSuccess) r = cuda_fail(ae, "cudaStreamCreate");
        if (!r && (ae = scratch_alloc(q.device, (void **)&dev, bytes, stream)) != cudaSuccess) r = cuda_fail(ae, "cudaMallocAsync"); // pooled: kept between calls
        if (!r) r = vpt_render_device(&q, spheres, n_spheres, dev, stream, &sts[k]);
        if (!r) {
            // this device's tiles are disjoint strided ranges of the frame: copy them straight into the caller's buffer
            const int owned = (n_tiles - k + n_devices - 1) / n_devices;
            const size_t pitch = tile_bytes * n_devices, off = tile_bytes * k;
            if (owned > 0) {
                const int full = ((size_t)(owned - 1) * pitch + off + tile_bytes <= bytes) ? owned : owned - 1;
                cudaError_t ce = cudaSuccess;
                if (full > 0) ce = cudaMemcpy2DAsync((char *)hdr_rgb + off, pitch, (char *)dev + off, pitch, tile_bytes, full, cudaMemcpyDeviceToHost, stream);
                if (ce == cudaSuccess && full < owned) { // last, partial tile of the frame
                    const size_t o2 = (size_t)full * pitch + off;
                    ce = cudaMemcpyAsync((char *)hdr_rgb + o2, (char *)dev + o2, bytes - o2, cudaMemcpyDeviceToHost, stream);
                }
                if (ce == cudaSuccess) ce = cudaStreamSynchronize(stream);
                if (ce != cudaSuccess) r = cuda_fail(ce, "copy tiles to host");
            }
        }
        if (dev) cudaFreeAsync(dev, stream);
        if (stream) { cudaStreamSynchronize(stream); cudaStreamDestroy(stream); }
        rcs[k] = r;
        errs[k] = g_last_cuda_error;
    };
    std::vector<std::thread> threads;
    for (int k = 0; k < n_devices; ++k) threads.emplace_back(worker, k);
    for (auto &t : threads) t.join();
    for (int k = 0; k < n_devices; ++k)
        if (rcs[k]) { g_last_cuda_error = errs[k]; return rcs[k]; }
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        for (int k = 0; k < n_devices; ++k) {
            stats->paths += sts[k].paths; stats->events += sts[k].events; stats->scene_scans += sts[k].scene_scans;
            stats->nonfinite += sts[k].nonfinite; stats->launches += sts[k].launches;
            stats->kernel_ms = std::max(stats->kernel_ms, sts[k].kernel_ms);
        }
        stats->total_ms = now_ms() - t0;
    }
    return VPT_OK;
}

// ---- host output stage: mathUtilities.h:34-45, rt.cpp:803,812-820 -----------------------------------------------------------
static inline int display_value(double x) {
    const double c = x < 0.0 ? 0.0 : (x > 1.0 ? 1.0 : x);
    return int(std::pow(c, 1.0 / 2.2) * 255 + .5);
}
int vpt_tonemap(const float *hdr_rgb, int32_t width, int32_t height, uint8_t *rgb8_out) {
    if (!hdr_rgb || !rgb8_out || width <= 0 || height <= 0) return VPT_ERR_INVALID_ARGUMENT;
    const size_t n = (size_t)width * height * 3;
    for (size_t i = 0; i < n; ++i) rgb8_out[i] = (uint8_t)display_value((double)hdr_rgb[i]);
    return VPT_OK;
}
int vpt_write_ppm(const float *hdr_rgb, int32_t width, int32_t height, const char *path) {
    if (!hdr_rgb || !path || width <= 0 || height <= 0) return VPT_ERR_INVALID_ARGUMENT;
    FILE *f = std::fopen(path, "w");
    if (!f) return VPT_ERR_IO;
    std::fprintf(f, "P3\n%d %d\n%d\n", width, height, 255);
    const size_t n = (size_t)width * height;
    for (size_t i = 0; i < n; ++i)
        std::fprintf(f, "%d %d %d ", display_value(hdr_rgb[3 * i]), display_value(hdr_rgb[3 * i + 1]), display_value(hdr_rgb[3 * i + 2]));
    return std::fclose(f) == 0 ? VPT_OK : VPT_ERR_IO;
}

int vpt_write_pfm(const float *hdr_rgb, int32_t width, int32_t height, const char *path) {
    if (!hdr_rgb || !path || width <= 0 || height <= 0) return VPT_ERR_INVALID_ARGUMENT;
    FILE *f = std::fopen(path, "wb");
    if (!f) return VPT_ERR_IO;
    std::fprintf(f, "PF\n%d %d\n-1.0\n", width, height);
    bool ok = true;
    for (int32_t row = height - 1; row >= 0 && ok; --row) // PFM stores the bottom row first; x86-64 and the GPU hosts are little-endian
        ok = std::fwrite(hdr_rgb + (size_t)row * width * 3, sizeof(float), (size_t)width * 3, f) == (size_t)width * 3;
    return (std::fclose(f) == 0 && ok) ? VPT_OK : VPT_ERR_IO;
}

// ---- unit kernels -----------------------------------------------------------------------------------------------------------------
static const int kUnitStrides[VPT_UNIT_COUNT_][2] = {
    {7, 1}, {6, 3}, {6, 1}, {7, 1}, {2, 4}, {2, 3}, {9, 6}, {2, 1}, {5, 4}, {7, 4}, {13, 6}, {3, 3}, {10, 3}, {11, 3}, {19, 3}, {9, 6}, {8, 4}, {4, 3}, {127, 4}, {8, 4}, {11, 3}, {6, 7},
};
int vpt_unit_strides(int32_t fn, int32_t *in_stride, int32_t *out_stride) {
    if (fn < 0 || fn >= VPT_UNIT_COUNT_) return VPT_ERR_INVALID_ARGUMENT;
    if (in_stride) *in_stride = kUnitStrides[fn][0];
    if (out_stride) *out_stride = kUnitStrides[fn][1];
    return VPT_OK;
}

int vpt_unit(int32_t fn, const vpt_params *p, const vpt_sphere *spheres, int32_t n_spheres, int32_t n, const double *in, int32_t in_stride,
             double *out, int32_t out_stride) {
    if (fn < 0 || fn >= VPT_UNIT_COUNT_ || !in || !out || n <= 0) return VPT_ERR_INVALID_ARGUMENT;
    if (in_stride < kUnitStrides[fn][0] || out_stride < kUnitStrides[fn][1]) return VPT_ERR_INVALID_ARGUMENT;
    int rc = validate_params(p, fn == VPT_UNIT_CAMERA_RAY);
    if (rc) return rc;
    rc = validate_scene(spheres, n_spheres, p->method);
    if (rc) return rc;
    DeviceGuard guard(p->device);
    if (guard.rc) return guard.rc;
    LaunchParams lp;
    vpt_params q = *p;
    if (q.width <= 0) q.width = 1;
    if (q.height <= 0) q.height = 1;
    if (q.spp <= 0) q.spp = 1;
    build_launch(&q, lp);
    double *din = nullptr, *dout = nullptr;
    const size_t bi = (size_t)n * in_stride * sizeof(double), bo = (size_t)n * out_stride * sizeof(double);
    CUDA_TRY(cudaMalloc(&din, bi));
    cudaError_t ce = cudaMalloc(&dout, bo);
    if (ce != cudaSuccess) { cudaFree(din); return cuda_fail(ce, "cudaMalloc"); }
    int launch_rc = 0;
    do {
        if ((ce = cudaMemcpy(din, in, bi, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((ce = cudaMemset(dout, 0, bo)) != cudaSuccess) break;
        if (p->precision == VPT_PRECISION_FP32) {
            SceneF sc;
            build_scene_f32(spheres, n_spheres, sc);
            ConstsF cf;
            build_consts_f32(lp, sc.n_emitters, cf);
            launch_rc = launch_unit_f32(fn, sc, lp, cf, n, din, in_stride, dout, out_stride, nullptr);
        } else {
            SceneD sc;
            build_scene_f64(spheres, n_spheres, sc);
            launch_rc = launch_unit_f64(fn, sc, lp, n, din, in_stride, dout, out_stride, nullptr);
        }
        if (launch_rc) { ce = (cudaError_t)launch_rc; break; }
        if ((ce = cudaDeviceSynchronize()) != cudaSuccess) break;
        ce = cudaMemcpy(out, dout, bo, cudaMemcpyDeviceToHost);
    } while (0);
    cudaFree(din);
    cudaFree(dout);
    if (ce != cudaSuccess) return cuda_fail(ce, "vpt_unit");
    return VPT_OK;
}

int vpt_philox(int32_t device, int32_t n, const uint32_t *ctr, const uint32_t *key, uint32_t *out) {
    if (!ctr || !key || !out || n <= 0) return VPT_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(device);
    if (guard.rc) return guard.rc;
    uint32_t *dc = nullptr, *dk = nullptr, *dout = nullptr;
    cudaError_t ce;
    do {
        if ((ce = cudaMalloc(&dc, (size_t)n * 16)) != cudaSuccess) break;
        if ((ce = cudaMalloc(&dk, (size_t)n * 8)) != cudaSuccess) break;
        if ((ce = cudaMalloc(&dout, (size_t)n * 16)) != cudaSuccess) break;
        if ((ce = cudaMemcpy(dc, ctr, (size_t)n * 16, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((ce = cudaMemcpy(dk, key, (size_t)n * 8, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        const int lrc = launch_philox(n, dc, dk, dout, nullptr);
        if (lrc) { ce = (cudaError_t)lrc; break; }
        ce = cudaMemcpy(out, dout, (size_t)n * 16, cudaMemcpyDeviceToHost);
    } while (0);
    cudaFree(dc); cudaFree(dk); cudaFree(dout);
    if (ce != cudaSuccess) return cuda_fail(ce, "vpt_philox");
    return VPT_OK;
}

int vpt_measure_fp32_peak(int32_t device, double *tflops_out, double *sm_clock_mhz_out) {
    if (!tflops_out) return VPT_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(device);
    if (guard.rc) return guard.rc;
    int sms = 0, khz = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    CUDA_TRY(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device));
    float *sink = nullptr;
    CUDA_TRY(cudaMalloc(&sink, 16));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * 8, threads = 256, iters = 8192;
    double best = 0;
    cudaError_t ce = cudaSuccess;
    for (int rep = 0; rep < 6 && ce == cudaSuccess; ++rep) {
        cudaEventRecord(e0);
        const int lrc = launch_fma_peak(sink, blocks, threads, iters, nullptr);
        if (lrc) { ce = (cudaError_t)lrc; break; }
        cudaEventRecord(e1);
        ce = cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = (double)blocks * threads * iters * kFmaPeakFlopsPerThreadIter;
        if (rep > 0 && ms > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(sink);
    if (ce != cudaSuccess) return cuda_fail(ce, "fma peak");
    *tflops_out = best;
    if (sm_clock_mhz_out) *sm_clock_mhz_out = khz / 1000.0;
    return VPT_OK;
}

void *vpt_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (bytes == 0 || cudaHostAlloc(&p, bytes, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void vpt_host_free(void *p) { if (p) cudaFreeHost(p); }

int vpt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char *vpt_strerror(int status) {
    switch (status) {
    case VPT_OK: return "ok";
    case VPT_ERR_INVALID_ARGUMENT: return "invalid argument";
    case VPT_ERR_SCENE: return "invalid scene (sphere count, emitter count or a non-finite / negative field)";
    case VPT_ERR_UNSUPPORTED: return "unsupported (material 3; the superseded MEGA_SCAN / WAVEFRONT kernels; wavefront kernel in fp64; quirks in fp32 precision)";
    case VPT_ERR_NO_DEVICE: return "no usable CUDA device (this library has no CPU fallback)";
    case VPT_ERR_CUDA: return "CUDA runtime error (see vpt_last_cuda_error)";
    case VPT_ERR_IO: return "I/O error";
    default: return "unknown status";
    }
}
const char *vpt_last_cuda_error(void) { return g_last_cuda_error.c_str(); }
void vpt_trim(void) {
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    for (int d = 0; d < kMaxDevices; ++d)
        if (g_pools[d]) { cudaMemPoolDestroy(g_pools[d]); g_pools[d] = nullptr; }
}
/* development aid, not part of include/vpt.h: the in-kernel cycle counters of the last render with stats on this thread
 * (all zero unless the library was built with -DVPT_SMWAVE_PROFILE) */
int vpt_debug_counters(unsigned long long *out, int32_t cap) {
    const int n = cap < kDebugCounters ? cap : kDebugCounters;
    for (int i = 0; i < n; ++i) out[i] = g_last_debug[i];
    return n;
}
const char *vpt_version(void) { return "vpt_b200 0.1 (sm_100a)"; }

} // extern "C"
#pragma GCC visibility pop
