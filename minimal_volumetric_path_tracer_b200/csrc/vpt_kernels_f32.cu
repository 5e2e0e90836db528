// vpt_kernels_f32.cu -- FP32 kernels of libvpt_b200 for sm_100a: the product kernel's launch (vpt_smwave.cuh), the one-thread-per-pixel
// megakernel kept as the measured alternative, the ray marcher, the unit kernels behind vpt_unit(), the Philox test kernel and the FFMA
// peak probe.  Every estimator evaluated here is vpt_stages.cuh -- the same functions the product kernel runs.
#include <cuda_runtime.h>
#include <mutex>
#include <vector>
#include "vpt_smwave.cuh"
#include "vpt_march.cuh"

namespace vpt {

using namespace f32;

// ---- context of the megakernel and the unit kernels: ONE thread drives one record through the stages ---------------------------------
struct ThreadCtx {
    const SmScene &S;
    const ConstsF &k;
    uint32_t key0, key1;
    unsigned events = 0, scans = 0, nonfinite = 0;
    double acc[3] = {0.0, 0.0, 0.0}; // radiance collected (the pixel's sum in the megakernel)
    __device__ ThreadCtx(const SmScene &S_, const ConstsF &k_, uint32_t k0, uint32_t k1) : S(S_), k(k_), key0(k0), key1(k1) {}
    __device__ __forceinline__ float4 rnd(const Rec &r, uint32_t block) const {
        const uint4 b = philox_block(r.pixel, r.sample, r.depth, block, key0, key1);
        return make_float4(u32_to_unit_f32(b.x), u32_to_unit_f32(b.y), u32_to_unit_f32(b.z), u32_to_unit_f32(b.w));
    }
    __device__ __forceinline__ float4 jitter(const Rec &r) const {
        const uint4 b = philox_block(r.pixel, r.sample, kJitterBounce, 0, key0, key1);
        return make_float4(u32_to_unit_f32(b.x), u32_to_unit_f32(b.y), 0.0f, 0.0f);
    }
    __device__ __forceinline__ bool scan(F3 o, F3 d, float &t, int &id) { return scan_sm(S, o, d, t, id); }
    __device__ __forceinline__ void add(const Rec &, F3 L) {
        if (!(fabsf(L.x) < kSmMaxContribution && fabsf(L.y) < kSmMaxContribution && fabsf(L.z) < kSmMaxContribution)) { ++nonfinite; return; }
        acc[0] += (double)L.x; acc[1] += (double)L.y; acc[2] += (double)L.z;
    }
    __device__ __forceinline__ void last_step() {}
    __device__ __forceinline__ void vertex(const Rec &, int) {}
};

// ---- the megakernel (VPT_KERNEL_MEGA): one thread owns one pixel and walks its samples, each path through trace_path --------------------
// Kept as the measured alternative to the wavefronts (north_star: "megakernel or wavefront, chosen by measurement"): lanes of a warp sit in
// different stages of different paths, so the scans run at a third of the lanes (profiles/r1_summary.md v0).
template <int METHOD>
__global__ void __launch_bounds__(kThreadsPerBlock) render_f32_kernel(const __grid_constant__ SceneF sc, const __grid_constant__ LaunchParams lp,
                                                                       const __grid_constant__ ConstsF cf, float *__restrict__ hdr, Counters *__restrict__ counters) {
    SmScene &S = *reinterpret_cast<SmScene *>(smwave_smem);
    stage_scene(S, sc, (int)threadIdx.x, (int)blockDim.x);
    __syncthreads();
    stage_scene_tables(S, (int)threadIdx.x, (int)blockDim.x);
    __syncthreads();
    const long long tile = (long long)blockIdx.x * lp.tile_count + lp.tile_rank;
    const long long pixel = tile * kTile + threadIdx.x;
    if (pixel >= lp.n_pixels) return;
    ThreadCtx c(S, cf, lp.key0, lp.key1);
    for (int s = lp.sample_begin; s < lp.sample_end; ++s) {
        Rec r;
        r.aux = 0u;
        if (stage_gen(c, true, (uint32_t)pixel, (uint32_t)s, lp.width, lp.height, r)) trace_path<METHOD>(c, r);
    }
    float *out = hdr + pixel * 3;
    out[0] = (float)(c.acc[0] * lp.out_scale);
    out[1] = (float)(c.acc[1] * lp.out_scale);
    out[2] = (float)(c.acc[2] * lp.out_scale);
    if (!counters) return;
    atomicAdd(&counters->events, (unsigned long long)c.events); atomicAdd(&counters->scans, (unsigned long long)c.scans);
    if (c.nonfinite) atomicAdd(&counters->nonfinite, (unsigned long long)c.nonfinite);
    atomicAdd(&counters->paths, (unsigned long long)(lp.sample_end - lp.sample_begin));
}

// ---- SM-wide wavefront (vpt_smwave.cuh) -----------------------------------------------------------------------------------------
template <int METHOD, int SLOTS>
__global__ void __launch_bounds__(kSmThreads, 1) render_f32_smwave_kernel(const __grid_constant__ SceneF sc, const __grid_constant__ LaunchParams lp,
                                                                           const __grid_constant__ ConstsF cf, float *__restrict__ hdr, Counters *__restrict__ counters,
                                                                           int log_p, int n_owned_tiles, int n_items, int zero) {
    SmShared<SLOTS> &M = sm_shared<SLOTS>();
    const int tid = (int)threadIdx.x;
    stage_scene(M.scene, sc, tid, kSmThreads);
    for (int i = tid; i < kSmPool; i += kSmThreads) M.meta[i] = 0u;
    if (tid < 32) { // record 0 backs the idle lanes of partial batches before it is first allocated: give it in-range values
        M.ox[0] = M.oy[0] = M.oz[0] = 0.0f; M.dx[0] = M.dy[0] = 0.0f; M.dz[0] = 1.0f; M.br[0] = M.bg[0] = M.bb[0] = 0.0f;
        M.sample[0] = 0u; M.xd[0] = M.xs[0] = 0.5f;
    }
    __syncthreads();
    stage_scene_tables(M.scene, tid, kSmThreads);
    SmWave<METHOD, SLOTS> wf(M, cf, lp, log_p, n_owned_tiles, zero);
    wf.init(n_items);
    __syncthreads();
    wf.run(hdr, n_items, kSmFixInv);
    if (!counters) return;
    unsigned long long ev = wf.events, scn = wf.scans, nf = wf.nonfinite, np = wf.paths;
    for (int off = 16; off > 0; off >>= 1) {
        ev += __shfl_down_sync(0xffffffffu, ev, off); scn += __shfl_down_sync(0xffffffffu, scn, off);
        nf += __shfl_down_sync(0xffffffffu, nf, off); np += __shfl_down_sync(0xffffffffu, np, off);
    }
    if ((tid & 31) == 0) {
        atomicAdd(&counters->events, ev); atomicAdd(&counters->scans, scn); atomicAdd(&counters->paths, np);
        if (nf) atomicAdd(&counters->nonfinite, nf);
#ifdef VPT_SMWAVE_PROFILE
        for (int i = 0; i < 24; ++i) if (wf.prof[i]) atomicAdd(&counters->dbg[i], wf.prof[i]);
#endif
    }
}

// ---- Philox round keys of the product kernel in constant memory (vpt_philox.cuh philox_block_ck) ------------------------------------------
// One schedule per device at a time.  Launches with the same seed share it (each waits for the copy that set it); a launch with another
// seed first waits -- on the device, through events -- for every launch that may still read the old one.  Nothing here blocks the host.
namespace {
struct KeySlot {
    std::mutex m;
    bool valid = false;
    uint32_t k0 = 0, k1 = 0;
    cudaEvent_t set_ev = nullptr;
    std::vector<cudaEvent_t> users, spare;
};
KeySlot g_key_slots[64];

cudaError_t philox_keys_begin(KeySlot &K, cudaStream_t st, uint32_t k0, uint32_t k1) { // K.m is held until philox_keys_end
    cudaError_t e;
    if (!K.set_ev && (e = cudaEventCreateWithFlags(&K.set_ev, cudaEventDisableTiming)) != cudaSuccess) return e;
    if (K.valid && K.k0 == k0 && K.k1 == k1) return cudaStreamWaitEvent(st, K.set_ev, 0);
    for (cudaEvent_t ev : K.users) {
        if ((e = cudaStreamWaitEvent(st, ev, 0)) != cudaSuccess) return e;
        K.spare.push_back(ev);
    }
    K.users.clear();
    uint32_t ks[20];
    for (uint32_t i = 0; i < 10; ++i) { ks[2 * i] = k0 + i * 0x9E3779B9u; ks[2 * i + 1] = k1 + i * 0xBB67AE85u; }
    K.valid = false;
    if ((e = cudaMemcpyToSymbolAsync(c_philox_ks, ks, sizeof(ks), 0, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
    if ((e = cudaEventRecord(K.set_ev, st)) != cudaSuccess) return e;
    K.valid = true; K.k0 = k0; K.k1 = k1;
    return cudaSuccess;
}
cudaError_t philox_keys_end(KeySlot &K, cudaStream_t st) { // after the launch: remember it as a reader of the current schedule
    if (K.users.size() >= 16) { // forget the readers that have finished
        std::vector<cudaEvent_t> busy;
        for (cudaEvent_t ev : K.users) (cudaEventQuery(ev) == cudaSuccess ? K.spare : busy).push_back(ev);
        cudaGetLastError(); // (cudaErrorNotReady is not an error)
        K.users.swap(busy);
    }
    cudaEvent_t ev = nullptr;
    cudaError_t e = cudaSuccess;
    if (!K.spare.empty()) { ev = K.spare.back(); K.spare.pop_back(); }
    else e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e == cudaSuccess && (e = cudaEventRecord(ev, st)) == cudaSuccess) K.users.push_back(ev);
    return e;
}
} // namespace

template <int METHOD, int SLOTS>
static int launch_smwave_slots(const SceneF &scene, const LaunchParams &lp, const ConstsF &cf, float *hdr_dev, Counters *counters_dev, cudaStream_t st, int n_owned_tiles) {
    int dev = 0, n_sm = 0;
    cudaError_t e;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return (int)e;
    if ((e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return (int)e;
    if ((e = cudaFuncSetAttribute(render_f32_smwave_kernel<METHOD, SLOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmShared<SLOTS>))) != cudaSuccess) return (int)e;
    const int log_p = 7; // work item: one 128-pixel tile
    const int n_items = n_owned_tiles;
    const int grid = n_items < n_sm ? n_items : n_sm;
#ifdef VPT_PHILOX_ARG_KEYS
    render_f32_smwave_kernel<METHOD, SLOTS><<<grid, kSmThreads, sizeof(SmShared<SLOTS>), st>>>(scene, lp, cf, hdr_dev, counters_dev, log_p, n_owned_tiles, n_items, 0);
    return (int)cudaGetLastError();
#else
    if (dev < 0 || dev >= 64) return (int)cudaErrorInvalidDevice;
    KeySlot &K = g_key_slots[dev];
    std::lock_guard<std::mutex> lock(K.m);
    if ((e = philox_keys_begin(K, st, lp.key0, lp.key1)) != cudaSuccess) return (int)e;
    render_f32_smwave_kernel<METHOD, SLOTS><<<grid, kSmThreads, sizeof(SmShared<SLOTS>), st>>>(scene, lp, cf, hdr_dev, counters_dev, log_p, n_owned_tiles, n_items, 0);
    if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
    return (int)philox_keys_end(K, st);
#endif
}
// Work items in flight per CTA (vpt_smsched.cuh): six below 96 samples per pixel, four below 384, two from there on.  Measured on 1024x768,
// equi-angular, Mpaths/s with 2 / 4 / 6 items: 16 spp - / 3221 / 4102, 32 spp 3196 / 5112 / 6461, 64 spp 4948 / 7640 / 8089 (two 256-pixel
// items: 6999), 128 spp 6971 / 8461 / 8311, 256 spp 8398 / 8559 / 8408, 512 spp 8599 / 8510 / -, 1024 spp 8631 / 8532 / -.
template <int METHOD>
static int launch_smwave(const SceneF &scene, const LaunchParams &lp, const ConstsF &cf, float *hdr_dev, Counters *counters_dev, cudaStream_t st, int n_owned_tiles) {
#ifdef VPT_ITEM_SLOTS_RUN
    const int slots = VPT_ITEM_SLOTS_RUN;
#else
    const int spp = lp.sample_end - lp.sample_begin;
    const int slots = spp < 96 ? 6 : (spp < 384 ? 4 : 2);
#endif
    if (slots > 4) return launch_smwave_slots<METHOD, 6>(scene, lp, cf, hdr_dev, counters_dev, st, n_owned_tiles);
    if (slots > 2) return launch_smwave_slots<METHOD, 4>(scene, lp, cf, hdr_dev, counters_dev, st, n_owned_tiles);
    return launch_smwave_slots<METHOD, 2>(scene, lp, cf, hdr_dev, counters_dev, st, n_owned_tiles);
}

// ---- ray-marching reference solver (vpt_march.cuh): one thread per pixel ---------------------------------------------------------
__global__ void __launch_bounds__(kThreadsPerBlock) render_f32_march_kernel(const __grid_constant__ SceneF sc, const __grid_constant__ LaunchParams lp,
                                                                             const __grid_constant__ ConstsF cf, float *__restrict__ hdr, Counters *__restrict__ counters) {
    SmScene &S = *reinterpret_cast<SmScene *>(smwave_smem);
    stage_scene(S, sc, (int)threadIdx.x, (int)blockDim.x);
    __syncthreads();
    const long long tile = (long long)blockIdx.x * lp.tile_count + lp.tile_rank;
    const long long pixel = tile * kTile + threadIdx.x;
    if (pixel >= lp.n_pixels) return;
    double acc[3] = {0, 0, 0};
    unsigned scans = 0, nonfinite = 0;
    for (int s = lp.sample_begin; s < lp.sample_end; ++s) {
        const uint4 j = philox_block((uint32_t)pixel, (uint32_t)s, kJitterBounce, 0, lp.key0, lp.key1);
        const F3 d = camera_dir(cf, (uint32_t)pixel, lp.width, lp.height, u32_to_unit_f32(j.x), u32_to_unit_f32(j.y));
        double L[3]; unsigned n_steps;
        ray_march3(S, mk(cf.cam_o[0], cf.cam_o[1], cf.cam_o[2]), d, lp.march_step, cf.march_source, cf.sigma_t, cf.sigma_s, L, n_steps, scans);
        if (isfinite(L[0] + L[1] + L[2])) { acc[0] += L[0]; acc[1] += L[1]; acc[2] += L[2]; } else ++nonfinite;
    }
    float *out = hdr + pixel * 3;
    out[0] = (float)(acc[0] * lp.out_scale); out[1] = (float)(acc[1] * lp.out_scale); out[2] = (float)(acc[2] * lp.out_scale);
    if (!counters) return;
    atomicAdd(&counters->events, (unsigned long long)(lp.sample_end - lp.sample_begin)); atomicAdd(&counters->scans, (unsigned long long)scans);
    if (nonfinite) atomicAdd(&counters->nonfinite, (unsigned long long)nonfinite);
    atomicAdd(&counters->paths, (unsigned long long)(lp.sample_end - lp.sample_begin));
}
int launch_march_f32(const SceneF &scene, const LaunchParams &lp, const ConstsF &cf, float *hdr_dev, Counters *counters_dev, void *stream, int n_blocks) {
    render_f32_march_kernel<<<n_blocks, kThreadsPerBlock, sizeof(SmScene), (cudaStream_t)stream>>>(scene, lp, cf, hdr_dev, counters_dev);
    return (int)cudaGetLastError();
}

int launch_render_f32(const SceneF &scene, const LaunchParams &lp, const ConstsF &cf, float *hdr_dev, Counters *counters_dev, void *stream, int n_blocks, int kernel) {
    cudaStream_t st = (cudaStream_t)stream;
    if (kernel == VPT_KERNEL_MEGA) {
        switch (lp.method) {
        case 0: render_f32_kernel<0><<<n_blocks, kThreadsPerBlock, sizeof(SmScene), st>>>(scene, lp, cf, hdr_dev, counters_dev); break;
        case 1: render_f32_kernel<1><<<n_blocks, kThreadsPerBlock, sizeof(SmScene), st>>>(scene, lp, cf, hdr_dev, counters_dev); break;
        case 4: render_f32_kernel<4><<<n_blocks, kThreadsPerBlock, sizeof(SmScene), st>>>(scene, lp, cf, hdr_dev, counters_dev); break;
        default: render_f32_kernel<2><<<n_blocks, kThreadsPerBlock, sizeof(SmScene), st>>>(scene, lp, cf, hdr_dev, counters_dev); break;
        }
        return (int)cudaGetLastError();
    }
    switch (lp.method) { // VPT_KERNEL_WAVEFRONT_SM
    case 0: return launch_smwave<0>(scene, lp, cf, hdr_dev, counters_dev, st, n_blocks);
    case 1: return launch_smwave<1>(scene, lp, cf, hdr_dev, counters_dev, st, n_blocks);
    case 4: return launch_smwave<4>(scene, lp, cf, hdr_dev, counters_dev, st, n_blocks);
    default: return launch_smwave<2>(scene, lp, cf, hdr_dev, counters_dev, st, n_blocks);
    }
}

// ---- unit kernels (include/vpt.h vpt_unit_fn) ---------------------------------------------------------------------------
__device__ __forceinline__ F3 ld3(const double *p) { return mk((float)p[0], (float)p[1], (float)p[2]); }
__device__ __forceinline__ void st3(double *p, F3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
__device__ __forceinline__ void st3d(double *p, const double *v) { p[0] = v[0]; p[1] = v[1]; p[2] = v[2]; }

// A stage context whose random numbers are an EXPLICIT list of uniforms in the reference's consumption order (e.g. the reference's own
// erand48 sequence, or the draws of one golden vector): per bounce  roulette, light pick, distance[, decision]  then
//   medium vertex:   2 cone numbers (NEE), 2 phase-function numbers
//   surface vertex:  per area light 2 cone numbers (+ 1 for a dielectric, misSamplingFunctions.h:116), 2 (dielectric: 1) for the BSDF-sampled
//                    light term of MISv2, 2 (dielectric: 1) for bdsf
// vertex() deals the vertex's numbers into the Philox block layout the stages ask for (vpt_philox.cuh slot table); block 0 of the next
// bounce is dealt on demand, so that a path ended by the roulette has consumed exactly the reference's count.
struct ListCtx {
    const SmScene &S;
    const ConstsF &k;
    const double *u; int n; int i = 0; bool overrun = false;
    unsigned events = 0, scans = 0, nonfinite = 0;
    double acc[3] = {0.0, 0.0, 0.0};
    float tab[12][4]; // blocks 1 .. 11 of the current vertex (block 10 + : the dielectric's per-light draws, slots 40 ..)
    __device__ ListCtx(const SmScene &S_, const ConstsF &k_, const double *u_, int n_) : S(S_), k(k_), u(u_), n(n_) {
        for (int b = 0; b < 12; ++b) tab[b][0] = tab[b][1] = tab[b][2] = tab[b][3] = 0.5f;
    }
    __device__ float next() { if (i >= n) { overrun = true; return 0.0f; } /* 0 < q: the next roulette draw ends the path */ return (float)u[i++]; }
    __device__ float4 rnd(const Rec &r, uint32_t block) {
        if (block == 0u) { // the header of bounce r.depth, in the reference's order and with its short-circuits
            if ((k.max_depth > 0 && (int)r.depth >= k.max_depth) || r.depth >= (uint32_t)VPT_MAX_DEPTH) return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            float4 h = make_float4(next(), 0.0f, 0.0f, 0.5f);
            if (h.x < k.q) return h;
            h.y = next(); h.z = next();
            if (k.method != 0) h.w = next();
            return h;
        }
        const uint32_t b = block < 12u ? block : 11u;
        return make_float4(tab[b][0], tab[b][1], tab[b][2], tab[b][3]);
    }
    __device__ float4 jitter(const Rec &) { return make_float4(0.5f, 0.5f, 0.0f, 0.0f); }
    __device__ __forceinline__ bool scan(F3 o, F3 d, float &t, int &id) { return scan_sm(S, o, d, t, id); }
    __device__ __forceinline__ void add(const Rec &, F3 L) {
        if (!(fabsf(L.x) < kSmMaxContribution && fabsf(L.y) < kSmMaxContribution && fabsf(L.z) < kSmMaxContribution)) { ++nonfinite; return; }
        acc[0] += (double)L.x; acc[1] += (double)L.y; acc[2] += (double)L.z;
    }
    __device__ __forceinline__ void last_step() {}
    __device__ void set_slot(uint32_t slot, float v) { const uint32_t b = slot >> 2; if (b >= 1u && b < 12u) tab[b][slot & 3u] = v; }
    __device__ void vertex(const Rec &r, int dest) {
        if (dest == SQ_MED_POINT || dest == SQ_MED_AREA) {
            set_slot(S_NEE, next()); set_slot(S_NEE + 1, next()); set_slot(S_PHASE, next()); set_slot(S_PHASE + 1, next());
        } else if (dest == SQ_SURF_P || dest == SQ_SURF_L || dest == SQ_SURF_F) {
            const bool diel = S.mats[r.hid].material == 2;
            for (int a = 0; a < S.n_area; ++a) {
                set_slot(S_AREA + 2 * a, next()); set_slot(S_AREA + 2 * a + 1, next());
                if (diel) set_slot(S_DIEL + a, next());
            }
            set_slot(S_MIS, next()); if (!diel) set_slot(S_MIS + 1, next());
            set_slot(S_BSDF, next()); if (!diel) set_slot(S_BSDF + 1, next());
        }
    }
};

__global__ void unit_f32_kernel(int fn, const __grid_constant__ SceneF sc, const __grid_constant__ LaunchParams lp, const __grid_constant__ ConstsF cf, int n, const double *__restrict__ in,
                                int in_stride, double *__restrict__ out, int out_stride) {
    SmScene &PS = *reinterpret_cast<SmScene *>(smwave_smem); // the product kernel's scene (dynamic shared memory): same staging, same scan
    stage_scene(PS, sc, (int)threadIdx.x, (int)blockDim.x);
    __syncthreads();
    stage_scene_tables(PS, (int)threadIdx.x, (int)blockDim.x);
    __syncthreads();
    const MatF *mats = PS.mats;
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    const double *a = in + (size_t)row * in_stride;
    double *o = out + (size_t)row * out_stride;
    Rec r;
    r.o = mk(0, 0, 0); r.d = mk(0, 0, 1); r.beta = mk(1, 1, 1); r.pixel = r.sample = r.depth = r.src = r.hid = r.aux = 0u; r.xi_dist = r.xi_decide = 0.5f;
    switch (fn) {
    case VPT_UNIT_SPHERE_INTERSECT: o[0] = sphere_t_sm(PS, (int)a[0], ld3(a + 1), ld3(a + 4)); break; // the scan's own root arithmetic, one sphere
    case VPT_UNIT_INTERSECT: { // the product kernel's scan (vpt_scan.cuh); on a miss the reference leaves id untouched (0)
        float t = 0.0f; int id = 0;
        const bool h = scan_sm(PS, ld3(a), ld3(a + 3), t, id);
        o[0] = h; o[1] = h ? t : 0.0; o[2] = h ? id : 0;
    } break;
    case VPT_UNIT_VISIBILITY: { // as the stages' point-light shadow rays: nothing hit before distance * (1 - 1e-4)
        const F3 light = ld3(a), lx = light - ld3(a + 3);
        const float d2 = dot(lx, lx), inv = rsqrtf(d2);
        float t; int id;
        const bool h = scan_sm(PS, light, lx * (-inv), t, id);
        o[0] = !h || t > d2 * inv * (1.0f - 1e-4f);
    } break;
    case VPT_UNIT_TRANSMITTANCE: { // the stages' transmittance (vpt_stages.cuh transmit: ex2.approx)
        const F3 v = ld3(a + 3) - ld3(a);
        o[0] = transmit((float)a[6] * sqrtf(dot(v, v)));
    } break;
    case VPT_UNIT_FREE_FLIGHT: { // stage_primary<0>'s distance and the stages' transmittance
        const float st = (float)a[0], inv_st = (float)(1.0 / a[0]), xi = (float)a[1];
        const float d = -logf(1.0f - xi) * inv_st;
        const float e = transmit(st * d);
        o[0] = d; o[1] = st * e; o[2] = 1.0f - e; o[3] = e;
    } break;
    case VPT_UNIT_PHASE_SAMPLE: st3(o, phase_sample((float)a[0], (float)a[1])); break;
    case VPT_UNIT_EQUIANGULAR: {
        const MatF &src = mats[(int)a[0]];
        const float tmax = (float)fmin(a[1], (double)kMaxFloat);
        const F3 org = ld3(a + 2), dir = ld3(a + 5);
        const float xi = (float)a[8];
        const F3 light = mk(src.px, src.py, src.pz);
        float D, dth, tl; // the stages' form (vpt_f32.cuh equiangular_sample): the two angles are reported for the comparison only
        const float dist = equiangular_sample(light, org, dir, tmax, xi, D, dth, tl);
        const float thA = atan2f(-dot(light - org, dir), D);
        o[0] = D; o[1] = thA; o[2] = thA + dth; o[3] = tl; o[4] = dist;
        o[5] = D / (dth * (tl * tl + D * D));
    } break;
    case VPT_UNIT_MIS_DISTANCE: {
        const MatF &src = mats[(int)a[0]];
        const float tmax = (float)fmin(a[1], (double)kMaxFloat), st = (float)a[8];
        float dist, inv_pdf;
        const bool surface = mis_distance(mk(src.px, src.py, src.pz), ld3(a + 2), ld3(a + 5), tmax, transmit(st * tmax), st, 1.0f / st, (float)a[9], (float)a[10], dist, inv_pdf);
        o[0] = surface; o[1] = dist; o[2] = surface ? 1.0f : 1.0f / inv_pdf;
    } break;
    case VPT_UNIT_DIELECTRIC: {
        const Frame fr = make_frame(ld3(a));
        const DielF di = dielectric_setup(unit(to_local(fr, ld3(a + 3))));
        st3(o, unit(to_world(fr, di.wt))); st3(o + 3, unit(to_world(fr, di.wr))); o[6] = di.F;
    } break;
    case VPT_UNIT_POWER_HEURISTIC: o[0] = power_heuristic((float)a[0], (float)a[1]); break;
    case VPT_UNIT_COSINE_HEMISPHERE: {
        const F3 nrm = ld3(a);
        const F3 w = unit(to_world(make_frame(nrm), cosine_local((float)a[3], (float)a[4])));
        st3(o, w); o[3] = dot(nrm, w) * kInvPi;
    } break;
    case VPT_UNIT_CONE_SAMPLE: {
        const float rr = (float)a[3], dist = (float)a[4];
        const float omc = one_minus_cos_max(rr * rr / (dist * dist));
        st3(o, cone_sample(ld3(a), omc, (float)a[5], (float)a[6]));
        o[3] = 1.0f / (kTwoPi * omc);
    } break;
    case VPT_UNIT_MICROFACET: {
        MatF m{};
        for (int c = 0; c < 3; ++c) { m.eta[c] = (float)a[c]; m.kappa[c] = (float)a[3 + c]; }
        m.alpha = (float)a[6]; m.material = 1;
        const F3 wi = ld3(a + 7), wo = ld3(a + 10);
        const F3 wh = unit(wi + wo);
        st3(o, facet_brdf(m, wi, wh, wo));
        o[3] = facet_pdf(wo, wh, m.alpha); o[4] = beckmann(wh, m.alpha); o[5] = smith_g1(wi, wh, m.alpha) * smith_g1(wo, wh, m.alpha);
    } break;
    case VPT_UNIT_FACET_NORMAL: st3(o, facet_normal((float)a[0], (float)a[1], (float)a[2])); break;
    // ---- the composite functions: ONE record pushed through the product's stage function, contributions captured -------------------
    case VPT_UNIT_MEDIUM_NEE: { // stage_med<POINT / AREA>: (free)SingleScattering = the stage's radiance contribution for throughput T sigma_s (or 1)
        const int sid = (int)a[3];
        const float sigma_s = (float)a[5], T = (float)a[6], pS = (float)a[7];
        ConstsF kk = cf; kk.sigma_t = (float)a[4]; kk.n_emitters = 1.0f / pS; kk.q = 2.0f; // (the roulette after the vertex always ends the path)
        ListCtx c(PS, kk, a + 8, 2);
        r.o = ld3(a); r.src = (uint32_t)sid;
        if (T >= 0.0f) r.beta = mk(T * sigma_s, T * sigma_s, T * sigma_s);
        const bool point = mats[sid].r == 0.0f;
        c.vertex(r, point ? SQ_MED_POINT : SQ_MED_AREA);
        if (point) stage_med<true>(c, true, r); else stage_med<false>(c, true, r);
        st3d(o, c.acc);
    } break;
    case VPT_UNIT_POINT_LIGHT: { // stage_surf_p: bare pLight (no transmittance, no 1 / probSource, no 1 / cp)
        ConstsF kk = cf; kk.sigma_t = 0.0f; kk.n_emitters = 1.0f; kk.inv_cp = 1.0f;
        ListCtx c(PS, kk, a, 0);
        r.hid = (uint32_t)a[0]; r.o = ld3(a + 1); r.d = ld3(a + 7); r.src = (uint32_t)a[10];
        stage_surf_p(c, true, r);
        st3d(o, c.acc);
    } break;
    case VPT_UNIT_SURFACE_MIS: { // stage_surf<LAMBERT / FACET>: MISv2 = the stage's radiance contribution for unit throughput and cp = 1
        ConstsF kk = cf; kk.sigma_t = (float)a[10]; kk.inv_cp = 1.0f; kk.q = 2.0f;
        ListCtx c(PS, kk, a + 11, 8);
        r.hid = (uint32_t)a[0]; r.o = ld3(a + 1); r.d = ld3(a + 7);
        const bool facet = mats[r.hid].material != 0;
        c.vertex(r, facet ? SQ_SURF_F : SQ_SURF_L);
        if (facet) stage_surf<true>(c, true, r); else stage_surf<false>(c, true, r);
        st3d(o, c.acc);
    } break;
    case VPT_UNIT_BSDF_SAMPLE: { // stage_surf's scatter step: bdsf folded with its use (fs cos / pdf) and the new direction, for cp = 1
        ConstsF kk = cf; kk.inv_cp = 1.0f; kk.q = 2.0f;
        ListCtx c(PS, kk, a, 0);
        const MatF &obj = mats[(int)a[0]];
        r.hid = (uint32_t)a[0]; r.o = mk(obj.px, obj.py, obj.pz) + ld3(a + 1) * obj.r; r.d = ld3(a + 4);
        c.set_slot(S_BSDF, (float)a[7]); c.set_slot(S_BSDF + 1, (float)a[8]);
        if (obj.material != 0) stage_surf<true>(c, true, r); else stage_surf<false>(c, true, r);
        st3(o, r.beta); st3(o + 3, r.d);
    } break;
    case VPT_UNIT_RADIANCE: { // a whole path on the Philox stream (pixel, sample) from a given ray: roulette of bounce 0, then the stages
        ThreadCtx c(PS, cf, lp.key0, lp.key1);
        r.o = ld3(a); r.d = ld3(a + 3); r.pixel = (uint32_t)a[6]; r.sample = (uint32_t)a[7];
        if (roulette(c, true, r) == SQ_PRIMARY) trace_path_method(lp.method, c, r);
        st3d(o, c.acc); o[3] = c.events;
    } break;
    case VPT_UNIT_RADIANCE_LIST: { // a whole path on an explicit list of uniforms in the reference's consumption order
        ListCtx c(PS, cf, a + 7, min((int)a[6], 120));
        r.o = ld3(a); r.d = ld3(a + 3);
        if (roulette(c, true, r) == SQ_PRIMARY) trace_path_method(lp.method, c, r);
        st3d(o, c.acc); o[3] = c.overrun ? -1.0 : (double)c.i;
    } break;
    case VPT_UNIT_CAMERA_RAY: { // stage_gen's camera ray for pixel (x, y) (y counted from the bottom, rt.cpp:773) and jitter (xi1, xi2)
        const uint32_t pixel = (uint32_t)((lp.height - 1 - (int)a[1]) * lp.width + (int)a[0]);
        st3(o, camera_dir(cf, pixel, lp.width, lp.height, (float)a[2], (float)a[3]));
    } break;
    case VPT_UNIT_RAYMARCH: {
        double L[3]; unsigned n_steps, scans = 0;
        ray_march3(PS, ld3(a), ld3(a + 3), a[6], (int)a[7], cf.sigma_t, cf.sigma_s, L, n_steps, scans);
        o[0] = L[0]; o[1] = L[1]; o[2] = L[2]; o[3] = n_steps;
    } break;
    default: break;
    }
}

int launch_unit_f32(int fn, const SceneF &scene, const LaunchParams &lp, const ConstsF &cf, int n, const double *in_dev, int in_stride, double *out_dev, int out_stride, void *stream) {
    const int tpb = 64;
    unit_f32_kernel<<<(n + tpb - 1) / tpb, tpb, sizeof(SmScene), (cudaStream_t)stream>>>(fn, scene, lp, cf, n, in_dev, in_stride, out_dev, out_stride);
    return (int)cudaGetLastError();
}

// ---- Philox known-answer kernel -------------------------------------------------------------------------------------------
__global__ void philox_kernel(int n, const uint32_t *__restrict__ ctr, const uint32_t *__restrict__ key, uint32_t *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 r = philox4x32_10(make_uint4(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3]), key[2 * i], key[2 * i + 1]);
    out[4 * i] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
}
int launch_philox(int n, const uint32_t *ctr_dev, const uint32_t *key_dev, uint32_t *out_dev, void *stream) {
    philox_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n, ctr_dev, key_dev, out_dev);
    return (int)cudaGetLastError();
}

// ---- FP32 roofline denominator: register-resident FFMA chains ------------------------------------------
// 16 independent accumulators x 8 FFMA each per loop iteration = kFmaPeakFlopsPerThreadIter flops per thread per iteration.
__global__ void __launch_bounds__(256) fma_peak_kernel(float *__restrict__ sink, int iters) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = (float)(threadIdx.x + i) * 1e-3f;
    const float a = 0.999f + (float)blockIdx.x * 1e-9f, b = 1e-3f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep)
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 12345.678f) sink[0] = s; // never true; keeps the chains alive
}
int launch_fma_peak(float *sink_dev, int n_blocks, int n_threads, int iters, void *stream) {
    fma_peak_kernel<<<n_blocks, n_threads, 0, (cudaStream_t)stream>>>(sink_dev, iters);
    return (int)cudaGetLastError();
}

} // namespace vpt
