// vpt_kernels_f64.cu -- FP64 "REF mode" kernels (compiled with -fmad=false: every product and sum rounds exactly as in the
// reference's strict-IEEE CPU build).  Same thread/pixel/sample structure and the same Philox stream as the FP32
// megakernel (vpt_kernels_f32.cu); the scene lives in shared memory as the reference's own FP64 sphere records.
#include <cuda_runtime.h>
#include "vpt_f64.cuh"
#include "vpt_smwave_f64.cuh"

namespace vpt {

using namespace f64;

__device__ __forceinline__ Ctx make_ctx(const SceneD &sc, const SphereD *shared_spheres, const LaunchParams &lp) {
    Ctx c;
    c.s = shared_spheres;
    c.n_spheres = sc.n_spheres;
    c.n_emitters = sc.n_emitters;
    c.n_volumes = 0;
    for (int i = 0; i < sc.n_spheres; ++i) c.n_volumes += sc.s[i].material == 3;
    c.emitters = sc.emitters;
    c.quirks = lp.quirks;
    c.sigma_a = lp.sigma_a; c.sigma_s = lp.sigma_s; c.sigma_t = lp.sigma_a + lp.sigma_s;
    c.cp = lp.continue_prob; c.q = 1 - lp.continue_prob;
    c.method = lp.method; c.max_depth = lp.max_depth;
    return c;
}
__device__ __forceinline__ void stage_scene(const SceneD &sc, SphereD *dst) {
    for (int i = threadIdx.x; i < sc.n_spheres * (int)(sizeof(SphereD) / 4); i += blockDim.x)
        reinterpret_cast<uint32_t *>(dst)[i] = reinterpret_cast<const uint32_t *>(sc.s)[i];
    __syncthreads();
}
// rt.cpp:787
__device__ __forceinline__ D3 camera_dir(const LaunchParams &lp, int x, int y, double xi1, double xi2) {
    const D3 v = v3(lp.cam_cx) * ((static_cast<double>(x) + xi1 - 0.5) / lp.width - .5) + v3(lp.cam_cy) * ((static_cast<double>(y) + xi2 - 0.5) / lp.height - .5) + v3(lp.cam_d);
    return unit(v);
}

__global__ void __launch_bounds__(kThreadsPerBlock) render_f64_kernel(const __grid_constant__ SceneD sc, const __grid_constant__ LaunchParams lp,
                                                                       float *__restrict__ hdr, Counters *__restrict__ counters) {
    __shared__ SphereD spheres[kMaxSpheres];
    stage_scene(sc, spheres);
    const long long tile = (long long)blockIdx.x * lp.tile_count + lp.tile_rank;
    const long long pixel = tile * kTile + threadIdx.x;
    if (pixel >= lp.n_pixels) return;
    const Ctx c = make_ctx(sc, spheres, lp);
    const int row = (int)(pixel / lp.width), col = (int)(pixel - (long long)row * lp.width);
    const int cam_y = lp.height - 1 - row;

    double acc_r = 0, acc_g = 0, acc_b = 0;
    Tally tally{0u, 0u};
    unsigned nonfinite = 0;
    Path p;
    Rng rng;
    int s = lp.sample_begin;
    bool active = false;
    p.depth = 0;
    if (c.method == VPT_METHOD_RAYMARCH) { // rt.cpp:791: one deterministic march per jittered camera ray
        for (; s < lp.sample_end; ++s) {
            rng.start((uint32_t)pixel, (uint32_t)s, lp.key0, lp.key1);
            double j1, j2;
            rng.jitter_f64(j1, j2);
            const D3 L = ray_march3(c, v3(lp.cam_o), camera_dir(lp, col, cam_y, j1, j2), lp.march_step, lp.march_source, tally);
            ++tally.events;
            if (isfinite(L.x + L.y + L.z)) { acc_r += L.x; acc_g += L.y; acc_b += L.z; } else ++nonfinite;
        }
    }
    if (c.method == VPT_METHOD_VOLUME_SPHERES) { // explicitPathRecursive2 (vptShadeMethods.h:398): one legacy path per jittered camera ray
        for (; s < lp.sample_end; ++s) {
            rng.start((uint32_t)pixel, (uint32_t)s, lp.key0, lp.key1);
            double j1, j2;
            rng.jitter_f64(j1, j2);
            const D3 L = volume_spheres_radiance(c, v3(lp.cam_o), camera_dir(lp, col, cam_y, j1, j2), rng, tally);
            if (isfinite(L.x + L.y + L.z)) { acc_r += L.x; acc_g += L.y; acc_b += L.z; } else ++nonfinite;
        }
    }
    for (;;) {
        bool have = false;
        for (;;) {
            if (!active) {
                if (s >= lp.sample_end) break;
                rng.start((uint32_t)pixel, (uint32_t)s, lp.key0, lp.key1);
                double j1, j2;
                rng.jitter_f64(j1, j2);
                p.o = v3(lp.cam_o); p.d = camera_dir(lp, col, cam_y, j1, j2);
                p.beta = mk(1, 1, 1); p.L = mk(0, 0, 0); p.depth = 0;
                active = true; ++s;
            }
            rng.begin_bounce((uint32_t)p.depth);
            const bool too_deep = (c.max_depth > 0 && p.depth >= c.max_depth) || p.depth >= VPT_MAX_DEPTH;
            if (too_deep || rng.next_f64(S_RR) < c.q) {
                if (isfinite(p.L.x + p.L.y + p.L.z)) { acc_r += p.L.x; acc_g += p.L.y; acc_b += p.L.z; } else ++nonfinite;
                active = false;
                continue;
            }
            have = true;
            break;
        }
        if (!have) break;
        if (vertex(c, p, rng, tally)) {
            ++p.depth;
        } else {
            if (isfinite(p.L.x + p.L.y + p.L.z)) { acc_r += p.L.x; acc_g += p.L.y; acc_b += p.L.z; } else ++nonfinite;
            active = false;
        }
    }
    float *out = hdr + pixel * 3;
    out[0] = (float)(acc_r * lp.out_scale);
    out[1] = (float)(acc_g * lp.out_scale);
    out[2] = (float)(acc_b * lp.out_scale);
    if (!counters) return;
    atomicAdd(&counters->events, (unsigned long long)tally.events);
    atomicAdd(&counters->scans, (unsigned long long)tally.scans);
    if (nonfinite) atomicAdd(&counters->nonfinite, (unsigned long long)nonfinite);
    atomicAdd(&counters->paths, (unsigned long long)(lp.sample_end - lp.sample_begin));
}

// ---- SM-wide wavefront, FP64 reference mode (vpt_smwave_f64.cuh) ------------------------------------------------------------------------
extern __shared__ __align__(16) unsigned char smwave_f64_smem[];
__global__ void __launch_bounds__(kSmdThreads, 1) render_f64_smwave_kernel(const __grid_constant__ SceneD sc, const __grid_constant__ LaunchParams lp,
                                                                            float *__restrict__ hdr, Counters *__restrict__ counters,
                                                                            int log_p, int n_owned_tiles, int n_items, int zero) {
    SmSharedD &M = *reinterpret_cast<SmSharedD *>(smwave_f64_smem);
    const int tid = (int)threadIdx.x;
    for (int i = tid; i < kSmdPool; i += kSmdThreads) M.meta[i] = 0u;
    if (tid == 0) { M.ox[0] = M.oy[0] = M.oz[0] = 0.0; M.dx[0] = M.dy[0] = 0.0; M.dz[0] = 1.0; M.br[0] = M.bg[0] = M.bb[0] = 0.0; M.sample[0] = 0u; }
    stage_scene(sc, M.spheres); // (ends with __syncthreads)
    if (tid == 0) M.ctx = make_ctx(sc, M.spheres, lp);
    __syncthreads();
    const Ctx &c = M.ctx;
    SmWaveD wf(M, c, lp, log_p, n_owned_tiles, zero);
    wf.init(n_items);
    __syncthreads();
    wf.run(hdr, n_items, kSmdFixInv);
    if (!counters) return;
    unsigned long long ev = wf.tally.events, scn = wf.tally.scans, nf = wf.nonfinite, np = wf.paths;
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

// kernel: VPT_KERNEL_MEGA = one thread per pixel (also the ray marcher's kernel), anything else = the SM-wide wavefront
int launch_render_f64(const SceneD &scene, const LaunchParams &lp, float *hdr_dev, Counters *counters_dev, void *stream, int n_blocks, int kernel) {
    cudaStream_t st = (cudaStream_t)stream;
    if (kernel == VPT_KERNEL_MEGA || lp.method == VPT_METHOD_RAYMARCH || lp.method == VPT_METHOD_VOLUME_SPHERES) {
        render_f64_kernel<<<n_blocks, kThreadsPerBlock, 0, st>>>(scene, lp, hdr_dev, counters_dev);
        return (int)cudaGetLastError();
    }
    int dev = 0, n_sm = 0;
    cudaError_t e;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return (int)e;
    if ((e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return (int)e;
    if ((e = cudaFuncSetAttribute(render_f64_smwave_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmSharedD))) != cudaSuccess) return (int)e;
    const int log_p = 7; // work item: one 128-pixel tile
    const int tiles_per_item = 1 << (log_p - 7);
    const int n_items = (n_blocks + tiles_per_item - 1) / tiles_per_item;
    const int grid = n_items < n_sm ? n_items : n_sm;
    render_f64_smwave_kernel<<<grid, kSmdThreads, sizeof(SmSharedD), st>>>(scene, lp, hdr_dev, counters_dev, log_p, n_blocks, n_items, 0);
    return (int)cudaGetLastError();
}

// ---- unit kernels ---------------------------------------------------------------------------------------------------------
struct ListRng {
    const double *u; int i; int n = 1 << 30; bool overrun = false;
    __device__ double next_f64(uint32_t = 0) { if (i >= n) { overrun = true; return 0.0; } /* 0 < q: the next roulette draw ends the path */ return u[i++]; }
    __device__ void begin_bounce(uint32_t) {}
};
__device__ __forceinline__ void st3(double *p, D3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

__global__ void unit_f64_kernel(int fn, const __grid_constant__ SceneD sc, const __grid_constant__ LaunchParams lp, int n, const double *__restrict__ in,
                                int in_stride, double *__restrict__ out, int out_stride) {
    __shared__ SphereD spheres[kMaxSpheres];
    stage_scene(sc, spheres);
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    const double *a = in + (size_t)row * in_stride;
    double *o = out + (size_t)row * out_stride;
    Ctx c = make_ctx(sc, spheres, lp);
    Tally tl{0u, 0u};
    switch (fn) {
    case VPT_UNIT_SPHERE_INTERSECT: o[0] = sphere_t(c.s[(int)a[0]], v3(a + 1), v3(a + 4)); break;
    case VPT_UNIT_INTERSECT: {
        double t = 0; int id = 0;
        const bool h = scan(c, v3(a), v3(a + 3), t, id, tl);
        o[0] = h; o[1] = t; o[2] = id;
    } break;
    case VPT_UNIT_VISIBILITY: o[0] = visible(c, v3(a), v3(a + 3), tl); break;
    case VPT_UNIT_TRANSMITTANCE: o[0] = transmittance(v3(a), v3(a + 3), a[6]); break;
    case VPT_UNIT_FREE_FLIGHT: {
        const double st = a[0], xi = a[1];
        const double d = -log(1 - xi) / st;
        o[0] = d; o[1] = st * exp(st * d * -1.0); o[2] = 1.0 - exp(-st * d); o[3] = exp(-st * d);
    } break;
    case VPT_UNIT_PHASE_SAMPLE: st3(o, phase_sample(a[0], a[1])); break;
    case VPT_UNIT_EQUIANGULAR: {
        const D3 org = v3(a + 2), dir = v3(a + 5);
        const D3 dv = pos(c.s[(int)a[0]]) - org;
        const double len = sqrt(dot(dv, dv));
        const double proj = dot(dv, dir) / dot(dir, dir);
        const double D = sqrt(len * len - proj * proj);
        const double thA = atan2(0.0 - proj, D), thB = atan2(a[1] - proj, D);
        const double tl_ = D * tan((1 - a[8]) * thA + a[8] * thB);
        o[0] = D; o[1] = thA; o[2] = thB; o[3] = tl_; o[4] = tl_ + proj; o[5] = D / fabs(thB - thA) / (tl_ * tl_ + D * D);
    } break;
    case VPT_UNIT_MIS_DISTANCE: {
        double dist, pdf;
        const bool surface = mis_distance(pos(c.s[(int)a[0]]), v3(a + 2), v3(a + 5), a[1], a[8], a[9], a[10], dist, pdf);
        o[0] = surface; o[1] = dist; o[2] = pdf;
    } break;
    case VPT_UNIT_DIELECTRIC: {
        const D3 n = v3(a), wo = v3(a + 3);
        const D3 wt = unit(refract_dielectric(1.0, 1.5, wo, n));
        st3(o, wt); st3(o + 3, unit(reflect_dielectric(wo, n))); o[6] = fresnel_dielectric(1.0, 1.5, dot(n, wt), dot(n, wo));
    } break;
    case VPT_UNIT_POWER_HEURISTIC: o[0] = power_heuristic(a[0], a[1]); break;
    case VPT_UNIT_COSINE_HEMISPHERE: {
        const D3 w = cosine_hemisphere(v3(a), a[3], a[4]);
        st3(o, w); o[3] = cosine_pdf(dot(v3(a), w));
    } break;
    case VPT_UNIT_CONE_SAMPLE: {
        const double cm = sqrt(1 - (a[3] / a[4]) * (a[3] / a[4]));
        st3(o, cone_sample(v3(a), cm, a[5], a[6])); o[3] = cone_pdf(cm);
    } break;
    case VPT_UNIT_MICROFACET: {
        SphereD m{};
        for (int k = 0; k < 3; ++k) { m.eta[k] = a[k]; m.kappa[k] = a[3 + k]; }
        const D3 wi = v3(a + 7), wo = v3(a + 10), nl = mk(0, 0, 1);
        const D3 wh = unit(wi + wo);
        st3(o, facet_brdf(m, wi, wh, wo, a[6], nl));
        o[3] = facet_pdf(wo, wh, a[6], nl); o[4] = beckmann(dot(nl, wh), a[6]); o[5] = smith_g1(nl, wi, wh, a[6]) * smith_g1(nl, wo, wh, a[6]);
    } break;
    case VPT_UNIT_FACET_NORMAL: st3(o, facet_normal(a[0], a[1], a[2])); break;
    case VPT_UNIT_MEDIUM_NEE: {
        c.sigma_t = a[4]; c.sigma_s = a[5];
        ListRng lr{a + 8, 0};
        st3(o, medium_direct(c, v3(a), (int)a[3], a[7], a[6] >= 0, a[6], lr, tl));
    } break;
    case VPT_UNIT_POINT_LIGHT: {
        const SphereD &obj = c.s[(int)a[0]], &src = c.s[(int)a[10]];
        st3(o, point_light_direct(c, obj, v3(a + 1), v3(a + 4), v3(a + 7), rad(src), pos(src), obj.alpha, tl));
    } break;
    case VPT_UNIT_SURFACE_MIS: {
        const SphereD &obj = c.s[(int)a[0]];
        c.sigma_t = a[10];
        ListRng lr{a + 11, 0};
        st3(o, surface_direct_mis(c, obj, v3(a + 1), v3(a + 4), v3(a + 7), obj.alpha, lr, tl));
    } break;
    case VPT_UNIT_BSDF_SAMPLE: {
        const SphereD &obj = c.s[(int)a[0]];
        const D3 nrm = v3(a + 1);
        ListRng lr{a + 7, 0};
        D3 wi; double pdf;
        const D3 fs = bsdf_sample(obj, wi, v3(a + 4), nrm, pdf, lr);
        wi = unit(wi);
        st3(o, fs * dot(nrm, wi) * (1 / pdf)); st3(o + 3, wi);
    } break;
    case VPT_UNIT_RADIANCE: {
        Path p; p.o = v3(a); p.d = v3(a + 3); p.beta = mk(1, 1, 1); p.L = mk(0, 0, 0); p.depth = 0;
        Rng rng; rng.start((uint32_t)a[6], (uint32_t)a[7], lp.key0, lp.key1);
        if (c.method == VPT_METHOD_VOLUME_SPHERES) { st3(o, volume_spheres_radiance(c, p.o, p.d, rng, tl)); o[3] = tl.events; break; }
        for (;;) {
            rng.begin_bounce((uint32_t)p.depth);
            if ((c.max_depth > 0 && p.depth >= c.max_depth) || p.depth >= VPT_MAX_DEPTH || rng.next_f64(S_RR) < c.q) break;
            if (!vertex(c, p, rng, tl)) break;
            ++p.depth;
        }
        st3(o, p.L); o[3] = tl.events;
    } break;
    case VPT_UNIT_RADIANCE_LIST: {
        Path p; p.o = v3(a); p.d = v3(a + 3); p.beta = mk(1, 1, 1); p.L = mk(0, 0, 0); p.depth = 0;
        ListRng rng{a + 7, 0, min((int)a[6], 120)};
        if (c.method == VPT_METHOD_VOLUME_SPHERES) { st3(o, volume_spheres_radiance(c, p.o, p.d, rng, tl)); o[3] = rng.overrun ? -1.0 : (double)rng.i; break; }
        for (;;) {
            if ((c.max_depth > 0 && p.depth >= c.max_depth) || rng.next_f64(S_RR) < c.q) break;
            if (!vertex(c, p, rng, tl)) break;
            ++p.depth;
        }
        st3(o, p.L); o[3] = rng.overrun ? -1.0 : (double)rng.i;
    } break;
    case VPT_UNIT_CAMERA_RAY: st3(o, camera_dir(lp, (int)a[0], (int)a[1], a[2], a[3])); break;
    case VPT_UNIT_RAYMARCH: st3(o, ray_march3(c, v3(a), v3(a + 3), a[6], (int)a[7], tl, o + 3)); break;
    default: break;
    }
}

int launch_unit_f64(int fn, const SceneD &scene, const LaunchParams &lp, int n, const double *in_dev, int in_stride, double *out_dev, int out_stride, void *stream) {
    const int tpb = 64;
    unit_f64_kernel<<<(n + tpb - 1) / tpb, tpb, 0, (cudaStream_t)stream>>>(fn, scene, lp, n, in_dev, in_stride, out_dev, out_stride);
    return (int)cudaGetLastError();
}

} // namespace vpt
