// vpt_march.cuh -- FP32 ray-marching reference solver (VPT_METHOD_RAYMARCH): rayMarching3, rayMarchingMethods.h:330-384, the commented
// line rt.cpp:791.  Constant-step Riemann sum of the single scattering from the centre of ONE source along the camera ray up to the
// first surface; deterministic except for the pixel jitter.  One thread per pixel; every step is one scene scan (the product
// kernel's scan over the float4 records in shared memory, vpt_smwave.cuh), so a frame costs about t / step = 3000 scans per sample.
// FP32 semantics as everywhere in this precision: r == 0 spheres are not ray-intersected, visibility = nothing hit before
// distance * (1 - 1e-4).  The per-step terms are fp32, their sum is accumulated in double (3000 terms).
#pragma once
#include "vpt_scan.cuh"

namespace vpt {
namespace f32 {

// L[3] and the number of steps.  (The reference attenuates each sample by the transmittance from the SURFACE point x to the sample,
// rayMarchingMethods.h:350, not from the ray origin; reproduced as written.)
__device__ __forceinline__ void ray_march3(const SmScene &S, F3 o, F3 d, double step_d, int source, float sigma_t, float sigma_s, double L[3], unsigned &n_steps,
                                           unsigned &n_scans) {
    L[0] = L[1] = L[2] = 0.0;
    n_steps = 0;
    float t; int id;
    ++n_scans;
    if (!scan_sm(S, o, d, t, id)) return;
    const MatF &src = S.mats[source];
    const F3 light = mk(src.px, src.py, src.pz);
    const float step = (float)step_d;
    const double steps = (double)t / step_d; // the loop bound in double: one step more or less is 3e-4 of the sum
    const float scale = kInv4Pi * sigma_s * step;
    double acc = 0.0; // the three channels share everything but the source's radiance
    unsigned i = 0;
    for (; (double)i < steps; ++i) {
        const float ti = step * (float)i;
        const F3 xt = fma3(d, ti, o);
        const F3 lx = light - xt;
        const float d2 = dot(lx, lx), inv = rsqrtf(d2), dist = d2 * inv;
        float th; int hid;
        ++n_scans;
        const bool hit = scan_sm(S, light, lx * (-inv), th, hid);
        if (!hit || th > dist * (1.0f - 1e-4f)) // visibility (pathTracingUtilities.h:39-53)
            acc += (double)(expf(-sigma_t * ((t - ti) + dist)) * inv * inv * scale); // T(x, xt) * T(xt, light) / |wc|^2 * phase * sigma_s * step
    }
    n_steps = i;
    L[0] = acc * (double)src.lr; L[1] = acc * (double)src.lg; L[2] = acc * (double)src.lb;
}

} // namespace f32
} // namespace vpt
