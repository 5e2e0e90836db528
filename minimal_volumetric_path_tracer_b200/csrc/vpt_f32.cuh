// vpt_f32.cuh -- FP32 building blocks of the performance path: vector algebra, local frames, the reference's sampling routines and pdfs,
// the Beckmann conductor model, the dielectric as the reference writes it, the equi-angular and MIS distance samplers.
//
// Written for the GPU from the reference's behaviour (file:line cited per function), not translated: angles are never
// materialised where the reference goes acos -> sin/cos (algebraic forms instead).  The scene scan lives in vpt_scan.cuh, the estimators
// built from these blocks -- ONE copy for every FP32 kernel and for the unit kernels -- in vpt_stages.cuh.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include "vpt_internal.h"
#include "vpt_philox.cuh"

namespace vpt {
namespace f32 {

constexpr float kPi = 3.14159265358979323846f;
constexpr float kInvPi = 0.31830988618379067154f;
constexpr float kInv4Pi = 0.07957747154594766788f;
constexpr float kTwoPi = 6.28318530717958647692f;
constexpr float kMaxFloat = 3.402823466e+38f; // MAXFLOAT, vptShadeMethods.h:1287
constexpr float kEps = 1e-4f;                 // Sphere.h:34, pathTracingUtilities.h:20

struct F3 { float x, y, z; };
__device__ __forceinline__ F3 mk(float x, float y, float z) { return F3{x, y, z}; }
__device__ __forceinline__ F3 operator+(F3 a, F3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ F3 operator-(F3 a, F3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ F3 operator*(F3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ F3 operator-(F3 a) { return mk(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot(F3 a, F3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ F3 had(F3 a, F3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ F3 cross(F3 a, F3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ F3 fma3(F3 a, float s, F3 b) { return mk(fmaf(a.x, s, b.x), fmaf(a.y, s, b.y), fmaf(a.z, s, b.z)); } // a*s + b
__device__ __forceinline__ F3 unit(F3 a) { return a * rsqrtf(dot(a, a)); }

// coordinateSystem (mathUtilities.h:10-19): t from n, s = t x n
struct Frame { F3 s, t, n; };
__device__ __forceinline__ Frame make_frame(F3 n) {
    Frame f; f.n = n;
    if (fabsf(n.x) > fabsf(n.y)) { const float inv = rsqrtf(n.x * n.x + n.z * n.z); f.t = mk(n.z * inv, 0.0f, -n.x * inv); }
    else { const float inv = rsqrtf(n.y * n.y + n.z * n.z); f.t = mk(0.0f, n.z * inv, -n.y * inv); }
    f.s = cross(f.t, n);
    return f;
}
__device__ __forceinline__ F3 to_local(const Frame &f, F3 w) { return mk(dot(f.s, w), dot(f.t, w), dot(f.n, w)); } // coordinateTraspose :21-30
__device__ __forceinline__ F3 to_world(const Frame &f, F3 l) { return fma3(f.s, l.x, fma3(f.t, l.y, f.n * l.z)); }

// sin / cos of 2 pi x for x in (0, 1): MUFU.SIN / MUFU.COS after an exact range reduction to (-pi, pi] (absolute error ~5e-7)
__device__ __forceinline__ void fast_sincos2pi(float x, float &s, float &c) {
    const float a = (x - rintf(x)) * kTwoPi; // x - round(x) is exact
    s = __sinf(a); c = __cosf(a);
}

// ---- sampling ------------------------------------------------------------------------------------------------------
// isotropicPhaseSample (vptSamplingFunctions.h:34-46): cos = 1 - 2 xi1, sin = 2 sqrt(xi1 (1 - xi1)), phi = 2 pi xi2
__device__ __forceinline__ F3 phase_sample(float xi1, float xi2) {
    const float c = 1.0f - 2.0f * xi1, s = 2.0f * sqrtf(xi1 * (1.0f - xi1));
    float sp, cp; fast_sincos2pi(xi2, sp, cp);
    return mk(s * cp, s * sp, c);
}
// cosineHemispheric (samplingFunctions.h:47-62): cos = sqrt(1 - xi1), sin = sqrt(xi1); local direction
__device__ __forceinline__ F3 cosine_local(float xi1, float xi2) {
    const float c = sqrtf(1.0f - xi1), s = sqrtf(xi1);
    float sp, cp; fast_sincos2pi(xi2, sp, cp);
    return mk(s * cp, s * sp, c);
}
// uniform cone: solidAngle (samplingFunctions.h:65-82) with 1 - cos(theta_max) given exactly (omc_max), so that the
// small light cones do not lose their digits in fp32.  cos = 1 - xi1 * omc_max.
__device__ __forceinline__ F3 cone_sample(F3 axis, float omc_max, float xi1, float xi2) {
    const float omc = xi1 * omc_max;
    const float c = 1.0f - omc, s = sqrtf(omc * (2.0f - omc));
    float sp, cp; fast_sincos2pi(xi2, sp, cp);
    const Frame f = make_frame(axis);
    return unit(to_world(f, mk(s * cp, s * sp, c)));
}
// 1 - sqrt(1 - s2) without cancellation; s2 = (r / distance)^2
__device__ __forceinline__ float one_minus_cos_max(float s2) { return s2 / (1.0f + sqrtf(1.0f - s2)); }
// powerHeuristics (misSamplingFunctions.h:12-16)
__device__ __forceinline__ float power_heuristic(float f, float g) { const float f2 = f * f, g2 = g * g; return f2 / (f2 + g2); }

// equiAngularParams2 (volumetricBasicFunctions.h:209-223) + equiAngularSample (vptSamplingFunctions.h:54) without materialising the
// two angles: with a = -proj, b = tmax - proj (the segment ends relative to the light's projection, b > a) and D the light's distance
// from the ray,  theta_B - theta_A = atan2((b - a) D, D^2 + a b)  (both angles lie in (-pi/2, pi/2]) and
// tan(theta_A + xi (theta_B - theta_A)) follows from the addition theorem with tan(theta_A) = a / D: one atan2f and one tanf instead
// of two atan2f and a tanf.  Returns the sampled distance along the ray; t_local is measured from the projection point.
__device__ __forceinline__ float equiangular_sample(F3 light, F3 o, F3 d, float tmax, float xi, float &D, float &dtheta, float &t_local) {
    const F3 dv = light - o;
    const float proj = dot(dv, d);
    const F3 perp = fma3(d, -proj, dv);
    D = sqrtf(dot(perp, perp));
    const float a = -proj, b = fminf(tmax, 1e18f) - proj; // (a miss has tmax = MAXFLOAT: keep the products finite, theta_B is pi/2 to 1e-15)
    dtheta = atan2f((b - a) * D, fmaf(a, b, D * D));
    const float tau = tanf(xi * dtheta);
    t_local = D * fmaf(D, tau, a) / fmaf(-a, tau, D);
    return __fadd_rn(t_local, proj); // (explicit roundings wherever a product meets a sum: every kernel that inlines this code must round alike)
}

// VPT_METHOD_MIS_DISTANCE (not in the reference, whose "MIS" method :1345 is the equi-angular estimator again):
// one-sample MIS of the reference's two distance techniques with the balance heuristic.  Both end at the surface with probability
// Tr = exp(-sigma_t t) and otherwise place a medium vertex on [0, t): free flight with density sigma_t exp(-sigma_t s)
// (freeFlightProb, vptSamplingFunctions.h:20), equi-angular with equiAngularProb(s) (1 - Tr) (:60, vptShadeMethods.h:1093).  xd < Tr: surface
// (returns true); xd < (1 + Tr) / 2: free flight restricted to [0, t), s = -log(1 - xi (1 - Tr)) / sigma_t; else the equi-angular sample.
// inv_pdf = 1 / mixture density = 2 / (sigma_t exp(-sigma_t s) + equiAngularProb(s) (1 - Tr)); the medium vertex then goes through method
// 1's code.  Restated in FP64 in oracle/vpt_oracle.hpp::mis_distance.
__device__ __forceinline__ bool mis_distance(F3 light, F3 o, F3 d, float tmax, float Tr, float sigma_t, float inv_sigma_t, float xi, float xd, float &dist, float &inv_pdf) {
    dist = 0.0f; inv_pdf = 1.0f;
    if (xd < Tr) return true;
    const F3 dv = light - o;
    const float proj = dot(dv, d);
    const F3 perp = fma3(d, -proj, dv);
    const float D2 = dot(perp, perp), D = sqrtf(D2);
    const float a = -proj, b = fminf(tmax, 1e18f) - proj; // as equiangular_sample
    const float dtheta = atan2f((b - a) * D, fmaf(a, b, D2));
    float tl;
    if (xd < fmaf(0.5f, Tr, 0.5f)) {
        dist = -logf(fmaf(-xi, 1.0f - Tr, 1.0f)) * inv_sigma_t;
        tl = dist - proj;
    } else {
        const float tau = tanf(xi * dtheta);
        tl = D * fmaf(D, tau, a) / fmaf(-a, tau, D);
        dist = __fadd_rn(tl, proj);
    }
    const float p_equi = D * (1.0f - Tr) / (dtheta * fmaf(tl, tl, D2));
    inv_pdf = 2.0f / fmaf(sigma_t, expf(-sigma_t * dist), p_equi); // p_free + p_equi
    return false;
}

// ---- Beckmann conductor microfacet model (microFacetUtilities.h), local frame n = +z ----------------------------------
__device__ __forceinline__ float fresnel_channel(float c, float s2, float eta, float kappa) { // fresnelSpectre :11-18 (s2 = sin^2)
    const float e2k2 = eta * eta - kappa * kappa - s2;
    const float a2b2 = sqrtf(fmaf(e2k2, e2k2, 4.0f * eta * eta * kappa * kappa));
    const float a = sqrtf(0.5f * (a2b2 + e2k2));
    const float c2 = c * c, two_ac = 2.0f * a * c;
    const float perp = (a2b2 + c2 - two_ac) / (a2b2 + c2 + two_ac);
    const float s4 = s2 * s2, x = a2b2 * c2 + s4, y = two_ac * s2;
    const float par = perp * (x - y) / (x + y);
    return 0.5f * (par + perp);
}
__device__ __forceinline__ F3 fresnel_conductor(float cos_h, const float *eta, const float *kappa) { // fresnel :21-29
    const float s2 = fmaxf(1.0f - cos_h * cos_h, 0.0f);
    return mk(fresnel_channel(cos_h, s2, eta[0], kappa[0]), fresnel_channel(cos_h, s2, eta[1], kappa[1]), fresnel_channel(cos_h, s2, eta[2], kappa[2]));
}
__device__ __forceinline__ float beckmann(F3 wh, float alpha) { // NDF :34-45 with cos = wh.z, sin^2 = x^2 + y^2
    if (!(wh.z >= 0.0f)) return 0.0f;
    const float c2 = wh.z * wh.z, s2 = fmaf(wh.x, wh.x, wh.y * wh.y), a2 = alpha * alpha;
    return expf(-s2 / (c2 * a2)) / (kPi * a2 * c2 * c2);
}
__device__ __forceinline__ float smith_g1(F3 wv, F3 wh, float alpha) { // Gn :47-61
    const float c = wv.z;
    const float s = sqrtf(fmaf(wv.x, wv.x, wv.y * wv.y));
    const float a = c / (alpha * s);
    if (dot(wv, wh) / c > 0.0f) {
        if (a < 1.6f) return (3.535f * a + 2.181f * a * a) / (1.0f + 2.276f * a + 2.577f * a * a);
        return 1.0f;
    }
    return 0.0f;
}
__device__ __forceinline__ float facet_pdf(F3 wo, F3 wh, float alpha) { // microFacetProb :86-92
    return beckmann(wh, alpha) * wh.z / (4.0f * fabsf(dot(wo, wh)));
}
__device__ __forceinline__ F3 facet_brdf(const MatF &m, F3 wi, F3 wh, F3 wo) { // frMicroFacet :95-100
    const float g = smith_g1(wi, wh, m.alpha) * smith_g1(wo, wh, m.alpha);
    const float k = beckmann(wh, m.alpha) * g / (4.0f * fabsf(wi.z) * fabsf(wo.z));
    return fresnel_conductor(dot(wi, wh), m.eta, m.kappa) * k;
}
// vectorFacet :71-84: tan^2 = -alpha^2 log(1 - xi1)
__device__ __forceinline__ F3 facet_normal(float alpha, float xi1, float xi2) {
    const float t2 = -alpha * alpha * logf(1.0f - xi1);
    const float c = rsqrtf(1.0f + t2), s = sqrtf(t2) * c;
    float sp, cp; fast_sincos2pi(xi2, sp, cp);
    return mk(s * cp, s * sp, c);
}

// ---- material 2 (dielectric) AS WRITTEN in the reference, local frame n = +z -----------------------------------------------------------
// refraxDielectric (microFacetUtilities.h:122-141): local x, y scaled by -etat/etai = -1.5, z = sqrt(1 - (1/1.5)^2 (1 - cos_i^2)) - 1, then
// normalised by the callers; reflexDielectric (:117-120); fresnelDie(1, 1.5, n.wt, n.wo) (:107-112).  Not Snell's law -- the reference's
// formulas, restated so that a scene using material 2 renders what the reference renders (FP64 statement: vpt_f64.cuh, oracle/vpt_oracle.hpp).
// fp32 care: 1 - cos_i^2 = x^2 + y^2 for the unit wo, and sqrt(1 - a) - 1 = -a / (1 + sqrt(1 - a)).
struct DielF { F3 wr, wt; float F; };
__device__ __forceinline__ DielF dielectric_setup(F3 wo_l) {
    DielF r;
    const float a = (wo_l.x * wo_l.x + wo_l.y * wo_l.y) * (1.0f / 2.25f);
    const float ct = -a / (1.0f + sqrtf(1.0f - a));
    r.wt = unit(mk(-1.5f * wo_l.x, -1.5f * wo_l.y, ct));
    r.wr = mk(-wo_l.x, -wo_l.y, wo_l.z);
    const float ci = wo_l.z, cn = r.wt.z;
    const float par = (1.5f * ci - cn) / (1.5f * ci + cn), perp = (ci - 1.5f * cn) / (ci + 1.5f * cn);
    r.F = 0.5f * (par * par + perp * perp);
    return r;
}
// the BSDF-sampled term of MISv2 for a dielectric (softDielectric, samplingFunctions.h:209-235; misSamplingFunctions.h:144-152), after the
// scan along the chosen direction: Le / |n.w| (x 1.5^2 when refracted), weighted by powerHeuristics(gpdf left by the light loop, cone pdf)
__device__ __forceinline__ F3 dielectric_direct(const MatF &em, F3 x, float cos_w, bool refracted, float gpdf_loop) {
    const F3 g = mk(em.lr, em.lg, em.lb) * ((refracted ? 2.25f : 1.0f) / fabsf(cos_w));
    if (!(g.x > 0.0f && g.y > 0.0f && g.z > 0.0f)) return mk(0, 0, 0);
    const F3 cx = mk(em.px, em.py, em.pz) - x;
    const float omc = one_minus_cos_max(em.r * em.r / dot(cx, cx));
    return g * power_heuristic(gpdf_loop, 1.0f / (kTwoPi * omc));
}

// Surface BRDF value for an incoming local direction wi (Lambert c/pi or microfacet)
__device__ __forceinline__ F3 brdf_eval(const MatF &obj, F3 wi_l, F3 wo_l) {
    if (obj.material == 1) return facet_brdf(obj, wi_l, unit(wi_l + wo_l), wo_l);
    return mk(obj.cr, obj.cg, obj.cb) * kInvPi;
}

// bdsf (vptShadeMethods.h:16-59) folded with its use at :1323-1327: returns fs * cos / pdf and the unit direction wi.
__device__ __forceinline__ F3 bsdf_sample(const MatF &obj, const Frame &fr, F3 wo_l, float xi1, float xi2, F3 &wi) {
    if (obj.material == 2) { // :26-46: fs cos / pdf = (F / n.wi) (n.wi) / F = 1 reflected, (1 - F) 1.5^2 / (1 - F) refracted; xi2 unused
        const DielF di = dielectric_setup(wo_l);
        const bool reflected = xi1 < di.F;
        wi = unit(to_world(fr, reflected ? di.wr : di.wt));
        const float w = reflected ? 1.0f : 2.25f;
        return mk(w, w, w);
    }
    if (obj.material == 1) {
        const F3 wh = facet_normal(obj.alpha, xi1, xi2);
        const F3 wi_l = unit(fma3(wh, 2.0f * dot(wh, wo_l), -wo_l));
        const F3 fs = facet_brdf(obj, wi_l, wh, wo_l);
        const float pdf = facet_pdf(wo_l, wh, obj.alpha);
        wi = unit(to_world(fr, wi_l));
        return fs * (wi_l.z / pdf);
    }
    wi = unit(to_world(fr, cosine_local(xi1, xi2)));
    return mk(obj.cr, obj.cg, obj.cb); // c/pi * cos / (cos/pi)
}

} // namespace f32
} // namespace vpt
