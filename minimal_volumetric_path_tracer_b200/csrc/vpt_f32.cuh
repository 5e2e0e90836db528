// vpt_f32.cuh -- FP32 device implementation of the reference's per-path radiance estimators (the performance path).
//
// Written for the GPU from the reference's behaviour (file:line cited per function), not translated: geometry uses a
// re-anchored, cancellation-free ray/sphere form so the r = 1e5 wall spheres survive fp32; angles are never
// materialised where the reference goes acos -> sin/cos (algebraic forms instead); local frames are built once per
// surface event.  Semantics = the reference with its two FP64-rounding-decided behaviours replaced by their
// well-defined alternative (include/vpt.h VPT_QUIRK_*): r == 0 spheres are never ray-intersected and visibility uses
// `t > distance * (1 - 1e-4)`.  The draw order of random numbers is the reference's, so that the
// FP64 CPU oracle consumes the identical stream.
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

// ---- Sphere::intersect (Sphere.h:27-37) in the re-anchored form -------------------------------------------------------
// returns the reference's choice of root: the near one unless it is negative or within 1e-4 of the origin, else the far
// one (which may be negative); 0 when the ray misses.
__device__ __forceinline__ float sphere_t(const GeomF &g, F3 o, F3 d) {
    const F3 oq = mk(o.x - g.qx, o.y - g.qy, o.z - g.qz);
    const F3 m = mk(g.mx, g.my, g.mz);
    const float c = fmaf(2.0f, dot(oq, m), dot(oq, oq)) + g.c0;
    const F3 op = oq + m;
    const float b = dot(op, d);
    float det;
    if (g.big) {
        det = fmaf(b, b, -c);
    } else { // r^2 - |op - (op.d) d|^2 : no cancellation for small far-away spheres
        const F3 l = fma3(d, -b, op);
        det = g.r2 - dot(l, l);
    }
    if (det < 0.0f) return 0.0f;
    const float sq = sqrtf(det);
    const float q = -(b + copysignf(sq, b)); // the root without cancellation; the other one is c / q
    const float other = c / q;
    const float t_near = fminf(q, other), t_far = fmaxf(q, other);
    return (t_near < 0.0f || fabsf(t_near) < kEps) ? t_far : t_near;
}

// intersect (pathTracingUtilities.h:12-36).  Returns hit; t and id (caller's sphere index) only written on a hit.
__device__ __forceinline__ bool scan(const SceneF &sc, F3 o, F3 d, float &t, int &id, unsigned &n_scans) {
    float best = CUDART_INF_F;
    int best_id = -1;
    ++n_scans;
    for (int i = 0; i < sc.n_geom; ++i) {
        const float ti = sphere_t(sc.geom[i], o, d);
        if (ti > kEps && ti < best) { best = ti; best_id = sc.geom[i].id; }
    }
    if (best_id < 0) return false;
    t = best; id = best_id;
    return true;
}

// visibility (pathTracingUtilities.h:39-53), well-defined variant: nothing hit before distance * (1 - 1e-4)
__device__ __forceinline__ bool visible(const SceneF &sc, F3 light, F3 x, unsigned &n_scans) {
    const F3 lx = light - x;
    const float d2 = dot(lx, lx);
    const float inv = rsqrtf(d2);
    float t; int id;
    if (!scan(sc, light, lx * (-inv), t, id, n_scans)) return true;
    return t > d2 * inv * (1.0f - 1e-4f);
}

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
    return t_local + proj;
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
        dist = -logf(1.0f - xi * (1.0f - Tr)) * inv_sigma_t;
        tl = dist - proj;
    } else {
        const float tau = tanf(xi * dtheta);
        tl = D * fmaf(D, tau, a) / fmaf(-a, tau, D);
        dist = tl + proj;
    }
    const float p_free = sigma_t * expf(-sigma_t * dist), p_equi = D * (1.0f - Tr) / (dtheta * fmaf(tl, tl, D2));
    inv_pdf = 2.0f / (p_free + p_equi);
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

// ---- per-path state -----------------------------------------------------------------------------------------------------
struct Consts { // derived once per launch from LaunchParams
    float sigma_t, inv_sigma_t, sigma_s, albedo_over_cp, inv_cp, q;
    float n_emitters; // 1 / probSource
    int method, max_depth;
};
struct Path { F3 o, d, beta, L; int depth; };
struct Tally { unsigned events, scans; };

// Surface BRDF value for an incoming local direction wi (Lambert c/pi or microfacet)
__device__ __forceinline__ F3 brdf_eval(const MatF &obj, F3 wi_l, F3 wo_l) {
    if (obj.material == 1) return facet_brdf(obj, wi_l, unit(wi_l + wo_l), wo_l);
    return mk(obj.cr, obj.cg, obj.cb) * kInvPi;
}

// pLight (vptShadeMethods.h:62-91) * transmittance * 1/probSource, as used at :1316 / :1113 / :1444.
// For an area source the reference's visibility ray starts at the sphere centre and hits that sphere at t = r, so the
// term is zero whenever the shaded point lies outside the source sphere.
__device__ __forceinline__ F3 point_light_direct(const SceneF &sc, const MatF &obj, const MatF &src, F3 x, const Frame &fr, F3 wo_l, const Consts &k, unsigned &n_scans) {
    const F3 light = mk(src.px, src.py, src.pz);
    const F3 lx = light - x;
    const float d2 = dot(lx, lx);
    if (src.r > 0.0f && d2 > src.r * src.r) return mk(0, 0, 0);
    if (!visible(sc, light, x, n_scans)) return mk(0, 0, 0);
    const float inv = rsqrtf(d2), dist = d2 * inv;
    const F3 wi = lx * inv;
    const F3 wi_l = unit(to_local(fr, wi));
    const F3 f = brdf_eval(obj, wi_l, wo_l);
    const float scale = dot(fr.n, wi) / d2 * expf(-k.sigma_t * dist) * k.n_emitters;
    return had(mk(src.lr, src.lg, src.lb), f) * scale;
}

// MISv2 (misSamplingFunctions.h:96-170), materials 0 and 1
template <class RngT>
__device__ __forceinline__ F3 surface_direct_mis(const SceneF &sc, const MatF *mats, const MatF &obj, F3 x, const Frame &fr, F3 wo_l,
                                                 const Consts &k, RngT &rng, unsigned &n_scans) {
    F3 total = mk(0, 0, 0);
    if (obj.material == 2) { // light-sampled terms are zero (fr = 0, samplingFunctions.h:190); the loop only leaves its last gpdf behind (:110-118)
        const DielF di = dielectric_setup(wo_l);
        float gpdf_loop = 0.0f;
        for (int a = 0; a < sc.n_area; ++a) {
            rng.next_f32(S_AREA + 2 * a); rng.next_f32(S_AREA + 2 * a + 1);
            gpdf_loop = rng.next_f32(S_DIEL + a) > di.F ? 1.0f - di.F : di.F;
        }
        const bool refracted = !(rng.next_f32(S_MIS) < di.F);
        const F3 w_l = refracted ? di.wt : di.wr;
        float t; int hit_id;
        if (scan(sc, x, unit(to_world(fr, w_l)), t, hit_id, n_scans)) total = dielectric_direct(mats[hit_id], x, w_l.z, refracted, gpdf_loop);
        return total;
    }
    float omc_last = 1.0f; // 1 - costhetaMax of the last light visited (reference: stale variable, :162); 1 = "cos 0"
    for (int a = 0; a < sc.n_area; ++a) {
        const int lid = sc.area[a];
        const MatF &src = mats[lid];
        const float xi1 = rng.next_f32(S_AREA + 2 * a), xi2 = rng.next_f32(S_AREA + 2 * a + 1);
        const F3 cx = mk(src.px, src.py, src.pz) - x;
        const float len2 = dot(cx, cx), inv_len = rsqrtf(len2);
        const float omc_max = one_minus_cos_max(src.r * src.r / len2);
        omc_last = omc_max;
        const F3 wi = cone_sample(cx * inv_len, omc_max, xi1, xi2);
        float t; int hit_id = 0;
        scan(sc, x, wi, t, hit_id, n_scans);
        if (hit_id != lid) continue; // Le = 0 (samplingFunctions.h:199-200)
        const F3 wi_l = unit(to_local(fr, wi));
        const float cos_i = dot(fr.n, wi);
        const float inv_fpdf = kTwoPi * omc_max, fpdf = 1.0f / inv_fpdf;
        F3 f; float gpdf;
        if (obj.material == 1) {
            const F3 wh = unit(wi_l + wo_l);
            f = facet_brdf(obj, wi_l, wh, wo_l);
            gpdf = facet_pdf(wo_l, wh, obj.alpha);
        } else {
            f = mk(obj.cr, obj.cg, obj.cb) * kInvPi;
            gpdf = cos_i * kInvPi;
        }
        const float Tr = expf(-k.sigma_t * len2 * inv_len);
        const float w = power_heuristic(fpdf, gpdf);
        total = total + had(mk(src.lr, src.lg, src.lb), f) * (cos_i * inv_fpdf * Tr * w);
    }
    // one BSDF sample (uniform :250 / microfacet :97)
    const float xi1 = rng.next_f32(S_MIS), xi2 = rng.next_f32(S_MIS + 1);
    if (obj.material == 1) {
        const F3 wh = facet_normal(obj.alpha, xi1, xi2);
        const F3 wi_l = unit(fma3(wh, 2.0f * dot(wh, wo_l), -wo_l));
        const F3 wi = unit(to_world(fr, wi_l));
        float t; int hit_id;
        if (scan(sc, x, wi, t, hit_id, n_scans)) {
            const MatF &src = mats[hit_id];
            if (src.emits) {
                const float gpdf = facet_pdf(wo_l, wh, obj.alpha);
                const F3 g = had(mk(src.lr, src.lg, src.lb), facet_brdf(obj, wi_l, wh, wo_l)) * (wi_l.z / gpdf);
                if (g.x > 0.0f) {
                    const F3 cx = mk(src.px, src.py, src.pz) - x;
                    omc_last = one_minus_cos_max(src.r * src.r / dot(cx, cx));
                }
                const float w = power_heuristic(gpdf, 1.0f / (kTwoPi * omc_last));
                total = total + g * w;
            }
        }
    } else {
        const F3 wi = unit(to_world(fr, cosine_local(xi1, xi2)));
        float t; int hit_id;
        if (scan(sc, x, wi, t, hit_id, n_scans)) {
            const MatF &src = mats[hit_id];
            const F3 g = had(mk(src.lr, src.lg, src.lb), mk(obj.cr, obj.cg, obj.cb)); // Le*c/pi * cos / (cos/pi)
            if (g.x > 0.0f && g.y > 0.0f && g.z > 0.0f) {
                const F3 cx = mk(src.px, src.py, src.pz) - x;
                const float omc = one_minus_cos_max(src.r * src.r / dot(cx, cx));
                const float w = power_heuristic(dot(fr.n, wi) * kInvPi, 1.0f / (kTwoPi * omc));
                total = total + g * w;
            }
        }
    }
    return total;
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

// freeSingleScattering (volumetricBasicFunctions.h:284-340) / singleScattering (:225-281) without the 1/probSource,
// transmitanceXT and sigma_s factors (the caller applies them).  The reference always draws two cone numbers (slots S_NEE); the point-light case needs none.
template <class RngT>
__device__ __forceinline__ F3 medium_direct(const SceneF &sc, const MatF &src, int src_id, F3 xt, const Consts &k, RngT &rng, unsigned &n_scans) {
    const F3 light = mk(src.px, src.py, src.pz);
    const F3 wc = light - xt;
    const float d2 = dot(wc, wc), inv = rsqrtf(d2);
    if (src.r == 0.0f) { // the reference draws two cone numbers here too (sequential list streams must still consume them)
        rng.next_f32(S_NEE); rng.next_f32(S_NEE + 1);
        if (!visible(sc, light, xt, n_scans)) return mk(0, 0, 0);
        return mk(src.lr, src.lg, src.lb) * (expf(-k.sigma_t * d2 * inv) * kInv4Pi / d2);
    }
    const float xi1 = rng.next_f32(S_NEE), xi2 = rng.next_f32(S_NEE + 1);
    const float omc_max = one_minus_cos_max(src.r * src.r / d2);
    const F3 wl = cone_sample(wc * inv, omc_max, xi1, xi2);
    float t; int hit_id = -1;
    scan(sc, xt, wl, t, hit_id, n_scans);
    if (hit_id != src_id) return mk(0, 0, 0);
    return mk(src.lr, src.lg, src.lb) * (expf(-k.sigma_t * t) * kInv4Pi * kTwoPi * omc_max);
}

// One path vertex after a successful roulette draw: iterativeVPTracerFree (vptShadeMethods.h:1263-1340),
// explicitVPTracerRecursive (:1014-1149) and MISVPTTracerRecursive (:1345-1481) in throughput form.
// Returns false when the path ends here.
template <int METHOD, class RngT>
__device__ __forceinline__ bool vertex(const SceneF &sc, const MatF *mats, const Consts &k, Path &p, RngT &rng, Tally &tally) {
    ++tally.events;
    float t; int id = 0;
    const bool hit = scan(sc, p.o, p.d, t, id, tally.scans);
    if (!hit) t = kMaxFloat;

    const int pick = min((int)(rng.next_f32(S_SRC) * k.n_emitters), sc.n_emitters - 1);
    const int src_id = sc.emitters[pick];
    const MatF &src = mats[src_id];

    bool surface;
    float dist, inv_pdf = 1.0f;
    if (METHOD == 0) {
        dist = -logf(1.0f - rng.next_f32(S_DIST)) * k.inv_sigma_t; // freeFlightSample, vptSamplingFunctions.h:11
        surface = dist > t;
    } else if (METHOD == 4) {
        const float xi = rng.next_f32(S_DIST);
        surface = mis_distance(mk(src.px, src.py, src.pz), p.o, p.d, t, expf(-k.sigma_t * t), k.sigma_t, k.inv_sigma_t, xi, rng.next_f32(S_DECIDE), dist, inv_pdf);
    } else {
        // equiAngularParams2 (volumetricBasicFunctions.h:209-223) + equiAngularProb (vptSamplingFunctions.h:60)
        const float Tr = expf(-k.sigma_t * t); // TrActual :1046 / psurf :1407 (0 on a miss)
        const F3 dv = mk(src.px, src.py, src.pz) - p.o;
        const float proj = dot(dv, p.d);
        const F3 perp = fma3(p.d, -proj, dv);
        const float D = sqrtf(dot(perp, perp));
        const float thA = atan2f(-proj, D), thB = atan2f(t - proj, D);
        const float xi = rng.next_f32(S_DIST);
        const float tl = D * tanf((1.0f - xi) * thA + xi * thB);
        dist = tl + proj;
        inv_pdf = fabsf(thB - thA) * (tl * tl + D * D) / (D * (1.0f - Tr));
        const float xs = rng.next_f32(S_DECIDE);
        surface = (METHOD == 1) ? (xs <= Tr) : (xs < Tr);
    }

    if (surface) {
        const MatF &obj = mats[id];
        if (obj.emits) { // :1308-1313
            if (p.depth == 0) p.L = had(mk(obj.lr, obj.lg, obj.lb), p.beta);
            return false;
        }
        const F3 x = fma3(p.d, t, p.o);
        const Frame fr = make_frame(unit(x - mk(obj.px, obj.py, obj.pz)));
        const F3 wo_l = unit(to_local(fr, -p.d));
        const F3 Ld_point = point_light_direct(sc, obj, src, x, fr, wo_l, k, tally.scans);
        const F3 Ld = surface_direct_mis(sc, mats, obj, x, fr, wo_l, k, rng, tally.scans);
        p.L = p.L + had(Ld_point + Ld, p.beta) * k.inv_cp;
        const float xi1 = rng.next_f32(S_BSDF), xi2 = obj.material == 2 ? 0.0f : rng.next_f32(S_BSDF + 1); // (the dielectric draws one number)
        F3 wi;
        const F3 weight = bsdf_sample(obj, fr, wo_l, xi1, xi2, wi);
        p.beta = had(p.beta, weight) * k.inv_cp;
        p.d = wi;
        p.o = x;
    } else {
        const F3 xt = fma3(p.d, dist, p.o);
        const F3 Ld = medium_direct(sc, src, src_id, xt, k, rng, tally.scans) * k.n_emitters;
        const float xi1 = rng.next_f32(S_PHASE), xi2 = rng.next_f32(S_PHASE + 1);
        if (METHOD == 0) {
            p.L = p.L + had(Ld, p.beta) * k.albedo_over_cp;
            p.beta = p.beta * k.albedo_over_cp;
        } else {
            const float w = k.sigma_s * expf(-k.sigma_t * dist) * inv_pdf * k.inv_cp; // sigma_s * T / (pdf * cp)
            p.L = p.L + had(Ld, p.beta) * w;
            p.beta = p.beta * w;
        }
        p.o = xt;
        p.d = phase_sample(xi1, xi2);
    }
    return true;
}

} // namespace f32
} // namespace vpt
