// vpt_f64.cuh -- FP64 "REF mode" device implementation: the reference's estimators in the reference's own operation
// order (this translation unit is compiled with -fmad=false, so no contraction changes a rounding), including its two
// rounding-decided behaviours when the corresponding VPT_QUIRK_* bit is set.  Purpose: renders that are statistically
// indistinguishable from the as-shipped CPU reference, and per-path agreement with the FP64
// CPU oracle to libm level.  B200 keeps a full-rate FP64 pipe (unlike sm_103), so this mode is a usable product path,
// just slower than vpt_f32.cuh.
//
// Undefined behaviour of the reference that is defined here (same choices as oracle/vpt_oracle.hpp): `idHitted` /
// `sourceid` on a miss = -1; MISv2's stale `costhetaMax` starts at 0; up to VPT_MAX_EMITTERS emitters.
#pragma once
#include <cuda_runtime.h>
#include "vpt_internal.h"
#include "vpt_philox.cuh"

// The estimators' building blocks are compiled OUT OF LINE: with every double-precision transcendental, division and scene scan inlined at
// every call site the wavefront kernel was 698 KB of SASS and spent 34 % of its warp-state samples waiting for instructions (`no_inst`)
// and 50 % at the round barrier behind the warps that did (profiles/r2_summary.md).  A call does not change an IEEE operation.
#define VPT_F64_OUTLINE static __device__ __noinline__
// Building blocks with ONE call site per stage of the SM pipeline are inlined into it (their by-reference arguments stay in registers):
// the vertex parts 899 -> 1066 Mpaths/s, medium_direct + bsdf_sample + point_light_direct 1071 -> 1130 (surface_direct_mis / light_sampled_direct as well:
// 1068 / 1054 against 1100 -- spills).  -DVPT_F64_OUTLINE_STAGE_BLOCKS puts them out of line again.
#ifdef VPT_F64_OUTLINE_STAGE_BLOCKS
#define VPT_F64_STAGE_BLOCK VPT_F64_OUTLINE
#else
#define VPT_F64_STAGE_BLOCK static __device__ __forceinline__
#endif
namespace vpt {
namespace f64 {

constexpr double kPi = 3.14159265358979323846;
constexpr double kMaxFloat = 3.40282346638528859811704183484516925e+38;
constexpr double kDblMax = 1.7976931348623157e+308;

// libm in double precision, one out-of-line copy each (exp alone is ~60 instructions, sincos / tan / atan2 150-300 with their slow paths)
VPT_F64_OUTLINE double m_exp(double x) { return exp(x); }
VPT_F64_OUTLINE double m_log(double x) { return log(x); }
VPT_F64_OUTLINE double m_atan2(double y, double x) { return atan2(y, x); }
VPT_F64_OUTLINE double m_atan(double x) { return atan(x); }
VPT_F64_OUTLINE double m_tan(double x) { return tan(x); }
VPT_F64_OUTLINE double m_acos(double x) { return acos(x); }
VPT_F64_OUTLINE void m_sincos(double x, double *s, double *c) { sincos(x, s, c); }

struct D3 { double x, y, z; };
__device__ __forceinline__ D3 mk(double x, double y, double z) { return D3{x, y, z}; }
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ D3 operator*(D3 a, double s) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ double dot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }             // Vector.h:27
__device__ __forceinline__ D3 had(D3 a, D3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }                // Vector.h:30
__device__ __forceinline__ D3 cross(D3 a, D3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); } // :24
VPT_F64_OUTLINE D3 unit(D3 a) { return a * (1.0 / sqrt(a.x * a.x + a.y * a.y + a.z * a.z)); }      // Vector.h:33

__device__ __forceinline__ D3 pos(const SphereD &s) { return mk(s.px, s.py, s.pz); }
__device__ __forceinline__ D3 rad(const SphereD &s) { return mk(s.lr, s.lg, s.lb); }
__device__ __forceinline__ D3 alb(const SphereD &s) { return mk(s.cr, s.cg, s.cb); }
__device__ __forceinline__ D3 v3(const double *p) { return mk(p[0], p[1], p[2]); }

struct Ctx { // everything a path needs besides its own state
    const SphereD *s; // shared-memory copy of the scene
    int n_spheres, n_emitters, n_volumes; // n_volumes: material-3 (volumetric) spheres, VPT_METHOD_VOLUME_SPHERES only
    const int *emitters;
    unsigned quirks;
    double sigma_a, sigma_s, sigma_t, cp, q;
    int method, max_depth;
};
struct Tally { unsigned events, scans; };

// Sphere::intersect, Sphere.h:27-37
__device__ __forceinline__ double sphere_t(const SphereD &s, D3 o, D3 d) {
    const D3 op = o - pos(s);
    const double b = dot(op, d);
    const double det = b * b - dot(op, op) + s.r * s.r;
    if (det < 0) return 0.0;
    const double root = sqrt(det);
    const double t_far = -b + root, t_near = -b - root;
    if (t_near < 0 || fabs(t_near) < 0.0001) return t_far;
    return t_near;
}
// intersect, pathTracingUtilities.h:12-36 (id only written on a hit; t = 0 on a miss)
VPT_F64_OUTLINE bool scan(const Ctx &c, D3 o, D3 d, double &t, int &id, Tally &tl) {
    double best = kDblMax;
    bool any = false;
    ++tl.scans;
    const bool skip_r0 = !(c.quirks & VPT_QUIRK_R0_FALLTHROUGH);
    for (int i = 0; i < c.n_spheres; ++i) {
        if (skip_r0 && c.s[i].r == 0) continue;
        const double ti = sphere_t(c.s[i], o, d);
        if (ti > 0 && fabs(ti) > 0.0001) {
            any = true;
            if (ti < best) { best = ti; id = i; }
        }
    }
    t = any ? best : 0;
    return any;
}
// visibility, pathTracingUtilities.h:39-53
VPT_F64_OUTLINE bool visible(const Ctx &c, D3 light, D3 x, Tally &tl) {
    D3 lx = light - x;
    const double distance = sqrt(dot(lx, lx));
    lx = unit(lx);
    lx = lx * -1;
    int id = 0;
    double t;
    scan(c, light, lx, t, id, tl);
    if (c.quirks & VPT_QUIRK_EXACT_VISIBILITY) return t > distance || t == 0;
    return t == 0 || t > distance * (1.0 - 1e-4);
}
// rayTracer, pathTracingUtilities.h:56-64
VPT_F64_OUTLINE D3 first_hit_radiance(const Ctx &c, D3 x, D3 wi, int &source, Tally &tl) {
    double t;
    int id = 0;
    if (!scan(c, x, wi, t, id, tl)) return mk(0, 0, 0);
    source = id;
    return rad(c.s[id]);
}
// cosinethetaMax, pathTracingUtilities.h:66-73
VPT_F64_OUTLINE double cone_cos(const Ctx &c, int source, D3 x) {
    const double radius = c.s[source].r;
    const D3 cx = pos(c.s[source]) - x;
    const double len = sqrt(dot(cx, cx));
    return sqrt(1 - (radius / len) * (radius / len));
}
// coordinateSystem / coordinateTraspose, mathUtilities.h:10-30
VPT_F64_OUTLINE void frame(D3 n, D3 &s, D3 &t) {
    if (fabs(n.x) > fabs(n.y)) {
        const double inv = 1.0 / sqrt(n.x * n.x + n.z * n.z);
        t = mk(n.z * inv, 0.0, -n.x * inv);
    } else {
        const double inv = 1.0 / sqrt(n.y * n.y + n.z * n.z);
        t = mk(0.0, n.z * inv, -n.y * inv);
    }
    s = cross(t, n);
}
VPT_F64_OUTLINE D3 to_local(D3 n, D3 w) {
    D3 s, t;
    frame(n, s, t);
    return mk(s.x, t.x, n.x) * w.x + mk(s.y, t.y, n.y) * w.y + mk(s.z, t.z, n.z) * w.z;
}
VPT_F64_OUTLINE D3 from_local(D3 n, D3 l) {
    D3 s, t;
    frame(n, s, t);
    return s * l.x + t * l.y + n * l.z;
}
VPT_F64_OUTLINE D3 sph(double theta, double phi) {
    double st, ct, sp, cp;
    m_sincos(theta, &st, &ct);
    m_sincos(phi, &sp, &cp);
    return mk(st * cp, st * sp, ct);
}
// vptSamplingFunctions.h:34-46 / samplingFunctions.h:47-62 / :65-82
VPT_F64_OUTLINE D3 phase_sample(double xi1, double xi2) { return unit(sph(m_acos(1 - 2 * xi1), 2 * kPi * xi2)); }
VPT_F64_OUTLINE D3 cosine_hemisphere(D3 n, double xi1, double xi2) { return unit(from_local(n, sph(m_acos(sqrt(1 - xi1)), 2 * kPi * xi2))); }
VPT_F64_OUTLINE D3 cone_sample(D3 wc, double cos_max, double e0, double xi2) { return unit(from_local(wc, sph(m_acos((1 - e0) + e0 * cos_max), 2 * kPi * xi2))); }
VPT_F64_OUTLINE double cone_pdf(double cos_max) { return 1 / (2 * kPi * (1 - cos_max)); } // samplingFunctions.h:85
__device__ __forceinline__ double cosine_pdf(double c) { return c * 1 / kPi; }                       // samplingFunctions.h:92
__device__ __forceinline__ double phase_value() { return 1 / (4 * kPi); }                            // volumetricBasicFunctions.h:59
VPT_F64_OUTLINE double transmittance(D3 a, D3 b, double sigma_t) {                        // volumetricBasicFunctions.h:14-21
    const D3 v = b - a;
    return m_exp(sigma_t * sqrt(dot(v, v)) * -1.0);
}
// rayMarching3, rayMarchingMethods.h:330-384: constant-step Riemann sum of the single scattering from the source's centre.
// (The reference attenuates each sample by the transmittance from the SURFACE point x to the sample, :350, not from the ray origin;
// reproduced as written.)  n_steps (nullable) receives the number of loop iterations.
VPT_F64_OUTLINE D3 ray_march3(const Ctx &c, D3 o, D3 d, double step, int source, Tally &tl, double *n_steps = nullptr) {
    double t;
    int id = 0;
    if (n_steps) *n_steps = 0;
    if (!scan(c, o, d, t, id, tl)) return mk(0, 0, 0);
    const D3 x = o + d * t;
    D3 Li = mk(0, 0, 0);
    const double steps = t / step;
    const D3 light = pos(c.s[source]);
    int i = 0;
    for (; i < steps; i++) {
        const D3 xt = o + d * step * i;
        const double T = transmittance(x, xt, c.sigma_a + c.sigma_s);
        const double phase = phase_value();
        const D3 wc = light - xt;
        const double normwc = dot(wc, wc);
        if (visible(c, light, xt, tl)) {
            const D3 Le = rad(c.s[source]) * (1 / normwc);
            const D3 Ls = Le * (phase * transmittance(xt, light, c.sigma_a + c.sigma_s));
            Li = Li + Ls * (T) * c.sigma_s * step;
        }
    }
    if (n_steps) *n_steps = i;
    return Li;
}
VPT_F64_OUTLINE double power_heuristic(double f, double g) { const double f2 = f * f, g2 = g * g; return f2 / (f2 + g2); }

// microFacetUtilities.h
VPT_F64_OUTLINE double fresnel_channel(double c, double s, double eta, double kappa) { // :11-18
    const double a2b2 = sqrt((eta * eta - kappa * kappa - s * s) * (eta * eta - kappa * kappa - s * s) + 4 * eta * eta * kappa * kappa);
    const double a = sqrt(0.5 * (a2b2 + eta * eta - kappa * kappa - s * s));
    const double perp = (a2b2 + c * c - 2 * a * c) / (a2b2 + c * c + 2 * a * c);
    const double par = perp * (a2b2 * c * c + s * s * s * s - 2 * a * c * s * s) / (a2b2 * c * c + s * s * s * s + 2 * a * c * s * s);
    return 0.5 * (par + perp);
}
VPT_F64_OUTLINE D3 fresnel_conductor(double ch, const double *eta, const double *kappa) { // :21-29
    const double sh = sqrt(1 - ch * ch);
    return mk(fresnel_channel(ch, sh, eta[0], kappa[0]), fresnel_channel(ch, sh, eta[1], kappa[1]), fresnel_channel(ch, sh, eta[2], kappa[2]));
}
VPT_F64_OUTLINE double beckmann(double c, double alpha) { // NDF :34-45
    if (c >= 0) {
        const double s = sqrt(1 - c * c);
        const double fac1 = kPi * alpha * alpha * c * c * c * c;
        const double tg = s / c;
        const double fac2 = m_exp((-1 * tg * tg) / (alpha * alpha));
        return (1 / fac1) * fac2;
    }
    return 0;
}
VPT_F64_OUTLINE double smith_g1(D3 n, D3 wv, D3 wh, double alpha) { // Gn :47-61
    const double s = sqrt(1 - dot(n, wv) * dot(n, wv));
    const double tg = s / dot(n, wv);
    const double a = 1 / (alpha * tg);
    if (dot(wv, wh) / dot(wv, n) > 0) {
        if (a < 1.6) return (3.535 * a + 2.181 * a * a) / (1 + 2.276 * a + 2.577 * a * a);
        return 1;
    }
    return 0;
}
VPT_F64_OUTLINE D3 facet_normal(double alpha, double xi1, double xi2) { // vectorFacet :71-84
    return unit(sph(m_atan(sqrt(-alpha * alpha * m_log(1 - xi1))), 2 * kPi * xi2));
}
VPT_F64_OUTLINE double facet_pdf(D3 wo, D3 wh, double alpha, D3 n) { // microFacetProb :86-92
    const double num = dot(wh, n);
    const double den = 4 * fabs(dot(wo, wh));
    return beckmann(dot(wh, n), alpha) * num / den;
}
VPT_F64_OUTLINE D3 facet_brdf(const SphereD &m, D3 wi, D3 wh, D3 wo, double alpha, D3 n) { // frMicroFacet :95-100
    const double den = (4 * fabs(dot(n, wi)) * fabs(dot(n, wo)));
    const double g = smith_g1(n, wi, wh, alpha) * smith_g1(n, wo, wh, alpha);
    return fresnel_conductor(dot(wi, wh), m.eta, m.kappa) * beckmann(dot(n, wh), alpha) * g * (1 / den);
}

// ---- material 2 (dielectric), restated AS WRITTEN in the reference -----------------------------------------------------------------
VPT_F64_OUTLINE double fresnel_dielectric(double etai, double etat, double ct, double ci) { // fresnelDie :107-112
    const double par = ((etat * ci - etai * ct) / (etat * ci + etai * ct)) * ((etat * ci - etai * ct) / (etat * ci + etai * ct));
    const double perp = ((etai * ci - etat * ct) / (etai * ci + etat * ct)) * ((etai * ci - etat * ct) / (etai * ci + etat * ct));
    return 0.5 * (par + perp);
}
__device__ __forceinline__ D3 reflect_dielectric(D3 wi, D3 n) { return wi * -1 + n * dot(n, wi) * 2; } // reflexDielectric :117-120
VPT_F64_OUTLINE D3 refract_dielectric(double etai, double etat, D3 wi, D3 n) {              // refraxDielectric :122-141
    const D3 wl = to_local(n, wi);
    const double ratio = etat / etai * -1;
    const double cosinei = dot(wi, n);
    const double invratio = etai / etat;
    const double cosinet = sqrt(1 - invratio * invratio * (1 - cosinei * cosinei)) - 1;
    return from_local(n, mk(wl.x * ratio, wl.y * ratio, cosinet));
}

// muestreoSA -> solidAngle(L), samplingFunctions.h:238-247 and :163-206
template <class RngT>
VPT_F64_OUTLINE D3 light_sampled_direct(const Ctx &c, int light, D3 x, const SphereD &obj, D3 n, D3 wray, double alpha, D3 &wi_out, double &cos_max_out,
                                                   RngT &rng, Tally &tl, uint32_t slot) {
    const SphereD &src = c.s[light];
    D3 cx = pos(src) - x;
    const double len = sqrt(dot(cx, cx));
    cx = cx * (1 / len);
    const double cos_max = sqrt(1 - (src.r / len) * (src.r / len));
    cos_max_out = cos_max;
    const double e0 = rng.next_f64(slot), e1 = rng.next_f64(slot + 1);
    const D3 wi = cone_sample(cx, cos_max, e0, e1);
    wi_out = wi;
    const D3 wil = unit(to_local(n, wi));
    const D3 wol = unit(to_local(n, wray * -1));
    const D3 wh = unit(wil + wol);
    D3 fr;
    if (obj.material == 0) fr = alb(obj) * (1 / kPi);
    else if (obj.material == 2) fr = mk(0, 0, 0); // samplingFunctions.h:190-193
    else fr = facet_brdf(obj, wil, wh, wol, alpha, mk(0, 0, 1));
    double t;
    int id = 0;
    scan(c, x, wi, t, id, tl);
    const D3 Le = (light == id) ? rad(c.s[id]) : mk(0, 0, 0);
    return had(Le, fr) * dot(n, wi) * (1 / cone_pdf(cos_max));
}

// MISv2, misSamplingFunctions.h:96-170
template <class RngT>
VPT_F64_OUTLINE D3 surface_direct_mis(const Ctx &c, const SphereD &obj, D3 x, D3 n, D3 wray, double alpha, RngT &rng, Tally &tl) {
    D3 total = mk(0, 0, 0);
    D3 wo = wray * -1;
    double cos_max = 0, gpdf_loop = 0; // (the reference's function-wide `gpdf`: its dielectric branch reads what the light loop left, :148)
    uint32_t area_index = 0;
    for (int light = 0; light < c.n_spheres; ++light) {
        if (c.s[light].r > 0 && c.s[light].lr > 0) {
            D3 wi_light;
            const D3 f = light_sampled_direct(c, light, x, obj, n, wray, alpha, wi_light, cos_max, rng, tl, S_AREA + 2 * area_index++) * transmittance(x, pos(c.s[light]), c.sigma_t);
            const double fpdf = cone_pdf(cos_max);
            double gpdf;
            if (obj.material == 0) gpdf = cosine_pdf(dot(n, wi_light));
            else if (obj.material == 2) { // :110-118
                const D3 wt = unit(refract_dielectric(1.0, 1.5, wo, n));
                gpdf = fresnel_dielectric(1.0, 1.5, dot(n, wt), dot(n, wo));
                if (rng.next_f64(S_DIEL + (area_index - 1)) > gpdf) gpdf = 1 - gpdf;
            } else gpdf = facet_pdf(wo, unit(wi_light + wo), alpha, n);
            gpdf_loop = gpdf;
            total = total + f * power_heuristic(fpdf, gpdf);
        }
    }
    D3 g;
    double wg;
    if (obj.material == 0) {
        const double xi1 = rng.next_f64(S_MIS), xi2 = rng.next_f64(S_MIS + 1);
        const D3 wi = unit(cosine_hemisphere(n, xi1, xi2)); // uniform, samplingFunctions.h:250-261
        int source = -1;
        const D3 Le = first_hit_radiance(c, x, wi, source, tl);
        g = mk(0, 0, 0) + had(Le, alb(obj) * (1 / kPi)) * dot(n, wi) * (1 / cosine_pdf(dot(n, wi)));
        const double gpdf = cosine_pdf(dot(n, wi));
        if (g.x > 0 && g.y > 0 && g.z > 0) wg = power_heuristic(gpdf, cone_pdf(cone_cos(c, source, x)));
        else wg = 0;
    } else if (obj.material == 2) { // softDielectric(1.5, 1.0, wo, ...), samplingFunctions.h:209-235; MISv2 :144-152
        const D3 wt = unit(refract_dielectric(1.0, 1.5, wo, n));
        const double F = fresnel_dielectric(1.0, 1.5, dot(n, wt), dot(n, wo));
        int source = -1;
        if (rng.next_f64(S_MIS) < F) {
            const D3 wr = unit(reflect_dielectric(wo, n));
            g = first_hit_radiance(c, x, wr, source, tl) * (1 / fabs(dot(n, wr)));
        } else {
            const double ratio = 1.5 / 1.0;
            g = first_hit_radiance(c, x, wt, source, tl) * (1 / fabs(dot(n, wt))) * ratio * ratio;
        }
        if (g.x > 0 && g.y > 0 && g.z > 0) wg = power_heuristic(gpdf_loop, cone_pdf(cone_cos(c, source, x)));
        else wg = 0;
    } else {
        const double xi1 = rng.next_f64(S_MIS), xi2 = rng.next_f64(S_MIS + 1);
        const D3 wh = facet_normal(alpha, xi1, xi2);
        wo = unit(to_local(n, wo));
        const D3 nl = mk(0, 0, 1); // microfacet, samplingFunctions.h:97-118
        const D3 wi = unit(wo * (-1) + wh * 2 * dot(wh, wo));
        const D3 wig = unit(from_local(n, wi));
        int source = -1;
        const D3 Le = first_hit_radiance(c, x, wig, source, tl);
        g = had(Le, facet_brdf(obj, wi, wh, wo, alpha, nl)) * dot(nl, wi) * (1 / facet_pdf(wo, wh, alpha, nl));
        const double gpdf = facet_pdf(wo, wh, alpha, nl);
        if (g.x > 0) cos_max = cone_cos(c, source, x);
        wg = power_heuristic(gpdf, cone_pdf(cos_max));
    }
    return total + g * wg;
}

// ---- material 3 (volumetric spheres), only reachable through VPT_METHOD_VOLUME_SPHERES ------------------------------------------------
// Sphere::intersectVPT, Sphere.h:39-45: both roots as they are (0, 0 on a miss)
__device__ __forceinline__ void sphere_t2(const SphereD &s, D3 o, D3 d, double &t1, double &t2) {
    const D3 op = o - pos(s);
    const double b = dot(op, d);
    const double det = b * b - dot(op, op) + s.r * s.r;
    if (det < 0) { t1 = 0.0; t2 = 0.0; return; }
    t2 = -b + sqrt(det);
    t1 = -b - sqrt(det);
}
// intersectVPT, volumetricBasicFunctions.h:64-91: intersect() that ignores material-3 spheres
VPT_F64_OUTLINE bool scan_vpt(const Ctx &c, D3 o, D3 d, double &t, int &id, Tally &tl) {
    double best = kDblMax;
    bool any = false;
    ++tl.scans;
    const bool skip_r0 = !(c.quirks & VPT_QUIRK_R0_FALLTHROUGH);
    for (int i = 0; i < c.n_spheres; ++i) {
        if (c.s[i].material == 3 || (skip_r0 && c.s[i].r == 0)) continue;
        const double ti = sphere_t(c.s[i], o, d);
        if (ti > 0 && fabs(ti) > 0.0001) {
            any = true;
            if (ti < best) { best = ti; id = i; }
        }
    }
    t = any ? best : 0;
    return any;
}
// visibilityVPT, volumetricBasicFunctions.h:94-106
VPT_F64_OUTLINE bool visible_vpt(const Ctx &c, D3 light, D3 x, Tally &tl) {
    D3 lx = light - x;
    const double distance = sqrt(dot(lx, lx));
    lx = unit(lx);
    lx = lx * -1;
    int id = 0;
    double t;
    scan_vpt(c, light, lx, t, id, tl);
    if (c.quirks & VPT_QUIRK_EXACT_VISIBILITY) return t > distance || t == 0;
    return t == 0 || t > distance * (1.0 - 1e-4);
}
// multipleT, volumetricBasicFunctions.h:26-58, as written (the segment's end is not checked; a sphere wholly behind x1 multiplies by
// exp(-sigma_t t_near) with t_near < 0)
VPT_F64_OUTLINE double multiple_t(const Ctx &c, D3 x1, D3 x2, double sigma_t) {
    double T = 1;
    const D3 w = unit(x2 - x1);
    for (int i = 0; i < c.n_spheres; ++i) {
        if (c.s[i].material != 3) continue;
        double t1, t2;
        sphere_t2(c.s[i], x1, w, t1, t2);
        if (t2 < 0) T = T * m_exp(-sigma_t * t1);
        if (t2 - t1 > 0) T = T * m_exp(-sigma_t * (t2 - t1));
    }
    return T;
}

// pLight, vptShadeMethods.h:62-91.  Without material-3 spheres visibilityVPT == visibility: the second branch (:70-74) would repeat the
// scan with the same answer and is skipped.
VPT_F64_STAGE_BLOCK D3 point_light_direct(const Ctx &c, const SphereD &obj, D3 x, D3 n, D3 wray, D3 I, D3 light, double alpha, Tally &tl) {
    D3 Le = mk(0, 0, 0);
    if (visible(c, light, x, tl)) Le = I * (1 / dot(light - x, light - x));
    else if (c.n_volumes > 0 && visible_vpt(c, light, x, tl)) {
        Le = I * (1 / dot(light - x, light - x));
        Le = Le * multiple_t(c, x, light, 0.05 + 0.009);
    }
    D3 wi = unit(light - x);
    D3 wo = wray * -1;
    wo = to_local(n, wo);
    wi = to_local(n, wi);
    wi = unit(wi);
    wo = unit(wo);
    const D3 wh = unit(wi + wo);
    D3 fr;
    if (obj.material == 1) fr = facet_brdf(obj, wi, wh, wo, alpha, mk(0, 0, 1));
    else fr = alb(obj) * (1 / kPi);
    return had(Le, fr) * dot(n, unit(light - x));
}

// bdsf, vptShadeMethods.h:16-59
template <class RngT>
VPT_F64_STAGE_BLOCK D3 bsdf_sample(const SphereD &obj, D3 &wi_out, D3 wray, D3 n, double &pdf, RngT &rng) {
    const D3 wo = wray * -1;
    if (obj.material == 2) { // :26-46 (ONE draw)
        const D3 wt = unit(refract_dielectric(1.0, 1.5, wo, n));
        const double F = fresnel_dielectric(1.0, 1.5, dot(n, wt), dot(n, wo));
        if (rng.next_f64(S_BSDF) < F) {
            const D3 wi = unit(reflect_dielectric(wo, n));
            pdf = F; wi_out = wi;
            return mk(1, 1, 1) * (1 / dot(n, wi)) * F;
        }
        pdf = 1 - F; wi_out = wt;
        return mk(1, 1, 1) * (1 / dot(n, wt)) * (1 - F) * 1.5 * 1.5;
    }
    const double xi1 = rng.next_f64(S_BSDF), xi2 = rng.next_f64(S_BSDF + 1);
    if (obj.material == 0) {
        const D3 wi = cosine_hemisphere(n, xi1, xi2);
        pdf = cosine_pdf(dot(n, wi));
        wi_out = wi;
        return alb(obj) * (1 / kPi);
    }
    D3 wh = facet_normal(obj.alpha, xi1, xi2);
    wh = from_local(n, wh);
    const D3 wi = wo * (-1) + wh * 2 * dot(wh, wo);
    pdf = facet_pdf(wo, wh, obj.alpha, n);
    wi_out = wi;
    return facet_brdf(obj, wi, wh, wo, obj.alpha, n);
}

// freeSingleScattering (volumetricBasicFunctions.h:284-340) / singleScattering (:225-281)
template <class RngT>
VPT_F64_STAGE_BLOCK D3 medium_direct(const Ctx &c, D3 xt, int source, double prob_source, bool equi, double T_xt, RngT &rng, Tally &tl) {
    const SphereD &src = c.s[source];
    D3 Ld = mk(0, 0, 0);
    if (src.r == 0) {
        const D3 light = pos(src);
        if (visible(c, light, xt, tl)) {
            D3 Le = rad(src);
            const double d2 = dot(light - xt, light - xt);
            Le = Le * (1 / d2);
            const D3 Ls = Le * transmittance(xt, light, c.sigma_t) * phase_value();
            Ld = equi ? Ls * T_xt * c.sigma_s * (1 / prob_source) : Ls * (1 / prob_source);
        }
    }
    D3 wc = pos(src) - xt;
    const double len = sqrt(dot(wc, wc));
    wc = wc * (1 / len);
    const double cos_max = sqrt(1 - src.r / len * (src.r / len));
    const double e0 = rng.next_f64(S_NEE), e1 = rng.next_f64(S_NEE + 1);
    const D3 wl = cone_sample(wc, cos_max, e0, e1);
    const double prob_wl = cone_pdf(cos_max);
    double dist;
    int hit_id = -1;
    scan(c, xt, wl, dist, hit_id, tl);
    if (source == hit_id) {
        const D3 Ls = rad(src) * m_exp(c.sigma_t * dist * -1.0) * phase_value();
        Ld = equi ? Ls * T_xt * c.sigma_s * (1 / prob_wl) * (1 / prob_source) : Ls * (1 / prob_wl) * (1 / prob_source);
    }
    return Ld;
}

// VPT_METHOD_MIS_DISTANCE: one-sample MIS (balance heuristic) of the free-flight and equi-angular distance techniques -- not in the
// reference; the FP64 form of vpt_f32.cuh::mis_distance, statement by statement oracle/vpt_oracle.hpp::mis_distance.  Returns true for a
// surface vertex; otherwise the distance and the mixture density.
VPT_F64_OUTLINE bool mis_distance(D3 light, D3 o, D3 d, double t, double sigma_t, double xi, double xd, double &dist, double &pdf) {
    dist = 0; pdf = 1;
    const double Tr = m_exp(sigma_t * t * -1.0);
    if (xd < Tr) return true;
    const D3 dv = light - o;
    const double len = sqrt(dot(dv, dv));
    const double proj = dot(dv, d) / dot(d, d);
    const double D = sqrt(len * len - proj * proj);
    const double thA = m_atan2(0.0 - proj, D), thB = m_atan2(t - proj, D);
    double t_local;
    if (xd < 0.5 + 0.5 * Tr) { dist = -m_log(1 - xi * (1 - Tr)) / sigma_t; t_local = dist - proj; }
    else { t_local = D * m_tan((1 - xi) * thA + xi * thB); dist = t_local + proj; }
    pdf = 0.5 * (sigma_t * m_exp(sigma_t * dist * -1.0) + D / fabs(thB - thA) / (t_local * t_local + D * D) * (1.0 - Tr));
    return false;
}

struct Path { D3 o, d, beta, L; int depth; };

// punctualVolumetric, rayMarchingMethods.h:12-32
VPT_F64_OUTLINE D3 punctual_volumetric(const Ctx &c, int source, D3 x, double phase, double sigma_t, double sigma_s, Tally &tl) {
    const D3 light = pos(c.s[source]);
    if (!visible_vpt(c, light, x, tl)) return mk(0, 0, 0);
    D3 Le = rad(c.s[source]);
    const double d2 = dot(light - x, light - x);
    Le = Le * (1 / d2);
    const D3 Ls = Le * phase * multiple_t(c, x, light, sigma_t);
    return Ls * sigma_s;
}
// intersectV2, volumetricBasicFunctions.h:109-134: the nearest sphere by its NEAR root only (a sphere the ray starts inside is never hit)
VPT_F64_OUTLINE bool scan_v2(const Ctx &c, D3 o, D3 d, double &t1, double &t2, int &id, Tally &tl) {
    double best = kDblMax;
    bool any = false;
    ++tl.scans;
    const bool skip_r0 = !(c.quirks & VPT_QUIRK_R0_FALLTHROUGH);
    for (int i = 0; i < c.n_spheres; ++i) {
        if (skip_r0 && c.s[i].r == 0) continue;
        double a, b;
        sphere_t2(c.s[i], o, d, a, b);
        if (a > 0 && fabs(a) > 0.0001) {
            any = true;
            if (a < best) { best = a; t1 = a; t2 = b; id = i; }
        }
    }
    if (!any) { t1 = 0; t2 = 0; }
    return any;
}
// VPT_METHOD_VOLUME_SPHERES = explicitPathRecursive2, vptShadeMethods.h:398-497, in throughput form (statement by statement
// oracle/vpt_oracle.hpp::volume_spheres_radiance, which is pinned on the unmodified reference): a surface path tracer in vacuum whose
// material-3 spheres are ray-marched in 100 steps with the function's own sigma_a = 0.05, sigma_s = 0.009; roulette q = 0.1 after the
// direct light; an emitter that is hit returns black at any depth.
template <class RngT>
VPT_F64_OUTLINE D3 volume_spheres_radiance(const Ctx &c, D3 ro, D3 rd, RngT &rng, Tally &tl) {
    const double sigma_a = 0.05, sigma_s = 0.009;
    const double sigma_t = sigma_a + sigma_s;
    D3 L = mk(0, 0, 0), beta = mk(1, 1, 1);
    for (int bounce = 0, guard = 0; guard < 100000 && bounce < VPT_MAX_DEPTH; ++guard) {
        double t, t2;
        int id = 0;
        if (!scan_v2(c, ro, rd, t, t2, id, tl)) break;
        if (c.s[id].lr > 0) break;
        ++tl.events;
        const D3 x = ro + rd * t;
        if (c.s[id].material == 3) {
            const int steps = 100;
            const double distance = t2 - t;
            const double step = distance / steps;
            D3 Ls = mk(0, 0, 0), xt = x;
            for (int i = 0; i < steps; i++) {
                xt = x + rd * step * i;
                for (int light = 0; light < c.n_spheres; light++)
                    if (c.s[light].r == 0)
                        Ls = punctual_volumetric(c, light, xt, phase_value(), sigma_t, sigma_s, tl) * step * transmittance(x, xt, sigma_t) + Ls;
            }
            L = L + had(Ls, beta);
            beta = beta * transmittance(x, xt, sigma_t);
            ro = xt;
            continue;
        }
        rng.begin_bounce((uint32_t)bounce);
        const SphereD &obj = c.s[id];
        const D3 n = unit(x - pos(obj));
        const D3 wo = rd * -1;
        D3 Ld = mk(0, 0, 0);
        for (int light = 0; light < c.n_spheres; light++)
            if (c.s[light].r == 0) Ld = point_light_direct(c, obj, x, n, rd, rad(c.s[light]), pos(c.s[light]), obj.alpha, tl) + Ld;
        {   // MIS (misSamplingFunctions.h:19-93) = MISv2 with transmittance 1
            Ctx vac = c;
            vac.sigma_t = 0.0;
            Ld = surface_direct_mis(vac, obj, x, n, rd, obj.alpha, rng, tl) + Ld;
        }
        L = L + had(Ld, beta);
        const double q = 0.1, continueprob = 1.0 - q;
        if (rng.next_f64(S_RR) < q) break;
        D3 wi, fs;
        double prob;
        if (obj.material == 0) {
            const double xi1 = rng.next_f64(S_BSDF), xi2 = rng.next_f64(S_BSDF + 1);
            wi = cosine_hemisphere(n, xi1, xi2);
            fs = alb(obj) * (1 / kPi);
            prob = cosine_pdf(dot(n, wi));
        } else {
            const double alpha = 0.001;
            const double xi1 = rng.next_f64(S_BSDF), xi2 = rng.next_f64(S_BSDF + 1);
            D3 wh = facet_normal(alpha, xi1, xi2);
            wh = from_local(n, wh);
            wi = wo * (-1) + wh * 2 * (dot(wh, wo));
            fs = facet_brdf(obj, wi, wh, wo, alpha, n);
            prob = facet_pdf(wo, wh, alpha, n);
        }
        const double cosine = dot(n, wi);
        beta = had(beta, fs) * fabs(cosine) * (1 / (prob * continueprob));
        ro = x; rd = wi;
        ++bounce;
    }
    return L;
}

// One path vertex after a successful roulette draw (vptShadeMethods.h:1263-1340 / :1014-1149 / :1345-1481 in throughput form), in three
// PARTS -- the cut points of the wavefront kernel (vpt_smwave_f64.cuh), which re-queues a path between them; vertex() below runs them
// back to back for the one-thread-per-pixel kernel and the unit kernels.  Same statements, same order, same roundings either way.
enum : int { V_END = 0, V_SURFACE = 1, V_MEDIUM = 2 };
struct VertexPlan { // what the first part hands to the second
    int id, source;        // hit object, picked source
    D3 x;                  // the vertex: surface point xs or medium point xt
    double T, pdf_medium;  // medium vertex (equi-angular methods): transmittance origin -> xt, distance pdf
};

// part 1: scene scan, light pick, distance sampling, surface-or-medium decision.  Lc: radiance this part contributes (a directly seen
// emitter at depth 0, :1308-1313), before it the caller's p.L is untouched.
template <class RngT>
__device__ __forceinline__ int vertex_primary_inl(const Ctx &c, const Path &p, RngT &rng, Tally &tl, VertexPlan &vp, D3 &Lc) {
    ++tl.events;
    Lc = mk(0, 0, 0);
    double t;
    int id = 0;
    const bool hit = scan(c, p.o, p.d, t, id, tl);
    if (!hit) t = kMaxFloat;
    const D3 xs = p.o + p.d * t;
    double Tr = 0;
    if (c.method == 1 && hit) Tr = transmittance(p.o, xs, c.sigma_t);
    if (c.n_emitters == 0) return V_END;
    const int source = c.emitters[static_cast<int>(rng.next_f64(S_SRC) * c.n_emitters)];

    bool surface;
    double dist, pdf_medium = 1;
    if (c.method == 0) {
        dist = -m_log(1 - rng.next_f64(S_DIST)) / c.sigma_t;
        surface = dist > t;
    } else if (c.method == 4) {
        const double xi = rng.next_f64(S_DIST);
        surface = mis_distance(pos(c.s[source]), p.o, p.d, t, c.sigma_t, xi, rng.next_f64(S_DECIDE), dist, pdf_medium);
    } else {
        if (c.method == 2) Tr = m_exp(c.sigma_t * t * -1.0);
        // equiAngularParams2, volumetricBasicFunctions.h:209-223
        const D3 dv = pos(c.s[source]) - p.o;
        const double len = sqrt(dot(dv, dv));
        const double proj = dot(dv, p.d) / dot(p.d, p.d);
        const double D = sqrt(len * len - proj * proj);
        const double thA = m_atan2(0.0 - proj, D), thB = m_atan2(t - proj, D);
        const double xi = rng.next_f64(S_DIST);
        const double t_local = D * m_tan((1 - xi) * thA + xi * thB);
        dist = t_local + proj;
        pdf_medium = D / fabs(thB - thA) / (t_local * t_local + D * D) * (1.0 - Tr);
        const double xs_ = rng.next_f64(S_DECIDE);
        surface = (c.method == 1) ? (xs_ <= Tr) : (xs_ < Tr);
    }
    vp.id = id; vp.source = source; vp.pdf_medium = pdf_medium; vp.T = 0;
    if (surface) {
        const SphereD &obj = c.s[id];
        if (obj.emits) {
            if (p.depth == 0) Lc = had(rad(obj), p.beta);
            return V_END;
        }
        vp.x = xs;
        return V_SURFACE;
    }
    vp.x = p.o + p.d * dist;
    if (c.method != 0) vp.T = transmittance(p.o, vp.x, c.sigma_t);
    return V_MEDIUM;
}

// part 2, surface vertex: pLight + MISv2 + bdsf (:1316-1327).  Lc: the vertex's direct light; p becomes the scattered ray.
template <class RngT>
__device__ __forceinline__ void vertex_surface_inl(const Ctx &c, Path &p, const VertexPlan &vp, RngT &rng, Tally &tl, D3 &Lc) {
    const SphereD &obj = c.s[vp.id];
    const D3 xs = vp.x;
    const D3 n = unit(xs - pos(obj));
    const double prob_source = 1.0 / c.n_emitters;
    const SphereD &src = c.s[vp.source];
    const double Trs = transmittance(xs, pos(src), c.sigma_t);
    const D3 Ld_point = point_light_direct(c, obj, xs, n, p.d, rad(src), pos(src), obj.alpha, tl) * Trs * (1 / prob_source);
    const D3 Ld = surface_direct_mis(c, obj, xs, n, p.d, obj.alpha, rng, tl);
    D3 wi;
    double pdf;
    const D3 fs = bsdf_sample(obj, wi, p.d, n, pdf, rng);
    wi = unit(wi);
    const double cosine = dot(n, wi);
    Lc = had(Ld_point + Ld, p.beta) * (1 / c.cp);
    p.beta = had(p.beta, fs) * (1 / c.cp) * cosine * (1 / pdf);
    p.o = xs; p.d = wi;
}

// part 2, medium vertex: (free)SingleScattering + isotropicPhaseSample (:1330-1337 / :1120-1135)
template <class RngT>
__device__ __forceinline__ void vertex_medium_inl(const Ctx &c, Path &p, const VertexPlan &vp, RngT &rng, Tally &tl, D3 &Lc) {
    const D3 xt = vp.x;
    const double prob_source = 1.0 / c.n_emitters;
    const bool equi = c.method != 0;
    const D3 Ld = medium_direct(c, xt, vp.source, prob_source, equi, equi ? vp.T : 0.0, rng, tl);
    if (!equi) {
        const double xi1 = rng.next_f64(S_PHASE), xi2 = rng.next_f64(S_PHASE + 1);
        Lc = had(Ld, p.beta) * (c.sigma_s / c.sigma_t) * (1 / c.cp);
        p.beta = p.beta * (c.sigma_s / c.sigma_t) * (1 / c.cp);
        p.o = xt; p.d = phase_sample(xi1, xi2);
    } else {
        const double T = vp.T;
        const double xi1 = rng.next_f64(S_PHASE), xi2 = rng.next_f64(S_PHASE + 1);
        Lc = had(Ld * (1 / vp.pdf_medium) * (1 / c.cp), p.beta);
        p.beta = p.beta * c.sigma_s * T * (1 / c.cp) * (1 / vp.pdf_medium);
        p.o = xt; p.d = phase_sample(xi1, xi2);
    }
}
// the same parts out of line (the sequential kernels and the unit kernels, which run them back to back inside one loop)
template <class RngT>
VPT_F64_OUTLINE int vertex_primary(const Ctx &c, const Path &p, RngT &rng, Tally &tl, VertexPlan &vp, D3 &Lc) { return vertex_primary_inl(c, p, rng, tl, vp, Lc); }
template <class RngT>
VPT_F64_OUTLINE void vertex_surface(const Ctx &c, Path &p, const VertexPlan &vp, RngT &rng, Tally &tl, D3 &Lc) { vertex_surface_inl(c, p, vp, rng, tl, Lc); }
template <class RngT>
VPT_F64_OUTLINE void vertex_medium(const Ctx &c, Path &p, const VertexPlan &vp, RngT &rng, Tally &tl, D3 &Lc) { vertex_medium_inl(c, p, vp, rng, tl, Lc); }

// the whole vertex; returns false when the path ends here
template <class RngT>
__device__ __forceinline__ bool vertex(const Ctx &c, Path &p, RngT &rng, Tally &tl) {
    VertexPlan vp;
    D3 Lc;
    const int kind = vertex_primary(c, p, rng, tl, vp, Lc);
    if (kind == V_END) {
        if (Lc.x != 0 || Lc.y != 0 || Lc.z != 0) p.L = Lc; // (a directly seen emitter counts at depth 0 only: nothing collected before)
        return false;
    }
    if (kind == V_SURFACE) vertex_surface(c, p, vp, rng, tl, Lc);
    else vertex_medium(c, p, vp, rng, tl, Lc);
    p.L = p.L + Lc;
    return true;
}

} // namespace f64
} // namespace vpt
