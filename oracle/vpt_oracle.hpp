/* oracle/vpt_oracle.hpp -- TEST INFRASTRUCTURE ("oracle level 1"), not product code.
 *
 * A CPU restatement, in FP64 and in the reference's operation order, of the per-pixel radiance
 * loop of gabo99cas/minimal_volumetric_path_tracer (src/rt.cpp:767-805 and everything it calls).
 * Each function cites the reference file:line it follows.  It exists so that the CUDA product can
 * be checked on identical random numbers:
 *
 *   - driven by an explicit draw stream (the reference's erand48 sequence) it reproduces the
 *     reference's own per-path results to rounding -- that is how this restatement is PINNED
 *     (tests/test_oracle_pinning.py against oracle/_ref/libvpt_l0.so here, and against the
 *     committed vectors in tests/golden/ everywhere);
 *   - driven by the Philox stream convention of oracle/philox.h it is the common-random-number
 *     truth for the device kernels.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.  The product
 * (minimal_volumetric_path_tracer_b200/) never includes, links or calls anything in oracle/.
 *
 * Deliberate deviations from the reference, all of them definitions of behaviour the reference
 * leaves undefined:
 *   - more than 4 emitters: the reference overflows `int arr[4]` (vptShadeMethods.h:1293); here up
 *     to VPT_ORACLE_MAX_EMITTERS are listed.
 *   - `idHitted` / `sourceid` are uninitialised on a miss (volumetricBasicFunctions.h:324,
 *     samplingFunctions.h:251): here they are -1 (matches nothing); none of the values read from
 *     them in the reference can change a result when the ray missed.
 *   - MISv2's `costhetaMax` is uninitialised when the scene has no area light and the BSDF sample
 *     of a microfacet surface returns g.x <= 0 (misSamplingFunctions.h:162): here it starts at 0.
 *   - material 2 (dielectric): MISv2 reads `gpdf` uninitialised when no area light with radiance.x > 0 exists
 *     (misSamplingFunctions.h:148): here it starts at 0 (weight 0).  Everything else of material 2 is restated AS WRITTEN, including
 *     refraxDielectric's `sqrt(..) - 1` and its -etat/etai scaling (microFacetUtilities.h:130-134).
 *   - material 3 (volumetric sphere) is outside the hot-path scope table: in the three active methods bdsf
 *     (vptShadeMethods.h:16-59) leaves `prob` and the new direction unset for it; scenes using it are rejected by the C API.
 * Two reference behaviours are decided by FP64 rounding; both are
 * reproduced by default and can be switched to a well-defined alternative:
 *   quirk R0_FALLTHROUGH   on: as reference.  off: scene scans skip r == 0 spheres, so the point-light
 *                          in-medium connection is never overwritten by the solid-angle block.
 *   quirk EXACT_VISIBILITY on: `t > distance` as reference.  off: `t > distance * (1 - 1e-4)`.
 */
#ifndef VPT_ORACLE_HPP
#define VPT_ORACLE_HPP

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include "philox.h"

namespace vpt_oracle {

constexpr int VPT_ORACLE_MAX_EMITTERS = 16;
constexpr double kMaxFloat = 3.40282346638528859811704183484516925e+38; /* MAXFLOAT, vptShadeMethods.h:1287 */
constexpr double kPi = 3.14159265358979323846;                          /* M_PI */
/* Philox slot of every draw (oracle/philox.h); sequential streams (erand48, explicit lists) ignore the slot, so the
 * CONSUMPTION ORDER below stays the reference's. */
enum : unsigned { S_RR = 0, S_SRC = 1, S_DIST = 2, S_DECIDE = 3, S_NEE = 4, S_PHASE = 6, S_BSDF = 4, S_MIS = 6, S_AREA = 8,
                  S_DIEL = 40 /* + a: the dielectric's extra draw per area light (misSamplingFunctions.h:116); S_AREA + 2a stays below 40 */ };

enum : unsigned { QUIRK_R0_FALLTHROUGH = 1u, QUIRK_EXACT_VISIBILITY = 2u };

/* ---- Vector.h:10-34 ---------------------------------------------------------------------------- */
struct Vec {
    double x = 0, y = 0, z = 0;
    Vec() = default;
    Vec(double a, double b, double c) : x(a), y(b), z(c) {}
};
static inline Vec operator+(const Vec &a, const Vec &b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline Vec operator-(const Vec &a, const Vec &b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline Vec operator*(const Vec &a, double s) { return {a.x * s, a.y * s, a.z * s}; }
static inline double dot(const Vec &a, const Vec &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline Vec had(const Vec &a, const Vec &b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }               /* mult */
static inline Vec cross(const Vec &a, const Vec &b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; } /* % */
static inline Vec unit(const Vec &a) { return a * (1.0 / std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z)); }   /* normalize */

struct Ray { Vec o, d; };

/* ---- Sphere.h:12-37 ---------------------------------------------------------------------------- */
struct Sphere {
    double r; Vec p, c, radiance; int material; Vec eta, kappa; double alpha;
    bool emits() const { return radiance.x > 0 || radiance.y > 0 || radiance.z > 0; } /* vptShadeMethods.h:1296 */
};

/* Sphere::intersect, Sphere.h:27-37 */
static inline double sphere_t(const Sphere &s, const Ray &ray) {
    const Vec op = ray.o - s.p;
    const double b = dot(op, ray.d);
    const double det = b * b - dot(op, op) + s.r * s.r;
    if (det < 0) return 0.0;
    const double root = std::sqrt(det);
    const double far_t = -b + root, near_t = -b - root;
    if (near_t < 0 || std::fabs(near_t) < 0.0001) return far_t;
    return near_t;
}

struct Scene {
    std::vector<Sphere> s;
    unsigned quirks = QUIRK_R0_FALLTHROUGH | QUIRK_EXACT_VISIBILITY;
    uint64_t scans = 0; /* scene scans performed (statistics only) */
};

/* intersect, pathTracingUtilities.h:12-36.  `id` is only written on a hit; t = 0 on a miss. */
static inline bool scan(Scene &sc, const Ray &ray, double &t, int &id) {
    double best = __DBL_MAX__;
    bool any = false;
    ++sc.scans;
    const bool skip_r0 = !(sc.quirks & QUIRK_R0_FALLTHROUGH);
    for (size_t i = 0; i < sc.s.size(); ++i) {
        if (skip_r0 && sc.s[i].r == 0) continue;
        const double ti = sphere_t(sc.s[i], ray);
        if (ti > 0 && std::fabs(ti) > 0.0001) {
            any = true;
            if (ti < best) { best = ti; id = (int)i; }
        }
    }
    t = any ? best : 0;
    return any;
}

/* visibility, pathTracingUtilities.h:39-53 */
static inline bool visible(Scene &sc, const Vec &light, const Vec &x) {
    Vec lx = light - x;
    const double distance = std::sqrt(dot(lx, lx));
    lx = unit(lx);
    lx = lx * -1;
    int id = 0;
    double t;
    scan(sc, Ray{light, lx}, t, id);
    if (sc.quirks & QUIRK_EXACT_VISIBILITY) return t > distance || t == 0;
    return t == 0 || t > distance * (1.0 - 1e-4);
}

/* rayTracer, pathTracingUtilities.h:56-64 */
static inline Vec first_hit_radiance(Scene &sc, const Vec &x, const Vec &wi, int &source) {
    double t;
    int id = 0;
    if (!scan(sc, Ray{x, wi}, t, id)) return Vec();
    source = id;
    return sc.s[id].radiance;
}

/* cosinethetaMax, pathTracingUtilities.h:66-73 */
static inline double cone_cos(const Scene &sc, int source, const Vec &x) {
    const double radius = sc.s[source].r;
    const Vec cx = sc.s[source].p - x;
    const double len = std::sqrt(dot(cx, cx));
    return std::sqrt(1 - (radius / len) * (radius / len));
}

/* coordinateSystem, mathUtilities.h:10-19 */
static inline void frame(const Vec &n, Vec &s, Vec &t) {
    if (std::fabs(n.x) > std::fabs(n.y)) {
        const double inv = 1.0 / std::sqrt(n.x * n.x + n.z * n.z);
        t = Vec(n.z * inv, 0.0, -n.x * inv);
    } else {
        const double inv = 1.0 / std::sqrt(n.y * n.y + n.z * n.z);
        t = Vec(0.0, n.z * inv, -n.y * inv);
    }
    s = cross(t, n);
}
/* coordinateTraspose, mathUtilities.h:21-30: world -> local (s, t, n) */
static inline Vec to_local(const Vec &n, const Vec &w) {
    Vec s, t;
    frame(n, s, t);
    const Vec col0(s.x, t.x, n.x), col1(s.y, t.y, n.y), col2(s.z, t.z, n.z);
    return col0 * w.x + col1 * w.y + col2 * w.z;
}
static inline Vec from_local(const Vec &n, double a, double b, double c) { /* s*x1 + t*y1 + n*z1, samplingFunctions.h:58 */
    Vec s, t;
    frame(n, s, t);
    return s * a + t * b + n * c;
}

/* ---- vptSamplingFunctions.h ---------------------------------------------------------------------- */
template <class Rng> static inline double free_flight_sample(Rng &rng, double sigma_t) { /* :11-16 */
    const double xi = rng.next(S_DIST);
    return -std::log(1 - xi) / sigma_t;
}
static inline double free_flight_pdf(double sigma_t, double d) { return sigma_t * std::exp(sigma_t * d * -1.0); } /* :20 */
static inline double pdf_success(double sigma_t, double tmax) { return 1.0 - std::exp(-sigma_t * tmax); }           /* :24 */
static inline double pdf_failure(double sigma_t, double tmax) { return std::exp(-sigma_t * tmax); }                 /* :29 */
static inline Vec sph(double theta, double phi) { return Vec(std::sin(theta) * std::cos(phi), std::sin(theta) * std::sin(phi), std::cos(theta)); }
template <class Rng> static inline Vec phase_sample(Rng &rng) { /* isotropicPhaseSample :34-46 */
    const double xi1 = rng.next(S_PHASE);
    const double xi2 = rng.next(S_PHASE + 1);
    return unit(sph(std::acos(1 - 2 * xi1), 2 * kPi * xi2));
}
static inline double phase_value() { return 1 / (4 * kPi); } /* isotropicPhaseFunction, volumetricBasicFunctions.h:59 */
template <class Rng> static inline double equiangular_sample(Rng &rng, double D, double a, double b) { /* :54-57 */
    const double xi = rng.next(S_DIST);
    return D * std::tan((1 - xi) * a + xi * b);
}
static inline double equiangular_pdf(double D, double a, double b, double t) { return D / std::fabs(b - a) / (t * t + D * D); } /* :60 */

/* transmitance, volumetricBasicFunctions.h:14-21 */
static inline double transmittance(const Vec &x1, const Vec &x2, double sigma_t) {
    const Vec v = x2 - x1;
    const double d = std::sqrt(dot(v, v));
    return std::exp(sigma_t * d * -1.0);
}

/* equiAngularParams2, volumetricBasicFunctions.h:209-223 */
struct EquiAngular { double D, thetaA, thetaB, t_local, t_ray; };
template <class Rng> static inline EquiAngular equiangular_setup(Rng &rng, const Scene &sc, int source, double tMax, const Ray &r) {
    EquiAngular e;
    const Vec dv = sc.s[source].p - r.o;
    const double len = std::sqrt(dot(dv, dv));
    const double proj = dot(dv, r.d) / dot(r.d, r.d);
    e.D = std::sqrt(len * len - proj * proj);
    e.thetaA = std::atan2(0.0 - proj, e.D);
    e.thetaB = std::atan2(tMax - proj, e.D);
    const double xi = rng.next(S_DIST);
    e.t_local = e.D * std::tan((1 - xi) * e.thetaA + xi * e.thetaB);
    e.t_ray = e.t_local + proj;
    return e;
}

/* ---- samplingFunctions.h -------------------------------------------------------------------------- */
template <class Rng> static inline Vec cosine_hemisphere(Rng &rng, const Vec &n, unsigned slot) { /* :47-62 */
    const double theta = std::acos(std::sqrt(1 - rng.next(slot)));
    const double phi = 2 * kPi * rng.next(slot + 1);
    const Vec l = sph(theta, phi);
    return unit(from_local(n, l.x, l.y, l.z));
}
template <class Rng> static inline Vec cone_sample(Rng &rng, const Vec &wc, double cos_max, unsigned slot) { /* solidAngle, :65-82 */
    const double e0 = rng.next(slot);
    const double theta = std::acos((1 - e0) + e0 * cos_max);
    const double phi = 2 * kPi * rng.next(slot + 1);
    const Vec l = sph(theta, phi);
    return unit(from_local(wc, l.x, l.y, l.z));
}
static inline double cone_pdf(double cos_max) { return 1 / (2 * kPi * (1 - cos_max)); } /* solidAngleProb :85 */
static inline double cosine_pdf(double c) { return c * 1 / kPi; }                       /* hemiCosineProb :92 */

/* ---- microFacetUtilities.h ------------------------------------------------------------------------- */
static inline double fresnel_channel(double c, double s, double eta, double kappa) { /* fresnelSpectre :11-18 */
    const double a2b2 = std::sqrt((eta * eta - kappa * kappa - s * s) * (eta * eta - kappa * kappa - s * s) + 4 * eta * eta * kappa * kappa);
    const double a = std::sqrt(0.5 * (a2b2 + eta * eta - kappa * kappa - s * s));
    const double perp = (a2b2 + c * c - 2 * a * c) / (a2b2 + c * c + 2 * a * c);
    const double par = perp * (a2b2 * c * c + s * s * s * s - 2 * a * c * s * s) / (a2b2 * c * c + s * s * s * s + 2 * a * c * s * s);
    return 0.5 * (par + perp);
}
static inline Vec fresnel_conductor(double cos_h, const Vec &eta, const Vec &kappa) { /* fresnel :21-29 */
    const double sin_h = std::sqrt(1 - cos_h * cos_h);
    return Vec(fresnel_channel(cos_h, sin_h, eta.x, kappa.x), fresnel_channel(cos_h, sin_h, eta.y, kappa.y), fresnel_channel(cos_h, sin_h, eta.z, kappa.z));
}
static inline double beckmann(double c, double alpha) { /* NDF :34-45 */
    if (c >= 0) {
        const double s = std::sqrt(1 - c * c);
        const double fac1 = kPi * alpha * alpha * c * c * c * c;
        const double tang = s / c;
        const double fac2 = std::exp((-1 * tang * tang) / (alpha * alpha));
        return (1 / fac1) * fac2;
    }
    return 0;
}
static inline double smith_g1(const Vec &n, const Vec &wv, const Vec &wh, double alpha) { /* Gn :47-61 */
    const double s = std::sqrt(1 - dot(n, wv) * dot(n, wv));
    const double tg = s / dot(n, wv);
    const double a = 1 / (alpha * tg);
    if (dot(wv, wh) / dot(wv, n) > 0) {
        if (a < 1.6) {
            const double num = 3.535 * a + 2.181 * a * a;
            const double den = 1 + 2.276 * a + 2.577 * a * a;
            return num / den;
        }
        return 1;
    }
    return 0;
}
static inline double smith_g(const Vec &n, const Vec &wi, const Vec &wo, const Vec &wh, double alpha) { /* G_smith :63-68 */
    const double g1 = smith_g1(n, wi, wh, alpha);
    const double g2 = smith_g1(n, wo, wh, alpha);
    return g1 * g2;
}
template <class Rng> static inline Vec facet_normal(Rng &rng, double alpha, unsigned slot) { /* vectorFacet :71-84 */
    const double theta = std::atan(std::sqrt(-alpha * alpha * std::log(1 - rng.next(slot))));
    const double phi = 2 * kPi * rng.next(slot + 1);
    return unit(sph(theta, phi));
}
static inline double facet_pdf(const Vec &wo, const Vec &wh, double alpha, const Vec &n) { /* microFacetProb :86-92 */
    const double num = dot(wh, n);
    const double den = 4 * std::fabs(dot(wo, wh));
    return beckmann(dot(wh, n), alpha) * num / den;
}
static inline Vec facet_brdf(const Vec &eta, const Vec &kappa, const Vec &wi, const Vec &wh, const Vec &wo, double alpha, const Vec &n) { /* frMicroFacet :95-100 */
    const double den = (4 * std::fabs(dot(n, wi)) * std::fabs(dot(n, wo)));
    return fresnel_conductor(dot(wi, wh), eta, kappa) * beckmann(dot(n, wh), alpha) * smith_g(n, wi, wo, wh, alpha) * (1 / den);
}
static inline double fresnel_dielectric(double etai, double etat, double ct, double ci) { /* fresnelDie :107-112 */
    const double par = ((etat * ci - etai * ct) / (etat * ci + etai * ct)) * ((etat * ci - etai * ct) / (etat * ci + etai * ct));
    const double perp = ((etai * ci - etat * ct) / (etai * ci + etat * ct)) * ((etai * ci - etat * ct) / (etai * ci + etat * ct));
    return 0.5 * (par + perp);
}

/* reflexDielectric :117-120 (wi = outgoing direction, -ray.d) */
static inline Vec reflect_dielectric(const Vec &wi, const Vec &n) { return wi * -1 + n * dot(n, wi) * 2; }
/* refraxDielectric :122-141, as written */
static inline Vec refract_dielectric(double etai, double etat, const Vec &wi, const Vec &n) {
    const Vec wl = to_local(n, wi);
    const double ratio = etat / etai * -1;
    const double cosinei = dot(wi, n);
    const double invratio = etai / etat;
    const double cosinet = std::sqrt(1 - invratio * invratio * (1 - cosinei * cosinei)) - 1;
    return from_local(n, wl.x * ratio, wl.y * ratio, cosinet);
}

/* powerHeuristics, misSamplingFunctions.h:12-16 */
static inline double power_heuristic(double f, double g) {
    const double f2 = f * f, g2 = g * g;
    return f2 / (f2 + g2);
}

/* ---- direct light at a surface --------------------------------------------------------------------- */
/* muestreoSA -> solidAngle(L), samplingFunctions.h:238-247 and :163-206 */
template <class Rng>
static inline Vec light_sampled_direct(Rng &rng, Scene &sc, int light, const Vec &x, const Sphere &obj, const Vec &n, const Vec &wray,
                                       Vec &wi_out, double &cos_max_out, double alpha, unsigned slot) {
    const Sphere &src = sc.s[light];
    Vec cx = src.p - x;
    const double len = std::sqrt(dot(cx, cx));
    cx = cx * (1 / len);
    const double cos_max = std::sqrt(1 - (src.r / len) * (src.r / len));
    cos_max_out = cos_max;
    /* solidAngle(L) */
    const Vec wi = cone_sample(rng, cx, cos_max, slot);
    wi_out = wi;
    Vec wil = to_local(n, wi);
    Vec wol = to_local(n, wray * -1);
    wil = unit(wil);
    wol = unit(wol);
    const Vec wh = unit(wil + wol);
    Vec fr;
    if (obj.material == 0) fr = obj.c * (1 / kPi);
    else if (obj.material == 2) fr = Vec(0, 0, 0);
    else fr = facet_brdf(obj.eta, obj.kappa, wil, wh, wol, alpha, Vec(0, 0, 1));
    double t;
    int id = 0;
    scan(sc, Ray{x, wi}, t, id);
    const Vec Le = (light == id) ? sc.s[id].radiance : Vec();
    return had(Le, fr) * dot(n, wi) * (1 / cone_pdf(cos_max));
}

/* uniform, samplingFunctions.h:250-261 */
template <class Rng>
static inline Vec bsdf_sampled_direct_lambert(Rng &rng, Scene &sc, const Vec &n, const Vec &x, const Vec &albedo, Vec &wi_out, int &source) {
    Vec wi = cosine_hemisphere(rng, n, S_MIS);
    wi = unit(wi);
    int sid = -1;
    const Vec Le = first_hit_radiance(sc, x, wi, sid);
    source = sid;
    const Vec L = Vec() + had(Le, albedo * (1 / kPi)) * dot(n, wi) * (1 / cosine_pdf(dot(n, wi)));
    wi_out = wi;
    return L;
}

/* microfacet, samplingFunctions.h:97-118 (wray = incoming ray direction, wh local) */
static inline Vec bsdf_sampled_direct_facet(Scene &sc, const Vec &x, const Vec &wray, const Vec &wh, const Vec &n, const Sphere &obj, double alpha, int &source) {
    const Vec nl(0, 0, 1);
    Vec wo = wray * (-1);
    wo = unit(to_local(n, wo));
    Vec wi = unit(wo * (-1) + wh * 2 * dot(wh, wo));
    const Vec wig = unit(from_local(n, wi.x, wi.y, wi.z));
    int sid = -1;
    const Vec Le = first_hit_radiance(sc, x, wig, sid);
    source = sid;
    const Vec fr = facet_brdf(obj.eta, obj.kappa, wi, wh, wo, alpha, nl);
    return had(Le, fr) * dot(nl, wi) * (1 / facet_pdf(wo, wh, alpha, nl));
}

/* softDielectric, samplingFunctions.h:209-235 (called as softDielectric(1.5, 1.0, wo, ...): etat = 1.5, etai = 1.0) */
template <class Rng>
static inline Vec bsdf_sampled_direct_dielectric(Rng &rng, Scene &sc, double etat, double etai, const Vec &wi, const Vec &n, const Vec &x, int &source) {
    Vec Ld;
    int sid = -1;
    Vec wt = refract_dielectric(etai, etat, wi, n);
    wt = unit(wt);
    const double F = fresnel_dielectric(etai, etat, dot(n, wt), dot(n, wi));
    if (rng.next(S_MIS) < F) {
        Vec wr = reflect_dielectric(wi, n);
        wr = unit(wr);
        Ld = first_hit_radiance(sc, x, wr, sid) * (1 / std::fabs(dot(n, wr)));
    } else {
        const double ratio = etat / etai;
        Ld = first_hit_radiance(sc, x, wt, sid) * (1 / std::fabs(dot(n, wt))) * ratio * ratio;
    }
    source = sid;
    return Ld;
}

/* MISv2, misSamplingFunctions.h:96-170 */
template <class Rng>
static inline Vec surface_direct_mis(Rng &rng, Scene &sc, const Sphere &obj, const Vec &x, const Vec &n, const Vec &wray, double alpha, double sigma_t) {
    Vec total;
    Vec wo = wray * -1;
    double cos_max = 0; /* reference: uninitialised */
    double gpdf_loop = 0; /* reference: `gpdf` is function-wide and uninitialised; the dielectric branch reads what the light loop left */
    const int count = (int)sc.s.size();
    unsigned area_index = 0;
    for (int light = 0; light < count; ++light) {
        if (sc.s[light].r > 0 && sc.s[light].radiance.x > 0) {
            Vec wi_light;
            const Vec f = light_sampled_direct(rng, sc, light, x, obj, n, wray, wi_light, cos_max, alpha, S_AREA + 2 * area_index++) * transmittance(x, sc.s[light].p, sigma_t);
            const double fpdf = cone_pdf(cos_max);
            double gpdf;
            if (obj.material == 0) gpdf = cosine_pdf(dot(n, wi_light));
            else if (obj.material == 2) { /* :110-118 */
                Vec wt = refract_dielectric(1.0, 1.5, wo, n);
                wt = unit(wt);
                gpdf = fresnel_dielectric(1.0, 1.5, dot(n, wt), dot(n, wo));
                if (rng.next(S_DIEL + (area_index - 1)) > gpdf) gpdf = 1 - gpdf;
            } else {
                const Vec wh = unit(wi_light + wo);
                gpdf = facet_pdf(wo, wh, alpha, n);
            }
            gpdf_loop = gpdf;
            const double wf = power_heuristic(fpdf, gpdf);
            total = total + f * wf;
        }
    }
    Vec g;
    double wg;
    if (obj.material == 0) {
        Vec wi_b;
        int source = -1;
        g = bsdf_sampled_direct_lambert(rng, sc, n, x, obj.c, wi_b, source);
        const double gpdf = cosine_pdf(dot(n, wi_b));
        if (g.x > 0 && g.y > 0 && g.z > 0) {
            cos_max = cone_cos(sc, source, x);
            const double fpdf = cone_pdf(cos_max);
            wg = power_heuristic(gpdf, fpdf);
        } else wg = 0;
    } else if (obj.material == 2) { /* :144-152 */
        int source = -1;
        g = bsdf_sampled_direct_dielectric(rng, sc, 1.5, 1.0, wo, n, x, source);
        if (g.x > 0 && g.y > 0 && g.z > 0) {
            cos_max = cone_cos(sc, source, x);
            const double fpdf = cone_pdf(cos_max);
            wg = power_heuristic(gpdf_loop, fpdf);
        } else wg = 0;
    } else {
        const Vec wh = facet_normal(rng, alpha, S_MIS);
        wo = unit(to_local(n, wo));
        int source = -1;
        g = bsdf_sampled_direct_facet(sc, x, wray, wh, n, obj, alpha, source);
        const double gpdf = facet_pdf(wo, wh, alpha, Vec(0, 0, 1));
        if (g.x > 0) cos_max = cone_cos(sc, source, x);
        const double fpdf = cone_pdf(cos_max);
        wg = power_heuristic(gpdf, fpdf);
    }
    return total + g * wg;
}
/* Sphere::intersectVPT, Sphere.h:39-45: both roots as they are (0, 0 on a miss) */
static inline void sphere_t2(const Sphere &s, const Ray &ray, double &t1, double &t2) {
    const Vec op = ray.o - s.p;
    const double b = dot(op, ray.d);
    const double det = b * b - dot(op, op) + s.r * s.r;
    if (det < 0) { t1 = 0.0; t2 = 0.0; return; }
    t2 = -b + std::sqrt(det);
    t1 = -b - std::sqrt(det);
}
/* intersectVPT, volumetricBasicFunctions.h:64-91: intersect() that ignores material-3 (volumetric) spheres */
static inline bool scan_vpt(Scene &sc, const Ray &ray, double &t, int &id) {
    double best = __DBL_MAX__;
    bool any = false;
    ++sc.scans;
    const bool skip_r0 = !(sc.quirks & QUIRK_R0_FALLTHROUGH);
    for (size_t i = 0; i < sc.s.size(); ++i) {
        if (sc.s[i].material == 3 || (skip_r0 && sc.s[i].r == 0)) continue;
        const double ti = sphere_t(sc.s[i], ray);
        if (ti > 0 && std::fabs(ti) > 0.0001) {
            any = true;
            if (ti < best) { best = ti; id = (int)i; }
        }
    }
    t = any ? best : 0;
    return any;
}
/* visibilityVPT, volumetricBasicFunctions.h:94-106 */
static inline bool visible_vpt(Scene &sc, const Vec &light, const Vec &x) {
    Vec lx = light - x;
    const double distance = std::sqrt(dot(lx, lx));
    lx = unit(lx);
    lx = lx * -1;
    int id = 0;
    double t;
    scan_vpt(sc, Ray{light, lx}, t, id);
    if (sc.quirks & QUIRK_EXACT_VISIBILITY) return t > distance || t == 0;
    return t == 0 || t > distance * (1.0 - 1e-4);
}
/* multipleT, volumetricBasicFunctions.h:26-58, as written: every material-3 sphere on the LINE through x1 towards x2 (the segment's end
 * is not checked) attenuates by its full chord; a sphere wholly behind x1 multiplies by exp(-sigma_t * t_near) with t_near < 0 */
static inline double multiple_t(const Scene &sc, const Vec &x1, const Vec &x2, double sigma_t) {
    double T = 1;
    const Ray r{x1, unit(x2 - x1)};
    for (size_t i = 0; i < sc.s.size(); ++i) {
        if (sc.s[i].material != 3) continue;
        double t1, t2;
        sphere_t2(sc.s[i], r, t1, t2);
        if (t2 < 0) T = T * std::exp(-sigma_t * t1);
        if (t2 - t1 > 0) T = T * std::exp(-sigma_t * (t2 - t1));
    }
    return T;
}

/* pLight, vptShadeMethods.h:62-91.  Without material-3 spheres visibilityVPT == visibility: the second branch (:70-74) repeats the
 * scan and cannot fire. */
static inline Vec point_light_direct(Scene &sc, const Sphere &obj, const Vec &x, const Vec &n, const Vec &wray, const Vec &I, const Vec &light, double alpha) {
    Vec Le;
    if (visible(sc, light, x)) Le = I * (1 / dot(light - x, light - x));
    else if (visible_vpt(sc, light, x)) {
        Le = I * (1 / dot(light - x, light - x));
        Le = Le * multiple_t(sc, x, light, 0.05 + 0.009);
    }
    Vec wi = unit(light - x);
    Vec wo = wray * -1;
    wo = to_local(n, wo);
    wi = to_local(n, wi);
    wi = unit(wi);
    wo = unit(wo);
    const Vec wh = unit(wi + wo);
    Vec fr;
    if (obj.material == 1) fr = facet_brdf(obj.eta, obj.kappa, wi, wh, wo, alpha, Vec(0, 0, 1));
    else fr = obj.c * (1 / kPi);
    return had(Le, fr) * dot(n, unit(light - x));
}

/* bdsf, vptShadeMethods.h:16-59 */
template <class Rng>
static inline Vec bsdf_sample(Rng &rng, const Scene &sc, Vec &wi_out, const Vec &wray, const Vec &n, double &pdf, int id) {
    const Sphere &obj = sc.s[id];
    const Vec wo = wray * -1;
    Vec fs;
    if (obj.material == 0) {
        const Vec wi = cosine_hemisphere(rng, n, S_BSDF);
        fs = obj.c * (1 / kPi);
        pdf = cosine_pdf(dot(n, wi));
        wi_out = wi;
    } else if (obj.material == 2) { /* :26-46 */
        Vec wt = refract_dielectric(1.0, 1.5, wo, n);
        wt = unit(wt);
        const double F = fresnel_dielectric(1.0, 1.5, dot(n, wt), dot(n, wo));
        Vec wi;
        if (rng.next(S_BSDF) < F) {
            wi = reflect_dielectric(wo, n);
            wi = unit(wi);
            fs = Vec(1, 1, 1) * (1 / dot(n, wi)) * F;
            pdf = F;
        } else {
            wi = wt;
            fs = Vec(1, 1, 1) * (1 / dot(n, wi)) * (1 - F) * 1.5 * 1.5;
            pdf = 1 - F;
        }
        wi_out = wi;
    } else {
        const double alpha = obj.alpha;
        Vec wh = facet_normal(rng, alpha, S_BSDF);
        wh = from_local(n, wh.x, wh.y, wh.z);
        const Vec wi = wo * (-1) + wh * 2 * dot(wh, wo);
        fs = facet_brdf(obj.eta, obj.kappa, wi, wh, wo, alpha, n);
        pdf = facet_pdf(wo, wh, alpha, n);
        wi_out = wi;
    }
    return fs;
}

/* ---- in-medium next-event estimation ------------------------------------------------------------- */
/* shared body of freeSingleScattering (volumetricBasicFunctions.h:284-340, in_medium_factor = 1) and
 * singleScattering (:225-281, in_medium_factor = transmitanceXT * sigma_s applied in the reference's order) */
template <class Rng>
static inline Vec medium_direct(Rng &rng, Scene &sc, const Vec &xt, int source, double sigma_t, double prob_source, bool equi, double sigma_s, double T_xt) {
    const Sphere &src = sc.s[source];
    Vec Ld;
    if (src.r == 0) {
        const Vec light = src.p;
        if (visible(sc, light, xt)) {
            Vec Le = src.radiance;
            const double d2 = dot(light - xt, light - xt);
            Le = Le * (1 / d2);
            const Vec Ls = Le * transmittance(xt, light, sigma_t) * phase_value();
            Ld = equi ? Ls * T_xt * sigma_s * (1 / prob_source) : Ls * (1 / prob_source);
        }
    }
    Vec wc = src.p - xt;
    const double len = std::sqrt(dot(wc, wc));
    wc = wc * (1 / len);
    const double cos_max = std::sqrt(1 - src.r / len * (src.r / len));
    const Vec wl = cone_sample(rng, wc, cos_max, S_NEE);
    const double prob_wl = cone_pdf(cos_max);
    double dist;
    int hit_id = -1;
    scan(sc, Ray{xt, wl}, dist, hit_id);
    if (source == hit_id) {
        const Vec Le = src.radiance;
        const double Tr = std::exp(sigma_t * dist * -1.0);
        const Vec Ls = Le * Tr * phase_value();
        Ld = equi ? Ls * T_xt * sigma_s * (1 / prob_wl) * (1 / prob_source) : Ls * (1 / prob_wl) * (1 / prob_source);
    }
    return Ld;
}

/* ---- estimators ------------------------------------------------------------------------------------ */
/* rayMarching3, rayMarchingMethods.h:330-384 (the commented call src/rt.cpp:791): constant-step Riemann sum of the single scattering
 * from the centre of sphere `source`.  As written in the reference: each sample is attenuated by the transmittance from the SURFACE
 * point x to the sample (:350), not from the ray origin.  n_steps (nullable) = loop iterations. */
static inline Vec ray_march3(Scene &sc, const Ray &r, double sigma_a, double sigma_s, double step, int source, uint64_t *n_steps = nullptr) {
    double t;
    int id = 0;
    if (n_steps) *n_steps = 0;
    if (!scan(sc, r, t, id)) return Vec();
    const Vec x = r.o + r.d * t;
    Vec Li;
    const double steps = t / step;
    int i = 0;
    for (; i < steps; i++) {
        const Vec xt = r.o + r.d * step * i;
        const double T = transmittance(x, xt, sigma_a + sigma_s);
        const double phase = phase_value();
        const Vec wc = sc.s[source].p - xt;
        const double normwc = dot(wc, wc);
        if (visible(sc, sc.s[source].p, xt)) {
            const Vec Le = sc.s[source].radiance * (1 / normwc);
            const Vec Ls = Le * (phase * transmittance(xt, sc.s[source].p, sigma_a + sigma_s));
            Li = Li + Ls * (T) * sigma_s * step;
        }
    }
    if (n_steps) *n_steps = (uint64_t)i;
    return Li;
}

/* Method 4, distance-sampling MIS.  NOT in the reference: its "MIS" method (vptShadeMethods.h:1345) is the
 * equi-angular estimator again.  Both of the reference's distance techniques end at the surface with probability Tr = exp(-sigma_t t) and
 * otherwise place a medium vertex on [0, t) -- free flight with density sigma_t exp(-sigma_t s) (freeFlightProb, vptSamplingFunctions.h:20),
 * equi-angular with equiAngularProb(s) (1 - Tr) (:60, vptShadeMethods.h:1093); each density has mass 1 - Tr.  Here ONE technique is chosen
 * with probability 1/2 and the sample is weighted with the balance heuristic, i.e. divided by the mixture density
 *     p(s) = (sigma_t exp(-sigma_t s) + equiAngularProb(s) (1 - Tr)) / 2.
 * Draws as in method 1: xi = S_DIST, then xd = S_DECIDE;  xd < Tr: surface;  xd < (1 + Tr) / 2: free flight restricted to [0, t),
 * s = -log(1 - xi (1 - Tr)) / sigma_t;  else equi-angular, s = proj + D tan((1 - xi) thetaA + xi thetaB).  Everything after the
 * distance is method 1's code (singleScattering with 1 / p). */
struct MisDistance { bool surface; double dist, pdf; };
static inline MisDistance mis_distance(const Scene &sc, int source, double t, const Ray &r, double sigma_t, double xi, double xd) {
    MisDistance m{false, 0, 1};
    const double Tr = std::exp(sigma_t * t * -1.0); /* 0 on a miss (t = MAXFLOAT) */
    if (xd < Tr) { m.surface = true; return m; }
    const Vec dv = sc.s[source].p - r.o;
    const double len = std::sqrt(dot(dv, dv));
    const double proj = dot(dv, r.d) / dot(r.d, r.d);
    const double D = std::sqrt(len * len - proj * proj);
    const double thetaA = std::atan2(0.0 - proj, D), thetaB = std::atan2(t - proj, D);
    double t_local;
    if (xd < 0.5 + 0.5 * Tr) {
        m.dist = -std::log(1 - xi * (1 - Tr)) / sigma_t;
        t_local = m.dist - proj;
    } else {
        t_local = D * std::tan((1 - xi) * thetaA + xi * thetaB);
        m.dist = t_local + proj;
    }
    m.pdf = 0.5 * (free_flight_pdf(sigma_t, m.dist) + equiangular_pdf(D, thetaA, thetaB, t_local) * (1.0 - Tr));
    return m;
}

struct PathStats { uint64_t events = 0; };

/* punctualVolumetric, rayMarchingMethods.h:12-32 */
static inline Vec punctual_volumetric(Scene &sc, int source, const Vec &x, double phase, double sigma_t, double sigma_s) {
    const Vec light = sc.s[source].p;
    if (!visible_vpt(sc, light, x)) return Vec();
    Vec Le = sc.s[source].radiance;
    const double d2 = dot(light - x, light - x);
    Le = Le * (1 / d2);
    const Vec Ls = Le * phase * multiple_t(sc, x, light, sigma_t);
    return Ls * sigma_s;
}
/* intersectV2, volumetricBasicFunctions.h:109-134: the nearest sphere by its NEAR root only -- a sphere the ray starts inside (near root
 * negative) is never hit -- with both of its roots */
static inline bool scan_v2(Scene &sc, const Ray &ray, double &t1, double &t2, int &id) {
    double best = __DBL_MAX__;
    bool any = false;
    ++sc.scans;
    const bool skip_r0 = !(sc.quirks & QUIRK_R0_FALLTHROUGH);
    for (size_t i = 0; i < sc.s.size(); ++i) {
        if (skip_r0 && sc.s[i].r == 0) continue;
        double a, b;
        sphere_t2(sc.s[i], ray, a, b);
        if (a > 0 && std::fabs(a) > 0.0001) {
            any = true;
            if (a < best) { best = a; t1 = a; t2 = b; id = (int)i; }
        }
    }
    if (!any) { t1 = 0; t2 = 0; }
    return any;
}
/* VPT_METHOD_VOLUME_SPHERES = explicitPathRecursive2, vptShadeMethods.h:398-497 (the reference's only estimator that
 * handles material 3): a surface path tracer in VACUUM -- no global medium; sigma_a = 0.05, sigma_s = 0.009 are its own literals for the
 * inside of volumetric spheres -- in throughput form.  A material-3 sphere is ray-marched in 100 steps (single scattering from the point
 * lights, punctualVolumetric) and the ray goes on from the LAST SAMPLE POINT (not the exit point) with the chord's transmittance; surface
 * vertices take pLight for every point light + the legacy MIS (= MISv2 without transmittance) and roulette with q = 0.1 AFTER the direct
 * light; an emitter that is hit returns black at any depth (`radiance.x > 0` only); materials 1 and 2 both scatter as a microfacet with the
 * literal alpha = 0.001.  `events` counts surface vertices and marched spheres. */
template <class Rng>
static inline Vec volume_spheres_radiance(Rng &rng, Scene &sc, Ray ray, PathStats *stats = nullptr) {
    const double sigma_a = 0.05, sigma_s = 0.009;
    const double sigma_t = sigma_a + sigma_s;
    Vec L, beta(1, 1, 1);
    const int n_spheres = (int)sc.s.size();
    for (int bounce = 0, guard = 0; guard < 100000; ++guard) {
        double t, t2;
        int id = 0;
        if (!scan_v2(sc, ray, t, t2, id)) break;
        if (sc.s[id].radiance.x > 0) break;
        if (stats) ++stats->events;
        const Vec x = ray.o + ray.d * t;
        if (sc.s[id].material == 3) {
            const int steps = 100;
            const double distance = t2 - t;
            const double step = distance / steps;
            Vec Ls, xt;
            for (int i = 0; i < steps; i++) {
                xt = x + ray.d * step * i;
                for (int light = 0; light < n_spheres; light++)
                    if (sc.s[light].r == 0)
                        Ls = punctual_volumetric(sc, light, xt, phase_value(), sigma_t, sigma_s) * step * transmittance(x, xt, sigma_t) + Ls;
            }
            L = L + had(Ls, beta);
            beta = beta * transmittance(x, xt, sigma_t);
            ray = Ray{xt, ray.d};
            continue;
        }
        rng.begin_bounce(bounce);
        const Sphere &obj = sc.s[id];
        const Vec n = unit(x - obj.p);
        const Vec wo = ray.d * -1;
        Vec Ld;
        for (int light = 0; light < n_spheres; light++)
            if (sc.s[light].r == 0) Ld = point_light_direct(sc, obj, x, n, ray.d, sc.s[light].radiance, sc.s[light].p, obj.alpha) + Ld;
        Ld = surface_direct_mis(rng, sc, obj, x, n, ray.d, obj.alpha, 0.0) + Ld; /* MIS (misSamplingFunctions.h:19-93) = MISv2 with transmittance 1 */
        L = L + had(Ld, beta);
        const double q = 0.1, continueprob = 1.0 - q;
        if (rng.next(S_RR) < q) break;
        Vec wi, fs;
        double prob;
        if (obj.material == 0) {
            wi = cosine_hemisphere(rng, n, S_BSDF);
            fs = obj.c * (1 / kPi);
            prob = cosine_pdf(dot(n, wi));
        } else {
            const double alpha = 0.001;
            Vec wh = facet_normal(rng, alpha, S_BSDF);
            Vec s_, t_;
            frame(n, s_, t_);
            wh = s_ * wh.x + t_ * wh.y + n * wh.z;
            wi = wo * (-1) + wh * 2 * (dot(wh, wo));
            fs = facet_brdf(obj.eta, obj.kappa, wi, wh, wo, alpha, n);
            prob = facet_pdf(wo, wh, alpha, n);
        }
        const double cosine = dot(n, wi);
        beta = had(beta, fs) * std::fabs(cosine) * (1 / (prob * continueprob));
        ray = Ray{x, wi};
        ++bounce;
    }
    return L;
}

struct Settings {
    int method = 0;              /* 0 free-flight (vptShadeMethods.h:1263), 1 equi-angular (:1014), 2 "MIS" (:1345),
                                    4 distance-sampling MIS (not in the reference, SURVEY.md 8f-4: see mis_distance below) */
    double sigma_a = 0.001, sigma_s = 0.009; /* src/rt.cpp:794 */
    double continue_prob = 0.6;  /* vptShadeMethods.h:1275 */
    int max_depth = 0;           /* <= 0: unlimited (reference) */
};

/* One camera path.  The reference's recursion (methods 1, 2) is unrolled into throughput form
 * (L += beta * A ; beta *= B): identical up to the rounding of the re-associated sum. */
template <class Rng>
static inline Vec radiance(Rng &rng, Scene &sc, Ray ray, const Settings &cfg, PathStats *stats = nullptr) {
    if (cfg.method == 5) return volume_spheres_radiance(rng, sc, ray, stats); /* explicitPathRecursive2: its own literals, no global medium */
    Vec L, beta(1, 1, 1);
    const double sigma_t = cfg.sigma_a + cfg.sigma_s;
    const double cp = cfg.continue_prob, q = 1 - cp;
    const int n_spheres = (int)sc.s.size();
    for (int depth = 0; cfg.max_depth <= 0 || depth < cfg.max_depth; ++depth) {
        rng.begin_bounce(depth);
        if (rng.next(S_RR) < q) break; /* :1282 / :1022 / :1353 */
        if (stats) ++stats->events;

        double t;
        int id = 0;
        const bool hit = scan(sc, ray, t, id);
        if (!hit) t = kMaxFloat;
        const Vec xs = ray.o + ray.d * t;
        double Tr = 0;
        if (cfg.method == 1 && hit) Tr = transmittance(ray.o, xs, sigma_t); /* :1046 */
        const Vec n = unit(xs - sc.s[id].p);

        int emitters[VPT_ORACLE_MAX_EMITTERS];
        int count = 0;
        for (int i = 0; i < n_spheres && count < VPT_ORACLE_MAX_EMITTERS; ++i)
            if (sc.s[i].emits()) emitters[count++] = i;
        if (count == 0) break;
        const double prob_source = 1.0 / count;
        const int source = emitters[static_cast<int>(rng.next(S_SRC) * count)];

        bool surface;
        double dist = 0, pdf_medium = 1;
        if (cfg.method == 0) {
            dist = free_flight_sample(rng, sigma_t); /* :1305 */
            surface = dist > t;
        } else if (cfg.method == 4) {
            const double xi = rng.next(S_DIST);
            const MisDistance m = mis_distance(sc, source, t, ray, sigma_t, xi, rng.next(S_DECIDE));
            surface = m.surface; dist = m.dist; pdf_medium = m.pdf;
        } else {
            if (cfg.method == 2) Tr = std::exp(sigma_t * t * -1.0); /* psurf :1407 */
            const EquiAngular e = equiangular_setup(rng, sc, source, t, ray);
            pdf_medium = equiangular_pdf(e.D, e.thetaA, e.thetaB, e.t_local) * (1.0 - Tr); /* :1093 / :1411 */
            dist = e.t_ray;
            const double xi = rng.next(S_DECIDE);
            surface = (cfg.method == 1) ? (xi <= Tr) : (xi < Tr); /* :1096 / :1414 */
        }

        if (surface) {
            const Sphere &obj = sc.s[id];
            if (obj.emits()) { /* :1308-1313 / :1099-1106 */
                if (depth == 0) L = had(obj.radiance, beta);
                break;
            }
            const Sphere &src = sc.s[source];
            const double Trs = transmittance(xs, src.p, sigma_t);
            const Vec Ld_point = point_light_direct(sc, obj, xs, n, ray.d, src.radiance, src.p, obj.alpha) * Trs * (1 / prob_source);
            const Vec Ld = surface_direct_mis(rng, sc, obj, xs, n, ray.d, obj.alpha, sigma_t);
            Vec wi;
            double pdf;
            const Vec fs = bsdf_sample(rng, sc, wi, ray.d, n, pdf, id);
            wi = unit(wi);
            const double cosine = dot(n, wi);
            L = L + had(Ld_point + Ld, beta) * (1 / cp);
            beta = had(beta, fs) * (1 / cp) * cosine * (1 / pdf);
            ray = Ray{xs, wi};
        } else {
            const Vec xt = ray.o + ray.d * dist;
            if (cfg.method == 0) {
                const Vec Ld = medium_direct(rng, sc, xt, source, sigma_t, prob_source, false, 0, 0);
                const Vec wi = phase_sample(rng);
                L = L + had(Ld, beta) * (cfg.sigma_s / sigma_t) * (1 / cp);
                beta = beta * (cfg.sigma_s / sigma_t) * (1 / cp);
                ray = Ray{xt, wi};
            } else {
                const double T = transmittance(ray.o, xt, sigma_t);
                const Vec Ld = medium_direct(rng, sc, xt, source, sigma_t, prob_source, true, cfg.sigma_s, T);
                const Vec wi = phase_sample(rng);
                L = L + had(Ld * (1 / pdf_medium) * (1 / cp), beta);
                beta = beta * cfg.sigma_s * T * (1 / cp) * (1 / pdf_medium);
                ray = Ray{xt, wi};
            }
        }
    }
    return L;
}

/* ---- camera, src/rt.cpp:752-759 and :787 (parameters are the reference's literals by default) -------- */
struct Camera {
    Vec o, d, cx, cy;
    int w, h;
    Camera(int w_, int h_, Vec o_ = Vec(0, 11.2, 214), Vec dir = Vec(0, -0.042612, -1), double fov = 0.5095) : w(w_), h(h_) {
        o = o_;
        d = unit(dir);
        cx = Vec(w * fov / h, 0., 0.);
        cy = unit(cross(cx, d)) * fov;
    }
    Ray ray(int x, int y, double xi1, double xi2) const {
        const Vec v = cx * ((static_cast<double>(x) + xi1 - 0.5) / w - .5) + cy * ((static_cast<double>(y) + xi2 - 0.5) / h - .5) + d;
        return Ray{o, unit(v)};
    }
};

/* ---- random streams ---------------------------------------------------------------------------------- */
/* explicit list of uniforms (e.g. an erand48 sequence captured from the reference) */
struct ListRng {
    const double *u; size_t n, i = 0; bool overrun = false;
    ListRng(const double *u_, size_t n_) : u(u_), n(n_) {}
    void begin_bounce(int) {}
    double next(unsigned = 0) { if (i >= n) { overrun = true; return 0.0; } /* 0 < q: the next roulette draw ends the path */ return u[i++]; }
};
/* POSIX erand48, sequential (the reference's generator, Vector.h:38) */
struct Erand48Rng {
    unsigned short s[3]; uint64_t draws = 0;
    Erand48Rng(unsigned a, unsigned b, unsigned c) { s[0] = (unsigned short)a; s[1] = (unsigned short)b; s[2] = (unsigned short)c; }
    void begin_bounce(int) {}
    double next(unsigned = 0) { ++draws; return erand48(s); }
};
/* Philox4x32-10 keyed (pixel, sample, bounce), one fixed slot per purpose: the product's stream convention (oracle/philox.h). */
struct PhiloxRng {
    uint32_t key[2], ctr[4], out[4]; uint32_t loaded = 0xffffffffu;
    PhiloxRng(uint64_t seed, uint32_t pixel, uint32_t sample) {
        key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
        ctr[0] = pixel; ctr[1] = sample; ctr[2] = 0; ctr[3] = 0;
    }
    void begin_bounce(int b) { ctr[2] = (uint32_t)b; loaded = 0xffffffffu; }
    double next(unsigned slot) {
        if ((slot >> 2) != loaded) { ctr[3] = slot >> 2; vpt_philox4x32_10(ctr, key, out); loaded = slot >> 2; }
        return vpt_u32_to_unit(out[slot & 3u]);
    }
    void jitter(double &a, double &b) { /* pseudo-bounce 0xffffffff, slots 0 and 1 */
        const uint32_t c[4] = {ctr[0], ctr[1], 0xffffffffu, 0};
        uint32_t o[4];
        vpt_philox4x32_10(c, key, o);
        a = vpt_u32_to_unit(o[0]); b = vpt_u32_to_unit(o[1]);
    }
};

} /* namespace vpt_oracle */
#endif
