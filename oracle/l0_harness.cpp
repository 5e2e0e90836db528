/* oracle/l0_harness.cpp -- TEST INFRASTRUCTURE ("oracle level 0"), not product code.
 *
 * A C-ABI window onto the UNMODIFIED reference headers under $REF/include (never copied into this
 * repository): every extern "C" function below simply calls the reference function named in its
 * comment, so tests and tools can obtain the reference's own answers (unit values, per-path
 * radiance, whole renders).  Built only where /root/reference exists (oracle/Makefile, target
 * `l0`), into oracle/_ref/libvpt_l0.so.  Nothing in the product imports or links this.
 *
 * The only own logic here is
 *   (1) the thread-local seedable erand48 wrapper (the reference shares one global seed between
 *       OpenMP threads, include/Vector.cpp:8 + src/rt.cpp:767),
 *   (2) a restatement of the camera / pixel loop of src/rt.cpp:752-805 (main() itself cannot be
 *       called as a function), and
 *   (3) two opt-in hooks that neutralise the two rounding-decided behaviours of the reference
 *        so that a "robust" reference render exists to compare
 *       the fp32 product path against.  With the hooks off (default) every call goes to the
 *       reference's own code.
 */
#include <cstdint>
#include <cstring>
#include <omp.h>

#include "Sphere.h"
#include "mathUtilities.h"
#include "pathTracingUtilities.h" /* the reference's own intersect()/visibility() are defined here */

/* ---- thread-local RNG (1) ------------------------------------------------------------------ */
static thread_local unsigned short tls_seed[3] = {0, 0, 0};
static thread_local uint64_t tls_draws = 0;
static thread_local const double *tls_inject = nullptr; /* explicit uniforms served before the LCG (l0_inject) */
static thread_local size_t tls_inject_n = 0, tls_inject_i = 0;
#undef erand48
double vpt_l0_erand48(unsigned short *) {
    ++tls_draws;
    if (tls_inject_i < tls_inject_n) return tls_inject[tls_inject_i++];
    return erand48(tls_seed);
}
#define erand48(s) vpt_l0_erand48(s)

/* ---- hooks (3) ----------------------------------------------------------------------------- */
static int g_robust_visibility = 0; /* 1: t > distance*(1-1e-4) instead of the exact t > distance   */
static int g_skip_r0 = 0;           /* 1: scene scans ignore r == 0 spheres (kills the fall-through) */

static inline bool vpt_l0_hook_intersect(const Ray &r, double &t, int &id) {
    if (!g_skip_r0) return (intersect)(r, t, id); /* reference: pathTracingUtilities.h:12 */
    /* same selection rule as pathTracingUtilities.h:12-36, minus point-light spheres */
    double best = __DBL_MAX__;
    bool any = false;
    for (size_t i = 0; i < spheres.size(); i++) {
        if (spheres[i].r == 0) continue;
        const double ti = spheres[i].intersect(r);
        if (ti > 0 && std::abs(ti) > 0.0001) {
            any = true;
            if (ti < best) { best = ti; id = (int)i; }
        }
    }
    t = any ? best : 0;
    return any;
}

static inline bool vpt_l0_hook_visibility(Point light, Point x) {
    if (!g_robust_visibility) return (visibility)(light, x); /* reference: pathTracingUtilities.h:39 */
    Vector lx = light - x;
    const double distance = std::sqrt(lx.dot(lx));
    lx.normalize();
    Ray back(light, lx * -1);
    int id = 0;
    double t;
    vpt_l0_hook_intersect(back, t, id);
    return t == 0 || t > distance * (1.0 - 1e-4);
}

/* 3-argument calls are the free function (hooked); 1-argument calls are Sphere::intersect (left alone:
 * a macro is not re-expanded inside its own expansion). */
#define VPT_L0_PICK(_1, _2, _3, NAME, ...) NAME
#define intersect(...) VPT_L0_PICK(__VA_ARGS__, vpt_l0_hook_intersect, vpt_l0_bad_arity, intersect)(__VA_ARGS__)
#define visibility(a, b) vpt_l0_hook_visibility(a, b)

#include "samplingFunctions.h"
#include "vptSamplingFunctions.h"
#include "volumetricBasicFunctions.h"
#include "misSamplingFunctions.h"
#include "vptShadeMethods.h" /* pulls rayMarchingMethods.h (non-inline definitions: one TU only) */

/* include/Sphere.cpp:7-22 as linked.  Captured on first use (not at static-init time: the order of dynamic initialisation
 * across translation units is unspecified, `spheres` might still be empty then); always called before any modification. */
static const std::vector<Sphere> &default_scene() { static const std::vector<Sphere> d = spheres; return d; }

static inline Vector V(const double *p) { return Vector(p[0], p[1], p[2]); }
static inline void put(double *o, const Vector &v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }

extern "C" {

/* ---- state --------------------------------------------------------------------------------- */
void l0_seed(unsigned s0, unsigned s1, unsigned s2) {
    tls_seed[0] = (unsigned short)s0; tls_seed[1] = (unsigned short)s1; tls_seed[2] = (unsigned short)s2;
    tls_draws = 0;
}
uint64_t l0_draws(void) { return tls_draws; }
/* the next n draws of this thread return u[0..n) (caller keeps u alive); resets the draw counter */
void l0_inject(const double *u, int n) { tls_inject = u; tls_inject_n = (size_t)n; tls_inject_i = 0; tls_draws = 0; }
double l0_erand48(void) { return vpt_l0_erand48(nullptr); }
void l0_set_hooks(int robust_visibility, int skip_r0) { g_robust_visibility = robust_visibility; g_skip_r0 = skip_r0; }

/* scene = the reference's global `spheres` (include/Sphere.h:49); 18 doubles per sphere:
 * r, p[3], c[3], radiance[3], material, eta[3], kappa[3], alpha */
int l0_scene_size(void) { return (int)spheres.size(); }
void l0_scene_get(int i, double *o) {
    const Sphere &s = spheres[i];
    o[0] = s.r; put(o + 1, s.p); put(o + 4, s.c); put(o + 7, s.radiance); o[10] = s.material;
    put(o + 11, s.eta); put(o + 14, s.kappa); o[17] = s.alpha;
}
void l0_scene_set(int n, const double *d) {
    (void)default_scene();
    spheres.clear();
    for (int i = 0; i < n; i++, d += 18)
        spheres.emplace_back(d[0], V(d + 1), V(d + 4), V(d + 7), (int)d[10], V(d + 11), V(d + 14), d[17]);
}
void l0_scene_reset(void) { spheres = default_scene(); }

/* ---- geometry: Sphere.h:27, pathTracingUtilities.h:12/39/56/66 ------------------------------ */
double l0_sphere_intersect(int i, const double *o, const double *d) { return spheres[i].intersect(Ray(V(o), V(d))); }
int l0_intersect(const double *o, const double *d, double *t, int *id) {
    int idd = *id; double tt = 0;
    bool h = intersect(Ray(V(o), V(d)), tt, idd);
    *t = tt; *id = idd; return h;
}
int l0_visibility(const double *light, const double *x) { return visibility(V(light), V(x)); }
double l0_cosinethetaMax(int id, const double *x) { return cosinethetaMax(id, V(x)); }
/* mathUtilities.h:10/21/43 */
void l0_coordinateSystem(const double *n, double *s, double *t) { Vector nn = V(n), ss, tt; coordinateSystem(nn, ss, tt); put(s, ss); put(t, tt); }
void l0_coordinateTraspose(const double *n, double *w) { Vector ww = V(w); coordinateTraspose(V(n), ww); put(w, ww); }
int l0_toDisplayValue(double x) { return toDisplayValue(x); }
double l0_clamp(double x) { return clamp(x); }

/* ---- volume sampling: vptSamplingFunctions.h, volumetricBasicFunctions.h --------------------- */
double l0_transmitance(const double *a, const double *b, double st) { return transmitance(V(a), V(b), st); }
double l0_freeFlightSample(double st) { return freeFlightSample(st); }
double l0_freeFlightProb(double st, double d) { return freeFlightProb(st, d); }
double l0_pdfSuccess(double st, double t) { return pdfSuccess(st, t); }
double l0_pdfFailure(double st, double t) { return pdfFailure(st, t); }
void l0_isotropicPhaseSample(double *o) { put(o, isotropicPhaseSample()); }
double l0_isotropicPhaseProb(void) { return isotropicPhaseProb(); }
double l0_equiAngularSample(double D, double a, double b) { return equiAngularSample(D, a, b); }
double l0_equiAngularProb(double D, double a, double b, double t) { return equiAngularProb(D, a, b, t); }
double l0_equiAngularParams2(int src, double tMax, const double *o, const double *d, double *out4) {
    double D, a, b, st;
    double r = equiAngularParams2(src, tMax, Ray(V(o), V(d)), D, a, b, st);
    out4[0] = D; out4[1] = a; out4[2] = b; out4[3] = st; return r;
}
void l0_freeSingleScattering(const double *xt, int src, double st, double pS, double *o) { put(o, freeSingleScattering(V(xt), src, st, pS)); }
void l0_singleScattering(const double *xt, int src, double st, double ss, double T, double pS, double *o) { put(o, singleScattering(V(xt), src, st, ss, T, pS)); }

/* ---- surface sampling: samplingFunctions.h, microFacetUtilities.h, misSamplingFunctions.h ----- */
void l0_cosineHemispheric(const double *n, double *o) { put(o, cosineHemispheric(V(n))); }
void l0_solidAngleDir(const double *wc, double cmax, double *o) { put(o, solidAngle(V(wc), cmax)); }
double l0_solidAngleProb(double cmax) { return solidAngleProb(cmax); }
double l0_hemiCosineProb(double c) { return hemiCosineProb(c); }
void l0_vectorFacet(double alpha, double *o) { put(o, vectorFacet(alpha)); }
double l0_NDF(double c, double a) { return NDF(c, a); }
void l0_fresnel(double c, const double *eta, const double *kappa, double *o) { put(o, fresnel(c, V(eta), V(kappa))); }
double l0_G_smith(const double *n, const double *wi, const double *wo, const double *wh, double a) { return G_smith(V(n), V(wi), V(wo), V(wh), a); }
double l0_microFacetProb(const double *wo, const double *wh, double a, const double *n) { return microFacetProb(V(wo), V(wh), a, V(n)); }
void l0_frMicroFacet(const double *eta, const double *kappa, const double *wi, const double *wh, const double *wo, double a, const double *n, double *o) {
    put(o, frMicroFacet(V(eta), V(kappa), V(wi), V(wh), V(wo), a, V(n)));
}
double l0_fresnelDie(double ei, double et, double ct, double ci) { return fresnelDie(ei, et, ct, ci); }
void l0_reflexDielectric(const double *wi, const double *n, double *o) { put(o, reflexDielectric(V(wi), V(n))); }
void l0_refraxDielectric(double ei, double et, const double *wi, const double *n, double *o) { put(o, refraxDielectric(ei, et, V(wi), V(n))); }
double l0_powerHeuristics(double f, double g) { return powerHeuristics(f, g); }
void l0_muestreoSA(int light, const double *x, int obj, const double *n, const double *wray, double alpha, double *L, double *wi, double *cmax) {
    Vector aux; double cm = 0;
    put(L, muestreoSA(spheres[light], V(x), light, spheres[obj], V(n), V(wray), aux, cm, alpha));
    put(wi, aux); *cmax = cm;
}
void l0_MISv2(int obj, const double *x, const double *n, const double *wray, double alpha, double st, double *o) {
    put(o, MISv2(spheres[obj], V(x), V(n), V(wray), alpha, st));
}
/* vptShadeMethods.h:16/62 */
void l0_pLight(int obj, const double *x, const double *n, const double *wray, const double *I, const double *light, double alpha, double *o) {
    put(o, pLight(spheres[obj], V(x), V(n), V(wray), V(I), V(light), alpha));
}
void l0_bdsf(const double *wray, const double *n, int id, double *wi, double *prob, double *fs) {
    Vector aux; double p = 0;
    put(fs, bdsf(aux, V(wray), V(n), p, id));
    put(wi, aux); *prob = p;
}

/* ---- estimators: vptShadeMethods.h:1263 / 1014 / 1345 / 1153 ---------------------------------- */
static inline Color radiance(int method, const Ray &r, double sa, double ss) {
    switch (method) {
    case 0: return iterativeVPTracerFree(r, sa, ss);
    case 1: return explicitVPTracerRecursive(r, sa, ss, 0);
    case 2: return MISVPTTracerRecursive(r, sa, ss, 0);
    case 5: return explicitPathRecursive2(r, 0); /* vptShadeMethods.h:398: the legacy estimator with volumetric (material 3) spheres */
    default: return explicitVPTracerRecursiveFree(r, sa, ss, 0);
    }
}
void l0_radiance(int method, const double *o, const double *d, double sa, double ss, double *out) { put(out, radiance(method, Ray(V(o), V(d)), sa, ss)); }
/* rayMarchingMethods.h:330 */
void l0_rayMarching3(const double *o, const double *d, double sa, double ss, double step, int src, double *out) { put(out, rayMarching3(Ray(V(o), V(d)), sa, ss, step, src)); }

/* ---- camera + pixel loop: restates src/rt.cpp:752-805 (own code) ------------------------------ */
struct Cam { Point o; Vector d, cx, cy; };
static Cam make_cam(int w, int h) {
    Cam c;
    c.o = Point(0, 11.2, 214);                      /* rt.cpp:755 */
    c.d = Vector(0, -0.042612, -1).normalize();
    c.cx = Vector(w * 0.5095 / h, 0., 0.);          /* rt.cpp:758 */
    c.cy = (c.cx % c.d).normalize() * 0.5095;       /* rt.cpp:759 */
    return c;
}
void l0_camera(int w, int h, double *o, double *d, double *cx, double *cy) { Cam c = make_cam(w, h); put(o, c.o); put(d, c.d); put(cx, c.cx); put(cy, c.cy); }
void l0_camera_ray(int w, int h, int x, int y, double xi1, double xi2, double *dir) {
    Cam c = make_cam(w, h);
    Vector v = c.cx * ((static_cast<double>(x) + xi1 - 0.5) / w - .5) + c.cy * ((static_cast<double>(y) + xi2 - 0.5) / h - .5) + c.d; /* rt.cpp:787 */
    put(dir, v.normalize());
}

static inline uint64_t splitmix(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31);
}

/* sum / sumsq: w*h*3 doubles, index (h-y-1)*w+x as rt.cpp:773, UNCLAMPED per-pixel sums over spp samples
 * (sumsq nullable).  Each image row gets its own erand48 stream derived from (seed, y): the result does not
 * depend on the thread count.  Returns the number of random draws. */
uint64_t l0_render(int w, int h, int spp, int method, double sa, double ss, uint64_t seed, int nthreads, double *sum, double *sumsq) {
    const Cam c = make_cam(w, h);
    uint64_t draws = 0;
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : draws)
    for (int y = 0; y < h; y++) {
        const uint64_t k = splitmix(seed * 0x100000001B3ull + (uint64_t)y);
        l0_seed((unsigned)(k & 0xffff), (unsigned)((k >> 16) & 0xffff), (unsigned)((k >> 32) & 0xffff));
        for (int x = 0; x < w; x++) {
            const size_t idx = ((size_t)(h - y - 1) * w + x) * 3;
            double s[3] = {0, 0, 0}, q[3] = {0, 0, 0};
            for (int i = 0; i < spp; i++) {
                const double xi1 = vpt_l0_erand48(nullptr), xi2 = vpt_l0_erand48(nullptr);
                Vector dir = c.cx * ((static_cast<double>(x) + xi1 - 0.5) / w - .5) + c.cy * ((static_cast<double>(y) + xi2 - 0.5) / h - .5) + c.d;
                const Color v = radiance(method, Ray(c.o, dir.normalize()), sa, ss);
                s[0] += v.x; s[1] += v.y; s[2] += v.z;
                q[0] += v.x * v.x; q[1] += v.y * v.y; q[2] += v.z * v.z;
            }
            for (int k3 = 0; k3 < 3; k3++) { sum[idx + k3] = s[k3]; if (sumsq) sumsq[idx + k3] = q[k3]; }
        }
        draws += tls_draws;
    }
    return draws;
}

} /* extern "C" */
