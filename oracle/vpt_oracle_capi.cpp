/* oracle/vpt_oracle_capi.cpp -- TEST INFRASTRUCTURE.  C-ABI window onto oracle/vpt_oracle.hpp (the FP64 CPU
 * restatement) for tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.  Not product code.
 *
 * Scenes are passed as n x 18 doubles: r, p[3], c[3], radiance[3], material, eta[3], kappa[3], alpha
 * (the field order of the reference's Sphere constructor, include/Sphere.h:23).
 * Functions that consume random numbers take an explicit list of uniforms `u[nu]` so that the reference
 * (oracle/l0_harness.cpp, l0_inject), this restatement and the device kernels see identical inputs. */
#include <cstring>
#include <omp.h>
#include "vpt_oracle.hpp"

using namespace vpt_oracle;

static inline Vec V(const double *p) { return Vec(p[0], p[1], p[2]); }
static inline void put(double *o, const Vec &v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }
static Sphere S(const double *d) { return Sphere{d[0], V(d + 1), V(d + 4), V(d + 7), (int)d[10], V(d + 11), V(d + 14), d[17]}; }
static Scene make_scene(const double *d, int n, unsigned quirks) {
    Scene sc;
    for (int i = 0; i < n; ++i) sc.s.push_back(S(d + 18 * i));
    sc.quirks = quirks;
    return sc;
}
static int check_scene(const double *d, int n) {
    if (n <= 0) return -1;
    int emitters = 0;
    for (int i = 0; i < n; ++i) {
        const Sphere s = S(d + 18 * i);
        if (s.material != 0 && s.material != 1 && s.material != 2) return -2;
        if (s.emits()) ++emitters;
    }
    if (emitters > VPT_ORACLE_MAX_EMITTERS) return -3;
    return 0;
}

extern "C" {

void l1_philox(const uint32_t *ctr, const uint32_t *key, uint32_t *out) { vpt_philox4x32_10(ctr, key, out); }
int l1_check_scene(const double *scene, int n) { return check_scene(scene, n); }

/* ---- geometry ---------------------------------------------------------------------------------------- */
double l1_sphere_intersect(const double *sph, const double *o, const double *d) { return sphere_t(S(sph), Ray{V(o), V(d)}); }
int l1_intersect(const double *scene, int n, unsigned quirks, const double *o, const double *d, double *t, int *id) {
    Scene sc = make_scene(scene, n, quirks);
    return scan(sc, Ray{V(o), V(d)}, *t, *id);
}
int l1_visibility(const double *scene, int n, unsigned quirks, const double *light, const double *x) {
    Scene sc = make_scene(scene, n, quirks);
    return visible(sc, V(light), V(x));
}
double l1_cosinethetaMax(const double *scene, int n, int id, const double *x) { Scene sc = make_scene(scene, n, 3); return cone_cos(sc, id, V(x)); }
void l1_coordinateSystem(const double *n, double *s, double *t) { Vec ss, tt; frame(V(n), ss, tt); put(s, ss); put(t, tt); }
void l1_coordinateTraspose(const double *n, double *w) { put(w, to_local(V(n), V(w))); }

/* ---- volume sampling ------------------------------------------------------------------------------------ */
double l1_transmitance(const double *a, const double *b, double st) { return transmittance(V(a), V(b), st); }
double l1_freeFlightSample(double st, const double *u, int nu) { ListRng r(u, nu); return free_flight_sample(r, st); }
double l1_freeFlightProb(double st, double d) { return free_flight_pdf(st, d); }
double l1_pdfSuccess(double st, double t) { return pdf_success(st, t); }
double l1_pdfFailure(double st, double t) { return pdf_failure(st, t); }
void l1_isotropicPhaseSample(const double *u, int nu, double *o) { ListRng r(u, nu); put(o, phase_sample(r)); }
double l1_equiAngularSample(double D, double a, double b, const double *u, int nu) { ListRng r(u, nu); return equiangular_sample(r, D, a, b); }
double l1_equiAngularProb(double D, double a, double b, double t) { return equiangular_pdf(D, a, b, t); }
double l1_equiAngularParams2(const double *scene, int n, int src, double tMax, const double *o, const double *d, const double *u, int nu, double *out4) {
    Scene sc = make_scene(scene, n, 3);
    ListRng r(u, nu);
    const EquiAngular e = equiangular_setup(r, sc, src, tMax, Ray{V(o), V(d)});
    out4[0] = e.D; out4[1] = e.thetaA; out4[2] = e.thetaB; out4[3] = e.t_local;
    return e.t_ray;
}
/* material 2's direction / Fresnel functions on n rows (n[3], wo[3]) -> (normalize(refraxDielectric(1, 1.5, wo, n))[3], normalize(reflexDielectric(wo, n))[3],
 * fresnelDie(1, 1.5, n.wt, n.wo)) -- the call pattern of bdsf (vptShadeMethods.h:27-29) */
void l1_dielectric(int n, const double *rows, double *out) {
    for (int i = 0; i < n; ++i) {
        const Vec nn = V(rows + 6 * i), wo = V(rows + 6 * i + 3);
        const Vec wt = unit(refract_dielectric(1.0, 1.5, wo, nn));
        put(out + 7 * i, wt); put(out + 7 * i + 3, unit(reflect_dielectric(wo, nn)));
        out[7 * i + 6] = fresnel_dielectric(1.0, 1.5, dot(nn, wt), dot(nn, wo));
    }
}
/* method 4's distance decision on n rows (source, tMax, o[3], d[3], sigma_t, xi, xd) -> (surface, dist, pdf) */
void l1_mis_distance(const double *scene, int n_spheres, int n, const double *rows, double *out) {
    Scene sc = make_scene(scene, n_spheres, 0);
    for (int i = 0; i < n; ++i) {
        const double *a = rows + 11 * i;
        const MisDistance m = mis_distance(sc, (int)a[0], a[1], Ray{V(a + 2), V(a + 5)}, a[8], a[9], a[10]);
        out[3 * i] = m.surface; out[3 * i + 1] = m.dist; out[3 * i + 2] = m.pdf;
    }
}
void l1_freeSingleScattering(const double *scene, int n, unsigned quirks, const double *xt, int src, double st, double pS, const double *u, int nu, double *o) {
    Scene sc = make_scene(scene, n, quirks);
    ListRng r(u, nu);
    put(o, medium_direct(r, sc, V(xt), src, st, pS, false, 0, 0));
}
void l1_singleScattering(const double *scene, int n, unsigned quirks, const double *xt, int src, double st, double ss, double T, double pS, const double *u, int nu, double *o) {
    Scene sc = make_scene(scene, n, quirks);
    ListRng r(u, nu);
    put(o, medium_direct(r, sc, V(xt), src, st, pS, true, ss, T));
}

/* ---- surface sampling ----------------------------------------------------------------------------------- */
void l1_cosineHemispheric(const double *n, const double *u, int nu, double *o) { ListRng r(u, nu); put(o, cosine_hemisphere(r, V(n), 0)); }
void l1_solidAngleDir(const double *wc, double cmax, const double *u, int nu, double *o) { ListRng r(u, nu); put(o, cone_sample(r, V(wc), cmax, 0)); }
double l1_solidAngleProb(double cmax) { return cone_pdf(cmax); }
double l1_hemiCosineProb(double c) { return cosine_pdf(c); }
void l1_vectorFacet(double alpha, const double *u, int nu, double *o) { ListRng r(u, nu); put(o, facet_normal(r, alpha, 0)); }
double l1_NDF(double c, double a) { return beckmann(c, a); }
void l1_fresnel(double c, const double *eta, const double *kappa, double *o) { put(o, fresnel_conductor(c, V(eta), V(kappa))); }
double l1_G_smith(const double *n, const double *wi, const double *wo, const double *wh, double a) { return smith_g(V(n), V(wi), V(wo), V(wh), a); }
double l1_microFacetProb(const double *wo, const double *wh, double a, const double *n) { return facet_pdf(V(wo), V(wh), a, V(n)); }
void l1_frMicroFacet(const double *eta, const double *kappa, const double *wi, const double *wh, const double *wo, double a, const double *n, double *o) {
    put(o, facet_brdf(V(eta), V(kappa), V(wi), V(wh), V(wo), a, V(n)));
}
double l1_fresnelDie(double ei, double et, double ct, double ci) { return fresnel_dielectric(ei, et, ct, ci); }
double l1_powerHeuristics(double f, double g) { return power_heuristic(f, g); }
void l1_muestreoSA(const double *scene, int n, unsigned quirks, int light, const double *x, int obj, const double *nrm, const double *wray, double alpha,
                   const double *u, int nu, double *L, double *wi, double *cmax) {
    Scene sc = make_scene(scene, n, quirks);
    ListRng r(u, nu);
    Vec w;
    put(L, light_sampled_direct(r, sc, light, V(x), sc.s[obj], V(nrm), V(wray), w, *cmax, alpha, 0));
    put(wi, w);
}
void l1_MISv2(const double *scene, int n, unsigned quirks, int obj, const double *x, const double *nrm, const double *wray, double alpha, double st,
              const double *u, int nu, double *o) {
    Scene sc = make_scene(scene, n, quirks);
    ListRng r(u, nu);
    put(o, surface_direct_mis(r, sc, sc.s[obj], V(x), V(nrm), V(wray), alpha, st));
}
void l1_pLight(const double *scene, int n, unsigned quirks, int obj, const double *x, const double *nrm, const double *wray, const double *I, const double *light, double alpha, double *o) {
    Scene sc = make_scene(scene, n, quirks);
    put(o, point_light_direct(sc, sc.s[obj], V(x), V(nrm), V(wray), V(I), V(light), alpha));
}
void l1_bdsf(const double *scene, int n, const double *wray, const double *nrm, int id, const double *u, int nu, double *wi, double *prob, double *fs) {
    Scene sc = make_scene(scene, n, 3);
    ListRng r(u, nu);
    Vec w;
    put(fs, bsdf_sample(r, sc, w, V(wray), V(nrm), *prob, id));
    put(wi, w);
}

/* ---- estimators ----------------------------------------------------------------------------------------- */
static Settings settings(int method, double sa, double ss, double cp, int max_depth) {
    Settings c; c.method = method; c.sigma_a = sa; c.sigma_s = ss; c.continue_prob = cp; c.max_depth = max_depth; return c;
}
/* one path, explicit uniforms; returns number of draws used (negative if the list ran out) */
int l1_radiance_list(const double *scene, int n, unsigned quirks, int method, double sa, double ss, double cp, int max_depth,
                     const double *o, const double *d, const double *u, int nu, double *out) {
    Scene sc = make_scene(scene, n, quirks);
    ListRng r(u, nu);
    put(out, radiance(r, sc, Ray{V(o), V(d)}, settings(method, sa, ss, cp, max_depth)));
    return r.overrun ? -1 : (int)r.i;
}
/* one path, the reference's generator seeded like oracle/l0_harness.cpp l0_seed */
uint64_t l1_radiance_erand48(const double *scene, int n, unsigned quirks, int method, double sa, double ss, double cp, int max_depth,
                             const double *o, const double *d, unsigned s0, unsigned s1, unsigned s2, double *out) {
    Scene sc = make_scene(scene, n, quirks);
    Erand48Rng r(s0, s1, s2);
    put(out, radiance(r, sc, Ray{V(o), V(d)}, settings(method, sa, ss, cp, max_depth)));
    return r.draws;
}
/* `count` paths with given rays on the product's Philox streams (pixel[i], sample[i]); (the pixel-jitter slots are simply
 * not used), exactly as the device's VPT_UNIT_RADIANCE does. */
void l1_radiance_philox(const double *scene, int n, unsigned quirks, int method, double sa, double ss, double cp, int max_depth, uint64_t seed,
                        int count, const double *o, const double *d, const uint32_t *pixel, const uint32_t *sample, double *out, uint64_t *events) {
    const Settings cfg = settings(method, sa, ss, cp, max_depth);
#pragma omp parallel
    {
        Scene sc = make_scene(scene, n, quirks);
#pragma omp for schedule(static)
        for (int i = 0; i < count; ++i) {
            PhiloxRng r(seed, pixel[i], sample[i]);
            PathStats st;
            put(out + 3 * i, radiance(r, sc, Ray{V(o + 3 * i), V(d + 3 * i)}, cfg, &st));
            if (events) events[i] = st.events;
        }
    }
}

/* rayMarching3 on `count` rays; out: count x 4 (L[3], steps) */
void l1_ray_march3(const double *scene, int n, unsigned quirks, double sa, double ss, double step, int source, int count, const double *o, const double *d, double *out) {
#pragma omp parallel
    {
        Scene sc = make_scene(scene, n, quirks);
#pragma omp for schedule(dynamic, 1)
        for (int i = 0; i < count; ++i) {
            uint64_t steps = 0;
            put(out + 4 * i, ray_march3(sc, Ray{V(o + 3 * i), V(d + 3 * i)}, sa, ss, step, source, &steps));
            out[4 * i + 3] = (double)steps;
        }
    }
}
/* whole frame with rayMarching3 per jittered camera ray (rt.cpp:791), Philox jitter as l1_render */
void l1_render_march(const double *scene, int n, unsigned quirks, double sa, double ss, double step, int source,
                     int w, int h, const double *cam_o, const double *cam_dir, double fov, uint64_t seed, int sample_begin, int sample_end, int nthreads, double *sum) {
    const Camera cam(w, h, V(cam_o), V(cam_dir), fov);
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel
    {
        Scene sc = make_scene(scene, n, quirks);
#pragma omp for schedule(dynamic, 1)
        for (int row = 0; row < h; ++row) {
            const int y = h - 1 - row;
            for (int x = 0; x < w; ++x) {
                const size_t pix = (size_t)row * w + x;
                double s[3] = {0, 0, 0};
                for (int k = sample_begin; k < sample_end; ++k) {
                    PhiloxRng r(seed, (uint32_t)pix, (uint32_t)k);
                    double xi1, xi2;
                    r.jitter(xi1, xi2);
                    const Vec v = ray_march3(sc, cam.ray(x, y, xi1, xi2), sa, ss, step, source);
                    s[0] += v.x; s[1] += v.y; s[2] += v.z;
                }
                for (int c = 0; c < 3; ++c) sum[pix * 3 + c] = s[c];
            }
        }
    }
}

/* camera, src/rt.cpp:752-759,787 */
void l1_camera(int w, int h, const double *cam_o, const double *cam_dir, double fov, double *o, double *d, double *cx, double *cy) {
    const Camera c(w, h, V(cam_o), V(cam_dir), fov);
    put(o, c.o); put(d, c.d); put(cx, c.cx); put(cy, c.cy);
}
void l1_camera_ray(int w, int h, const double *cam_o, const double *cam_dir, double fov, int x, int y, double xi1, double xi2, double *dir) {
    const Camera c(w, h, V(cam_o), V(cam_dir), fov);
    put(dir, c.ray(x, y, xi1, xi2).d);
}

/* Whole render on the product's Philox streams: restates the pixel loop src/rt.cpp:767-805.
 * sum / sumsq: w*h*3 doubles, storage index (h-y-1)*w+x (rt.cpp:773), UNCLAMPED sums over samples
 * [sample_begin, sample_end); pixel id of the stream = storage index.  stats3 (nullable): paths, events, scans. */
void l1_render(const double *scene, int n, unsigned quirks, int method, double sa, double ss, double cp, int max_depth,
               int w, int h, const double *cam_o, const double *cam_dir, double fov, uint64_t seed, int sample_begin, int sample_end,
               int nthreads, double *sum, double *sumsq, uint64_t *stats3) {
    const Settings cfg = settings(method, sa, ss, cp, max_depth);
    const Camera cam(w, h, V(cam_o), V(cam_dir), fov);
    uint64_t events = 0, scans = 0;
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel reduction(+ : events, scans)
    {
        Scene sc = make_scene(scene, n, quirks);
#pragma omp for schedule(dynamic, 1)
        for (int row = 0; row < h; ++row) {
            const int y = h - 1 - row;
            for (int x = 0; x < w; ++x) {
                const size_t pix = (size_t)row * w + x;
                double s[3] = {0, 0, 0}, q[3] = {0, 0, 0};
                for (int k = sample_begin; k < sample_end; ++k) {
                    PhiloxRng r(seed, (uint32_t)pix, (uint32_t)k);
                    double xi1, xi2;
                    r.jitter(xi1, xi2);
                    PathStats st;
                    const Vec v = radiance(r, sc, cam.ray(x, y, xi1, xi2), cfg, &st);
                    events += st.events;
                    s[0] += v.x; s[1] += v.y; s[2] += v.z;
                    q[0] += v.x * v.x; q[1] += v.y * v.y; q[2] += v.z * v.z;
                }
                for (int c = 0; c < 3; ++c) { sum[pix * 3 + c] = s[c]; if (sumsq) sumsq[pix * 3 + c] = q[c]; }
            }
        }
        scans += sc.scans;
    }
    if (stats3) { stats3[0] = (uint64_t)w * h * (uint64_t)(sample_end - sample_begin); stats3[1] = events; stats3[2] = scans; }
}

/* host-side output stage, mathUtilities.h:34-45 */
int l1_toDisplayValue(double x) {
    const double c = x < 0.0 ? 0.0 : (x > 1.0 ? 1.0 : x);
    return int(std::pow(c, 1.0 / 2.2) * 255 + .5);
}

} /* extern "C" */
