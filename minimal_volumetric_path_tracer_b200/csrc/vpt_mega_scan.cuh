// vpt_mega_scan.cuh -- the scan-converged FP32 megakernel (VPT_KERNEL_MEGA_SCAN).
//
// Measured on B200 (profiles/r1_mega_v0_*): the one-vertex-per-iteration megakernel ran with 9.9 of 32 lanes active and 47 %
// of its stall samples in `no_inst`: the all-sphere scan sat inlined at eight places inside divergent shading branches and the
// 91 KB of SASS thrashed the instruction cache.  Here every loop iteration performs exactly ONE scene scan, at one code site,
// for all live lanes.  A lane is a small state machine: the primary ray of a path vertex, then that vertex's next-event
// queries one at a time -- point-light visibility, one cone-sampled ray per area light, the BSDF-sampled ray of the
// reference's MISv2 -- each "set up a ray and a pending contribution, scan, resolve".  Set-up code is shared between medium
// and surface vertices where the reference does the same thing twice.  Semantics, random-number slots and results are those
// of vpt_f32.cuh (same functions, file:line cited there); only the order in which a vertex's independent direct-light terms
// are added differs, i.e. fp32 rounding.
#pragma once
#include "vpt_f32.cuh"

namespace vpt {
namespace f32 {

// Sphere::intersect + the `t > 1e-4` acceptance of intersect(), branch-free: returns the accepted distance or +inf.
template <bool BIG>
__device__ __forceinline__ float sphere_hit(const GeomF &g, F3 o, F3 d) {
    const F3 oq = mk(o.x - g.qx, o.y - g.qy, o.z - g.qz);
    float c, b, det;
    if (BIG) {
        const F3 m = mk(g.mx, g.my, g.mz);
        c = fmaf(2.0f, dot(oq, m), dot(oq, oq)) + g.c0;
        b = dot(oq + m, d);
        det = fmaf(b, b, -c);
    } else {
        c = dot(oq, oq) - g.r2;
        b = dot(oq, d);
        const F3 l = fma3(d, -b, oq);
        det = g.r2 - dot(l, l);
    }
    const float sq = det * rsqrtf(det);            // NaN when det <= 0: every comparison below is then false
    const float q = -(b + copysignf(sq, b));
    const float other = __fdividef(c, q);
    const float t_near = fminf(q, other), t_far = fmaxf(q, other);
    const float t = t_near >= kEps ? t_near : t_far; // Sphere.h:34
    return t > kEps ? t : CUDART_INF_F;              // pathTracingUtilities.h:20 (false for NaN)
}

__device__ __forceinline__ bool scan_fast(const SceneF &sc, F3 o, F3 d, float &t, int &id) {
    float best = CUDART_INF_F;
    int best_id = -1;
    const int nb = sc.n_big, ng = sc.n_geom;
    for (int i = 0; i < nb; ++i) {
        const float ti = sphere_hit<true>(sc.geom[i], o, d);
        if (ti < best) { best = ti; best_id = sc.geom[i].id; }
    }
    for (int i = nb; i < ng; ++i) {
        const float ti = sphere_hit<false>(sc.geom[i], o, d);
        if (ti < best) { best = ti; best_id = sc.geom[i].id; }
    }
    t = best; id = best_id;
    return best_id >= 0;
}

enum : int { PH_NEW = 0, PH_PRIMARY, PH_MED_POINT, PH_MED_AREA, PH_SURF_POINT, PH_SURF_AREA, PH_MIS, PH_DONE };
enum : int { NX_NONE = 0, NX_POINT, NX_AREA, NX_MIS, NX_FINISH_MED, NX_FINISH_SURF };

struct CameraF { F3 o, d, cx, cy; float inv_w, inv_h; };

template <int METHOD>
__device__ __forceinline__ void render_pixel_scan(const SceneF &sc, const MatF *mats, const Consts &k, const CameraF &cam, float fx, float fy, uint32_t pixel,
                                                  int sample_begin, int sample_end, uint32_t key0, uint32_t key1, double acc[3], Tally &tally, unsigned &nonfinite) {
    F3 o = mk(0, 0, 0), d = mk(0, 0, 1), beta = mk(1, 1, 1), L = mk(0, 0, 0);
    F3 qo = o, qd = d, C = mk(0, 0, 0);
    float lim = 0.0f, w = 0.0f;
    int want = -1, id = 0, src = 0, phase = PH_NEW, depth = 0, a = 0;
    int s = sample_begin;
    uint32_t sample = 0;
    bool active = false;
    uint4 blk = make_uint4(0, 0, 0, 0); // Philox block 0 of the bounce until the vertex is classified, block 1 afterwards
    uint2 area_hi = make_uint2(0, 0);   // lanes 2,3 of the last area-light block (the odd light of a pair)

    for (;;) {
        if (phase == PH_NEW) { // (re)generate until a vertex survives roulette -- cheap, no scan inside
            for (;;) {
                if (!active) {
                    if (s >= sample_end) { phase = PH_DONE; break; }
                    sample = (uint32_t)s++; depth = 0; beta = mk(1, 1, 1); L = mk(0, 0, 0); active = true;
                }
                blk = philox_block(pixel, sample, (uint32_t)depth, 0, key0, key1);
                if ((k.max_depth > 0 && depth >= k.max_depth) || u32_to_unit_f32(blk.x) < k.q) { // vptShadeMethods.h:1282
                    if (isfinite(L.x + L.y + L.z)) { acc[0] += L.x; acc[1] += L.y; acc[2] += L.z; } else ++nonfinite;
                    active = false;
                    continue;
                }
                if (depth == 0) { // camera ray, rt.cpp:787 (jitter: pseudo-bounce block)
                    const uint4 j = philox_block(pixel, sample, kJitterBounce, 0, key0, key1);
                    const float u = (fx + u32_to_unit_f32(j.x) - 0.5f) * cam.inv_w - 0.5f, v = (fy + u32_to_unit_f32(j.y) - 0.5f) * cam.inv_h - 0.5f;
                    o = cam.o; d = unit(fma3(cam.cx, u, fma3(cam.cy, v, cam.d)));
                }
                qo = o; qd = d; phase = PH_PRIMARY;
                break;
            }
        }
        if (phase == PH_DONE) break;

        // ---- the one scene scan of this iteration ------------------------------------------------------------------------
        float t; int hid;
        const bool hit = scan_fast(sc, qo, qd, t, hid);
        ++tally.scans;

        // ---- resolve what the ray was for ------------------------------------------------------------------------------------
        int next = NX_NONE;
        if (phase == PH_PRIMARY) {
            ++tally.events;
            if (!hit) { t = kMaxFloat; hid = 0; }
            const int pick = min((int)(u32_to_unit_f32(blk.y) * k.n_emitters), sc.n_emitters - 1);
            src = sc.emitters[pick];
            bool surface; float dist, inv_pdf = 1.0f;
            if (METHOD == 0) {
                dist = -logf(1.0f - u32_to_unit_f32(blk.z)) * k.inv_sigma_t;
                surface = dist > t;
            } else if (METHOD == 4) { // distance-sampling MIS (vpt_f32.cuh mis_distance)
                const MatF &ls = mats[src];
                surface = mis_distance(mk(ls.px, ls.py, ls.pz), o, d, t, expf(-k.sigma_t * t), k.sigma_t, k.inv_sigma_t, u32_to_unit_f32(blk.z), u32_to_unit_f32(blk.w), dist, inv_pdf);
            } else {
                const MatF &sm = mats[src];
                const float Tr = expf(-k.sigma_t * t);
                const F3 dv = mk(sm.px, sm.py, sm.pz) - o;
                const float proj = dot(dv, d);
                const F3 perp = fma3(d, -proj, dv);
                const float D = sqrtf(dot(perp, perp));
                const float thA = atan2f(-proj, D), thB = atan2f(t - proj, D);
                const float xi = u32_to_unit_f32(blk.z);
                const float tl = D * tanf((1.0f - xi) * thA + xi * thB);
                dist = tl + proj;
                inv_pdf = fabsf(thB - thA) * (tl * tl + D * D) / (D * (1.0f - Tr));
                const float xs = u32_to_unit_f32(blk.w);
                surface = (METHOD == 1) ? (xs <= Tr) : (xs < Tr);
            }
            if (surface && mats[hid].emits) { // :1308-1313: a directly seen emitter ends the path
                if (depth == 0) L = had(mk(mats[hid].lr, mats[hid].lg, mats[hid].lb), beta);
                if (isfinite(L.x + L.y + L.z)) { acc[0] += L.x; acc[1] += L.y; acc[2] += L.z; } else ++nonfinite;
                active = false; phase = PH_NEW;
                continue;
            }
            blk = philox_block(pixel, sample, (uint32_t)depth, 1, key0, key1);
            const MatF &sm = mats[src];
            if (surface) {
                o = fma3(d, t, o); id = hid; a = 0;
                const F3 lx = mk(sm.px, sm.py, sm.pz) - o;
                // pLight: zero for an area source unless the point lies inside it (see vpt_f32.cuh point_light_direct)
                next = (sm.r > 0.0f && dot(lx, lx) > sm.r * sm.r) ? (sc.n_area > 0 ? NX_AREA : NX_MIS) : NX_POINT;
                phase = PH_SURF_POINT; // "a surface vertex": refined below
            } else {
                o = fma3(d, dist, o);
                w = (METHOD == 0) ? k.albedo_over_cp : k.sigma_s * expf(-k.sigma_t * fabsf(dist)) * inv_pdf * k.inv_cp;
                next = (sm.r == 0.0f) ? NX_POINT : NX_AREA;
                phase = PH_MED_POINT; // "a medium vertex"
            }
        } else if (phase == PH_MIS) {
            // BSDF-sampled direct light (uniform :250 / microfacet :97) and its power-heuristic weight, misSamplingFunctions.h:124-167
            if (hit && mats[hid].emits) {
                const MatF &obj = mats[id]; const MatF &em = mats[hid];
                const F3 n = unit(o - mk(obj.px, obj.py, obj.pz));
                const F3 cx = mk(em.px, em.py, em.pz) - o;
                float omc = one_minus_cos_max(em.r * em.r / dot(cx, cx));
                if (obj.material == 1) {
                    const Frame fr = make_frame(n);
                    const F3 wo_l = unit(to_local(fr, -d)), wi_l = unit(to_local(fr, qd));
                    const F3 wh = unit(wi_l + wo_l);
                    const float gpdf = facet_pdf(wo_l, wh, obj.alpha);
                    const F3 g = had(mk(em.lr, em.lg, em.lb), facet_brdf(obj, wi_l, wh, wo_l)) * (wi_l.z / gpdf);
                    if (!(g.x > 0.0f)) { // the reference's stale costhetaMax: the last area light visited, or cos 0
                        omc = 1.0f;
                        if (sc.n_area > 0) { const MatF &la = mats[sc.area[sc.n_area - 1]]; const F3 c2 = mk(la.px, la.py, la.pz) - o; omc = one_minus_cos_max(la.r * la.r / dot(c2, c2)); }
                    }
                    L = L + had(g, beta) * (power_heuristic(gpdf, 1.0f / (kTwoPi * omc)) * k.inv_cp);
                } else {
                    const F3 g = had(mk(em.lr, em.lg, em.lb), mk(obj.cr, obj.cg, obj.cb));
                    if (g.x > 0.0f && g.y > 0.0f && g.z > 0.0f)
                        L = L + had(g, beta) * (power_heuristic(dot(n, qd) * kInvPi, 1.0f / (kTwoPi * omc)) * k.inv_cp);
                }
            }
            next = NX_FINISH_SURF;
        } else {
            // a next-event ray: success = the wanted light is what the ray reached (area) / nothing before the point (point light)
            const int reached = hit ? hid : (phase == PH_SURF_AREA ? 0 : -1); // reference: id stays 0 on a miss (samplingFunctions.h:196)
            const bool ok = (want >= 0) ? (reached == want) : (!hit || t > lim);
            if (ok) L = L + C * (phase == PH_MED_AREA ? expf(-k.sigma_t * t) : 1.0f);
            if (phase == PH_MED_POINT || phase == PH_MED_AREA) next = NX_FINISH_MED;
            else { if (phase == PH_SURF_AREA) ++a; next = (a < sc.n_area) ? NX_AREA : NX_MIS; }
        }

        // ---- set up the next query of this vertex, or finish the vertex ------------------------------------------------------
        const bool medium = (phase == PH_MED_POINT || phase == PH_MED_AREA);
        if (next == NX_POINT) { // pLight (:62-91) at a surface / the r == 0 branch of (free)SingleScattering in the medium
            const MatF &sm = mats[src];
            const F3 light = mk(sm.px, sm.py, sm.pz);
            const F3 lx = light - o;
            const float d2 = dot(lx, lx), inv = rsqrtf(d2), dist = d2 * inv;
            const float atten = expf(-k.sigma_t * dist) / d2 * k.n_emitters;
            const F3 Le = mk(sm.lr, sm.lg, sm.lb);
            if (medium) {
                C = had(Le, beta) * (atten * kInv4Pi * w);
                phase = PH_MED_POINT;
            } else {
                const MatF &obj = mats[id];
                const F3 n = unit(o - mk(obj.px, obj.py, obj.pz));
                const F3 wi = lx * inv;
                F3 f = mk(obj.cr, obj.cg, obj.cb) * kInvPi;
                if (obj.material == 1) { const Frame fr = make_frame(n); f = brdf_eval(obj, unit(to_local(fr, wi)), unit(to_local(fr, -d))); }
                C = had(had(Le, f), beta) * (dot(n, wi) * atten * k.inv_cp);
                phase = PH_SURF_POINT;
            }
            qo = light; qd = lx * (-inv); lim = dist * (1.0f - 1e-4f); want = -1;
        } else if (next == NX_AREA) { // cone-sampled ray to an area light: muestreoSA (:238) at a surface, the solid-angle block in the medium
            const int lid = medium ? src : sc.area[a];
            const MatF &sm = mats[lid];
            float xi1, xi2;
            if (medium) { xi1 = u32_to_unit_f32(blk.x); xi2 = u32_to_unit_f32(blk.y); }
            else {
                if ((a & 1) == 0) { const uint4 r = philox_block(pixel, sample, (uint32_t)depth, 2 + (a >> 1), key0, key1); xi1 = u32_to_unit_f32(r.x); xi2 = u32_to_unit_f32(r.y); area_hi = make_uint2(r.z, r.w); }
                else { xi1 = u32_to_unit_f32(area_hi.x); xi2 = u32_to_unit_f32(area_hi.y); }
            }
            const F3 cx = mk(sm.px, sm.py, sm.pz) - o;
            const float len2 = dot(cx, cx), inv_len = rsqrtf(len2);
            const float omc_max = one_minus_cos_max(sm.r * sm.r / len2);
            const F3 wi = cone_sample(cx * inv_len, omc_max, xi1, xi2);
            const F3 Le = mk(sm.lr, sm.lg, sm.lb);
            if (medium) {
                C = had(Le, beta) * (kInv4Pi * kTwoPi * omc_max * k.n_emitters * w); // x exp(-sigma_t t) once the hit distance is known
                phase = PH_MED_AREA;
            } else {
                const MatF &obj = mats[id];
                const F3 n = unit(o - mk(obj.px, obj.py, obj.pz));
                const float cos_i = dot(n, wi);
                F3 f = mk(obj.cr, obj.cg, obj.cb) * kInvPi;
                float gpdf = cos_i * kInvPi;
                if (obj.material == 1) {
                    const Frame fr = make_frame(n);
                    const F3 wo_l = unit(to_local(fr, -d)), wi_l = unit(to_local(fr, wi));
                    const F3 wh = unit(wi_l + wo_l);
                    f = facet_brdf(obj, wi_l, wh, wo_l); gpdf = facet_pdf(wo_l, wh, obj.alpha);
                }
                const float inv_fpdf = kTwoPi * omc_max;
                const float wmis = power_heuristic(1.0f / inv_fpdf, gpdf);
                C = had(had(Le, f), beta) * (cos_i * inv_fpdf * expf(-k.sigma_t * len2 * inv_len) * wmis * k.inv_cp);
                phase = PH_SURF_AREA;
            }
            qo = o; qd = wi; want = lid;
        } else if (next == NX_MIS) { // the BSDF-sampled ray of MISv2: slots S_MIS = lanes 2,3 of block 1
            const MatF &obj = mats[id];
            const Frame fr = make_frame(unit(o - mk(obj.px, obj.py, obj.pz)));
            const float xi1 = u32_to_unit_f32(blk.z), xi2 = u32_to_unit_f32(blk.w);
            F3 wi_l;
            if (obj.material == 1) { const F3 wo_l = unit(to_local(fr, -d)); const F3 wh = facet_normal(obj.alpha, xi1, xi2); wi_l = unit(fma3(wh, 2.0f * dot(wh, wo_l), -wo_l)); }
            else wi_l = cosine_local(xi1, xi2);
            qo = o; qd = unit(to_world(fr, wi_l)); phase = PH_MIS;
        } else if (next == NX_FINISH_SURF) { // bdsf (:16-59): slots S_BSDF = lanes 0,1 of block 1
            const MatF &obj = mats[id];
            const Frame fr = make_frame(unit(o - mk(obj.px, obj.py, obj.pz)));
            F3 wi;
            const F3 weight = bsdf_sample(obj, fr, unit(to_local(fr, -d)), u32_to_unit_f32(blk.x), u32_to_unit_f32(blk.y), wi);
            beta = had(beta, weight) * k.inv_cp;
            d = wi; ++depth; phase = PH_NEW;
        } else if (next == NX_FINISH_MED) { // isotropicPhaseSample: slots S_PHASE = lanes 2,3 of block 1
            d = phase_sample(u32_to_unit_f32(blk.z), u32_to_unit_f32(blk.w));
            beta = beta * w;
            ++depth; phase = PH_NEW;
        }
    }
}

} // namespace f32
} // namespace vpt
