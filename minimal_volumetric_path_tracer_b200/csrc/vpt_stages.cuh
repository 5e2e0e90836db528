// vpt_stages.cuh -- THE estimators of the FP32 path, once: the reference's three shade methods (iterativeVPTracerFree vptShadeMethods.h:1263-1340,
// explicitVPTracerRecursive :1014-1149, MISVPTTracerRecursive :1345-1481, in throughput form) plus VPT_METHOD_MIS_DISTANCE, cut into the STAGES
// between which a path's record is re-queued.  Every FP32 kernel runs exactly this code --
//   * the product kernel (vpt_smwave.cuh: records in shared memory, stage batches claimed by warps),
//   * the HBM wavefront (vpt_hbmwave.cuh: records in HBM queues, one kernel per stage),
//   * the megakernel and the unit kernels (vpt_kernels_f32.cu: one thread drives one record through the stages, trace_path below) --
// through a context type C that supplies the record's random numbers, the scene scan, the radiance sink and the counters:
//   float4 C::rnd(const Rec &, uint32_t block)   uniforms of Philox block `block` of (pixel, sample, bounce = rec.depth)    (vpt_philox.cuh slot table)
//   float4 C::jitter(const Rec &)                the pixel-jitter block (pseudo-bounce 0xffffffff)
//   bool   C::scan(F3 o, F3 d, float &t, int &id)  nearest accepted hit (vpt_scan.cuh); the stages count the scans of their active lanes
//   void   C::add(const Rec &, F3 contribution)  radiance arriving at the record's pixel (rt.cpp:794: the sum over the path's vertices)
//   void   C::last_step()                        called once per stage before its last block of work (the product kernel claims its next batch here)
//   const SmScene &S;  const ConstsF &k;  unsigned events, scans;
//
// A stage takes `act` (false: an idle lane that only keeps the warp's scans in step; its record holds in-range dummies) and returns where
// the record goes next: a stage (SQ_*), kDestFree (the path ended) or -1 (idle lane).
//
// Semantics = the reference with its two FP64-rounding-decided behaviours replaced by their well-defined alternative (include/vpt.h
// VPT_QUIRK_*): r == 0 spheres are never ray-intersected and visibility uses `t > distance * (1 - 1e-4)`.  Radiance is not carried in the
// record: every vertex's direct-light terms go to the pixel sum when they are computed (the estimators only ever ADD to L:
// vptShadeMethods.h:1321,1335) -- a record is origin, direction, throughput, its Philox counter and 3 words of stage hand-over.
#pragma once
#include "vpt_scan.cuh"

namespace vpt {
namespace f32 {

struct Rec {
    F3 o, d, beta;             // ray origin (the current vertex once PRIMARY has placed it), direction, throughput
    uint32_t pixel, sample;    // Philox counter words 0, 1
    uint32_t depth;            // bounce = Philox counter word 2
    uint32_t src, hid;         // picked source (drawn with the roulette, slot 1), hit object (set by PRIMARY for surface vertices): sphere indices
    float xi_dist, xi_decide;  // slots 2, 3 of the bounce (distance sample, surface-or-medium decision), drawn with the roulette
    uint32_t aux;              // the context's own (the product kernel: pixel-in-item and item slot)
};

// exp(-x) of every transmittance on this path: ex2.approx (<= 2 ulp + the rounding of x log2 e: <= 1.5e-6 relative for x <= 20;
// VPT_UNIT_TRANSMITTANCE / VPT_UNIT_FREE_FLIGHT evaluate exactly this function)
__device__ __forceinline__ float transmit(float x) { return __expf(-x); }

// ---- roulette for the next bounce (vptShadeMethods.h:1282; also at depth 0) + the draws that come with it --------------------------------
// the record already holds the new origin / direction / throughput / depth; block 0 of the new bounce = roulette, light pick, distance, decision
template <class C>
__device__ __forceinline__ int roulette_with(C &c, bool act, Rec &r, float4 u) {
    bool alive = false;
    if (act) {
        alive = !(c.k.max_depth > 0 && (int)r.depth >= c.k.max_depth) && r.depth < (uint32_t)VPT_MAX_DEPTH && !(u.x < c.k.q);
        if (alive) {
            r.src = (uint32_t)c.S.emitters[min((int)(u.y * c.k.n_emitters), c.S.n_emitters - 1)]; // :1293-1304
            r.xi_dist = u.z; r.xi_decide = u.w;
        }
    }
    return alive ? SQ_PRIMARY : (act ? kDestFree : -1);
}
template <class C>
__device__ __forceinline__ int roulette(C &c, bool act, Rec &r) {
    c.last_step();
    float4 u = make_float4(0, 0, 0, 0);
    if (act) u = c.rnd(r, 0);
    return roulette_with(c, act, r, u);
}

// camera ray of storage pixel `pixel` with jitter (j1, j2), rt.cpp:773,787 (storage row 0 is the top of the image)
__device__ __forceinline__ F3 camera_dir(const ConstsF &k, uint32_t pixel, int width, int height, float j1, float j2) {
    // row = pixel / width without the integer-division sequence where the pixel index is exact in fp32: a float estimate (off by less than
    // one row: relative error 1.2e-7 of at most 2^24) corrected by one step
    int row, col;
    if (pixel < (1u << 24)) {
        row = (int)((float)pixel * k.inv_w); col = (int)pixel - row * width;
        if (col < 0) { --row; col += width; } else if (col >= width) { ++row; col -= width; }
    } else { row = (int)(pixel / (unsigned)width); col = (int)pixel - row * width; }
    const float fx = (float)col, fy = (float)(height - 1 - row);
    const float u = (fx + j1 - 0.5f) * k.inv_w - 0.5f, v = (fy + j2 - 0.5f) * k.inv_h - 0.5f;
    return unit(mk(fmaf(k.cam_cx[0], u, fmaf(k.cam_cy[0], v, k.cam_d[0])), fmaf(k.cam_cx[1], u, fmaf(k.cam_cy[1], v, k.cam_d[1])),
                   fmaf(k.cam_cx[2], u, fmaf(k.cam_cy[2], v, k.cam_d[2]))));
}

// ---- GEN: camera sample (pixel, sample) -> record at the camera, or nothing if the first roulette ends the path (rt.cpp:773-794) ---------
template <class C>
__device__ __forceinline__ bool stage_gen(C &c, bool mine, uint32_t pixel, uint32_t sample, int width, int height, Rec &r) {
    r.pixel = pixel; r.sample = sample; r.depth = 0u;
    bool alive = false;
    if (mine) {
        const float4 u = c.rnd(r, 0);
        alive = !(u.x < c.k.q);
        if (alive) {
            const float4 j = c.jitter(r);
            r.d = camera_dir(c.k, pixel, width, height, j.x, j.y);
            r.o = mk(c.k.cam_o[0], c.k.cam_o[1], c.k.cam_o[2]);
            r.beta = mk(1.0f, 1.0f, 1.0f);
            r.src = (uint32_t)c.S.emitters[min((int)(u.y * c.k.n_emitters), c.S.n_emitters - 1)];
            r.xi_dist = u.z; r.xi_decide = u.w; r.hid = 0u;
        }
    }
    return alive;
}

// K camera samples per lane as straight-line code: the same operations on the same values as stage_gen, per sample, but nothing is
// branched around -- the K Philox / camera chains are independent and interleave (the kernel is latency bound), and a lane whose sample
// dies at the roulette (40 % at the default continue probability) would idle through its neighbours' work anyway.
template <int K, class C>
__device__ __forceinline__ void stage_gen_k(C &c, const bool *mine, const uint32_t *pixel, const uint32_t *sample, int width, int height, Rec *r, bool *alive) {
    float4 u[K], j[K];
#pragma unroll
    for (int h = 0; h < K; ++h) { r[h].pixel = pixel[h]; r[h].sample = sample[h]; r[h].depth = 0u; }
#pragma unroll
    for (int h = 0; h < K; ++h) u[h] = c.rnd(r[h], 0);
#pragma unroll
    for (int h = 0; h < K; ++h) j[h] = c.jitter(r[h]);
#pragma unroll
    for (int h = 0; h < K; ++h) {
        alive[h] = mine[h] && !(u[h].x < c.k.q);
        r[h].d = camera_dir(c.k, pixel[h], width, height, j[h].x, j[h].y);
        r[h].o = mk(c.k.cam_o[0], c.k.cam_o[1], c.k.cam_o[2]);
        r[h].beta = mk(1.0f, 1.0f, 1.0f);
        r[h].src = (uint32_t)c.S.emitters[min((int)(u[h].y * c.k.n_emitters), c.S.n_emitters - 1)];
        r[h].xi_dist = u[h].z; r[h].xi_decide = u[h].w; r[h].hid = 0u;
    }
}

// ---- PRIMARY: scan of the path ray, distance sampling, surface-or-medium decision --------------------------------------------------------
template <int METHOD, class C>
__device__ __forceinline__ int stage_primary(C &c, bool act, Rec &r) {
    float t; int hid;
    const bool hit = c.scan(r.o, r.d, t, hid);
    int dest = -1;
    if (act) {
        ++c.events; ++c.scans;
        if (!hit) { t = kMaxFloat; hid = 0; } // :1287 (id stays 0)
        const MatF &sm = c.S.mats[r.src];
        bool surface; float dist, inv_pdf = 1.0f;
        if (METHOD == 0) {
            dist = -logf(1.0f - r.xi_dist) * c.k.inv_sigma_t; // freeFlightSample, vptSamplingFunctions.h:11
            surface = dist > t;
        } else if (METHOD == 4) { // distance-sampling MIS (vpt_f32.cuh mis_distance)
            surface = mis_distance(mk(sm.px, sm.py, sm.pz), r.o, r.d, t, transmit(c.k.sigma_t * t), c.k.sigma_t, c.k.inv_sigma_t, r.xi_dist, r.xi_decide, dist, inv_pdf);
        } else { // equiAngularParams2 (volumetricBasicFunctions.h:209-223) + equiAngularProb (vptSamplingFunctions.h:60)
            const float Tr = transmit(c.k.sigma_t * t); // TrActual :1046 / psurf :1407 (0 on a miss)
            float D, dth, tl;
            dist = equiangular_sample(mk(sm.px, sm.py, sm.pz), r.o, r.d, t, r.xi_dist, D, dth, tl);
            inv_pdf = dth * fmaf(tl, tl, D * D) / (D * (1.0f - Tr));
            surface = (METHOD == 1) ? (r.xi_decide <= Tr) : (r.xi_decide < Tr); // :1096 / :1426
        }
        const MatF &obj = c.S.mats[hid];
        if (surface && obj.emits) { // :1308-1313: a directly seen emitter ends the path; it counts only at depth 0
            if (r.depth == 0u) c.add(r, had(mk(obj.lr, obj.lg, obj.lb), r.beta));
            dest = kDestFree;
        } else if (surface) {
            r.o = fma3(r.d, t, r.o);
            r.hid = (uint32_t)hid;
            const F3 lx = mk(sm.px, sm.py, sm.pz) - r.o;
            const bool to_sp = !(sm.r > 0.0f && dot(lx, lx) > sm.r * sm.r); // pLight is zero for an area source seen from outside it (its own sphere blocks the centre)
            dest = to_sp ? SQ_SURF_P : (obj.material != 0 ? SQ_SURF_F : SQ_SURF_L);
        } else { // medium vertex: throughput factor (sigma_s / sigma_t) / cp (:1335) or sigma_s T / (pdf cp) (:1130)
            r.o = fma3(r.d, dist, r.o);
            const float w = (METHOD == 0) ? c.k.albedo_over_cp : c.k.sigma_s * transmit(c.k.sigma_t * fabsf(dist)) * inv_pdf * c.k.inv_cp;
            r.beta = r.beta * w;
            dest = sm.r == 0.0f ? SQ_MED_POINT : SQ_MED_AREA;
        }
    }
    return dest;
}

// ---- MED: (free)SingleScattering (volumetricBasicFunctions.h:284-340 / :225-281) + isotropicPhaseSample (vptSamplingFunctions.h:34) + roulette
// r.beta already carries the vertex's factor; 1 / probSource = n_emitters.  POINT: the source is a point light (shadow ray from the light
// to the vertex), else an area light (cone-sampled ray towards its sphere; counts if it is the first thing hit).
template <bool POINT, class C>
__device__ __forceinline__ int stage_med(C &c, bool act, Rec &r) {
    const float4 b1 = c.rnd(r, 1);
    const int src = (int)r.src;
    const MatF &sm = c.S.mats[src];
    const F3 light = mk(sm.px, sm.py, sm.pz);
    const F3 lx = light - r.o;
    const float d2 = dot(lx, lx), inv = rsqrtf(d2);
    F3 qo, qd, Cn; float lim = 0.0f;
    if (POINT) {
        const float dist = d2 * inv;
        Cn = had(mk(sm.lr, sm.lg, sm.lb), r.beta) * (transmit(c.k.sigma_t * dist) / d2 * c.k.n_emitters * kInv4Pi);
        qo = light; qd = lx * (-inv); lim = dist * (1.0f - 1e-4f);
    } else {
        const float omc_max = one_minus_cos_max(sm.r * sm.r / d2);
        qd = cone_sample(lx * inv, omc_max, b1.x, b1.y);
        qo = r.o;
        Cn = had(mk(sm.lr, sm.lg, sm.lb), r.beta) * (kInv4Pi * kTwoPi * omc_max * c.k.n_emitters);
    }
    float t; int hid;
    const bool hit = POINT ? scan_sm_light(c.S, src, qo, qd, t, hid) : c.scan(qo, qd, t, hid);
    if (act) {
        ++c.scans;
        if (POINT) { if (!hit || t > lim) c.add(r, Cn); }
        else if (hit && hid == src) c.add(r, Cn * transmit(c.k.sigma_t * t));
    }
    r.d = phase_sample(b1.z, b1.w);
    r.depth += 1u;
    return roulette(c, act, r);
}

// ---- SURF_P: pLight (vptShadeMethods.h:62-91) x transmittance / probSource as used at :1316 / :1113 / :1444 ------------------------------
// microfacet BRDF for world-space directions (rare: kept out of line so that the Lambert stages stay small)
static __device__ __noinline__ F3 facet_eval_world(const MatF &obj, F3 n_, F3 wi, F3 d) {
    const Frame fr = make_frame(n_);
    return brdf_eval(obj, unit(to_local(fr, wi)), unit(to_local(fr, -d)));
}
template <class C>
__device__ __forceinline__ int stage_surf_p(C &c, bool act, Rec &r) {
    const MatF &sm = c.S.mats[r.src];
    const MatF &obj = c.S.mats[r.hid];
    const F3 light = mk(sm.px, sm.py, sm.pz);
    const F3 lx = light - r.o;
    const float d2 = dot(lx, lx), inv = rsqrtf(d2), dist = d2 * inv;
    const F3 n_ = unit(r.o - mk(obj.px, obj.py, obj.pz));
    const F3 wi = lx * inv;
    F3 f = mk(obj.cr, obj.cg, obj.cb) * kInvPi;
    if (obj.material == 1) f = facet_eval_world(obj, n_, wi, r.d);
    const F3 Cn = had(had(mk(sm.lr, sm.lg, sm.lb), f), r.beta) * (dot(n_, wi) * transmit(c.k.sigma_t * dist) / d2 * c.k.n_emitters * c.k.inv_cp);
    float t; int hid;
    const bool hit = scan_sm_light(c.S, (int)r.src, light, lx * (-inv), t, hid);
    c.scans += act ? 1u : 0u;
    if (act && (!hit || t > dist * (1.0f - 1e-4f))) c.add(r, Cn);
    return act ? (obj.material != 0 ? SQ_SURF_F : SQ_SURF_L) : -1;
}

// ---- SURF: MISv2 (misSamplingFunctions.h:96-170) + bdsf (vptShadeMethods.h:16-59) + roulette ----------------------------------------------
// FACET = false: Lambert (material 0); true: Beckmann conductor (1) and the dielectric as written in the reference (2).
// The vertex's next-event rays -- one cone-sampled ray per area light (muestreoSA) and the BSDF-sampled ray -- all start at the vertex:
// they are scanned together, two area lights per pass and the last one or two together with the BSDF-sampled ray (scan_sm_n: one pass
// over the spheres, origin part of every sphere test shared).  The terms are added in the reference's order (lights, then the BSDF term).
struct LightRay { F3 wi; float omc_max, len2, inv_len; int lid; };
template <bool FACET, class C>
__device__ __forceinline__ int stage_surf(C &c, bool act, Rec &r) {
    const F3 o = r.o, d = r.d, beta = r.beta;
    const MatF &obj = c.S.mats[r.hid];
    const F3 n_ = unit(o - mk(obj.px, obj.py, obj.pz));
    const Frame fr = make_frame(n_);
    const F3 wo_l = FACET ? unit(to_local(fr, -d)) : mk(0, 0, 1);
    const F3 albedo = mk(obj.cr, obj.cg, obj.cb);
    F3 L = mk(0.0f, 0.0f, 0.0f); // this vertex's direct light, before throughput and 1 / cp
    // material 2 shares this stage with the microfacet: its light-sampled terms are zero (samplingFunctions.h:190), its lanes only run the
    // light scans in step with the others
    const bool diel = FACET && obj.material == 2;
    DielF di; di.F = 0.0f; di.wr = di.wt = mk(0, 0, 1);
    if (diel) di = dielectric_setup(wo_l);
    const int n_area = c.S.n_area;
    float omc_tail = 1.0f; // (FACET: 1 - costhetaMax of the last light of the ray-by-ray loop)

    // muestreoSA (misSamplingFunctions.h:105-118): aim at area light a, and what a hit of it contributes
    auto aim = [&](int a, float xi1, float xi2) -> LightRay {
        LightRay q;
        q.lid = c.S.area[a];
        const MatF &sm = c.S.mats[q.lid];
        const F3 cx = mk(sm.px, sm.py, sm.pz) - o;
        q.len2 = dot(cx, cx); q.inv_len = rsqrtf(q.len2);
        q.omc_max = one_minus_cos_max(sm.r * sm.r / q.len2);
        q.wi = cone_sample(cx * q.inv_len, q.omc_max, xi1, xi2);
        return q;
    };
    auto shade_light = [&](const LightRay &q, bool hit, int hid) {
        c.scans += act ? 1u : 0u;
        if (act && (hit ? hid : 0) == q.lid && !diel) { // id stays 0 on a miss, samplingFunctions.h:196
            const MatF &sm = c.S.mats[q.lid];
            const float cos_i = dot(n_, q.wi);
            F3 f = albedo * kInvPi;
            float gpdf = cos_i * kInvPi;
            if (FACET) { const F3 wi_l = unit(to_local(fr, q.wi)); const F3 wh = unit(wi_l + wo_l); f = facet_brdf(obj, wi_l, wh, wo_l); gpdf = facet_pdf(wo_l, wh, obj.alpha); }
            const float inv_fpdf = kTwoPi * q.omc_max;
            const float wmis = power_heuristic(1.0f / inv_fpdf, gpdf);
            L = L + had(mk(sm.lr, sm.lg, sm.lb), f) * (cos_i * inv_fpdf * transmit(c.k.sigma_t * q.len2 * q.inv_len) * wmis);
        }
    };

    // (the rare microfacet / dielectric stage scans ray by ray: one copy of the scan instead of four keeps its code small)
    constexpr bool kFuse = !FACET;
    int a = 0;
    if (!kFuse) {
        float4 ra = make_float4(0, 0, 0, 0);
        for (; a < n_area; ++a) {
            if ((a & 1) == 0) ra = c.rnd(r, 2 + (a >> 1));
            const LightRay q = aim(a, (a & 1) ? ra.z : ra.x, (a & 1) ? ra.w : ra.y);
            float t; int hid;
            const bool hit = c.scan(o, q.wi, t, hid);
            shade_light(q, hit, hid);
            if (a == n_area - 1) omc_tail = q.omc_max;
        }
    }
    while (kFuse && n_area - a > 2) { // two area lights per pass while more than two are left (their cone numbers share one Philox block)
        const float4 ra = c.rnd(r, 2 + (a >> 1));
        const LightRay q0 = aim(a, ra.x, ra.y), q1 = aim(a + 1, ra.z, ra.w);
        RaysN<2> rays;
        rays.d[0] = q0.wi; rays.d[1] = q1.wi;
        scan_sm_n<2>(c.S, o, rays);
        shade_light(q0, rays.hit[0], rays.id[0]);
        shade_light(q1, rays.hit[1], rays.id[1]);
        a += 2;
    }
    // the last pass: the remaining 0, 1 or 2 area lights and the BSDF-sampled ray of MISv2 (:124-167)
    const int left = n_area - a;
    LightRay q0, q1;
    q0.wi = q1.wi = mk(0, 0, 1); q0.omc_max = q1.omc_max = 1.0f; q0.len2 = q1.len2 = q0.inv_len = q1.inv_len = 1.0f; q0.lid = q1.lid = -1;
    if (left > 0) {
        const float4 ra = c.rnd(r, 2 + (a >> 1));
        q0 = aim(a, ra.x, ra.y);
        if (left > 1) q1 = aim(a + 1, ra.z, ra.w);
    }
    const float omc_last = n_area == 0 ? 1.0f : (left > 1 ? q1.omc_max : (left > 0 ? q0.omc_max : omc_tail)); // 1 - costhetaMax of the last light visited (:162); 1 = "cos 0"
    const float4 b1 = c.rnd(r, 1);
    // Lambert: the next bounce's roulette block is drawn here, next to the vertex's own block -- two independent Philox chains side by side
    // (+0.3 %; the microfacet stage has no registers to spare for it)
    float4 un = make_float4(0, 0, 0, 0);
    if (!FACET) { Rec rn = r; rn.depth += 1u; un = c.rnd(rn, 0); }
    // slots S_MIS = lanes 2, 3 of block 1
    const float xi1 = b1.z, xi2 = b1.w;
    F3 wi_l, wh = mk(0, 0, 1);
    float gpdf_loop = 0.0f; // the pdf the reference's light loop leaves behind for the dielectric's BSDF term (misSamplingFunctions.h:110-118,148)
    bool refracted = false;
    if (FACET) {
        wh = facet_normal(obj.alpha, xi1, xi2); wi_l = unit(fma3(wh, 2.0f * dot(wh, wo_l), -wo_l));
        if (diel) { // softDielectric (samplingFunctions.h:209-235): reflect with probability F, else the reference's refraction
            if (n_area > 0) {
                const uint32_t slot = S_DIEL + (uint32_t)n_area - 1u;
                const float4 bd = c.rnd(r, slot >> 2);
                const float xg = (slot & 3u) == 0u ? bd.x : (slot & 3u) == 1u ? bd.y : (slot & 3u) == 2u ? bd.z : bd.w;
                gpdf_loop = xg > di.F ? 1.0f - di.F : di.F;
            }
            refracted = !(xi1 < di.F);
            wi_l = refracted ? di.wt : di.wr;
        }
    } else wi_l = cosine_local(xi1, xi2);
    const F3 wi_b = unit(to_world(fr, wi_l));
    bool hit_b; int hid_b;
    if (kFuse && left == 2) {
        RaysN<3> rays;
        rays.d[0] = q0.wi; rays.d[1] = q1.wi; rays.d[2] = wi_b;
        scan_sm_n<3>(c.S, o, rays);
        shade_light(q0, rays.hit[0], rays.id[0]);
        shade_light(q1, rays.hit[1], rays.id[1]);
        hit_b = rays.hit[2]; hid_b = rays.id[2];
    } else if (kFuse && left == 1) {
        RaysN<2> rays;
        rays.d[0] = q0.wi; rays.d[1] = wi_b;
        scan_sm_n<2>(c.S, o, rays);
        shade_light(q0, rays.hit[0], rays.id[0]);
        hit_b = rays.hit[1]; hid_b = rays.id[1];
    } else { float t_b; hit_b = c.scan(o, wi_b, t_b, hid_b); }
    c.scans += act ? 1u : 0u;
    if (act && hit_b && c.S.mats[hid_b].emits) {
        const MatF &em = c.S.mats[hid_b];
        const F3 cx = mk(em.px, em.py, em.pz) - o;
        float omc = one_minus_cos_max(em.r * em.r / dot(cx, cx));
        if (diel) L = L + dielectric_direct(em, o, wi_l.z, refracted, gpdf_loop);
        else if (FACET) {
            const float gpdf = facet_pdf(wo_l, wh, obj.alpha);
            const F3 g = had(mk(em.lr, em.lg, em.lb), facet_brdf(obj, wi_l, wh, wo_l)) * (wi_l.z / gpdf);
            if (!(g.x > 0.0f)) omc = omc_last; // the reference's stale costhetaMax (:162)
            L = L + g * power_heuristic(gpdf, 1.0f / (kTwoPi * omc));
        } else {
            const F3 g = had(mk(em.lr, em.lg, em.lb), albedo); // Le c/pi cos / (cos/pi)
            if (g.x > 0.0f && g.y > 0.0f && g.z > 0.0f) L = L + g * power_heuristic(dot(n_, wi_b) * kInvPi, 1.0f / (kTwoPi * omc));
        }
    }
    if (act) c.add(r, had(L, beta) * c.k.inv_cp); // :1321
    F3 wi, weight; // bdsf (:16-59): slots S_BSDF = lanes 0,1 of block 1
    if (FACET) weight = bsdf_sample(obj, fr, wo_l, b1.x, b1.y, wi);
    else { wi = unit(to_world(fr, cosine_local(b1.x, b1.y))); weight = albedo; } // c/pi * cos / (cos/pi)
    r.beta = had(beta, weight) * c.k.inv_cp; // :1326
    r.d = wi;
    r.depth += 1u;
    if (!FACET) { c.last_step(); return roulette_with(c, act, r, un); }
    return roulette(c, act, r);
}

// ---- one thread drives one (alive) record through its stages until the path ends: the megakernel and the unit kernels ---------------------
// `c.vertex(r, dest)` is told where PRIMARY sent the record before that stage runs (contexts with explicit random-number lists arrange
// the vertex's draws then).
template <int METHOD, class C>
__device__ __forceinline__ void trace_path(C &c, Rec &r) {
    int dest = SQ_PRIMARY;
    while (dest == SQ_PRIMARY) {
        dest = stage_primary<METHOD>(c, true, r);
        c.vertex(r, dest);
        if (dest == SQ_SURF_P) dest = stage_surf_p(c, true, r);
        if (dest == SQ_MED_POINT) dest = stage_med<true>(c, true, r);
        else if (dest == SQ_MED_AREA) dest = stage_med<false>(c, true, r);
        else if (dest == SQ_SURF_L) dest = stage_surf<false>(c, true, r);
        else if (dest == SQ_SURF_F) dest = stage_surf<true>(c, true, r);
    }
}
template <class C>
__device__ __forceinline__ void trace_path_method(int method, C &c, Rec &r) {
    if (method == 0) trace_path<0>(c, r);
    else if (method == 1) trace_path<1>(c, r);
    else if (method == 4) trace_path<4>(c, r);
    else trace_path<2>(c, r);
}

} // namespace f32
} // namespace vpt
