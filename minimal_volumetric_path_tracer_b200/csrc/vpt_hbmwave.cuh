// vpt_hbmwave.cuh -- the classic multi-kernel wavefront (VPT_KERNEL_WAVEFRONT_HBM): path state in HBM queues, one kernel per stage.
//
// north_star: "separate extend / medium-sample / NEE-shadow / scatter kernels compact SoA path-state queues with warp ballot / prefix
// scans ... achieved HBM GB/s on the wavefront queues".  Built as the measured alternative to the on-chip wavefront (vpt_smwave.cuh):
//   * a path record (72 B: origin, direction, throughput, radiance so far, pixel, sample, depth, three Philox words) lives IN the queue of
//     the stage that has to process it next: structure-of-arrays per queue, so that both the reads of a stage (thread i reads entry i) and
//     its writes (compacted per block: match.any groups the lanes of a warp by destination queue, the groups take offsets from
//     shared-memory counters, ONE global atomic per block and destination reserves the entries -- per-warp global atomics on the six
//     queue counters serialise in L2 and were the bottleneck of the first version) are coalesced;
//   * one ROUND = generate (tops the pool up with new camera samples) -> PRIMARY -> SURF_P -> MED_POINT -> MED_AREA -> SURF_L -> SURF_F;
//     every kernel is a grid-stride loop over its input queue, whose length it reads from device memory, so the host enqueues rounds
//     without synchronising; no stage writes the queue it reads, and the kernel that follows a stage resets that stage's counter;
//   * per-pixel sums: 64-bit fixed-point (2^-30) global atomics (RED.64 in L2): order independent, bit-reproducible, same value as the
//     other kernels' sums up to the fixed-point rounding;
//   * stage arithmetic, random-number slots and semantics are those of vpt_smwave.cuh (file:line citations in vpt_f32.cuh).
#pragma once
#include "vpt_smwave.cuh"

namespace vpt {
namespace f32 {

enum : int { HF_OX = 0, HF_OY, HF_OZ, HF_DX, HF_DY, HF_DZ, HF_BR, HF_BG, HF_BB, HF_LR, HF_LG, HF_LB, HF_FLOATS };
enum : int { HU_PIXEL = 0, HU_SAMPLE, HU_DEPTH, HU_R1, HU_R2, HU_R3, HU_WORDS };
constexpr int kHbmThreads = 256;

struct HbmState { // device pointers and sizes, passed by value
    float *f;                  // [SQ_COUNT][HF_FLOATS][cap]
    uint32_t *u;               // [SQ_COUNT][HU_WORDS][cap]
    unsigned *count;           // [SQ_COUNT] queue lengths; [8] samples to generate this round; [9] rounds that found nothing to do
    unsigned long long *cursor; // [0] next camera sample index of this launch, [1] first sample index of this round's generation
    unsigned long long *acc;   // [n_pixels][3] fixed-point sums
    unsigned long long *tally; // events, scans, nonfinite, paths
    unsigned cap;              // records per queue = paths in flight
    unsigned n_owned_pixels;   // pixels this launch renders (tile shard)
    unsigned long long n_total; // camera samples of this launch = n_owned_pixels * samples
};

struct Rec {
    F3 o, d, beta, L;
    uint32_t pixel, sample, depth, r1, r2, r3;
};

__device__ __forceinline__ Rec hbm_load(const HbmState &H, int q, unsigned i) {
    const float *f = H.f + (size_t)q * HF_FLOATS * H.cap + i;
    const uint32_t *u = H.u + (size_t)q * HU_WORDS * H.cap + i;
    const size_t c = H.cap;
    Rec r;
    r.o = mk(f[HF_OX * c], f[HF_OY * c], f[HF_OZ * c]); r.d = mk(f[HF_DX * c], f[HF_DY * c], f[HF_DZ * c]);
    r.beta = mk(f[HF_BR * c], f[HF_BG * c], f[HF_BB * c]); r.L = mk(f[HF_LR * c], f[HF_LG * c], f[HF_LB * c]);
    r.pixel = u[HU_PIXEL * c]; r.sample = u[HU_SAMPLE * c]; r.depth = u[HU_DEPTH * c]; r.r1 = u[HU_R1 * c]; r.r2 = u[HU_R2 * c]; r.r3 = u[HU_R3 * c];
    return r;
}
__device__ __forceinline__ void hbm_store(const HbmState &H, int q, unsigned i, const Rec &r) {
    float *f = H.f + (size_t)q * HF_FLOATS * H.cap + i;
    uint32_t *u = H.u + (size_t)q * HU_WORDS * H.cap + i;
    const size_t c = H.cap;
    f[HF_OX * c] = r.o.x; f[HF_OY * c] = r.o.y; f[HF_OZ * c] = r.o.z; f[HF_DX * c] = r.d.x; f[HF_DY * c] = r.d.y; f[HF_DZ * c] = r.d.z;
    f[HF_BR * c] = r.beta.x; f[HF_BG * c] = r.beta.y; f[HF_BB * c] = r.beta.z; f[HF_LR * c] = r.L.x; f[HF_LG * c] = r.L.y; f[HF_LB * c] = r.L.z;
    u[HU_PIXEL * c] = r.pixel; u[HU_SAMPLE * c] = r.sample; u[HU_DEPTH * c] = r.depth; u[HU_R1 * c] = r.r1; u[HU_R2 * c] = r.r2; u[HU_R3 * c] = r.r3;
}
__device__ __forceinline__ void hbm_add_radiance(const HbmState &H, uint32_t pixel, F3 L, unsigned &nonfinite) {
    if (!isfinite(L.x + L.y + L.z)) { ++nonfinite; return; }
    unsigned long long *a = H.acc + (size_t)pixel * 3;
    if (L.x != 0.0f) atomicAdd(a + 0, (unsigned long long)__float2ll_rn(L.x * kSmFixScale));
    if (L.y != 0.0f) atomicAdd(a + 1, (unsigned long long)__float2ll_rn(L.y * kSmFixScale));
    if (L.z != 0.0f) atomicAdd(a + 2, (unsigned long long)__float2ll_rn(L.z * kSmFixScale));
}

struct HbmCtx { // what every stage needs
    const SmScene &S;
    const SceneF &sc;
    const ConstsF &k;
    const LaunchParams &lp;
    const HbmState &H;
    unsigned events = 0, scans = 0, nonfinite = 0, paths = 0;
    int dest = -1; // where the record of this lane goes next (SQ_*), -1: nowhere; the kernel's block-wide emit step writes it
    __device__ HbmCtx(const SmScene &S_, const SceneF &sc_, const ConstsF &k_, const LaunchParams &lp_, const HbmState &H_) : S(S_), sc(sc_), k(k_), lp(lp_), H(H_) {}

    // roulette for the next bounce (vptShadeMethods.h:1282); survivors go back to PRIMARY with the block-0 words of the new bounce
    __device__ __forceinline__ void continue_or_end(bool act, Rec &r) {
        bool alive = false;
        if (act) {
            const uint4 b0 = philox_block(r.pixel, r.sample, r.depth, 0, lp.key0, lp.key1);
            alive = !(k.max_depth > 0 && (int)r.depth >= k.max_depth) && !(u32_to_unit_f32(b0.x) < k.q);
            if (alive) { r.r1 = b0.y; r.r2 = b0.z; r.r3 = b0.w; }
            else hbm_add_radiance(H, r.pixel, r.L, nonfinite);
        }
        dest = alive ? SQ_PRIMARY : -1;
    }

    // ---- PRIMARY ----
    template <int METHOD>
    __device__ __forceinline__ void primary(bool act, Rec &r) {
        float t; int hid;
        const bool hit = scan_sm(S, r.o, r.d, t, hid);
        dest = -1;
        if (act) {
            ++scans; ++events;
            if (!hit) { t = kMaxFloat; hid = 0; }
            const int pick = min((int)(u32_to_unit_f32(r.r1) * k.n_emitters), sc.n_emitters - 1);
            const int src = sc.emitters[pick];
            const MatF &sm = S.mats[src];
            bool surface; float dist, inv_pdf = 1.0f;
            if (METHOD == 0) {
                dist = -logf(1.0f - u32_to_unit_f32(r.r2)) * k.inv_sigma_t;
                surface = dist > t;
            } else if (METHOD == 4) { // distance-sampling MIS (vpt_f32.cuh mis_distance)
                surface = mis_distance(mk(sm.px, sm.py, sm.pz), r.o, r.d, t, __expf(-k.sigma_t * t), k.sigma_t, k.inv_sigma_t, u32_to_unit_f32(r.r2), u32_to_unit_f32(r.r3), dist, inv_pdf);
            } else {
                const float Tr = __expf(-k.sigma_t * t);
                float D, dth, tl;
                dist = equiangular_sample(mk(sm.px, sm.py, sm.pz), r.o, r.d, t, u32_to_unit_f32(r.r2), D, dth, tl);
                inv_pdf = dth * (tl * tl + D * D) / (D * (1.0f - Tr));
                const float xs = u32_to_unit_f32(r.r3);
                surface = (METHOD == 1) ? (xs <= Tr) : (xs < Tr);
            }
            const MatF &obj = S.mats[hid];
            if (surface && obj.emits) {
                const F3 L = r.depth == 0 ? had(mk(obj.lr, obj.lg, obj.lb), r.beta) : r.L;
                hbm_add_radiance(H, r.pixel, L, nonfinite);
            } else if (surface) {
                r.o = fma3(r.d, t, r.o);
                const F3 lx = mk(sm.px, sm.py, sm.pz) - r.o;
                const bool to_sp = !(sm.r > 0.0f && dot(lx, lx) > sm.r * sm.r);
                dest = to_sp ? SQ_SURF_P : (obj.material != 0 ? SQ_SURF_F : SQ_SURF_L);
                r.r1 = (uint32_t)src | ((uint32_t)hid << 8);
            } else {
                r.o = fma3(r.d, dist, r.o);
                const float w = (METHOD == 0) ? k.albedo_over_cp : k.sigma_s * __expf(-k.sigma_t * fabsf(dist)) * inv_pdf * k.inv_cp;
                r.r2 = __float_as_uint(w);
                r.r1 = (uint32_t)src;
                dest = sm.r == 0.0f ? SQ_MED_POINT : SQ_MED_AREA;
            }
        }
    }

    // ---- MED ----
    template <bool POINT>
    __device__ __forceinline__ void med(bool act, Rec &r) {
        const float w = __uint_as_float(r.r2);
        const int src = act ? (int)(r.r1 & 0xffu) : 0;
        const uint4 b1 = philox_block(r.pixel, r.sample, r.depth, 1, lp.key0, lp.key1);
        const MatF &sm = S.mats[src];
        const F3 light = mk(sm.px, sm.py, sm.pz);
        const F3 lx = light - r.o;
        const float d2 = dot(lx, lx), inv = rsqrtf(d2);
        F3 qo, qd, C; float lim = 0.0f;
        if (POINT) {
            const float dist = d2 * inv;
            C = had(mk(sm.lr, sm.lg, sm.lb), r.beta) * (__expf(-k.sigma_t * dist) / d2 * k.n_emitters * kInv4Pi * w);
            qo = light; qd = lx * (-inv); lim = dist * (1.0f - 1e-4f);
        } else {
            const float omc_max = one_minus_cos_max(sm.r * sm.r / d2);
            qd = cone_sample(lx * inv, omc_max, u32_to_unit_f32(b1.x), u32_to_unit_f32(b1.y));
            qo = r.o;
            C = had(mk(sm.lr, sm.lg, sm.lb), r.beta) * (kInv4Pi * kTwoPi * omc_max * k.n_emitters * w);
        }
        float t; int hid;
        const bool hit = scan_sm(S, qo, qd, t, hid);
        if (act) {
            ++scans;
            if (POINT) { if (!hit || t > lim) r.L = r.L + C; }
            else if (hit && hid == src) r.L = r.L + C * __expf(-k.sigma_t * t);
        }
        r.d = phase_sample(u32_to_unit_f32(b1.z), u32_to_unit_f32(b1.w));
        r.beta = r.beta * w;
        r.depth += 1;
        continue_or_end(act, r);
    }

    // ---- SURF_P ----
    __device__ __forceinline__ void surf_p(bool act, Rec &r) {
        const uint32_t ids = act ? r.r1 : 0u;
        const MatF &sm = S.mats[ids & 0xffu];
        const MatF &obj = S.mats[ids >> 8];
        const F3 light = mk(sm.px, sm.py, sm.pz);
        const F3 lx = light - r.o;
        const float d2 = dot(lx, lx), inv = rsqrtf(d2), dist = d2 * inv;
        const F3 n_ = unit(r.o - mk(obj.px, obj.py, obj.pz));
        const F3 wi = lx * inv;
        const bool facet = obj.material == 1;
        F3 f = mk(obj.cr, obj.cg, obj.cb) * kInvPi;
        if (facet) { const Frame fr = make_frame(n_); f = brdf_eval(obj, unit(to_local(fr, wi)), unit(to_local(fr, -r.d))); }
        const F3 C = had(had(mk(sm.lr, sm.lg, sm.lb), f), r.beta) * (dot(n_, wi) * __expf(-k.sigma_t * dist) / d2 * k.n_emitters * k.inv_cp);
        float t; int hid;
        const bool hit = scan_sm(S, light, lx * (-inv), t, hid);
        if (act) {
            ++scans;
            if (!hit || t > dist * (1.0f - 1e-4f)) r.L = r.L + C;
        }
        dest = act ? (obj.material != 0 ? SQ_SURF_F : SQ_SURF_L) : -1;
    }

    // ---- SURF ----
    template <bool FACET>
    __device__ __forceinline__ void surf(bool act, Rec &r) {
        const int id = act ? (int)(r.r1 >> 8) : 0;
        const MatF &obj = S.mats[id];
        const F3 o = r.o, d = r.d, beta = r.beta;
        F3 L = r.L;
        const F3 n_ = unit(o - mk(obj.px, obj.py, obj.pz));
        const Frame fr = make_frame(n_);
        const F3 wo_l = FACET ? unit(to_local(fr, -d)) : mk(0, 0, 1);
        const F3 albedo = mk(obj.cr, obj.cg, obj.cb);
        float omc_last = 1.0f;
        // material 2 (dielectric, as written in the reference: vpt_f32.cuh dielectric_setup) shares this stage with the microfacet: its
        // light-sampled terms are zero (samplingFunctions.h:190), the loop below only runs its scans in step with the other lanes
        const bool diel = FACET && obj.material == 2;
        DielF di; di.F = 0.0f; di.wr = di.wt = mk(0, 0, 1);
        if (diel) di = dielectric_setup(wo_l);
        float gpdf_loop = 0.0f; // the pdf the reference's light loop leaves behind for the dielectric's BSDF term (misSamplingFunctions.h:110-118,148)
        bool refracted = false;
        uint4 ra = make_uint4(0, 0, 0, 0);
        for (int a = 0; a < sc.n_area; ++a) {
            if ((a & 1) == 0) ra = philox_block(r.pixel, r.sample, r.depth, 2 + (a >> 1), lp.key0, lp.key1);
            const float xi1 = u32_to_unit_f32((a & 1) ? ra.z : ra.x), xi2 = u32_to_unit_f32((a & 1) ? ra.w : ra.y);
            const int lid = sc.area[a];
            const MatF &sm = S.mats[lid];
            const F3 cx = mk(sm.px, sm.py, sm.pz) - o;
            const float len2 = dot(cx, cx), inv_len = rsqrtf(len2);
            const float omc_max = one_minus_cos_max(sm.r * sm.r / len2);
            omc_last = omc_max;
            const F3 wi = cone_sample(cx * inv_len, omc_max, xi1, xi2);
            float t; int hid;
            const bool hit = scan_sm(S, o, wi, t, hid);
            if (act) {
                ++scans;
                if ((hit ? hid : 0) == lid && !diel) {
                    const float cos_i = dot(n_, wi);
                    F3 f = albedo * kInvPi;
                    float gpdf = cos_i * kInvPi;
                    if (FACET) { const F3 wi_l = unit(to_local(fr, wi)); const F3 wh = unit(wi_l + wo_l); f = facet_brdf(obj, wi_l, wh, wo_l); gpdf = facet_pdf(wo_l, wh, obj.alpha); }
                    const float inv_fpdf = kTwoPi * omc_max;
                    const float wmis = power_heuristic(1.0f / inv_fpdf, gpdf);
                    L = L + had(had(mk(sm.lr, sm.lg, sm.lb), f), beta) * (cos_i * inv_fpdf * __expf(-k.sigma_t * len2 * inv_len) * wmis * k.inv_cp);
                }
            }
        }
        const uint4 b1 = philox_block(r.pixel, r.sample, r.depth, 1, lp.key0, lp.key1);
        {
            const float xi1 = u32_to_unit_f32(b1.z), xi2 = u32_to_unit_f32(b1.w);
            F3 wi_l, wh = mk(0, 0, 1);
            if (FACET) {
                wh = facet_normal(obj.alpha, xi1, xi2); wi_l = unit(fma3(wh, 2.0f * dot(wh, wo_l), -wo_l));
                if (diel) { // softDielectric (samplingFunctions.h:209-235): reflect with probability F, else the reference's refraction
                    if (sc.n_area > 0) {
                        const uint32_t slot = S_DIEL + (uint32_t)sc.n_area - 1u;
                        const float xg = u32_to_unit_f32(pick_lane(philox_block(r.pixel, r.sample, r.depth, slot >> 2, lp.key0, lp.key1), slot & 3u));
                        gpdf_loop = xg > di.F ? 1.0f - di.F : di.F;
                    }
                    refracted = !(xi1 < di.F);
                    wi_l = refracted ? di.wt : di.wr;
                }
            } else wi_l = cosine_local(xi1, xi2);
            const F3 wi = unit(to_world(fr, wi_l));
            float t; int hid;
            const bool hit = scan_sm(S, o, wi, t, hid);
            if (act) {
                ++scans;
                if (hit && S.mats[hid].emits) {
                    const MatF &em = S.mats[hid];
                    const F3 cx = mk(em.px, em.py, em.pz) - o;
                    float omc = one_minus_cos_max(em.r * em.r / dot(cx, cx));
                    if (diel) L = L + had(dielectric_direct(em, o, wi_l.z, refracted, gpdf_loop), beta) * k.inv_cp;
                    else if (FACET) {
                        const float gpdf = facet_pdf(wo_l, wh, obj.alpha);
                        const F3 g = had(mk(em.lr, em.lg, em.lb), facet_brdf(obj, wi_l, wh, wo_l)) * (wi_l.z / gpdf);
                        if (!(g.x > 0.0f)) omc = omc_last;
                        L = L + had(g, beta) * (power_heuristic(gpdf, 1.0f / (kTwoPi * omc)) * k.inv_cp);
                    } else {
                        const F3 g = had(mk(em.lr, em.lg, em.lb), albedo);
                        if (g.x > 0.0f && g.y > 0.0f && g.z > 0.0f)
                            L = L + had(g, beta) * (power_heuristic(dot(n_, wi) * kInvPi, 1.0f / (kTwoPi * omc)) * k.inv_cp);
                    }
                }
            }
        }
        F3 wi, weight;
        if (FACET) weight = bsdf_sample(obj, fr, wo_l, u32_to_unit_f32(b1.x), u32_to_unit_f32(b1.y), wi);
        else { wi = unit(to_world(fr, cosine_local(u32_to_unit_f32(b1.x), u32_to_unit_f32(b1.y)))); weight = albedo; }
        r.L = L;
        r.beta = had(beta, weight) * k.inv_cp;
        r.d = wi;
        r.depth += 1;
        continue_or_end(act, r);
    }

    // ---- generation: camera sample g of the launch -> owned pixel g % n_owned, sample g / n_owned ----
    __device__ __forceinline__ void generate(bool mine, unsigned long long g, Rec &r) {
        bool alive = false;
        if (mine) {
            const unsigned long long s = g / H.n_owned_pixels;
            const unsigned op = (unsigned)(g - s * H.n_owned_pixels);                 // owned pixel index
            const unsigned tile = op >> 7;
            const unsigned pixel = (tile * (unsigned)lp.tile_count + (unsigned)lp.tile_rank) * (unsigned)kTile + (op & (unsigned)(kTile - 1));
            if (pixel < (unsigned)lp.n_pixels) {
                ++paths;
                r.pixel = pixel; r.sample = (uint32_t)lp.sample_begin + (uint32_t)s; r.depth = 0;
                const uint4 b0 = philox_block(pixel, r.sample, 0u, 0, lp.key0, lp.key1);
                alive = !(k.max_depth > 0 && 0 >= k.max_depth) && !(u32_to_unit_f32(b0.x) < k.q);
                if (alive) {
                    const uint4 j = philox_block(pixel, r.sample, kJitterBounce, 0, lp.key0, lp.key1);
                    const int row = (int)(pixel / (unsigned)lp.width), col = (int)pixel - row * lp.width;
                    const float fx = (float)col, fy = (float)(lp.height - 1 - row);
                    const float u = (fx + u32_to_unit_f32(j.x) - 0.5f) * k.inv_w - 0.5f, v = (fy + u32_to_unit_f32(j.y) - 0.5f) * k.inv_h - 0.5f;
                    r.d = unit(mk(fmaf(k.cam_cx[0], u, fmaf(k.cam_cy[0], v, k.cam_d[0])), fmaf(k.cam_cx[1], u, fmaf(k.cam_cy[1], v, k.cam_d[1])),
                                  fmaf(k.cam_cx[2], u, fmaf(k.cam_cy[2], v, k.cam_d[2]))));
                    r.o = mk(k.cam_o[0], k.cam_o[1], k.cam_o[2]);
                    r.beta = mk(1, 1, 1); r.L = mk(0, 0, 0);
                    r.r1 = b0.y; r.r2 = b0.z; r.r3 = b0.w;
                }
            }
        }
        dest = alive ? SQ_PRIMARY : -1;
    }

    __device__ __forceinline__ void flush_tally() {
        unsigned long long ev = events, scn = scans, nf = nonfinite, np = paths;
        for (int off = 16; off > 0; off >>= 1) {
            ev += __shfl_down_sync(0xffffffffu, ev, off); scn += __shfl_down_sync(0xffffffffu, scn, off);
            nf += __shfl_down_sync(0xffffffffu, nf, off); np += __shfl_down_sync(0xffffffffu, np, off);
        }
        if ((threadIdx.x & 31) == 0) {
            if (ev) atomicAdd(&H.tally[0], ev);
            if (scn) atomicAdd(&H.tally[1], scn);
            if (nf) atomicAdd(&H.tally[2], nf);
            if (np) atomicAdd(&H.tally[3], np);
        }
    }
};

} // namespace f32
} // namespace vpt
