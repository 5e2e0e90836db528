// vpt_wavefront.cuh -- warp-local wavefront FP32 kernel (VPT_KERNEL_WAVEFRONT).
//
// The wavefront decomposition of the radiance loop (extend / medium-sample / NEE-shadow / scatter stages over compacted
// path queues) done ON CHIP: every warp is an independent worker that owns 32 pixels and keeps a pool of kPoolSlots path
// records plus one index queue per stage in shared memory (B200: 227 KB per SM).  A stage pops up to 32 records of the SAME
// kind, so the lanes of the warp run the same code on live data -- the ncu profiles of the two megakernels
// (profiles/r1_summary.md) showed 9-10 of 32 lanes active and 40-47 % instruction-fetch stalls caused by lanes sitting in
// different phases of different vertex kinds.  Queues are compacted with warp ballots / prefix popcounts; nothing goes
// through HBM except the final 12 B per pixel, and no block-level synchronisation is needed after the scene is staged.
//
// Stages (records flow PRIMARY -> {MED | SURF_P -> SURF | SURF} -> PRIMARY until roulette or an emitter ends the path):
//   refill    32 new camera samples at once (lane = pixel of the warp's tile): roulette of bounce 0 for all of them in one
//             go (vptShadeMethods.h:1282), camera rays for the survivors (rt.cpp:787)
//   PRIMARY   scene scan of the path ray, light pick, distance sampling (free-flight / equi-angular), surface-or-medium
//   MED       medium vertex: next-event ray (point-light visibility or cone sample), scan, phase-function scatter, roulette
//   SURF_P    surface vertex whose picked source needs the pLight term (vptShadeMethods.h:62): visibility scan
//   SURF      surface vertex: one cone-sampled scan per area light + the BSDF-sampled scan of MISv2, bdsf scatter, roulette
// Random-number slots, formulas and semantics are exactly those of vpt_f32.cuh / vpt_mega_scan.cuh.
//
// Finished paths are added to per-pixel 64-bit fixed-point accumulators (2^-30 resolution) with shared-memory atomics:
// integer addition is order-independent, so the image stays bit-reproducible although the order in which paths finish is
// data dependent.
#pragma once
#include "vpt_mega_scan.cuh"

namespace vpt {
namespace f32 {

constexpr int kPoolSlots = 128; // path records per warp (power of two, <= 256: queue entries are bytes)
constexpr int kWarpsPerBlock = kThreadsPerBlock / 32;
enum : int { Q_PRIMARY = 0, Q_MED = 1, Q_SURF_P = 2, Q_SURF = 3, Q_COUNT = 4 };
constexpr float kFixScale = 1073741824.0f; // 2^30
constexpr double kFixInv = 1.0 / 1073741824.0;

struct WarpPool {
    float ox[kPoolSlots], oy[kPoolSlots], oz[kPoolSlots];
    float dx[kPoolSlots], dy[kPoolSlots], dz[kPoolSlots];
    float br[kPoolSlots], bg[kPoolSlots], bb[kPoolSlots]; // throughput
    float lr[kPoolSlots], lg[kPoolSlots], lb[kPoolSlots]; // radiance collected so far
    float w[kPoolSlots];                                   // medium vertex: throughput factor applied at the scatter
    uint32_t sample[kPoolSlots];
    uint32_t r1[kPoolSlots], r2[kPoolSlots], r3[kPoolSlots]; // Philox block 0 of the bounce, words y,z,w (light pick, distance, decision)
    uint32_t meta[kPoolSlots];                              // pixel_local | depth << 8
    uint32_t ids[kPoolSlots];                               // picked source | hit object << 8
    unsigned long long acc[32][3];
    uint8_t queue[Q_COUNT][kPoolSlots];
    uint8_t freelist[kPoolSlots];
};

__device__ __noinline__ bool scan_call(const SceneF &sc, float ox, float oy, float oz, float dx, float dy, float dz, float &t, int &id) {
    return scan_fast(sc, mk(ox, oy, oz), mk(dx, dy, dz), t, id);
}

template <int METHOD>
struct Wavefront {
    const SceneF &sc;
    const MatF *mats;
    const Consts &k;
    const CameraF &cam;
    WarpPool &P;
    const uint32_t key0, key1;
    const uint32_t pixel_base; // first pixel of this warp
    const int n_valid;         // pixels of this warp inside the image
    const int width, height;
    const unsigned lane;
    // warp-uniform scheduler state
    int qhead[Q_COUNT], qcount[Q_COUNT], free_count;
    // per-lane statistics
    unsigned events = 0, scans = 0, nonfinite = 0, paths = 0;

    __device__ Wavefront(const SceneF &sc_, const MatF *mats_, const Consts &k_, const CameraF &cam_, WarpPool &P_, uint32_t key0_, uint32_t key1_,
                         uint32_t pixel_base_, int n_valid_, int width_, int height_)
        : sc(sc_), mats(mats_), k(k_), cam(cam_), P(P_), key0(key0_), key1(key1_), pixel_base(pixel_base_), n_valid(n_valid_), width(width_), height(height_),
          lane(threadIdx.x & 31u) {
        for (int q = 0; q < Q_COUNT; ++q) { qhead[q] = 0; qcount[q] = 0; }
        // zero the whole pool: lanes without a record in a partial batch read slot 0, whose indices must always be valid
        uint32_t *raw = reinterpret_cast<uint32_t *>(&P);
        for (int i = lane; i < (int)(sizeof(WarpPool) / 4); i += 32) raw[i] = 0u;
        __syncwarp();
        for (int i = lane; i < kPoolSlots; i += 32) P.freelist[i] = (uint8_t)i;
        free_count = kPoolSlots;
        __syncwarp();
    }

    // ---- queue primitives (warp-synchronous) ----------------------------------------------------------------------------
    __device__ __forceinline__ int pop(int q, int n) {
        const int slot = ((int)lane < n) ? (int)P.queue[q][(qhead[q] + (int)lane) & (kPoolSlots - 1)] : -1;
        qhead[q] = (qhead[q] + n) & (kPoolSlots - 1); qcount[q] -= n;
        return slot;
    }
    __device__ __forceinline__ void push(int q, bool flag, int slot) {
        const unsigned m = __ballot_sync(0xffffffffu, flag);
        if (flag) P.queue[q][(qhead[q] + qcount[q] + __popc(m & ((1u << lane) - 1u))) & (kPoolSlots - 1)] = (uint8_t)slot;
        qcount[q] += __popc(m);
    }
    __device__ __forceinline__ int alloc(bool flag) {
        const unsigned m = __ballot_sync(0xffffffffu, flag);
        const int slot = flag ? (int)P.freelist[free_count - 1 - __popc(m & ((1u << lane) - 1u))] : -1;
        free_count -= __popc(m);
        return slot;
    }
    __device__ __forceinline__ void release(bool flag, int slot) {
        const unsigned m = __ballot_sync(0xffffffffu, flag);
        if (flag) P.freelist[free_count + __popc(m & ((1u << lane) - 1u))] = (uint8_t)slot;
        free_count += __popc(m);
    }
    __device__ __forceinline__ void finish_path(int pixel_local, F3 L) { // rt.cpp:794: pixelValue += f(...)
        if (isfinite(L.x + L.y + L.z)) {
            atomicAdd(&P.acc[pixel_local][0], (unsigned long long)__float2ll_rn(L.x * kFixScale));
            atomicAdd(&P.acc[pixel_local][1], (unsigned long long)__float2ll_rn(L.y * kFixScale));
            atomicAdd(&P.acc[pixel_local][2], (unsigned long long)__float2ll_rn(L.z * kFixScale));
        } else ++nonfinite;
    }
    // roulette for the next bounce; on survival the record (already holding o) gets its new direction, throughput and block-0 words
    __device__ __forceinline__ void continue_or_end(bool act, int slot, int pixel_local, uint32_t sample, int depth, F3 d, F3 beta, F3 L) {
        bool alive = false;
        if (act) {
            const uint4 b0 = philox_block(pixel_base + pixel_local, sample, (uint32_t)depth, 0, key0, key1);
            alive = !(k.max_depth > 0 && depth >= k.max_depth) && !(u32_to_unit_f32(b0.x) < k.q); // vptShadeMethods.h:1282
            if (alive) {
                P.dx[slot] = d.x; P.dy[slot] = d.y; P.dz[slot] = d.z;
                P.br[slot] = beta.x; P.bg[slot] = beta.y; P.bb[slot] = beta.z;
                P.lr[slot] = L.x; P.lg[slot] = L.y; P.lb[slot] = L.z;
                P.r1[slot] = b0.y; P.r2[slot] = b0.z; P.r3[slot] = b0.w;
                P.meta[slot] = (uint32_t)pixel_local | ((uint32_t)depth << 8);
            } else finish_path(pixel_local, L);
        }
        push(Q_PRIMARY, alive, slot);
        release(act && !alive, slot);
    }

    // ---- refill: 32 camera samples, lane = pixel ------------------------------------------------------------------------------
    __device__ __forceinline__ void refill(uint32_t sample) {
        const bool valid = (int)lane < n_valid;
        bool alive = false;
        uint4 b0 = make_uint4(0, 0, 0, 0);
        if (valid) {
            ++paths;
            b0 = philox_block(pixel_base + lane, sample, 0u, 0, key0, key1);
            alive = !(k.max_depth > 0 && 0 >= k.max_depth) && !(u32_to_unit_f32(b0.x) < k.q);
        }
        const int slot = alloc(alive);
        if (alive) {
            const uint4 j = philox_block(pixel_base + lane, sample, kJitterBounce, 0, key0, key1);
            const uint32_t pixel = pixel_base + lane;
            const int row = (int)(pixel / (uint32_t)width), col = (int)(pixel - (uint32_t)row * (uint32_t)width);
            const float fx = (float)col, fy = (float)(height - 1 - row); // rt.cpp:773
            const float u = (fx + u32_to_unit_f32(j.x) - 0.5f) * cam.inv_w - 0.5f, v = (fy + u32_to_unit_f32(j.y) - 0.5f) * cam.inv_h - 0.5f;
            const F3 d = unit(fma3(cam.cx, u, fma3(cam.cy, v, cam.d)));
            P.ox[slot] = cam.o.x; P.oy[slot] = cam.o.y; P.oz[slot] = cam.o.z;
            P.dx[slot] = d.x; P.dy[slot] = d.y; P.dz[slot] = d.z;
            P.br[slot] = 1.0f; P.bg[slot] = 1.0f; P.bb[slot] = 1.0f;
            P.lr[slot] = 0.0f; P.lg[slot] = 0.0f; P.lb[slot] = 0.0f;
            P.sample[slot] = sample;
            P.r1[slot] = b0.y; P.r2[slot] = b0.z; P.r3[slot] = b0.w;
            P.meta[slot] = lane;
        }
        push(Q_PRIMARY, alive, slot);
    }

    // ---- PRIMARY -------------------------------------------------------------------------------------------------------------------
    __device__ __forceinline__ void stage_primary(int n) {
        const int slot = pop(Q_PRIMARY, n);
        const bool act = slot >= 0;
        const int s = act ? slot : 0;
        F3 o = mk(P.ox[s], P.oy[s], P.oz[s]);
        const F3 d = mk(P.dx[s], P.dy[s], P.dz[s]);
        float t; int hid;
        const bool hit = scan_call(sc, o.x, o.y, o.z, d.x, d.y, d.z, t, hid);
        bool to_med = false, to_surf_p = false, to_surf = false, ended = false;
        if (act) {
            ++scans; ++events;
            if (!hit) { t = kMaxFloat; hid = 0; }
            const int pick = min((int)(u32_to_unit_f32(P.r1[s]) * k.n_emitters), sc.n_emitters - 1);
            const int src = sc.emitters[pick];
            const MatF &sm = mats[src];
            bool surface; float dist, inv_pdf = 1.0f;
            if (METHOD == 0) {
                dist = -logf(1.0f - u32_to_unit_f32(P.r2[s])) * k.inv_sigma_t; // freeFlightSample
                surface = dist > t;
            } else if (METHOD == 4) { // distance-sampling MIS (vpt_f32.cuh mis_distance)
                const MatF &ls = sm;
                surface = mis_distance(mk(ls.px, ls.py, ls.pz), o, d, t, expf(-k.sigma_t * t), k.sigma_t, k.inv_sigma_t, u32_to_unit_f32(P.r2[s]), u32_to_unit_f32(P.r3[s]), dist, inv_pdf);
            } else { // equiAngularParams2 + equiAngularProb
                const float Tr = expf(-k.sigma_t * t);
                const F3 dv = mk(sm.px, sm.py, sm.pz) - o;
                const float proj = dot(dv, d);
                const F3 perp = fma3(d, -proj, dv);
                const float D = sqrtf(dot(perp, perp));
                const float thA = atan2f(-proj, D), thB = atan2f(t - proj, D);
                const float xi = u32_to_unit_f32(P.r2[s]);
                const float tl = D * tanf((1.0f - xi) * thA + xi * thB);
                dist = tl + proj;
                inv_pdf = fabsf(thB - thA) * (tl * tl + D * D) / (D * (1.0f - Tr));
                const float xs = u32_to_unit_f32(P.r3[s]);
                surface = (METHOD == 1) ? (xs <= Tr) : (xs < Tr);
            }
            if (surface && mats[hid].emits) { // :1308-1313
                const uint32_t meta = P.meta[s];
                const F3 L = (meta >> 8) == 0 ? had(mk(mats[hid].lr, mats[hid].lg, mats[hid].lb), mk(P.br[s], P.bg[s], P.bb[s])) : mk(P.lr[s], P.lg[s], P.lb[s]);
                finish_path((int)(meta & 0xffu), L);
                ended = true;
            } else if (surface) {
                o = fma3(d, t, o);
                const F3 lx = mk(sm.px, sm.py, sm.pz) - o;
                to_surf_p = !(sm.r > 0.0f && dot(lx, lx) > sm.r * sm.r); // pLight is zero for an area source seen from outside it
                to_surf = !to_surf_p;
                P.ids[s] = (uint32_t)src | ((uint32_t)hid << 8);
            } else {
                o = fma3(d, dist, o);
                P.w[s] = (METHOD == 0) ? k.albedo_over_cp : k.sigma_s * expf(-k.sigma_t * fabsf(dist)) * inv_pdf * k.inv_cp;
                P.ids[s] = (uint32_t)src;
                to_med = true;
            }
            if (!ended) { P.ox[s] = o.x; P.oy[s] = o.y; P.oz[s] = o.z; }
        }
        push(Q_MED, to_med, slot);
        push(Q_SURF_P, to_surf_p, slot);
        push(Q_SURF, to_surf, slot);
        release(ended, slot);
    }

    // ---- MED: (free)SingleScattering + isotropicPhaseSample + roulette ----------------------------------------------------------
    __device__ __forceinline__ void stage_med(int n) {
        const int slot = pop(Q_MED, n);
        const bool act = slot >= 0;
        const int s = act ? slot : 0;
        const F3 o = mk(P.ox[s], P.oy[s], P.oz[s]);
        F3 beta = mk(P.br[s], P.bg[s], P.bb[s]);
        F3 L = mk(P.lr[s], P.lg[s], P.lb[s]);
        const float w = P.w[s];
        const int src = (int)(P.ids[s] & 0xffu);
        const uint32_t sample = P.sample[s], meta = P.meta[s];
        const int pixel_local = (int)(meta & 0xffu), depth = (int)(meta >> 8);
        const uint4 b1 = philox_block(pixel_base + pixel_local, sample, (uint32_t)depth, 1, key0, key1);
        const MatF &sm = mats[src];
        const F3 light = mk(sm.px, sm.py, sm.pz);
        const F3 lx = light - o;
        const float d2 = dot(lx, lx), inv = rsqrtf(d2);
        const bool point = sm.r == 0.0f;
        F3 qo, qd, C; float lim = 0.0f;
        if (point) {
            const float dist = d2 * inv;
            C = had(mk(sm.lr, sm.lg, sm.lb), beta) * (expf(-k.sigma_t * dist) / d2 * k.n_emitters * kInv4Pi * w);
            qo = light; qd = lx * (-inv); lim = dist * (1.0f - 1e-4f);
        } else {
            const float omc_max = one_minus_cos_max(sm.r * sm.r / d2);
            qd = cone_sample(lx * inv, omc_max, u32_to_unit_f32(b1.x), u32_to_unit_f32(b1.y));
            qo = o;
            C = had(mk(sm.lr, sm.lg, sm.lb), beta) * (kInv4Pi * kTwoPi * omc_max * k.n_emitters * w);
        }
        float t; int hid;
        const bool hit = scan_call(sc, qo.x, qo.y, qo.z, qd.x, qd.y, qd.z, t, hid);
        if (act) {
            ++scans;
            if (point) { if (!hit || t > lim) L = L + C; }
            else if (hit && hid == src) L = L + C * expf(-k.sigma_t * t);
        }
        const F3 d = phase_sample(u32_to_unit_f32(b1.z), u32_to_unit_f32(b1.w));
        beta = beta * w;
        continue_or_end(act, slot, pixel_local, sample, depth + 1, d, beta, L);
    }

    // ---- SURF_P: pLight ---------------------------------------------------------------------------------------------------------------
    __device__ __forceinline__ void stage_surf_p(int n) {
        const int slot = pop(Q_SURF_P, n);
        const bool act = slot >= 0;
        const int s = act ? slot : 0;
        const F3 o = mk(P.ox[s], P.oy[s], P.oz[s]);
        const F3 d = mk(P.dx[s], P.dy[s], P.dz[s]);
        const F3 beta = mk(P.br[s], P.bg[s], P.bb[s]);
        const uint32_t ids = P.ids[s];
        const MatF &sm = mats[ids & 0xffu];
        const MatF &obj = mats[ids >> 8];
        const F3 light = mk(sm.px, sm.py, sm.pz);
        const F3 lx = light - o;
        const float d2 = dot(lx, lx), inv = rsqrtf(d2), dist = d2 * inv;
        const F3 n_ = unit(o - mk(obj.px, obj.py, obj.pz));
        const F3 wi = lx * inv;
        F3 f = mk(obj.cr, obj.cg, obj.cb) * kInvPi;
        if (obj.material == 1) { const Frame fr = make_frame(n_); f = brdf_eval(obj, unit(to_local(fr, wi)), unit(to_local(fr, -d))); }
        const F3 C = had(had(mk(sm.lr, sm.lg, sm.lb), f), beta) * (dot(n_, wi) * expf(-k.sigma_t * dist) / d2 * k.n_emitters * k.inv_cp);
        const F3 qd = lx * (-inv);
        float t; int hid;
        const bool hit = scan_call(sc, light.x, light.y, light.z, qd.x, qd.y, qd.z, t, hid);
        if (act) {
            ++scans;
            if (!hit || t > dist * (1.0f - 1e-4f)) { P.lr[s] += C.x; P.lg[s] += C.y; P.lb[s] += C.z; }
        }
        push(Q_SURF, act, slot);
    }

    // ---- SURF: MISv2 + bdsf + roulette --------------------------------------------------------------------------------------------------
    __device__ __forceinline__ void stage_surf(int n) {
        const int slot = pop(Q_SURF, n);
        const bool act = slot >= 0;
        const int s = act ? slot : 0;
        const F3 o = mk(P.ox[s], P.oy[s], P.oz[s]);
        const F3 d = mk(P.dx[s], P.dy[s], P.dz[s]);
        F3 beta = mk(P.br[s], P.bg[s], P.bb[s]);
        F3 L = mk(P.lr[s], P.lg[s], P.lb[s]);
        const int id = (int)(P.ids[s] >> 8);
        const uint32_t sample = P.sample[s], meta = P.meta[s];
        const int pixel_local = (int)(meta & 0xffu), depth = (int)(meta >> 8);
        const uint32_t pixel = pixel_base + pixel_local;
        const MatF &obj = mats[id];
        const F3 n_ = unit(o - mk(obj.px, obj.py, obj.pz));
        const Frame fr = make_frame(n_);
        const F3 wo_l = unit(to_local(fr, -d));
        const bool facet = obj.material == 1;
        const F3 albedo = mk(obj.cr, obj.cg, obj.cb);
        float omc_last = 1.0f;
        uint4 ra = make_uint4(0, 0, 0, 0);
        for (int a = 0; a < sc.n_area; ++a) { // muestreoSA for every area light (misSamplingFunctions.h:105-118)
            if ((a & 1) == 0) ra = philox_block(pixel, sample, (uint32_t)depth, 2 + (a >> 1), key0, key1);
            const float xi1 = u32_to_unit_f32((a & 1) ? ra.z : ra.x), xi2 = u32_to_unit_f32((a & 1) ? ra.w : ra.y);
            const int lid = sc.area[a];
            const MatF &sm = mats[lid];
            const F3 cx = mk(sm.px, sm.py, sm.pz) - o;
            const float len2 = dot(cx, cx), inv_len = rsqrtf(len2);
            const float omc_max = one_minus_cos_max(sm.r * sm.r / len2);
            omc_last = omc_max;
            const F3 wi = cone_sample(cx * inv_len, omc_max, xi1, xi2);
            float t; int hid;
            const bool hit = scan_call(sc, o.x, o.y, o.z, wi.x, wi.y, wi.z, t, hid);
            if (act) {
                ++scans;
                if ((hit ? hid : 0) == lid) { // id stays 0 on a miss, samplingFunctions.h:196
                    const float cos_i = dot(n_, wi);
                    F3 f = albedo * kInvPi;
                    float gpdf = cos_i * kInvPi;
                    if (facet) { const F3 wi_l = unit(to_local(fr, wi)); const F3 wh = unit(wi_l + wo_l); f = facet_brdf(obj, wi_l, wh, wo_l); gpdf = facet_pdf(wo_l, wh, obj.alpha); }
                    const float inv_fpdf = kTwoPi * omc_max;
                    const float wmis = power_heuristic(1.0f / inv_fpdf, gpdf);
                    L = L + had(had(mk(sm.lr, sm.lg, sm.lb), f), beta) * (cos_i * inv_fpdf * expf(-k.sigma_t * len2 * inv_len) * wmis * k.inv_cp);
                }
            }
        }
        const uint4 b1 = philox_block(pixel, sample, (uint32_t)depth, 1, key0, key1);
        { // the BSDF-sampled term of MISv2 (:124-167): slots S_MIS = lanes 2,3 of block 1
            const float xi1 = u32_to_unit_f32(b1.z), xi2 = u32_to_unit_f32(b1.w);
            F3 wi_l, wh = mk(0, 0, 1);
            if (facet) { wh = facet_normal(obj.alpha, xi1, xi2); wi_l = unit(fma3(wh, 2.0f * dot(wh, wo_l), -wo_l)); }
            else wi_l = cosine_local(xi1, xi2);
            const F3 wi = unit(to_world(fr, wi_l));
            float t; int hid;
            const bool hit = scan_call(sc, o.x, o.y, o.z, wi.x, wi.y, wi.z, t, hid);
            if (act) {
                ++scans;
                if (hit && mats[hid].emits) {
                    const MatF &em = mats[hid];
                    const F3 cx = mk(em.px, em.py, em.pz) - o;
                    float omc = one_minus_cos_max(em.r * em.r / dot(cx, cx));
                    if (facet) {
                        const float gpdf = facet_pdf(wo_l, wh, obj.alpha);
                        const F3 g = had(mk(em.lr, em.lg, em.lb), facet_brdf(obj, wi_l, wh, wo_l)) * (wi_l.z / gpdf);
                        if (!(g.x > 0.0f)) omc = omc_last; // the reference's stale costhetaMax (:162)
                        L = L + had(g, beta) * (power_heuristic(gpdf, 1.0f / (kTwoPi * omc)) * k.inv_cp);
                    } else {
                        const F3 g = had(mk(em.lr, em.lg, em.lb), albedo);
                        if (g.x > 0.0f && g.y > 0.0f && g.z > 0.0f)
                            L = L + had(g, beta) * (power_heuristic(dot(n_, wi) * kInvPi, 1.0f / (kTwoPi * omc)) * k.inv_cp);
                    }
                }
            }
        }
        F3 wi; // bdsf (:16-59): slots S_BSDF = lanes 0,1 of block 1
        const F3 weight = bsdf_sample(obj, fr, wo_l, u32_to_unit_f32(b1.x), u32_to_unit_f32(b1.y), wi);
        beta = had(beta, weight) * k.inv_cp;
        continue_or_end(act, slot, pixel_local, sample, depth + 1, wi, beta, L);
    }

    // ---- scheduler --------------------------------------------------------------------------------------------------------------------
    __device__ __forceinline__ void run(int sample_begin, int sample_end) {
        int s_next = sample_begin;
        for (;;) {
            __syncwarp();
            if (s_next < sample_end && free_count >= 32) { refill((uint32_t)s_next++); continue; }
            int best = 0, best_count = qcount[0]; // constant indices only: the queue counters must stay in registers
#pragma unroll
            for (int q = 1; q < Q_COUNT; ++q) if (qcount[q] > best_count) { best = q; best_count = qcount[q]; }
            const int n = min(best_count, 32);
            if (n == 0) break;
            switch (best) {
            case Q_PRIMARY: stage_primary(n); break;
            case Q_MED: stage_med(n); break;
            case Q_SURF_P: stage_surf_p(n); break;
            default: stage_surf(n); break;
            }
        }
    }
};

} // namespace f32
} // namespace vpt
