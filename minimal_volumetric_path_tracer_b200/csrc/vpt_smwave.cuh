// vpt_smwave.cuh -- SM-wide wavefront FP32 kernel (VPT_KERNEL_WAVEFRONT_SM): one persistent CTA per SM, lock-step stages.
//
// Why (profiles/r1_summary.md, "wave v2"): the warp-local wavefront (vpt_wavefront.cuh) fixed SIMT efficiency (28 of 32 lanes
// active) but only 40 % of the issue slots were used -- 42 % of all stall samples were `no_inst`: sixteen warps per SM, each in a
// different stage of a 62 KB kernel, thrash the 32 KB L1.5 / 6 KB L0 instruction caches.  Here ALL warps of an SM run the SAME stage
// at the same time on one shared pool of path records:
//   * one CTA of kSmThreads threads per SM, grid = number of SMs (persistent); work items = groups of pixel tiles, handed out
//     statically (item j -> CTA j % gridDim.x), two items in flight per CTA so that a draining item overlaps the next one;
//   * kSmPool path records in shared memory (SoA, 68 B each), one index queue (ring) per stage;
//   * work proceeds in ROUNDS: at a barrier thread 0 snapshots every queue's tail and decides how many new camera samples to
//     generate into the free records; then every warp claims 32-record batches from the snapshot with one shared-memory atomic per
//     batch -- longest stages first, so the round ends evenly -- runs them and pushes the survivors to the next stage's queue
//     (warp ballot + one atomic per warp); what is pushed during a round is consumed in the next one.  Batches are full warps
//     except while a work item drains, so the lanes stay busy, and one round costs two barriers for about 2000 records;
//   * queues are split by what diverges: medium vertex with a point / an area source, surface vertex needing the pLight shadow ray,
//     Lambert / microfacet surface vertex;
//   * the scene's scan records are staged in shared memory as float4 (broadcast LDS.128); ordinary spheres (r < 64) use the
//     direct roots -b -+ sqrt(det) with det = r^2 - |op - (op.d)d|^2 (21 instructions per test), huge ones the re-anchored
//     cancellation-free form of vpt_f32.cuh.
// Random-number slots, formulas and semantics are exactly those of vpt_f32.cuh (file:line citations there); per-pixel sums use
// the order-independent 2^-30 fixed-point accumulators of vpt_wavefront.cuh, so images are bit-reproducible.
#pragma once
#include "vpt_mega_scan.cuh"

namespace vpt {
namespace f32 {

constexpr int kSmThreads = 768;
constexpr int kSmPool = 2048;         // path records per CTA (power of two)
constexpr int kSmMaxItemPixels = 256; // pixels per work item (power of two multiple of kTile)
constexpr float kSmFixScale = 1073741824.0f; // 2^30
constexpr double kSmFixInv = 1.0 / 1073741824.0;
constexpr float kSimpleRootMaxR2 = 64.0f * 64.0f;

enum : int { SQ_PRIMARY = 0, SQ_MED_POINT, SQ_MED_AREA, SQ_SURF_P, SQ_SURF_L, SQ_SURF_F, SQ_COUNT, ST_GEN = SQ_COUNT, ST_FLUSH, ST_EXIT };

struct SmShared {
    // ---- path records (SoA) ----
    float ox[kSmPool], oy[kSmPool], oz[kSmPool];
    float dx[kSmPool], dy[kSmPool], dz[kSmPool];
    float br[kSmPool], bg[kSmPool], bb[kSmPool]; // throughput
    float lr[kSmPool], lg[kSmPool], lb[kSmPool]; // radiance collected so far
    uint32_t sample[kSmPool];
    uint32_t meta[kSmPool];                      // pixel-in-item (bits 0-8) | item slot (bit 9) | depth << 10
    uint32_t r1[kSmPool];                        // Philox block 0 word y (light pick); after PRIMARY: picked source | hit object << 8
    uint32_t r2[kSmPool];                        // word z (distance);            after PRIMARY (medium vertex): throughput factor w
    uint32_t r3[kSmPool];                        // word w (surface/medium decision)
    uint16_t queue[SQ_COUNT][kSmPool];
    uint16_t freelist[kSmPool];
    unsigned long long acc[2][kSmMaxItemPixels][3];
    // ---- scene ----
    MatF mats[kMaxSpheres];
    float4 ga[2 * kMaxSpheres]; // general-form spheres: (qx qy qz c0) (mx my mz -)
    float4 gb[kMaxSpheres];     // direct-root spheres: (px py pz r^2)
    int gid[2 * kMaxSpheres];   // scan order -> caller's sphere index (general ones first)
    int n_ga, n_gb;
    // ---- control ----
    unsigned q_head[SQ_COUNT]; // claim counters: a warp takes entries [old, old + 32) with one atomicAdd
    unsigned q_tail[SQ_COUNT]; // push counters
    unsigned q_end[SQ_COUNT];  // this round's snapshot: entries below it may be claimed
    unsigned free_head, free_tail; // the free-record ring: allocate at the head (only below this round's snapshot), release at the tail
    unsigned gen_claim, gen_count, gen_begin; // this round's new camera samples: claim counter, how many, first sample index of the item
    int gen_slot, flush_slot, exit_flag;
    unsigned t_cursor[2], t_done[2];
    int t_item[2]; // -1: slot idle
    int next_item;
};

extern __shared__ __align__(16) unsigned char smwave_smem[]; // the CTA's one SmShared (dynamic shared memory)
__device__ __forceinline__ SmShared &sm_shared() { return *reinterpret_cast<SmShared *>(smwave_smem); }

// nearest accepted hit over all spheres (pathTracingUtilities.h:12-36 with Sphere.h:27-37): distance (+inf: none) and scan index.
// One out-of-line copy: it is called from seven places and must stay resident in the instruction cache.
struct ScanHit { float t; int index; };
static __device__ __noinline__ ScanHit scan_sm_call(float ox, float oy, float oz, float dx, float dy, float dz) {
    const SmShared &S = sm_shared();
    const F3 o = mk(ox, oy, oz), d = mk(dx, dy, dz);
    float best = CUDART_INF_F;
    int bi = -1;
    const int na = S.n_ga, nb = S.n_gb;
    for (int i = 0; i < na; ++i) {
        const float4 a = S.ga[2 * i], m = S.ga[2 * i + 1];
        const F3 oq = mk(o.x - a.x, o.y - a.y, o.z - a.z);
        const F3 op = mk(oq.x + m.x, oq.y + m.y, oq.z + m.z);
        const float b = dot(op, d);
        const float c = fmaf(oq.x, op.x + m.x, fmaf(oq.y, op.y + m.y, fmaf(oq.z, op.z + m.z, a.w))); // |op|^2 - r^2 without cancellation
        const float det = fmaf(b, b, -c);
        const float sq = det * rsqrtf(det); // NaN when det <= 0: every comparison below is then false
        const float q = -(b + copysignf(sq, b));
        const float other = __fdividef(c, q);
        const float tn = fminf(q, other), tf = fmaxf(q, other);
        const float ti = tn >= kEps ? tn : tf; // Sphere.h:34
        if (ti > kEps && ti < best) { best = ti; bi = i; }
    }
    for (int i = 0; i < nb; ++i) {
        const float4 a = S.gb[i];
        const F3 oq = mk(o.x - a.x, o.y - a.y, o.z - a.z);
        const float b = dot(oq, d);
        const F3 l = fma3(d, -b, oq);
        const float det = fmaf(-l.x, l.x, fmaf(-l.y, l.y, fmaf(-l.z, l.z, a.w)));
        const float sq = det * rsqrtf(det);
        const float tn = -b - sq, tf = sq - b;
        const float ti = tn >= kEps ? tn : tf;
        if (ti > kEps && ti < best) { best = ti; bi = na + i; }
    }
    return ScanHit{best, bi};
}
__device__ __forceinline__ bool scan_sm(const SmShared &S, F3 o, F3 d, float &t, int &id) {
    const ScanHit h = scan_sm_call(o.x, o.y, o.z, d.x, d.y, d.z);
    t = h.t;
    id = h.index >= 0 ? S.gid[h.index] : -1;
    return h.index >= 0;
}

template <int METHOD>
struct SmWave {
    SmShared &S;
    const SceneF &sc;
    const ConstsF &k;
    const LaunchParams &lp;
    const int tid, lane;
    const int log_p, item_pixels, n_owned_tiles;
    unsigned events = 0, scans = 0, nonfinite = 0, paths = 0;

    __device__ SmWave(SmShared &S_, const SceneF &sc_, const ConstsF &k_, const LaunchParams &lp_, int log_p_, int n_owned_)
        : S(S_), sc(sc_), k(k_), lp(lp_), tid((int)threadIdx.x), lane((int)threadIdx.x & 31), log_p(log_p_), item_pixels(1 << log_p_), n_owned_tiles(n_owned_) {}

    // ---- work items: item j = owned tiles [j * K, (j + 1) * K), K = item_pixels / kTile -----------------------------------------
    __device__ __forceinline__ long long item_pixel(int item, int pl) const {
        const long long owned = ((long long)item << (log_p - 7)) + (pl >> 7);
        if (owned >= n_owned_tiles) return -1;
        const long long pixel = (owned * lp.tile_count + lp.tile_rank) * kTile + (pl & (kTile - 1));
        return pixel < lp.n_pixels ? pixel : -1;
    }

    // ---- queue / pool primitives (warp-aggregated shared-memory atomics) ------------------------------------------------------------
    __device__ __forceinline__ void push(int q, bool flag, int slot) {
        const unsigned m = __ballot_sync(0xffffffffu, flag);
        if (m == 0u) return;
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(&S.q_tail[q], (unsigned)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (flag) S.queue[q][(base + __popc(m & ((1u << lane) - 1u))) & (kSmPool - 1)] = (uint16_t)slot;
    }
    __device__ __forceinline__ int alloc(bool flag) { // the round's snapshot guarantees enough free records below free_tail
        const unsigned m = __ballot_sync(0xffffffffu, flag);
        if (m == 0u) return -1;
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(&S.free_head, (unsigned)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        return flag ? (int)S.freelist[(base + __popc(m & ((1u << lane) - 1u))) & (kSmPool - 1)] : -1;
    }
    __device__ __forceinline__ void release(bool flag, int slot) {
        const unsigned m = __ballot_sync(0xffffffffu, flag);
        if (m == 0u) return;
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(&S.free_tail, (unsigned)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (flag) S.freelist[(base + __popc(m & ((1u << lane) - 1u))) & (kSmPool - 1)] = (uint16_t)slot;
    }
    // a path ended: add its radiance to the pixel (rt.cpp:794) -- the caller reports it to count_done()
    __device__ __forceinline__ void add_radiance(uint32_t meta, F3 L) {
        if (!isfinite(L.x + L.y + L.z)) { ++nonfinite; return; }
        unsigned long long *a = S.acc[(meta >> 9) & 1u][meta & 0x1ffu];
        if (L.x != 0.0f) atomicAdd(a + 0, (unsigned long long)__float2ll_rn(L.x * kSmFixScale));
        if (L.y != 0.0f) atomicAdd(a + 1, (unsigned long long)__float2ll_rn(L.y * kSmFixScale));
        if (L.z != 0.0f) atomicAdd(a + 2, (unsigned long long)__float2ll_rn(L.z * kSmFixScale));
    }
    __device__ __forceinline__ void count_done(bool ended, uint32_t meta) {
        const unsigned m0 = __ballot_sync(0xffffffffu, ended && ((meta >> 9) & 1u) == 0u);
        const unsigned m1 = __ballot_sync(0xffffffffu, ended && ((meta >> 9) & 1u) == 1u);
        if (lane == 0) {
            if (m0) atomicAdd(&S.t_done[0], (unsigned)__popc(m0));
            if (m1) atomicAdd(&S.t_done[1], (unsigned)__popc(m1));
        }
    }
    __device__ __forceinline__ uint32_t pixel_of(uint32_t meta) const {
        return (uint32_t)item_pixel(S.t_item[(meta >> 9) & 1u], (int)(meta & 0x1ffu));
    }
    // roulette for the next bounce (vptShadeMethods.h:1282); a surviving record (already holding o) gets its new direction,
    // throughput, radiance-so-far and the block-0 words of the new bounce
    __device__ __forceinline__ void continue_or_end(bool act, int slot, uint32_t pixel, uint32_t sample, uint32_t meta, F3 d, F3 beta, F3 L) {
        bool alive = false;
        if (act) {
            const uint32_t depth = meta >> 10;
            const uint4 b0 = philox_block(pixel, sample, depth, 0, lp.key0, lp.key1);
            alive = !(k.max_depth > 0 && (int)depth >= k.max_depth) && !(u32_to_unit_f32(b0.x) < k.q);
            if (alive) {
                S.dx[slot] = d.x; S.dy[slot] = d.y; S.dz[slot] = d.z;
                S.br[slot] = beta.x; S.bg[slot] = beta.y; S.bb[slot] = beta.z;
                S.lr[slot] = L.x; S.lg[slot] = L.y; S.lb[slot] = L.z;
                S.r1[slot] = b0.y; S.r2[slot] = b0.z; S.r3[slot] = b0.w;
                S.meta[slot] = meta;
            } else add_radiance(meta, L);
        }
        push(SQ_PRIMARY, alive, slot);
        release(act && !alive, slot);
        count_done(act && !alive, meta);
    }

    // ---- GEN: n <= 32 new camera samples of item slot b, lane i takes sample index g0 + i -------------------------------------------
    __device__ __forceinline__ void stage_gen(int b, unsigned g0, int n) {
        const bool mine = lane < n;
        const unsigned g = g0 + (unsigned)lane;
        const int pl = (int)(g & (unsigned)(item_pixels - 1));
        const uint32_t sample = (uint32_t)lp.sample_begin + (g >> log_p);
        const long long pixel = mine ? item_pixel(S.t_item[b], pl) : -1;
        bool alive = false;
        uint4 b0 = make_uint4(0, 0, 0, 0);
        if (pixel >= 0) {
            ++paths;
            b0 = philox_block((uint32_t)pixel, sample, 0u, 0, lp.key0, lp.key1);
            alive = !(k.max_depth > 0 && 0 >= k.max_depth) && !(u32_to_unit_f32(b0.x) < k.q); // the roulette of bounce 0, vptShadeMethods.h:1282
        }
        const int slot = alloc(alive);
        if (alive) {
            const uint4 j = philox_block((uint32_t)pixel, sample, kJitterBounce, 0, lp.key0, lp.key1);
            const int row = (int)(pixel / lp.width), col = (int)(pixel - (long long)row * lp.width);
            const float fx = (float)col, fy = (float)(lp.height - 1 - row); // rt.cpp:773
            const float u = (fx + u32_to_unit_f32(j.x) - 0.5f) * k.inv_w - 0.5f, v = (fy + u32_to_unit_f32(j.y) - 0.5f) * k.inv_h - 0.5f;
            const F3 d = unit(mk(fmaf(k.cam_cx[0], u, fmaf(k.cam_cy[0], v, k.cam_d[0])), fmaf(k.cam_cx[1], u, fmaf(k.cam_cy[1], v, k.cam_d[1])),
                                 fmaf(k.cam_cx[2], u, fmaf(k.cam_cy[2], v, k.cam_d[2])))); // rt.cpp:787
            S.ox[slot] = k.cam_o[0]; S.oy[slot] = k.cam_o[1]; S.oz[slot] = k.cam_o[2];
            S.dx[slot] = d.x; S.dy[slot] = d.y; S.dz[slot] = d.z;
            S.br[slot] = 1.0f; S.bg[slot] = 1.0f; S.bb[slot] = 1.0f;
            S.lr[slot] = 0.0f; S.lg[slot] = 0.0f; S.lb[slot] = 0.0f;
            S.sample[slot] = sample;
            S.r1[slot] = b0.y; S.r2[slot] = b0.z; S.r3[slot] = b0.w;
            S.meta[slot] = (uint32_t)pl | ((uint32_t)b << 9);
        }
        push(SQ_PRIMARY, alive, slot);
        // samples of pixels outside the image and paths killed by the first roulette are finished already
        const unsigned m = __ballot_sync(0xffffffffu, mine && !alive);
        if (lane == 0 && m) atomicAdd(&S.t_done[b], (unsigned)__popc(m));
    }

    // ---- PRIMARY: scan of the path ray, light pick, distance sampling, surface-or-medium decision --------------------------------
    __device__ __forceinline__ void stage_primary(int slot) {
        const bool act = slot >= 0;
        const int s = act ? slot : 0;
        F3 o = mk(S.ox[s], S.oy[s], S.oz[s]);
        const F3 d = mk(S.dx[s], S.dy[s], S.dz[s]);
        float t; int hid;
        const bool hit = scan_sm(S, o, d, t, hid);
        bool to_mp = false, to_ma = false, to_sp = false, to_sl = false, to_sf = false, ended = false;
        const uint32_t meta = S.meta[s];
        if (act) {
            ++scans; ++events;
            if (!hit) { t = kMaxFloat; hid = 0; }
            const int pick = min((int)(u32_to_unit_f32(S.r1[s]) * k.n_emitters), sc.n_emitters - 1);
            const int src = sc.emitters[pick];
            const MatF &sm = S.mats[src];
            bool surface; float dist, inv_pdf = 1.0f;
            if (METHOD == 0) {
                dist = -logf(1.0f - u32_to_unit_f32(S.r2[s])) * k.inv_sigma_t; // freeFlightSample, vptSamplingFunctions.h:11
                surface = dist > t;
            } else { // equiAngularParams2 (volumetricBasicFunctions.h:209-223) + equiAngularProb (vptSamplingFunctions.h:60)
                const float Tr = expf(-k.sigma_t * t);
                const F3 dv = mk(sm.px, sm.py, sm.pz) - o;
                const float proj = dot(dv, d);
                const F3 perp = fma3(d, -proj, dv);
                const float D = sqrtf(dot(perp, perp));
                const float thA = atan2f(-proj, D), thB = atan2f(t - proj, D);
                const float xi = u32_to_unit_f32(S.r2[s]);
                const float tl = D * tanf((1.0f - xi) * thA + xi * thB);
                dist = tl + proj;
                inv_pdf = fabsf(thB - thA) * (tl * tl + D * D) / (D * (1.0f - Tr));
                const float xs = u32_to_unit_f32(S.r3[s]);
                surface = (METHOD == 1) ? (xs <= Tr) : (xs < Tr);
            }
            const MatF &obj = S.mats[hid];
            if (surface && obj.emits) { // :1308-1313: a directly seen emitter ends the path
                const F3 L = (meta >> 10) == 0 ? had(mk(obj.lr, obj.lg, obj.lb), mk(S.br[s], S.bg[s], S.bb[s])) : mk(S.lr[s], S.lg[s], S.lb[s]);
                add_radiance(meta, L);
                ended = true;
            } else if (surface) {
                o = fma3(d, t, o);
                const F3 lx = mk(sm.px, sm.py, sm.pz) - o;
                to_sp = !(sm.r > 0.0f && dot(lx, lx) > sm.r * sm.r); // pLight is zero for an area source seen from outside it
                to_sf = !to_sp && obj.material == 1;
                to_sl = !to_sp && !to_sf;
                S.r1[s] = (uint32_t)src | ((uint32_t)hid << 8);
            } else {
                o = fma3(d, dist, o);
                const float w = (METHOD == 0) ? k.albedo_over_cp : k.sigma_s * expf(-k.sigma_t * fabsf(dist)) * inv_pdf * k.inv_cp;
                S.r2[s] = __float_as_uint(w);
                S.r1[s] = (uint32_t)src;
                to_mp = sm.r == 0.0f; to_ma = !to_mp;
            }
            if (!ended) { S.ox[s] = o.x; S.oy[s] = o.y; S.oz[s] = o.z; }
        }
        push(SQ_MED_POINT, to_mp, slot);
        push(SQ_MED_AREA, to_ma, slot);
        push(SQ_SURF_P, to_sp, slot);
        push(SQ_SURF_L, to_sl, slot);
        push(SQ_SURF_F, to_sf, slot);
        release(ended, slot);
        count_done(ended, meta);
    }

    // ---- MED: (free)SingleScattering (volumetricBasicFunctions.h:284-340 / :225-281) + isotropicPhaseSample + roulette ----------
    template <bool POINT>
    __device__ __forceinline__ void stage_med(int slot) {
        const bool act = slot >= 0;
        const int s = act ? slot : 0;
        const F3 o = mk(S.ox[s], S.oy[s], S.oz[s]);
        F3 beta = mk(S.br[s], S.bg[s], S.bb[s]);
        F3 L = mk(S.lr[s], S.lg[s], S.lb[s]);
        const float w = __uint_as_float(S.r2[s]);
        const int src = act ? (int)(S.r1[s] & 0xffu) : 0; // idle lanes run the scan on whatever slot 0 holds: keep their indices in range
        const uint32_t sample = S.sample[s], meta = S.meta[s];
        const uint32_t pixel = pixel_of(meta);
        const uint4 b1 = philox_block(pixel, sample, meta >> 10, 1, lp.key0, lp.key1);
        const MatF &sm = S.mats[src];
        const F3 light = mk(sm.px, sm.py, sm.pz);
        const F3 lx = light - o;
        const float d2 = dot(lx, lx), inv = rsqrtf(d2);
        F3 qo, qd, C; float lim = 0.0f;
        if (POINT) {
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
        const bool hit = scan_sm(S, qo, qd, t, hid);
        if (act) {
            ++scans;
            if (POINT) { if (!hit || t > lim) L = L + C; }
            else if (hit && hid == src) L = L + C * expf(-k.sigma_t * t);
        }
        const F3 d = phase_sample(u32_to_unit_f32(b1.z), u32_to_unit_f32(b1.w));
        beta = beta * w;
        continue_or_end(act, slot, pixel, sample, meta + (1u << 10), d, beta, L);
    }

    // ---- SURF_P: pLight (vptShadeMethods.h:62-91) ------------------------------------------------------------------------------------
    __device__ __forceinline__ void stage_surf_p(int slot) {
        const bool act = slot >= 0;
        const int s = act ? slot : 0;
        const F3 o = mk(S.ox[s], S.oy[s], S.oz[s]);
        const F3 d = mk(S.dx[s], S.dy[s], S.dz[s]);
        const F3 beta = mk(S.br[s], S.bg[s], S.bb[s]);
        const uint32_t ids = act ? S.r1[s] : 0u;
        const MatF &sm = S.mats[ids & 0xffu];
        const MatF &obj = S.mats[ids >> 8];
        const F3 light = mk(sm.px, sm.py, sm.pz);
        const F3 lx = light - o;
        const float d2 = dot(lx, lx), inv = rsqrtf(d2), dist = d2 * inv;
        const F3 n_ = unit(o - mk(obj.px, obj.py, obj.pz));
        const F3 wi = lx * inv;
        const bool facet = obj.material == 1;
        F3 f = mk(obj.cr, obj.cg, obj.cb) * kInvPi;
        if (facet) f = facet_eval_world(obj, n_, wi, d);
        const F3 C = had(had(mk(sm.lr, sm.lg, sm.lb), f), beta) * (dot(n_, wi) * expf(-k.sigma_t * dist) / d2 * k.n_emitters * k.inv_cp);
        const F3 qd = lx * (-inv);
        float t; int hid;
        const bool hit = scan_sm(S, light, qd, t, hid);
        if (act) {
            ++scans;
            if (!hit || t > dist * (1.0f - 1e-4f)) { S.lr[s] += C.x; S.lg[s] += C.y; S.lb[s] += C.z; }
        }
        push(SQ_SURF_L, act && !facet, slot);
        push(SQ_SURF_F, act && facet, slot);
    }
    // microfacet BRDF for world-space directions (rare: kept out of line so that the Lambert stages stay small)
    static __device__ __noinline__ F3 facet_eval_world(const MatF &obj, F3 n_, F3 wi, F3 d) {
        const Frame fr = make_frame(n_);
        return brdf_eval(obj, unit(to_local(fr, wi)), unit(to_local(fr, -d)));
    }

    // ---- SURF: MISv2 (misSamplingFunctions.h:96-170) + bdsf (vptShadeMethods.h:16-59) + roulette -------------------------------------
    template <bool FACET>
    __device__ __forceinline__ void stage_surf(int slot) {
        const bool act = slot >= 0;
        const int s = act ? slot : 0;
        const F3 o = mk(S.ox[s], S.oy[s], S.oz[s]);
        const F3 d = mk(S.dx[s], S.dy[s], S.dz[s]);
        F3 beta = mk(S.br[s], S.bg[s], S.bb[s]);
        F3 L = mk(S.lr[s], S.lg[s], S.lb[s]);
        const int id = act ? (int)(S.r1[s] >> 8) : 0;
        const uint32_t sample = S.sample[s], meta = S.meta[s];
        const uint32_t depth = meta >> 10;
        const uint32_t pixel = pixel_of(meta);
        const MatF &obj = S.mats[id];
        const F3 n_ = unit(o - mk(obj.px, obj.py, obj.pz));
        const Frame fr = make_frame(n_);
        const F3 wo_l = FACET ? unit(to_local(fr, -d)) : mk(0, 0, 1);
        const F3 albedo = mk(obj.cr, obj.cg, obj.cb);
        float omc_last = 1.0f;
        uint4 ra = make_uint4(0, 0, 0, 0);
        for (int a = 0; a < sc.n_area; ++a) { // muestreoSA for every area light (misSamplingFunctions.h:105-118)
            if ((a & 1) == 0) ra = philox_block(pixel, sample, depth, 2 + (a >> 1), lp.key0, lp.key1);
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
                if ((hit ? hid : 0) == lid) { // id stays 0 on a miss, samplingFunctions.h:196
                    const float cos_i = dot(n_, wi);
                    F3 f = albedo * kInvPi;
                    float gpdf = cos_i * kInvPi;
                    if (FACET) { const F3 wi_l = unit(to_local(fr, wi)); const F3 wh = unit(wi_l + wo_l); f = facet_brdf(obj, wi_l, wh, wo_l); gpdf = facet_pdf(wo_l, wh, obj.alpha); }
                    const float inv_fpdf = kTwoPi * omc_max;
                    const float wmis = power_heuristic(1.0f / inv_fpdf, gpdf);
                    L = L + had(had(mk(sm.lr, sm.lg, sm.lb), f), beta) * (cos_i * inv_fpdf * expf(-k.sigma_t * len2 * inv_len) * wmis * k.inv_cp);
                }
            }
        }
        const uint4 b1 = philox_block(pixel, sample, depth, 1, lp.key0, lp.key1);
        { // the BSDF-sampled term of MISv2 (:124-167): slots S_MIS = lanes 2,3 of block 1
            const float xi1 = u32_to_unit_f32(b1.z), xi2 = u32_to_unit_f32(b1.w);
            F3 wi_l, wh = mk(0, 0, 1);
            if (FACET) { wh = facet_normal(obj.alpha, xi1, xi2); wi_l = unit(fma3(wh, 2.0f * dot(wh, wo_l), -wo_l)); }
            else wi_l = cosine_local(xi1, xi2);
            const F3 wi = unit(to_world(fr, wi_l));
            float t; int hid;
            const bool hit = scan_sm(S, o, wi, t, hid);
            if (act) {
                ++scans;
                if (hit && S.mats[hid].emits) {
                    const MatF &em = S.mats[hid];
                    const F3 cx = mk(em.px, em.py, em.pz) - o;
                    float omc = one_minus_cos_max(em.r * em.r / dot(cx, cx));
                    if (FACET) {
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
        F3 weight;
        if (FACET) weight = bsdf_sample(obj, fr, wo_l, u32_to_unit_f32(b1.x), u32_to_unit_f32(b1.y), wi);
        else { wi = unit(to_world(fr, cosine_local(u32_to_unit_f32(b1.x), u32_to_unit_f32(b1.y)))); weight = albedo; } // c/pi * cos / (cos/pi)
        beta = had(beta, weight) * k.inv_cp;
        continue_or_end(act, slot, pixel, sample, meta + (1u << 10), wi, beta, L);
    }

    // ---- item bookkeeping ---------------------------------------------------------------------------------------------------------------
    // every thread: write the finished item's pixels and clear its accumulators; thread 0: load the next item into the slot
    __device__ __forceinline__ void flush_item(int b, float *__restrict__ hdr, int n_items) {
        const int item = S.t_item[b];
        for (int pl = tid; pl < item_pixels; pl += kSmThreads) {
            const long long pixel = item_pixel(item, pl);
            if (pixel >= 0) {
                float *out = hdr + pixel * 3;
                for (int c = 0; c < 3; ++c) out[c] = (float)((double)(long long)S.acc[b][pl][c] * kSmFixInv * lp.out_scale);
            }
            S.acc[b][pl][0] = 0ull; S.acc[b][pl][1] = 0ull; S.acc[b][pl][2] = 0ull;
        }
        __syncthreads(); // everyone has read t_item[b]
        if (tid == 0) {
            const int next = S.next_item;
            if (next < n_items) { S.t_item[b] = next; S.next_item = next + (int)gridDim.x; S.t_cursor[b] = 0u; S.t_done[b] = 0u; }
            else S.t_item[b] = -1;
        }
    }

    // ---- one round: thread 0 snapshots the queues and plans the generation ---------------------------------------------------------
    __device__ __forceinline__ void plan_round(unsigned item_total) {
        int flush = -1, gen = -1, active = 0;
        for (int b = 0; b < 2; ++b) {
            if (S.t_item[b] < 0) continue;
            ++active;
            if (S.t_cursor[b] == item_total) { if (S.t_done[b] == item_total && flush < 0) flush = b; }
            else if (gen < 0 || S.t_item[b] < S.t_item[gen]) gen = b;
        }
        S.flush_slot = flush;
        unsigned n_gen = 0;
        if (gen >= 0) {
            n_gen = min(S.free_tail - S.free_head, item_total - S.t_cursor[gen]);
            if (n_gen < item_total - S.t_cursor[gen]) n_gen &= ~31u; // full warps only, except for the last samples of an item
            S.gen_begin = S.t_cursor[gen];
            S.t_cursor[gen] += n_gen;
        }
        S.gen_slot = gen; S.gen_count = n_gen; S.gen_claim = 0u;
        // while new samples keep coming only full 32-record batches are handed out (the remainder waits for the next round);
        // once generation has stopped (an item drains) everything goes
        unsigned work = n_gen;
        for (int q = 0; q < SQ_COUNT; ++q) {
            const unsigned head = S.q_end[q]; // everything below the previous snapshot was consumed
            unsigned count = S.q_tail[q] - head;
            if (n_gen != 0u) count &= ~31u;
            S.q_head[q] = head; S.q_end[q] = head + count;
            work += count;
        }
        S.exit_flag = (work == 0u && flush < 0 && active == 0) ? 1 : 0;
    }

    // claim the next 32 entries of queue q below this round's snapshot; returns the number of entries (0: none left)
    __device__ __forceinline__ int claim(int q, unsigned end, unsigned &start) {
        if ((int)(end - *(volatile unsigned *)&S.q_head[q]) <= 0) return 0;
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(&S.q_head[q], 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        start = base;
        const int left = (int)(end - base);
        return left <= 0 ? 0 : min(left, 32);
    }

    __device__ __forceinline__ void run(float *__restrict__ hdr, int n_items) {
        const unsigned item_total = (unsigned)item_pixels * (unsigned)(lp.sample_end - lp.sample_begin);
        for (;;) {
            __syncthreads(); // (A) the previous round's pushes / releases / counters are visible
            if (tid == 0) plan_round(item_total);
            __syncthreads(); // (B) the plan is visible
            if (S.exit_flag) break;
            if (S.flush_slot >= 0) flush_item(S.flush_slot, hdr, n_items); // its records are all finished; the round below only touches the other item
            unsigned end[SQ_COUNT];
#pragma unroll
            for (int q = 0; q < SQ_COUNT; ++q) end[q] = S.q_end[q];
            const unsigned gen_count = S.gen_count, gen_begin = S.gen_begin;
            const int gen_slot = S.gen_slot;
            for (;;) { // claim batches, longest stage first, until the snapshot is used up
                unsigned start; int n;
                if ((n = claim(SQ_SURF_F, end[SQ_SURF_F], start)) > 0) { stage_surf<true>(lane < n ? (int)S.queue[SQ_SURF_F][(start + lane) & (kSmPool - 1)] : -1); continue; }
                if ((n = claim(SQ_SURF_L, end[SQ_SURF_L], start)) > 0) { stage_surf<false>(lane < n ? (int)S.queue[SQ_SURF_L][(start + lane) & (kSmPool - 1)] : -1); continue; }
                if ((n = claim(SQ_PRIMARY, end[SQ_PRIMARY], start)) > 0) { stage_primary(lane < n ? (int)S.queue[SQ_PRIMARY][(start + lane) & (kSmPool - 1)] : -1); continue; }
                if ((n = claim(SQ_MED_AREA, end[SQ_MED_AREA], start)) > 0) { stage_med<false>(lane < n ? (int)S.queue[SQ_MED_AREA][(start + lane) & (kSmPool - 1)] : -1); continue; }
                if ((n = claim(SQ_MED_POINT, end[SQ_MED_POINT], start)) > 0) { stage_med<true>(lane < n ? (int)S.queue[SQ_MED_POINT][(start + lane) & (kSmPool - 1)] : -1); continue; }
                if ((n = claim(SQ_SURF_P, end[SQ_SURF_P], start)) > 0) { stage_surf_p(lane < n ? (int)S.queue[SQ_SURF_P][(start + lane) & (kSmPool - 1)] : -1); continue; }
                if (*(volatile unsigned *)&S.gen_claim < gen_count) {
                    unsigned base = 0;
                    if (lane == 0) base = atomicAdd(&S.gen_claim, 32u);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (base < gen_count) { stage_gen(gen_slot, gen_begin + base, (int)min(32u, gen_count - base)); continue; }
                }
                break;
            }
        }
    }
};

} // namespace f32
} // namespace vpt
