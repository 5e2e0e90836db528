// vpt_smwave.cuh -- the FP32 product pipeline of the SM-wide wavefront (VPT_KERNEL_WAVEFRONT_SM, the AUTO kernel): record layout and stage
// batches on top of the scheduler vpt_smsched.cuh.  The stage arithmetic is vpt_stages.cuh (shared with every other FP32 kernel and the
// unit kernels), the scan vpt_scan.cuh.
//   * kSmPool path records in shared memory, SoA, 52 B each: origin, direction, throughput, sample, one packed word (pixel-in-item, item
//     slot, picked source, hit object, depth) and the two uniforms drawn with the roulette.  No radiance in the record: a stage adds its
//     vertex's direct light to the pixel's 2^-30 fixed-point sum right away;
//   * a stage batch loads only the fields its stage reads and stores only what the stage changed.
#pragma once
#include "vpt_smsched.cuh"
#include "vpt_stages.cuh"

namespace vpt {
namespace f32 {

#ifndef VPT_SM_THREADS
#define VPT_SM_THREADS 768 // by measurement; tools/build_variant.py builds other values for comparison
#endif
#ifndef VPT_SM_POOL
#define VPT_SM_POOL 3072   // path records per CTA (multiple of 32; rings are indexed modulo this): 2048 -8.5 %, 2560 -3.4 %, 3200 +0.1 %
#endif
constexpr int kSmThreads = VPT_SM_THREADS;
constexpr int kSmPool = VPT_SM_POOL;
constexpr float kSmFixScale = 1073741824.0f; // 2^30
constexpr double kSmFixInv = 1.0 / 1073741824.0;
constexpr float kSmMaxContribution = 4294967296.0f; // 2^32: a contribution at or above it (or NaN) is dropped and counted (vpt_stats.nonfinite)

template <int SLOTS>
struct SmShared {
    SmScene scene; // first: scan_sm_call finds it at the start of the dynamic shared memory
    // ---- path records (SoA) ----
    float ox[kSmPool], oy[kSmPool], oz[kSmPool];
    float dx[kSmPool], dy[kSmPool], dz[kSmPool];
    float br[kSmPool], bg[kSmPool], bb[kSmPool]; // throughput
    uint32_t sample[kSmPool];
    uint32_t meta[kSmPool];                      // meta_pack(): pixel-in-item, item slot, picked source, hit object, depth
    float xd[kSmPool], xs[kSmPool];              // uniforms of slots 2 (distance) and 3 (decision) of the record's bounce
    SmCtl<kSmPool, SLOTS> ctl;
};
static_assert(sizeof(SmShared<kMaxItemSlots>) <= 232448, "one CTA per SM: at most 227 KB of shared memory");
template <int SLOTS>
__device__ __forceinline__ SmShared<SLOTS> &sm_shared() { return *reinterpret_cast<SmShared<SLOTS> *>(smwave_smem); }

// The pipeline, and at the same time the stages' context (vpt_stages.cuh): random numbers from Philox, scans over the staged scene,
// radiance into the work item's fixed-point sums, the claim for the next batch issued at a stage's last step (SmSched::last_step).
template <int METHOD, int SLOTS>
struct SmWave : SmSched<SmWave<METHOD, SLOTS>, kSmPool, kSmThreads, SLOTS> {
    using Base = SmSched<SmWave<METHOD, SLOTS>, kSmPool, kSmThreads, SLOTS>;
    using Base::lane; using Base::lp; using Base::route; using Base::count_done; using Base::alloc; using Base::pixel_of; using Base::item_pixel;
    using Base::last_step; using Base::log_p; using Base::item_pixels; using Base::Q;
    SmShared<SLOTS> &M;
    const SmScene &S;
    const ConstsF &k;
    unsigned events = 0, scans = 0, nonfinite = 0, paths = 0;

    __device__ SmWave(SmShared<SLOTS> &M_, const ConstsF &k_, const LaunchParams &lp_, int log_p_, int n_owned_, int zero)
        : Base(M_.ctl, lp_, log_p_, n_owned_, zero), M(M_), S(M_.scene), k(k_) {}

    // ---- the stages' context ------------------------------------------------------------------------------------------------------
    __device__ __forceinline__ float4 rnd(const Rec &r, uint32_t block) const {
#ifdef VPT_PHILOX_ARG_KEYS
        const uint4 b = philox_block(r.pixel, r.sample, r.depth, block, lp.key0, lp.key1);
#else
        const uint4 b = philox_block_ck(r.pixel, r.sample, r.depth, block);
#endif
        return make_float4(u32_to_unit_f32(b.x), u32_to_unit_f32(b.y), u32_to_unit_f32(b.z), u32_to_unit_f32(b.w));
    }
    __device__ __forceinline__ float4 jitter(const Rec &r) const {
#ifdef VPT_PHILOX_ARG_KEYS
        const uint4 b = philox_block(r.pixel, r.sample, kJitterBounce, 0, lp.key0, lp.key1);
#else
        const uint4 b = philox_block_ck(r.pixel, r.sample, kJitterBounce, 0);
#endif
        return make_float4(u32_to_unit_f32(b.x), u32_to_unit_f32(b.y), 0.0f, 0.0f);
    }
    __device__ __forceinline__ bool scan(F3 o, F3 d, float &t, int &id) { return scan_sm(S, o, d, t, id); }
    // radiance arriving at the record's pixel (rt.cpp:794)
    __device__ __forceinline__ void add(const Rec &r, F3 L) {
        if (!(fabsf(L.x) < kSmMaxContribution && fabsf(L.y) < kSmMaxContribution && fabsf(L.z) < kSmMaxContribution)) { ++nonfinite; return; } // NaN, Inf, out of range
        unsigned long long *a = Base::pixel_acc(r.aux);
        if (L.x != 0.0f) Base::add_fixed(a + 0, __float2ll_rn(L.x * kSmFixScale));
        if (L.y != 0.0f) Base::add_fixed(a + 1, __float2ll_rn(L.y * kSmFixScale));
        if (L.z != 0.0f) Base::add_fixed(a + 2, __float2ll_rn(L.z * kSmFixScale));
    }

    // a stage that ends with the roulette: the survivor's new direction / throughput / draws go back into its record
    __device__ __forceinline__ void finish_vertex(int slot, const Rec &r, int dest, bool store_beta) {
        if (dest == SQ_PRIMARY) {
            M.dx[slot] = r.d.x; M.dy[slot] = r.d.y; M.dz[slot] = r.d.z;
            if (store_beta) { M.br[slot] = r.beta.x; M.bg[slot] = r.beta.y; M.bb[slot] = r.beta.z; }
            M.xd[slot] = r.xi_dist; M.xs[slot] = r.xi_decide;
            M.meta[slot] = meta_pack(r.aux, r.src, 0u, r.depth);
        }
        route(dest, slot);
        count_done(dest == kDestFree, r.aux);
    }

    // ---- GEN: n <= 32 new camera samples of item slot b, lane i takes sample index g0 + i -------------------------------------------
    __device__ __forceinline__ void run_gen(int b, unsigned g0, int n) {
        const bool mine = lane < n;
        const unsigned g = g0 + (unsigned)lane;
        const int pl = (int)(g & (unsigned)(item_pixels - 1));
        const uint32_t sample = (uint32_t)lp.sample_begin + (g >> log_p);
        const int pixel = mine ? item_pixel(Q.t_item[b], pl) : -1;
        if (pixel >= 0) ++paths;
        Rec r;
        r.aux = meta_aux(pl, b);
        const bool alive = stage_gen(*this, pixel >= 0, (uint32_t)pixel, sample, lp.width, lp.height, r);
        const int slot = alloc(alive);
        if (alive) {
            M.ox[slot] = r.o.x; M.oy[slot] = r.o.y; M.oz[slot] = r.o.z;
            M.dx[slot] = r.d.x; M.dy[slot] = r.d.y; M.dz[slot] = r.d.z;
            M.br[slot] = 1.0f; M.bg[slot] = 1.0f; M.bb[slot] = 1.0f;
            M.sample[slot] = sample;
            M.xd[slot] = r.xi_dist; M.xs[slot] = r.xi_decide;
            M.meta[slot] = meta_pack(r.aux, r.src, 0u, 0u);
        }
        route(alive ? SQ_PRIMARY : -1, slot);
        Base::count_stillborn(b, mine, alive);
    }

    // n <= kTailGen samples, kTailGen / 32 per lane (g0 + lane, g0 + 32 + lane, ...), straight-line: the tail fill's batches
    __device__ __forceinline__ void run_gen_wide(int b, unsigned g0, int n) {
        constexpr int K = kTailGen / 32;
        bool mine[K], alive[K];
        uint32_t pixel[K], sample[K];
        Rec r[K];
        const int item = Q.t_item[b];
#pragma unroll
        for (int h = 0; h < K; ++h) {
            const unsigned g = g0 + (unsigned)(lane + 32 * h);
            const int pl = (int)(g & (unsigned)(item_pixels - 1));
            sample[h] = (uint32_t)lp.sample_begin + (g >> log_p);
            const int px = lane + 32 * h < n ? item_pixel(item, pl) : -1;
            mine[h] = px >= 0; pixel[h] = mine[h] ? (uint32_t)px : 0u;
            if (mine[h]) ++paths;
            r[h].aux = meta_aux(pl, b);
        }
        stage_gen_k<K>(*this, mine, pixel, sample, lp.width, lp.height, r, alive);
        int slot[K];
        Base::template alloc_push<K>(SQ_PRIMARY, alive, slot);
        unsigned dead = 0;
#pragma unroll
        for (int h = 0; h < K; ++h) {
            if (alive[h]) {
                const int s = slot[h];
                M.ox[s] = r[h].o.x; M.oy[s] = r[h].o.y; M.oz[s] = r[h].o.z;
                M.dx[s] = r[h].d.x; M.dy[s] = r[h].d.y; M.dz[s] = r[h].d.z;
                M.br[s] = 1.0f; M.bg[s] = 1.0f; M.bb[s] = 1.0f;
                M.sample[s] = sample[h];
                M.xd[s] = r[h].xi_dist; M.xs[s] = r[h].xi_decide;
                M.meta[s] = meta_pack(r[h].aux, r[h].src, 0u, 0u);
            }
            dead += (unsigned)__popc(__ballot_sync(0xffffffffu, lane + 32 * h < n && !alive[h]));
        }
        if (lane == 0 && dead) smem_red(&Q.t_done[b], dead, Base::lz);
    }

    // ---- the stage batches: load the fields the stage reads, run it (vpt_stages.cuh), store what it changed, route ---------------------
    // (idle lanes of a partial batch run on record 0's values: in range, never stored)
    template <int STAGE>
    __device__ __forceinline__ void run_stage(int slot) {
        const bool act = slot >= 0;
        const int s = act ? slot : 0;
        Rec r;
        const uint32_t meta = M.meta[s];
        r.aux = meta; r.src = (meta >> 10) & 31u; r.hid = (meta >> 15) & 31u; r.depth = meta >> 20;
        r.o = mk(M.ox[s], M.oy[s], M.oz[s]); r.beta = mk(M.br[s], M.bg[s], M.bb[s]);
        if (STAGE == SQ_PRIMARY) {
            r.d = mk(M.dx[s], M.dy[s], M.dz[s]);
            r.xi_dist = M.xd[s]; r.xi_decide = M.xs[s];
            const int dest = stage_primary<METHOD>(*this, act, r);
            if (dest >= 0 && dest != kDestFree) {
                M.ox[s] = r.o.x; M.oy[s] = r.o.y; M.oz[s] = r.o.z;
                if (dest == SQ_MED_POINT || dest == SQ_MED_AREA) { M.br[s] = r.beta.x; M.bg[s] = r.beta.y; M.bb[s] = r.beta.z; }
                else M.meta[s] = meta_pack(meta, r.src, r.hid, r.depth);
            }
            last_step();
            route(dest, slot);
            count_done(dest == kDestFree, meta);
        } else if (STAGE == SQ_MED_POINT || STAGE == SQ_MED_AREA) {
            r.sample = M.sample[s]; r.pixel = pixel_of(meta);
            const int dest = stage_med<STAGE == SQ_MED_POINT>(*this, act, r);
            finish_vertex(slot, r, dest, false);
        } else if (STAGE == SQ_SURF_P) {
            r.d = mk(M.dx[s], M.dy[s], M.dz[s]);
            const int dest = stage_surf_p(*this, act, r);
            last_step();
            route(dest, slot);
        } else {
            r.d = mk(M.dx[s], M.dy[s], M.dz[s]);
            r.sample = M.sample[s]; r.pixel = pixel_of(meta);
            const int dest = stage_surf<STAGE == SQ_SURF_F>(*this, act, r);
            finish_vertex(slot, r, dest, true);
        }
    }
};

} // namespace f32
} // namespace vpt
