// vpt_hbmwave.cuh -- the classic multi-kernel wavefront (VPT_KERNEL_WAVEFRONT_HBM): path state in HBM queues, one kernel per stage.
//
// north_star: "separate extend / medium-sample / NEE-shadow / scatter kernels compact SoA path-state queues with warp ballot / prefix
// scans ... achieved HBM GB/s on the wavefront queues".  Built as the measured alternative to the on-chip wavefront (vpt_smwave.cuh):
//   * a path record (60 B: origin, direction, throughput, pixel, sample, depth, picked source / hit object, the two uniforms drawn with the
//     roulette) lives IN the queue of the stage that has to process it next: structure-of-arrays per queue, so that both the reads of a stage
//     (thread i reads entry i) and its writes (compacted per block: match.any groups the lanes of a warp by destination queue, the groups
//     take offsets from shared-memory counters, ONE global atomic per block and destination reserves the entries -- per-warp global atomics
//     on the six queue counters serialise in L2 and were the bottleneck of the first version) are coalesced;
//   * one ROUND = generate (tops the pool up with new camera samples) -> PRIMARY -> SURF_P -> MED_POINT -> MED_AREA -> SURF_L -> SURF_F;
//     every kernel is a grid-stride loop over its input queue, whose length it reads from device memory, so the host enqueues rounds
//     without synchronising; no stage writes the queue it reads, and the kernel that follows a stage resets that stage's counter;
//   * per-pixel sums: 64-bit fixed-point (2^-30) global atomics (RED.64 in L2): order independent, bit-reproducible, and -- since both
//     wavefronts run the SAME stage code (vpt_stages.cuh) on the same Philox streams -- bit-identical to the on-chip wavefront's image.
#pragma once
#include "vpt_smwave.cuh"

namespace vpt {
namespace f32 {

enum : int { HF_OX = 0, HF_OY, HF_OZ, HF_DX, HF_DY, HF_DZ, HF_BR, HF_BG, HF_BB, HF_XD, HF_XS, HF_FLOATS };
enum : int { HU_PIXEL = 0, HU_SAMPLE, HU_DEPTH, HU_IDS, HU_WORDS };
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

__device__ __forceinline__ Rec hbm_load(const HbmState &H, int q, unsigned i) {
    const float *f = H.f + (size_t)q * HF_FLOATS * H.cap + i;
    const uint32_t *u = H.u + (size_t)q * HU_WORDS * H.cap + i;
    const size_t c = H.cap;
    Rec r;
    r.o = mk(f[HF_OX * c], f[HF_OY * c], f[HF_OZ * c]); r.d = mk(f[HF_DX * c], f[HF_DY * c], f[HF_DZ * c]);
    r.beta = mk(f[HF_BR * c], f[HF_BG * c], f[HF_BB * c]); r.xi_dist = f[HF_XD * c]; r.xi_decide = f[HF_XS * c];
    r.pixel = u[HU_PIXEL * c]; r.sample = u[HU_SAMPLE * c]; r.depth = u[HU_DEPTH * c];
    const uint32_t ids = u[HU_IDS * c];
    r.src = ids & 0xffu; r.hid = ids >> 8; r.aux = 0u;
    return r;
}
__device__ __forceinline__ void hbm_store(const HbmState &H, int q, unsigned i, const Rec &r) {
    float *f = H.f + (size_t)q * HF_FLOATS * H.cap + i;
    uint32_t *u = H.u + (size_t)q * HU_WORDS * H.cap + i;
    const size_t c = H.cap;
    f[HF_OX * c] = r.o.x; f[HF_OY * c] = r.o.y; f[HF_OZ * c] = r.o.z; f[HF_DX * c] = r.d.x; f[HF_DY * c] = r.d.y; f[HF_DZ * c] = r.d.z;
    f[HF_BR * c] = r.beta.x; f[HF_BG * c] = r.beta.y; f[HF_BB * c] = r.beta.z; f[HF_XD * c] = r.xi_dist; f[HF_XS * c] = r.xi_decide;
    u[HU_PIXEL * c] = r.pixel; u[HU_SAMPLE * c] = r.sample; u[HU_DEPTH * c] = r.depth; u[HU_IDS * c] = r.src | (r.hid << 8);
}

struct HbmCtx { // the stages' context (vpt_stages.cuh): Philox, the staged scene, global fixed-point pixel sums
    const SmScene &S;
    const ConstsF &k;
    const LaunchParams &lp;
    const HbmState &H;
    unsigned events = 0, scans = 0, nonfinite = 0, paths = 0;
    __device__ HbmCtx(const SmScene &S_, const ConstsF &k_, const LaunchParams &lp_, const HbmState &H_) : S(S_), k(k_), lp(lp_), H(H_) {}

    __device__ __forceinline__ float4 rnd(const Rec &r, uint32_t block) const {
        const uint4 b = philox_block(r.pixel, r.sample, r.depth, block, lp.key0, lp.key1);
        return make_float4(u32_to_unit_f32(b.x), u32_to_unit_f32(b.y), u32_to_unit_f32(b.z), u32_to_unit_f32(b.w));
    }
    __device__ __forceinline__ float4 jitter(const Rec &r) const {
        const uint4 b = philox_block(r.pixel, r.sample, kJitterBounce, 0, lp.key0, lp.key1);
        return make_float4(u32_to_unit_f32(b.x), u32_to_unit_f32(b.y), 0.0f, 0.0f);
    }
    __device__ __forceinline__ bool scan(F3 o, F3 d, float &t, int &id) { return scan_sm(S, o, d, t, id); }
    __device__ __forceinline__ void add(const Rec &r, F3 L) {
        if (!(fabsf(L.x) < kSmMaxContribution && fabsf(L.y) < kSmMaxContribution && fabsf(L.z) < kSmMaxContribution)) { ++nonfinite; return; }
        unsigned long long *a = H.acc + (size_t)r.pixel * 3;
        if (L.x != 0.0f) atomicAdd(a + 0, (unsigned long long)__float2ll_rn(L.x * kSmFixScale));
        if (L.y != 0.0f) atomicAdd(a + 1, (unsigned long long)__float2ll_rn(L.y * kSmFixScale));
        if (L.z != 0.0f) atomicAdd(a + 2, (unsigned long long)__float2ll_rn(L.z * kSmFixScale));
    }
    __device__ __forceinline__ void last_step() {}

    // ---- generation: camera sample g of the launch -> owned pixel g % n_owned, sample g / n_owned ----
    __device__ __forceinline__ int generate(bool mine, unsigned long long g, Rec &r) {
        bool inside = false;
        uint32_t pixel = 0u, sample = 0u;
        if (mine) {
            const unsigned long long s = g / H.n_owned_pixels;
            const unsigned op = (unsigned)(g - s * H.n_owned_pixels);                 // owned pixel index
            pixel = ((op >> 7) * (unsigned)lp.tile_count + (unsigned)lp.tile_rank) * (unsigned)kTile + (op & (unsigned)(kTile - 1));
            sample = (uint32_t)lp.sample_begin + (uint32_t)s;
            inside = pixel < (unsigned)lp.n_pixels;
        }
        if (inside) ++paths;
        r.aux = 0u;
        return stage_gen(*this, inside, pixel, sample, lp.width, lp.height, r) ? SQ_PRIMARY : -1;
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
