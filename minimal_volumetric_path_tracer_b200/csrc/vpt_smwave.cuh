// vpt_smwave.cuh -- SM-wide wavefront FP32 kernel (VPT_KERNEL_WAVEFRONT_SM, the AUTO kernel): one persistent CTA per SM, path state on chip.
//
// Why (profiles/r1_summary.md): the warp-local wavefront (vpt_wavefront.cuh) fixed SIMT efficiency (28 of 32 lanes active) but used only
// 40 % of the issue slots -- 42 % of all stall samples were `no_inst`: sixteen warps per SM, each in a different stage of a 62 KB
// kernel, thrash the 32 KB L1.5 / 6 KB L0 instruction caches.  Here the warps of an SM share ONE pool of path records and move through
// the stages together:
//   * one CTA of kSmThreads threads per SM, grid = number of SMs (persistent); work items = groups of pixel tiles x all samples, handed
//     out statically (item j -> CTA j % gridDim.x), two items in flight per CTA so that a draining item overlaps the next one;
//   * kSmPool path records in shared memory (SoA, 68 B each), one index queue (ring) per stage, one ring of free records;
//   * work proceeds in ROUNDS: at a barrier warp 0 snapshots every queue's tail (one load of the 20 control words, then shuffles) into a
//     table of 32-record batches ordered longest stage first; every warp claims batches with ONE shared-memory atomic each, runs them
//     and routes the survivors to the next stage's ring (match.any groups the lanes by destination: one atomic per group); what is
//     pushed during a round is consumed in the next one.  Only full batches are handed out while samples remain.  A warp that finds
//     the table used up generates new camera samples into the free records instead of idling at the barrier (tail fill).  The rank order
//     also keeps the 24 warps inside two or three stages at a time -- a barrier-free variant lost 40 % to instruction-cache misses;
//   * queues are split by what diverges: medium vertex with a point / an area source, surface vertex needing the pLight shadow ray,
//     Lambert / microfacet surface vertex;
//   * the scene's scan records are staged in shared memory as float4 (broadcast LDS.128); ordinary spheres (r < 64) use the direct
//     roots -b -+ sqrt(det) with det = r^2 - |op - (op.d)d|^2, huge ones the re-anchored cancellation-free form of vpt_f32.cuh; the
//     nearest accepted root is selected with one three-input unsigned minimum per sphere (see scan_sm_call).
// Random-number slots, formulas and semantics are exactly those of vpt_f32.cuh (file:line citations there); per-pixel sums are 64-bit
// fixed-point (2^-30) accumulators in shared memory, built from native 32-bit atomics: integer adds are order independent, so images
// are bit-reproducible although the order in which paths finish is data dependent.
#pragma once
#include "vpt_mega_scan.cuh"

namespace vpt {
namespace f32 {

#ifndef VPT_SM_THREADS
#define VPT_SM_THREADS 768 // by measurement; tools/build_variant.py builds other values for comparison
#endif
constexpr int kSmThreads = VPT_SM_THREADS;
constexpr int kSmPool = 2048;         // path records per CTA (power of two)
constexpr int kSmMaxItemPixels = 256; // pixels per work item (power of two multiple of kTile)
constexpr float kSmFixScale = 1073741824.0f; // 2^30
constexpr double kSmFixInv = 1.0 / 1073741824.0;
constexpr float kSimpleRootMaxR2 = 64.0f * 64.0f;

enum : int { SQ_PRIMARY = 0, SQ_MED_POINT, SQ_MED_AREA, SQ_SURF_P, SQ_SURF_L, SQ_SURF_F, SQ_COUNT };
// Claim order of a round's batches (one nibble per rank, SQ_COUNT = generation), longest stage first so that a round ends evenly:
// SURF_F, SURF_L, PRIMARY, MED_AREA, MED_POINT, SURF_P, generation
constexpr unsigned kRankStage = 0x6312045u;

// the scene as the scan and the shading code read it; first in the CTA's shared memory (the unit kernels stage only this part).
// Scan records are stored in PAIRS for the packed FP32 instructions of sm_100 (FFMA2 / FADD2 / FMUL2: two lanes per issue slot -- the
// kernel is issue bound, not FMA-pipe bound, profiles/r1_summary.md): component c of spheres 2j and 2j+1 sits in one 64-bit half of a float4.
struct SmScene {
    MatF mats[kMaxSpheres];
    float4 ga[2 * kMaxSpheres]; // general-form pair j: (qx0 qx1 qy0 qy1) (qz0 qz1 c0_0 c0_1) (mx0 mx1 my0 my1) (mz0 mz1 - -)
    float4 gb[kMaxSpheres];     // direct-root pair j:  (px0 px1 py0 py1) (pz0 pz1 r2_0 r2_1)
    int gid[2 * kMaxSpheres + 2]; // scan slot -> caller's sphere index (general pairs first; -1: the padding slot of an odd class)
    int n_pa, n_pb;             // pairs per class
};
// cooperative staging by the whole block (call, then __syncthreads).  Scan order: general-form spheres (huge / re-anchored ones and anything
// with r >= 64) first, in scene order, then the direct-root ones; every scan record finds its place with one pass over its predecessors.
// An odd class is padded with a record that no ray can hit (negative discriminant for every ray).
__device__ __forceinline__ void stage_scene(SmScene &S, const SceneF &sc, int tid, int n_threads) {
    for (int i = tid; i < sc.n_spheres * (int)(sizeof(MatF) / 4); i += n_threads)
        reinterpret_cast<uint32_t *>(S.mats)[i] = reinterpret_cast<const uint32_t *>(sc.mat)[i];
    int n_general = 0;
    for (int g = 0; g < sc.n_geom; ++g) n_general += (sc.geom[g].big || sc.geom[g].r2 >= kSimpleRootMaxR2);
    const int n_direct = sc.n_geom - n_general, n_pa = (n_general + 1) >> 1, n_pb = (n_direct + 1) >> 1;
    float *ga = reinterpret_cast<float *>(S.ga), *gb = reinterpret_cast<float *>(S.gb);
    if (tid < sc.n_geom) {
        const GeomF &G = sc.geom[tid];
        const bool general = G.big || G.r2 >= kSimpleRootMaxR2;
        int before_same = 0;
        for (int g = 0; g < tid; ++g) before_same += ((sc.geom[g].big || sc.geom[g].r2 >= kSimpleRootMaxR2) == general);
        const int pair = before_same >> 1, h = before_same & 1;
        if (general) {
            float *r = ga + 16 * pair + h;
            r[0] = G.qx; r[2] = G.qy; r[4] = G.qz; r[6] = G.c0; r[8] = G.mx; r[10] = G.my; r[12] = G.mz; r[14] = 0.0f;
            S.gid[before_same] = G.id;
        } else {
            float *r = gb + 8 * pair + h;
            r[0] = G.qx; r[2] = G.qy; r[4] = G.qz; r[6] = G.r2;
            S.gid[2 * n_pa + before_same] = G.id;
        }
    }
    if (tid == 0) {
        S.n_pa = n_pa; S.n_pb = n_pb;
        if (n_general & 1) { // c = |oq|^2 + 1e30 > b^2: never hit
            float *r = ga + 16 * (n_pa - 1) + 1;
            r[0] = 0.0f; r[2] = 0.0f; r[4] = 0.0f; r[6] = 1e30f; r[8] = 0.0f; r[10] = 0.0f; r[12] = 0.0f; r[14] = 0.0f;
            S.gid[n_general] = -1;
        }
        if (n_direct & 1) { // r^2 = -1: never hit
            float *r = gb + 8 * (n_pb - 1) + 1;
            r[0] = 0.0f; r[2] = 0.0f; r[4] = 0.0f; r[6] = -1.0f;
            S.gid[2 * n_pa + n_direct] = -1;
        }
    }
}

struct SmShared {
    SmScene scene; // first: scan_sm_call and the unit kernels find it at the start of the dynamic shared memory
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
    // ---- control ----
    // this round's batch table by rank (kRankStage): batch k belongs to the last rank with rb_first[rank] <= k and covers entries
    // [rb_begin + 32 j, min(rb_end, +32)), j = k - rb_first
    __align__(16) unsigned rb_first[8]; // [7] = number of batches of the round
    unsigned rb_begin[8], rb_end[8];
    unsigned round_claim;      // next unclaimed batch of the round: one atomicAdd per batch
    unsigned tail_budget, tail_used; // camera samples that warps out of batches may still generate in this round (tail fill)
    // the 20 words the round plan reads, contiguous: warp 0 fetches them with ONE load (lane i reads word i, see plan_round)
    __align__(16) unsigned q_tail[SQ_COUNT]; // [0..5]   push counters
    unsigned q_end[SQ_COUNT];                // [6..11]  entries below it have been handed out
    unsigned free_head, free_tail;           // [12,13]  the free-record ring: allocate at the head (only below the round's snapshot), release at the tail
    int t_item[2];                           // [14,15]  work item of the slot, -1: slot idle
    unsigned t_cursor[2], t_done[2];         // [16..19] camera samples generated / paths finished
    unsigned ctl_pad[12];
    int gen_slot, flush_slot, exit_flag;
#ifdef VPT_SMWAVE_PROFILE
    long long dbg_arrive[32];
#endif
    int next_item;
};
static_assert(offsetof(SmShared, t_done) - offsetof(SmShared, q_tail) == 18 * sizeof(unsigned), "plan_round reads the control words by index");
static_assert(offsetof(SmShared, free_tail) - offsetof(SmShared, q_tail) == 13 * sizeof(unsigned) && offsetof(SmShared, freelist) - offsetof(SmShared, queue) == SQ_COUNT * kSmPool * sizeof(uint16_t), "route() addresses the free ring as queue number SQ_COUNT");

// shared-memory atomic add issued by ONE lane (the callers aggregate over the warp themselves): plain ATOMS.ADD, without the
// compiler's own warp-aggregation wrapper around atomicAdd
// `lane_zero` = laneid * (a kernel argument that is always 0): ptxas wraps an atomic on a provably warp-uniform address in its own
// leader election (VOTEU / FLO / POPC / S2R / SHFL, ~13 instructions per site, 7 % of all executed instructions in the first
// profile); the callers have already elected lane 0, so the address is made formally lane dependent.
__device__ __forceinline__ unsigned smem_add(unsigned *p, unsigned v, unsigned lane_zero = 0u) {
    unsigned old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"((unsigned)__cvta_generic_to_shared(p) + lane_zero), "r"(v));
    return old;
}
__device__ __forceinline__ void smem_red(unsigned *p, unsigned v, unsigned lane_zero = 0u) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p) + lane_zero), "r"(v));
}

extern __shared__ __align__(16) unsigned char smwave_smem[]; // the CTA's one SmShared (dynamic shared memory)
__device__ __forceinline__ SmShared &sm_shared() { return *reinterpret_cast<SmShared *>(smwave_smem); }

// nearest accepted hit over all spheres (pathTracingUtilities.h:12-36 with Sphere.h:27-37): distance (+inf: none) and scan index.
// One out-of-line copy: it is called from seven places and must stay resident in the instruction cache.
struct ScanHit { float t; int index; };
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; } // MUFU.RCP, as __fdividef
// sign flip that ptxas folds into the operand modifiers of the packed instructions: `neg.f32` without .ftz (under -ftz=true the compiler's
// own negation is neg.ftz = a separate flushing FADD per lane; the consuming .FTZ instruction flushes anyway)
__device__ __forceinline__ float neg_fold(float x) { float r; asm("neg.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(neg_fold(a.x), neg_fold(a.y)); }
__device__ __forceinline__ float2 lo2(float4 v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(float4 v) { return make_float2(v.z, v.w); }
static __device__ __noinline__ ScanHit scan_sm_call(float ox, float oy, float oz, float dx, float dy, float dz) {
    const SmScene &S = *reinterpret_cast<const SmScene *>(smwave_smem);
    // Selection without compares: the reference accepts the near root unless it is below 1e-4, else the far one, and then requires
    // t > 1e-4 (Sphere.h:34, pathTracingUtilities.h:20) = the smallest root above 1e-4.  For w = root - 1e-4 the valid candidates are
    // exactly the positive floats, whose bit patterns order like unsigned integers, while negative values (sign bit) and the NaN of a
    // negative discriminant (0x7fffffff) compare above +inf: ONE three-input unsigned minimum per sphere replaces four compares and
    // selects on the half-rate ALU pipe (ncu: ALU 41 % busy, FMA 25 %); the index follows with one compare and one select.
    // Two spheres per iteration: every add / multiply / fma below is one packed instruction for both (same IEEE roundings as the scalar
    // form, so the roots are bit-identical to a one-sphere-at-a-time scan); the ray's components are broadcast operands.
    unsigned best = 0x7f800000u; // +inf
    int bi = -1;
    const int na = S.n_pa, nb = S.n_pb;
    const float2 ox2 = make_float2(ox, ox), oy2 = make_float2(oy, oy), oz2 = make_float2(oz, oz);
    const float2 dx2 = make_float2(dx, dx), dy2 = make_float2(dy, dy), dz2 = make_float2(dz, dz);
    const float2 meps = make_float2(-kEps, -kEps);
    for (int j = 0; j < na; ++j) {
        const float4 A = S.ga[4 * j], B = S.ga[4 * j + 1], C = S.ga[4 * j + 2], E = S.ga[4 * j + 3];
        const float2 mx = lo2(C), my = hi2(C), mz = lo2(E);
        const float2 oqx = __fadd2_rn(ox2, neg2(lo2(A))), oqy = __fadd2_rn(oy2, neg2(hi2(A))), oqz = __fadd2_rn(oz2, neg2(lo2(B)));
        const float2 opx = __fadd2_rn(oqx, mx), opy = __fadd2_rn(oqy, my), opz = __fadd2_rn(oqz, mz);
        const float2 b = __ffma2_rn(opx, dx2, __ffma2_rn(opy, dy2, __fmul2_rn(opz, dz2)));
        const float2 c = __ffma2_rn(oqx, __fadd2_rn(opx, mx), __ffma2_rn(oqy, __fadd2_rn(opy, my), __ffma2_rn(oqz, __fadd2_rn(opz, mz), hi2(B)))); // |op|^2 - r^2 without cancellation
        const float2 det = __ffma2_rn(b, b, neg2(c));
        const float2 sq = __fmul2_rn(det, make_float2(rsqrtf(det.x), rsqrtf(det.y))); // NaN when det <= 0
        const float2 q = __fadd2_rn(neg2(b), neg2(make_float2(copysignf(sq.x, b.x), copysignf(sq.y, b.y)))); // the root without cancellation; the other one is c / q
        const float2 w1 = __fadd2_rn(q, meps), w2 = __ffma2_rn(c, make_float2(rcp_approx(q.x), rcp_approx(q.y)), meps);
        unsigned k = __vimin3_u32(best, __float_as_uint(w1.x), __float_as_uint(w2.x));
        if (k != best) bi = 2 * j;
        best = k;
        k = __vimin3_u32(best, __float_as_uint(w1.y), __float_as_uint(w2.y));
        if (k != best) bi = 2 * j + 1;
        best = k;
    }
    for (int j = 0; j < nb; ++j) {
        const float4 A = S.gb[2 * j], B = S.gb[2 * j + 1];
        const float2 oqx = __fadd2_rn(ox2, neg2(lo2(A))), oqy = __fadd2_rn(oy2, neg2(hi2(A))), oqz = __fadd2_rn(oz2, neg2(lo2(B)));
        const float2 b = __ffma2_rn(oqx, dx2, __ffma2_rn(oqy, dy2, __fmul2_rn(oqz, dz2)));
        const float2 nb2 = neg2(b);
        const float2 lx = __ffma2_rn(dx2, nb2, oqx), ly = __ffma2_rn(dy2, nb2, oqy), lz = __ffma2_rn(dz2, nb2, oqz);
        const float2 det = __ffma2_rn(neg2(lx), lx, __ffma2_rn(neg2(ly), ly, __ffma2_rn(neg2(lz), lz, hi2(B))));
        const float2 sq = __fmul2_rn(det, make_float2(rsqrtf(det.x), rsqrtf(det.y)));
        const float2 nbe = __fadd2_rn(nb2, meps);
        const float2 w1 = __fadd2_rn(nbe, neg2(sq)), w2 = __fadd2_rn(nbe, sq);
        unsigned k = __vimin3_u32(best, __float_as_uint(w1.x), __float_as_uint(w2.x));
        if (k != best) bi = 2 * (na + j);
        best = k;
        k = __vimin3_u32(best, __float_as_uint(w1.y), __float_as_uint(w2.y));
        if (k != best) bi = 2 * (na + j) + 1;
        best = k;
    }
    return ScanHit{__uint_as_float(best) + kEps, bi};
}
__device__ __forceinline__ bool scan_sm(const SmScene &S, F3 o, F3 d, float &t, int &id) {
    const ScanHit h = scan_sm_call(o.x, o.y, o.z, d.x, d.y, d.z);
    t = h.t;
    id = h.index >= 0 ? S.gid[h.index] : -1;
    return h.index >= 0;
}

// optional in-kernel timing (-DVPT_SMWAVE_PROFILE): per-warp cycle sums, added to Counters::dbg at the end
//   dbg[0..7]  cycles inside batches of stage q (SQ_* order, 6 = generation, 7 = tail-fill generation)     dbg[8..15]  batches of stage q
//   dbg[16] cycles waiting at barrier (A)   dbg[17] cycles from (A) to (B) (planning)   dbg[18] cycles in the claim loop outside batches
//   dbg[19] total cycles of all warps   dbg[20] rounds (per CTA, summed)   dbg[21] cycles flushing
#ifdef VPT_SMWAVE_PROFILE
#define SMW_T(var) const long long var = clock64()
#define SMW_ADD(i, v) prof[i] += (unsigned long long)(v)
#else
#define SMW_T(var)
#define SMW_ADD(i, v)
#endif

template <int METHOD>
struct SmWave {
    SmShared &S;
    const SceneF &sc;
    const ConstsF &k;
    const LaunchParams &lp;
    const int tid, lane;
    const unsigned lz; // lane * 0, opaque to the compiler (see smem_add)
    const int log_p, item_pixels, n_owned_tiles;
    unsigned events = 0, scans = 0, nonfinite = 0, paths = 0;
    unsigned next_raw = 0; // lane 0: the next batch of the round, claimed while the tail of the current one is still running (claim_early)
#ifdef VPT_SMWAVE_PROFILE
    unsigned long long prof[24] = {};
#endif

    __device__ SmWave(SmShared &S_, const SceneF &sc_, const ConstsF &k_, const LaunchParams &lp_, int log_p_, int n_owned_, int zero)
        : S(S_), sc(sc_), k(k_), lp(lp_), tid((int)threadIdx.x), lane((int)threadIdx.x & 31), lz((threadIdx.x & 31u) * (unsigned)zero), log_p(log_p_),
          item_pixels(1 << log_p_), n_owned_tiles(n_owned_) {}

    // ---- work items: item j = owned tiles [j * K, (j + 1) * K), K = item_pixels / kTile -----------------------------------------
    // (32-bit arithmetic: n_pixels is an int32, so tile and pixel indices fit; -1 = outside the image / not this rank's tile)
    __device__ __forceinline__ int item_pixel(int item, int pl) const {
        const unsigned owned = ((unsigned)item << (log_p - 7)) + ((unsigned)pl >> 7);
        if (owned >= (unsigned)n_owned_tiles) return -1;
        const unsigned pixel = (owned * (unsigned)lp.tile_count + (unsigned)lp.tile_rank) * (unsigned)kTile + ((unsigned)pl & (unsigned)(kTile - 1));
        return pixel < (unsigned)lp.n_pixels ? (int)pixel : -1;
    }

    // ---- queue / pool primitives (warp-aggregated shared-memory atomics) ------------------------------------------------------------
    __device__ __forceinline__ void push(int q, bool flag, int slot) {
        const unsigned m = __ballot_sync(0xffffffffu, flag);
        if (m == 0u) return;
        unsigned base = 0;
        if (lane == 0) base = smem_add(&S.q_tail[q], (unsigned)__popc(m), lz);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (flag) S.queue[q][(base + __popc(m & ((1u << lane) - 1u))) & (kSmPool - 1)] = (uint16_t)slot;
    }
    // Route every lane's record in ONE step: dest = a stage queue (SQ_*), kFree (the record goes back to the free ring) or -1 (nothing).
    // Lanes with the same destination find each other with match.any; the lowest lane of each group reserves the group's entries
    // with one atomic (all group leaders in the same instruction), the others take their rank behind it.  Replaces up to six
    // ballot / atomic / store sequences per stage (PRIMARY feeds five queues and the free ring).
    static constexpr int kFree = SQ_COUNT;
    __device__ __forceinline__ void route(int dest, int slot) {
        const unsigned grp = __match_any_sync(0xffffffffu, dest);
        if (dest < 0) return;
        const int leader = __ffs(grp) - 1;
        unsigned base = 0;
        // counters: q_tail[0..5] are words 0..5 of the control block, free_tail is word 13; rings: queue[0..5] and, right behind them, freelist
        if (lane == leader) base = smem_add(&S.q_tail[0] + (dest == kFree ? 13 : dest), (unsigned)__popc(grp), lz);
        base = __shfl_sync(grp, base, leader);
        (&S.queue[0][0])[dest * kSmPool + ((base + __popc(grp & ((1u << lane) - 1u))) & (kSmPool - 1))] = (uint16_t)slot;
    }
    __device__ __forceinline__ int alloc(bool flag) { // the round's snapshot guarantees enough free records below free_tail
        const unsigned m = __ballot_sync(0xffffffffu, flag);
        if (m == 0u) return -1;
        unsigned base = 0;
        if (lane == 0) base = smem_add(&S.free_head, (unsigned)__popc(m), lz);
        base = __shfl_sync(0xffffffffu, base, 0);
        return flag ? (int)S.freelist[(base + __popc(m & ((1u << lane) - 1u))) & (kSmPool - 1)] : -1;
    }
    // a path ended: add its radiance to the pixel (rt.cpp:794) -- the caller reports it to count_done()
    __device__ __forceinline__ void add_radiance(uint32_t meta, F3 L) {
        if (!isfinite(L.x + L.y + L.z)) { ++nonfinite; return; }
        unsigned long long *a = S.acc[(meta >> 9) & 1u][meta & 0x1ffu];
        if (L.x != 0.0f) add_fixed(a + 0, __float2ll_rn(L.x * kSmFixScale));
        if (L.y != 0.0f) add_fixed(a + 1, __float2ll_rn(L.y * kSmFixScale));
        if (L.z != 0.0f) add_fixed(a + 2, __float2ll_rn(L.z * kSmFixScale));
    }
    // 64-bit two's-complement add from native 32-bit shared-memory atomics (a 64-bit atomicAdd on shared memory is a CAS loop):
    // low word first, its carry goes into the high word; the sum modulo 2^64 does not depend on the order of the adds
    static __device__ __forceinline__ void add_fixed(unsigned long long *acc, long long v) {
        unsigned *w = reinterpret_cast<unsigned *>(acc);
        const unsigned lo = (unsigned)v;
        unsigned hi = (unsigned)((unsigned long long)v >> 32);
        const unsigned old = smem_add(w, lo);
        hi += (old + lo < old) ? 1u : 0u;
        if (hi) smem_red(w + 1, hi);
    }
    __device__ __forceinline__ void count_done(bool ended, uint32_t meta) {
        const unsigned m0 = __ballot_sync(0xffffffffu, ended && ((meta >> 9) & 1u) == 0u);
        const unsigned m1 = __ballot_sync(0xffffffffu, ended && ((meta >> 9) & 1u) == 1u);
        if (lane == 0) {
            if (m0) smem_red(&S.t_done[0], (unsigned)__popc(m0), lz);
            if (m1) smem_red(&S.t_done[1], (unsigned)__popc(m1), lz);
        }
    }
    __device__ __forceinline__ uint32_t pixel_of(uint32_t meta) const {
        return (uint32_t)item_pixel(S.t_item[(meta >> 9) & 1u], (int)(meta & 0x1ffu));
    }
    // The claim for the NEXT batch is issued a few hundred cycles before the current one ends (at the start of its last step: the roulette's
    // Philox block or the final routing), so that the atomic's round trip is over when the claim loop needs it; early enough to hide the
    // latency, late enough not to commit a warp to work while others idle (claiming at the START of a batch measured 15 % slower).
    __device__ __forceinline__ void claim_early() { if (lane == 0) next_raw = smem_add(&S.round_claim, 1u, lz); }
    // roulette for the next bounce (vptShadeMethods.h:1282); a surviving record (already holding o) gets its new direction,
    // throughput, radiance-so-far and the block-0 words of the new bounce
    __device__ __forceinline__ void continue_or_end(bool act, int slot, uint32_t pixel, uint32_t sample, uint32_t meta, F3 d, F3 beta, F3 L) {
        claim_early();
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
        route(alive ? SQ_PRIMARY : (act ? kFree : -1), slot);
        count_done(act && !alive, meta);
    }

    // ---- GEN: n <= 32 new camera samples of item slot b, lane i takes sample index g0 + i -------------------------------------------
    __device__ __forceinline__ void stage_gen(int b, unsigned g0, int n) {
        const bool mine = lane < n;
        const unsigned g = g0 + (unsigned)lane;
        const int pl = (int)(g & (unsigned)(item_pixels - 1));
        const uint32_t sample = (uint32_t)lp.sample_begin + (g >> log_p);
        const int pixel = mine ? item_pixel(S.t_item[b], pl) : -1;
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
            const int row = (int)((unsigned)pixel / (unsigned)lp.width), col = pixel - row * lp.width;
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
        if (lane == 0 && m) smem_red(&S.t_done[b], (unsigned)__popc(m), lz);
    }

    // ---- PRIMARY: scan of the path ray, light pick, distance sampling, surface-or-medium decision --------------------------------
    __device__ __forceinline__ void stage_primary(int slot) {
        const bool act = slot >= 0;
        const int s = act ? slot : 0;
        F3 o = mk(S.ox[s], S.oy[s], S.oz[s]);
        const F3 d = mk(S.dx[s], S.dy[s], S.dz[s]);
        float t; int hid;
        const bool hit = scan_sm(S.scene, o, d, t, hid);
        bool to_mp = false, to_ma = false, to_sp = false, to_sl = false, to_sf = false, ended = false;
        const uint32_t meta = S.meta[s];
        if (act) {
            ++scans; ++events;
            if (!hit) { t = kMaxFloat; hid = 0; }
            const int pick = min((int)(u32_to_unit_f32(S.r1[s]) * k.n_emitters), sc.n_emitters - 1);
            const int src = sc.emitters[pick];
            const MatF &sm = S.scene.mats[src];
            bool surface; float dist, inv_pdf = 1.0f;
            if (METHOD == 0) {
                dist = -logf(1.0f - u32_to_unit_f32(S.r2[s])) * k.inv_sigma_t; // freeFlightSample, vptSamplingFunctions.h:11
                surface = dist > t;
            } else if (METHOD == 4) { // distance-sampling MIS (vpt_f32.cuh mis_distance)
                surface = mis_distance(mk(sm.px, sm.py, sm.pz), o, d, t, __expf(-k.sigma_t * t), k.sigma_t, k.inv_sigma_t, u32_to_unit_f32(S.r2[s]), u32_to_unit_f32(S.r3[s]), dist, inv_pdf);
            } else { // equiAngularParams2 (volumetricBasicFunctions.h:209-223) + equiAngularProb (vptSamplingFunctions.h:60)
                const float Tr = __expf(-k.sigma_t * t);
                float D, dth, tl;
                dist = equiangular_sample(mk(sm.px, sm.py, sm.pz), o, d, t, u32_to_unit_f32(S.r2[s]), D, dth, tl);
                inv_pdf = dth * (tl * tl + D * D) / (D * (1.0f - Tr));
                const float xs = u32_to_unit_f32(S.r3[s]);
                surface = (METHOD == 1) ? (xs <= Tr) : (xs < Tr);
            }
            const MatF &obj = S.scene.mats[hid];
            if (surface && obj.emits) { // :1308-1313: a directly seen emitter ends the path
                const F3 L = (meta >> 10) == 0 ? had(mk(obj.lr, obj.lg, obj.lb), mk(S.br[s], S.bg[s], S.bb[s])) : mk(S.lr[s], S.lg[s], S.lb[s]);
                add_radiance(meta, L);
                ended = true;
            } else if (surface) {
                o = fma3(d, t, o);
                const F3 lx = mk(sm.px, sm.py, sm.pz) - o;
                to_sp = !(sm.r > 0.0f && dot(lx, lx) > sm.r * sm.r); // pLight is zero for an area source seen from outside it
                to_sf = !to_sp && obj.material != 0; // microfacet and dielectric
                to_sl = !to_sp && !to_sf;
                S.r1[s] = (uint32_t)src | ((uint32_t)hid << 8);
            } else {
                o = fma3(d, dist, o);
                const float w = (METHOD == 0) ? k.albedo_over_cp : k.sigma_s * __expf(-k.sigma_t * fabsf(dist)) * inv_pdf * k.inv_cp;
                S.r2[s] = __float_as_uint(w);
                S.r1[s] = (uint32_t)src;
                to_mp = sm.r == 0.0f; to_ma = !to_mp;
            }
            if (!ended) { S.ox[s] = o.x; S.oy[s] = o.y; S.oz[s] = o.z; }
        }
        claim_early();
        route(to_mp ? SQ_MED_POINT : to_ma ? SQ_MED_AREA : to_sp ? SQ_SURF_P : to_sl ? SQ_SURF_L : to_sf ? SQ_SURF_F : ended ? kFree : -1, slot);
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
        const MatF &sm = S.scene.mats[src];
        const F3 light = mk(sm.px, sm.py, sm.pz);
        const F3 lx = light - o;
        const float d2 = dot(lx, lx), inv = rsqrtf(d2);
        F3 qo, qd, C; float lim = 0.0f;
        if (POINT) {
            const float dist = d2 * inv;
            C = had(mk(sm.lr, sm.lg, sm.lb), beta) * (__expf(-k.sigma_t * dist) / d2 * k.n_emitters * kInv4Pi * w);
            qo = light; qd = lx * (-inv); lim = dist * (1.0f - 1e-4f);
        } else {
            const float omc_max = one_minus_cos_max(sm.r * sm.r / d2);
            qd = cone_sample(lx * inv, omc_max, u32_to_unit_f32(b1.x), u32_to_unit_f32(b1.y));
            qo = o;
            C = had(mk(sm.lr, sm.lg, sm.lb), beta) * (kInv4Pi * kTwoPi * omc_max * k.n_emitters * w);
        }
        float t; int hid;
        const bool hit = scan_sm(S.scene, qo, qd, t, hid);
        if (act) {
            ++scans;
            if (POINT) { if (!hit || t > lim) L = L + C; }
            else if (hit && hid == src) L = L + C * __expf(-k.sigma_t * t);
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
        const MatF &sm = S.scene.mats[ids & 0xffu];
        const MatF &obj = S.scene.mats[ids >> 8];
        const F3 light = mk(sm.px, sm.py, sm.pz);
        const F3 lx = light - o;
        const float d2 = dot(lx, lx), inv = rsqrtf(d2), dist = d2 * inv;
        const F3 n_ = unit(o - mk(obj.px, obj.py, obj.pz));
        const F3 wi = lx * inv;
        const bool facet = obj.material == 1;
        F3 f = mk(obj.cr, obj.cg, obj.cb) * kInvPi;
        if (facet) f = facet_eval_world(obj, n_, wi, d);
        const F3 C = had(had(mk(sm.lr, sm.lg, sm.lb), f), beta) * (dot(n_, wi) * __expf(-k.sigma_t * dist) / d2 * k.n_emitters * k.inv_cp);
        const F3 qd = lx * (-inv);
        float t; int hid;
        const bool hit = scan_sm(S.scene, light, qd, t, hid);
        if (act) {
            ++scans;
            if (!hit || t > dist * (1.0f - 1e-4f)) { S.lr[s] += C.x; S.lg[s] += C.y; S.lb[s] += C.z; }
        }
        claim_early();
        route(act ? (obj.material != 0 ? SQ_SURF_F : SQ_SURF_L) : -1, slot);
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
        const MatF &obj = S.scene.mats[id];
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
        for (int a = 0; a < sc.n_area; ++a) { // muestreoSA for every area light (misSamplingFunctions.h:105-118)
            if ((a & 1) == 0) ra = philox_block(pixel, sample, depth, 2 + (a >> 1), lp.key0, lp.key1);
            const float xi1 = u32_to_unit_f32((a & 1) ? ra.z : ra.x), xi2 = u32_to_unit_f32((a & 1) ? ra.w : ra.y);
            const int lid = sc.area[a];
            const MatF &sm = S.scene.mats[lid];
            const F3 cx = mk(sm.px, sm.py, sm.pz) - o;
            const float len2 = dot(cx, cx), inv_len = rsqrtf(len2);
            const float omc_max = one_minus_cos_max(sm.r * sm.r / len2);
            omc_last = omc_max;
            const F3 wi = cone_sample(cx * inv_len, omc_max, xi1, xi2);
            float t; int hid;
            const bool hit = scan_sm(S.scene, o, wi, t, hid);
            if (act) {
                ++scans;
                if ((hit ? hid : 0) == lid && !diel) { // id stays 0 on a miss, samplingFunctions.h:196
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
        const uint4 b1 = philox_block(pixel, sample, depth, 1, lp.key0, lp.key1);
        { // the BSDF-sampled term of MISv2 (:124-167): slots S_MIS = lanes 2,3 of block 1
            const float xi1 = u32_to_unit_f32(b1.z), xi2 = u32_to_unit_f32(b1.w);
            F3 wi_l, wh = mk(0, 0, 1);
            if (FACET) {
                wh = facet_normal(obj.alpha, xi1, xi2); wi_l = unit(fma3(wh, 2.0f * dot(wh, wo_l), -wo_l));
                if (diel) { // softDielectric (samplingFunctions.h:209-235): reflect with probability F, else the reference's refraction
                    if (sc.n_area > 0) {
                        const uint32_t slot = S_DIEL + (uint32_t)sc.n_area - 1u;
                        const float xg = u32_to_unit_f32(pick_lane(philox_block(pixel, sample, depth, slot >> 2, lp.key0, lp.key1), slot & 3u));
                        gpdf_loop = xg > di.F ? 1.0f - di.F : di.F;
                    }
                    refracted = !(xi1 < di.F);
                    wi_l = refracted ? di.wt : di.wr;
                }
            } else wi_l = cosine_local(xi1, xi2);
            const F3 wi = unit(to_world(fr, wi_l));
            float t; int hid;
            const bool hit = scan_sm(S.scene, o, wi, t, hid);
            if (act) {
                ++scans;
                if (hit && S.scene.mats[hid].emits) {
                    const MatF &em = S.scene.mats[hid];
                    const F3 cx = mk(em.px, em.py, em.pz) - o;
                    float omc = one_minus_cos_max(em.r * em.r / dot(cx, cx));
                    if (diel) L = L + had(dielectric_direct(em, o, wi_l.z, refracted, gpdf_loop), beta) * k.inv_cp;
                    else if (FACET) {
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
            const int pixel = item_pixel(item, pl);
            if (pixel >= 0) {
                float *out = hdr + (size_t)pixel * 3;
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

    // ---- one round: warp 0 snapshots the queues (lane = claim rank) and plans the generation --------------------------------------
    // Every other warp waits for this, so the dependent chain is kept short: ONE shared-memory load fetches all control words (lane i
    // reads word i), everything else is register shuffles, and every lane computes the few scalar decisions redundantly.
    __device__ __forceinline__ void plan_round(unsigned item_total) {
        const unsigned v = (&S.q_tail[0])[lane]; // words 0..19 are the control block, the padding behind it is never used
        const int item0 = (int)__shfl_sync(0xffffffffu, v, 14), item1 = (int)__shfl_sync(0xffffffffu, v, 15);
        unsigned cur0 = __shfl_sync(0xffffffffu, v, 16), cur1 = __shfl_sync(0xffffffffu, v, 17);
        const unsigned done0 = __shfl_sync(0xffffffffu, v, 18), done1 = __shfl_sync(0xffffffffu, v, 19);
        const unsigned n_free = __shfl_sync(0xffffffffu, v, 13) - __shfl_sync(0xffffffffu, v, 12);
        cur0 = min(cur0, item_total); cur1 = min(cur1, item_total); // (the tail fill may overshoot)
        int flush = -1, gen = -1;
        if (item1 >= 0) { if (cur1 == item_total) { if (done1 == item_total) flush = 1; } else gen = 1; }
        if (item0 >= 0) { if (cur0 == item_total) { if (done0 == item_total) flush = 0; } else if (gen < 0 || item0 < item1) gen = 0; }
        const unsigned gen_begin = gen == 0 ? cur0 : cur1;
        const unsigned left = gen >= 0 ? item_total - gen_begin : 0u;
        // while new samples keep coming only full 32-record batches are handed out (the remainder waits for the next round);
        // once generation has stopped (an item drains) everything goes
        const int q = (kRankStage >> (4 * min(lane, 6))) & 0xf; // lanes 0..5: the queue of that rank
        const unsigned tail = __shfl_sync(0xffffffffu, v, q & 7), handed = __shfl_sync(0xffffffffu, v, 6 + (q & 7));
        unsigned begin = 0, end = 0;
        if (lane < SQ_COUNT) {
            unsigned count = tail - handed;
            if (left != 0u) count &= ~31u;
            begin = handed; end = handed + count;
            S.q_end[q] = end;
        }
        unsigned queued = end - begin;
#pragma unroll
        for (int off = 1; off < 8; off <<= 1) queued += __shfl_xor_sync(0xffffffffu, queued, off);
        // generation: normally left to the tail fill -- warps that find the round's batches all claimed generate camera samples instead
        // of idling at the barrier -- and planned as batches of the round only when the queues cannot keep every warp busy
        unsigned n_gen = 0;
        if (queued < (unsigned)kSmThreads) {
            n_gen = min(n_free, left);
            if (n_gen < left) n_gen &= ~31u; // full warps only, except for the last samples of an item
        }
        if (lane == SQ_COUNT) { begin = gen_begin; end = gen_begin + n_gen; }
        const unsigned nb = (end - begin + 31u) >> 5;
        unsigned incl = nb;
#pragma unroll
        for (int off = 1; off < 8; off <<= 1) { const unsigned u = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += u; }
        if (lane < 8) { S.rb_first[lane] = incl - nb; S.rb_begin[lane] = begin; S.rb_end[lane] = end; }
        const unsigned total = __shfl_sync(0xffffffffu, incl, 7);
        if (lane == 0) {
            S.t_cursor[0] = cur0 + (gen == 0 ? n_gen : 0u); S.t_cursor[1] = cur1 + (gen == 1 ? n_gen : 0u);
            S.flush_slot = flush; S.gen_slot = gen; S.round_claim = (unsigned)(kSmThreads / 32); S.tail_budget = n_free - n_gen; S.tail_used = 0u;
            S.exit_flag = (total == 0u && flush < 0 && gen < 0) ? 1 : 0; // nothing queued, nothing to generate, nothing to write out
        }
    }

    __device__ __forceinline__ void run(float *__restrict__ hdr, int n_items) {
        const unsigned item_total = (unsigned)item_pixels * (unsigned)(lp.sample_end - lp.sample_begin);
        SMW_T(t_start);
        for (;;) {
            SMW_T(t0);
#ifdef VPT_SMWAVE_PROFILE
            if (lane == 0) S.dbg_arrive[tid >> 5] = t0;
#endif
            __syncthreads(); // (A) the previous round's pushes / releases / counters are visible
            SMW_T(t1);
            if (tid < 32) plan_round(item_total);
            __syncthreads(); // (B) the plan is visible
            SMW_T(t2);
#ifdef VPT_SMWAVE_PROFILE
            if (tid == 0) { // arrival spread at (A): last arrival minus mean arrival (x warps = idle warp-cycles), and last arrival -> (B) passed
                long long last = 0, sum = 0;
                for (int w = 0; w < kSmThreads / 32; ++w) { const long long a = S.dbg_arrive[w]; last = a > last ? a : last; sum += a; }
                prof[23] += (unsigned long long)(last * (kSmThreads / 32) - sum);
                prof[22] += (unsigned long long)((t2 - last) * (kSmThreads / 32));
            }
#endif
            SMW_ADD(16, t1 - t0); SMW_ADD(17, t2 - t1); SMW_ADD(20, tid == 0);
            if (S.exit_flag) break;
            if (S.flush_slot >= 0) flush_item(S.flush_slot, hdr, n_items); // its records are all finished; the round below only touches the other item
            SMW_T(t3);
            SMW_ADD(21, t3 - t2);
#ifdef VPT_SMWAVE_PROFILE
            long long in_batches = 0;
#define SMW_BATCH(q, call) { const long long b0 = clock64(); call; const long long b1 = clock64(); prof[q] += b1 - b0; prof[8 + q] += 1; in_batches += b1 - b0; }
#else
#define SMW_BATCH(q, call) { call; }
#endif
            const unsigned total = S.rb_first[7];
            const int gen_slot = S.gen_slot;
            // the first batch of a round needs no atomic: warp w takes batch w, the claim counter starts behind those (plan_round); the batch
            // table does not change during a round: read it once
            unsigned kb = (unsigned)tid >> 5;
            const uint4 f0 = *reinterpret_cast<const uint4 *>(&S.rb_first[0]), f1 = *reinterpret_cast<const uint4 *>(&S.rb_first[4]);
            while (kb < total) {
                const int rank = (kb >= f0.y) + (kb >= f0.z) + (kb >= f0.w) + (kb >= f1.x) + (kb >= f1.y) + (kb >= f1.z);
                const unsigned start = S.rb_begin[rank] + ((kb - S.rb_first[rank]) << 5);
                const int n = (int)min(32u, S.rb_end[rank] - start);
                const unsigned e = (start + (unsigned)lane) & (kSmPool - 1);
                switch (rank) {
                case 0: SMW_BATCH(SQ_SURF_F, stage_surf<true>(lane < n ? (int)S.queue[SQ_SURF_F][e] : -1)); break;
                case 1: SMW_BATCH(SQ_SURF_L, stage_surf<false>(lane < n ? (int)S.queue[SQ_SURF_L][e] : -1)); break;
                case 2: SMW_BATCH(SQ_PRIMARY, stage_primary(lane < n ? (int)S.queue[SQ_PRIMARY][e] : -1)); break;
                case 3: SMW_BATCH(SQ_MED_AREA, stage_med<false>(lane < n ? (int)S.queue[SQ_MED_AREA][e] : -1)); break;
                case 4: SMW_BATCH(SQ_MED_POINT, stage_med<true>(lane < n ? (int)S.queue[SQ_MED_POINT][e] : -1)); break;
                case 5: SMW_BATCH(SQ_SURF_P, stage_surf_p(lane < n ? (int)S.queue[SQ_SURF_P][e] : -1)); break;
                default: SMW_BATCH(6, stage_gen(gen_slot, start, n)); claim_early(); break;
                }
                // (claiming the next batch before running this one hides the atomic's latency but commits warps too early: measured slower)
                // (taking ALL batches round-robin without atomics: 6750 against 7340 -- the dynamic claim is what balances 12 000-cycle SURF_L batches
                // against 4 000-cycle ones)
                kb = __shfl_sync(0xffffffffu, next_raw, 0);
            }
            // tail fill: the round's batches are all claimed; instead of idling at the barrier generate camera samples from the budget
            // the plan left over (one free record per sample is guaranteed), they are consumed in the next round
            if (gen_slot >= 0) {
                const unsigned budget = S.tail_budget;
                for (;;) {
                    unsigned base = item_total;
                    if (lane == 0 && *(volatile unsigned *)&S.tail_used + 32u <= budget && smem_add(&S.tail_used, 32u, lz) + 32u <= budget)
                        base = smem_add(&S.t_cursor[gen_slot], 32u, lz);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (base >= item_total) break;
                    SMW_BATCH(7, stage_gen(gen_slot, base, (int)min(32u, item_total - base)));
                }
            }
#ifdef VPT_SMWAVE_PROFILE
            prof[18] += clock64() - t3 - in_batches;
#endif
        }
#ifdef VPT_SMWAVE_PROFILE
        prof[19] += clock64() - t_start;
#endif
    }
};

} // namespace f32
} // namespace vpt
