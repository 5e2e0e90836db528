// vpt_smsched.cuh -- the SM-wide wavefront SCHEDULER: one persistent CTA per SM, a pool of path records in shared memory, rounds of
// 32-record stage batches claimed by warps.  Precision-agnostic: the FP32 product pipeline (vpt_smwave.cuh) and the FP64 reference-mode
// pipeline (vpt_smwave_f64.cuh) plug their record layout and their stage functions into it (CRTP).
//
// Why (profiles/r1_summary.md): a per-warp wavefront fixed SIMT efficiency (28 of 32 lanes active) but used only 40 % of the issue slots --
// 42 % of all stall samples were `no_inst`: sixteen warps per SM, each in a different stage of a 62 KB kernel, thrash the 32 KB L1.5 /
// 6 KB L0 instruction caches.  Here the warps of an SM share ONE pool of path records and move through the stages together:
//   * one CTA per SM, grid = number of SMs (persistent); work items = groups of pixel tiles x all samples, handed out statically
//     (item j -> CTA j % gridDim.x), SLOTS = two or four items in flight per CTA so that draining items overlap the next ones;
//   * POOL path records in shared memory (layout: the pipeline's), one index queue (ring) per stage, one ring of free records.  A record
//     carries NO radiance: every vertex's direct light goes straight into the pixel's fixed-point sum;
//   * work proceeds in ROUNDS: at a barrier warp 0 snapshots every queue's tail (one load of the 20 control words, then shuffles) into a
//     table of 32-record batches ordered longest stage first; every warp claims batches with ONE shared-memory atomic each, runs them
//     and routes the survivors to the next stage's ring (match.any groups the lanes by destination: one atomic per group); what is
//     pushed during a round is consumed in the next one.  Only full batches are handed out while samples remain.  A warp that finds
//     the table used up generates new camera samples into the free records instead of idling at the barrier (tail fill).  The rank order
//     also keeps the warps inside two or three stages at a time -- a barrier-free variant lost 40 % to instruction-cache misses;
//   * queues are split by what diverges: medium vertex with a point / an area source, surface vertex needing the pLight shadow ray,
//     Lambert / microfacet surface vertex.
// Per-pixel sums are 64-bit fixed-point accumulators in shared memory, built from native 32-bit atomics: integer adds are order
// independent, so images are bit-reproducible although the order in which paths finish is data dependent.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include "vpt_internal.h"

namespace vpt {

#ifndef VPT_TAIL_GEN
#define VPT_TAIL_GEN 64 // camera samples per tail-fill batch: 64 = two per lane (run_gen_wide), 32 = one per lane
#endif
constexpr int kTailGen = VPT_TAIL_GEN;
constexpr int kSmMaxItemPixels = 128; // pixels per work item (power of two multiple of kTile)
// Work items in flight per CTA (template parameter SLOTS of the scheduler, chosen per launch): more of them keep the pool full at low sample
// counts, where an item's samples are generated in a few rounds while its last paths take tens of rounds to finish (64 spp: four slots
// +8 .. 10 % over two, six another +6 %; 16 spp: six +27 % over four); fewer cost less in the plan and per batch (1024 spp: two slots
// +1.2 % over four).  3 KB of pixel sums per slot.
#ifndef VPT_MAX_ITEM_SLOTS
#define VPT_MAX_ITEM_SLOTS 6
#endif
constexpr int kMaxItemSlots = VPT_MAX_ITEM_SLOTS; // (the plan reads 14 + 3 * SLOTS control words with one load of the warp: at most 6)
// Claim order of a round's batches (one nibble per rank, SQ_COUNT = generation), longest stage first so that a round ends evenly:
// SURF_F, SURF_L, PRIMARY, MED_AREA, MED_POINT, SURF_P, generation
constexpr unsigned kRankStage = 0x6312045u;

// meta word of a record: pixel-in-item (bits 0-6) | item slot (bits 7-9) | picked source (10-14) | hit object (15-19) | depth (20-31)
static_assert(kMaxSpheres <= 32 && VPT_MAX_DEPTH <= 4095 && kSmMaxItemPixels <= 128 && kSmMaxItemPixels >= kTile && kMaxItemSlots <= 8, "meta word layout");
__device__ __forceinline__ uint32_t meta_slot(uint32_t meta) { return (meta >> 7) & 7u; }
__device__ __forceinline__ uint32_t meta_pixel(uint32_t meta) { return meta & 0x7fu; }
__device__ __forceinline__ uint32_t meta_aux(int pixel_in_item, int item_slot) { return (uint32_t)pixel_in_item | ((uint32_t)item_slot << 7); }
__device__ __forceinline__ uint32_t meta_pack(uint32_t aux, uint32_t src, uint32_t hid, uint32_t depth) { return (aux & 0x3ffu) | (src << 10) | (hid << 15) | (depth << 20); }

// queues, pixel sums and control words of one CTA (the pipeline's shared-memory struct holds one, next to its records and its scene)
template <int POOL, int SLOTS>
struct SmCtl {
    static_assert(POOL % 32 == 0 && POOL <= 65536, "queues hold 16-bit record indices, batches are 32 records");
    static_assert(SLOTS >= 2 && SLOTS <= kMaxItemSlots, "item slots");
    uint16_t queue[SQ_COUNT][POOL];
    uint16_t freelist[POOL];
    unsigned long long acc[SLOTS][kSmMaxItemPixels][3];
    // this round's batch table by rank (kRankStage): batch k belongs to the last rank with rb_first[rank] <= k and covers entries
    // [rb_begin + 32 j, min(rb_end, +32)), j = k - rb_first
    __align__(16) unsigned rb_first[8]; // [7] = number of batches of the round
    unsigned rb_begin[8], rb_end[8];
    // ... and spelled out per batch, so that a claim costs ONE load: rank (bits 0-2) | records in the batch (3-8) | bits 9-20: ring index of the
    // batch's first queue entry or, for a generation batch (rank 6), its number j (samples [rb_begin[6] + 32 j, + 32)) | bits 21-31: the round
    // (mod 2048).  Written right after the plan by ALL warps (a few entries each: SmSched::run), so that the plan itself stays short.
    static constexpr int kDescMax = 2 * (POOL / 32) + 16;
    static_assert(POOL <= 4096, "batch descriptors keep a ring index in 12 bits");
    unsigned desc[kDescMax];
    unsigned round_no;
    unsigned round_claim;      // next unclaimed batch of the round: one atomicAdd per batch
    unsigned tail_limit;       // camera-sample cursor up to which warps out of batches may generate in this round (tail fill)
    // the 20 words the round plan reads, contiguous: warp 0 fetches them with ONE load (lane i reads word i, see plan_round)
    __align__(16) unsigned q_tail[SQ_COUNT]; // [0..5]   push counters
    unsigned q_end[SQ_COUNT];                // [6..11]  entries below it have been handed out
    unsigned free_head, free_tail;           // [12,13]  the free-record ring: allocate at the head (only below the round's snapshot), release at the tail
    int t_item[SLOTS];                       // [14 ..]  work item of the slot, -1: slot idle
    unsigned t_cursor[SLOTS], t_done[SLOTS]; // camera samples generated / paths finished
    unsigned ctl_pad[19 - 3 * SLOTS];
    int gen_slot, flush_slot, exit_flag;
#ifdef VPT_SMWAVE_PROFILE
    long long dbg_arrive[32];
#endif
    int next_item;
};

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

// The pipeline D derives from SmSched<D, POOL, THREADS, SLOTS> and provides
//   template <int STAGE> void run_stage(int slot)     one lane's record of a batch of stage STAGE (slot < 0: idle lane); loads the record, runs the
//                                                     stage, stores what changed, calls route() / count_done() -- every lane of the warp must call both
//   void run_gen(int item_slot, unsigned g0, int n)   n <= 32 new camera samples of the item in slot item_slot, lane i takes sample index g0 + i
//   void run_gen_wide(int item_slot, unsigned g0, int n)  the same for n <= kTailGen, lane i takes g0 + i, g0 + 32 + i, ... (the tail fill's batches)
template <class D, int POOL, int THREADS, int SLOTS>
struct SmSched {
    using Ctl = SmCtl<POOL, SLOTS>;
    Ctl &Q;
    const LaunchParams &lp;
    const int tid, lane;
    const unsigned lz; // lane * 0, opaque to the compiler (see smem_add)
    const int log_p, item_pixels, n_owned_tiles;
    unsigned next_raw = 0; // lane 0: the next batch of the round, claimed while the tail of the current one is still running (last_step)
#ifdef VPT_SMWAVE_PROFILE
    unsigned long long prof[24] = {};
#endif
    // Ring counters only grow between plans; every plan pulls a ring's two marks (pushed / handed out, free head / tail) back by the pool size
    // once the smaller one has passed it.  A ring never holds more than POOL entries and a round hands out at most POOL, so every counter a
    // round sees is below 2 * POOL and `counter mod POOL` is one compare and one subtract (POOL need not be a power of two).
    static __device__ __forceinline__ unsigned ring_index(unsigned counter) { return counter >= (unsigned)POOL ? counter - (unsigned)POOL : counter; }

    __device__ SmSched(Ctl &Q_, const LaunchParams &lp_, int log_p_, int n_owned_, int zero)
        : Q(Q_), lp(lp_), tid((int)threadIdx.x), lane((int)threadIdx.x & 31), lz((threadIdx.x & 31u) * (unsigned)zero), log_p(log_p_),
          item_pixels(1 << log_p_), n_owned_tiles(n_owned_) {}
    __device__ __forceinline__ D &self() { return *static_cast<D *>(this); }

    // cooperative initialisation by the whole block (then __syncthreads)
    __device__ __forceinline__ void init(int n_items) {
        for (int i = tid; i < POOL; i += THREADS) Q.freelist[i] = (uint16_t)i;
        for (int i = tid; i < SLOTS * kSmMaxItemPixels * 3; i += THREADS) (&Q.acc[0][0][0])[i] = 0ull;
        if (tid == 0) {
            for (int q = 0; q < SQ_COUNT; ++q) { Q.q_tail[q] = 0u; Q.q_end[q] = 0u; }
            Q.free_head = 0u; Q.free_tail = (unsigned)POOL;
            for (int b = 0; b < SLOTS; ++b) {
                const int item = (int)blockIdx.x + b * (int)gridDim.x;
                Q.t_item[b] = item < n_items ? item : -1; Q.t_cursor[b] = 0u; Q.t_done[b] = 0u;
            }
            Q.next_item = (int)blockIdx.x + SLOTS * (int)gridDim.x;
            Q.gen_slot = -1; Q.tail_limit = 0u; Q.round_no = 0u;
        }
        for (int i = tid; i < Ctl::kDescMax; i += THREADS) Q.desc[i] = 0xffffffffu; // (no round carries the tag 2047 before the 2047th)
    }

    // The claim for the NEXT batch is issued a few hundred cycles before the current one ends (at the start of its last step: the roulette's
    // Philox block or the final routing), so that the atomic's round trip is over when the claim loop needs it; early enough to hide the
    // latency, late enough not to commit a warp to work while others idle (claiming at the START of a batch measured 15 % slower).
    __device__ __forceinline__ void last_step() { if (lane == 0) next_raw = smem_add(&Q.round_claim, 1u, lz); }

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
    // the fixed-point sums of the pixel a record belongs to (meta: pixel-in-item, item slot)
    __device__ __forceinline__ unsigned long long *pixel_acc(uint32_t meta) { return Q.acc[meta_slot(meta)][meta_pixel(meta)]; }

    // ---- work items: item j = owned tiles [j * K, (j + 1) * K), K = item_pixels / kTile -----------------------------------------
    // (32-bit arithmetic: n_pixels is an int32, so tile and pixel indices fit; -1 = outside the image / not this rank's tile)
    __device__ __forceinline__ int item_pixel(int item, int pl) const {
        const unsigned owned = ((unsigned)item << (log_p - 7)) + ((unsigned)pl >> 7);
        if (owned >= (unsigned)n_owned_tiles) return -1;
        const unsigned pixel = (owned * (unsigned)lp.tile_count + (unsigned)lp.tile_rank) * (unsigned)kTile + ((unsigned)pl & (unsigned)(kTile - 1));
        return pixel < (unsigned)lp.n_pixels ? (int)pixel : -1;
    }
    __device__ __forceinline__ uint32_t pixel_of(uint32_t meta) const { return (uint32_t)item_pixel(Q.t_item[meta_slot(meta)], (int)meta_pixel(meta)); }

    // ---- queue / pool primitives (warp-aggregated shared-memory atomics) ------------------------------------------------------------
    // Route every lane's record in ONE step: dest = a stage queue (SQ_*), kDestFree (the record goes back to the free ring) or -1 (nothing).
    // Lanes with the same destination find each other with match.any; the lowest lane of each group reserves the group's entries
    // with one atomic (all group leaders in the same instruction), the others take their rank behind it.
    __device__ __forceinline__ void route(int dest, int slot) {
        const unsigned grp = __match_any_sync(0xffffffffu, dest);
        if (dest < 0) return;
        const int leader = __ffs(grp) - 1;
        unsigned base = 0;
        // counters: q_tail[0..5] are words 0..5 of the control block, free_tail is word 13; rings: queue[0..5] and, right behind them, freelist
        if (lane == leader) base = smem_add(&Q.q_tail[0] + (dest == kDestFree ? 13 : dest), (unsigned)__popc(grp), lz);
        base = __shfl_sync(grp, base, leader);
        (&Q.queue[0][0])[dest * POOL + (int)ring_index(base + __popc(grp & ((1u << lane) - 1u)))] = (uint16_t)slot;
    }
    __device__ __forceinline__ int alloc(bool flag) { // the round's snapshot guarantees enough free records below free_tail
        const unsigned m = __ballot_sync(0xffffffffu, flag);
        if (m == 0u) return -1;
        unsigned base = 0;
        if (lane == 0) base = smem_add(&Q.free_head, (unsigned)__popc(m), lz);
        base = __shfl_sync(0xffffffffu, base, 0);
        return flag ? (int)Q.freelist[ring_index(base + __popc(m & ((1u << lane) - 1u)))] : -1;
    }
    // K new records per lane (run_gen_wide): all taken from the free ring and pushed to queue q, one atomic each for the whole warp
    template <int K>
    __device__ __forceinline__ void alloc_push(int q, const bool *f, int *s) {
        unsigned m[K], n = 0;
#pragma unroll
        for (int h = 0; h < K; ++h) { m[h] = __ballot_sync(0xffffffffu, f[h]); n += (unsigned)__popc(m[h]); s[h] = -1; }
        if (n == 0u) return;
        unsigned from = 0, to = 0;
        if (lane == 0) { from = smem_add(&Q.free_head, n, lz); to = smem_add(&Q.q_tail[q], n, lz); }
        from = __shfl_sync(0xffffffffu, from, 0); to = __shfl_sync(0xffffffffu, to, 0);
        const unsigned below = (1u << lane) - 1u;
        unsigned before = 0;
#pragma unroll
        for (int h = 0; h < K; ++h) {
            const unsigned rank = before + (unsigned)__popc(m[h] & below);
            if (f[h]) { s[h] = (int)Q.freelist[ring_index(from + rank)]; Q.queue[q][ring_index(to + rank)] = (uint16_t)s[h]; }
            before += (unsigned)__popc(m[h]);
        }
    }
    __device__ __forceinline__ void count_done(bool ended, uint32_t meta) {
        if (SLOTS > 2 && !__any_sync(0xffffffffu, ended)) return;
#pragma unroll
        for (int b = 0; b < SLOTS; ++b) {
            const unsigned m = __ballot_sync(0xffffffffu, ended && meta_slot(meta) == (uint32_t)b);
            if (lane == 0 && m) smem_red(&Q.t_done[b], (unsigned)__popc(m), lz);
        }
    }
    // generation bookkeeping: samples of pixels outside the image and paths killed by the first roulette are finished already
    __device__ __forceinline__ void count_stillborn(int item_slot, bool mine, bool alive) {
        const unsigned m = __ballot_sync(0xffffffffu, mine && !alive);
        if (lane == 0 && m) smem_red(&Q.t_done[item_slot], (unsigned)__popc(m), lz);
    }

    // ---- item bookkeeping ---------------------------------------------------------------------------------------------------------------
    // every thread: write the finished item's pixels and clear its accumulators; thread 0: load the next item into the slot
    __device__ __forceinline__ void flush_item(int b, float *__restrict__ hdr, int n_items, double fix_inv) {
        const int item = Q.t_item[b];
        for (int pl = tid; pl < item_pixels; pl += THREADS) {
            const int pixel = item_pixel(item, pl);
            if (pixel >= 0) {
                float *out = hdr + (size_t)pixel * 3;
                for (int c = 0; c < 3; ++c) out[c] = (float)((double)(long long)Q.acc[b][pl][c] * fix_inv * lp.out_scale);
            }
            Q.acc[b][pl][0] = 0ull; Q.acc[b][pl][1] = 0ull; Q.acc[b][pl][2] = 0ull;
        }
        __syncthreads(); // everyone has read t_item[b]
        if (tid == 0) {
            const int next = Q.next_item;
            if (next < n_items) { Q.t_item[b] = next; Q.next_item = next + (int)gridDim.x; Q.t_cursor[b] = 0u; Q.t_done[b] = 0u; }
            else Q.t_item[b] = -1;
        }
    }

    // ---- one round: warp 0 snapshots the queues (lane = claim rank) and plans the generation --------------------------------------
    // Every other warp waits for this, so the dependent chain is kept short: ONE shared-memory load fetches all control words (lane i
    // reads word i), everything else is register shuffles, and every lane computes the few scalar decisions redundantly.
    __device__ __forceinline__ void plan_round(unsigned item_total) {
        static_assert(offsetof(Ctl, t_item) - offsetof(Ctl, q_tail) == 14 * sizeof(unsigned) &&
                          offsetof(Ctl, t_done) - offsetof(Ctl, q_tail) == (14 + 2 * SLOTS) * sizeof(unsigned) && 14 + 3 * SLOTS <= 32,
                      "plan_round reads the control words by index");
        static_assert(offsetof(Ctl, free_tail) - offsetof(Ctl, q_tail) == 13 * sizeof(unsigned) &&
                          offsetof(Ctl, freelist) - offsetof(Ctl, queue) == SQ_COUNT * POOL * sizeof(uint16_t),
                      "route() addresses the free ring as queue number SQ_COUNT");
        unsigned v = (&Q.q_tail[0])[lane]; // words 0 .. 13 + 3 * SLOTS are the control block, the padding behind it is never used
        { // keep the ring counters below 2 * POOL: lanes 0..5 / 6..11 hold a queue's pushed / handed-out marks, 12 / 13 the free ring's head / tail
            const unsigned low = __shfl_sync(0xffffffffu, v, lane < 6 ? lane + 6 : (lane == 13 ? 12 : lane)); // the smaller mark of the pair
            if (lane < 14 && low >= (unsigned)POOL) { v -= (unsigned)POOL; (&Q.q_tail[0])[lane] = v; }
        }
        const int prev_gen = Q.gen_slot;
        const unsigned prev_limit = Q.tail_limit; // the tail fill overshoots its limit by the claims that found nothing
        const unsigned n_free = __shfl_sync(0xffffffffu, v, 13) - __shfl_sync(0xffffffffu, v, 12);
        // the slot to write out (all samples generated, all paths finished; the lowest such slot, one per round) and the slot to generate
        // from (the oldest item that still has samples)
        unsigned cur[SLOTS];
        int flush = -1, gen = -1, gen_item = 0x7fffffff;
        unsigned gen_begin = 0;
#pragma unroll
        for (int b = SLOTS - 1; b >= 0; --b) {
            const int item = (int)__shfl_sync(0xffffffffu, v, 14 + b);
            cur[b] = __shfl_sync(0xffffffffu, v, 14 + SLOTS + b);
            const unsigned done = __shfl_sync(0xffffffffu, v, 14 + 2 * SLOTS + b);
            if (prev_gen == b) cur[b] = min(cur[b], prev_limit);
            if (item >= 0) {
                if (cur[b] == item_total) { if (done == item_total) flush = b; }
                else if (item <= gen_item) { gen = b; gen_item = item; gen_begin = cur[b]; }
            }
        }
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
            Q.q_end[q] = end;
        }
        unsigned queued = end - begin;
#pragma unroll
        for (int off = 1; off < 8; off <<= 1) queued += __shfl_xor_sync(0xffffffffu, queued, off);
        // generation: normally left to the tail fill -- warps that find the round's batches all claimed generate camera samples instead
        // of idling at the barrier -- and planned as batches of the round only when the queues cannot keep every warp busy
        unsigned n_gen = 0;
        if (queued < (unsigned)THREADS) {
            n_gen = min(n_free, left);
            if (n_gen < left) n_gen &= ~31u; // full warps only, except for the last samples of an item
        }
        if (lane == SQ_COUNT) { begin = gen_begin; end = gen_begin + n_gen; }
        const unsigned nb = (end - begin + 31u) >> 5;
        unsigned incl = nb;
#pragma unroll
        for (int off = 1; off < 8; off <<= 1) { const unsigned u = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += u; }
        const unsigned first = incl - nb;
        if (lane < 8) { Q.rb_first[lane] = first; Q.rb_begin[lane] = begin; Q.rb_end[lane] = end; }
        const unsigned total = __shfl_sync(0xffffffffu, incl, 7);
        if (lane == 0) {
            const unsigned after = gen_begin + n_gen;                       // the cursor after the planned generation batches
            const unsigned budget = (n_free - n_gen) & ~31u;                // free records the tail fill may use: one per sample, whole batches
            const unsigned limit = min(after + budget, item_total);         // (a last, partial batch of the item needs fewer records than it claims)
#pragma unroll
            for (int b = 0; b < SLOTS; ++b) Q.t_cursor[b] = gen == b ? after : cur[b];
            Q.tail_limit = gen >= 0 ? limit : 0u;
            Q.flush_slot = flush; Q.gen_slot = gen; Q.round_claim = (unsigned)(THREADS / 32); Q.round_no = Q.round_no + 1u;
            Q.exit_flag = (total == 0u && flush < 0 && gen < 0) ? 1 : 0; // nothing queued, nothing to generate, nothing to write out
        }
    }

    __device__ __forceinline__ void run(float *__restrict__ hdr, int n_items, double fix_inv) {
        const unsigned item_total = (unsigned)item_pixels * (unsigned)(lp.sample_end - lp.sample_begin);
        SMW_T(t_start);
        for (;;) {
            SMW_T(t0);
#ifdef VPT_SMWAVE_PROFILE
            if (lane == 0) Q.dbg_arrive[tid >> 5] = t0;
#endif
            __syncthreads(); // (A) the previous round's pushes / releases / counters are visible
            SMW_T(t1);
            if (tid < 32) plan_round(item_total);
            __syncthreads(); // (B) the plan is visible
            SMW_T(t2);
#ifdef VPT_SMWAVE_PROFILE
            if (tid == 0) { // arrival spread at (A): last arrival minus mean arrival (x warps = idle warp-cycles), and last arrival -> (B) passed
                long long last = 0, sum = 0;
                for (int w = 0; w < THREADS / 32; ++w) { const long long a = Q.dbg_arrive[w]; last = a > last ? a : last; sum += a; }
                prof[23] += (unsigned long long)(last * (THREADS / 32) - sum);
                prof[22] += (unsigned long long)((t2 - last) * (THREADS / 32));
            }
#endif
            SMW_ADD(16, t1 - t0); SMW_ADD(17, t2 - t1); SMW_ADD(20, tid == 0);
            if (Q.exit_flag) break;
            if (Q.flush_slot >= 0) flush_item(Q.flush_slot, hdr, n_items, fix_inv); // its records are all finished; the round below only touches the other item
            SMW_T(t3);
            SMW_ADD(21, t3 - t2);
#ifdef VPT_SMWAVE_PROFILE
            long long in_batches = 0;
#define SMW_BATCH(q, call) { const long long b0 = clock64(); call; const long long b1 = clock64(); prof[q] += b1 - b0; prof[8 + q] += 1; in_batches += b1 - b0; }
#else
#define SMW_BATCH(q, call) { call; }
#endif
            const unsigned total = Q.rb_first[7];
            const int gen_slot = Q.gen_slot;
            const unsigned gen_begin = Q.rb_begin[SQ_COUNT];
            const unsigned tag = (Q.round_no & 0x7ffu) << 21;
            // the batch table (rb_first / rb_begin / rb_end by rank) -> one descriptor per batch
            const uint4 f0 = *reinterpret_cast<const uint4 *>(&Q.rb_first[0]), f1 = *reinterpret_cast<const uint4 *>(&Q.rb_first[4]);
            auto make_desc = [&](unsigned k) -> unsigned {
                const int rank = (k >= f0.y) + (k >= f0.z) + (k >= f0.w) + (k >= f1.x) + (k >= f1.y) + (k >= f1.z);
                const unsigned j = k - Q.rb_first[rank], start = Q.rb_begin[rank] + (j << 5);
                const unsigned n = min(32u, Q.rb_end[rank] - start);
                return (unsigned)rank | (n << 3) | ((rank == SQ_COUNT ? j : ring_index(start)) << 9) | tag;
            };
            { // every warp spells out its share of the table (the plan, on which all warps wait, stays short); a reader checks the round tag
                constexpr int kPer = (Ctl::kDescMax + THREADS / 32 - 1) / (THREADS / 32);
                const unsigned k = (unsigned)(tid >> 5) * (unsigned)kPer + (unsigned)lane;
                if (lane < kPer && k < total) Q.desc[k] = make_desc(k);
            }
            // the first batch of a round needs no atomic and no table: warp w takes batch w, the claim counter starts behind those (plan_round)
            unsigned kb = (unsigned)tid >> 5;
            unsigned dsc = kb < total ? make_desc(kb) : 0u;
            while (kb < total) {
                const int rank = (int)(dsc & 7u), n = (int)((dsc >> 3) & 63u);
                const unsigned idx = (dsc >> 9) & 0xfffu;
                const unsigned start = gen_begin + (idx << 5);             // (generation batches only)
                const unsigned e = ring_index(idx + (unsigned)lane);       // this lane's queue entry
                switch (rank) {
                case 0: SMW_BATCH(SQ_SURF_F, self().template run_stage<SQ_SURF_F>(lane < n ? (int)Q.queue[SQ_SURF_F][e] : -1)); break;
                case 1: SMW_BATCH(SQ_SURF_L, self().template run_stage<SQ_SURF_L>(lane < n ? (int)Q.queue[SQ_SURF_L][e] : -1)); break;
                case 2: SMW_BATCH(SQ_PRIMARY, self().template run_stage<SQ_PRIMARY>(lane < n ? (int)Q.queue[SQ_PRIMARY][e] : -1)); break;
                case 3: SMW_BATCH(SQ_MED_AREA, self().template run_stage<SQ_MED_AREA>(lane < n ? (int)Q.queue[SQ_MED_AREA][e] : -1)); break;
                case 4: SMW_BATCH(SQ_MED_POINT, self().template run_stage<SQ_MED_POINT>(lane < n ? (int)Q.queue[SQ_MED_POINT][e] : -1)); break;
                case 5: SMW_BATCH(SQ_SURF_P, self().template run_stage<SQ_SURF_P>(lane < n ? (int)Q.queue[SQ_SURF_P][e] : -1)); break;
                default: SMW_BATCH(6, self().run_gen(gen_slot, start, n)); last_step(); break;
                }
                // (claiming the next batch before running this one hides the atomic's latency but commits warps too early: measured slower)
                // (taking ALL batches round-robin without atomics: 6750 against 7340 -- the dynamic claim is what balances 12 000-cycle SURF_L batches
                // against 4 000-cycle ones)
                kb = __shfl_sync(0xffffffffu, next_raw, 0);
                if (kb < total) do { dsc = *(volatile unsigned *)&Q.desc[kb]; } while ((dsc ^ tag) >> 21); // (written thousands of cycles ago: never loops)
            }
            // tail fill: the round's batches are all claimed; instead of idling at the barrier generate camera samples up to the limit the
            // plan set (one free record per sample is guaranteed), they are consumed in the next round.  One atomic per batch.
            if (gen_slot >= 0) {
                const unsigned limit = Q.tail_limit;
                for (;;) {
                    unsigned base = limit;
                    if (lane == 0 && *(volatile unsigned *)&Q.t_cursor[gen_slot] < limit) base = smem_add(&Q.t_cursor[gen_slot], (unsigned)kTailGen, lz);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (base >= limit) break;
                    // (the last samples of a round in 32-sample batches, for a finer grain where the round ends: 8655 against 8686 -- the one-per-lane
                    // form is slower per sample than the gain in arrival spread)
                    if (kTailGen > 32) { SMW_BATCH(7, self().run_gen_wide(gen_slot, base, (int)min((unsigned)kTailGen, limit - base))); }
                    else { SMW_BATCH(7, self().run_gen(gen_slot, base, (int)min(32u, limit - base))); }
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

} // namespace vpt
