// vpt_philox.cuh -- counter-based random stream of the renderer (replaces the reference's shared erand48 state,
// Vector.h:38 / Vector.cpp:8 / rt.cpp:746).
//
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11).  Stream convention (DESIGN.md "RNG"):
//   key     = (seed & 0xffffffff, seed >> 32)
//   counter = (pixel, sample, bounce, block)       block = slot / 4, lane = slot % 4
//   uniform = (2 * (word >> 9) + 1) * 2^-24        in (0,1), never 0 or 1 (the reference's 48-bit erand48 practically never
//                                                  returns 0 either); exact in fp32 and fp64 -> both precisions see the same numbers
// One fixed slot per PURPOSE inside a bounce, so a kernel may generate blocks where it needs them instead of in the
// reference's consumption order (the FP64 CPU oracle uses the same table, oracle/philox.h):
//   0 roulette   1 light pick   2 distance (free-flight xi or equi-angular xi)   3 equi-angular surface/medium decision
//   medium vertex : 4,5 NEE cone sample      6,7 phase-function sample
//   surface vertex: 4,5 BSDF sample (next direction)   6,7 BSDF-sampled direct light (MISv2)   8+2a, 9+2a cone sample of area light a
//                   dielectric (material 2): 4 and 6 only (one draw each), 40+a reflect-or-refract pdf choice per area light a
//   pixel jitter (rt.cpp:787): slots 0,1 of the pseudo-bounce 0xffffffff
#pragma once
#include <stdint.h>

namespace vpt {

enum : uint32_t { S_RR = 0, S_SRC = 1, S_DIST = 2, S_DECIDE = 3, S_NEE = 4, S_PHASE = 6, S_BSDF = 4, S_MIS = 6, S_AREA = 8,
                  S_DIEL = 40 }; // + a: a dielectric surface's extra draw per area light a (misSamplingFunctions.h:116); S_AREA + 2a stays below 40
constexpr uint32_t kJitterBounce = 0xffffffffu;

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}
// out-of-line copy: the megakernels call this from several places and must stay small enough for the instruction cache
static __device__ __noinline__ uint4 philox_block(uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t block, uint32_t k0, uint32_t k1) {
    return philox4x32_10(make_uint4(pixel, sample, bounce, block), k0, k1);
}
// The same block with the ten round keys (key + round * Weyl constant) read from constant memory: `hi ^ c ^ c[bank][imm]` is ONE LOP3 with a
// constant-bank operand, where the version above spends two more instructions per round bumping the keys (UIADD3: 3.4 % of the product
// kernel's executed instructions, profiles/r2_summary.md).  The host sets the schedule of the render's seed before the launch
// (vpt_kernels_f32.cu philox_keys_begin); used by the product kernel only.
static __constant__ uint32_t c_philox_ks[20];
#ifdef VPT_PHILOX_INLINE // experiment (tools/build_variant.py)
static __device__ __forceinline__ uint4 philox_block_ck(
#else
static __device__ __noinline__ uint4 philox_block_ck(
#endif
    uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t block) {
    uint4 c = make_uint4(pixel, sample, bounce, block);
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ c_philox_ks[2 * round], lo1, hi0 ^ c.w ^ c_philox_ks[2 * round + 1], lo0);
    }
    return c;
}
// (2k + 1) 2^-24 with k = w >> 9, without the integer-to-float conversion: 1.k (k in the mantissa) minus (1 - 2^-24) is exactly that value
__device__ __forceinline__ float u32_to_unit_f32(uint32_t w) { return __uint_as_float(0x3f800000u | (w >> 9)) - 0.99999994039535522461f; }
__device__ __forceinline__ double u32_to_unit_f64(uint32_t w) { return (double)(2u * (w >> 9) + 1u) * 5.9604644775390625e-8; }
__device__ __forceinline__ uint32_t pick_lane(const uint4 &b, uint32_t lane) { return lane == 0 ? b.x : lane == 1 ? b.y : lane == 2 ? b.z : b.w; }

struct Rng {
    uint32_t pixel, sample, bounce, k0, k1;
    uint32_t loaded; // block currently held in buf (0xffffffff: none)
    uint4 buf;

    __device__ __forceinline__ void start(uint32_t pixel_, uint32_t sample_, uint32_t key0, uint32_t key1) {
        pixel = pixel_; sample = sample_; bounce = 0; k0 = key0; k1 = key1; loaded = 0xffffffffu;
    }
    __device__ __forceinline__ void begin_bounce(uint32_t b) { bounce = b; loaded = 0xffffffffu; }
    __device__ __forceinline__ uint32_t word(uint32_t slot) {
        const uint32_t blk = slot >> 2;
        if (blk != loaded) { buf = philox_block(pixel, sample, bounce, blk, k0, k1); loaded = blk; }
        return pick_lane(buf, slot & 3u);
    }
    __device__ __forceinline__ float next_f32(uint32_t slot) { return u32_to_unit_f32(word(slot)); }
    __device__ __forceinline__ double next_f64(uint32_t slot) { return u32_to_unit_f64(word(slot)); }
    __device__ __forceinline__ void jitter_f32(float &a, float &b) {
        const uint4 r = philox_block(pixel, sample, kJitterBounce, 0, k0, k1);
        a = u32_to_unit_f32(r.x); b = u32_to_unit_f32(r.y);
    }
    __device__ __forceinline__ void jitter_f64(double &a, double &b) {
        const uint4 r = philox_block(pixel, sample, kJitterBounce, 0, k0, k1);
        a = u32_to_unit_f64(r.x); b = u32_to_unit_f64(r.y);
    }
};

} // namespace vpt
