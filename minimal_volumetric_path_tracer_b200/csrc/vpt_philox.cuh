// vpt_philox.cuh -- counter-based random stream of the renderer (replaces the reference's shared erand48 state,
// Vector.h:38 / Vector.cpp:8 / rt.cpp:746).
//
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11).  Stream convention (DESIGN.md "RNG"):
//   key     = (seed & 0xffffffff, seed >> 32)
//   counter = (pixel, sample, bounce, block)       block = draw_index / 4, lane = draw_index % 4
//   uniform = (2 * (word >> 9) + 1) * 2^-24        in (0,1), never 0 or 1 (the reference's 48-bit erand48 practically never
//                                                  returns 0 either); exact in fp32 and fp64 -> both precisions see the same numbers
// Draw order inside a bounce is the reference's consumption order (SURVEY.md section 8a pseudo-code).  Bounce 0 starts
// with the two pixel-jitter draws of rt.cpp:787.
#pragma once
#include <stdint.h>

namespace vpt {

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

struct Rng {
    uint32_t pixel, sample, bounce, k0, k1;
    uint32_t idx;    // next draw index inside this bounce
    uint32_t loaded; // block currently held in buf (0xffffffff: none)
    uint4 buf;

    __device__ __forceinline__ void start(uint32_t pixel_, uint32_t sample_, uint32_t key0, uint32_t key1) {
        pixel = pixel_; sample = sample_; bounce = 0; k0 = key0; k1 = key1; idx = 0; loaded = 0xffffffffu;
    }
    __device__ __forceinline__ void begin_bounce(uint32_t b) {
        if (b == 0) return; // bounce 0 continues after the jitter draws
        bounce = b; idx = 0; loaded = 0xffffffffu;
    }
    __device__ __forceinline__ void skip(uint32_t n) { idx += n; }
    __device__ __forceinline__ uint32_t next_word() {
        const uint32_t blk = idx >> 2;
        if (blk != loaded) { buf = philox4x32_10(make_uint4(pixel, sample, bounce, blk), k0, k1); loaded = blk; }
        const uint32_t lane = idx & 3u;
        ++idx;
        return lane == 0 ? buf.x : lane == 1 ? buf.y : lane == 2 ? buf.z : buf.w;
    }
    __device__ __forceinline__ float next_f32() { return (float)(2u * (next_word() >> 9) + 1u) * 5.9604644775390625e-8f; }
    __device__ __forceinline__ double next_f64() { return (double)(2u * (next_word() >> 9) + 1u) * 5.9604644775390625e-8; }
};

} // namespace vpt
