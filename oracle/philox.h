/* oracle/philox.h -- TEST INFRASTRUCTURE.  Philox4x32-10 (Salmon et al., "Parallel random numbers: as
 * easy as 1, 2, 3", SC'11), written out independently of the device implementation in
 * minimal_volumetric_path_tracer_b200/csrc/vpt_philox.cuh so that the two check each other
 * (known-answer vectors of the Random123 distribution are in tests/test_oracle_units.py).
 *
 * Stream convention shared with the product (DESIGN.md "RNG"):
 *   key     = (seed & 0xffffffff, seed >> 32)
 *   counter = (pixel, sample, bounce, block)     block = slot / 4, lane = slot % 4
 *   slots of one bounce (a fixed slot per PURPOSE, so that a GPU kernel may generate blocks out of consumption order):
 *       0 roulette   1 light pick   2 distance (free-flight xi or equi-angular xi)   3 equi-angular surface/medium decision
 *       medium vertex : 4,5 NEE cone sample      6,7 phase-function sample
 *       surface vertex: 4,5 BSDF sample (next direction)   6,7 BSDF-sampled direct light (MISv2)   8+2a, 9+2a cone sample of area light a
 *                       dielectric (material 2): 4 and 6 only (one draw each), 40+a reflect-or-refract pdf choice per area light a
 *   pixel jitter (rt.cpp:787): slots 0,1 of the pseudo-bounce 0xffffffff
 *   uniform = (2 * (word >> 9) + 1) * 2^-24      in (0, 1): never 0 or 1, exactly representable in fp32 and fp64
 */
#ifndef VPT_ORACLE_PHILOX_H
#define VPT_ORACLE_PHILOX_H
#include <stdint.h>

static inline void vpt_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; ++round) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline double vpt_u32_to_unit(uint32_t w) { return (double)(2u * (w >> 9) + 1u) * (1.0 / 16777216.0); }

#endif
