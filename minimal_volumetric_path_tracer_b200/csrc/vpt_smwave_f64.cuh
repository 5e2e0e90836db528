// vpt_smwave_f64.cuh -- the FP64 reference-mode pipeline of the SM-wide wavefront (VPT_PRECISION_FP64_REF with kernel AUTO): the only mode whose
// image equals src/rt.cpp as shipped (both rounding-decided quirks available) on the scheduler of the product kernel (vpt_smsched.cuh).
//
// The one-thread-per-pixel FP64 kernel ran with about 10 of 32 lanes active and 378 KB of SASS (every double-precision transcendental inlined
// at every divergent call site).  Here a batch of 32 records runs ONE part of a vertex (vpt_f64.cuh: vertex_primary, vertex_medium,
// vertex_surface -- the very functions the sequential kernels and the FP64 unit kernels compose into vertex(), so the arithmetic, its order
// and its roundings are the reference-faithful ones) with all lanes in the same code:
//   PRIMARY   scan, light pick, distance sampling, surface-or-medium decision                     -> MED_POINT / MED_AREA / SURF_L / SURF_F / end
//   MED_*     (free)SingleScattering (incl. the r = 0 fall-through of VPT_QUIRK_R0_FALLTHROUGH), phase sample, roulette   -> PRIMARY / end
//   SURF_*    pLight, MISv2, bdsf, roulette (Lambert / microfacet-or-dielectric batches)          -> PRIMARY / end
// (the FP32 pipeline's separate pLight stage is not used: the reference evaluates pLight for every surface vertex, vptShadeMethods.h:1316).
// Records: 80 B (origin, direction, throughput in double -- a medium vertex parks its transmittance and distance pdf in the direction
// slots, which it does not need --, sample, packed word); random numbers are regenerated from the Philox counter in every part.
// Radiance goes to 2^-34 fixed-point pixel sums (order independent: reruns are bit-identical).
#pragma once
#include "vpt_smsched.cuh"
#include "vpt_f64.cuh"

namespace vpt {
namespace f64 {

#ifndef VPT_SMD_THREADS
#define VPT_SMD_THREADS 512
#endif
constexpr int kSmdThreads = VPT_SMD_THREADS;
constexpr int kSmdPool = 2048;
#ifndef VPT_SMD_SLOTS
#define VPT_SMD_SLOTS 4 // (six: 16 spp 715 against 577 Mpaths/s, but 64 spp 1082 against 1097 and 256 spp 1103 against 1126)
#endif
constexpr int kSmdSlots = VPT_SMD_SLOTS; // work items in flight (vpt_smsched.cuh): reference mode is used at low sample counts
constexpr double kSmdFixScale = 17179869184.0;       // 2^34
constexpr double kSmdFixInv = 1.0 / 17179869184.0;
constexpr double kSmdMaxContribution = 268435456.0;  // 2^28: a contribution at or above it (or NaN) is dropped and counted (vpt_stats.nonfinite)

struct SmSharedD {
    SphereD spheres[kMaxSpheres];
    Ctx ctx; // ONE copy per CTA: the out-of-line building blocks take it by reference, and a per-thread copy would live in local memory
    double ox[kSmdPool], oy[kSmdPool], oz[kSmdPool];
    double dx[kSmdPool], dy[kSmdPool], dz[kSmdPool]; // direction; medium vertex between PRIMARY and MED: dx = transmittance, dy = distance pdf
    double br[kSmdPool], bg[kSmdPool], bb[kSmdPool];
    uint32_t sample[kSmdPool];
    uint32_t meta[kSmdPool];
    SmCtl<kSmdPool, kSmdSlots> ctl;
};
static_assert(sizeof(SmSharedD) <= 232448, "one CTA per SM: at most 227 KB of shared memory");

struct SmWaveD : SmSched<SmWaveD, kSmdPool, kSmdThreads, kSmdSlots> {
    using Base = SmSched<SmWaveD, kSmdPool, kSmdThreads, kSmdSlots>;
    SmSharedD &M;
    const Ctx &c;
    Tally tally{0u, 0u};
    unsigned nonfinite = 0, paths = 0;

    __device__ SmWaveD(SmSharedD &M_, const Ctx &c_, const LaunchParams &lp_, int log_p_, int n_owned_, int zero)
        : Base(M_.ctl, lp_, log_p_, n_owned_, zero), M(M_), c(c_) {}

    __device__ __forceinline__ void add(uint32_t meta, D3 L) {
        if (!(fabs(L.x) < kSmdMaxContribution && fabs(L.y) < kSmdMaxContribution && fabs(L.z) < kSmdMaxContribution)) { ++nonfinite; return; }
        unsigned long long *a = pixel_acc(meta);
        if (L.x != 0.0) add_fixed(a + 0, __double2ll_rn(L.x * kSmdFixScale));
        if (L.y != 0.0) add_fixed(a + 1, __double2ll_rn(L.y * kSmdFixScale));
        if (L.z != 0.0) add_fixed(a + 2, __double2ll_rn(L.z * kSmdFixScale));
    }
    // roulette of the next bounce (vptShadeMethods.h:1282): the path goes on to PRIMARY or ends
    __device__ __forceinline__ bool survives(Rng &rng, uint32_t depth) {
        rng.begin_bounce(depth);
        if ((c.max_depth > 0 && (int)depth >= c.max_depth) || depth >= (uint32_t)VPT_MAX_DEPTH) return false;
        return !(rng.next_f64(S_RR) < c.q);
    }

    // rt.cpp:787
    __device__ __forceinline__ D3 camera_dir(int x, int y, double xi1, double xi2) const {
        const D3 v = v3(lp.cam_cx) * ((static_cast<double>(x) + xi1 - 0.5) / lp.width - .5) + v3(lp.cam_cy) * ((static_cast<double>(y) + xi2 - 0.5) / lp.height - .5) + v3(lp.cam_d);
        return unit(v);
    }

    __device__ __forceinline__ void run_gen(int b, unsigned g0, int n) {
        const bool mine = lane < n;
        const unsigned g = g0 + (unsigned)lane;
        const int pl = (int)(g & (unsigned)(item_pixels - 1));
        const uint32_t sample = (uint32_t)lp.sample_begin + (g >> log_p);
        const int pixel = mine ? item_pixel(Q.t_item[b], pl) : -1;
        bool alive = false;
        Rng rng;
        if (pixel >= 0) {
            ++paths;
            rng.start((uint32_t)pixel, sample, lp.key0, lp.key1);
            alive = survives(rng, 0u);
        }
        const int slot = alloc(alive);
        if (alive) {
            double j1, j2;
            rng.jitter_f64(j1, j2);
            const int row = (int)((unsigned)pixel / (unsigned)lp.width), col = pixel - row * lp.width;
            const D3 d = camera_dir(col, lp.height - 1 - row, j1, j2);
            M.ox[slot] = lp.cam_o[0]; M.oy[slot] = lp.cam_o[1]; M.oz[slot] = lp.cam_o[2];
            M.dx[slot] = d.x; M.dy[slot] = d.y; M.dz[slot] = d.z;
            M.br[slot] = 1.0; M.bg[slot] = 1.0; M.bb[slot] = 1.0;
            M.sample[slot] = sample;
            M.meta[slot] = meta_pack(meta_aux(pl, b), 0u, 0u, 0u);
        }
        route(alive ? SQ_PRIMARY : -1, slot);
        count_stillborn(b, mine, alive);
    }

    __device__ __forceinline__ void run_gen_wide(int b, unsigned g0, int n) { // (one sample per lane at a time in FP64: registers)
        for (int h = 0; h < n; h += 32) run_gen(b, g0 + (unsigned)h, min(n - h, 32));
    }

    template <int STAGE>
    __device__ __forceinline__ void run_stage(int slot) {
        if (STAGE == SQ_SURF_P) { last_step(); route(-1, slot); return; } // never queued in this pipeline
        const bool act = slot >= 0;
        const int s = act ? slot : 0;
        const uint32_t meta = M.meta[s];
        const uint32_t depth = meta >> 20;
        Path p;
        p.o = mk(M.ox[s], M.oy[s], M.oz[s]); p.d = mk(M.dx[s], M.dy[s], M.dz[s]); p.beta = mk(M.br[s], M.bg[s], M.bb[s]);
        p.L = mk(0, 0, 0); p.depth = (int)depth;
        Rng rng;
        rng.start(pixel_of(meta), M.sample[s], lp.key0, lp.key1);
        rng.begin_bounce(depth);
        if (STAGE == SQ_PRIMARY) {
            int dest = -1;
            if (act) {
                VertexPlan vp;
                D3 Lc;
                const int kind = vertex_primary_inl(c, p, rng, tally, vp, Lc);
                if (kind == V_END) { add(meta, Lc); dest = kDestFree; }
                else {
                    M.ox[s] = vp.x.x; M.oy[s] = vp.x.y; M.oz[s] = vp.x.z;
                    M.meta[s] = meta_pack(meta, (uint32_t)vp.source, (uint32_t)vp.id, depth);
                    if (kind == V_MEDIUM) { M.dx[s] = vp.T; M.dy[s] = vp.pdf_medium; dest = c.s[vp.source].r == 0 ? SQ_MED_POINT : SQ_MED_AREA; }
                    else dest = c.s[vp.id].material != 0 ? SQ_SURF_F : SQ_SURF_L;
                }
            }
            last_step();
            route(dest, slot);
            count_done(dest == kDestFree, meta);
        } else {
            int dest = -1;
            if (act) {
                VertexPlan vp;
                vp.source = (int)((meta >> 10) & 31u); vp.id = (int)((meta >> 15) & 31u); vp.x = p.o;
                D3 Lc;
                if (STAGE == SQ_MED_POINT || STAGE == SQ_MED_AREA) { vp.T = p.d.x; vp.pdf_medium = p.d.y; vertex_medium_inl(c, p, vp, rng, tally, Lc); }
                else { vp.T = 0; vp.pdf_medium = 1; vertex_surface_inl(c, p, vp, rng, tally, Lc); }
                add(meta, Lc);
            }
            last_step();
            if (act) {
                if (survives(rng, depth + 1u)) {
                    M.ox[s] = p.o.x; M.oy[s] = p.o.y; M.oz[s] = p.o.z;
                    M.dx[s] = p.d.x; M.dy[s] = p.d.y; M.dz[s] = p.d.z;
                    M.br[s] = p.beta.x; M.bg[s] = p.beta.y; M.bb[s] = p.beta.z;
                    M.meta[s] = meta_pack(meta, 0u, 0u, depth + 1u);
                    dest = SQ_PRIMARY;
                } else dest = kDestFree;
            }
            route(dest, slot);
            count_done(dest == kDestFree, meta);
        }
    }
};

} // namespace f64
} // namespace vpt
