// vpt_scan.cuh -- the scene as every FP32 kernel reads it (shared memory) and THE all-sphere scan of the FP32 path.
//
// intersect (pathTracingUtilities.h:12-36) over Sphere::intersect (Sphere.h:27-37), FP32 semantics (include/vpt.h): r == 0 spheres have no
// scan record.  One out-of-line copy (scan_sm_call) serves the product wavefront, the HBM wavefront, the megakernel, the ray marcher and
// the unit kernels; the per-pair root arithmetic lives in pair_general_* / pair_direct_*, which VPT_UNIT_SPHERE_INTERSECT evaluates directly.
//   * scan records are staged in shared memory as float4 and read with broadcast LDS.128;
//   * ordinary spheres (r < 64) take the roots directly as -b -+ sqrt(det) with det = r^2 - |op - (op.d)d|^2 (no cancellation for small
//     far-away spheres); spheres with r >= 64 (the r = 1e5 walls of Sphere.cpp:11-15 above all) use the re-anchored cancellation-free
//     form: the host picks an anchor q on the sphere nearest the scene and m = q - p, then c = |o-q|^2 + 2 (o-q).m + c0 and the far root
//     is c / q' with q' = -(b + sign(b) sqrt(det));
//   * two spheres per iteration in the packed FP32 instructions of sm_100 (FFMA2 / FADD2 / FMUL2: one issue slot, two lanes);
//   * the nearest accepted root is selected with one three-input unsigned minimum per sphere (see scan_sm_call).
#pragma once
#include "vpt_f32.cuh"

namespace vpt {
namespace f32 {

constexpr float kSimpleRootMaxR2 = 64.0f * 64.0f; // = vpt_api.cpp kHuge^2: below it the direct-root form, from it on the re-anchored general form

// Scan records are stored in PAIRS: component c of spheres 2j and 2j+1 sits in one 64-bit half of a float4.
struct SmScene {
    MatF mats[kMaxSpheres];
    float4 ga[2 * kMaxSpheres]; // general-form pair j: (qx0 qx1 qy0 qy1) (qz0 qz1 c0_0 c0_1) (mx0 mx1 my0 my1) (mz0 mz1 - -)
    float4 gb[kMaxSpheres];     // direct-root pair j:  (px0 px1 py0 py1) (pz0 pz1 r2_0 r2_1)
    int gid[2 * kMaxSpheres + 2]; // scan slot -> caller's sphere index (general pairs first; -1: the padding slot of an odd class)
    int n_pa, n_pb;             // pairs per class
    int n_emitters, n_area;
    int emitters[kMaxEmitters]; // spheres with any radiance channel > 0 (vptShadeMethods.h:1296), in index order
    int area[kMaxEmitters];     // spheres with r > 0 && radiance.x > 0 (misSamplingFunctions.h:106), in index order
    // Shadow rays of a point light start AT the light (visibility, pathTracingUtilities.h:39-53): the origin part of every sphere test is
    // the same for all of them.  Tables of it for the first kLightTables point lights (stage_scene_tables), used by scan_sm_light.
    int light_slot[kMaxSpheres];                           // sphere index -> table, -1: none
    float4 lg[4][kMaxSpheres / 2][2];                      // general-form pair j seen from the light: PairG (opx opy) (opz c)
    float4 ld[4][kMaxSpheres / 2][2];                      // direct-root pair j: PairD (oqx oqy) (oqz r2)
};
constexpr int kLightTables = 4;
// cooperative staging by the whole block (call, then __syncthreads).  Scan order: general-form spheres first, in scene order, then the
// direct-root ones; every scan record finds its place with one pass over its predecessors.  An odd class is padded with a record that no
// ray can hit (negative discriminant for every ray).
__device__ __forceinline__ void stage_scene(SmScene &S, const SceneF &sc, int tid, int n_threads) {
    for (int i = tid; i < sc.n_spheres * (int)(sizeof(MatF) / 4); i += n_threads)
        reinterpret_cast<uint32_t *>(S.mats)[i] = reinterpret_cast<const uint32_t *>(sc.mat)[i];
    for (int i = tid; i < kMaxEmitters; i += n_threads) { S.emitters[i] = sc.emitters[i]; S.area[i] = sc.area[i]; }
    for (int i = tid; i < kMaxSpheres; i += n_threads) { // the first kLightTables point lights (r == 0 emitters), in index order, get a table
        int slot = -1;
        if (i < sc.n_spheres && sc.mat[i].emits && sc.mat[i].r == 0.0f) {
            slot = 0;
            for (int j = 0; j < i; ++j) slot += (sc.mat[j].emits && sc.mat[j].r == 0.0f);
            if (slot >= kLightTables) slot = -1;
        }
        S.light_slot[i] = slot;
    }
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
        S.n_pa = n_pa; S.n_pb = n_pb; S.n_emitters = sc.n_emitters; S.n_area = sc.n_area;
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

extern __shared__ __align__(16) unsigned char smwave_smem[]; // every kernel's dynamic shared memory starts with its SmScene

struct ScanHit { float t; int index; };
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; } // MUFU.RCP, as __fdividef
// sign flip that ptxas folds into the operand modifiers of the packed instructions: `neg.f32` without .ftz (under -ftz=true the compiler's
// own negation is neg.ftz = a separate flushing FADD per lane; the consuming .FTZ instruction flushes anyway)
__device__ __forceinline__ float neg_fold(float x) { float r; asm("neg.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(neg_fold(a.x), neg_fold(a.y)); }
__device__ __forceinline__ float2 lo2(float4 v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(float4 v) { return make_float2(v.z, v.w); }

// The pair arithmetic in two parts: what depends on the ray's ORIGIN only (shared by all rays that start at the same point -- the three
// next-event rays of a surface vertex, scan_sm_n) and what depends on its direction.  Every add / multiply / fma is one packed instruction for
// both spheres of the pair (same IEEE roundings as a scalar, one-sphere-at-a-time form); the ray's components are broadcast operands.
struct Org2 { float2 x, y, z; };  // origin, each component broadcast to both packed lanes
struct Dir2 { float2 x, y, z; };
__device__ __forceinline__ Org2 org2(F3 o) { return Org2{make_float2(o.x, o.x), make_float2(o.y, o.y), make_float2(o.z, o.z)}; }
__device__ __forceinline__ Dir2 dir2(F3 d) { return Dir2{make_float2(d.x, d.x), make_float2(d.y, d.y), make_float2(d.z, d.z)}; }
struct PairG { float2 opx, opy, opz, c; }; // general form: o - p and |o - p|^2 - r^2 (without cancellation)
struct PairD { float2 oqx, oqy, oqz, r2; }; // direct-root form: o - p and r^2
__device__ __forceinline__ PairG pair_general_origin(const float4 *__restrict__ rec, const Org2 &o) {
    const float4 A = rec[0], B = rec[1], C = rec[2], E = rec[3];
    const float2 mx = lo2(C), my = hi2(C), mz = lo2(E);
    const float2 oqx = __fadd2_rn(o.x, neg2(lo2(A))), oqy = __fadd2_rn(o.y, neg2(hi2(A))), oqz = __fadd2_rn(o.z, neg2(lo2(B)));
    PairG g;
    g.opx = __fadd2_rn(oqx, mx); g.opy = __fadd2_rn(oqy, my); g.opz = __fadd2_rn(oqz, mz);
    g.c = __ffma2_rn(oqx, __fadd2_rn(g.opx, mx), __ffma2_rn(oqy, __fadd2_rn(g.opy, my), __ffma2_rn(oqz, __fadd2_rn(g.opz, mz), hi2(B))));
    return g;
}
// The two roots of both spheres of a pair, each MINUS 1e-4 (w = root - eps: see scan_sm_call).  NaN when the ray misses (det < 0).
__device__ __forceinline__ void pair_general_dir(const PairG &g, const Dir2 &d, float2 &w1, float2 &w2) {
    const float2 meps = make_float2(-kEps, -kEps);
    const float2 b = __ffma2_rn(g.opx, d.x, __ffma2_rn(g.opy, d.y, __fmul2_rn(g.opz, d.z)));
    const float2 det = __ffma2_rn(b, b, neg2(g.c));
    const float2 sq = __fmul2_rn(det, make_float2(rsqrtf(det.x), rsqrtf(det.y))); // NaN when det <= 0
    const float2 q = __fadd2_rn(neg2(b), neg2(make_float2(copysignf(sq.x, b.x), copysignf(sq.y, b.y)))); // the root without cancellation; the other one is c / q
    w1 = __fadd2_rn(q, meps);
    w2 = __ffma2_rn(g.c, make_float2(rcp_approx(q.x), rcp_approx(q.y)), meps);
}
__device__ __forceinline__ PairD pair_direct_origin(const float4 *__restrict__ rec, const Org2 &o) {
    const float4 A = rec[0], B = rec[1];
    return PairD{__fadd2_rn(o.x, neg2(lo2(A))), __fadd2_rn(o.y, neg2(hi2(A))), __fadd2_rn(o.z, neg2(lo2(B))), hi2(B)};
}
__device__ __forceinline__ void pair_direct_dir(const PairD &g, const Dir2 &d, float2 &w1, float2 &w2) {
    const float2 meps = make_float2(-kEps, -kEps);
    const float2 b = __ffma2_rn(g.oqx, d.x, __ffma2_rn(g.oqy, d.y, __fmul2_rn(g.oqz, d.z)));
    const float2 nb2 = neg2(b);
    const float2 lx = __ffma2_rn(d.x, nb2, g.oqx), ly = __ffma2_rn(d.y, nb2, g.oqy), lz = __ffma2_rn(d.z, nb2, g.oqz);
    const float2 det = __ffma2_rn(neg2(lx), lx, __ffma2_rn(neg2(ly), ly, __ffma2_rn(neg2(lz), lz, g.r2)));
    const float2 sq = __fmul2_rn(det, make_float2(rsqrtf(det.x), rsqrtf(det.y)));
    const float2 nbe = __fadd2_rn(nb2, meps);
    w1 = __fadd2_rn(nbe, neg2(sq));
    w2 = __fadd2_rn(nbe, sq);
}
// the smaller accepted root of the pair's two spheres against the best so far (see scan_sm_call), slot = the pair's first scan slot
__device__ __forceinline__ void pair_select(unsigned &best, int &bi, int slot, float2 w1, float2 w2) {
    unsigned k = __vimin3_u32(best, __float_as_uint(w1.x), __float_as_uint(w2.x));
    if (k != best) bi = slot;
    best = k;
    k = __vimin3_u32(best, __float_as_uint(w1.y), __float_as_uint(w2.y));
    if (k != best) bi = slot + 1;
    best = k;
}

// nearest accepted hit over all spheres: distance (+inf: none) and scan index.
// Selection without compares: the reference accepts the near root unless it is below 1e-4, else the far one, and then requires
// t > 1e-4 (Sphere.h:34, pathTracingUtilities.h:20) = the smallest root above 1e-4.  For w = root - 1e-4 the valid candidates are
// exactly the positive floats, whose bit patterns order like unsigned integers, while negative values (sign bit) and the NaN of a
// negative discriminant (0x7fffffff) compare above +inf: ONE three-input unsigned minimum per sphere replaces four compares and
// selects on the half-rate ALU pipe; the index follows with one compare and one select.
// One copy per call site.  Round 1 kept the scan out of line (its megakernels had it at eight divergent sites of one loop body); in the
// staged kernels a stage has one or two scan sites and every warp of the SM is in the same stage, so the copies cost little instruction
// cache and save the call (argument / return moves were 4 % of the executed instructions) and let the loads of the first pair overlap the
// stage's own arithmetic: +2.6 % (equi-angular) / +3.7 % (free flight), profiles/r2_summary.md.  -DVPT_SCAN_OUTLINE restores the call.
#ifdef VPT_SCAN_OUTLINE
#define VPT_SCAN_LINKAGE __noinline__
#else
#define VPT_SCAN_LINKAGE __forceinline__
#endif
#ifdef VPT_SCAN_UNROLL // experiment (tools/build_variant.py -DVPT_SCAN_UNROLL=2): pairs per loop iteration
#define VPT_SCAN_STR(x) #x
#define VPT_SCAN_UNROLL_N(n) _Pragma(VPT_SCAN_STR(unroll n))
#define VPT_SCAN_UNROLL_PRAGMA VPT_SCAN_UNROLL_N(VPT_SCAN_UNROLL)
#else
#define VPT_SCAN_UNROLL_PRAGMA
#endif
static __device__ VPT_SCAN_LINKAGE ScanHit scan_sm_call(float ox, float oy, float oz, float dx, float dy, float dz) {
    const SmScene &S = *reinterpret_cast<const SmScene *>(smwave_smem);
    unsigned best = 0x7f800000u; // +inf
    int bi = -1;
    const int na = S.n_pa, nb = S.n_pb;
    const Org2 o = org2(mk(ox, oy, oz));
    const Dir2 d = dir2(mk(dx, dy, dz));
    VPT_SCAN_UNROLL_PRAGMA
    for (int j = 0; j < na; ++j) {
        float2 w1, w2;
        pair_general_dir(pair_general_origin(&S.ga[4 * j], o), d, w1, w2);
        pair_select(best, bi, 2 * j, w1, w2);
    }
    VPT_SCAN_UNROLL_PRAGMA
    for (int j = 0; j < nb; ++j) {
        float2 w1, w2;
        pair_direct_dir(pair_direct_origin(&S.gb[2 * j], o), d, w1, w2);
        pair_select(best, bi, 2 * (na + j), w1, w2);
    }
    return ScanHit{__uint_as_float(best) + kEps, bi};
}
// N rays from ONE origin in one pass over the spheres: the records are loaded once and the origin part of every pair (o - p, and for the
// general form |o - p|^2 - r^2: 28 of its 46 instructions) is computed once; the N direction parts are independent chains (instruction-
// level parallelism for a kernel that is latency bound at six warps per scheduler).  Per ray the same operations on the same values as
// scan_sm: identical results.  Used by the surface stages, whose next-event rays (one per area light + the BSDF-sampled one) share the
// vertex (vpt_stages.cuh stage_surf).
template <int N>
struct RaysN { F3 d[N]; bool hit[N]; int id[N]; }; // in: directions; out: hit and sphere index per ray
struct RayBest { Dir2 d; unsigned best; int bi; };
__device__ __forceinline__ RayBest ray_best(F3 d) { return RayBest{dir2(d), 0x7f800000u, -1}; }
__device__ __forceinline__ void ray_general(RayBest &r, const PairG &g, int slot) {
    float2 w1, w2;
    pair_general_dir(g, r.d, w1, w2);
    pair_select(r.best, r.bi, slot, w1, w2);
}
__device__ __forceinline__ void ray_direct(RayBest &r, const PairD &g, int slot) {
    float2 w1, w2;
    pair_direct_dir(g, r.d, w1, w2);
    pair_select(r.best, r.bi, slot, w1, w2);
}
// (scalars, not arrays indexed in unrolled loops: that form crashed the compiler's front end now and then)
template <int N>
__device__ __forceinline__ void scan_sm_n(const SmScene &S, F3 o, RaysN<N> &q) {
    static_assert(N == 2 || N == 3, "two or three rays");
    RayBest r0 = ray_best(q.d[0]), r1 = ray_best(q.d[1]), r2 = ray_best(q.d[N - 1]);
    const int na = S.n_pa, nb = S.n_pb;
    const Org2 oo = org2(o);
    for (int j = 0; j < na; ++j) {
        const PairG g = pair_general_origin(&S.ga[4 * j], oo);
        ray_general(r0, g, 2 * j);
        ray_general(r1, g, 2 * j);
        if (N > 2) ray_general(r2, g, 2 * j);
    }
    for (int j = 0; j < nb; ++j) {
        const PairD g = pair_direct_origin(&S.gb[2 * j], oo);
        ray_direct(r0, g, 2 * (na + j));
        ray_direct(r1, g, 2 * (na + j));
        if (N > 2) ray_direct(r2, g, 2 * (na + j));
    }
    q.hit[0] = r0.bi >= 0; q.id[0] = r0.bi >= 0 ? S.gid[r0.bi] : -1;
    q.hit[1] = r1.bi >= 0; q.id[1] = r1.bi >= 0 ? S.gid[r1.bi] : -1;
    if (N > 2) { q.hit[N - 1] = r2.bi >= 0; q.id[N - 1] = r2.bi >= 0 ? S.gid[r2.bi] : -1; }
}
__device__ __forceinline__ bool scan_sm(const SmScene &S, F3 o, F3 d, float &t, int &id) {
    const ScanHit h = scan_sm_call(o.x, o.y, o.z, d.x, d.y, d.z);
    t = h.t;
    id = h.index >= 0 ? S.gid[h.index] : -1;
    return h.index >= 0;
}

// second staging phase (after the __syncthreads that follows stage_scene; then __syncthreads again): the per-light origin tables
__device__ __forceinline__ void stage_scene_tables(SmScene &S, int tid, int n_threads) {
    const int n_pairs = S.n_pa + S.n_pb;
    for (int w = tid; w < kMaxSpheres * n_pairs; w += n_threads) {
        const int i = w / n_pairs, j = w - i * n_pairs, k = S.light_slot[i];
        if (k < 0) continue;
        const Org2 o = org2(mk(S.mats[i].px, S.mats[i].py, S.mats[i].pz));
        if (j < S.n_pa) {
            const PairG g = pair_general_origin(&S.ga[4 * j], o);
            S.lg[k][j][0] = make_float4(g.opx.x, g.opx.y, g.opy.x, g.opy.y); S.lg[k][j][1] = make_float4(g.opz.x, g.opz.y, g.c.x, g.c.y);
        } else {
            const PairD g = pair_direct_origin(&S.gb[2 * (j - S.n_pa)], o);
            S.ld[k][j - S.n_pa][0] = make_float4(g.oqx.x, g.oqx.y, g.oqy.x, g.oqy.y); S.ld[k][j - S.n_pa][1] = make_float4(g.oqz.x, g.oqz.y, g.r2.x, g.r2.y);
        }
    }
}
// scan of a ray that starts at the centre of emitter `src` (a point light's shadow ray): the origin part of every pair comes from the light's
// table -- same values as scan_sm computes, 28 of a general pair's 46 instructions less.  Lights without a table (more than kLightTables
// point lights, area sources) take scan_sm; the choice is made per warp.
__device__ __forceinline__ bool scan_sm_light(const SmScene &S, int src, F3 light, F3 d, float &t, int &id) {
    const int k = S.light_slot[src];
    if (__any_sync(0xffffffffu, k < 0)) return scan_sm(S, light, d, t, id);
    unsigned best = 0x7f800000u; // +inf
    int bi = -1;
    const int na = S.n_pa, nb = S.n_pb;
    const Dir2 dd = dir2(d);
    for (int j = 0; j < na; ++j) {
        const float4 a = S.lg[k][j][0], b = S.lg[k][j][1];
        float2 w1, w2;
        pair_general_dir(PairG{lo2(a), hi2(a), lo2(b), hi2(b)}, dd, w1, w2);
        pair_select(best, bi, 2 * j, w1, w2);
    }
    for (int j = 0; j < nb; ++j) {
        const float4 a = S.ld[k][j][0], b = S.ld[k][j][1];
        float2 w1, w2;
        pair_direct_dir(PairD{lo2(a), hi2(a), lo2(b), hi2(b)}, dd, w1, w2);
        pair_select(best, bi, 2 * (na + j), w1, w2);
    }
    t = __uint_as_float(best) + kEps;
    id = bi >= 0 ? S.gid[bi] : -1;
    return bi >= 0;
}

// Sphere::intersect (Sphere.h:27-37) for ONE sphere through the scan's own pair arithmetic: the near root unless it is negative or within
// 1e-4 of the origin, else the far one (which may be negative); 0 when the ray misses or the sphere has no scan record (r == 0).
__device__ __forceinline__ float sphere_t_sm(const SmScene &S, int sphere, F3 o, F3 d) {
    const Org2 oo = org2(o);
    const Dir2 dd = dir2(d);
    const int n_slots = 2 * (S.n_pa + S.n_pb);
    for (int slot = 0; slot < n_slots; ++slot) {
        if (S.gid[slot] != sphere) continue;
        float2 w1, w2;
        if (slot < 2 * S.n_pa) pair_general_dir(pair_general_origin(&S.ga[4 * (slot >> 1)], oo), dd, w1, w2);
        else pair_direct_dir(pair_direct_origin(&S.gb[2 * ((slot - 2 * S.n_pa) >> 1)], oo), dd, w1, w2);
        const float a = ((slot & 1) ? w1.y : w1.x) + kEps, b = ((slot & 1) ? w2.y : w2.x) + kEps;
        if (!(a == a) || !(b == b)) return 0.0f;
        const float t_near = fminf(a, b), t_far = fmaxf(a, b);
        return (t_near < 0.0f || fabsf(t_near) < kEps) ? t_far : t_near;
    }
    return 0.0f;
}

} // namespace f32
} // namespace vpt
