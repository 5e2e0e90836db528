/* include/vpt.h -- C-ABI boundary of the B200-native volumetric path tracer (libvpt_b200.so).
 *
 * The reference (gabo99cas/minimal_volumetric_path_tracer) has no plugin/FFI interface: its hot path is
 * the OpenMP pixel loop src/rt.cpp:767-805, which calls one shade method per sample
 *     Color iterativeVPTracerFree      (const Ray&, double sigma_a, double sigma_s)        vptShadeMethods.h:1263
 *     Color explicitVPTracerRecursive  (const Ray&, double sigma_a, double sigma_s, int)   vptShadeMethods.h:1014
 *     Color MISVPTTracerRecursive      (const Ray&, double sigma_a, double sigma_s, int)   vptShadeMethods.h:1345
 * with the scene in the global `std::vector<Sphere> spheres` (Sphere.h:49, Sphere.cpp:7-22) and the RNG in the
 * global `seed` (Vector.h:38).  vpt_render() replaces that loop in ONE call: the host keeps the set-up
 * (rt.cpp:744-762) and the tonemap / PPM write (rt.cpp:808-829).  INTEGRATION.md shows the call site.
 *
 * Conventions: plain C types only; every function returns 0 (VPT_OK) or a negative vpt_status and never throws;
 * the caller owns every buffer; the only state the library keeps is a per-device scratch-memory pool for the host-buffer entry
 * points (released by vpt_trim()); there is NO CPU fallback -- without a usable CUDA device every compute entry point returns
 * VPT_ERR_NO_DEVICE / VPT_ERR_CUDA.
 */
#ifndef VPT_H
#define VPT_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VPT_MAX_SPHERES 32  /* reference: unbounded std::vector */
#define VPT_MAX_EMITTERS 16 /* reference: `int arr[4]` overflows beyond 4 (vptShadeMethods.h:1293); here a checked limit */
#define VPT_MAX_DEPTH 4095  /* hard bounce cap of every kernel (the on-chip path record keeps the depth in 12 bits).  The reference has no depth limit
                               (roulette only, continue probability 0.6: P(depth > 4095) = 0.6^4095 = 0); see vpt_params.max_depth */
#define VPT_MAX_SAMPLES_PER_CALL (1 << 24) /* samples per pixel rendered by ONE call (sample_end - sample_begin); larger renders are split into calls */

typedef enum {
    VPT_OK = 0,
    VPT_ERR_INVALID_ARGUMENT = -1, /* null pointer, non-positive size, bad enum, sample range outside [0, spp] ... */
    VPT_ERR_SCENE = -2,            /* n_spheres out of range, too many emitters, negative radius, non-finite field */
    VPT_ERR_UNSUPPORTED = -3,      /* material 3 outside VPT_METHOD_VOLUME_SPHERES; that method in fp32 precision; the superseded MEGA_SCAN / WAVEFRONT kernels;
                                      the HBM wavefront in fp64; quirks requested in fp32 precision */
    VPT_ERR_NO_DEVICE = -4,        /* no CUDA device / bad ordinal */
    VPT_ERR_CUDA = -5,             /* a CUDA runtime call failed; vpt_last_cuda_error() has the text */
    VPT_ERR_IO = -6                /* vpt_write_ppm could not write */
} vpt_status;

/* Sphere.h:14-21, same fields, same order as the Sphere constructor (Sphere.h:23). */
typedef struct {
    double r;           /* 0 = point light */
    double p[3];        /* centre */
    double c[3];        /* Lambert albedo */
    double radiance[3]; /* > 0 in any channel = emitter (vptShadeMethods.h:1296) */
    int32_t material;   /* 0 Lambert, 1 Beckmann conductor microfacet, 2 dielectric exactly as the reference writes it (bdsf vptShadeMethods.h:26-46,
                           softDielectric samplingFunctions.h:209, refraxDielectric microFacetUtilities.h:122 -- not Snell's law; used by no scene of
                           Sphere.cpp); 3 volumetric sphere: only with VPT_METHOD_VOLUME_SPHERES (in the reference's three active methods bdsf leaves
                           its pdf and direction unset for it: VPT_ERR_UNSUPPORTED there) */
    int32_t _pad;
    double eta[3], kappa[3];
    double alpha; /* Beckmann roughness */
} vpt_sphere;

/* RAYMARCH = rayMarching3 (rayMarchingMethods.h:330-384, the commented line rt.cpp:791): deterministic constant-step Riemann sum of
 * the single scattering from ONE source (vpt_params.march_source, treated as a point at its centre) along the camera ray up to the
 * first surface; only the pixel jitter is random.  A noise-free check of the medium next-event term.
 * MIS_DISTANCE is NOT in the reference: the reference's "MIS" method (vptShadeMethods.h:1345) is numerically its equi-angular
 * estimator again.  MIS_DISTANCE is the combination its name promises: per vertex ONE of the two distance techniques -- free flight
 * (freeFlightSample, vptSamplingFunctions.h:11) or equi-angular towards the picked source (equiAngularParams2, volumetricBasicFunctions.h:209) --
 * is chosen with probability 1/2 and weighted with the balance heuristic (the sample is divided by the mean of the two densities); everything
 * else (roulette, surface shading, next-event estimation, random-number slots) is method 1's.  Same expectation as methods 0-2, lower variance
 * than either.
 * VOLUME_SPHERES = explicitPathRecursive2 (vptShadeMethods.h:398-497), the reference's only estimator that handles material 3 (volumetric
 * spheres): a legacy surface path tracer in VACUUM -- the global medium (sigma_a, sigma_s, continue_prob of vpt_params) is not
 * used; sigma_a = 0.05, sigma_s = 0.009 and the roulette q = 0.1 are the function's own literals.  A material-3 sphere is ray-marched in 100
 * steps with single scattering from the point lights (punctualVolumetric rayMarchingMethods.h:12, multipleT volumetricBasicFunctions.h:26,
 * Sphere::intersectVPT Sphere.h:39) and the ray continues behind it; spheres are hit from OUTSIDE only (intersectV2 :109 keeps the near
 * root), so it needs a scene without an enclosing room (scenes/scene_volume_spheres.txt).  FP64_REF precision only. */
enum { VPT_METHOD_FREE_FLIGHT = 0, VPT_METHOD_EQUIANGULAR = 1, VPT_METHOD_MIS = 2, VPT_METHOD_RAYMARCH = 3, VPT_METHOD_MIS_DISTANCE = 4, VPT_METHOD_VOLUME_SPHERES = 5 };
enum { VPT_PRECISION_FP32 = 0, VPT_PRECISION_FP64_REF = 1 };
enum { VPT_OUTPUT_SUM = 0, VPT_OUTPUT_MEAN = 1 };
/* FP32 kernel variants (DESIGN.md "Kernels"), same results to fp32 rounding, AUTO = the fastest by measurement (profiles/):
 * MEGA = one thread per pixel, one path vertex per loop iteration (scene scans inside divergent shading code);
 * MEGA_SCAN = scan-converged state machine, exactly one scene scan per loop iteration executed by all lanes;
 * WAVEFRONT = on-chip wavefront, one pool of path records and per-stage queues PER WARP in shared memory;
 * WAVEFRONT_SM = on-chip wavefront, one persistent CTA per SM, one pool per SM, the warps move through the stages together;
 * WAVEFRONT_HBM = the classic multi-kernel wavefront: SoA path-state queues in HBM, one kernel per stage (synchronises the stream). */
enum { VPT_KERNEL_AUTO = 0, VPT_KERNEL_MEGA = 1, VPT_KERNEL_WAVEFRONT = 2, VPT_KERNEL_MEGA_SCAN = 3, VPT_KERNEL_WAVEFRONT_SM = 4, VPT_KERNEL_WAVEFRONT_HBM = 5 };
/* Two behaviours of the reference are decided by FP64 rounding (DESIGN.md "Parity hazards"):
 *  R0_FALLTHROUGH   the in-medium point-light connection is overwritten with 0 whenever the r == 0 sphere registers as
 *                   hit in the solid-angle block (volumetricBasicFunctions.h:310-337 / :251-278);
 *  EXACT_VISIBILITY visibility() tests `t > distance` with no epsilon (pathTracingUtilities.h:48).
 * They exist only in FP64_REF precision, where the reference's operation order is reproduced; FP32 precision implements
 * the well-defined alternative (r == 0 spheres are not ray-intersected; `t > distance * (1 - 1e-4)`). */
enum { VPT_QUIRK_R0_FALLTHROUGH = 1u, VPT_QUIRK_EXACT_VISIBILITY = 2u, VPT_QUIRKS_REFERENCE = 3u, VPT_QUIRKS_NONE = 0u };

typedef struct {
    int32_t width, height;            /* rt.cpp:752 (1024 x 768) */
    int32_t spp;                      /* argv[1], rt.cpp:784: samples per pixel of the WHOLE render */
    int32_t sample_begin, sample_end; /* this call renders samples [begin, end); 0,0 = [0, spp) (shard / resume) */
    int32_t tile_rank, tile_count;    /* this call renders pixel tiles with tile_id % tile_count == tile_rank; 0,0 = all */
    int32_t method;                   /* VPT_METHOD_*: which line of rt.cpp:791-796 is active */
    int32_t max_depth;                /* <= 0: roulette only, as the reference (needs continue_prob <= 0.99; paths are cut at VPT_MAX_DEPTH bounces, which
                                         roulette reaches with probability < 1e-17); otherwise 1 .. VPT_MAX_DEPTH.  continue_prob > 0.99 needs max_depth > 0:
                                         a roulette that never fires would never end a path in a scene without emitter geometry */
    double sigma_a, sigma_s;          /* 0.001, 0.009 (rt.cpp:794) */
    double continue_prob;             /* 0.6 (vptShadeMethods.h:1275) */
    double cam_o[3], cam_dir[3], fov; /* (0,11.2,214), (0,-0.042612,-1), 0.5095 (rt.cpp:755-759) */
    uint64_t seed;                    /* Philox key; replaces getentropy (rt.cpp:746) */
    uint32_t quirks;                  /* VPT_QUIRK_* (FP64_REF only) */
    int32_t precision;                /* VPT_PRECISION_* */
    int32_t output;                   /* VPT_OUTPUT_SUM: per-pixel sum over the rendered samples; MEAN: sum / spp (rt.cpp:800) */
    int32_t kernel;                   /* VPT_KERNEL_* */
    int32_t device;                   /* CUDA ordinal */
    int32_t march_source;             /* VPT_METHOD_RAYMARCH: sphere index of the source; 7 in the commented call rt.cpp:791 (an area light: its own
                                         sphere blocks the centre-origin shadow ray, the image is black outside it); 8 is the point light */
    double march_step;                /* VPT_METHOD_RAYMARCH: step length, 0.1 in rt.cpp:791 */
} vpt_params;

typedef struct {
    uint64_t paths;       /* camera paths traced = pixels rendered x samples */
    uint64_t events;      /* path vertices that survived Russian roulette */
    uint64_t scene_scans; /* all-sphere intersection passes */
    uint64_t nonfinite;   /* radiance contributions that were NaN/Inf or >= 2^32 in magnitude (outside the range of the 64-bit fixed-point pixel sums,
                             2^-30 .. 2^33) and were dropped (0 in a healthy run; FP64_REF: whole paths) */
    double kernel_ms;     /* CUDA-event time of the render kernels only */
    double total_ms;      /* wall time of the call including copies */
    uint64_t launches;    /* kernels launched by this call */
} vpt_stats;

/* The reference's literals (see the field comments); spp = 64, precision FP32, quirks NONE, output MEAN. */
void vpt_default_params(vpt_params *p);
/* Sphere.cpp:11-22 restated as data; returns the sphere count (10) or VPT_ERR_INVALID_ARGUMENT if cap is too small. */
int vpt_default_scene(vpt_sphere *out, int32_t cap);

/* Scene files (scenes without recompiling; the reference edits Sphere.cpp:7-106): text, one sphere per line, 18 numbers in the
 * argument order of the Sphere constructor (Sphere.h:23) -- radius, centre[3], albedo[3], radiance[3], material, eta[3], kappa[3], alpha --
 * separated by blanks or commas; `#` starts a comment.  Returns the sphere count, VPT_ERR_IO (cannot read), VPT_ERR_SCENE (malformed line,
 * more than `cap` spheres).  scenes/ holds the reference's active scene and its five commented alternates. */
int vpt_load_scene(const char *path, vpt_sphere *out, int32_t cap);

/* Replaces the pixel loop rt.cpp:767-805.  hdr_rgb: HOST buffer, width*height*3 floats, pixel index (h-y-1)*w+x as
 * rt.cpp:773 (row 0 = top of the image), UNCLAMPED linear radiance; pixels outside this call's tiles are written as 0.
 * Any host memory works.  If hdr_rgb is page-locked and mapped (vpt_host_alloc, cudaHostAlloc, cudaHostRegister) the kernel stores its
 * pixels straight into it over PCIe while it renders (no staging copy after the kernel); pageable memory gets a device frame + one copy.
 * The calling thread's current CUDA device is left unchanged. */
int vpt_render(const vpt_params *p, const vpt_sphere *spheres, int32_t n_spheres, float *hdr_rgb, vpt_stats *stats /* nullable */);
/* Same, but hdr_rgb is a DEVICE buffer on p->device and the work is enqueued on `cuda_stream` (a cudaStream_t, may be 0).
 * With stats == NULL the call does not synchronise; with stats != NULL it synchronises the stream before returning. */
int vpt_render_device(const vpt_params *p, const vpt_sphere *spheres, int32_t n_spheres, float *hdr_rgb_device, void *cuda_stream, vpt_stats *stats);
/* One frame sharded over n_devices GPUs of this process by interleaved pixel tiles (one host thread per device).  The tiles
 * are disjoint, so each device copies its own strided tile ranges straight into the caller's buffer: no reduction step.
 * (Multi-process sharding -- one rank per GPU and one NCCL reduce of the HDR buffers -- lives in the host layer,
 * minimal_volumetric_path_tracer_b200/distributed.py.) */
int vpt_render_multi(const vpt_params *p, const vpt_sphere *spheres, int32_t n_spheres, const int32_t *devices, int32_t n_devices, float *hdr_rgb, vpt_stats *stats);

/* Host output stage, byte-identical to rt.cpp:803,812-820 + mathUtilities.h:34-45: clamp, gamma 2.2, "P3" text.
 * hdr_rgb holds per-pixel MEAN radiance. */
int vpt_tonemap(const float *hdr_rgb, int32_t width, int32_t height, uint8_t *rgb8_out);
int vpt_write_ppm(const float *hdr_rgb, int32_t width, int32_t height, const char *path);
/* The frame before the tonemap as a binary colour PFM ("PF", scale -1.0 = little-endian floats, rows bottom to top): the side output
 * SURVEY.md 8(f)-1 asks for, so that two renders can be compared without the 8-bit clamp (the reference has no counterpart: it only
 * keeps the clamped values, rt.cpp:800).  hdr_rgb as above (row 0 = top of the image). */
int vpt_write_pfm(const float *hdr_rgb, int32_t width, int32_t height, const char *path);

/* Unit kernels: one device thread per row evaluates the device implementation of one reference function, so that it
 * can be compared with the reference's C++ function on identical inputs (tolerance 1e-5 relative in FP32).
 * `in` is n x in_stride doubles, `out` is n x out_stride doubles (host).  Row layouts: see vpt_unit below. */
typedef enum {
    VPT_UNIT_SPHERE_INTERSECT = 0, /* Sphere.h:27            in: sphere_index, o[3], d[3]              out: t */
    VPT_UNIT_INTERSECT = 1,        /* pathTracingUtilities.h:12  in: o[3], d[3]                        out: hit, t, id */
    VPT_UNIT_VISIBILITY = 2,       /* pathTracingUtilities.h:39  in: light[3], x[3]                    out: visible */
    VPT_UNIT_TRANSMITTANCE = 3,    /* volumetricBasicFunctions.h:14  in: x1[3], x2[3], sigma_t         out: T */
    VPT_UNIT_FREE_FLIGHT = 4,      /* vptSamplingFunctions.h:11-31  in: sigma_t, xi                    out: d, freeFlightProb(d), pdfSuccess(d), pdfFailure(d) */
    VPT_UNIT_PHASE_SAMPLE = 5,     /* vptSamplingFunctions.h:34  in: xi1, xi2                          out: w[3] */
    VPT_UNIT_EQUIANGULAR = 6,      /* volumetricBasicFunctions.h:209 + vptSamplingFunctions.h:60
                                      in: source_index, tMax, o[3], d[3], xi                          out: D, thetaA, thetaB, t_local, t_ray, pdf */
    VPT_UNIT_POWER_HEURISTIC = 7,  /* misSamplingFunctions.h:12  in: f, g                              out: w */
    VPT_UNIT_COSINE_HEMISPHERE = 8,/* samplingFunctions.h:47     in: n[3], xi1, xi2                    out: w[3], hemiCosineProb(n.w) */
    VPT_UNIT_CONE_SAMPLE = 9,      /* samplingFunctions.h:65,85  in: wc[3], r, dist, xi1, xi2 (cos_max = sqrt(1-(r/dist)^2))  out: w[3], solidAngleProb */
    VPT_UNIT_MICROFACET = 10,      /* microFacetUtilities.h:21-100  in: eta[3], kappa[3], alpha, wi[3], wo[3] (local, n = z; wh = normalize(wi+wo))
                                                                                                     out: fr[3], microFacetProb, NDF, G_smith */
    VPT_UNIT_FACET_NORMAL = 11,    /* microFacetUtilities.h:71   in: alpha, xi1, xi2                   out: wh[3] */
    VPT_UNIT_MEDIUM_NEE = 12,      /* volumetricBasicFunctions.h:284 (T_xt < 0) / :225 (T_xt >= 0)
                                      in: xt[3], source_index, sigma_t, sigma_s, T_xt, prob_source, xi1, xi2   out: Ld[3] */
    VPT_UNIT_POINT_LIGHT = 13,     /* vptShadeMethods.h:62       in: obj_index, x[3], n[3], wray[3], source_index   out: L[3] */
    VPT_UNIT_SURFACE_MIS = 14,     /* misSamplingFunctions.h:96  in: obj_index, x[3], n[3], wray[3], sigma_t, xi[8]   out: L[3] */
    VPT_UNIT_BSDF_SAMPLE = 15,     /* vptShadeMethods.h:16       in: obj_index, n[3], wray[3], xi1, xi2      out: fs[3], wi[3], pdf */
    VPT_UNIT_RADIANCE = 16,        /* the three shade methods    in: o[3], d[3], pixel, sample  (Philox stream (pixel, sample), jitter draws skipped;
                                      method, sigma_*, continue_prob, max_depth, seed, quirks from `p`)  out: L[3], events */
    VPT_UNIT_CAMERA_RAY = 17,      /* rt.cpp:787                 in: x, y, xi1, xi2 (width, height, camera from `p`)   out: d[3] */
    VPT_UNIT_RADIANCE_LIST = 18,   /* the three shade methods on an EXPLICIT list of uniforms (e.g. the reference's erand48 sequence)
                                      in: o[3], d[3], n_u, u[120]                                      out: L[3], draws used (-1: list too short) */
    VPT_UNIT_RAYMARCH = 19,        /* rayMarchingMethods.h:330   in: o[3], d[3], step, source_index (sigma_* from `p`)            out: L[3], steps */
    VPT_UNIT_MIS_DISTANCE = 20,    /* VPT_METHOD_MIS_DISTANCE's distance decision (not in the reference; free-flight vptSamplingFunctions.h:11-20 and
                                      equi-angular volumetricBasicFunctions.h:209 + vptSamplingFunctions.h:60 under the balance heuristic)
                                      in: source_index, tMax, o[3], d[3], sigma_t, xi, xd                out: surface (0/1), distance, mixture pdf */
    VPT_UNIT_DIELECTRIC = 21,      /* microFacetUtilities.h:107-141 as used by bdsf / softDielectric / MISv2 for material 2 (etai 1, etat 1.5)
                                      in: n[3], wo[3] (unit, world; wo = -ray direction)   out: normalize(refraxDielectric)[3], normalize(reflexDielectric)[3], fresnelDie */
    VPT_UNIT_COUNT_
} vpt_unit_fn;
int vpt_unit(int32_t fn, const vpt_params *p, const vpt_sphere *spheres, int32_t n_spheres, int32_t n, const double *in, int32_t in_stride,
             double *out, int32_t out_stride);
int vpt_unit_strides(int32_t fn, int32_t *in_stride, int32_t *out_stride);

/* Philox4x32-10 on the device (the stream the renders use): ctr/key/out are host arrays of n x 4 / n x 2 / n x 4 words. */
int vpt_philox(int32_t device, int32_t n, const uint32_t *ctr, const uint32_t *key, uint32_t *out);

/* Register-resident FFMA microbenchmark: the measured FP32 roofline denominator. */
int vpt_measure_fp32_peak(int32_t device, double *tflops_out, double *sm_clock_mhz_out /* nullable */);

/* Page-locked, device-mapped host memory for the frame buffers handed to vpt_render / vpt_render_multi (optional, see vpt_render). */
void *vpt_host_alloc(size_t bytes);
void vpt_host_free(void *p);

int vpt_device_count(void);
const char *vpt_strerror(int status);
const char *vpt_last_cuda_error(void); /* thread-local text of the last CUDA failure seen by this thread */
const char *vpt_version(void);
void vpt_trim(void); /* give the per-device scratch pools (frame buffers of vpt_render between calls) back to the driver */

#ifdef __cplusplus
}
#endif
#endif /* VPT_H */
