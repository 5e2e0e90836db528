// vpt_kernels_hbm.cu -- kernels and host loop of the multi-kernel HBM wavefront (VPT_KERNEL_WAVEFRONT_HBM, vpt_hbmwave.cuh).
#include <cuda_runtime.h>
#include <cstdlib>
#include "vpt_hbmwave.cuh"

namespace vpt {

using namespace f32;

enum : int { HK_GEN = SQ_COUNT };

// one thread: close the previous round, open the next one
__global__ void hbm_plan_kernel(HbmState H) {
    H.count[SQ_SURF_F] = 0u; // consumed by the last kernel of the previous round
    const unsigned long long cur = H.cursor[0];
    const unsigned in_flight = H.count[SQ_PRIMARY];
    const unsigned long long left = H.n_total - cur;
    unsigned room = H.cap - in_flight;
    room &= ~31u; // whole warps
    const unsigned n_new = (unsigned)(left < (unsigned long long)room ? left : (unsigned long long)room);
    H.cursor[1] = cur; H.cursor[0] = cur + n_new;
    H.count[8] = n_new;
    if (n_new == 0u && in_flight == 0u) H.count[9] += 1u; // nothing left: the host stops enqueueing rounds when it sees this
}

template <int METHOD, int STAGE>
__global__ void __launch_bounds__(kHbmThreads) hbm_stage_kernel(const __grid_constant__ SceneF sc, const __grid_constant__ LaunchParams lp, const __grid_constant__ ConstsF cf,
                                                                 const __grid_constant__ HbmState H, int reset_q) {
    if (blockIdx.x == 0 && threadIdx.x == 0 && reset_q >= 0) H.count[reset_q] = 0u; // the queue the previous kernel consumed
    const unsigned n = STAGE == HK_GEN ? H.count[8] : H.count[STAGE];
    const unsigned per_pass = gridDim.x * blockDim.x;
    if (blockIdx.x * blockDim.x >= n) return; // nothing for this block (an empty queue costs a launch, not a scene staging)
    __shared__ unsigned blk_cnt[8], blk_base[8];
    SmScene &S = *reinterpret_cast<SmScene *>(smwave_smem);
    stage_scene(S, sc, (int)threadIdx.x, (int)blockDim.x);
    if (threadIdx.x < 8) blk_cnt[threadIdx.x] = 0u;
    __syncthreads();
    if (STAGE == SQ_MED_POINT || STAGE == SQ_SURF_P) { stage_scene_tables(S, (int)threadIdx.x, (int)blockDim.x); __syncthreads(); }
    HbmCtx C(S, cf, lp, H);
    const unsigned lane = threadIdx.x & 31u;
    for (unsigned first = blockIdx.x * blockDim.x; first < n; first += per_pass) { // block-uniform trip count
        const unsigned i = first + threadIdx.x;
        const bool act = i < n;
        Rec r;
        int dest;
        if (STAGE == HK_GEN) dest = C.generate(act, H.cursor[1] + i, r);
        else {
            if (act) r = hbm_load(H, STAGE, i);
            else { r.o = mk(0, 0, 0); r.d = mk(0, 0, 1); r.beta = mk(0, 0, 0); r.xi_dist = r.xi_decide = 0.5f; r.pixel = r.sample = r.depth = r.src = r.hid = r.aux = 0u; }
            if (STAGE == SQ_PRIMARY) dest = stage_primary<METHOD>(C, act, r);
            else if (STAGE == SQ_MED_POINT) dest = stage_med<true>(C, act, r);
            else if (STAGE == SQ_MED_AREA) dest = stage_med<false>(C, act, r);
            else if (STAGE == SQ_SURF_P) dest = stage_surf_p(C, act, r);
            else if (STAGE == SQ_SURF_L) dest = stage_surf<false>(C, act, r);
            else dest = stage_surf<true>(C, act, r);
            if (dest == kDestFree) dest = -1; // the path ended: its record simply is not written anywhere
        }
        // ---- block-wide emit: offsets inside the block from shared-memory counters, one global atomic per destination and block ----
        const unsigned grp = __match_any_sync(0xffffffffu, dest);
        unsigned local = 0;
        if (dest >= 0) {
            const int leader = __ffs(grp) - 1;
            if ((int)lane == leader) local = atomicAdd(&blk_cnt[dest], (unsigned)__popc(grp));
            local = __shfl_sync(grp, local, leader) + __popc(grp & ((1u << lane) - 1u));
        }
        __syncthreads();
        if (threadIdx.x < SQ_COUNT && blk_cnt[threadIdx.x]) blk_base[threadIdx.x] = atomicAdd(&H.count[threadIdx.x], blk_cnt[threadIdx.x]);
        __syncthreads();
        if (dest >= 0) hbm_store(H, dest, blk_base[dest] + local, r);
        if (threadIdx.x < 8) blk_cnt[threadIdx.x] = 0u;
        __syncthreads();
    }
    if (H.tally) C.flush_tally();
}

// fixed-point sums -> float frame, owned pixels only
__global__ void hbm_resolve_kernel(const __grid_constant__ LaunchParams lp, HbmState H, float *__restrict__ hdr) {
    for (unsigned op = blockIdx.x * blockDim.x + threadIdx.x; op < H.n_owned_pixels; op += gridDim.x * blockDim.x) {
        const unsigned pixel = ((op >> 7) * (unsigned)lp.tile_count + (unsigned)lp.tile_rank) * (unsigned)kTile + (op & (unsigned)(kTile - 1));
        if (pixel >= (unsigned)lp.n_pixels) continue;
        for (int c = 0; c < 3; ++c) hdr[(size_t)pixel * 3 + c] = (float)((double)(long long)H.acc[(size_t)pixel * 3 + c] * kSmFixInv * lp.out_scale);
    }
}

template <int METHOD, int STAGE>
static cudaError_t launch_stage(const SceneF &scene, const LaunchParams &lp, const ConstsF &cf, const HbmState &H, int reset_q, int grid, cudaStream_t st) {
    hbm_stage_kernel<METHOD, STAGE><<<grid, kHbmThreads, sizeof(SmScene), st>>>(scene, lp, cf, H, reset_q);
    return cudaGetLastError();
}

template <int METHOD>
static int run_hbmwave(const SceneF &scene, const LaunchParams &lp, const ConstsF &cf, float *hdr_dev, Counters *counters_dev, cudaStream_t st, int n_owned_tiles,
                       uint64_t *launches) {
    int dev = 0, n_sm = 0;
    cudaError_t e;
#define HBM_TRY(x) do { e = (x); if (e != cudaSuccess) goto fail; } while (0)
    unsigned *flags_host = nullptr;
    void *block = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    HbmState H{};
    {
        if ((e = cudaGetDevice(&dev)) != cudaSuccess) return (int)e;
        if ((e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return (int)e;
        unsigned cap = 1u << 23; // 8 Mi paths in flight = 3.6 GB of queues: best of 1 / 2 / 4 / 6 / 8 / 12 Mi by measurement
#ifdef VPT_DEV_KNOBS // profiling builds only (tools/build_variant.py -DVPT_DEV_KNOBS): the shipped library does not read the environment
        if (const char *cap_env = std::getenv("VPT_HBM_CAP")) cap = (unsigned)std::strtoul(cap_env, nullptr, 10);
        cap = cap < (1u << 16) ? (1u << 16) : (cap > (1u << 25) ? (1u << 25) : cap);
#endif
        cap = (cap + 1023u) & ~1023u;
        H.cap = cap;
        H.n_owned_pixels = (unsigned)n_owned_tiles * (unsigned)kTile;
        H.n_total = (unsigned long long)H.n_owned_pixels * (unsigned long long)(lp.sample_end - lp.sample_begin);
        const size_t bytes_f = (size_t)SQ_COUNT * HF_FLOATS * cap * 4, bytes_u = (size_t)SQ_COUNT * HU_WORDS * cap * 4, bytes_acc = (size_t)lp.n_pixels * 3 * 8;
        const size_t bytes_ctl = 256;
        HBM_TRY((cudaError_t)scratch_alloc_(dev, &block, bytes_f + bytes_u + bytes_acc + bytes_ctl, (void *)st));
        char *p = (char *)block;
        H.f = (float *)p; p += bytes_f;
        H.u = (uint32_t *)p; p += bytes_u;
        H.acc = (unsigned long long *)p; p += bytes_acc;
        H.count = (unsigned *)p; H.cursor = (unsigned long long *)(p + 64); H.tally = counters_dev ? (unsigned long long *)(p + 128) : nullptr;
        HBM_TRY(cudaMemsetAsync(H.acc, 0, bytes_acc + bytes_ctl, st));
        HBM_TRY(cudaMallocHost((void **)&flags_host, 64));
        flags_host[0] = flags_host[1] = 0u;
        HBM_TRY(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
        HBM_TRY(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
        int ctas_per_sm = 4; // best of 3 / 4 / 6 / 8 / 16 by measurement
#ifdef VPT_DEV_KNOBS
        if (const char *grid_env = std::getenv("VPT_HBM_GRID")) ctas_per_sm = std::atoi(grid_env);
        ctas_per_sm = ctas_per_sm < 1 ? 1 : (ctas_per_sm > 32 ? 32 : ctas_per_sm);
#endif
        const int grid = n_sm * ctas_per_sm;
        // rounds are enqueued in batches; after each batch the "nothing left" counter is copied to the host, and the host looks at the
        // copy of the batch BEFORE the one it has just enqueued, so the GPU never waits for the host
        const int kBatch = 16;
        const long long max_batches = (long long)(H.n_total / (cap / 64u + 1u) + 4096ull) / kBatch + 2; // a bug must not enqueue rounds forever
        for (long long batch = 0;; ++batch) {
            if (batch > max_batches) { e = cudaErrorLaunchTimeout; goto fail; }
            for (int r = 0; r < kBatch; ++r) {
                hbm_plan_kernel<<<1, 1, 0, st>>>(H);
                HBM_TRY(cudaGetLastError());
                HBM_TRY((launch_stage<METHOD, HK_GEN>(scene, lp, cf, H, -1, grid, st)));
                HBM_TRY((launch_stage<METHOD, SQ_PRIMARY>(scene, lp, cf, H, -1, grid, st)));
                HBM_TRY((launch_stage<METHOD, SQ_SURF_P>(scene, lp, cf, H, SQ_PRIMARY, grid, st)));
                HBM_TRY((launch_stage<METHOD, SQ_MED_POINT>(scene, lp, cf, H, SQ_SURF_P, grid, st)));
                HBM_TRY((launch_stage<METHOD, SQ_MED_AREA>(scene, lp, cf, H, SQ_MED_POINT, grid, st)));
                HBM_TRY((launch_stage<METHOD, SQ_SURF_L>(scene, lp, cf, H, SQ_MED_AREA, grid, st)));
                HBM_TRY((launch_stage<METHOD, SQ_SURF_F>(scene, lp, cf, H, SQ_SURF_L, grid, st)));
                if (launches) *launches += 8;
            }
            HBM_TRY(cudaMemcpyAsync(&flags_host[batch & 1], &H.count[9], 4, cudaMemcpyDeviceToHost, st));
            HBM_TRY(cudaEventRecord(ev[batch & 1], st));
            if (batch > 0) {
                HBM_TRY(cudaEventSynchronize(ev[(batch - 1) & 1]));
                if (flags_host[(batch - 1) & 1] != 0u) break;
            }
        }
        hbm_resolve_kernel<<<n_sm * 4, 256, 0, st>>>(lp, H, hdr_dev);
        HBM_TRY(cudaGetLastError());
        if (launches) *launches += 1;
        if (counters_dev) HBM_TRY(cudaMemcpyAsync(counters_dev, H.tally, 32, cudaMemcpyDeviceToDevice, st));
    }
    e = cudaSuccess;
fail:
    if (block) cudaFreeAsync(block, st);
    if (ev[0]) cudaEventDestroy(ev[0]);
    if (ev[1]) cudaEventDestroy(ev[1]);
    if (flags_host) { cudaStreamSynchronize(st); cudaFreeHost(flags_host); }
    scratch_trim_(dev, (size_t)256 << 20); // the queues (3.6 GB) go back to the driver; a frame buffer's worth stays pooled
    return (int)e;
#undef HBM_TRY
}

int launch_hbmwave_f32(const SceneF &scene, const LaunchParams &lp, const ConstsF &cf, float *hdr_dev, Counters *counters_dev, void *stream, int n_owned_tiles,
                       uint64_t *launches) {
    cudaStream_t st = (cudaStream_t)stream;
    switch (lp.method) {
    case 0: return run_hbmwave<0>(scene, lp, cf, hdr_dev, counters_dev, st, n_owned_tiles, launches);
    case 1: return run_hbmwave<1>(scene, lp, cf, hdr_dev, counters_dev, st, n_owned_tiles, launches);
    case 4: return run_hbmwave<4>(scene, lp, cf, hdr_dev, counters_dev, st, n_owned_tiles, launches);
    default: return run_hbmwave<2>(scene, lp, cf, hdr_dev, counters_dev, st, n_owned_tiles, launches);
    }
}

} // namespace vpt
