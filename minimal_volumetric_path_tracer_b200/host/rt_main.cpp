// rt_main.cpp -- C++ host that keeps the reference's command line (`./rt <spp>` -> image.ppm in the cwd, progress on
// stderr, "elapsed time: Ns" on stdout; src/rt.cpp:744-830) and replaces its OpenMP pixel loop (rt.cpp:767-805) with one
// call through the C-ABI of include/vpt.h.  Everything the reference hard-codes is an optional flag here.
//
//   rt <spp> [--method free|equi|mis|march|mis-distance] [--march-step x] [--march-source i] [--scene file] [--size WxH] [--sigma-a x] [--sigma-s x] [--seed n] [--ref] [--gpus n]
//            [--max-depth n] [--continue-prob x] [-o image.ppm] [--pfm image.pfm]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "vpt.h"

static int usage() {
    std::fprintf(stderr, "usage: rt <spp> [--method free|equi|mis|march|mis-distance] [--march-step x] [--march-source i] [--scene file] [--size WxH] [--sigma-a x] [--sigma-s x] [--seed n] [--ref] [--gpus n] "
                         "[--max-depth n] [--continue-prob x] [-o image.ppm] [--pfm image.pfm]\n");
    return 2;
}

int main(int argc, char **argv) {
    if (argc < 2) return usage(); // the reference segfaults here (atoi(NULL), rt.cpp:784)
    const auto start = std::chrono::system_clock::now();
    vpt_params p;
    vpt_default_params(&p);
    p.spp = std::atoi(argv[1]);
    std::string out = "image.ppm", scene_path, pfm;
    int gpus = 1;
    for (int i = 2; i < argc; ++i) {
        const std::string a = argv[i];
        auto val = [&](const char *name) -> const char * { if (i + 1 >= argc) { std::fprintf(stderr, "%s needs a value\n", name); std::exit(2); } return argv[++i]; };
        if (a == "--method") {
            const std::string m = val("--method");
            p.method = m == "free" ? VPT_METHOD_FREE_FLIGHT : m == "equi" ? VPT_METHOD_EQUIANGULAR : m == "mis" ? VPT_METHOD_MIS : m == "march" ? VPT_METHOD_RAYMARCH : m == "mis-distance" ? VPT_METHOD_MIS_DISTANCE : -1;
        } else if (a == "--size") { if (std::sscanf(val("--size"), "%dx%d", &p.width, &p.height) != 2) return usage(); }
        else if (a == "--sigma-a") p.sigma_a = std::atof(val("--sigma-a"));
        else if (a == "--sigma-s") p.sigma_s = std::atof(val("--sigma-s"));
        else if (a == "--march-step") p.march_step = std::atof(val("--march-step"));   // rt.cpp:791: 0.1
        else if (a == "--march-source") p.march_source = std::atoi(val("--march-source")); // rt.cpp:791: 7
        else if (a == "--scene") scene_path = val("--scene");   // instead of editing Sphere.cpp and recompiling
        else if (a == "--seed") p.seed = std::strtoull(val("--seed"), nullptr, 10);
        else if (a == "--max-depth") p.max_depth = std::atoi(val("--max-depth"));
        else if (a == "--continue-prob") p.continue_prob = std::atof(val("--continue-prob"));
        else if (a == "--ref") { p.precision = VPT_PRECISION_FP64_REF; p.quirks = VPT_QUIRKS_REFERENCE; }
        else if (a == "--gpus") gpus = std::atoi(val("--gpus"));
        else if (a == "-o") out = val("-o");
        else if (a == "--pfm") pfm = val("--pfm");   // the frame before the tonemap, binary floats
        else return usage();
    }
    vpt_sphere scene[VPT_MAX_SPHERES];
    const int n = scene_path.empty() ? vpt_default_scene(scene, VPT_MAX_SPHERES) /* Sphere.cpp:11-22 */ : vpt_load_scene(scene_path.c_str(), scene, VPT_MAX_SPHERES);
    if (n < 0) { std::fprintf(stderr, "rt: %s: %s\n", scene_path.c_str(), vpt_strerror(n)); return 1; }
    std::vector<float> hdr((size_t)p.width * p.height * 3);
    vpt_stats st;
    int rc;
    if (gpus > 1) {
        std::vector<int32_t> dev(gpus);
        for (int k = 0; k < gpus; ++k) dev[k] = k;
        rc = vpt_render_multi(&p, scene, n, dev.data(), gpus, hdr.data(), &st);
    } else {
        rc = vpt_render(&p, scene, n, hdr.data(), &st);
    }
    if (rc != VPT_OK) {
        std::fprintf(stderr, "rt: %s %s\n", vpt_strerror(rc), vpt_last_cuda_error());
        return 1;
    }
    std::fprintf(stderr, "\r%5.2f%%\n", 100.0);
    rc = vpt_write_ppm(hdr.data(), p.width, p.height, out.c_str()); // rt.cpp:812-820
    if (rc != VPT_OK) { std::fprintf(stderr, "rt: cannot write %s\n", out.c_str()); return 1; }
    if (!pfm.empty() && vpt_write_pfm(hdr.data(), p.width, p.height, pfm.c_str()) != VPT_OK) { std::fprintf(stderr, "rt: cannot write %s\n", pfm.c_str()); return 1; }
    const std::chrono::duration<double> elapsed = std::chrono::system_clock::now() - start;
    std::fprintf(stderr, "paths %llu  events %llu  scans %llu  kernel %.3f ms  (%.1f Mpaths/s)\n", (unsigned long long)st.paths,
                 (unsigned long long)st.events, (unsigned long long)st.scene_scans, st.kernel_ms, st.paths / (st.kernel_ms * 1e3));
    std::printf("elapsed time: %gs\n", elapsed.count()); // rt.cpp:827
    return 0;
}
