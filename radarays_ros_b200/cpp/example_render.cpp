// example_render.cpp — the C ABI from plain C++ (no Python, no ROS): load a map file, set RadarParams, render frames.
//
//   make -C radarays_ros_b200/csrc            (builds radarays_ros_b200/example_render next to the library)
//   radarays_ros_b200/example_render map.ply out.pgm [x y z yaw] [n_poses]
//
// Mirrors what src/radar_simulator.cpp does around its backend: import the map (:149), load materials (Radar.cpp:220-226),
// take the dynamic-reconfigure defaults (cfg/RadarModel.cfg), call simulate() and hand out the mono8 polar image.
#include <radarays_b200.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

static void check(rr_ctx* ctx, int rc, const char* what)
{
    if (rc == RR_OK) return;
    std::fprintf(stderr, "%s failed (%d): %s\n", what, rc, rr_last_error(ctx));
    std::exit(1);
}

int main(int argc, char** argv)
{
    if (argc < 3) { std::fprintf(stderr, "usage: %s map.{ply,obj,dae} out.pgm [x y z yaw] [n_poses]\n", argv[0]); return 2; }
    const double x = argc > 6 ? std::atof(argv[3]) : 0.0, y = argc > 6 ? std::atof(argv[4]) : 0.0;
    const double z = argc > 6 ? std::atof(argv[5]) : 1.0, yaw = argc > 6 ? std::atof(argv[6]) : 0.0;
    const size_t n_poses = argc > 7 ? (size_t)std::atoi(argv[7]) : 1;

    rr_ctx* ctx = nullptr;
    check(nullptr, rr_create(&ctx, 0), "rr_create");
    uint32_t n_objects = 0;
    check(ctx, rr_set_mesh_file(ctx, argv[1], &n_objects), "rr_set_mesh_file");

    // materials: air + one wall material for every object (config/mulran_kaist02.yaml:10-18)
    const rr_material mats[2] = {{0.3f, 1.0f, 0.0f, 1.0f}, {0.0f, 1.0f, 0.0f, 3000.0f}};
    std::vector<int32_t> object_materials(n_objects, 1);
    check(ctx, rr_set_materials(ctx, mats, 2, object_materials.data(), object_materials.size(), 0), "rr_set_materials");

    rr_config cfg;
    rr_config_defaults(&cfg);
    cfg.include_motion = 0;
    check(ctx, rr_set_params(ctx, nullptr, &cfg), "rr_set_params");
    check(ctx, rr_set_beam_samples(ctx, nullptr, 0, 1), "rr_set_beam_samples");

    std::vector<rr_pose> poses(n_poses);
    for (size_t i = 0; i < n_poses; i++) {
        const double a = yaw + 0.01 * (double)i;
        poses[i] = rr_pose{0.f, 0.f, (float)std::sin(a / 2), (float)std::cos(a / 2), (float)(x + 0.05 * (double)i), (float)y, (float)z};
    }
    std::vector<uint8_t> img(n_poses * (size_t)cfg.n_cells * RR_N_ANGLES);
    rr_stats st;
    const auto t0 = std::chrono::steady_clock::now();
    check(ctx, rr_simulate(ctx, poses.data(), n_poses, 0, img.data(), &st), "rr_simulate");
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::printf("%zu frame(s): %llu casts, %llu hits, %llu returns, BVH %llu nodes built in %.1f ms, kernels %.3f ms, call %.3f ms\n",
                n_poses, (unsigned long long)st.n_casts, (unsigned long long)st.n_hits, (unsigned long long)st.n_signals,
                (unsigned long long)st.bvh_nodes, st.bvh_build_ms, st.kernel_ms, ms);

    FILE* f = std::fopen(argv[2], "wb");                       // first frame as binary PGM: rows = range bins, cols = azimuths
    if (!f) { std::perror(argv[2]); return 1; }
    std::fprintf(f, "P5\n%d %d\n255\n", RR_N_ANGLES, cfg.n_cells);
    std::fwrite(img.data(), 1, (size_t)cfg.n_cells * RR_N_ANGLES, f);
    std::fclose(f);
    rr_destroy(ctx);
    return 0;
}
