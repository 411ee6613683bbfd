/*
 * launcher_main.cpp — the CLI of the reference launchers: `rt_render <num_rays> <num_bounce>`
 * (optimized.cu:774-785, cpu_launcher.cpp:654-659; consumed by benchmark.py:20), writing a PNG in the CWD
 * and printing `Rendering time: <s> s` (optimized.cu:879-881). Extras beyond the reference:
 *   --profile optimized|cpu|array_bvh   knob set + scene layout of that reference program (default optimized)
 *   --width W --height H                (the reference hard-codes 512x512, optimized.cu:786-787)
 *   --obj PATH                          default cadnav.com_model/Models_F0202A090/cat.obj relative to the CWD (:802)
 *   --out FILE                          default image_optimized.png (:862) / image.png (cpu_launcher.cpp:719)
 *   --device D   --frames F             render F frames (kernel time is reported per frame)
 *   --out-pattern P   --orbit W         with --frames: write EVERY frame to P (a printf pattern with one %d), encoded and
 *                                       written in the background while the next frames render (rt_png_writer_*); between
 *                                       frames the light moves by W rad/s x 0.02 s (MoveLightSource, realtime_render.cu:1072-1090)
 *   --gpu-build                         build the BVH on the device (same tree; the reference builds on the host, :809-813)
 *   --stochastic                        the reference's own default: sigma 0.2 Box-Muller jitter + cosine-weighted
 *                                       indirect bounce on the cuRAND XORWOW stream of optimized.cu:745 (without the
 *                                       flag: the deterministic mode the parity contract is stated on; a notice says so)
 *   --gpus N                            one process, N devices: the scene is built on device 0 and broadcast (rt_scene_broadcast,
 *                                       NCCL), every device renders every N-th group of --row-group G (default 4) consecutive rows of
 *                                       each frame on its own thread, the bands are gathered to device 0 (rt_gather_framebuffer_groups)
 *   --mirror                            the mesh is a mirror (Geometry::mirror, optimized.cu:111): BASELINE.json configs[2]
 */
#include "scene.hpp"

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <thread>

/* --out-pattern: exactly one integer conversion (%d or %0Nd) and no other '%': the string goes to snprintf as a format */
static bool pattern_ok(const std::string& p) {
    int conv = 0;
    for (size_t i = 0; i < p.size(); i++) {
        if (p[i] != '%') continue;
        size_t j = i + 1;
        while (j < p.size() && p[j] >= '0' && p[j] <= '9') j++;
        if (j >= p.size() || p[j] != 'd') return false;
        conv++;
        i = j;
    }
    return conv == 1;
}

/* --gpus N: devices 1 .. N-1 (device 0 is driven by the caller's own scene). Row-interleaved bands, gathered to device 0. */
struct MultiGpu {
    int n = 1, W = 0, H = 0;
    std::vector<rt_comm*> comms;
    std::vector<rt_scene*> scenes; /* [0] is the caller's */
    std::vector<void*> bands;      /* device buffers, one per device */
    void* frame = nullptr;         /* device 0 */

    int G = 4; /* rows per group of the interleave (--row-group): neighbouring rows keep a warp's pixel tile a tile */
    int rows_of(int H_, int r, int n_) const { return rt_shard_rows(H_, r, n_, G, nullptr); }

    void init(int n_, rt_scene* root, int W_, int H_) {
        n = n_; W = W_; H = H_;
        comms.assign(n, nullptr);
        rtb200::check(rt_comm_init_all(comms.data(), n, nullptr));
        scenes.assign(n, nullptr);
        scenes[0] = root;
        bands.assign(n, nullptr);
        uint8_t handle[64];
        for (int d = 0; d < n; d++) {
            if (d > 0) rtb200::check(rt_scene_create(&scenes[d], d));
            rtb200::check(rt_peer_alloc(d, (size_t)std::max(rows_of(H, d, n), 1) * W * 3, &bands[d], handle));
        }
        rtb200::check(rt_peer_alloc(0, (size_t)W * H * 3, &frame, handle));
        each([&](int d) { rtb200::check(rt_scene_broadcast(scenes[d], comms[d], 0, nullptr)); });
    }
    template <typename F>
    void each(F f) { /* one thread per device: NCCL collectives of one process must be entered concurrently */
        std::vector<std::thread> th;
        std::vector<std::string> err(n);
        for (int d = 0; d < n; d++)
            th.emplace_back([&, d] {
                try { f(d); } catch (const std::exception& e) { err[d] = e.what(); }
            });
        for (auto& t : th) t.join();
        for (int d = 0; d < n; d++)
            if (!err[d].empty()) throw std::runtime_error("device " + std::to_string(d) + ": " + err[d]);
    }
    /* one frame: every device renders its rows, device 0 assembles; returns the slowest device's kernel time */
    rt_stats render(const rt_params& p, const float L[3], float intensity, uint8_t* host_rgb) {
        std::vector<rt_stats> st(n);
        each([&](int d) {
            rt_params q = p;
            rt_shard_rows(H, d, n, G, &q);
            rtb200::check(rt_scene_set_light(scenes[d], L, intensity));
            if (q.row_count > 0) rtb200::check(rt_render(scenes[d], &q, 0, (uint8_t*)bands[d], nullptr, nullptr, nullptr, nullptr, &st[d]));
            rtb200::check(rt_gather_framebuffer_groups(scenes[d], comms[d], bands[d], W, H, 3, G, d == 0 ? frame : nullptr, 0));
            if (d == 0) rtb200::check(rt_scene_push_rows(scenes[0], frame, host_rgb, W, 3, 0, 1, H)); /* D2H on the scene's stream */
            rtb200::check(rt_scene_sync(scenes[d], nullptr));
        });
        rt_stats out = st[0];
        for (int d = 1; d < n; d++) {
            out.kernel_ms = std::max(out.kernel_ms, st[d].kernel_ms);
            out.rays += st[d].rays;
            out.launches += st[d].launches;
        }
        return out;
    }
    void close() {
        for (int d = 0; d < n; d++) {
            if (bands[d]) rt_peer_free(d, bands[d]);
            if (comms[d]) rt_comm_destroy(comms[d]);
            if (d > 0 && scenes[d]) rt_scene_destroy(scenes[d]);
        }
        if (frame) rt_peer_free(0, frame);
    }
};

int main(int argc, char** argv) {
    std::vector<std::string> pos;
    std::string profile = "optimized", obj = "cadnav.com_model/Models_F0202A090/cat.obj", out;
    int W = 512, H = 512, device = 0, frames = 1, gpus = 1, row_group = 4;
    bool stochastic = false, gpu_build = false, mirror = false;
    std::string pattern;
    float orbit = 0.f;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char* { return (i + 1 < argc) ? argv[++i] : ""; };
        if (a == "--profile") profile = next();
        else if (a == "--width") W = atoi(next());
        else if (a == "--height") H = atoi(next());
        else if (a == "--obj") obj = next();
        else if (a == "--out") out = next();
        else if (a == "--device") device = atoi(next());
        else if (a == "--frames") frames = atoi(next());
        else if (a == "--stochastic") stochastic = true;
        else if (a == "--gpu-build") gpu_build = true;
        else if (a == "--gpus") gpus = atoi(next());
        else if (a == "--row-group") row_group = atoi(next());
        else if (a == "--mirror") mirror = true;
        else if (a == "--out-pattern") pattern = next();
        else if (a == "--orbit") orbit = (float)atof(next());
        else pos.push_back(a);
    }
    if (pos.size() != 2) {
        std::cout << "Invalid number of arguments!\nThe first argument is number of rays and the second argument is number of bounces.\n";
        return 0;
    }
    if (!pattern.empty() && !pattern_ok(pattern)) {
        std::cerr << "rt_render: --out-pattern must contain exactly one %d (or %0Nd) and no other conversion\n";
        return 2;
    }
    if (row_group < 1 || row_group > 64 || (row_group & (row_group - 1)) != 0) {
        std::cerr << "rt_render: --row-group must be a power of two <= 64\n";
        return 2;
    }
    if (gpus < 1 || (gpus > 1 && device != 0)) {
        std::cerr << "rt_render: --gpus N uses devices 0..N-1\n";
        return 2;
    }
    if (!stochastic)
        std::cerr << "rt_render: deterministic mode (sigma 0, no indirect bounce); the reference's `./optimized R B` always jitters and bounces: pass --stochastic for that\n";
    auto start_time = std::chrono::system_clock::now();
    const int num_rays = atoi(pos[0].c_str()), num_bounce = atoi(pos[1].c_str());
    if (out.empty()) out = (profile == "cpu") ? "image.png" : "image_optimized.png";
    try {
        rt_params p;
        rtb200::check(rt_params_profile(&p, profile.c_str(), W, H, num_rays, num_bounce));
        if (stochastic) {
            p.aa_sigma = 0.2f; /* optimized.cu:753 */
            p.indirect = 1;    /* optimized.cu:631-649 */
        }
        /* optimized.cu evaluates z inside the kernel (:748-749, CUDA's tanf); cpu_launcher.cpp and array_bvh.cu on the host */
        if (profile == "optimized") rtb200::check(rt_camera_z_device(device, W, (float)(3.14159265358979323846 / 3), &p.z));
        rt_sphere walls[6];
        int32_t mesh_id = 0;
        rtb200::check(rt_default_walls(profile.c_str(), walls, &mesh_id));

        rtb200::TriangleMeshHost mesh; /* cat */
        mesh.mirror = mirror;
        mesh.readOBJ(obj.c_str());
        if (profile == "optimized") mesh.rescale(0.6f, rtb200::Vector(0.f, -4.f, 0.f));       /* optimized.cu:804 */
        else if (profile == "array_bvh") mesh.rescale(0.6f, rtb200::Vector(0.f, -10.f, 0.f)); /* array_bvh.cu:1033 */
        if (gpu_build) mesh.buildBVHDevice(device);
        else mesh.buildBVH();

        rtb200::Scene scene(device);
        for (int k = 0, id = 0; k < 6; k++, id++) {
            if (id == mesh_id) {
                scene.addObject(mesh);
                id++;
            }
            const rt_sphere& w = walls[k];
            scene.addObject(rtb200::Sphere(rtb200::Vector(w.C[0], w.C[1], w.C[2]), w.R, rtb200::Vector(w.albedo[0], w.albedo[1], w.albedo[2])));
        }
        if (mesh_id >= 6) scene.addObject(mesh);

        std::vector<uint8_t> image((size_t)W * H * 3);
        rt_stats st{};
        rt_png_writer* writer = nullptr;
        if (!pattern.empty()) rtb200::check(rt_png_writer_create(&writer, 0, 0));
        float L[3] = {scene.L.x, scene.L.y, scene.L.z};
        double kernel_ms_sum = 0.;
        MultiGpu multi;
        if (gpus > 1) {
            scene.flush(); /* spheres + light onto device 0 before the blob is broadcast */
            multi.G = row_group;
            multi.init(gpus, scene.handle(), W, H);
        }
        for (int f = 0; f < frames; f++) {
            if (f > 0 && orbit != 0.f) {
                rt_move_light(L, orbit, 0.02f);
                scene.setLight(rtb200::Vector(L[0], L[1], L[2]), scene.intensity);
            }
            st = gpus > 1 ? multi.render(p, L, scene.intensity, image.data()) : scene.render(p, image.data());
            kernel_ms_sum += st.kernel_ms;
            if (writer) {
                char name[1024];
                snprintf(name, sizeof name, pattern.c_str(), f);
                rtb200::check(rt_png_writer_submit(writer, name, W, H, image.data())); /* copies; encode + write overlap the next frame */
            }
        }
        if (writer) {
            const int rc = rt_png_writer_wait(writer);
            rt_png_writer_destroy(writer);
            rtb200::check(rc);
        } else {
            rtb200::check(rt_write_png(out.c_str(), W, H, image.data()));
        }
        if (gpus > 1) multi.close();
        if (frames > 1) std::cerr << frames << " frames, mean kernel " << (kernel_ms_sum / frames) << " ms/frame\n";
        auto end_time = std::chrono::system_clock::now();
        std::chrono::duration<float> run_time = end_time - start_time;
        std::cout << "Rendering time: " << run_time.count() << " s\n";
        std::cerr << "kernel " << st.kernel_ms << " ms/frame, " << st.rays << " rays, " << (st.rays / (st.kernel_ms * 1e3)) << " Mrays/s\n";
    } catch (const std::exception& e) {
        std::cerr << "rt_render: " << e.what() << "\n";
        return 1;
    }
    return 0;
}
