#!/usr/bin/env python
"""bench.py — Mrays/s and ms/frame of the render hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path on the host cores

Workload (N = 1): BASELINE.json configs[1] — cadnav cat TriangleMesh with the array BVH + six wall spheres,
1920x1080, 1 sample/pixel, primary + shadow rays (4,147,200 rays/frame), optimized.cu knobs. A "step" is one
frame. For N > 1 the path shards by frame: every rank renders its own frames of the SAME workload (frame-parallel,
no data-path collective, "scaling": "weak"), so that the per-N values are comparable; the frames of a light-orbit
animation (BASELINE.json configs[3] style: the light's candidate bins are rebuilt every frame) and whole 4K depth-4
frames rendered frame-parallel are reported beside the headline (`animation_light_orbit`, `frames_4k_depth4`).

`value` times the render kernels with the scene resident in HBM (CUDA events on the launching stream, L2
flushed between steps). `e2e` goes through the C ABI with HOST buffers: every step re-uploads the mesh in the
reference interchange formats (H2D), renders, and copies the 8-bit frame back to pinned host memory (D2H).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CAT_REL = os.path.join("cadnav.com_model", "Models_F0202A090", "cat.obj")
W, H = 1920, 1080
METRIC = "Mrays/s"


def find_cat():
    for base in (os.environ.get("RT_REFERENCE_DIR", "/root/reference"), os.path.join(ROOT, "oracle", "_ref")):
        p = os.path.join(base, CAT_REL)
        if os.path.exists(p):
            return p
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes of ONE frame (all its render launches: wf_generate / wf_leaves / wf_shade per strip) from the newest
    committed `ncu --set full` summary of this workload (profiles/*_ncu_summary.json, written by tools/ncu_summary.py, one
    captured frame per file); None when no capture is committed. The render path has no single dominant kernel any more
    (generate 45 %, leaves 40 %, shade 15 % of a frame), so the roofline is stated for the frame."""
    import glob
    files = [f for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_summary.json"))) if "config5" not in f]
    if not files:
        return None, None
    d = json.load(open(files[-1]))
    vals = [l["dram_traffic_B"] for l in d["launches"] if l["kernel"].startswith("wf_")]
    return (int(sum(vals)) if vals else None), os.path.basename(files[-1])


def build_scene_host(rt):
    """Host side of the launcher (optimized.cu:801-813): load, rescale, build the BVH — with the product's host code."""
    cat = find_cat()
    if cat:
        mesh = rt.Mesh.read_obj(cat).rescale(0.6, (0.0, -4.0, 0.0)).build_bvh()
        name = "cadnav cat (3954 tris, 2019 BVH nodes)"
    else:  # the asset is not redistributable; without it a synthetic mesh of similar size keeps the bench runnable
        sys.path.insert(0, ROOT)
        from oracle import scenes
        v, t = scenes.torus(64, 31)
        mesh = rt.Mesh.from_arrays(v, t).build_bvh()
        name = "synthetic torus (cat.obj unavailable)"
    walls, mesh_id = rt.default_walls("optimized")
    return mesh, walls, mesh_id, name


def run_ours(args):
    import torch
    import raytracinggpu_b200 as rt

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if rt.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; raytracinggpu_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sc = rt.Scene(local)
    stream = torch.cuda.Stream()
    sc.set_stream(stream.cuda_stream)
    mesh, walls, mesh_id, mesh_name = build_scene_host(rt)  # every rank keeps the host arrays for its own e2e leg
    verts, recs, bvh = mesh.vertices, mesh.tri_records, mesh.arr_bvh
    blob_bytes = 0
    if world > 1:
        # the device scene is built and packed on rank 0 only and broadcast once
        from raytracinggpu_b200 import distributed as rtd
        if rank == 0:
            sc.set_spheres(walls)
            sc.set_mesh(verts, recs, bvh, id=mesh_id)
        blob_bytes = rtd.broadcast_scene(sc, src=0)
    else:
        sc.set_spheres(walls)
        sc.set_mesh(verts, recs, bvh, id=mesh_id)
    p = rt.params_profile("optimized", W, H, 1, 1)

    # frame-parallel for N > 1: rank r renders frames r, r+N, ... of a sequence whose frames are all configs[1] (static scene):
    # the same per-GPU work at every N. The light-orbit animation is timed separately below.
    n_frames = (args.warmup + args.steps) * world
    omega = 2 * np.pi / (240 * 0.02)
    lights = [(-10.0, 20.0, 40.0)] * n_frames

    rgb = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]

    def one_step(i, timed_idx=None):
        sc.set_light(lights[i * world + rank], 3e10)
        with torch.cuda.stream(stream):
            flush.zero_()  # L2 flush, outside the event pair
            if timed_idx is not None:
                ev[timed_idx][0].record(stream)
            sc.render_into(p, rgb=rgb, flags=rt.RT_RENDER_NO_SYNC)
            if timed_idx is not None:
                ev[timed_idx][1].record(stream)

    for i in range(args.warmup):
        one_step(i)
    stats0 = sc.sync()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    t0 = time.perf_counter()
    for i in range(args.steps):
        one_step(args.warmup + i, i)
    st = sc.sync()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall_s = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    rays_per_frame = int(st.rays)
    launches_per_step = int(st.launches)
    ms_list = [a.elapsed_time(b) for a, b in ev]
    ms_sum = float(sum(ms_list))
    if world > 1:
        t = torch.tensor([ms_sum], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_sum = float(t.item())
        r = torch.tensor([rays_per_frame * args.steps], device="cuda", dtype=torch.int64)
        dist.all_reduce(r)
        total_rays = int(r.item())
    else:
        total_rays = rays_per_frame * args.steps
    ms_per_step = ms_sum / args.steps
    value = total_rays / (ms_sum * 1e-3) / 1e6

    # ---- e2e through the C ABI with host buffers --------------------------------------------------------
    host_rgb = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()
    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(2):
        sc.set_mesh(verts, recs, bvh, id=mesh_id)
        sc.render_into(p, rgb=host_rgb.numpy())
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        sc.set_light(lights[(args.warmup + i % args.steps) * world + rank], 3e10)
        sc.set_mesh(verts, recs, bvh, id=mesh_id)          # H2D of the interchange arrays + device repack
        e2e_st = sc.render_into(p, rgb=host_rgb.numpy())    # render + D2H of the frame, synchronous
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = rays_per_frame * e2e_steps * world / e2e_s / 1e6
    h2d = int(verts.nbytes + recs.nbytes + bvh.nbytes)
    d2h = int(H * W * 3)

    # ---- extra lines: frames of a one-revolution light orbit (the light's bins are rebuilt every frame) and
    # whole 4K depth-4 frames, both frame-parallel over the ranks, device-timed, max over ranks
    def frame_parallel(params, frames, light_of):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(frames)]
        buf = torch.empty((params.H, params.W, 3), dtype=torch.uint8, device="cuda")
        rays_f = 0
        for i in range(3 + frames):
            sc.set_light(light_of(i), 3e10)
            with torch.cuda.stream(stream):
                flush.zero_()
                if i >= 3:
                    evs[i - 3][0].record(stream)
                sc.render_into(params, rgb=buf, flags=rt.RT_RENDER_NO_SYNC)
                if i >= 3:
                    evs[i - 3][1].record(stream)
            if i == 2:
                sc.sync()  # end of the warm-up: lets the library enlarge buffers the first frames found too small
        rays_f = int(sc.sync().rays)
        torch.cuda.synchronize()
        ms = float(sum(a.elapsed_time(b) for a, b in evs))
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return {"ms_per_frame_per_gpu": round(ms / frames, 4), "rays_per_frame": rays_f, "frames_per_gpu": frames,
                "mrays_per_s": round(rays_f * frames * world / (ms * 1e-3) / 1e6, 1), "scaling": "weak"}

    orbit = rt.sharding.light_positions((-10.0, 20.0, 40.0), (3 + 20) * world, omega, 0.02, rt.move_light)
    animation = frame_parallel(p, 20, lambda i: orbit[i * world + rank])
    animation["workload"] = "configs[1] scene, light on a one-revolution orbit, frame f on rank f mod N (light bins rebuilt every frame)"
    sc.set_light((-10.0, 20.0, 40.0), 3e10)

    # ---- extra line: the reference's BVH builder, host vs device, on the bench mesh
    bvh_build = None
    if rank == 0:
        cat = find_cat()
        if cat:
            def fresh():
                return rt.Mesh.read_obj(cat).rescale(0.6, (0.0, -4.0, 0.0))
            fresh().build_bvh_gpu(local)  # warm-up
            t0 = time.perf_counter(); fresh_h = fresh(); t1 = time.perf_counter(); fresh_h.build_bvh(); t2 = time.perf_counter()
            g = fresh(); t3 = time.perf_counter(); g.build_bvh_gpu(local); t4 = time.perf_counter()
            bvh_build = {"mesh": mesh_name, "host_builder_ms": round((t2 - t1) * 1e3, 3), "device_builder_wall_ms": round((t4 - t3) * 1e3, 3), "device_build_ms": round(g.build_ms, 3),
                         "identical": bool(np.array_equal(fresh_h.arr_bvh, g.arr_bvh) and np.array_equal(fresh_h.tri_records, g.tri_records)),
                         "note": "10 M triangles (configs[4] mesh): 2234 ms host, 36.9 ms on the device, see profiles/r01_notes.md"}

    work = sc.render(p, want=("rgb",), count_work=True)["stats"]  # instrumented pass for the roofline, not timed
    like_for_like = stochastic_vs_reference_kernel(rt, torch, sc) if (world == 1 and rank == 0) else None
    sharded = sharded_single_frame(rt, torch, sc, stream, world, rank, verts, recs, bvh, mesh_id, walls)
    # after sharded_single_frame every rank holds the mirror-cat scene: whole 4K depth-4 frames, one per rank at a time
    frames_4k = frame_parallel(rt.params_profile("optimized", 3840, 2160, 1, 4), 6, lambda i: (-10.0, 20.0, 40.0))
    frames_4k["workload"] = "BASELINE.json configs[2] frames (mirror cat 3840x2160, reflection depth 4), whole frames, frame-parallel over the ranks"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (render) ---------------------------------------------------------
    n_mesh_queries = rays_per_frame  # every ray tests the mesh root box once
    alg_bytes = 32 * (n_mesh_queries + 2 * work["node_visits"]) + 48 * work["tri_tests"] + H * W * 3
    alg_flop = 150 * rays_per_frame + 19 * (n_mesh_queries + 2 * work["node_visits"]) + 50 * work["tri_tests"]
    kernel_ms = float(np.mean(ms_list))
    peak, peak_src = measured_peaks()
    traffic, traffic_src = ncu_traffic()
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": traffic, "traffic_source": "dram__bytes_read+write summed over the render launches of one frame, profiles/%s" % traffic_src if traffic_src else None, "peak_source": peak_src,
                "note": "per FRAME (generate + leaves + shade launches; no single dominant kernel). Algorithmic bytes = SURVEY.md 8d: 32 B per box test and 48 B per triangle test of "
                        "the reference's own traversal (node visits / triangle tests counted by the instrumented tree search). The anchored-ray bins find the same leaves with a "
                        "quarter of the box tests, so the achieved figure is work the reference algorithm defines divided by the time this path needs (it can exceed the HBM peak: most of those bytes are never moved; measured DRAM traffic is `traffic`); the scene (0.4 MB) is "
                        "L1/L2-resident and the binding limits are instruction issue (generate) and LSU wavefronts (leaves), see profiles/ and fp32",
                "algorithmic_bytes_per_launch": int(alg_bytes), "node_visits": int(work["node_visits"]), "tri_tests": int(work["tri_tests"]),
                "fp32": {"achieved_tflops": round(alg_flop / (kernel_ms * 1e-3) / 1e12, 3), "peak_tflops": round(148 * 128 * 2 * 1.965e9 / 1e12, 1),
                         "frac": round(alg_flop / (kernel_ms * 1e-3) / (148 * 128 * 2 * 1.965e9), 4)}}

    # ---- CPU baseline: the reference's own classes on this box's host cores (bounded sample) ---------------
    # N = 1 only; torchrun exports OMP_NUM_THREADS=1, so the thread count is passed explicitly
    cpu_baseline = cpu_reference_sample(budget_s=12.0, threads=os.cpu_count() or 1) if world == 1 else None

    out = {"metric": METRIC, "value": round(value, 2), "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": round(ms_per_step, 5), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic",
           "config": {"workload": "BASELINE.json configs[1]: %s + 6 wall spheres, 1920x1080, 1 spp, primary+shadow rays, optimized.cu knobs" % mesh_name,
                      "rays_per_frame": rays_per_frame, "frames_per_step_per_gpu": 1, "sharding": "frame-parallel (every rank renders its own frames of this workload)" if world > 1 else "single GPU",
                      "l2": "256 MB memset between steps, outside the event pair", "timed_wall_s": round(wall_s, 4)},
           "e2e": {"value": round(e2e_value, 2), "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": round(e2e_s / e2e_steps * 1e3, 4),
                   "steps": e2e_steps},
           "gpu_launches": launches_per_step * args.steps, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
           "ms_per_frame": round(kernel_ms, 5), "scene_broadcast_bytes": blob_bytes, "single_frame_sharded": sharded,
           "animation_light_orbit": animation, "frames_4k_depth4": frames_4k, "bvh_build": bvh_build, "stochastic_vs_reference_gpu_kernel": like_for_like}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def stochastic_vs_reference_kernel(rt, torch, sc):
    """Like for like with the reference's GPU program: `./optimized 1 1` at 1920x1080 is the STOCHASTIC mode (sigma 0.2
    jitter + cuRAND stream). Times this library's stochastic mode (CUDA events inside rt_render) next to the unmodified
    optimized.cu kernel (oracle/_ref/ref_optimized: reference flags retargeted to sm_100a, CUDA events around its
    KernelLaunch) on this GPU. Extra information beside the headline; None when the compiled reference did not travel."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_optimized")
    cat = find_cat()
    if not (cat and os.path.exists(exe)):
        return None
    out = {}
    for rays, bounce in ((1, 1), (4, 3)):
        p = rt.params_profile("optimized", W, H, rays, bounce)
        p.aa_sigma, p.indirect = 0.2, 1
        rgb = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
        ms = []
        for i in range(8):
            st = sc.render_into(p, rgb=rgb)
            if i >= 3:
                ms.append(st.kernel_ms)
        try:
            r = subprocess.run([exe, cat, str(W), str(H), str(rays), str(bounce), "5"], capture_output=True, text=True, timeout=120)
            ref_ms = json.loads(r.stdout.strip().splitlines()[-1])["kernel_ms_median"]
        except Exception as e:  # noqa: BLE001
            ref_ms = None
        ours = float(np.median(ms))
        out["optimized_%d_%d" % (rays, bounce)] = {"this_ms": round(ours, 4), "reference_kernel_ms": ref_ms, "rays_per_frame": int(st.rays),
                                                   "speedup": round(ref_ms / ours, 2) if ref_ms else None}
    return out


def sharded_single_frame(rt, torch, sc, stream, world, rank, verts, recs, bvh, mesh_id, walls, frames=8):
    """BASELINE.json configs[2]: ONE 3840x2160 frame, mirror cat (reflection depth 4), rows interleaved over the ranks,
    bands all-gathered (NCCL) and de-interleaved on the device. Strong scaling of a single frame; timed with CUDA
    events around render + gather, max over ranks. Reported beside the headline, not as `value`."""
    import torch.distributed as dist
    from raytracinggpu_b200 import distributed as rtd
    W4, H4 = 3840, 2160
    if world > 1:
        if rank == 0:
            sc.set_mesh(verts, recs, bvh, mirror=1, id=mesh_id)
        rtd.broadcast_scene(sc, src=0)
    else:
        sc.set_mesh(verts, recs, bvh, mirror=1, id=mesh_id)
    sc.set_light((-10.0, 20.0, 40.0), 3e10)
    fg = rtd.FrameGather(H4, W4, world, rank, torch.device("cuda", torch.cuda.current_device()))
    p = fg.apply(rt.params_profile("optimized", W4, H4, 1, 4))
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(frames)]
    rays = 0
    for i in range(3 + frames):
        if world > 1:
            dist.barrier()
        with torch.cuda.stream(stream):
            k = i - 3
            if k >= 0:
                ev[k][0].record(stream)
            sc.render_into(p, rgb=fg.band, flags=rt.RT_RENDER_NO_SYNC)
            if k >= 0:
                ev[k][1].record(stream)
            fg.gather()
            if k >= 0:
                ev[k][2].record(stream)
        rays = int(sc.sync().rays)
    torch.cuda.synchronize()
    t = torch.tensor([sum(a.elapsed_time(c) for a, _, c in ev), sum(a.elapsed_time(b) for a, b, _ in ev)], device="cuda", dtype=torch.float64)
    r = torch.tensor([rays], device="cuda", dtype=torch.int64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(r)
    total_ms, render_ms = float(t[0].item()) / frames, float(t[1].item()) / frames
    # the same frame with the bands pushed into rank 0's frame buffer over NVLink (CUDA IPC peer copies + one barrier)
    push = None
    if world > 1:
        fp = rtd.FramePush(sc, H4, W4, world, rank, torch.device("cuda", torch.cuda.current_device()))
        pp = fp.apply(rt.params_profile("optimized", W4, H4, 1, 4))
        ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(frames)]
        for i in range(3 + frames):
            dist.barrier()
            with torch.cuda.stream(stream):
                k = i - 3
                if k >= 0:
                    ev2[k][0].record(stream)
                sc.render_into(pp, rgb=fp.band, flags=rt.RT_RENDER_NO_SYNC)
                if k >= 0:
                    ev2[k][1].record(stream)
                fp.push()
                if k >= 0:
                    ev2[k][2].record(stream)
            sc.sync()
        torch.cuda.synchronize()
        t2 = torch.tensor([sum(a.elapsed_time(c) for a, _, c in ev2), sum(a.elapsed_time(b) for a, b, _ in ev2)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        same = None
        if rank == 0:
            with torch.cuda.stream(stream):
                pushed = fp.frame_tensor()
            sc.sync()
            torch.cuda.synchronize()
            same = bool(torch.equal(pushed, fg.frame))
        fp.close()
        push = {"ms_per_frame": round(float(t2[0].item()) / frames, 4), "render_ms": round(float(t2[1].item()) / frames, 4),
                "push_and_barrier_ms": round(float(t2[0].item() - t2[1].item()) / frames, 4), "frame_equals_all_gather": same,
                "how": "rank 0 owns the frame (CUDA IPC handle broadcast once); every rank copies its band into it over NVLink, one barrier per frame"}
    return {"workload": "BASELINE.json configs[2]: mirror cat 3840x2160, reflection depth 4, rows interleaved over %d GPU(s), all-gather to every rank" % world,
            "p2p_push": push,
            "rays_per_frame": int(r.item()), "ms_per_frame": round(total_ms, 4), "render_ms": round(render_ms, 4), "gather_ms": round(total_ms - render_ms, 4),
            "mrays_per_s": round(int(r.item()) / (total_ms * 1e-3) / 1e6, 1), "gather_bytes_per_rank": int(fg.band.numel()), "frames": frames, "scaling": "strong"}


def cpu_reference_sample(budget_s=12.0, threads=0):
    """Time the reference's CPU render loop (oracle/_ref/libref_cpu.so = cpu_launcher.cpp's classes, OpenMP
    schedule(dynamic,1) over rows) on a bounded sample of the same workload; falls back to the oracle port."""
    from oracle import profiles, pyoracle, scenes
    cores = os.cpu_count() or 1
    cat = find_cat()
    rays = 2 * W * H
    if cat and pyoracle.ref_cpu_available():
        # scene_kind 1 = optimized.cu object order + mesh transform; num_bounce 0 in the recursive CPU code = one
        # segment + its shadow ray = the GPU program's num_bounce 1
        pyoracle.ref_cpu_render(cat, 1, 480, 270, 1, 0, threads, hits=False)  # warm-up
        n, spent = 0, 0.0
        while spent < budget_s and n < 30:
            r = pyoracle.ref_cpu_render(cat, 1, W, H, 1, 0, threads, hits=False)
            spent += r["seconds"]
            n += 1
        return {"value": round(rays * n / spent / 1e6, 3), "unit": "Mrays/s", "cores": cores if threads <= 0 else threads, "kind": "reference",
                "sample": "%d full 1920x1080 frames (%.2f s) by cpu_launcher.cpp's classes, render loop only" % (n, spent), "ms_per_frame": round(spent / n * 1e3, 2)}
    desc = scenes.cat_scene("optimized") or scenes.torus_scene("optimized")
    p = profiles.params("optimized", W, H, 1, 1)
    n, spent = 0, 0.0
    while spent < budget_s and n < 30:
        o = scenes.run_oracle(desc, p, threads=threads, want=("rgb",))
        spent += o["work"]["seconds"]
        n += 1
    return {"value": round(o["work"]["rays"] * n / spent / 1e6, 3), "unit": "Mrays/s", "cores": o["work"]["threads"], "kind": "port",
            "sample": "%d full 1920x1080 frames (%.2f s) by the oracle port" % (n, spent), "ms_per_frame": round(spent / n * 1e3, 2)}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores, same config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle
    cat = find_cat()
    cores = os.cpu_count() or 1
    rays = 2 * W * H
    use_ref = bool(cat and pyoracle.ref_cpu_available())
    if use_ref:
        def frame():
            return pyoracle.ref_cpu_render(cat, 1, W, H, 1, 0, cores, hits=False)["seconds"]
        kind = "reference"
    else:
        from oracle import profiles, scenes
        desc = scenes.cat_scene("optimized") or scenes.torus_scene("optimized")
        p = profiles.params("optimized", W, H, 1, 1)

        def frame():
            return scenes.run_oracle(desc, p, threads=cores, want=("rgb",))["work"]["seconds"]
        kind = "port"
    for _ in range(min(args.warmup, 3)):
        frame()
    t0 = time.perf_counter()
    times = [frame() for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    total = float(sum(times))
    value = rays * args.steps / total / 1e6
    out = {"impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": round(total / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic",
           "config": {"workload": "BASELINE.json configs[1]: cat + 6 wall spheres, 1920x1080, 1 spp, primary+shadow rays; each step = one full frame on the host CPU",
                      "rays_per_frame": rays, "wall_s": round(wall, 3)},
           "cpu_baseline": {"value": round(value, 3), "unit": "Mrays/s", "cores": cores, "kind": kind,
                            "sample": "%d full 1920x1080 frames, render loop only (cpu_launcher.cpp:695-718), OpenMP over rows" % args.steps},
           "e2e": {"value": round(value, 3), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        if args.steps > 40:
            args.steps = 40
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
