#!/usr/bin/env python
"""bench.py — Mrays/s and ms/frame of the render hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path on the host cores

Workload (N = 1): BASELINE.json configs[1] — cadnav cat TriangleMesh with the array BVH + six wall spheres,
1920x1080, 1 sample/pixel, primary + shadow rays (4,147,200 rays/frame), optimized.cu knobs. A "step" is one
frame. For N > 1 the path shards by frame: every rank renders its own frames of the SAME workload (frame-parallel,
no data-path collective, "scaling": "weak"), so that the per-N values are comparable. The other BASELINE.json configs
are reported beside the headline under "configs": [0] spheres scene 800x600, [2] one 4K depth-4 frame, groups of 4 rows interleaved
over the ranks (strong scaling, NCCL all-gather and NVLink push), [3] the 240-frame light-orbit animation of the spheres
scene at 1080p frame-parallel over the ranks, [4] the 10 M-triangle scene at 4K built on rank 0, broadcast once and
rendered the same way; plus a light-orbit animation of the cat scene and whole 4K depth-4 frames.

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
ROW_GROUP = int(os.environ.get("RT_ROW_GROUP", "4"))  # single frames over N ranks: groups of 4 consecutive rows dealt out in turn (rt_params.row_group)


def workload_string(mesh_name):
    """The same string in both arms (the driver compares them)."""
    return "BASELINE.json configs[1]: %s + 6 wall spheres, 1920x1080, 1 spp, primary+shadow rays, optimized.cu knobs" % mesh_name


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


def ncu_frame_summary(pattern="*_ncu_summary.json", exclude=("config5", "config4")):
    """Per-FRAME figures of the newest committed `ncu --set full` summary of the headline workload (profiles/*_ncu_summary.json,
    written by tools/ncu_summary.py, one captured frame per file): DRAM bytes summed over the frame's render launches, the
    serialised duration of those launches and the L2 throughput share. None when no capture is committed."""
    import glob
    files = [f for f in sorted(glob.glob(os.path.join(ROOT, "profiles", pattern))) if not any(x in f for x in exclude)]
    if not files:
        return None
    d = json.load(open(files[-1]))
    L = [l for l in d["launches"] if l["kernel"].split("<")[0].replace("rtk::", "").startswith(("wf_", "bins_", "accumulate"))]
    if not L:
        return None
    dur = sum(l["duration_us"] or 0 for l in L)
    return {"file": os.path.basename(files[-1]), "dram_bytes": int(sum(l["dram_traffic_B"] or 0 for l in L)), "serialised_us": round(dur, 1), "launches": len(L),
            "l2_throughput_pct_time_weighted": round(sum((l["l2_throughput_pct"] or 0) * (l["duration_us"] or 0) for l in L) / dur, 2) if dur else None,
            "kernel_share": {k: round(sum(l["duration_us"] or 0 for l in L if l["kernel"].split("<")[0].endswith(k)) / dur, 3) for k in ("wf_generate", "wf_leaves", "wf_shade")} if dur else None}


def pin_to_gpu_numa(local):
    """N > 1: bind this rank's host threads (and with them the first-touch placement of its pinned buffers) to the CPUs of its GPU's NUMA
    node: the e2e leg moves 6.5 MB per step between host and device on every rank at once. Best effort; returns what was done."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        path = "/sys/bus/pci/devices/%s:%s/local_cpulist" % (dom[-4:].lower(), rest.lower())
        cpus = set()
        for part in open(path).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "%d CPUs of the GPU's NUMA node" % len(cpus)
    except Exception as e:  # noqa: BLE001
        return "not pinned (%s)" % repr(e)[:80]
    return "not pinned"


def build_scene_host(rt):
    """Host side of the launcher (optimized.cu:801-813): load, rescale, build the BVH — with the product's host code."""
    cat = find_cat()
    if cat:
        mesh = rt.Mesh.read_obj(cat).rescale(0.6, (0.0, -4.0, 0.0)).build_bvh()
        name = "cadnav cat (3954 tris, 2019 BVH nodes)"
    else:  # the asset is not redistributable; without it a synthetic mesh of similar size keeps the bench runnable
        from raytracinggpu_b200 import synthetic
        v, t = synthetic.torus(64, 31)
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
    numa = None
    if world > 1:
        import torch.distributed as dist
        numa = pin_to_gpu_numa(local)
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
    # Every step uploads the interchange arrays (rt_scene_set_mesh from host memory: H2D + device repack) and renders into a pinned HOST
    # frame. The steps are enqueued back to back (RT_RENDER_NO_SYNC): the library copies frame k to the host on its copy stream while frame
    # k + 1 renders (two scratch sets), the uploads alternate between two pinned halves, and the one rt_scene_sync that ends the timed
    # region waits for every kernel and every copy. Two host frames alternate, as a consumer that reads frame k during step k + 1 needs.
    host_frames = [torch.empty((H, W, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
    host_rgb = host_frames[0]
    e2e_steps = max(3, min(args.steps, 50))

    def e2e_run(steps, pipelined):
        t0 = time.perf_counter()
        for i in range(steps):
            sc.set_light(lights[(args.warmup + i % args.steps) * world + rank], 3e10)
            sc.set_mesh(verts, recs, bvh, id=mesh_id)          # H2D of the interchange arrays + device repack
            if pipelined:
                sc.render_into(p, rgb=host_frames[i & 1].numpy(), flags=rt.RT_RENDER_NO_SYNC)
            else:
                sc.render_into(p, rgb=host_rgb.numpy())          # render + D2H of the frame, synchronous
        st_ = sc.sync() if pipelined else None
        return time.perf_counter() - t0, st_

    for _ in range(2):
        e2e_run(4, True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_s, e2e_st = e2e_run(e2e_steps, True)
    e2e_s = _max_over_ranks(torch, world, e2e_s)
    # the frames that reached the host are the frame the device-timed loop rendered
    assert torch.equal(host_frames[0], rgb.cpu()) and torch.equal(host_frames[1], rgb.cpu()), "e2e frames differ from the device-timed frame"
    e2e_run(2, False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_sync_s = _max_over_ranks(torch, world, e2e_run(e2e_steps, False)[0])
    e2e_value = rays_per_frame * e2e_steps * world / e2e_s / 1e6
    h2d = int(verts.nbytes + recs.nbytes + bvh.nbytes)
    d2h = int(H * W * 3)
    # where the step goes, with every rank doing the same at the same time (max over ranks): the mesh upload alone, the frame's D2H alone
    def timed(fn, n=10):
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return _max_over_ranks(torch, world, (time.perf_counter() - t0) / n * 1e3)
    e2e_breakdown = {"synchronous_step_ms": round(e2e_sync_s / e2e_steps * 1e3, 4),
                     "set_mesh_ms": round(timed(lambda: (sc.set_mesh(verts, recs, bvh, id=mesh_id), sc.sync())), 4),
                     "d2h_6MB_alone_ms": round(timed(lambda: host_rgb.copy_(rgb)), 4),
                     "render_to_device_buffer_sync_ms": round(timed(lambda: sc.render_into(p, rgb=rgb)), 4),
                     "note": "all ranks at once, max over ranks; on an 8-GPU board two GPUs share a PCIe switch uplink, which is what the D2H line shows at N = 8"}

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
    cfg0, cfg3 = spheres_configs(rt, torch, local, world, rank, flush)
    cfg4 = config4_ten_million(rt, torch, local, world, rank, flush)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (render) ---------------------------------------------------------
    n_mesh_queries = rays_per_frame  # every ray tests the mesh root box once
    alg_bytes = 32 * (n_mesh_queries + 2 * work["node_visits"]) + 48 * work["tri_tests"] + H * W * 3
    alg_flop = 150 * rays_per_frame + 19 * (n_mesh_queries + 2 * work["node_visits"]) + 50 * work["tri_tests"]
    kernel_ms = float(np.mean(ms_list))
    hbm_peak, hbm_src = measured_peaks()
    fma_peak = rt.fma_peak_tflops(local)  # measured on this box, in this run: 8 FMA chains per thread (rt_selftest_fma_peak)
    ncu = ncu_frame_summary()
    traffic = ncu["dram_bytes"] if ncu else None
    achieved_tf = alg_flop / (kernel_ms * 1e-3) / 1e12
    frame_bytes = H * W * 3
    roofline = {"bound": "fp32", "achieved": round(achieved_tf, 3), "peak": round(fma_peak, 2), "unit": "TFLOP/s", "frac": round(achieved_tf / fma_peak, 4),
                "traffic": traffic,
                "peak_source": "FMA micro-benchmark run by this bench (rt_selftest_fma_peak: 8 independent FFMA chains per thread, best of 5); nominal 148 SMs x 128 lanes x 2 x 1.965 GHz = 74.4",
                "note": "per FRAME (generate + leaves + shade launches of both row bands; no single dominant kernel). The scene (0.4 MB) is L1/L2-resident: the compulsory HBM traffic is "
                        "the 6.2 MB frame, so the binding roof is the FP32 pipe / instruction issue (SURVEY.md 8d). Algorithmic FLOP = SURVEY.md 8d's F_ray: 150 per ray (six sphere tests) "
                        "+ 19 per box test + 50 per triangle test of the reference's own traversal (node visits / triangle tests counted by the instrumented tree search).",
                "algorithmic_flop_per_frame": int(alg_flop), "algorithmic_bytes_per_frame": int(alg_bytes), "node_visits": int(work["node_visits"]), "tri_tests": int(work["tri_tests"]),
                "hbm": {"compulsory_bytes": frame_bytes, "measured_dram_bytes": traffic, "traffic_over_compulsory": round(traffic / frame_bytes, 2) if traffic else None,
                        "achieved_gbs": round(traffic / (kernel_ms * 1e-3) / 1e9, 1) if traffic else None, "peak_gbs": hbm_peak, "peak_source": hbm_src,
                        "frac": round(traffic / (kernel_ms * 1e-3) / 1e9 / hbm_peak, 4) if traffic else None,
                        "source": ("ncu --set full --cache-control none, dram__bytes_read+write summed over the render launches of one frame, profiles/%s" % ncu["file"]) if ncu else None},
                "ncu": ncu}

    # ---- CPU baseline: the reference's own classes on this box's host cores (bounded sample) ---------------
    # N = 1 only; torchrun exports OMP_NUM_THREADS=1, so the thread count is passed explicitly
    cpu_baseline = cpu_reference_sample(budget_s=12.0, threads=os.cpu_count() or 1) if world == 1 else None

    out = {"metric": METRIC, "value": round(value, 2), "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": round(ms_per_step, 5), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic",
           "config": {"workload": workload_string(mesh_name),
                      "rays_per_frame": rays_per_frame, "frames_per_step_per_gpu": 1, "sharding": "frame-parallel (every rank renders its own frames of this workload)" if world > 1 else "single GPU",
                      "l2": "256 MB memset between steps, outside the event pair", "timed_wall_s": round(wall_s, 4), "host_affinity": numa},
           "e2e": {"value": round(e2e_value, 2), "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": round(e2e_s / e2e_steps * 1e3, 4),
                   "steps": e2e_steps, "breakdown": e2e_breakdown},
           "gpu_launches": launches_per_step * args.steps, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
           "ms_per_frame": round(kernel_ms, 5), "scene_broadcast_bytes": blob_bytes,
           "configs": {"configs0_spheres_800x600": cfg0, "configs2_one_4k_depth4_frame_row_sharded": sharded, "configs3_spheres_animation_240_frames": cfg3,
                       "configs4_10M_triangles_4k_row_sharded": cfg4},
           "single_frame_sharded": sharded, "animation_light_orbit": animation, "frames_4k_depth4": frames_4k, "bvh_build": bvh_build,
           "stochastic_vs_reference_gpu_kernel": like_for_like}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def _timed_frames(rt, torch, sc, stream, flush, params, frames, light_of=None, warm=3):
    """`frames` frames of `params` on this rank, each between its own CUDA-event pair with an L2 flush before it (outside the pair).
    Returns (sum of ms, rays of the last frame, launches of the last frame)."""
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(frames)]
    rows = params.row_count if params.row_count > 0 else params.H
    buf = torch.empty((rows, params.W, 3), dtype=torch.uint8, device="cuda")
    for i in range(warm + frames):
        if light_of is not None:
            sc.set_light(light_of(i), 3e10)
        with torch.cuda.stream(stream):
            flush.zero_()
            if i >= warm:
                evs[i - warm][0].record(stream)
            sc.render_into(params, rgb=buf, flags=rt.RT_RENDER_NO_SYNC)
            if i >= warm:
                evs[i - warm][1].record(stream)
        if i == warm - 1:
            sc.sync()  # end of the warm-up: lets the library enlarge buffers the first frames found too small
    st = sc.sync()
    torch.cuda.synchronize()
    return float(sum(a.elapsed_time(b) for a, b in evs)), int(st.rays), int(st.launches)


def _max_over_ranks(torch, world, x):
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _sum_over_ranks(torch, world, x):
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], device="cuda", dtype=torch.int64)
    dist.all_reduce(t)
    return int(t.item())


def spheres_configs(rt, torch, local, world, rank, flush):
    """BASELINE.json configs[0] (spheres scene of cpu_launcher.cpp:668-678 with a point light, ONE 800x600 frame, 1 spp, cpu knobs)
    and configs[3] (the same scene at 1920x1080, light on a one-revolution orbit, 240 frames, frame f on rank f mod N). The scene has
    a mirror sphere and a refractive shell, so a path has up to 6 segments (num_bounce 5 in the recursive CPU program). Deterministic
    mode (sigma 0, no indirect bounce: what the parity tests pin); the CPU reference beside it always bounces (its own getColor)."""
    from raytracinggpu_b200 import synthetic
    sc = rt.Scene(local)
    stream = torch.cuda.Stream()
    sc.set_stream(stream.cuda_stream)
    sc.set_spheres(synthetic.spheres_scene_spheres(rt))
    sc.set_light((-10.0, 20.0, 40.0), 3e10)
    # configs[0]: every rank renders the same single frame (it does not shard: 0.48 Mpx); rank 0's numbers are reported
    p0 = rt.params_profile("cpu", 800, 600, 1, 5)
    ms0, rays0, l0 = _timed_frames(rt, torch, sc, stream, flush, p0, 20)
    host = torch.empty((600, 800, 3), dtype=torch.uint8).pin_memory()
    sc.render_into(p0, rgb=host.numpy())
    t0 = time.perf_counter()
    for _ in range(10):
        sc.render_into(p0, rgb=host.numpy())
    e2e0 = (time.perf_counter() - t0) / 10 * 1e3
    cfg0 = {"workload": "BASELINE.json configs[0]: spheres scene (6 walls + white / mirror / refractive-shell spheres) with point light, 800x600, 1 spp, cpu_launcher.cpp knobs, up to 6 path segments",
            "ms_per_frame": round(ms0 / 20, 5), "rays_per_frame": rays0, "mrays_per_s": round(rays0 * 20 / (ms0 * 1e-3) / 1e6, 1), "launches_per_frame": l0,
            "e2e_ms_per_frame_host_buffers": round(e2e0, 4), "n_gpus_used": 1}
    # configs[3]: 240 frames over the ranks
    total_frames = 240
    omega = 2 * np.pi / (total_frames * 0.02)
    orbit = rt.sharding.light_positions((-10.0, 20.0, 40.0), total_frames, omega, 0.02, rt.move_light)
    mine = rt.sharding.frames_for_rank(total_frames, rank, world)
    p3 = rt.params_profile("cpu", W, H, 1, 5)
    ms3, rays3, l3 = _timed_frames(rt, torch, sc, stream, flush, p3, len(mine), light_of=lambda i: orbit[mine[max(i - 3, 0)]])
    ms3_max = _max_over_ranks(torch, world, ms3)
    rays_total = _sum_over_ranks(torch, world, rays3 * len(mine))
    cfg3 = {"workload": "BASELINE.json configs[3]: circulating-light spheres animation, 240 frames at 1920x1080, frame f on rank f mod N (no data-path collective)",
            "frames": total_frames, "frames_per_gpu": len(mine), "ms_total_max_over_ranks": round(ms3_max, 3), "ms_per_frame_per_gpu": round(ms3 / max(len(mine), 1), 5),
            "frames_per_s": round(total_frames / (ms3_max * 1e-3), 1), "rays_per_frame": rays3, "mrays_per_s": round(rays_total / (ms3_max * 1e-3) / 1e6, 1),
            "launches_per_frame": l3, "scaling": "strong (240 frames whatever N)"}
    if rank == 0 and world == 1:
        try:  # the reference's own CPU classes on the same scene, bounded sample (oracle/_ref/libref_cpu.so, scene kind 2)
            from oracle import pyoracle
            if pyoracle.ref_cpu_available():
                th = os.cpu_count() or 1
                pyoracle.ref_cpu_set_light((-10.0, 20.0, 40.0))
                pyoracle.ref_cpu_render("", 2, 200, 150, 1, 5, th, hits=False)
                r0 = [pyoracle.ref_cpu_render("", 2, 800, 600, 1, 5, th, hits=False)["seconds"] for _ in range(5)]
                r3 = [pyoracle.ref_cpu_render("", 2, W, H, 1, 5, th, hits=False)["seconds"] for _ in range(3)]
                cfg0["cpu_reference"] = {"ms_per_frame": round(float(np.median(r0)) * 1e3, 2), "cores": th, "kind": "reference",
                                         "sample": "5 frames by cpu_launcher.cpp's classes (its getColor also takes the random indirect bounce)"}
                cfg3["cpu_reference"] = {"ms_per_frame": round(float(np.median(r3)) * 1e3, 2), "cores": th, "kind": "reference", "sample": "3 of the 240 frames"}
        except Exception as e:  # noqa: BLE001
            cfg0["cpu_reference"] = {"unavailable": repr(e)[:200]}
    sc.close()
    return cfg0, cfg3


def config4_ten_million(rt, torch, local, world, rank, flush):
    """BASELINE.json configs[4]: the cat instanced to 9,999,666 triangles (2,529 baked copies, raytracinggpu_b200.synthetic), 3840x2160,
    primary + shadow rays. The scene is built on rank 0 only (device BVH builder + upload), packed, broadcast ONCE, and every rank
    renders the rows r, r + N, ...; the bands go into rank 0's frame over NVLink (FramePush). The only scene that does not fit the
    126 MB L2 (0.85 GB blob): its roofline is HBM, stated from MEASURED DRAM bytes (committed ncu capture of this frame)."""
    import torch.distributed as dist
    from raytracinggpu_b200 import distributed as rtd, synthetic
    cat = find_cat()
    if not cat:
        return {"unavailable": "cat.obj not on this box"}
    W4, H4 = 3840, 2160
    sc = rt.Scene(local)
    stream = torch.cuda.Stream()
    sc.set_stream(stream.cuda_stream)
    build = {}
    if rank == 0:
        t0 = time.perf_counter()
        scales, offs = synthetic.instance_lattice()
        mesh = rt.Mesh.read_obj(cat).instance(scales, offs)
        t1 = time.perf_counter()
        mesh.build_bvh_gpu(local)
        t2 = time.perf_counter()
        walls, mesh_id = rt.default_walls("optimized")
        sc.set_spheres(walls)
        sc.set_mesh_from(mesh, id=mesh_id)  # the builder's arrays never leave the device (rt_scene_set_mesh_device: relayout + repack as kernels)
        sc.sync()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        nv, nt, nn = mesh.counts()
        build = {"triangles": nt, "bvh_nodes": nn, "host_instancing_ms": round((t1 - t0) * 1e3, 1), "upload_and_bvh_build_wall_ms": round((t2 - t1) * 1e3, 1),
                 "bvh_build_device_ms": round(mesh.build_ms, 2), "device_relayout_and_repack_ms": round((t3 - t2) * 1e3, 1),
                 "scene_ready_ms_after_the_arrays_exist": round((t3 - t1) * 1e3, 1),
                 "note": "round 1: host node relayout + 1.1 GB over PCIe, 1174 ms; builder results copied back and reordered on the host, 477 ms"}
        del mesh
    bcast_ms, blob = 0.0, sc.blob_size() if rank == 0 else 0
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        blob = rtd.broadcast_scene(sc, src=0)
        torch.cuda.synchronize()
        dist.barrier()
        bcast_ms = (time.perf_counter() - t0) * 1e3
    p = rt.params_profile("optimized", W4, H4, 1, 1)
    frames = 6
    out = {"workload": "BASELINE.json configs[4]: cat instanced to 9,999,666 triangles, 3840x2160, primary + shadow rays, BVH built on rank 0 and broadcast once, groups of %d rows interleaved over %d GPU(s)" % (ROW_GROUP, world),
           "build_on_rank0": build, "scene_blob_bytes": int(blob), "scene_broadcast_ms": round(bcast_ms, 2),
           "scene_broadcast_gbs": round(blob / (bcast_ms * 1e-3) / 1e9, 1) if bcast_ms > 0 else None, "scaling": "strong"}
    if world == 1:
        ms, rays, launches = _timed_frames(rt, torch, sc, stream, flush, p, frames, warm=2)
        out.update({"ms_per_frame": round(ms / frames, 4), "rays_per_frame": rays, "mrays_per_s": round(rays * frames / (ms * 1e-3) / 1e6, 1), "launches_per_frame": launches})
    else:
        fp = rtd.FramePush(sc, H4, W4, world, rank, torch.device("cuda", local), group=ROW_GROUP)
        pp = fp.apply(p)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(frames)]
        rays = 0
        for i in range(2 + frames):
            dist.barrier()
            with torch.cuda.stream(stream):
                k = i - 2
                flush.zero_()
                if k >= 0:
                    ev[k][0].record(stream)
                sc.render_into(pp, rgb=fp.band, flags=rt.RT_RENDER_NO_SYNC)
                if k >= 0:
                    ev[k][1].record(stream)
                fp.push()
                if k >= 0:
                    ev[k][2].record(stream)
            rays = int(sc.sync().rays)
        torch.cuda.synchronize()
        tot = _max_over_ranks(torch, world, float(sum(a.elapsed_time(c) for a, _, c in ev)))
        ren = _max_over_ranks(torch, world, float(sum(a.elapsed_time(b) for a, b, _ in ev)))
        rays_all = _sum_over_ranks(torch, world, rays)
        fp.close()
        out.update({"ms_per_frame": round(tot / frames, 4), "render_ms": round(ren / frames, 4), "push_and_barrier_ms": round((tot - ren) / frames, 4),
                    "rays_per_frame": rays_all, "mrays_per_s": round(rays_all * frames / (tot * 1e-3) / 1e6, 1)})
    ncu = ncu_frame_summary("*config4*ncu*.json", exclude=()) or ncu_frame_summary("*config5*ncu*.json", exclude=())
    if ncu and "ms_per_frame" in out and world == 1:
        hbm_peak, hbm_src = measured_peaks()
        gbs = ncu["dram_bytes"] / (out["ms_per_frame"] * 1e-3) / 1e9
        out["roofline"] = {"bound": "hbm", "achieved": round(gbs, 1), "peak": hbm_peak, "unit": "GB/s", "frac": round(gbs / hbm_peak, 4), "traffic": ncu["dram_bytes"],
                           "peak_source": hbm_src, "note": "MEASURED DRAM bytes of one frame (profiles/%s) over this run's frame time: the tree search of this scene is L2-latency / issue bound, not HBM bound" % ncu["file"]}
    sc.close()
    return out


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
        p.z = rt.camera_z_device(W)  # optimized.cu evaluates z inside the kernel (rt_camera_z_device)
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
    fg = rtd.FrameGather(H4, W4, world, rank, torch.device("cuda", torch.cuda.current_device()), group=ROW_GROUP)
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
        fp = rtd.FramePush(sc, H4, W4, world, rank, torch.device("cuda", torch.cuda.current_device()), group=ROW_GROUP)
        push_signal = "flags in rank 0's memory (rt_peer_signal / rt_peer_wait)" if fp.signal == "flags" else "one all-reduce per frame"
        pp = fp.apply(rt.params_profile("optimized", W4, H4, 1, 4))
        ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(frames)]
        kernel_ms2 = []
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
            st2 = sc.sync()
            if i >= 3:
                kernel_ms2.append(st2.kernel_ms)
        torch.cuda.synchronize()
        t2 = torch.tensor([sum(a.elapsed_time(c) for a, _, c in ev2), sum(a.elapsed_time(b) for a, b, _ in ev2)], device="cuda", dtype=torch.float64)
        per_rank = [torch.zeros(2, device="cuda", dtype=torch.float64) for _ in range(world)]
        dist.all_gather(per_rank, t2)
        lib_ms = torch.tensor([float(np.median(kernel_ms2))], device="cuda", dtype=torch.float64)  # rt_render's own event pair (graph launch to last kernel)
        lib_ranks = [torch.zeros(1, device="cuda", dtype=torch.float64) for _ in range(world)]
        dist.all_gather(lib_ranks, lib_ms)
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
                "push_and_barrier_ms": round(float(t2[0].item() - t2[1].item()) / frames, 4), "frame_equals_all_gather": same, "completion": push_signal,
                "render_ms_per_rank": [round(float(x[1].item()) / frames, 4) for x in per_rank],
                "kernels_only_ms_per_rank": [round(float(x.item()), 4) for x in lib_ranks],
                "how": "rank 0 owns the frame (CUDA IPC handle broadcast once); every rank copies its band into it over NVLink and writes a completion flag behind it"}
    return {"workload": "BASELINE.json configs[2]: mirror cat 3840x2160, reflection depth 4, groups of %d rows interleaved over %d GPU(s), all-gather to every rank" % (ROW_GROUP, world),
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
           "config": {"workload": workload_string("cadnav cat (3954 tris, 2019 BVH nodes)" if cat else "synthetic torus (cat.obj unavailable)"),
                      "arm": "each step = one full frame on the host CPU (cpu_launcher.cpp's classes, OpenMP over rows)", "rays_per_frame": rays, "wall_s": round(wall, 3)},
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
