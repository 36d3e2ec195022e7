#!/usr/bin/env python3
"""bench.py — the ray-query benchmark (BASELINE.json metric: Mrays/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload spheres|terrain|cornell] [--impl reference]

A step = one pass of the hot path (closest-hit traversal of one frame's worth of rays) over one
batch of synthetic rays.  Default workload = BASELINE.json configs[1]: the procedural 1 048 576-
triangle sphere field, 1920 x 1080 coherent primary rays generated on the device by the camera
kernel in the reference's 32 x 32 tile order, closest hit only.

Prints ONE JSON line (rank 0).  `value` is whole-job Mrays/s with the ray stream resident in HBM,
timed with CUDA events around the traversal kernel (L2 flushed and rays regenerated, untimed, before
every step); `e2e` is the same metric through the C ABI with page-locked HOST ray arrays, copies
inside the timed region; `roofline` relates the traversal kernel to the measured HBM bandwidth using
the oracle's per-ray node / triangle counts; `cpu_baseline` is the reference's own stream kernel
(oracle/_ref, compiled from the reference sources) on all host threads over the same rays.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from phosphorus_mk2_b200 import scenes  # noqa: E402
from phosphorus_mk2_b200.rays import HIT, RayBatch  # noqa: E402

S_NODE, S_TRI = 80, 48  # bytes of a packed node / triangle (csrc/phos_internal.hpp)
B_IO_CLOSEST = 56       # 32 B read (p, wi, d, flags) + 24 B written (d, flags, mesh, face, u, v) per ray


def make_workload(name: str):
    if name == "spheres":
        return scenes.sphere_field(), "configs[1]: procedural 1M-triangle tessellated sphere field, 1920x1080 coherent primary rays, closest-hit"
    if name == "terrain":
        return scenes.terrain(), "configs[2] geometry: procedural 10M-triangle displaced terrain, 1920x1080 primary rays, closest-hit"
    if name == "terrain_ggx":
        return scenes.terrain(glossy_fraction=0.1), "configs[3]: 10M-triangle terrain, diffuse + 10% GGX (roughness 0.2) face bands, sky emitter, 1920x1080"
    if name == "instanced30m":
        return scenes.instanced_field(), "configs[4]: 30M-triangle field of baked copies of one 4096-triangle object, 3840x2160"
    if name == "cornell":
        return scenes.cornell_box(), "configs[0]: synthetic Cornell box 512x512 primary rays, closest-hit"
    if name == "tiny":
        return scenes.sphere_field(4, 16, 8, 256, 256), "tiny smoke workload (4x4 spheres, 256x256)"
    raise SystemExit(f"unknown workload {name}")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        if not sm:
            return None
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if r[col].lower() == "active":
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(self.rows[0][2]), "samples": len(sm),
                "reasons": sorted(reasons)}


def tile_order_tiles(cam):
    from phosphorus_mk2_b200.device import make_tiles
    return make_tiles(cam.film_width, cam.film_height, 32)


def run_reference_arm(args, scene, label):
    """The reference's own CPU implementation of the path (oracle/_ref), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.pyoracle import Oracle, RefLib
    orc = Oracle()
    cam = scene.camera
    parts = [orc.camera_rays(cam, x, y, w, h) for (x, y, w, h) in tile_order_tiles(cam)]
    rays = RayBatch(sum(p.n for p in parts))
    o = 0
    for p in parts:
        for f in ("px", "py", "pz", "wx", "wy", "wz", "d", "flags"):
            getattr(rays, f)[o:o + p.n] = getattr(p, f)
        o += p.n
    if RefLib.available():
        ref = RefLib()
        rs = ref.scene(scene)
        build_s = rs.build()
        cores = ref.hardware_concurrency()
        kind = "reference"

        def step():
            return rs.trace(rays, "stream", threads=cores)[1]
    else:
        from phosphorus_mk2_b200.device import Accel
        acc = Accel(scene)
        build_s = acc.build_seconds
        nodes, packets = acc.nodes_array(), acc.packets_array()
        cores, kind = 1, "port"

        def step():
            t0 = time.perf_counter()
            orc.traverse(nodes, packets, rays)
            return time.perf_counter() - t0
    for _ in range(args.warmup):
        step()
    secs = [step() for _ in range(args.steps)]
    total = sum(secs)
    v = rays.n * args.steps / total / 1e6
    line = {"impl": "reference", "metric": "Mrays/s (closest-hit)", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": label, "rays_per_step": rays.n, "triangles": scene.num_triangles()},
            "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": kind,
                             "sample": f"whole frame ({rays.n} rays) per step, reference stream_mbvh_kernel_t, {cores} threads; "
                                       f"BVH build {build_s:.2f} s excluded"},
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_render_bench(args, scene, label):
    """Path-traced samples/s: a step = one frame of --spp samples per pixel at --depth, the frame split
    over the ranks by tiles (or sample ranges), one NCCL film reduce per frame inside the timed region."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cam = scene.camera
    n_px = cam.film_width * cam.film_height
    if args.impl == "reference":
        if rank != 0:
            return
        from oracle.pyoracle import RefLib
        ref = RefLib()
        rs = ref.scene(scene)
        spp = 4 if n_px > 512 * 512 else 16  # bounded sample: the CPU renderer at a reduced, perfect-square spp
        secs = []
        for i in range(args.warmup + args.steps):
            _, s_ = rs.render(spp, 1, args.depth, single_threaded=False)
            if i >= args.warmup:
                secs.append(s_)
        v = n_px * spp * len(secs) / sum(secs)
        print(json.dumps({"impl": "reference", "metric": "path-traced samples/s", "value": v, "unit": "samples/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs),
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": label, "spp": spp, "depth": args.depth, "pixels": n_px},
                          "cpu_baseline": {"value": v, "unit": "samples/s", "cores": ref.hardware_concurrency(), "kind": "reference",
                                           "sample": f"{spp} spp frame (reference cpu_t::start/join, all host threads)"},
                          "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from phosphorus_mk2_b200.device import Accel, CudaDevice, Options
    from phosphorus_mk2_b200.frame import film_tensor, reduce_film, samples_of_rank, tiles_of_rank
    dev = CudaDevice.make(Options(samples_per_pixel=args.spp, paths_per_sample=1, path_depth=args.depth), local)
    acc = Accel(scene)
    dev.preprocess(scene, acc)
    dev.upload_scene(scene)
    tiles = tile_order_tiles(cam)
    if args.partition == "tiles":
        my_tiles, my_range = tiles_of_rank(tiles, rank, world), (0, args.spp)
    else:
        my_tiles, my_range = tiles, samples_of_rank(args.spp, rank, world)
    film_t = film_tensor(dev) if dist else None

    def frame(read_back: bool):
        dev.film_clear()
        if my_range[1] > my_range[0] and my_tiles:
            dev.render(my_tiles, my_range[0], my_range[1], args.spp, seed=1)
        dev.synchronize()
        if dist:
            reduce_film(film_t, dist, 0)
            import torch
            torch.cuda.synchronize()
        if read_back and rank == 0:
            return dev.film_read()
        return None

    for _ in range(args.warmup):
        frame(False)
    if dist:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = dev.launch_count()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        frame(False)
    if dist:
        dist.barrier()
    dt = time.perf_counter() - t0
    launches = dev.launch_count() - l0
    t0 = time.perf_counter()
    img = None
    for _ in range(max(1, args.steps // 2)):
        img = frame(True)
    if dist:
        dist.barrier()
    dt_e2e = (time.perf_counter() - t0) / max(1, args.steps // 2)
    if dist:
        import torch
        t = torch.tensor([dt, dt_e2e], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, dt_e2e = float(t[0]), float(t[1])
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        v = n_px * args.spp * args.steps / dt
        print(json.dumps({"metric": "path-traced samples/s", "value": v, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": label, "spp": args.spp, "depth": args.depth, "pixels": n_px, "partition": args.partition,
                                     "timing": "wall clock around K frames (render + NCCL film reduce), barrier + sync both sides"},
                          "clocks": clocks, "gpu_launches": int(launches),
                          "e2e": {"value": n_px * args.spp / dt_e2e, "unit": "samples/s", "h2d_bytes_per_step": 16 * len(my_tiles) + 8 * args.spp,
                                  "d2h_bytes_per_step": 16 * n_px},
                          "image_mean": float(img[..., :3].mean()) if img is not None else None}), flush=True)
    dev.close()
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="spheres")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--render", action="store_true", help="path-traced samples/s (BASELINE configs[0]/[3]) instead of Mrays/s")
    ap.add_argument("--spp", type=int, default=16)
    ap.add_argument("--depth", type=int, default=8)
    ap.add_argument("--partition", default="tiles", choices=["tiles", "samples"])
    args = ap.parse_args()
    if args.impl != "reference" and not args.render:
        args.warmup = max(args.warmup, 3)  # timing hygiene for the kernel metric; frame benches take what they are given

    scene, label = make_workload(args.workload)
    if args.render:
        run_render_bench(args, scene, label)
        return
    if args.impl == "reference":
        run_reference_arm(args, scene, label)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from phosphorus_mk2_b200.device import Accel, CudaDevice, Options, pinned_ray_batch

    dev = CudaDevice.make(Options(), local)  # raises without the library or a GPU: no fallback
    t0 = time.perf_counter()
    acc = Accel(scene)
    dev.preprocess(scene, acc)
    dev.upload_scene(scene)
    st = dev.accel_stats()
    prep_s = time.perf_counter() - t0
    cam = scene.camera
    tiles = tile_order_tiles(cam)
    n = cam.film_width * cam.film_height
    drays = dev.device_rays(n)

    # ---- device-resident: value ------------------------------------------------------------------
    def dev_step(timed: bool) -> float:
        dev.flush_l2()                      # untimed: evict the BVH and the previous step's rays
        dev.camera_rays(tiles, drays)       # untimed: fresh primary rays (d = FLT_MAX, flags = 0)
        if not timed:
            dev.trace_device(drays)
            dev.synchronize()
            return 0.0
        dev.timer_begin()
        dev.trace_device(drays)
        return dev.timer_end()

    for _ in range(args.warmup):
        dev_step(False)
    if dist:
        dist.barrier()
    dev.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = dev.launch_count()
    wall0 = time.perf_counter()
    step_ms = [dev_step(True) for _ in range(args.steps)]
    dev.synchronize()
    if dist:
        dist.barrier()
    wall = time.perf_counter() - wall0
    launches = dev.launch_count() - launches0
    total_ms = float(sum(step_ms))
    if dist:
        import torch
        t = torch.tensor([total_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    value = world * n * args.steps / (total_ms * 1e-3) / 1e6

    # GPU's own fetch counts for the same batch (for the record; the roofline uses the oracle's counts)
    dev.camera_rays(tiles, drays)
    g_nodes, g_tris = dev.trace_count(drays)
    result = drays.download()
    hits = int(result.hit.sum())

    # ---- the same frame as occlusion queries (BASELINE metric "closest-hit, shadow"): any-hit Mrays/s ------------
    # SHADOW set on every camera ray (tmax = FLT_MAX): the traversal stops at the first accepted triangle.  Rank 0.
    any_hit = None
    if rank == 0:
        from phosphorus_mk2_b200.rays import SHADOW
        dev.camera_rays(tiles, drays)
        shadow_host = drays.download()
        shadow_host.flags[:] = SHADOW
        ms = []
        for i in range(3 + max(3, min(args.steps, 10))):
            dev.flush_l2()
            drays.upload(shadow_host)
            dev.timer_begin()
            dev.trace_device(drays)
            t = dev.timer_end()
            if i >= 3:
                ms.append(t)
        occluded = drays.download()
        any_hit = {"value": n / (float(np.mean(ms)) * 1e-3) / 1e6, "unit": "Mrays/s", "rays": n,
                   "occluded_fraction": float(occluded.hit.mean()), "verdict_equals_closest_hit": bool(np.array_equal(occluded.hit, result.hit))}

    # ---- end to end through the C ABI with pinned host arrays: e2e ------------------------------------
    dev.camera_rays(tiles, drays)
    pristine = drays.download()
    # one rank per GPU: stay on the CPUs next to this GPU before allocating the page-locked arrays (first touch),
    # so the up-copies and the zero-copy result stores do not cross the socket interconnect
    numa_cpus = dev._L.phos_cuda_bind_host_to_device(local)
    hrays = pinned_ray_batch(n)
    fields = ("px", "py", "pz", "wx", "wy", "wz", "d", "u", "v", "mesh", "face", "flags")

    def restore():
        for f in fields:
            getattr(hrays, f)[:] = getattr(pristine, f)

    for _ in range(2):
        restore()
        dev.trace(hrays)
    e2e_s = 0.0
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(e2e_steps):
        restore()
        if dist:
            dist.barrier()
        t1 = time.perf_counter()
        dev.trace(hrays)  # blocking: H2D + traversal + D2H
        e2e_s += time.perf_counter() - t1
    if dist:
        import torch
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = world * n * e2e_steps / e2e_s / 1e6
    e2e_ok = bool(np.array_equal(hrays.flags, result.flags) and np.array_equal(hrays.d.view(np.uint32), result.d.view(np.uint32)))

    # ---- CPU baseline + oracle counts (rank 0, N = 1 only) ------------------------------------------
    cpu = None
    roof = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.pyoracle import Oracle, RefLib  # the checker / CPU baseline leg only
        orc = Oracle()
        nodes, packets = acc.nodes_array(), acc.packets_array()
        # oracle traversal of EVERY ray of the frame (a few seconds): per-ray node / triangle counts for the
        # roofline, and the parity verdict of the whole batch
        sel = np.arange(n)
        sample = RayBatch(len(sel))
        for f in fields:
            getattr(sample, f)[:] = getattr(pristine, f)[sel]
        want, cnt = orc.traverse(nodes, packets, sample)
        n_node, n_tri = cnt.nodes / cnt.rays, cnt.triangles / cnt.rays
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from parity import mismatches
        got = RayBatch(len(sel))
        for f in fields:
            getattr(got, f)[:] = getattr(result, f)[sel]
        parity = {"sample_rays": int(len(sel)), "mismatch_vs_oracle": int(len(mismatches(sample, got, want)))}
        b_ray = B_IO_CLOSEST + n_node * S_NODE + n_tri * S_TRI
        peak, peak_src = peaks()
        ms_kernel = total_ms / args.steps
        achieved = n * b_ray / (ms_kernel * 1e-3) / 1e9
        traffic = None  # DRAM bytes per launch from the committed ncu capture of this workload, if there is one
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(args.workload)
        except Exception:
            pass
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "algorithmic_bytes_per_launch": n * b_ray,
                "kernel": "trace_kernel", "peak_source": peak_src, "bytes_per_ray": b_ray,
                "oracle_nodes_per_ray": n_node, "oracle_tris_per_ray": n_tri,
                "gpu_fetched_bytes_per_ray": B_IO_CLOSEST + (g_nodes * S_NODE + g_tris * S_TRI) / n,
                "gpu_nodes_per_ray": g_nodes / n, "gpu_tris_per_ray": g_tris / n}
        if RefLib.available():
            ref = RefLib()
            rs = ref.scene(scene)
            build_s = rs.build()
            cores = ref.hardware_concurrency()
            _, s1 = rs.trace(pristine, "stream", threads=cores)       # calibration / warm-up pass
            reps = int(max(1, min(40, 10.0 / max(s1, 1e-3))))
            secs = sum(rs.trace(pristine, "stream", threads=cores)[1] for _ in range(reps))
            cpu = {"value": n * reps / secs / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "reference",
                   "sample": f"the same {n}-ray frame x {reps} passes ({secs:.1f} s), reference stream_mbvh_kernel_t on {cores} threads, "
                             f"reference BVH build {build_s:.2f} s excluded"}
        else:
            t1 = time.perf_counter()
            orc.traverse(nodes, packets, sample)
            secs = time.perf_counter() - t1
            cpu = {"value": len(sel) / secs / 1e6, "unit": "Mrays/s", "cores": 1, "kind": "port",
                   "sample": f"the whole {len(sel)}-ray frame once, scalar oracle traversal"}

    if rank == 0:
        line = {"metric": "Mrays/s (closest-hit)", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": label, "rays_per_step_per_gpu": n, "triangles": scene.num_triangles(),
                           "packed_nodes": st.nodes, "packed_bytes": int(st.bytes_nodes + st.bytes_triangles),
                           "l2": "flushed (256 MiB memset) and rays regenerated before every timed step",
                           "timing": "CUDA events around the traversal kernel, summed over steps, max over ranks",
                           "hit_fraction": hits / n, "preprocess_s": prep_s, "wall_s_timed_loop": wall},
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": e2e, "unit": "Mrays/s", "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 8 * n + 16 * hits,
                        "steps": e2e_steps, "matches_device_path": e2e_ok, "host_cpus_bound": numa_cpus},
                "any_hit": any_hit,
                "roofline": roof, "cpu_baseline": cpu, "parity": parity}
        print(json.dumps(line), flush=True)
    dev.close()
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
