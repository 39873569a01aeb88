#!/usr/bin/env python
"""bench.py -- the hot-path benchmark (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c3|c4|c5|cornell]

A *step* is one progressive frame of the workload: `spp` passes of the wavefront path tracer over the whole
image (each pass = 1 path per pixel, <= depth vertices, 1 extension + <= 1 shadow ray per vertex) followed, for
N > 1, by the single NCCL reduce of the PixelData accumulator to rank 0.  Metric: Mrays/s, a ray being one
closest-hit/any-hit query exactly as the reference counts them (Kernel/TraceHelper.cu:176).

N = 1 : the whole image on one B200.  N > 1 : one process per GPU (torchrun), scene replicated, image split in
interleaved 64x64 tiles (tile % N == rank), "scaling": "strong" (the total work -- one image -- is fixed).

`value`   : device-timed (CUDA events on the launching stream) with the scene resident in HBM.
`e2e`     : the same frames through the public API with HOST buffers: per step the sample tables of every pass are
            generated on the host (DeviceSampleTables=0, the reference's UpdateKernel behaviour) and copied H2D from pinned
            memory, and the frame is resolved (default image pipeline) and the RGBA8 image copied D2H into pinned memory.
`roofline`: the traversal kernel (k_intersect, extension + shadow launches): algorithmic bytes (visit counts of an
            instrumented pass x SURVEY 8d's per-visit bytes) / live CUDA-event time of those launches.
`cpu_baseline` / `--impl reference`: the reference's CPU path (oracle/_ref when built, else the oracle port) on the
            host cores, on a bounded crop of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (scene kind, width, height, spp, MaxPathLength, description)
    "cornell": ("cornell", 256, 256, 1, 8, "Cornell-32 256x256 1spp depth 8 (configs[0])"),
    "c2": ("c2", 1920, 1080, 8, 8, "procedural 100K-triangle diffuse scene (99,854 tris) 1920x1080 8spp 8 bounces (configs[1])"),
    "c3": ("c3", 1920, 1080, 8, 8, "100K scene, diffuse/roughconductor/dielectric mix, 1920x1080 8spp 8 bounces (configs[2])"),
    "c4": ("c4", 1920, 1080, 8, 8, "procedural 1M-triangle clustered scene (999,854 tris) 1920x1080 8spp 8 bounces (configs[3])"),
    "c5": ("c5", 1920, 1080, 64, 32, "1M scene, microfacet mix, 32 bounces 64spp (configs[4])"),
}
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def baseline_metric():
    """The metric name as BASELINE.json states it (value = its Mrays/s part; its 'BVH kernel HBM GB/s vs peak' part is `roofline`)."""
    try:
        with open(os.path.join(ROOT, "BASELINE.json")) as f:
            return json.load(f)["metric"]
    except Exception:
        return "Mrays/s at 1920x1080x8spp, 8-bounce path trace; BVH kernel HBM GB/s vs peak"


def hbm_peak():
    """Measured HBM copy bandwidth (GB/s) from the driver-written MEASURED_PEAKS.json, else the profiling guide's fallback."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            d = json.load(f)

        def find(o):
            if isinstance(o, dict):
                for k in ("hbm_gbs", "hbm_gb_s", "hbm_GBps", "hbm_gbps"):
                    if k in o and isinstance(o[k], (int, float)):
                        return float(o[k])
                for k, v in o.items():
                    if isinstance(v, (int, float)) and "hbm" in k.lower() and 500.0 < float(v) < 20000.0:
                        return float(v)
                for v in o.values():
                    r = find(v)
                    if r:
                        return r
            elif isinstance(o, list):
                for v in o:
                    r = find(v)
                    if r:
                        return r
            return None
        v = find(d)
        if v:
            return v, "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv"); os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def crop_window(w, h, frac):
    """Centre crop holding `frac` of the pixels (same aspect)."""
    s = frac ** 0.5
    cw, ch = max(16, int(w * s)) // 8 * 8, max(16, int(h * s)) // 8 * 8
    x0, y0 = (w - cw) // 2, (h - ch) // 2
    return (x0, y0, x0 + cw, y0 + ch)


def cpu_reference_run(view, w, h, depth, window, n_passes, threads):
    """Time the reference's CPU path on `window`: oracle/_ref when present, else the oracle port."""
    import oracle_binding as ob
    kind = "port"
    try:
        import ref_binding as rb
        if rb.available():
            kind = "reference"
    except Exception:
        rb = None
    t0 = time.perf_counter()
    if kind == "reference":
        _, rays = rb.render(view, w, h, n_passes=n_passes, max_path_length=depth, window=window, n_threads=threads)
    else:
        _, rays = ob.render(view, w, h, n_passes=n_passes, max_path_length=depth, window=window, n_threads=threads)
    dt = time.perf_counter() - t0
    return kind, rays, dt


def static_config(name, scene, world, tile):
    """What both arms state identically about the workload (no measured values in here)."""
    kind, w, h, spp, depth, desc = WORKLOADS[name]
    return {"workload": desc, "width": w, "height": h, "spp": spp, "max_path_length": depth, "rr_start_depth": 5, "direct": True, "triangles": scene.n_triangles,
            "partition": "whole image" if world == 1 else f"interleaved {tile}x{tile} tiles, tile % {world} == rank; one NCCL reduce of PixelData (7*w*h f32) per step",
            "l2": "256 MiB buffer written between timed steps (L2 flush); per-pass queue/path-state working set ~0.5 GB > 126 MB L2"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref: its PathTrace / traceRay sources compiled here) on all host
    cores, rank 0 only.  Same workload and config as the b200 arm; every step renders a bounded sample of the frame -- a centre crop, one pass -- sized by
    a probe so that the whole run stays within a few minutes (a full 1 M-triangle frame is ~160 M rays, minutes per step on the host)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cudatracerlib_b200 import Scene, TILE
    kind, w, h, spp, depth, desc = WORKLOADS[args.workload]
    scene = Scene(kind, w, h)
    if args.ref_own_tree:
        if args.workload == "cornell":
            raise SystemExit("--ref-own-tree: the 100 K / 1 M workloads only")
        scene = scene_on_reference_tree(scene, w, h)
    threads = os.cpu_count() or 1
    if args.workload == "cornell":
        window = (0, 0, w, h)
    else:
        probe = crop_window(w, h, 1.0 / 256)
        _, prays, pdt = cpu_reference_run(scene.view, w, h, depth, probe, 1, threads)
        budget_s = max(0.5, min(6.0, args.ref_seconds / max(1, args.warmup + args.steps)))   # seconds of host work per step
        frac = min(1.0, (1.0 / 256) * budget_s / max(pdt, 1e-3))
        window = crop_window(w, h, frac)
    times, rays_tot, impl_kind = [], 0, "port"
    for i in range(args.warmup + args.steps):
        impl_kind, rays, dt = cpu_reference_run(scene.view, w, h, depth, window, 1, threads)
        if i >= args.warmup:
            times.append(dt); rays_tot += rays
    total = sum(times)
    v = rays_tot / total / 1e6
    sample = f"{window[2]-window[0]}x{window[3]-window[1]} centre crop of the {w}x{h} image, 1 pass per step (of {spp}), depth {depth}" + \
             (", mesh trees built by the reference's own SplitBVHBuilder" if args.ref_own_tree else ", trees of this repo's builder (same arrays as the b200 arm)")
    line = {"impl": "reference", "metric": baseline_metric(), "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / max(1, args.steps), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": static_config(args.workload, scene, max(1, args.gpus), args.tile or TILE),
            "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": threads, "kind": impl_kind, "sample": sample},
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def scene_on_reference_tree(scene, w, h):
    """The same scene with every mesh BVH built by the reference's OWN SplitBVHBuilder (through oracle/_ref's Mesh::CompileMesh -> .xmsh writer, then the
    .xmsh import): the CPU baseline as "the reference on its own tree".  Minutes for 1 M triangles (its builder is single-threaded)."""
    import ctypes as C
    import ref_binding as rb
    from cudatracerlib_b200 import api, Scene
    if not rb.available():
        raise SystemExit("--ref-own-tree needs oracle/_ref (built where /root/reference is mounted)")
    scene.setRebraid(0)
    tmp = tempfile.mkdtemp()
    tri_data = scene.array("tri_data"); meshes = scene.array("meshes"); nodes = scene.array("nodes")
    v = scene.view
    mats_all = (api.Material * v.n_materials).from_address(C.addressof(v.materials.contents))
    lights = (api.Light * max(1, v.n_lights_buf)).from_address(C.addressof(v.lights.contents)) if v.n_lights_buf else []
    paths = []
    for mi in range(len(meshes)):
        T = scene.mesh_triangles(mi)
        toff, moff = int(meshes[mi][0]), int(meshes[mi][4])
        mat_idx = ((tri_data[toff:toff + len(T), 1] >> 16) & 0xff).astype(np.int64)
        order = np.argsort(mat_idx, kind="stable"); used = int(mat_idx.max()) + 1
        V = T[order].reshape(-1, 3); I = np.arange(len(V), dtype=np.uint32)
        node = next(n for n in range(len(nodes)) if int(nodes[n][0]) == mi)
        em = np.zeros((used, 3), np.float32)
        for k in range(used):
            nli = mats_all[int(nodes[node][1]) + k].node_light_index
            if nli != 0xffffffff:
                em[k] = list(lights[int(nodes[node][4 + nli])].radiance)
        pth = os.path.join(tmp, f"m{mi}.xmsh")
        rb.write_xmsh(pth, V, I, np.bincount(mat_idx, minlength=used), [mats_all[moff + k] for k in range(used)], em)
        paths.append(pth)
    xf = scene.array("node_xf").reshape(-1, 16)
    cam = ((0, 0, -9.5), (0, 0, 0), (0, 1, 0), 60.0)   # camera of the 100 K / 1 M scenes (csrc/scene_builder.cpp)
    s2 = Scene.from_xmsh([paths[int(nodes[n][0])] for n in range(len(nodes))], *cam, w, h, node_xforms=xf)
    s2.setRebraid(0)
    return s2


def measure_workload(name, args, torch, dist, world, rank, local, dev, stream, steps, warmup, primary):
    """One workload on this rank's share of the image.  Returns (line fields of rank 0, scene)."""
    from cudatracerlib_b200 import Scene, PathTracer, traversal_bytes, TILE
    kind, w, h, spp, depth, desc = WORKLOADS[name]
    tile = args.tile if args.tile > 0 else TILE
    scene = Scene(kind, w, h)
    tracer = PathTracer(w, h, device=local)
    tracer.InitializeScene(scene)
    tracer.setParameter("MaxPathLength", depth)
    if args.sort_mode is not None:
        tracer.setParameter("SortMode", args.sort_mode)
    reupload = False
    for kv in args.set:
        k, v = kv.split("="); tracer.setParameter(k, int(v)); reupload |= k == "StagedTreeletNodes"
    if reupload:
        tracer.InitializeScene(scene)
    tracer.setStream(stream.cuda_stream)
    accum = torch.zeros(h * w * 7, dtype=torch.float32, device=dev)
    tracer.setAccumDevicePtr(accum.data_ptr())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    d_rgba = torch.empty(h * w * 4, dtype=torch.uint8, device=dev)
    host_rgba = torch.empty(h * w * 4, dtype=torch.uint8, pin_memory=True)
    table_bytes = 4096 * 30 * 12
    batch = min(spp, args.batch if args.batch > 0 else (16 if spp >= 32 else 8))   # frames of many passes (configs[4]): 16 per wavefront, the wavefronts overlapped on lanes (ctl_render_frame_tiled)
    if spp % batch:
        raise SystemExit(f"--batch {batch} does not divide spp {spp}")

    def render_pass(p, new_trace):
        # `batch` progressive passes fused into one wavefront (ctl_render_passes_tiled) on this rank's tiles
        tracer.DoPasses(batch, new_trace=new_trace, tile=(tile, tile), part=rank, n_parts=world)

    if world > 1:
        # the communicator lives behind the C ABI (csrc/ctl_comm.cu: NCCL loaded at run time); torch.distributed only carries the 128-byte id to the ranks
        box = [PathTracer.commUniqueId() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        tracer.commInitRank(box[0], rank, world)

    # Frames in flight: a frame that is ONE wavefront (spp == batch) cannot hide the drain of its own persistent traversal launches, the next frame's launches can
    # (ctl_comm_submit_frame / ctl_acquire_frame: every frame on a lane of its own -- stream, wavefront buffers, accumulator, sample tables -- its reduce on the
    # communication stream; frames are handed back in order).  Frames of several wavefronts (configs[4]) overlap their own wavefronts instead (ctl_comm_render_frame).
    fif = max(1, min(7, args.frames_in_flight)) if spp == batch else 1
    tracer.setParameter("FramesInFlight", fif)

    def read_back_frame():
        if rank == 0:
            # what an application reads per frame: the image after the (default) image pipeline, as in the reference's
            # applyImagePipeline -> RGBCOL (Kernel/ImagePipeline/ImagePipeline.cu:54-63); PixelData stays on the device
            tracer.resolveSRGB8Device(d_rgba.data_ptr())
            host_rgba.copy_(d_rgba, non_blocking=True)

    def frame(read_back):
        # spp passes on this rank's tiles + (N > 1) the one NCCL reduce of the accumulator to rank 0, all on `stream`: ctl_comm_render_frame
        tracer.commRenderFrame(spp, batch, tile, 0)
        if read_back:
            read_back_frame()

    def frames(n, read_back, flush_l2):
        """n steps.  fif == 1: one after the other; else as a pipeline with `fif` frames in flight -- every step submits one frame and (once the pipeline is full)
        acquires the oldest one (and reads it back); the pipeline is drained before returning, so all n frames are complete inside the caller's bracket."""
        for i in range(n):
            if flush_l2:
                flush.zero_()   # L2 flush between timed iterations (on `stream`; lanes fork from it)
            if fif == 1:
                frame(read_back)
                continue
            tracer.commSubmitFrame(spp, batch, tile, 0)
            if i >= fif - 1:
                tracer.acquireFrame()
                if read_back:
                    read_back_frame()
        while fif > 1 and tracer.framesInFlight():
            tracer.acquireFrame()
            if read_back:
                read_back_frame()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    clocks = ClockSampler(local)
    if primary:
        clocks.start()   # before the warm-up: short timed regions (8 GPUs) still collect samples under load
    # ---- warm-up
    r0 = tracer.getTotalRays()
    frames(warmup, False, False)
    rays_per_frame_local = (tracer.getTotalRays() - r0) // max(1, warmup)  # frames are identical (a new trace restarts the sample stream)
    sync_all()

    # ---- device-timed steps (value)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps if fif == 1 else 1)]
    sync_all()
    t_wall0 = time.perf_counter()
    if fif == 1:
        for i in range(steps):
            flush.zero_()  # L2 flush between timed iterations, outside the event bracket
            ev[i][0].record(stream)
            frame(False)
            ev[i][1].record(stream)
    else:   # the pipeline: ONE event bracket over exactly `steps` frames, fill and drain (and the L2 flushes) inside it
        ev[0][0].record(stream)
        frames(steps, False, True)
        ev[0][1].record(stream)
    sync_all()
    t_wall = time.perf_counter() - t_wall0
    clk = clocks.stop() if primary else None
    step_ms = [a.elapsed_time(b) for a, b in ev]
    dev_ms = float(sum(step_ms))
    t = torch.tensor([dev_ms, float(rays_per_frame_local)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_ms, rays_frame = float(tmax[0]), float(tsum[1])
    else:
        rays_frame = float(t[1])
    value = rays_frame * steps / (dev_ms * 1e-3) / 1e6
    launches_per_batch = tracer.stageTimes()[1]

    # ---- end-to-end steps through the public API with host buffers
    e2e_steps = max(2, min(steps, 20))
    tracer.setParameter("DeviceSampleTables", 0)   # e2e: the step's inputs (the pass sample tables) come from the host, like the reference's UpdateKernel
    frames(fif, True, False)   # untimed: every slot of the pipeline allocates its pinned table sets on first use
    sync_all()
    t0 = time.perf_counter()
    frames(e2e_steps, True, False)
    sync_all()
    e2e_s = time.perf_counter() - t0
    tracer.setParameter("DeviceSampleTables", 1)
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = rays_frame * e2e_steps / float(te[0]) / 1e6
    img_mean = float(host_rgba.view(h, w, 4)[:, :, :3].float().mean()) if rank == 0 else 0.0

    # ---- roofline of the traversal kernel (this rank's share), live CUDA-event stage times of the same batched frames
    tracer.setParameter("StageTimers", 1)
    ext_ms = sh_ms = 0.0
    n_batches = spp // batch
    for b_i in range(n_batches):
        render_pass(b_i, b_i == 0)
        tracer.synchronize()
        ms, _ = tracer.stageTimes()
        ext_ms += ms[1]; sh_ms += ms[3]
    tracer.setParameter("StageTimers", 0)
    stage_last = ms
    tracer.setInstrumented(1)
    render_pass(0, True)
    tracer.synchronize()
    e_cnt, s_cnt = tracer.visitCounts()
    tracer.setInstrumented(0)
    bytes_batch = traversal_bytes(e_cnt, e_cnt[3]) + traversal_bytes(s_cnt, s_cnt[3])   # one batch (all batches of a frame are alike up to RNG)
    trav_ms_batch = (ext_ms + sh_ms) / n_batches
    fused = bool(tracer.getParameter("FuseTraversal"))
    n_trav_launches = depth + 1 if fused else 2 * depth   # fused: ext(0), [shadow(b-1)+ext(b)] x (depth-1), shadow(depth-1)
    peak, peak_src = hbm_peak()
    achieved = bytes_batch / (trav_ms_batch * 1e-3) / 1e9
    tk = tracer.getParameter("TraversalKernel")
    n_rays_b = max(1, e_cnt[3] + s_cnt[3])
    roof = {"bound": "hbm", "kernel": {0: "k_intersect / k_intersect_fused", 1: "k_intersect_simple", 2: "k_intersect_staged"}.get(tk, "?") + " (all traversal launches of a wavefront)",
            "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "peak_source": peak_src, "traffic": None,
            "algorithmic_bytes_per_launch": bytes_batch / n_trav_launches, "avg_launch_ms": trav_ms_batch / n_trav_launches,
            "bytes_per_ray": bytes_batch / n_rays_b, "launches_per_batch": n_trav_launches, "passes_per_batch": batch,
            "visits_per_ray": {"inner_nodes": (e_cnt[0] + s_cnt[0]) / n_rays_b, "triangle_tests": (e_cnt[1] + s_cnt[1]) / n_rays_b, "instance_entries": (e_cnt[2] + s_cnt[2]) / n_rays_b},
            "traversal_share_of_batch": (stage_last[1] + stage_last[3]) / max(1e-9, sum(stage_last)),
            "stage_ms_last_batch": {"generate": stage_last[0], "extension": stage_last[1], "shade": stage_last[2], "shadow": stage_last[3], "finish": stage_last[4]},
            "note": "algorithmic bytes in the reference's record sizes (SURVEY 8d); the BVH is L1/L2-resident, so this is a traversal rate in bytes, not DRAM utilisation -- `traffic` is the measured DRAM side"}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(tp) as f:
            tj = json.load(f).get(name, {})
        roof["traffic"] = tj.get("dram_bytes_per_launch"); roof["traffic_source"] = tj.get("source")
    except Exception:
        pass

    # ---- CPU baseline (rank 0, N == 1, primary workload only): bounded crop on the host cores
    cpu = None
    if rank == 0 and world == 1 and primary and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        frac = args.cpu_frac if name != "cornell" else 1.0
        window = crop_window(w, h, frac)
        ckind, crays, cdt = cpu_reference_run(scene.view, w, h, depth, window, 1, threads)   # probe: 1 pass on the small crop
        n_p = 1
        if name != "cornell" and cdt < 8.0:
            # size the sample for ~12 s of CPU work: all spp passes on a centre crop of the matching size
            target_rays = 12.0 * crays / max(cdt, 1e-3)
            frac2 = min(1.0, target_rays / (crays / frac * spp))
            if frac2 > frac:
                window, n_p = crop_window(w, h, frac2), spp
            else:
                n_p = int(min(spp, max(1, round(target_rays / crays))))
            ckind, crays, cdt = cpu_reference_run(scene.view, w, h, depth, window, n_p, threads)
        cpu = {"value": crays / cdt / 1e6, "unit": "Mrays/s", "cores": threads, "kind": ckind,
               "sample": f"{window[2]-window[0]}x{window[3]-window[1]} centre crop of the {w}x{h} image, {n_p} pass(es), depth {depth}, {crays} rays in {cdt:.2f} s"}

    line = None
    if rank == 0:
        line = {
            "metric": baseline_metric(), "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": dev_ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": static_config(name, scene, world, tile),
            "rays_per_step": rays_frame, "frames_in_flight": fif,
            "timing": ("CUDA events on the launching stream around each step, summed" if fif == 1 else
                       f"one CUDA-event bracket on the launching stream over all {steps} steps: the steps run as a pipeline of {fif} frames in flight (fill and drain inside the bracket); ms_per_step = bracket / steps"),
            "passes_per_wavefront": batch, "wavefront_lanes": (tracer.getParameter("OverlapLanes") if tracer.getParameter("OverlapWavefronts") and spp > batch else 1),
            "scene_level": {"leaves": int(scene.view.n_nodes), "re_braided": bool(scene.view.node_alias)},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": spp * table_bytes, "d2h_bytes_per_step": h * w * 4,
                    "steps": e2e_steps, "note": "wall clock, the same pipeline of frames in flight as `value`; DeviceSampleTables=0: sample tables generated by the host XORWOW twin and copied H2D from pinned memory for every pass, the frame resolved by ctl_resolve_srgb8 (default image pipeline) and the RGBA8 image copied D2H to pinned memory every step"},
            "gpu_launches": int((launches_per_batch + 1) * (spp // batch) * steps),
            "wall_s_timed_region": t_wall, "image_mean_srgb8": img_mean, "roofline": roof,
        }
        if cpu:
            line["cpu_baseline"] = cpu
    tracer.close()
    del accum, flush, d_rgba
    torch.cuda.empty_cache()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS), help="default: configs[3], the 1 M-triangle scene the north star's targets are stated on")
    ap.add_argument("--cpu-frac", type=float, default=1.0 / 16, help="fraction of the image the CPU baseline renders per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra.configs lines (configs[1], [2], [4] on one GPU)")
    ap.add_argument("--ref-seconds", type=float, default=100.0, help="--impl reference: host seconds the whole run may take (sizes the per-step sample)")
    ap.add_argument("--ref-own-tree", action="store_true", help="--impl reference: mesh trees from the reference's own SplitBVHBuilder")
    ap.add_argument("--sort-mode", type=int, default=None)
    ap.add_argument("--set", action="append", default=[], metavar="KEY=INT", help="extra tracer parameter (ctl_set_param_i), for A/B runs")
    ap.add_argument("--tile", type=int, default=0, help="tile edge for the multi-GPU partition (0 = package default)")
    ap.add_argument("--frames-in-flight", type=int, default=3, help="one-wavefront frames (spp == batch): steps in flight at once (1 = every step finishes before the next starts)")
    ap.add_argument("--batch", type=int, default=0, help="progressive passes fused into one wavefront (must divide spp; 0 = 8, or 16 for frames of >= 32 passes)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3  # timing rule: W >= 3
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from cudatracerlib_b200 import build
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    stream = torch.cuda.Stream(device=dev)  # a real (non-default) stream: the tracer, NCCL and the timing events all use it
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0

    line = measure_workload(args.workload, args, torch, dist, world, rank, local, dev, stream, args.steps, args.warmup, True)
    # the other single-GPU configurations of BASELINE.json (configs[1], [2], [4]) next to the headline: shorter runs, same measurement
    if world == 1 and not args.no_extra and args.workload == "c4":
        extra = {}
        for name, st in (("c2", 5), ("c3", 5), ("c5", 2)):
            x = measure_workload(name, args, torch, dist, world, rank, local, dev, stream, st, 3, False)
            extra[name] = {k: x[k] for k in ("value", "unit", "steps", "warmup", "ms_per_step", "config", "rays_per_step", "frames_in_flight", "scene_level", "e2e", "roofline", "image_mean_srgb8")}
        line["extra"] = {"configs": extra}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
