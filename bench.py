#!/usr/bin/env python
"""Headline benchmark: frames/s of the rasterization hot path on BASELINE.json configs[1]
(1920x1080, createSphere(100, 501, 1000) = 1,000,000 triangles, one directional light,
Blinn-Phong, untextured), one turntable frame per step.

  python bench.py --gpus N --steps K --warmup W            the CUDA path (this repo)
  python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU renderer

N > 1 (under torchrun, one rank per GPU): the multi-view batch of BASELINE.json configs[4] —
every rank owns a replica of the scene and renders its own views; no data-path collective
(weak scaling). The timed region is bracketed by a barrier + device sync, per-rank device time is
measured with CUDA events on the stream the kernels run on (recorded by the library immediately
around each frame's launches, mr_set_timing_events), and the job time is the max over ranks. Before
the timed region every rank renders untimed for a quarter of a second on top of the W warm-up steps,
so that no GPU is still at idle clocks.

One JSON line on stdout (rank 0). `value` = frames/s with the scene resident in HBM and the L2
flushed between timed steps; `e2e` = frames/s through the drop-in Renderer API with host buffers
(per step: per-frame tables H2D, render, float RGB image D2H into page-locked host memory).
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]

import numpy as np  # noqa: E402

WIDTH, HEIGHT = 1920, 1080
LAT, LON = 501, 1000  # createSphere(100, 501, 1000): 1,000,000 triangles, 500,002 vertices
METRIC = "frames_per_sec_1080p_1Mtri"
WORKLOAD = "configs[1]: 1920x1080, createSphere(100,501,1000) = 1,000,000 triangles / 500,002 vertices, smooth normals, untextured, directional light, Blinn-Phong (shininess 12)"


def algorithmic_bytes(n_pos, n_nrm, n_uv, n_tri, index_arrays, w, h):
    """every input read once, every output pixel written once."""
    return 12 * n_pos + 12 * n_nrm + 8 * n_uv + 12 * n_tri * index_arrays + 16 * w * h


def issue_roofline(summary, sm_mhz, n_sm, ms_per_step):
    """The frame against the SMs' instruction issue rate - the roof that binds this pipeline (its kernels use 10-25 % of
    the DRAM bandwidth and half of their issue slots): warp instructions the two kernels execute per frame (ncu
    `smsp__inst_executed.sum` of the committed capture, same kernel source) / (SMs x 4 schedulers x SM clock). None
    if the capture does not carry the counts."""
    try:
        n = sum(float(summary[k]["warp_instructions"]) for k in ("k_geom", "k_raster"))
        per_s = float(n_sm) * 4.0 * float(sm_mhz) * 1e6
        t_min = n / per_s
        return {"bound": "issue slots", "warp_instructions_per_frame": n, "peak": per_s, "unit": "warp instructions/s",
                "achieved": n / (ms_per_step * 1e-3), "frac": t_min / (ms_per_step * 1e-3), "t_min_us": t_min * 1e6,
                "note": "time the frame's executed warp instructions need at one instruction per scheduler per cycle; the HBM time of "
                        "the frame's algorithmic bytes (roofline / frame_roofline) is about half of it"}
    except Exception:
        return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled while the render loops run: NVML polled every millisecond from
    a thread (the timed region of the default run is ~8 ms long, too short for `nvidia-smi -lms`), with
    `nvidia-smi` every 20 ms as the fallback when NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, uuid=None):
        self.index, self.uuid, self.rows, self.proc, self.nvml, self.stopping = index, uuid, [], None, None, False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + self.uuid).encode()) if self.uuid else None
            except Exception:
                h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.nvml = (pynvml, h, float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv, h, mx = self.nvml
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
        while not self.stopping:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                self.rows.append([time.perf_counter(), sm, mx, 0.0] + ["Active" if bits & b else "Not Active" for _, b in names])
            except Exception:
                pass
            time.sleep(0.001)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.perf_counter()] + [x.strip() for x in line.split(",")])

    def stop(self, t_load0=None, t0=None, t1=None):
        """Samples between t_load0 (start of the untimed render loop that precedes the timed region) and
        t1 (end of the timed region) are 'under load'; those between t0 and t1 fell inside the timed region."""
        self.stopping = True
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        elif not self.nvml:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi / NVML unavailable"]}
        sm, mx, reasons, inside = [], [], set(), []
        for r in list(self.rows):
            t, r = r[0], r[1:]
            if t_load0 is not None and not (t_load0 <= t <= t1):
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            if t0 is not None and t0 <= t <= t1:
                inside.append(float(r[0]))
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if str(v).lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(inside)) if inside else (float(np.median(sm)) if sm else None), "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_in_timed_region": len(inside),
                "source": "NVML polled every ms" if self.nvml else "nvidia-smi -lms 20",
                "window": "sm_mhz = median of the samples inside the timed region (of all samples under load if none fell inside); "
                          "reasons over the untimed quarter-second render loop and the timed region"}


def config_dict():
    """The same dictionary in both arms (the driver compares them key by key)."""
    return {"workload": WORKLOAD, "width": WIDTH, "height": HEIGHT, "triangles": 2 * LON * (LAT - 1), "data": "synthetic"}


def build_scene(be, frame=0):
    from minirender_b200 import scenes
    return scenes.sphere_scene(be, WIDTH, HEIGHT, LAT, LON, frame=frame)


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU renderer (oracle/_ref when present, else the C port)
# ------------------------------------------------------------------------------------------------
def _ref_worker(args):
    kind, warm, frames = args
    import minirender_b200 as m
    from minirender_b200 import scenes
    import pyoracle
    if kind == "reference":
        be = m.Backend(pyoracle.REF_PATH)
        setup = build_scene(be)
        r = setup.apply(m.Renderer(be))
        for i in range(warm):
            r.render()
        t0 = time.perf_counter()
        for i in frames:
            r.set_view(scenes.sphere_view(be, i))
            r.render()
        return time.perf_counter() - t0
    be = m.Backend()  # host-side flatten only; pixels come from the C restatement
    setup = build_scene(be)
    r = setup.apply(m.Renderer(be))
    out = None
    t0 = time.perf_counter()
    for k, i in enumerate([0] * warm + list(frames)):
        if k == warm:
            t0 = time.perf_counter()
        r.set_view(scenes.sphere_view(be, i))
        r.prepare()
        out = pyoracle.render_port(r.scene_desc_ptr(), r.frame_desc_ptr(), WIDTH, HEIGHT, into=out)
    return time.perf_counter() - t0


def cpu_reference_run(steps, warmup, procs):
    """Each step = one frame per worker process, `procs` workers in parallel (the reference is
    single-threaded; frames are independent, so this is all the host parallelism it can use)."""
    import multiprocessing as mp
    import pyoracle
    kind = "reference" if pyoracle.have_ref() else "port"
    if kind == "port":
        pyoracle.port()
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        # every worker builds the scene, renders `warmup` untimed frames, then times its own frames;
        # the job time is the slowest worker's timed loop (scene construction is not timed)
        times = pool.map(_ref_worker, [(kind, warmup, [p + procs * k for k in range(steps)]) for p in range(procs)])
    return kind, max(times)


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, cores)  # every host thread: one single-threaded reference renderer per core
    # every step = one frame per worker (~0.1-0.2 s each): --steps / --warmup are honoured up to a bound that keeps the
    # run within a few minutes; the line reports what actually ran
    steps = max(1, min(args.steps, 400))
    warm = max(1, min(args.warmup, 10))
    kind, t = cpu_reference_run(steps, warm, procs)
    fps = procs * steps / t
    n_tri = 2 * LON * (LAT - 1)
    line = {"metric": METRIC, "value": fps, "unit": "frames/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": 1000.0 * t / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "mtri_per_s": fps * n_tri / 1e6, "mpix_per_s": fps * WIDTH * HEIGHT / 1e6,
            "config": config_dict(),
            "notes": {"arm": "the reference's own src/Renderer.cpp (oracle/_ref) on the host: one frame per worker process per step, %d "
                             "single-threaded workers; warm-up frames per worker: %d" % (procs, warm)},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": procs, "kind": kind,
                             "sample": "%d frames per worker x %d workers" % (steps, procs)},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import torch
    import minirender_b200 as m
    from minirender_b200 import cabi, scenes

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    lib = cabi.load()
    try:
        uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        uuid = None
    sampler = ClockSampler(local_rank, uuid)
    if rank == 0:
        sampler.start()  # up and sampling long before the render loops begin
    be = m.Backend()
    setup = build_scene(be)
    r = setup.apply(m.Renderer(be))
    r.set_device(local_rank)
    ctx = r.context_ptr()
    # The kernels must run on the stream the events are recorded on. torch's *default* stream has
    # handle 0, which mr_set_stream reads as "use the context's own stream", so make a real one.
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    assert lib.mr_set_stream(ctx, C.c_void_p(stream.cuda_stream)) == 0

    def view_of(step):  # rank-interleaved turntable views
        return scenes.sphere_view(be, rank + world * step)

    K, W = max(1, args.steps), max(3, args.warmup)
    # ---- warm-up (also uploads the scene and settles queue sizes) ----
    for i in range(W):
        r.set_view(view_of(i))
        r.render()
    r.synchronize()
    # W frames are a fraction of a millisecond of GPU work: keep rendering (untimed) for a quarter of a second
    # so that every rank's GPU has left its idle clocks before the timed region (seen at N = 8: a GPU that had
    # been idle reported 83 us per frame against 65 us on its neighbours)
    t_ramp = time.perf_counter()
    n_ramp = 0
    while time.perf_counter() - t_ramp < 0.25:
        for i in range(64):
            r.set_view(view_of(i))
            r.render()
        r.synchronize()
        n_ramp += 64

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed: device time per step, L2 flushed (untimed) before every step ----
    # The events are recorded by the library itself, immediately before the frame's first and after its last
    # kernel launch (mr_set_timing_events): the bracket holds the three kernels and none of this loop's host
    # work, so a rank whose Python thread is descheduled for a while (8 ranks share the box's cores) does not
    # report host time as device time.
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    for e in starts + stops:
        e.record(stream)  # creates the underlying cudaEvent_t
    barrier()
    wall0 = time.perf_counter()
    for i in range(K):
        r.set_view(view_of(W + i))
        lib.mr_flush_l2(ctx)
        assert lib.mr_set_timing_events(ctx, C.c_void_p(starts[i].cuda_event), C.c_void_p(stops[i].cuda_event)) == 0
        r.render()
    barrier()
    wall_dev_loop = time.perf_counter() - wall0
    clocks = sampler.stop(t_ramp, wall0, time.perf_counter()) if rank == 0 else None
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, stops)]
    dev_ms = sum(step_ms)

    # ---- the same K steps back to back without the flush (warm L2, frames pipelined on the stream):
    # what a turntable loop that keeps its images on the device sees. Reported as an extra key. ----
    barrier()
    w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0.record(stream)
    for i in range(K):
        r.set_view(view_of(W + i))
        r.render()
    w1.record(stream)
    barrier()
    warm_ms = w0.elapsed_time(w1)

    # ---- per-kernel times (CUDA events between the stages, same stream, L2 flushed) ----
    lib.mr_set_debug(ctx, 2)
    r.prepare()
    assert lib.mr_profile_frame(ctx, r.frame_desc_ptr(), min(K, 20)) == 0, lib.mr_last_error(ctx)
    lib.mr_set_debug(ctx, 0)
    st = cabi.Stats()
    lib.mr_get_stats(ctx, C.byref(st))
    stage_ms = {"k_geom": st.ms_kernel[1], "k_raster": st.ms_kernel[4], "frame": st.ms_kernel[5]}

    # ---- end to end through the drop-in API: H2D tables, render, D2H float image ----
    # (a) the reference's own call sequence: setView, render(), getImage() — the copy blocks
    r.image_view()
    barrier()
    t0 = time.perf_counter()
    checksum = 0.0
    for i in range(K):
        r.set_view(view_of(W + K + i))
        r.render()
        img = r.image_view()  # Renderer::getImage(): blocking D2H into its page-locked host mirror
        checksum += float(img[HEIGHT // 2, WIDTH // 2, 0])
    barrier()
    e2e_block_s = time.perf_counter() - t0
    # (b) the same steps with two output slots: the D2H copy of frame i (copy stream, page-locked
    # destination) overlaps the kernels of frame i+1; every frame's image still lands on the host and
    # is read there before its buffer is reused
    host = torch.empty((2, HEIGHT, WIDTH, 3), dtype=torch.float32, pin_memory=True)
    hp = [C.cast(C.c_void_p(host[j].data_ptr()), cabi.F32P) for j in range(2)]
    assert lib.mr_set_output_slots(ctx, 2) == 0, lib.mr_last_error(ctx)
    for i in range(2):
        r.set_view(view_of(i)); r.render()
    r.synchronize()
    barrier()
    tick = [C.c_int(0), C.c_int(0)]
    t0 = time.perf_counter()
    for i in range(K):
        r.set_view(view_of(W + K + i))
        r.render()
        assert lib.mr_read_image_begin(ctx, hp[i & 1], C.byref(tick[i & 1])) == 0
        if i > 0:
            assert lib.mr_read_wait(ctx, tick[(i - 1) & 1]) == 0
            checksum += float(host[(i - 1) & 1, HEIGHT // 2, WIDTH // 2, 0])
    assert lib.mr_read_wait(ctx, tick[(K - 1) & 1]) == 0
    checksum += float(host[(K - 1) & 1, HEIGHT // 2, WIDTH // 2, 0])
    barrier()
    e2e_full_s = time.perf_counter() - t0
    # (b') the same loop, the host images kept as persistent mirrors: mr_read_image_dirty_begin copies only the rectangle
    # the new frame and the frame the buffer held may have drawn into (k_geom tracks it; the rest is background on both
    # sides). Every frame's complete float image is still on the host, bit-identical to a full read (checked below).
    for i in range(2):
        r.set_view(view_of(i)); r.render()
        assert lib.mr_read_image_dirty_begin(ctx, hp[i & 1], C.byref(tick[i & 1])) == 0, lib.mr_last_error(ctx)
        assert lib.mr_read_wait(ctx, tick[i & 1]) == 0
    barrier()
    d2h_total = 0
    t0 = time.perf_counter()
    for i in range(K):
        r.set_view(view_of(W + K + i))
        r.render()
        assert lib.mr_read_image_dirty_begin(ctx, hp[i & 1], C.byref(tick[i & 1])) == 0
        lib.mr_get_stats(ctx, C.byref(st))
        d2h_total += int(st.d2h_bytes)
        if i > 0:
            assert lib.mr_read_wait(ctx, tick[(i - 1) & 1]) == 0
            checksum += float(host[(i - 1) & 1, HEIGHT // 2, WIDTH // 2, 0])
    assert lib.mr_read_wait(ctx, tick[(K - 1) & 1]) == 0
    checksum += float(host[(K - 1) & 1, HEIGHT // 2, WIDTH // 2, 0])
    barrier()
    e2e_s = time.perf_counter() - t0
    full_img = torch.empty((HEIGHT, WIDTH, 3), dtype=torch.float32)
    assert lib.mr_read_image(ctx, C.cast(C.c_void_p(full_img.data_ptr()), cabi.F32P)) == 0
    mirror_identical = bool(torch.equal(full_img.view(torch.int32), host[(K - 1) & 1].view(torch.int32)))
    assert lib.mr_set_output_slots(ctx, 1) == 0
    lib.mr_get_stats(ctx, C.byref(st))
    h2d = int(st.h2d_bytes)
    d2h_full = WIDTH * HEIGHT * 12
    d2h = d2h_total // max(K, 1)
    # (c) the image in the format the reference writes to disk: savePPM's 8-bit truncation (io.cpp:358-361)
    # done on the device, 3 bytes per pixel over PCIe instead of 12 (mr_read_rgb8, blocking)
    host8 = torch.empty((HEIGHT, WIDTH, 3), dtype=torch.uint8, pin_memory=True)
    hp8 = C.c_void_p(host8.data_ptr())
    assert lib.mr_read_rgb8(ctx, hp8) == 0
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        r.set_view(view_of(W + K + i))
        r.render()
        assert lib.mr_read_rgb8(ctx, hp8) == 0
        checksum += float(host8[HEIGHT // 2, WIDTH // 2, 0])
    barrier()
    e2e_rgb8_s = time.perf_counter() - t0

    # (d) what the box can do: every rank copies image-sized device buffers into its page-locked host memory at the same
    # time, nothing else running. The e2e figures above are bounded by this (PCIe / host memory), not by the kernels.
    dimg = torch.empty((HEIGHT, WIDTH, 3), dtype=torch.float32, device="cuda")
    for i in range(3):
        host[i & 1].copy_(dimg, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        host[i & 1].copy_(dimg, non_blocking=True)
    torch.cuda.synchronize()
    d2h_only_s = time.perf_counter() - t0
    barrier()
    del dimg

    t = torch.tensor([dev_ms, e2e_s * 1000.0, warm_ms, e2e_block_s * 1000.0, e2e_rgb8_s * 1000.0, d2h_only_s * 1000.0, e2e_full_s * 1000.0], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max, warm_ms_max, e2e_block_ms_max, e2e_rgb8_ms_max, d2h_only_ms_max, e2e_full_ms_max = (float(x) for x in t)
    per_rank = [dev_ms / K]
    if dist is not None:
        mine = torch.tensor([dev_ms / K], dtype=torch.float64, device="cuda")
        everyone = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(everyone, mine)
        per_rank = [float(x[0]) for x in everyone]

    strips = None
    if world > 1 and not args.no_strips:
        # BASELINE.json configs[2] in the same run: one 4K frame of 10 M textured triangles, strips sharded over the ranks
        r.synchronize()
        strips = measure_strips(args, rank, local_rank, world, dist, min(K, 50), 5)

    if rank == 0:
        n_tri = int(st.triangles_in)
        fps = world * K / (dev_ms_max / 1000.0)
        e2e_fps = world * K / (e2e_ms_max / 1000.0)
        peak, peak_src = measured_peak()
        n_pos = (LAT - 1) * LON + 2
        frame_bytes = algorithmic_bytes(n_pos, n_pos, 0, n_tri, 2, WIDTH, HEIGHT)
        raster_bytes = 16 * WIDTH * HEIGHT  # image + depth written once by the dominant kernel
        dom = max(("k_geom", "k_raster"), key=lambda k: stage_ms[k])
        # each kernel's share of the frame's algorithmic bytes: k_geom reads positions, normals and both
        # index arrays, k_raster writes image + depth
        dom_bytes = {"k_geom": 24 * n_pos + 24 * n_tri, "k_raster": raster_bytes}[dom]
        ach = dom_bytes / (stage_ms[dom] * 1e-3) / 1e9
        # dram__bytes of the dominant kernel from the committed ncu capture - only if that capture was made on the kernel
        # source this run executes (profiles/ncu_summary.json records the file's hash: tools/make_ncu_summary.py)
        traffic, traffic_note, frame_traffic = None, "no ncu capture of this kernel source (profiles/ncu_summary.json is missing or older than mr_kernels.cu)", None
        fresh_summary = None
        try:
            import hashlib
            with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
                summ = json.load(f)
            with open(os.path.join(ROOT, "minirender_b200", "csrc", "mr_kernels.cu"), "rb") as f:
                fresh = summ.get("kernels_sha256") == hashlib.sha256(f.read()).hexdigest()
            if fresh:
                fresh_summary = summ
                traffic = summ.get(dom, {}).get("dram_bytes_per_launch")
                frame_traffic = summ.get("frame", {}).get("dram_bytes_per_frame_no_flush")
                traffic_note = summ.get(dom, {}).get("source")
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
            "extra_warmup_frames": n_ramp,
            "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "mtri_per_s": fps * n_tri / 1e6, "mpix_per_s": fps * WIDTH * HEIGHT / 1e6,
            "config": config_dict(),
            "notes": {"l2": "flushed (256 MiB write + 256 MiB read) before every timed step, outside the timed events",
                      "multi_gpu": "view batch: rank r renders views r, r+N, ... of a scene replica; no collective on the data path",
                      "parity": "depth/coverage/winner ids bit-exact, RGB <= 1 LSB vs the reference (tests/test_gpu_fullsize.py)"},
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "host_image_identical_to_full_read": mirror_identical,
                    "note": "per step: setView, render(), the frame's float RGB image brought up to date in page-locked host memory and read "
                            "there (mr_read_image_dirty_begin / mr_read_wait: the two host buffers persist, only the rectangle this frame and "
                            "the frame the buffer held may have drawn into crosses PCIe, the rest is background on both sides); two output "
                            "slots, so the copy of frame i overlaps the kernels of frame i+1"},
            "e2e_full_copy": {"value": world * K / (e2e_full_ms_max / 1000.0), "unit": "frames/s", "d2h_bytes_per_step": d2h_full,
                              "note": "the same loop copying the whole 24.9 MB image every step (mr_read_image_begin): bound by d2h_ceiling"},
            "e2e_blocking": {"value": world * K / (e2e_block_ms_max / 1000.0), "unit": "frames/s",
                             "note": "the reference's call sequence setView + render() + getImage(): the D2H copy blocks every step"},
            "e2e_rgb8": {"value": world * K / (e2e_rgb8_ms_max / 1000.0), "unit": "frames/s", "d2h_bytes_per_step": WIDTH * HEIGHT * 3,
                         "note": "setView + render() + mr_read_rgb8: the 8-bit image of savePPM (io.cpp:358-361) quantised on the device, "
                                 "blocking D2H of 3 bytes per pixel (not the headline: the reference API returns float RGB)"},
            "d2h_ceiling": {"value": world * K / (d2h_only_ms_max / 1000.0), "unit": "frames/s",
                            "gb_per_s": world * K * d2h_full / (d2h_only_ms_max / 1000.0) / 1e9,
                            "note": "every rank copying %d-byte float images from its GPU into its own page-locked host memory at the same time, "
                                    "no rendering: the ceiling of `e2e_full_copy` on this box (PCIe + host memory); e2e_full_copy / d2h_ceiling = %.2f"
                                    % (d2h_full, (world * K / (e2e_full_ms_max / 1000.0)) / (world * K / (d2h_only_ms_max / 1000.0)))},
            "warm_l2_pipelined": {"value": world * K / (warm_ms_max / 1000.0), "unit": "frames/s", "ms_per_step": warm_ms_max / K,
                                  "note": "same K steps back to back, no L2 flush (not the headline)"},
            "gpu_launches": int(st.kernels_launched) * K,
            "kernels_per_step": int(st.kernels_launched),
            "stage_ms": stage_ms,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": traffic, "traffic_source": traffic_note, "algorithmic_bytes_per_launch": dom_bytes, "peak_source": peak_src},
            "frame_roofline": {"algorithmic_bytes_per_frame": frame_bytes, "achieved": frame_bytes / (dev_ms_max / K * 1e-3) / 1e9,
                               "unit": "GB/s", "frac": frame_bytes / (dev_ms_max / K * 1e-3) / 1e9 / peak,
                               "traffic_no_flush": frame_traffic},
            "ms_per_step_per_rank": per_rank,
            "ms_per_step_rank0": {"min": min(step_ms), "median": float(np.median(step_ms)), "max": max(step_ms)},
            "clocks": clocks,
            "host_loop_wall_s": wall_dev_loop,
            "stats": {"triangles_in": n_tri, "records": int(st.records), "bin_entries": int(st.bin_entries)},
        }
        if strips is not None:
            line["strips4k"] = strips
        try:
            import torch
            issue = issue_roofline(fresh_summary, clocks.get("sm_mhz"), torch.cuda.get_device_properties(local_rank).multi_processor_count, dev_ms_max / K) if fresh_summary else None
            if issue:
                line["issue_roofline"] = issue
        except Exception:
            pass
        if world == 1 and not args.no_cpu_baseline:
            import pyoracle
            kind = "reference" if pyoracle.have_ref() else "port"
            n = 5
            per = _ref_worker((kind, 1, list(range(1, 1 + n)))) / n
            line["cpu_baseline"] = {"value": 1.0 / per, "unit": "frames/s", "cores": 1, "kind": kind,
                                    "sample": "%d frames of the same workload, single thread (the reference is single-threaded)" % n,
                                    "host_cpus": os.cpu_count()}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# optional: BASELINE.json configs[2] — one 3840x2160 frame of a 10M-triangle textured mesh, screen
# strips sharded across the ranks (strong scaling). Not the headline line; run with
#   --workload strips4k [--gather peer|nccl]
# ------------------------------------------------------------------------------------------------
def measure_strips(args, rank, local_rank, world, dist, steps, warmup):
    """One 3840x2160 frame of the 10 M-triangle textured sphere per step, rank r rendering the rows
    sharding.strip_rows(2160, r, N) straight into rank 0's framebuffer over NVLink (tile stores through peer memory),
    ranks joined on the device (mr_stream_signal / mr_stream_wait: no collective, no host sync per frame).
    Returns the dict that goes into the bench line (rank 0) or None."""
    import torch
    import minirender_b200 as m
    from minirender_b200 import cabi, scenes, sharding

    lib = cabi.load()
    be = m.Backend()
    W4, H4, LAT4, LON4 = 3840, 2160, 2237, 2236  # createSphere(r, 2237, 2236): 9,999,392 triangles
    setup = scenes.sphere_scene(be, W4, H4, lat=LAT4, lon=LON4, textured=True, d=330.0, tex_size=2048)
    r = setup.apply(m.Renderer(be))
    r.set_device(local_rank)
    ctx = r.context_ptr()
    stream = torch.cuda.current_stream()
    assert lib.mr_set_stream(ctx, C.c_void_p(stream.cuda_stream)) == 0
    view = lambda i: scenes.sphere_view(be, i, d=330.0)
    strips = sharding.all_strips(H4, world)
    calibration = None
    if world > 1 and not getattr(args, "equal_strips", False):
        # Balanced strips: equal heights give the ranks whose rows cross the middle of the sphere twice the triangles of
        # the outer ones. A few untimed rounds on each rank's own GPU (device time of its strip, all-gathered; every
        # rank computes the same new cuts) move the cuts until the strips take about the same time.
        calibration = []
        r.set_row_range(*strips[rank])
        r.set_view(view(0))
        r.render()  # (uploads the scene)
        r.synchronize()
        for _ in range(4):
            r.set_row_range(*strips[rank])
            r.set_view(view(0))
            r.prepare()
            assert lib.mr_profile_frame(ctx, r.frame_desc_ptr(), 3) == 0, lib.mr_last_error(ctx)
            st0 = cabi.Stats()
            lib.mr_get_stats(ctx, C.byref(st0))
            t_mine = float(st0.ms_kernel[5])
            if rank == 0:
                # the gathering rank also resets the peers' rows of a finished frame to the clear values: part of its loop
                bg3 = (C.c_float * 3)(*setup.background)
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c0.record(stream)
                for k_, (b_, e_) in enumerate(strips):
                    if k_ != 0 and e_ > b_:
                        assert lib.mr_clear_rows(ctx, bg3, b_, e_) == 0
                c1.record(stream)
                c1.synchronize()
                t_mine += c0.elapsed_time(c1)
            mine = torch.tensor([t_mine], dtype=torch.float64, device="cuda")
            everyone = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(everyone, mine)
            times = [float(x[0]) for x in everyone]
            calibration.append({"rows": [e - b for b, e in strips], "strip_ms": times})
            strips = sharding.balanced_strips(H4, strips, times)
    rb, re = strips[rank]
    r.clear()
    r.synchronize()
    close = join = None
    if world > 1:
        double = not getattr(args, "single_target", False)
        if double and rank == 0:
            # two framebuffers on the gathering rank: the peers store frame i + 1 while it waits for, consumes and clears frame i
            assert lib.mr_set_output_slots(ctx, 2) == 0
            bg3 = (C.c_float * 3)(*setup.background)
            for slot in (0, 1):
                assert lib.mr_clear_rows_slot(ctx, slot, bg3, 0, H4) == 0
            r.synchronize()
        dist.barrier()
        if double:
            targets, first, close = sharding.open_peer_targets(lib, ctx, rank, world, dist, dst=0)
            join = sharding.StripJoin(lib, ctx, rank, world, dist, dst=0, targets=targets, first=first)
        else:
            close = sharding.open_peer_target(lib, ctx, rank, world, dist, dst=0)
            join = sharding.StripJoin(lib, ctx, rank, world, dist, dst=0)
    if world > 1 and rank != 0:
        assert lib.mr_set_sparse_remote_stores(ctx, 1) == 0  # only the tiles this rank draws into cross NVLink
    if world > 1 and calibration is not None and not getattr(args, "local_calibration", False):
        # Second stage of the balancing, with the stores where they go in the timed loop: a peer's tile kernel is slower
        # when its tiles cross NVLink (and the more so the more tiles it touches), which the rounds above cannot see.
        # All ranks render their strips at the same time (the peers share rank 0's NVLink ingress), unsynchronised and
        # untimed; the cuts move again. Rank 0 then resets its framebuffers and tells the peers which one comes first.
        for _ in range(3):
            r.set_row_range(*strips[rank])
            r.set_view(view(0))
            if rank != 0:
                if double:
                    assert lib.mr_set_remote_target(ctx, *targets[join.first % 2]) == 0
                assert lib.mr_set_raster_gate(ctx, None, 0) == 0
            r.prepare()
            torch.cuda.synchronize()
            dist.barrier()
            # (nine timed frames per round: with seven peers storing into one GPU a strip's time varies by +-15 % from frame to frame)
            assert lib.mr_profile_frame(ctx, r.frame_desc_ptr(), 9) == 0, lib.mr_last_error(ctx)
            st0 = cabi.Stats()
            lib.mr_get_stats(ctx, C.byref(st0))
            t_mine = float(st0.ms_kernel[5])
            if rank == 0:
                bg3 = (C.c_float * 3)(*setup.background)
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c0.record(stream)
                for k_, (b_, e_) in enumerate(strips):
                    if k_ != 0 and e_ > b_:
                        assert lib.mr_clear_rows(ctx, bg3, b_, e_) == 0
                c1.record(stream)
                c1.synchronize()
                t_mine += c0.elapsed_time(c1)
            mine = torch.tensor([t_mine], dtype=torch.float64, device="cuda")
            everyone = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(everyone, mine)
            times = [float(x[0]) for x in everyone]
            calibration.append({"rows": [e - b for b, e in strips], "strip_ms": times, "stores": "into rank 0 over NVLink"})
            strips = sharding.balanced_strips(H4, strips, times)
        torch.cuda.synchronize()
        dist.barrier()
        first = [None]
        if rank == 0:
            bg3 = (C.c_float * 3)(*setup.background)
            if double:
                for slot in (0, 1):
                    assert lib.mr_clear_rows_slot(ctx, slot, bg3, 0, H4) == 0
                first = [(lib.mr_output_slot(ctx) + 1) % 2]
            else:
                assert lib.mr_clear_rows(ctx, bg3, 0, H4) == 0
            r.synchronize()
        dist.broadcast_object_list(first, src=0)
        if double:
            join.first = first[0]
        rb, re = strips[rank]
    r.set_row_range(rb, re)
    counter = [0]
    peer_rows = [s_ for k_, s_ in enumerate(strips) if k_ != 0 and s_[1] > s_[0]]
    total_frames = warmup + steps

    views = [view(i) for i in range(total_frames)]  # (composing a view through ctypes costs the host 40 us: not per frame)

    def frame(i):
        k = counter[0]
        counter[0] += 1
        r.set_view(views[i])
        if join:
            join.begin(k)
        r.render()
        if join:
            # rank 0 resets the peers' rows to the clear values before it lets them in again (not after the last frame:
            # that one is compared with a single-GPU render below)
            join.end(k, clear_rows=(setup.background, peer_rows) if (rank == 0 and k + 1 < total_frames) else None)

    for i in range(warmup):
        frame(i)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for i in range(steps):
        frame(warmup + i)
    e1.record(stream)
    enqueue = time.perf_counter() - t0  # host time spent issuing the frames (a rank whose queue never fills is host-bound)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    if dist is not None:
        dist.barrier()
    dev_ms = e0.elapsed_time(e1)
    # per-rank device time of its own strip (stage events, L2 not flushed: the working set exceeds it)
    lib.mr_set_debug(ctx, 0)
    r.prepare()
    assert lib.mr_profile_frame(ctx, r.frame_desc_ptr(), 5) == 0, lib.mr_last_error(ctx)
    st = cabi.Stats()
    lib.mr_get_stats(ctx, C.byref(st))
    mine = torch.tensor([dev_ms, wall * 1000.0, st.ms_kernel[5], st.ms_kernel[1], st.ms_kernel[4], float(st.tiles_stored), enqueue * 1000.0 / steps],
                        dtype=torch.float64, device="cuda")
    everyone = [mine]
    if dist is not None:
        everyone = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(everyone, mine)
    # ---- the assembled frame against the same frame rendered by rank 0 alone ----
    identical = None
    last = warmup + steps - 1
    if rank == 0:
        img, dep = sharding.device_tensors(lib, ctx, H4, W4, torch.device("cuda", local_rank))
        got_i, got_d = img.clone(), dep.clone()
    if dist is not None:
        dist.barrier()
    if close:
        close()
    if join:
        join.close()
    if rank == 0:
        r.set_row_range(0, H4)
        r.set_view(view(last))
        r.render()
        r.synchronize()
        img, dep = sharding.device_tensors(lib, ctx, H4, W4, torch.device("cuda", local_rank))
        identical = bool(torch.equal(img.view(torch.int32), got_i.view(torch.int32)) and torch.equal(dep.view(torch.int32), got_d.view(torch.int32)))
        # single-GPU time of the whole frame, for the strong-scaling factor
        lib.mr_set_debug(ctx, 0)
        r.prepare()
        assert lib.mr_profile_frame(ctx, r.frame_desc_ptr(), 5) == 0
        lib.mr_get_stats(ctx, C.byref(st))
        whole_ms = float(st.ms_kernel[5])
    if dist is not None:
        dist.barrier()
    if rank != 0:
        return None
    per_rank = [[float(x) for x in t] for t in everyone]
    ms = max(max(p[0], p[1]) for p in per_rank) / steps
    nvlink_bytes = int(sum(p[5] for p in per_rank[1:]) * 256 * 16)  # tiles the peers stored x 256 pixels x (12 + 4) bytes
    nvlink_bytes_dense = sum((e - b) * W4 * 16 for r_, (b, e) in enumerate(strips) if r_ != 0)
    return {"metric": "frames_per_sec_4k_10Mtri_strips", "value": 1000.0 / ms, "unit": "frames/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms, "scaling": "strong",
            "workload": "configs[2]: 3840x2160, createSphere(100,2237,2236) = 9,999,392 triangles, 2048x2048 float texture, one strip of "
                        "whole tile rows per rank (sort-first: every rank culls all clusters, sets up the ones its rows can see)",
            "gather": "tile stores into rank 0's framebuffer over NVLink (peer memory, touched tiles only: rank 0 clears the rest itself) + device-side "
                      "join (stream-ordered flags; rank 0 alternates between two framebuffers, so the peers store frame i + 1 while it waits for, "
                      "consumes and clears frame i; a peer's tile kernel only waits for the release of frame i - 2)" if world > 1 else "none",
            "nvlink_bytes_per_frame": nvlink_bytes, "nvlink_bytes_per_frame_if_every_tile_were_sent": nvlink_bytes_dense,
            "ms_per_step_device_rank0": per_rank[0][0] / steps, "ms_per_step_host_wall_max": max(p[1] for p in per_rank) / steps,
            "strip_device_ms_per_rank": [p[2] for p in per_rank], "strip_geom_ms_per_rank": [p[3] for p in per_rank],
            "strip_raster_ms_per_rank": [p[4] for p in per_rank],
            "host_enqueue_ms_per_frame_per_rank": [p[6] for p in per_rank],
            "strip_rows_per_rank": [e - b for b, e in strips],
            "strip_balancing": calibration if calibration is not None else "equal heights",
            "single_gpu_frame_ms": whole_ms, "speedup_vs_single_gpu_frame": whole_ms / ms,
            "assembled_frame_identical_to_single_gpu": identical,
            "mtri_per_s": 9.999392 / ms * 1e3, "mpix_per_s": W4 * H4 / ms / 1e3}


def run_strips(args, rank, local_rank, world):
    import torch
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    res = measure_strips(args, rank, local_rank, world, dist, max(1, args.steps), max(3, args.warmup))
    if rank == 0:
        res.update({"higher_is_better": True, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                    "config": {"workload": res.pop("workload"), "gather": res["gather"], "l2": "working set (> 1 GB) exceeds the L2"}})
        print(json.dumps(res), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strips", action="store_true", help="N > 1: skip the configs[2] strip-sharded frame (extra key strips4k)")
    ap.add_argument("--single-target", action="store_true", help="strips: one framebuffer on the gathering rank instead of two alternating ones")
    ap.add_argument("--equal-strips", action="store_true", help="strips of equal height instead of heights balanced by measured time")
    ap.add_argument("--local-calibration", action="store_true", help="balance the strips with local stores only (without the second stage that measures them with the stores going to rank 0)")
    ap.add_argument("--workload", default="sphere1m", choices=["sphere1m", "strips4k", "turntable2m"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload == "turntable2m":
        # BASELINE.json configs[4]: the 2M-triangle mesh of the 1024-view turntable batch; same code
        # path as the headline, views rank-interleaved (no data-path collective). Not the headline.
        global LAT, LON, WORKLOAD, METRIC
        LAT, LON = 1001, 1000
        WORKLOAD = ("configs[4]: turntable batch, 1920x1080, createSphere(100,1001,1000) = 2,000,000 triangles / 1,000,002 vertices, "
                    "views k = rank (mod N), smooth normals, untextured, directional light, Blinn-Phong")
        METRIC = "frames_per_sec_1080p_2Mtri_turntable"
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.workload == "strips4k":
        run_strips(args, rank, local_rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
