#!/usr/bin/env python3
"""bench.py — path samples/s of the wavefront path tracer on cbox (BASELINE configs C2 / C4).

Contract: `python bench.py --gpus N --steps K --warmup W` (torchrun for N > 1) prints ONE JSON line on rank 0.
  workload  c2 (default) = cbox 1280x720 @ 1024 spp, the configuration BASELINE.json quotes the metric on;
            c4 = cbox 4096x4096 @ 4096 spp (scenes/cbox/pt.json as the reference ships it): the multi-wave path,
            16.7 M pixels, 68.7 G samples per step — run it with --steps 1 --warmup 1
  step      = one full render of the workload (spp / 64 passes of 64 spp, pt.rs:1126-1149)
  value     = path samples / s, whole job, scene + sampler tables already resident in HBM, film left on device
              (N > 1: resolved on device and all-gathered over NCCL inside the timed region)
  e2e       = same metric through the public call with HOST buffers: scene upload (H2D, incl. host BVH build),
              sampler-table upload (H2D), render, resolve, NCCL gather (N > 1) and the D2H copy of the full frame on
              rank 0 — all inside the timed region
  roofline  = HBM roofline of the dominant kernel: algorithmic bytes per launch / CUDA-event time
  cpu_baseline = the CPU oracle (restatement of the reference's `-d cpu` path) on a bounded sample
`--impl reference` times the CPU oracle instead (the reference binary cannot be built here: DESIGN.md).
Multi-GPU: image rows interleaved in small blocks over the ranks (akari_render_b200/sharding.py), one rank per GPU,
scene replicated, no collective on the hot path; one NCCL all_gather assembles the HDR image at the end of every step.
Outside the timed region rank 0 re-renders one pass of the whole frame alone and checks that the gathered frame is
bit-identical (`gather_check`).
"""
import argparse
import json
import os
import subprocess
import sys
import threading

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {"c2": (1280, 720, 1024), "c4": (4096, 4096, 4096)}
# packed record sizes -> algorithmic bytes per unit of work
B_SAMPLE = 168                                # raygen W RAY + PATH + idx (120), accumulate R L + idx + film RMW (48)
B_TRACE_PER_SEG, B_SHADE_PER_SEG = 72, 228    # trace: R RAY + W HIT + idx; shade: R HIT + wo + PATH, W PATH + RAY + key/idx
B_TRACE_PER_SHADOW, B_SHADE_PER_SHADOW = 88, 64  # only where a shadow QUEUE exists (queued pipeline): shade W 64, trace R 64 + RMW 24
# what the fused pipeline moves by design: 64-byte record in + out per shaded hit, accumulators + film per sample
B_FUSED_PER_HIT, B_FUSED_PER_SAMPLE = 128, 96


def sample_clocks(stop, out):
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md)."""
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", os.environ.get("LOCAL_RANK", "0")],
                               capture_output=True, text=True, timeout=5)
            for line in r.stdout.strip().splitlines():
                f = [x.strip() for x in line.split(",")]
                if len(f) >= 6:
                    out.append(f)
        except Exception:
            pass
        stop.wait(0.2)


def summarize_clocks(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
    sm = sorted(int(s[0]) for s in samples if s[0].isdigit())
    mx = max(int(s[1]) for s in samples if s[1].isdigit())
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in samples)]
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons}


def workload_config(args, world, block_rows=None):
    w, h, spp = WORKLOADS[args.workload]
    spp = args.spp or spp
    scene = "cbox" if args.scene == "cbox" else "cbox + 8.5 K-triangle clutter (BVH path; documented stand-in for the absent classroom asset, not a BASELINE config)"
    cfg = {"workload": f"{scene} {w}x{h} @ {spp}spp, pmj02bn seed 0, gaussian r=1.5, max_depth 12, rr_depth 5, 64 spp/pass"
                       + (" (BASELINE config C2)" if args.workload == "c2" and spp == 1024 and args.scene == "cbox" else "")
                       + (" (BASELINE config C4)" if args.workload == "c4" and spp == 4096 and args.scene == "cbox" else ""),
           "parallelism": f"image rows interleaved x{world}" + (f" in blocks of {block_rows}" if block_rows else ""),
           "l2": "working set (wave state) > L2; no inter-step reuse (film cleared each step)",
           "wave_paths": args.wave or (1 << 22), "trace_mode": args.trace_mode, "pipeline": "queued" if args.fused == 2 else "auto (fused on flat scenes)"}
    if args.scene != "cbox":
        cfg["scene"] = args.scene
    return cfg


REF_SPP_PER_STEP = 4


def best_cpu_threads(oracle, scene, task, pmj, bn, width, height):
    """The GPU boxes expose many logical CPUs but deliver the throughput of far fewer (measured: the oracle peaks at
    16-32 threads and drops at 128: tools/cpu_scaling.py), so the CPU arm uses the thread count that is fastest on this
    host instead of blindly using os.cpu_count()."""
    n = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, n) if c <= n} | {n})
    best, best_rate = n, 0.0
    rows = max(1, min(height // 2, (1 << 19) // width))
    for c in cands:
        _, st, _ = oracle.render(scene.desc, width, height, task.pt, task.sampler, task.filter, pmj, bn, y0=0, y1=rows, spp_begin=0, spp_end=1, threads=c)
        rate = st.samples / max(st.seconds, 1e-9)
        if rate > best_rate:
            best, best_rate = c, rate
    return best


def run_reference(args, rank, world):
    """Reference arm.  The reference itself cannot be built here (Rust + un-vendored luisa_compute), so this
    times the CPU oracle — the restatement of its `-d cpu` path — with the fastest host thread count, on a bounded
    sample of the same workload: each step renders REF_SPP_PER_STEP samples per pixel of (a band of) the frame (cost is
    linear in spp and in pixels, pt.rs:1126-1149).  Rank 0 only; other ranks exit without work."""
    if rank != 0:
        return
    import akari_render_b200 as akr
    from oracle import binding as oracle
    width, height, spp = WORKLOADS[args.workload]
    spp = args.spp or spp
    scene = akr.load_scene(os.path.join(ROOT, "scenes", "cbox", "scene.json")).set_resolution(width, height)
    task = akr.RenderTask.from_file(os.path.join(ROOT, "scenes", "cbox", "pt.json"))
    task.pt.spp = spp
    pmj, bn = akr.sampler_tables()
    cores = best_cpu_threads(oracle, scene, task, pmj, bn, width, height)
    rows = min(height, max(1, (1280 * 720) // width))  # c4: a 225-row band in the middle of the frame = the pixel count of c2
    y0 = (height - rows) // 2
    vals = []
    for it in range(args.warmup + args.steps):
        s0 = (it * REF_SPP_PER_STEP) % spp
        _, st, _ = oracle.render(scene.desc, width, height, task.pt, task.sampler, task.filter, pmj, bn, y0=y0, y1=y0 + rows, spp_begin=s0,
                                 spp_end=s0 + REF_SPP_PER_STEP, threads=cores)
        if it >= args.warmup:
            vals.append((st.samples, st.seconds))  # the oracle's render-loop time == the window the reference times (pt.rs:1126-1157)
    samples = sum(v[0] for v in vals)
    secs = sum(v[1] for v in vals)
    value = samples / secs
    line = {
        "impl": "reference", "metric": "path samples/sec on cbox 1280x720" if args.workload == "c2" else f"path samples/sec on cbox {width}x{height}",
        "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, args.steps), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps x {REF_SPP_PER_STEP} of the {spp} spp of rows {y0}..{y0 + rows} of the {width}x{height} frame = {samples} samples, "
                                   f"{secs:.1f} s; CPU oracle (restatement of the reference's -d cpu path) on {cores} threads = the fastest "
                                   f"thread count on this host ({os.cpu_count()} logical CPUs)"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS), help="c2 = cbox 1280x720 @ 1024 spp (default, the BASELINE metric); c4 = cbox 4096^2 @ 4096 spp")
    ap.add_argument("--spp", type=int, default=0, help="override for quick experiments (the reported configs are 1024 / 4096)")
    ap.add_argument("--wave", type=int, default=1 << 26, help="paths in flight per wave (0 = engine default of 4 Mi)")
    ap.add_argument("--trace-mode", type=int, default=0, help="0 auto, 1 BVH, 2 flat list")
    ap.add_argument("--smem-node-kb", type=int, default=0, help="BVH scenes: KiB of top-of-tree nodes staged per CTA (0 = default)")
    ap.add_argument("--fused", type=int, default=0, help="0 auto (fused bounce kernels on flat scenes), 2 off (trace stage + shade stage + queues)")
    ap.add_argument("--block-rows", type=int, default=0, help="rows per interleaved block (0 = largest <= 8 that balances the ranks exactly)")
    ap.add_argument("--profile-stages", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--scene", default="cbox", choices=["cbox", "clutter"],
                    help="cbox = the BASELINE workload; clutter = cbox + 8.5 K-triangle spheres (BVH / divergence stress, not the headline)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import akari_render_b200 as akr
    from akari_render_b200.sharding import gather_index, gather_rows, interleaved_tile, max_tile_rows, pick_block_rows, tile_row_indices

    WIDTH, HEIGHT, spp = WORKLOADS[args.workload]
    spp = args.spp or spp
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    scene_path = os.path.join(ROOT, "scenes", "cbox", "scene.json")
    if args.scene == "clutter":  # generated on the fly by the test-suite's scene generator (experiments only)
        import tempfile
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import scene_variants
        scene_path = scene_variants.write_clutter(tempfile.mkdtemp(prefix=f"akr_clutter_{rank}_"))
    scene = akr.load_scene(scene_path).set_resolution(WIDTH, HEIGHT)
    task = akr.RenderTask.from_file(os.path.join(ROOT, "scenes", "cbox", "pt.json"))
    task.pt.spp = spp
    # image-plane shard: interleaved row blocks; per-GPU work shrinks with N => strong scaling of one frame
    # (the BASELINE metric is quoted on the fixed frame, so the frame is split and `scaling` is "strong").
    block_rows = args.block_rows or pick_block_rows(HEIGHT, world)
    tile = interleaved_tile(HEIGHT, world, rank, block_rows) if world > 1 else None
    my_rows = len(tile_row_indices(HEIGHT, world, rank, block_rows))
    max_rows = max_tile_rows(HEIGHT, world, block_rows)
    stream = torch.cuda.current_stream().cuda_stream
    pt = akr.PathTracer(local_rank, stream=stream)
    eng = dict(wave_size=args.wave, trace_mode=args.trace_mode, fused=args.fused, smem_node_kb=args.smem_node_kb)
    pt.set_engine_options(profile_stages=1 if args.profile_stages else 0, **eng)
    pt.upload_scene(scene)
    img_local = torch.zeros((max_rows, WIDTH, 3), device="cuda", dtype=torch.float32)
    gathered = torch.zeros((world, max_rows, WIDTH, 3), device="cuda", dtype=torch.float32) if world > 1 else None
    g_index = gather_index(HEIGHT, world, block_rows, device="cuda") if world > 1 else None
    host_img = torch.zeros((HEIGHT, WIDTH, 3), dtype=torch.float32).pin_memory()
    frame = [None]

    def render_all(n_spp):
        pt.begin(task, tile)
        done = 0
        while done < n_spp:
            cur = min(task.pt.spp_per_pass, n_spp - done)
            pt.render_pass(cur, blocking=False)
            done += cur

    def resolve_and_gather():
        pt.resolve_into_device(img_local.data_ptr(), my_rows * WIDTH * 3)
        if world > 1:  # the one collective: NCCL all_gather of the HDR rows, then the row permutation back to sensor order
            frame[0] = gather_rows(img_local, HEIGHT, WIDTH, rank, world, dist, block_rows=block_rows, out=gathered, index=g_index)
        else:
            frame[0] = img_local

    def step_device():
        """hot path with inputs resident: begin + passes + resolve on device (+ NCCL gather for N > 1)."""
        render_all(spp)
        resolve_and_gather()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, iters):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step_device()
    pt.reset_stats()
    clocks, stop = [], threading.Event()
    th = threading.Thread(target=sample_clocks, args=(stop, clocks), daemon=True)
    th.start()
    ms_total = timed(step_device, args.steps)
    stop.set()
    th.join(timeout=2)
    st = pt.stats()
    total_samples = WIDTH * HEIGHT * spp * args.steps
    value = total_samples / (ms_total * 1e-3)

    # ---- e2e: host buffers in, host frame out ----
    pmj, bn = akr.sampler_tables()
    # host buffers of the end-to-end leg live in PINNED memory (true async DMA): the sampler tables going in, the frame coming out
    pmj_pin = torch.from_numpy(pmj.view(np.int32)).pin_memory()
    bn_pin = torch.from_numpy(bn.view(np.int16)).pin_memory()
    pmj_h, bn_h = pmj_pin.numpy().view(np.uint32), bn_pin.numpy().view(np.uint16)

    def step_e2e():
        pt.upload_sampler_tables(pmj_h, bn_h)
        pt.upload_scene(scene)
        render_all(spp)
        resolve_and_gather()
        if rank == 0:  # the assembled HDR frame goes back to the host (what `util::write_image` would be handed, lib.rs:191-192)
            host_img.copy_(frame[0][:HEIGHT], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    e2e_value = total_samples / (ms_e2e * 1e-3)
    d = scene.desc.contents
    scene_bytes = int(d.shader_data_size)
    for i in range(d.n_meshes):
        m = d.meshes[i]
        scene_bytes += m.n_vertices * 12 + m.n_triangles * (12 + 24) + m.n_material_slots * 4
    h2d = pmj.nbytes + bn.nbytes + scene_bytes   # per rank (scene and tables are replicated)
    d2h = host_img.numel() * 4                   # rank 0: the whole frame

    # ---- N > 1: the gathered frame must equal the single-GPU frame bit for bit (one 64-spp pass, outside the timed region) ----
    gather_check = None
    check_spp = min(spp, task.pt.spp_per_pass)
    if world > 1:
        render_all(check_spp)
        resolve_and_gather()
        torch.cuda.synchronize()
        multi = frame[0][:HEIGHT].clone()
        dist.barrier()
        if rank == 0:
            pt.begin(task, None)
            pt.render_pass(check_spp, blocking=True)
            single = torch.zeros((HEIGHT, WIDTH, 3), device="cuda", dtype=torch.float32)
            pt.resolve_into_device(single.data_ptr(), single.numel())
            torch.cuda.synchronize()
            same = bool(torch.equal(single, multi))
            gather_check = {"spp": check_spp, "bit_identical_to_single_gpu_frame": same}
            del single
        dist.barrier()
        del multi

    # ---- roofline of the dominant kernel, measured live with per-stage CUDA events ----
    pt.set_engine_options(profile_stages=1, **eng)
    pt.reset_stats()
    pt.begin(task, tile)
    prof_spp = min(spp, task.pt.spp_per_pass)
    pt.render_pass(prof_spp, blocking=True)
    ps = pt.stats()
    pt.set_engine_options(profile_stages=0, **eng)
    names = ["raygen", "trace", "shade_lambert", "shade_conductor", "accumulate", "misc", "shade_general"]
    stage_ms = {names[i]: ps.gpu_ms_kernel[i] for i in range(7)}
    stage_launches = {names[i]: int(ps.launches_kernel[i]) for i in range(7)}
    shade_ms = stage_ms["shade_lambert"] + stage_ms["shade_conductor"] + stage_ms["shade_general"]
    shade_n = stage_launches["shade_lambert"] + stage_launches["shade_conductor"] + stage_launches["shade_general"]
    fused = stage_launches["trace"] == 0
    # algorithmic bytes of the profiled pass, per stage (every launch of the stage together), SURVEY 8(d) per-unit figures.
    # Fused pipeline: the bounce kernels do the shade AND the trace stage of a segment; there is no shadow queue, so no
    # shadow-queue bytes are charged (the camera rays' trace share belongs to the raygen kernel).
    if fused:
        trace_bytes = 0
        shade_bytes = ps.shaded_hits * B_SHADE_PER_SEG + (ps.segments - ps.samples) * B_TRACE_PER_SEG
        dom_name = "k_bounce<class> (shade + shadow ray + next ray)"
    else:
        trace_bytes = ps.segments * B_TRACE_PER_SEG + ps.shadow_rays * B_TRACE_PER_SHADOW
        shade_bytes = ps.shaded_hits * B_SHADE_PER_SEG + ps.shadow_rays * B_SHADE_PER_SHADOW
        dom_name = "k_shade<class>"
    if shade_ms >= stage_ms["trace"]:
        dom, dom_ms, dom_bytes, dom_n = dom_name, shade_ms, shade_bytes, shade_n
    else:
        dom, dom_ms, dom_bytes, dom_n = "k_trace", stage_ms["trace"], trace_bytes, stage_launches["trace"]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # measured DRAM traffic of the dominant stage, per launch: dram__bytes_read.sum + dram__bytes_write.sum from the committed
    # ncu launch list (profiles/launches_current.json <- tools/ncu_summary.py list).  Only quoted for the exact configuration
    # it was captured on (c2, one GPU, default engine options); null otherwise.
    traffic, traffic_src = None, None
    if world == 1 and args.workload == "c2" and args.scene == "cbox" and args.wave == (1 << 26) and args.fused == 0 and args.trace_mode == 0:
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "launches_current.json")))
            pref = "k_bounce" if fused else ("k_shade" if dom.startswith("k_shade") else "k_trace")
            ks = [v for k, v in tj.items() if k.startswith(pref)]
            if ks:
                traffic = sum(v["dram_read"] + v["dram_write"] for v in ks) / sum(v["launches"] for v in ks)
                traffic_src = "profiles/launches_current.json (ncu launch list of `bench.py --steps 1 --warmup 1 --spp 64`: five 64-spp passes, cold-cache, per launch)"
        except Exception:
            pass
    achieved = (dom_bytes / 1e9) / (dom_ms * 1e-3) if dom_ms > 0 else 0.0
    n_seg = st.segments / max(1, st.samples)
    s_ratio = st.shadow_rays / max(1, st.segments)
    hit_ratio = st.shaded_hits / max(1, st.segments)
    if fused:
        bytes_per_sample = B_SAMPLE + n_seg * (B_TRACE_PER_SEG + B_SHADE_PER_SEG * hit_ratio)
    else:
        bytes_per_sample = B_SAMPLE + n_seg * (B_TRACE_PER_SEG + B_SHADE_PER_SEG * hit_ratio + (B_TRACE_PER_SHADOW + B_SHADE_PER_SHADOW) * s_ratio)
    survey_model_bytes_per_sample = B_SAMPLE + n_seg * (300 + 152 * s_ratio)  # SURVEY 8(d) verbatim (charges a shadow queue)
    moved_bytes_per_sample = B_FUSED_PER_SAMPLE + n_seg * hit_ratio * B_FUSED_PER_HIT if fused else None
    my_rate = (WIDTH * my_rows * spp * args.steps) / (ms_total * 1e-3)  # this rank's samples per second
    pipeline_gbs = my_rate * bytes_per_sample / 1e9

    # per-rank path statistics (load balance of the split)
    per_rank = None
    if world > 1:
        v = torch.tensor([n_seg, float(my_rows)], device="cuda", dtype=torch.float64)
        allv = [torch.zeros_like(v) for _ in range(world)]
        dist.all_gather(allv, v)
        per_rank = {"n_seg": [round(float(t[0]), 4) for t in allv], "rows": [int(t[1]) for t in allv]}

    # ---- CPU baseline on rank 0, N = 1 only, bounded sample ----
    # Timed window = the oracle's own render loop (AkrOracleStats.seconds), the window the reference times itself
    # (Instant around each dispatch, pt.rs:1126-1157): scene preparation and the Python binding are excluded.
    cpu = None
    if rank == 0 and world == 1 and args.scene == "cbox" and args.cpu_seconds > 0:  # (the oracle brute-forces every triangle: only the headline scene is timed)
        from oracle import binding as oracle
        cores = best_cpu_threads(oracle, scene, task, pmj, bn, WIDTH, HEIGHT)
        rows = min(HEIGHT, max(1, (1280 * 720) // WIDTH))
        y0 = (HEIGHT - rows) // 2
        _, ost, _ = oracle.render(scene.desc, WIDTH, HEIGHT, task.pt, task.sampler, task.filter, pmj, bn, y0=y0, y1=y0 + rows, spp_begin=0, spp_end=1, threads=cores)
        rate = ost.samples / max(ost.seconds, 1e-6)
        n_spp = int(min(64, max(1, rate * args.cpu_seconds / (WIDTH * rows))))
        _, ost, _ = oracle.render(scene.desc, WIDTH, HEIGHT, task.pt, task.sampler, task.filter, pmj, bn, y0=y0, y1=y0 + rows, spp_begin=1, spp_end=1 + n_spp,
                                  threads=cores)
        cpu = {"value": ost.samples / ost.seconds, "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": f"{n_spp} of the {spp} spp of rows {y0}..{y0 + rows} of the {WIDTH}x{HEIGHT} frame ({ost.samples} samples, {ost.seconds:.1f} s), CPU oracle on "
                         f"{cores} threads (fastest thread count on this host, {os.cpu_count()} logical CPUs)"}

    if rank == 0:
        line = {
            "metric": ("path samples/sec on cbox 1280x720" if args.workload == "c2" else f"path samples/sec on cbox {WIDTH}x{HEIGHT}") +
                      ("" if args.scene == "cbox" else " + clutter"),
            "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world, block_rows if world > 1 else None),
            "clocks": summarize_clocks(clocks),
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "includes": "table + scene upload, render, resolve" + (", NCCL gather" if world > 1 else "") + ", full-frame D2H on rank 0"},
            "gpu_launches": int(st.kernel_launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": dom_bytes / max(1, dom_n),
                         "kernel": dom, "kernel_launches": dom_n, "kernel_ms": dom_ms, "kernel_algorithmic_bytes": dom_bytes,
                         "peak_source": "measured" if peaks else "fallback",
                         "bytes_model": "SURVEY 8(d) per-unit figures; no shadow-queue bytes where no shadow queue exists (fused pipeline)",
                         "pipeline_algorithmic_gbs": pipeline_gbs, "pipeline_frac": pipeline_gbs / peak, "bytes_per_sample": bytes_per_sample,
                         "survey_formula_bytes_per_sample": survey_model_bytes_per_sample,
                         "survey_formula_pipeline_frac": my_rate * survey_model_bytes_per_sample / 1e9 / peak,
                         "moved_by_design_bytes_per_sample": moved_bytes_per_sample,
                         "moved_by_design_pipeline_frac": (my_rate * moved_bytes_per_sample / 1e9 / peak) if moved_bytes_per_sample else None,
                         "n_seg": n_seg, "shadow_per_seg": s_ratio, "hit_per_seg": hit_ratio,
                         "stage_ms": stage_ms, "stage_launches": stage_launches, "profiled_spp": prof_spp, "rank": 0},
            "cpu_baseline": cpu,
        }
        if per_rank:
            line["per_rank"] = per_rank
        if gather_check:
            line["gather_check"] = gather_check
        print(json.dumps(line), flush=True)
    pt.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
