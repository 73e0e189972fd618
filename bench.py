#!/usr/bin/env python3
"""bench.py — path samples/s of the wavefront path tracer on cbox 1280x720 @ 1024 spp (BASELINE config C2).

Contract: `python bench.py --gpus N --steps K --warmup W` (torchrun for N > 1) prints ONE JSON line on rank 0.
  step      = one full render of the workload (all 1024 spp = 16 passes of 64 spp, pt.rs:1126-1149)
  value     = path samples / s, whole job, scene + sampler tables already resident in HBM, film left on device
  e2e       = same metric through the public call with HOST buffers: scene upload (H2D, incl. host BVH build),
              sampler-table upload (H2D), render, film download (D2H) — all inside the timed region
  roofline  = HBM roofline of the dominant kernel (k_shade): algorithmic bytes per launch / CUDA-event time
  cpu_baseline = the CPU oracle (restatement of the reference's `-d cpu` path) on a bounded sample
`--impl reference` times the CPU oracle instead (the reference binary cannot be built here: DESIGN.md).
Multi-GPU: the image plane is split into contiguous row bands, one rank per GPU, scene replicated, no
collective on the hot path; one NCCL all_gather assembles the HDR image at the end of every step.
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

WIDTH, HEIGHT, SPP = 1280, 720, 1024
# packed record sizes -> algorithmic bytes
B_SAMPLE, B_SEG, B_SHADOW = 168, 300, 152
# split of those per-unit figures over the two stages of one bounce (DESIGN.md "algorithmic bytes"):
B_TRACE_PER_SEG, B_SHADE_PER_SEG = 72, 228   # trace: R RAY + W HIT + idx; shade: R HIT + wo + PATH, W PATH + RAY + key/idx
B_TRACE_PER_SHADOW, B_SHADE_PER_SHADOW = 88, 64  # trace: R RAY+NEE+idx (64) + RMW L (24); shade: W RAY+NEE+idx (64)


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


def workload_config(spp, world, wave, trace_mode):
    return {"workload": f"cbox 1280x720 @ {spp}spp, pmj02bn seed 0, gaussian r=1.5, max_depth 12, rr_depth 5, 64 spp/pass",
            "parallelism": f"image rows x{world}", "l2": "working set (wave state) > L2; no inter-step reuse (film cleared each step)",
            "wave_paths": wave or (1 << 22), "trace_mode": trace_mode}


REF_SPP_PER_STEP = 4


def best_cpu_threads(oracle, scene, task, pmj, bn):
    """The GPU boxes expose 128 logical CPUs but deliver the throughput of far fewer (measured: the oracle peaks at
    16-32 threads, 2.5 M samples/s, and drops to 1.5 M at 128: tools/cpu_scaling.py), so the CPU arm uses the thread
    count that is fastest on this host instead of blindly using os.cpu_count()."""
    n = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, n) if c <= n} | {n})
    best, best_rate = n, 0.0
    for c in cands:
        _, st, _ = oracle.render(scene.desc, WIDTH, HEIGHT, task.pt, task.sampler, task.filter, pmj, bn, y0=0, y1=HEIGHT // 2, spp_begin=0, spp_end=1,
                                 threads=c)
        rate = st.samples / max(st.seconds, 1e-9)
        if rate > best_rate:
            best, best_rate = c, rate
    return best


def run_reference(args, rank, world):
    """Reference arm.  The reference itself cannot be built here (Rust + un-vendored luisa_compute), so
    this times the CPU oracle — the restatement of its `-d cpu` path — with every host thread, on a bounded sample of
    the same workload: each step renders REF_SPP_PER_STEP of the 1024 samples per pixel of the full 1280x720 frame
    (cost is linear in spp, pt.rs:1126-1149).  Rank 0 only; other ranks exit without work."""
    if rank != 0:
        return
    import akari_render_b200 as akr
    from oracle import binding as oracle
    scene = akr.load_scene(os.path.join(ROOT, "scenes", "cbox", "scene.json")).set_resolution(WIDTH, HEIGHT)
    task = akr.RenderTask.from_file(os.path.join(ROOT, "scenes", "cbox", "pt.json"))
    task.pt.spp = SPP
    pmj, bn = akr.sampler_tables()
    cores = best_cpu_threads(oracle, scene, task, pmj, bn)
    vals = []
    for it in range(args.warmup + args.steps):
        s0 = (it * REF_SPP_PER_STEP) % SPP
        _, st, _ = oracle.render(scene.desc, WIDTH, HEIGHT, task.pt, task.sampler, task.filter, pmj, bn, spp_begin=s0, spp_end=s0 + REF_SPP_PER_STEP,
                                 threads=cores)
        if it >= args.warmup:
            vals.append((st.samples, st.seconds))  # the oracle's render-loop time == the window the reference times (pt.rs:1126-1157)
    samples = sum(v[0] for v in vals)
    secs = sum(v[1] for v in vals)
    value = samples / secs
    line = {
        "impl": "reference", "metric": "path samples/sec on cbox 1280x720", "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, args.steps), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(SPP, args.gpus, args.wave, args.trace_mode),
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps x {REF_SPP_PER_STEP} of the 1024 spp of the full 1280x720 frame = {samples} samples, "
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
    ap.add_argument("--spp", type=int, default=SPP, help="override for quick experiments (the reported config is 1024)")
    ap.add_argument("--wave", type=int, default=1 << 26, help="paths in flight per wave (0 = engine default of 4 Mi)")
    ap.add_argument("--trace-mode", type=int, default=0, help="0 auto, 1 BVH, 2 flat list")
    ap.add_argument("--smem-node-kb", type=int, default=0, help="BVH scenes: KiB of top-of-tree nodes staged per CTA (0 = default)")
    ap.add_argument("--fused", type=int, default=0, help="0 auto (fused bounce kernels on flat scenes), 2 off (trace stage + shade stage + queues)")
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
    task.pt.spp = args.spp
    spp = args.spp
    # image-plane shard: contiguous row bands; per-GPU work shrinks with N => strong scaling of one frame.
    # Weak scaling as the contract defines it (fixed per-GPU work) = every rank renders the full 1280x720 band count / world? No:
    # the BASELINE metric is quoted on the fixed 1280x720 frame, so the frame is split and `scaling` is "strong".
    from akari_render_b200.sharding import gather_bands, max_band_rows, row_bands
    tile = row_bands(HEIGHT, world)[rank]
    my_rows = tile[1] - tile[0]
    stream = torch.cuda.current_stream().cuda_stream
    pt = akr.PathTracer(local_rank, stream=stream)
    eng = dict(wave_size=args.wave, trace_mode=args.trace_mode, fused=args.fused, smem_node_kb=args.smem_node_kb)
    pt.set_engine_options(profile_stages=1 if args.profile_stages else 0, **eng)
    pt.upload_scene(scene)
    max_rows = max_band_rows(HEIGHT, world)
    img_local = torch.zeros((max_rows, WIDTH, 3), device="cuda", dtype=torch.float32)
    gathered = torch.zeros((world, max_rows, WIDTH, 3), device="cuda", dtype=torch.float32) if world > 1 else None
    host_img = torch.zeros((HEIGHT, WIDTH, 3), dtype=torch.float32).pin_memory()

    def step_device():
        """hot path with inputs resident: begin + 16 passes + resolve on device (+ NCCL gather for N > 1)."""
        pt.begin(task, tile)
        done = 0
        while done < spp:
            cur = min(task.pt.spp_per_pass, spp - done)
            pt.render_pass(cur, blocking=False)
            done += cur
        pt.resolve_into_device(img_local.data_ptr(), my_rows * WIDTH * 3)
        if world > 1:
            gather_bands(img_local, HEIGHT, WIDTH, rank, world, dist, out=gathered)  # the one collective: NCCL all_gather of the HDR bands

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

    # ---- e2e: host buffers in, host film out ----
    pmj, bn = akr.sampler_tables()
    # host buffers of the end-to-end leg live in PINNED memory (true async DMA): the sampler tables going in, the film coming out
    pmj_pin = torch.from_numpy(pmj.view(np.int32)).pin_memory()
    bn_pin = torch.from_numpy(bn.view(np.int16)).pin_memory()
    pmj_h, bn_h = pmj_pin.numpy().view(np.uint32), bn_pin.numpy().view(np.uint16)
    film_pin = torch.empty(7 * WIDTH * my_rows, dtype=torch.float32).pin_memory()
    film_host = film_pin.numpy()

    def step_e2e():
        pt.upload_sampler_tables(pmj_h, bn_h)
        pt.upload_scene(scene)
        pt.begin(task, tile)
        done = 0
        while done < spp:
            cur = min(task.pt.spp_per_pass, spp - done)
            pt.render_pass(cur, blocking=False)
            done += cur
        pt._check(pt._lib.akr_b200_download_film(pt._ctx, film_host.ctypes.data, film_host.size))

    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    e2e_value = total_samples / (ms_e2e * 1e-3)
    d = scene.desc.contents
    scene_bytes = int(d.shader_data_size)
    for i in range(d.n_meshes):
        m = d.meshes[i]
        scene_bytes += m.n_vertices * 12 + m.n_triangles * (12 + 24) + m.n_material_slots * 4
    h2d = pmj.nbytes + bn.nbytes + scene_bytes
    d2h = film_host.nbytes

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
    # algorithmic bytes of the profiled pass, per stage (every launch of the stage together)
    trace_bytes = ps.segments * B_TRACE_PER_SEG + ps.shadow_rays * B_TRACE_PER_SHADOW
    shade_bytes = ps.segments * B_SHADE_PER_SEG + ps.shadow_rays * B_SHADE_PER_SHADOW
    if shade_ms >= stage_ms["trace"]:
        dom, dom_ms, dom_bytes, dom_n = "k_shade<class>", shade_ms, shade_bytes, shade_n
    else:
        dom, dom_ms, dom_bytes, dom_n = "k_trace", stage_ms["trace"], trace_bytes, stage_launches["trace"]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # measured DRAM traffic of the dominant stage, per launch: dram__bytes_read.sum + dram__bytes_write.sum from the committed
    # ncu launch list of this same command at 64 spp (profiles/launches_current.json <- tools/ncu_summary.py list)
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "launches_current.json")))
        pref = "k_shade" if dom.startswith("k_shade") else "k_trace"
        ks = [v for k, v in tj.items() if k.startswith(pref)]
        if ks:
            traffic = sum(v["dram_read"] + v["dram_write"] for v in ks) / sum(v["launches"] for v in ks)
            traffic_src = "profiles/launches_current.json (ncu, cold-cache, per launch)"
    except Exception:
        pass
    achieved = (dom_bytes / 1e9) / (dom_ms * 1e-3) if dom_ms > 0 else 0.0
    n_seg = st.segments / max(1, st.samples)
    s_ratio = st.shadow_rays / max(1, st.segments)
    bytes_per_sample = B_SAMPLE + n_seg * (B_SEG + B_SHADOW * s_ratio)
    pipeline_gbs = value / world * bytes_per_sample / 1e9

    # ---- CPU baseline on rank 0, N = 1 only, bounded sample ----
    # Timed window = the oracle's own render loop (AkrOracleStats.seconds), the window the reference times itself
    # (Instant around each dispatch, pt.rs:1126-1157): scene preparation and the Python binding are excluded.
    cpu = None
    if rank == 0 and world == 1 and args.scene == "cbox":  # (the oracle brute-forces every triangle: only the headline scene is timed)
        from oracle import binding as oracle
        cores = best_cpu_threads(oracle, scene, task, pmj, bn)
        _, ost, _ = oracle.render(scene.desc, WIDTH, HEIGHT, task.pt, task.sampler, task.filter, pmj, bn, spp_begin=0, spp_end=1, threads=cores)
        rate = ost.samples / max(ost.seconds, 1e-6)
        n_spp = int(min(64, max(1, rate * args.cpu_seconds / (WIDTH * HEIGHT))))
        _, ost, _ = oracle.render(scene.desc, WIDTH, HEIGHT, task.pt, task.sampler, task.filter, pmj, bn, spp_begin=1, spp_end=1 + n_spp, threads=cores)
        cpu = {"value": ost.samples / ost.seconds, "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": f"{n_spp} of the 1024 spp of the full 1280x720 frame ({ost.samples} samples, {ost.seconds:.1f} s), CPU oracle on {cores} threads "
                         f"(fastest thread count on this host, {os.cpu_count()} logical CPUs)"}

    if rank == 0:
        line = {
            "metric": "path samples/sec on cbox 1280x720", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(spp, world, args.wave, args.trace_mode), **({"scene": args.scene} if args.scene != "cbox" else {})),
            "clocks": summarize_clocks(clocks),
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(st.kernel_launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": dom_bytes / max(1, dom_n),
                         "kernel": dom, "kernel_launches": dom_n, "kernel_ms": dom_ms, "kernel_algorithmic_bytes": dom_bytes,
                         "peak_source": "measured" if peaks else "fallback",
                         "pipeline_algorithmic_gbs": pipeline_gbs, "pipeline_frac": pipeline_gbs / peak,
                         "n_seg": n_seg, "shadow_per_seg": s_ratio, "bytes_per_sample": bytes_per_sample,
                         "stage_ms": stage_ms, "stage_launches": stage_launches, "profiled_spp": prof_spp},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    pt.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
