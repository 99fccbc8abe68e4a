#!/usr/bin/env python
"""bench.py -- tiles/sec of the multi-scale + flip pseudo-mask fusion (BASELINE.json metric) on N B200s.

One "step" = one pass of the fused hot path (pisto_fuse_argmax_confusion) over one batch of synthetic tiles of
BASELINE config 2 (WSSS4LUAD pseudo-mask inference: 224x224 tiles, 3 classes, scales {0.75,1,1.25} x hflip = 6 stride-8
views, class-presence vector, background mask, uint8 label map + 32x32 logit export).

  python bench.py [--gpus N] [--steps K] [--warmup W]            ours, one process per GPU (torchrun for N > 1)
  python bench.py --impl reference ...                            the reference's own torch-CPU path (oracle/pipeline.py)

Prints ONE JSON line on stdout (contract in the task statement); diagnostics go to stderr.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "tiles/sec, multi-scale+flip pseudo-mask fusion"
UNIT = "tiles/s"
SCALES = (0.75, 1.0, 1.25)
T, C = 224, 3


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Native libraries write banners to fd 1 (NCCL prints its version there when
# NCCL_DEBUG is set), so fd 1 is pointed at stderr for the whole run and the JSON line goes to a private copy of the real stdout.
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def bytes_per_tile(sizes, with_gt=False):
    """Algorithmic (compulsory) HBM bytes per tile, SURVEY.md 8(d): views + bg + label (+ gt) + 32x32 logits."""
    return 4 * C * sum(h * h for h in sizes) + T * T * (2 + int(with_gt)) + 4 * C * 32 * 32


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons sampled WHILE the timed region runs: NVML polled from a thread every ~2 ms (a 20-step region is
    only ~20 ms long, nvidia-smi's own loop is too coarse for it); falls back to `nvidia-smi -lms` when NVML is unavailable."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.samples, self.rows, self.t = [], [], []
        self.proc, self.nvml, self.stop_flag = None, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        except Exception as e:
            log("NVML clock sampler unavailable, using nvidia-smi:", repr(e)[:120])
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception as e:  # pragma: no cover
            log("clock sampler unavailable:", e)

    def _poll(self):
        n = self.nvml
        bits = {"hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8), "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop_flag:
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
                r = int(get_reasons(self.h))
                self.samples.append((time.time(), sm, [k for k, b in bits.items() if r & b]))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip()); self.t.append(time.time())

    def stop(self, t0, t1):
        if self.nvml is not None:
            self.stop_flag = True
            self.th.join(timeout=1.0)
            sel = [x for x in self.samples if t0 <= x[0] <= t1] or self.samples[-3:]
            if not sel:
                return None
            sm = sorted(x[1] for x in sel)
            reasons = sorted({r for x in sel for r in x[2]})
            return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_sm, "reasons": reasons, "samples": len(sel), "source": "NVML, 2 ms poll inside the timed region"}
        if self.proc is None:
            return None
        time.sleep(0.12)
        self.proc.terminate()
        sel = [r for r, t in zip(self.rows, self.t) if t0 <= t <= t1 + 0.06] or self.rows[-3:]
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in sel:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except Exception:
                continue
            for nme, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 20"}


def cpu_reference_run(n_tiles, repeats=1, threads=None):
    """The reference's torch-CPU path on a bounded sample of the same workload, all host threads (or `threads`)."""
    from oracle import pipeline
    from pistoseg_b200 import synthetic
    torch.set_num_threads(threads or os.cpu_count())
    cfg = synthetic.cfg2(N=n_tiles, T=T, C=C, scales=SCALES)
    pres, bg = cfg["present"].numpy(), cfg["bg"].numpy()
    pipeline.pseudo_mask_batch([v[:8] for v in cfg["views"]], cfg["codes"], (T, T), pres[:8], bg[:8])  # warm
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        pipeline.pseudo_mask_batch(cfg["views"], cfg["codes"], (T, T), pres, bg)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n_tiles / best, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample: calibrate on 64 tiles, then size each step so that W + K steps take about args.cpu_budget seconds
    rate, _ = cpu_reference_run(64)
    n = int(max(32, min(args.cpu_tiles, args.cpu_budget * rate / (args.steps + args.warmup))))
    times = []
    for i in range(args.warmup + args.steps):
        v, dt = cpu_reference_run(n)
        if i >= args.warmup:
            times.append(dt)
    dt = sum(times) / len(times)
    val = n / dt
    sizes = [int(T * s) // 8 for s in SCALES for _ in (0, 1)]
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2: WSSS4LUAD pseudo-mask inference, 224x224 tiles, C=3, scales {0.75,1,1.25} x hflip (V=6 stride-8 views 21/28/35), present vector (40% single-label), bg mask, u8 labels + 32x32 logits",
                   "tiles_per_step": n, "views": sizes},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                         "sample": f"{n} tiles per step x {args.steps} steps; oracle/pipeline.py (literal torch-CPU restatement of infer_pseudo_masks.py:118-154, torch.set_num_threads({os.cpu_count()}))"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def pin_rank_to_cores(local, world):
    """One process per GPU: give every rank its own slice of the host cores BEFORE any pinned allocation, so that staging buffers
    are first touched (and their copy threads run) on cores this rank owns.  Returns the slice for the record."""
    try:
        avail = sorted(os.sched_getaffinity(0))
        per = max(1, len(avail) // max(world, 1))
        mine = avail[local * per:(local + 1) * per] or avail
        os.sched_setaffinity(0, mine)
        return {"cores": [mine[0], mine[-1]], "n": len(mine), "of": len(avail)}
    except Exception as e:  # pragma: no cover
        return {"error": str(e)}


def event_time(fn, steps, dist_mod, world, dev):
    """CUDA-event time of `steps` calls of fn, bracketed by barrier + synchronize, max over ranks (ms)."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist_mod.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist_mod.barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist_mod.all_reduce(t, op=dist_mod.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def run_cfg3(args, dist_mod, world, rank, dev):
    """BASELINE config 3 (strong scaling): 10 000 BCSS-shaped tiles (C = 4, 3 scales x flip, gt in {0..4}) in contiguous shards,
    per step: zero the matrix, fused kernel on the shard, ONE all-reduce (NCCL) of the int64 [4,4] confusion matrix -- all inside
    the CUDA-event region.  Rank 0 checks the merged matrix against the unsharded run (untimed)."""
    from pistoseg_b200 import dist as pdist, ops, synthetic
    from pistoseg_b200._lib import DECIDE_SOFTMAX
    N3 = 10000
    base = synthetic.cfg3(N=1000)  # seeded block, identical on every rank
    lo, hi = pdist.shard_range(N3, rank, world)

    def up(t, a, b):  # tiles [a, b) of the 10x repeated block
        idx = torch.arange(a, b) % t.shape[0]
        return t[idx].to(dev).contiguous()
    views, gt = [up(v, lo, hi) for v in base["views"]], up(base["gt"], lo, hi)
    conf = ops.new_confusion(4, dev)

    def kernel_only():
        ops.fuse_argmax_confusion(views, base["codes"], (224, 224), decide=DECIDE_SOFTMAX, gt=gt, conf=conf)

    def step():
        conf.zero_()
        kernel_only()
        pdist.all_reduce_confusion(conf)  # the one collective of the path

    steps = max(20, args.steps // 4)
    for _ in range(3):
        step()
    ms = event_time(step, steps, dist_mod, world, dev)
    ms_kernel = event_time(kernel_only, steps, dist_mod, world, dev)
    scratch = torch.zeros(16, dtype=torch.int64, device=dev)
    ms_ar = event_time(lambda: pdist.all_reduce_confusion(scratch), steps, dist_mod, world, dev) if world > 1 else 0.0
    # the same step replayed from a CUDA graph (memset + kernel + NCCL all-reduce captured once): what the launch overhead costs
    ms_graph = None
    try:
        gph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()
            with torch.cuda.graph(gph, stream=side):
                step()
        torch.cuda.current_stream().wait_stream(side)
        gph.replay()
        ms_graph = event_time(gph.replay, steps, dist_mod, world, dev)
    except Exception as e:  # pragma: no cover
        log("cfg3: CUDA-graph replay unavailable:", repr(e)[:200])
    step()
    torch.cuda.synchronize()
    merged = conf.clone()
    exact = None
    if rank == 0:
        fv, fgt = [up(v, 0, N3) for v in base["views"]], up(base["gt"], 0, N3)
        full = ops.new_confusion(4, dev)
        ops.fuse_argmax_confusion(fv, base["codes"], (224, 224), decide=DECIDE_SOFTMAX, gt=fgt, conf=full)
        torch.cuda.synchronize()
        exact = bool(torch.equal(full, merged))
        assert exact, "cfg3: merged confusion matrix differs from the single-GPU matrix"
    bpt = 4 * 4 * sum(v.shape[2] * v.shape[3] for v in views) + 2 * 224 * 224
    peak, _ = measured_peak()
    rate = N3 * steps / (ms * 1e-3)
    terms = {"kernel_ms_per_step": ms_kernel / steps, "allreduce_us": ms_ar / steps * 1e3, "step_ms": ms / steps}
    lim = "kernel (shard too small to matter)" if ms_kernel > 0.8 * ms else "launch + NCCL all-reduce latency against a sub-millisecond kernel"
    return {"workload": "cfg3: 10 000 BCSS-shaped tiles, C=4, V=6 (21/28/35 x hflip), gt in {0..4}, labels + confusion, contiguous shards, "
                        "confusion all-reduce(SUM, int64[16]) every step inside the timed region",
            "scaling": "strong", "n_tiles": N3, "n_gpus": world, "shard_rank0": [lo, hi], "steps": steps,
            "tiles_per_s": rate, "tiles_per_s_kernel_only": N3 * steps / (ms_kernel * 1e-3),
            "tiles_per_s_cuda_graph": (N3 * steps / (ms_graph * 1e-3)) if ms_graph else None,
            "hbm_frac_per_gpu": rate / world * bpt / 1e9 / peak, **terms, "merged_exact": exact, "limiting_term": lim,
            "note": "the reference reports every 100 batches (segmentation_test.py:130): one all-reduce per report point, not per step, "
                    "would make the collective vanish; per step is the worst case and is what is timed here"}


def run_cfg5(args, dist_mod, world, rank, dev, T5=1024):
    """BASELINE config 5 at T = 1024 (weak scaling): 4 classes, 5 scales x flip (V = 10), gt + labels + confusion, one all-reduce per step."""
    from pistoseg_b200 import dist as pdist, ops, synthetic
    from pistoseg_b200._lib import DECIDE_SOFTMAX
    n5 = 256
    base = synthetic.cfg5(N=8, T=T5)
    rep = n5 // 8
    views = [v.to(dev).repeat((rep, 1, 1, 1)).contiguous() for v in base["views"]]
    gt = base["gt"].to(dev).repeat((rep, 1, 1)).contiguous()
    conf = ops.new_confusion(4, dev)

    def step():
        conf.zero_()
        ops.fuse_argmax_confusion(views, base["codes"], (T5, T5), decide=DECIDE_SOFTMAX, gt=gt, conf=conf)
        pdist.all_reduce_confusion(conf)
    steps = 10
    for _ in range(2):
        step()
    ms = event_time(step, steps, dist_mod, world, dev)
    bpt = 4 * 4 * sum(v.shape[2] * v.shape[3] for v in views) + 2 * T5 * T5
    peak, _ = measured_peak()
    rate = world * n5 * steps / (ms * 1e-3)
    return {"workload": f"cfg5: T={T5}, C=4, scales [1,1.25,1.5,1.75,2] x hflip (V=10), gt + labels + confusion, {n5} tiles per GPU, all-reduce every step",
            "scaling": "weak", "n_gpus": world, "tiles_per_gpu": n5, "steps": steps, "tiles_per_s": rate, "ms_per_step": ms / steps,
            "bytes_per_tile": bpt, "hbm_frac_per_gpu": rate / world * bpt / 1e9 / peak}


def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    affinity = pin_rank_to_cores(local, world)
    from pistoseg_b200 import _lib, ops, synthetic
    from pistoseg_b200._lib import DECIDE_SOFTMAX, MASK_FILL

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; pistoseg_b200 has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    N = args.tiles
    sizes = synthetic.view_sizes(T, SCALES)
    # synthetic inputs: a 1024-tile seeded block generated on the CPU (identical bits to what the oracle sees),
    # tiled up to N on the device so that one step streams 2.8 GB (>> 126 MB L2) through the kernel
    base = synthetic.cfg2(N=min(N, 1024), T=T, C=C, scales=SCALES, single_frac=args.single_frac)
    rep = (N + base["views"][0].shape[0] - 1) // base["views"][0].shape[0]

    def up(t):
        return t.to(dev).repeat((rep,) + (1,) * (t.dim() - 1))[:N].contiguous()
    views = [up(v) for v in base["views"]]
    bg, present = up(base["bg"]), up(base["present"])
    codes = base["codes"]

    def step():
        return ops.fuse_argmax_confusion(views, codes, (T, T), mask_mode=MASK_FILL, decide=DECIDE_SOFTMAX, present=present, bg=bg,
                                         bg_match=1, bg_label=C, lowres=(32, 32))
    out = step()
    torch.cuda.synchronize()
    if rank == 0 and not args.skip_check:
        # cheap in-run sanity check against the oracle on the first 64 tiles (not timed)
        from oracle import fuse as ofuse
        sub = {k: [v[:64] for v in base["views"]] if k == "views" else base[k] for k in ("views", "codes")}
        fused = ofuse.fuse_views(sub["views"], sub["codes"], (T, T))
        lab = ofuse.pseudo_masks(fused, base["present"][:64].numpy(), base["bg"][:64].numpy())
        agree = float((out["labels"][:64].cpu().numpy() == lab).mean())
        assert agree >= 0.9999, f"label agreement with the oracle {agree}"
        assert torch.equal(out["lowres"][:64].cpu(), ofuse.lowres_32(fused)), "32x32 logits differ from the oracle"
        log(f"parity check on 64 tiles: label agreement {agree:.6f}, 32x32 logits bit-exact")

    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        step()
    l0 = _lib.launch_count(local)
    t_wall0 = time.time()
    ms = event_time(step, args.steps, dist, world, dev)
    t_wall1 = time.time()
    launches = _lib.launch_count(local) - l0
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None

    # ---- end to end: the C ABI's *_host call, copies inside the timed region; timed twice: by the library's CUDA events (first H2D
    # start -> last D2H end) and by the caller's wall clock around the Python call (perf_counter, includes ctypes / driver overhead).
    # Variant (a): every input in pinned HOST memory (what a caller holding numpy data does).  Variant (b): the reference's own
    # dataflow (infer_pseudo_masks.py:119-137) -- the logits are already on the GPU, only `tissue` / the label vector come from the
    # host and the labels + 32x32 logits go back.
    Ne = args.e2e_tiles
    repe = (Ne + base["views"][0].shape[0] - 1) // base["views"][0].shape[0]

    def pin(t):
        return t.repeat((repe,) + (1,) * (t.dim() - 1))[:Ne].contiguous().pin_memory()
    hviews = [pin(v) for v in base["views"]]
    hbg, hpres = pin(base["bg"]), pin(base["present"])
    hout = {"labels": torch.empty((Ne, T, T), dtype=torch.uint8).pin_memory(), "lowres": torch.empty((Ne, C, 32, 32), dtype=torch.float32).pin_memory()}
    dviews = [v[:Ne] if v.shape[0] >= Ne else v.repeat((repe, 1, 1, 1))[:Ne].contiguous() for v in views]

    def e2e_run(vws):
        def estep():
            ops.fuse_argmax_confusion_host(vws, codes, (T, T), mask_mode=MASK_FILL, decide=DECIDE_SOFTMAX, present=hpres, bg=hbg, bg_match=1,
                                           bg_label=C, lowres=(32, 32), chunk=args.e2e_chunk, device=local, out=hout)
            return _lib.last_pipeline_ms(local)
        for _ in range(2):
            estep()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e_ms, w0 = 0.0, time.perf_counter()
        for _ in range(args.e2e_steps):
            e_ms += estep()
        w_ms = (time.perf_counter() - w0) * 1e3
        if world > 1:
            t = torch.tensor([e_ms, w_ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms, w_ms = float(t[0].item()), float(t[1].item())
        return e_ms, w_ms
    e_ms, ew_ms = e2e_run(hviews)
    assert torch.equal(hout["labels"][:256], out["labels"][:256].cpu()), "e2e labels differ from the device-resident run"
    d_ms, dw_ms = e2e_run(dviews)
    assert torch.equal(hout["labels"][:256], out["labels"][:256].cpu()), "e2e (device logits) labels differ from the device-resident run"
    h2d_views = sum(v[0].numel() * 4 for v in hviews) * Ne
    h2d_masks = hbg[0].numel() * Ne + C * Ne
    d2h = (T * T + C * 32 * 32 * 4) * Ne

    cfg3 = run_cfg3(args, dist, world, rank, dev) if not args.no_extra else None
    cfg5 = run_cfg5(args, dist, world, rank, dev) if not args.no_extra else None

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    value = world * N * args.steps / (ms * 1e-3)
    bpt = bytes_per_tile(sizes)
    peak, peak_src = measured_peak()
    per_launch_ms = ms / max(launches, 1)
    achieved = bpt * N / (per_launch_ms * 1e-3) / 1e9
    traffic, other, kernel_name = None, None, "fuse_static_kernel<C=3,G=3,VPG=2,F=25,NB=2>"
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        if tj.get("tiles_per_launch"):
            traffic = tj["dram_bytes_per_launch"] * (N / tj["tiles_per_launch"])
        # the stride-8 multi-view path is not HBM-bound; the fractions of the resources that do bind it come from the
        # committed ncu capture of this kernel (they are properties of the kernel, not re-measured under the timer)
        other = tj.get("other_resources")
        kernel_name = tj.get("kernel", kernel_name)
    except Exception:
        pass
    e_val = world * Ne * args.e2e_steps / (e_ms * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2: WSSS4LUAD pseudo-mask inference, 224x224 tiles, C=3, scales {0.75,1,1.25} x hflip (V=6 stride-8 views 21/28/35), present vector (40% single-label), bg mask, u8 labels + 32x32 logits",
                   "tiles_per_step_per_gpu": N, "views": sizes, "parallelism": f"tile-sharded x{world}, no data-path collective",
                   "l2": f"inputs+outputs {bpt * N / 1e9:.2f} GB per step > 126 MB L2 (no flush needed)"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "bytes_per_tile": bpt, "kernel": kernel_name,
                     "other_resources": other,
                     "note": "achieved = algorithmic bytes/tile (SURVEY.md 8(d)) x tiles per launch / CUDA-event time per launch; traffic = ncu "
                             "dram read+write bytes of one launch (profiles/traffic.json) scaled to this launch size; the kernel is "
                             "issue/latency-bound, not DRAM-bound (DESIGN.md 4.1, profiles/)"},
        "e2e": {"value": e_val, "unit": UNIT, "h2d_bytes_per_step": h2d_views + h2d_masks, "d2h_bytes_per_step": d2h,
                "tiles_per_step_per_gpu": Ne, "api": "pistoseg_b200.ops.fuse_argmax_confusion_host -> pisto_fuse_argmax_confusion_host (pinned host buffers, 3-slot pipeline)",
                "timed_by": "CUDA events inside the library (first H2D start -> last D2H end), max over ranks",
                "wall_value": world * Ne * args.e2e_steps / (ew_ms * 1e-3), "wall_timed_by": "time.perf_counter() around the Python calls, max over ranks",
                "h2d_gbs_per_rank": (h2d_views + h2d_masks) * args.e2e_steps / (e_ms * 1e-3) / 1e9,
                "d2h_gbs_per_rank": d2h * args.e2e_steps / (e_ms * 1e-3) / 1e9},
        "e2e_device_logits": {"value": world * Ne * args.e2e_steps / (d_ms * 1e-3), "unit": UNIT, "wall_value": world * Ne * args.e2e_steps / (dw_ms * 1e-3),
                              "h2d_bytes_per_step": h2d_masks, "d2h_bytes_per_step": d2h, "tiles_per_step_per_gpu": Ne,
                              "dataflow": "the reference's own (infer_pseudo_masks.py:119-137): logits resident on the GPU, tissue mask + label vector in, labels + 32x32 logits out",
                              "d2h_gbs_per_rank": d2h * args.e2e_steps / (d_ms * 1e-3) / 1e9},
        "gpu_launches": launches,
        "clocks": clocks,
        "affinity": affinity,
    }
    if cfg3:
        line["cfg3"] = cfg3
    if cfg5:
        line["cfg5_T1024"] = cfg5
    if world == 1 and not args.no_cpu_baseline:
        v, dt = cpu_reference_run(args.cpu_tiles)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                "sample": f"{args.cpu_tiles} tiles of the same workload, oracle/pipeline.py (torch-CPU restatement of infer_pseudo_masks.py:118-154), {dt:.1f} s"}
        # the reference pins itself to 2 threads (infer_pseudo_masks.py:22-28): same port, 2 threads, smaller sample
        v2, dt2 = cpu_reference_run(256, threads=2)
        line["cpu_baseline"]["value_2_threads"] = v2
        line["cpu_baseline"]["sample_2_threads"] = f"256 tiles, torch.set_num_threads(2), {dt2:.1f} s"
        # the honest comparator: the same literal torch code on CUDA tensors of the same box -- which is how the
        # reference really runs (infer_pseudo_masks.py:118-137: eager kernels, two .cpu() syncs per tile).  Untimed-region only.
        try:
            from oracle import pipeline
            ng = 2048
            cfg = synthetic.cfg2(N=ng, T=T, C=C, scales=SCALES)
            gv = [x.to(dev) for x in cfg["views"]]
            pres_np, bg_np = cfg["present"].numpy(), cfg["bg"].numpy()
            pipeline.pseudo_mask_batch([x[:64] for x in gv], cfg["codes"], (T, T), pres_np[:64], bg_np[:64])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            pipeline.pseudo_mask_batch(gv, cfg["codes"], (T, T), pres_np, bg_np)
            torch.cuda.synchronize()
            dtg = time.perf_counter() - t0
            line["gpu_eager_baseline"] = {"value": ng / dtg, "unit": UNIT, "kind": "port of the reference's torch code run on CUDA tensors (eager, per-tile .cpu() syncs)",
                                          "sample": f"{ng} tiles, oracle/pipeline.py on cuda:{local}, {dtg:.2f} s, batch 32"}
        except Exception as e:  # pragma: no cover
            line["gpu_eager_baseline"] = {"unavailable": repr(e)[:200]}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tiles", type=int, default=16384, help="tiles per step per GPU")
    ap.add_argument("--e2e-tiles", type=int, default=8192)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-chunk", type=int, default=1024)
    ap.add_argument("--cpu-tiles", type=int, default=1536, help="bounded CPU sample")
    ap.add_argument("--cpu-budget", type=float, default=90.0, help="seconds of CPU work for the whole --impl reference run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the cfg3 / cfg5 blocks (profiling runs)")
    ap.add_argument("--skip-check", action="store_true")
    ap.add_argument("--single-frac", type=float, default=0.4, help="fraction of single-label tiles (SURVEY.md 8(d) cfg 2: 0.4)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
