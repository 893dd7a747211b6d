#!/usr/bin/env python
"""bench.py -- image pairs/s of the GFNet hot path (coarse match + local correlation at every scale /
iteration / pass + post-process + balanced sampling with kde + homography + corner error).

  python bench.py --gpus N --steps K --warmup W            our arm (one process per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W    the reference's own CPU functions on host cores, rank 0 only
  python bench.py --config map224|map672 ...               the other BASELINE.json workloads (native pyramids)

Default workload (BASELINE.json configs[1], the one the metric is quoted on): vis_ir.json, 448x448 pairs (+ the 560
upsample pass the reference always runs), num_itr = 2, 32 pairs per GPU (op batch 64, symmetric), synthetic pyramids /
random H.  One JSON line on stdout (rank 0).
"""
import argparse
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "image pairs/s at 448^2 (corr+kde+H)"
UNIT = "pairs/s"
NUM_ITR = 2
# config -> (pairs per GPU, scaling): map672 is BASELINE config 4's fixed global batch of 64 (strong scaling)
CONFIGS = {"visir448": (32, "weak"), "map224": (32, "weak"), "map672": (64, "strong")}


def load_synth():
    """gfnet_b200/synth.py without importing the package (the reference arm must not load our .so)."""
    spec = importlib.util.spec_from_file_location("gfb_synth_standalone", os.path.join(ROOT, "gfnet_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, timed=None):
        """Median SM clock over every sample between start() and now -- the sampler runs from before the warm-up to the end of
        the end-to-end loop, all of it under load (the timed region alone is ~50 ms: one 50 ms sample at best) -- plus how many
        of the samples fell inside ``timed = (t0, t1)`` (host clock)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        self.in_timed = sum(1 for (t, _) in self.rows if timed and timed[0] <= t <= timed[1] + 0.05)
        for (_, r) in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "samples_in_timed_region": self.in_timed,
                "window": "warm-up + timed steps + pipelined steps + end-to-end loop (all under load), 50 ms period",
                "reasons": sorted(reasons)}


def metric_name(args, synth):
    return METRIC.replace("448", str(synth.WORKLOADS[args.config][0]))


def config_dict(args, n_gpus, pairs_per_gpu, synth, extra=None):
    res, up, desc = synth.WORKLOADS[args.config]
    cfg = {"workload": desc, "name": args.config, "res": res, "upsample_res": up,
           "pairs_per_gpu": pairs_per_gpu, "global_pairs": pairs_per_gpu * n_gpus, "num_samples": 5000,
           "kde_points": 20000, "homography": "cv2.findHomography(RANSAC, 3 px, 0.99999) restated on the device (adaptive iterations)",
           "parallelism": f"pairs sharded over {n_gpus} GPU(s); one all_gather_into_tensor of the [steps,B,12] result rows after the last step",
           "l2": "inputs + outputs per step >> 126 MB L2 (no flush needed)"}
    if extra:
        cfg.update(extra)
    return cfg


def pairs_per_gpu(args, world):
    if args.pairs_per_gpu:
        return args.pairs_per_gpu
    base, scaling = CONFIGS[args.config]
    return max(1, base // world) if scaling == "strong" else base


def cpu_reference_run(synth, args, steps, warmup, sample_pairs=1):
    """The reference's CPU implementation of the path on this box's host cores: its OWN functions where they import
    (oracle.reference), the oracle port otherwise.  Returns (pairs/s, cpu_baseline dict)."""
    import torch
    from oracle import reference as R
    from oracle.pipeline import cpu_hot_path
    impl = "reference" if R.available() else "port"
    if impl == "reference":
        R.load_reference()                               # import cost outside the timed region
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    res, up, _ = synth.WORKLOADS[args.config]
    batches = [synth.PairBatch(sample_pairs, res=res, upsample_res=up, num_itr=NUM_ITR, seed=1234, device="cpu", pair_offset=i)
               for i in range(2)]
    for w in range(warmup):
        cpu_hot_path(batches[w % 2], impl=impl)
    timings = {}
    t0 = time.perf_counter()
    for s_ in range(steps):
        cpu_hot_path(batches[s_ % 2], timings=timings, impl=impl)
    dt = time.perf_counter() - t0
    value = sample_pairs * steps / dt
    what = ("the reference's own local_correlation / GFNet.corr_volume / pos_embed / sample (kde) (+ the torch calls of ConvRefiner.forward :537-555) imported from "
            + R.reference_root() if impl == "reference" else "oracle port of the reference's torch calls")
    base = {"value": value, "unit": UNIT, "cores": cores, "kind": impl,
            "sample": f"{sample_pairs} pair per step x {steps} steps of the same workload: {what}; torch {torch.__version__} CPU with "
                      f"{cores} threads, kde(half=False, down=8) as the reference does on CPU, cv2.findHomography RANSAC",
            "seconds_by_stage_per_pair": {k: v / (sample_pairs * steps) for k, v in timings.items()}}
    return value, base, dt


def gpu_reference_run(synth, args, dev, pairs=2, steps=3, warmup=1):
    """the reference's OWN functions on the SAME B200 (torch CUDA kernels; fp16 kde at down = 1 and cv2 on the
    host, exactly what the reference does on a GPU) -- a secondary baseline next to the CPU arm the task defines.  Bounded:
    `pairs` pairs per step.  Returns a dict or None when the reference sources are not staged."""
    import torch
    from oracle import reference as R
    from oracle.pipeline import cpu_hot_path
    if not R.available():
        return None
    R.load_reference()
    res, up, _ = synth.WORKLOADS[args.config]
    batches = [synth.PairBatch(pairs, res=res, upsample_res=up, num_itr=NUM_ITR, seed=1234, device=dev, pair_offset=i) for i in range(2)]
    for w in range(warmup):
        cpu_hot_path(batches[w % 2], impl="reference")
    torch.cuda.synchronize()
    timings = {}
    t0 = time.perf_counter()
    for s_ in range(steps):
        cpu_hot_path(batches[s_ % 2], timings=timings, impl="reference")
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"value": pairs * steps / dt, "unit": UNIT, "pairs_per_step": pairs, "steps": steps,
            "what": "the reference's own local_correlation / corr_volume / pos_embed / sample (kde half=True, down=1) and the torch "
                    "calls of ConvRefiner.forward :537-555 on this GPU (torch " + torch.__version__ + " CUDA kernels, fp32 features), "
                    "cv2.findHomography on the host; wall clock with a device synchronise on both sides",
            "seconds_by_stage_per_pair": {k: v / (pairs * steps) for k, v in timings.items()}}


def run_reference(args):
    """--impl reference: rank 0 times the reference's CPU path (see cpu_reference_run); other ranks exit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    synth = load_synth()
    value, base, dt = cpu_reference_run(synth, args, args.steps, args.warmup)
    ppg = pairs_per_gpu(args, args.gpus)
    print(json.dumps({"impl": "reference", "metric": metric_name(args, synth), "value": value, "unit": UNIT, "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
                      "higher_is_better": True, "scaling": CONFIGS[args.config][1], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": config_dict(args, args.gpus, ppg, synth, {"reference_sample_pairs_per_step": 1}),
                      "cpu_baseline": base,
                      "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def run_ours(args):
    import torch
    import torch.distributed as dist
    import gfnet_b200 as gf
    from gfnet_b200 import synth
    from gfnet_b200.pipeline import HotPath

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (the product has no CPU path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    from gfnet_b200.dist import gather_equal
    B = pairs_per_gpu(args, world)
    res, up, _ = synth.WORKLOADS[args.config]
    batch = synth.PairBatch(B, res=res, upsample_res=up, num_itr=NUM_ITR, seed=1234, device=dev, rank=rank)
    hp = HotPath()
    gen = torch.Generator(device=dev).manual_seed(4321 + rank)
    rows = torch.zeros((max(args.steps, 8), B, 12), device=dev, dtype=torch.float64)     # result rows of every timed step

    def step(b=batch, k=0):
        out = hp.run(b, generator=gen)
        rows[k].copy_(out["result"])
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)                # nvidia-smi needs ~0.2 s before its first sample
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    hp.timing = []                     # CUDA-event pairs around every local_correlation launch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_host0 = time.perf_counter()
    e0.record()
    for k in range(args.steps):
        out = step(k=k)
    all_rows = gather_equal(rows[:args.steps])      # the path's only collective: once, after the last step
    e1.record()
    t_host = time.perf_counter() - t_host0          # host time to ISSUE the steps (no sync): launch-bound if ~ device time
    barrier()
    assert all_rows.shape[0] == world
    t_timed = (t_host0, time.perf_counter())
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    lc_ms = sum(a.elapsed_time(b) for (_, a, b) in hp.timing)
    lc_by_scale = {}
    for (key, a, b_) in hp.timing:
        lc_by_scale[key] = lc_by_scale.get(key, 0.0) + a.elapsed_time(b_)
    n_lc = len(hp.timing)
    hp.timing = None
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    value = B * world * args.steps / (total_ms / 1e3)

    # ---- two-stream pipelining (reported next to `value`, which stays the plain back-to-back measurement the roofline is
    # taken from): the sparse tail of step k (post-process, top-k, kde, homography: per-pair kernels on 32 CTAs) runs on a
    # second stream under the dense part of step k + 1
    s_back = torch.cuda.Stream(device=dev)

    def step_pipelined(k):
        cur = torch.cuda.current_stream(dev)
        hp.run_front(batch)
        ev = torch.cuda.Event()
        ev.record(cur)
        with torch.cuda.stream(s_back):
            s_back.wait_event(ev)                  # the tail consumes the dense part's final flow in the real model
            o = hp.run_back(batch, generator=gen)
            rows[k].copy_(o["result"])

    for k in range(3):
        step_pipelined(k)
    torch.cuda.current_stream(dev).wait_stream(s_back)
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    hp.timing = []
    p0.record()
    for k in range(args.steps):
        step_pipelined(k)
    torch.cuda.current_stream(dev).wait_stream(s_back)
    p1.record()
    barrier()
    lc_ms_pipelined = sum(a.elapsed_time(b_) for (_, a, b_) in hp.timing)
    hp.timing = None
    pms = torch.tensor([p0.elapsed_time(p1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(pms, op=dist.ReduceOp.MAX)
    pipelined = {"value": B * world * args.steps / (float(pms.item()) / 1e3), "unit": UNIT, "ms_per_step": float(pms.item()) / args.steps,
                 "what": "same K steps, the sparse tail of step k on a second CUDA stream under the dense part of step k + 1",
                 "local_correlation_ms_per_step": lc_ms_pipelined / args.steps}

    # ---- end to end: host (pinned) inputs -> device -> path -> host result, every step.
    # Two device-side input sets: while step i computes, the copy engine uploads step i+1's inputs on a second
    # stream (every step's upload -- feature maps of both images once, flows, logits; the grid features are computed on
    # the device by the refiner-input assembly -- and its [B,12] read-back are inside the timed region).
    batch_b = synth.PairBatch(B, res=res, upsample_res=up, num_itr=NUM_ITR, seed=1234, device=dev, rank=rank)
    sets = [(batch, batch.tensors()), (batch_b, batch_b.tensors())]
    pinned = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t) for t in sets[0][1]]
    h2d = sum(t.numel() * t.element_size() for t in pinned)
    res_host = [torch.empty((B, 12), dtype=torch.float64, pin_memory=True) for _ in range(2)]
    d2h = res_host[0].numel() * res_host[0].element_size()
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)

    def upload(k, after=None):
        with torch.cuda.stream(copy_stream):
            if after is not None:
                copy_stream.wait_event(after)          # the set is free once the step that read it has finished
            for dst, src_ in zip(sets[k][1], pinned):
                dst.copy_(src_, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev

    def e2e_run(n):
        done = [None, None]
        ready = upload(0)
        for i in range(n):
            k = i & 1
            nxt = upload(k ^ 1, done[k ^ 1]) if i + 1 < n else None
            main_stream.wait_event(ready)
            sets[k][0].derive()                        # image-B side of the op batch: device-side concatenation of the uploaded maps
            o = step(sets[k][0], k=i % rows.shape[0])
            res_host[k].copy_(o["result"], non_blocking=True)
            done[k] = torch.cuda.Event()
            done[k].record(main_stream)
            ready = nxt

    e2e_steps = max(2, min(args.steps, 8))
    e2e_run(2)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    e2e_run(e2e_steps)
    f1.record()
    barrier()
    ems = torch.tensor([f0.elapsed_time(f1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    e2e_value = B * world * e2e_steps / (float(ems.item()) / 1e3)
    del batch_b, sets
    clocks = sampler.stop(t_timed) if rank == 0 else None

    if rank == 0:
        peaks, src = measured_peaks()
        lc_bytes = 0
        by_scale = {}
        for pi, scales in enumerate(batch.passes):
            for sc in scales:
                nb = gf.local_correlation_bytes(2 * B, sc["c"], sc["hs"], sc["hs"], sc["G"], sc["r"]) * len(sc["flows"])
                lc_bytes += nb
                k = f"pass{pi + 1}_scale{sc['scale']}"
                t_ms = lc_by_scale.get(k, 0.0) / args.steps
                by_scale[k] = {"GBps": nb / t_ms / 1e6 if t_ms else None, "frac": nb / t_ms / 1e6 / peaks["hbm_gbs"] if t_ms else None,
                               "ms_per_step": t_ms}
        achieved = lc_bytes * args.steps / (lc_ms / 1e3) / 1e9 if lc_ms else None
        traffic = None                     # DRAM bytes of the same launches from the committed ncu capture (tools/lc_traffic.py)
        if args.config == "visir448" and B == 32:
            try:
                traffic = json.load(open(os.path.join(ROOT, "profiles", "r2_lc_dram_traffic.json")))["dram_bytes_per_step"]
            except Exception:
                pass
        roofline = {"kernel": "local_correlation: lc_tc2_kernel (+ lc_prep_plan_kernel) at C = 64, lc_mma_kernel at C = 32, lc_rot_kernel at C = 16; all scales/iterations/passes of a step", "bound": "hbm",
                    "achieved": achieved, "peak": peaks["hbm_gbs"], "peak_source": src, "unit": "GB/s",
                    "frac": achieved / peaks["hbm_gbs"] if achieved else None, "traffic": traffic,
                    "traffic_source": "profiles/r2_lc_dram_traffic.json (ncu dram__bytes_read+write, all 14 calls of a step)" if traffic else None,
                    "algorithmic_bytes_per_step": lc_bytes, "launches_per_step": n_lc // max(args.steps, 1),
                    "share_of_step": lc_ms / total_ms if total_ms else None, "by_scale": by_scale}
        # CPU baseline on this box's host cores: bounded sample (1 pair x 2 steps) of the reference's own functions
        cpu, gpu_ref = None, None
        if not args.no_cpu_baseline:
            _, cpu, _ = cpu_reference_run(synth, args, 2, 1)
            try:
                gpu_ref = gpu_reference_run(synth, args, dev)
            except Exception as exc:                   # secondary figure: never fail the bench line over it
                gpu_ref = {"unavailable": repr(exc)[:200]}
        tail = full = None
        if world == 1 and not args.no_extras:
            try:
                tail = refiner_tail_run(synth, args, B, dev)
            except Exception as exc:
                tail = {"unavailable": repr(exc)[:200]}
            try:
                full = full_decoder_run(synth, args, B, dev, hp, batch)
            except Exception as exc:
                full = {"unavailable": repr(exc)[:200]}
        line = {"metric": metric_name(args, synth), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": CONFIGS[args.config][1], "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config_dict(args, world, B, synth),
                "host_issue_ms_per_step": t_host / args.steps * 1e3,
                "clocks": clocks, "gpu_launches": hp.kernel_launches(batch) * args.steps,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                        "overlap": "upload of step i+1 on a copy stream while step i computes (two device input sets)"},
                "pipelined": pipelined, "roofline": roofline, "cpu_baseline": cpu, "reference_cuda": gpu_ref, "refiner_tail": tail, "full_decoder": full,
                "ace_px_mean": float(out["err"].mean()), "solved": int(out["status"].sum())}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def refiner_tail_run(synth, args, B, dev):
    """Extra key (outside the BASELINE metric): the refiner convolution tails (model/network.py:557-563) a step of
    the reference's decoder would run between the correlation calls -- one RefinerBlocks call per scale and pass at op batch 2B,
    random-init weights, times num_itr.  Scale 1 (no local correlation, model/network.py:139-153) included."""
    import torch
    from gfnet_b200 import refiner as RF
    res, up, _ = synth.WORKLOADS[args.config]
    ddim = {16: 64, 8: 64, 4: 32, 2: 16, 1: 8}
    shapes = []
    for pi, u in enumerate((None, up) if up else (None,)):
        for (s, c, hs, G, r) in synth.pyramid_config(res, upsample_res=u):
            shapes.append((f"pass{pi + 1}_scale{s}", 2 * c + ddim[s] + (2 * r + 1) ** 2, G))
        shapes.append((f"pass{pi + 1}_scale1", 2 * 8 + ddim[1], synth.final_grid(res, u) if u else 8 * (res // 14)))
    tot_ms, tot_flops, launches, rows = 0.0, 0.0, 0, {}
    for name, c, G in shapes:
        torch.manual_seed(c)
        blocks = [torch.nn.Sequential(torch.nn.Conv2d(c, c, 5, 1, 2, groups=c), torch.nn.BatchNorm2d(c), torch.nn.ReLU(inplace=True),
                                      torch.nn.Conv2d(c, c, 1, 1, 0)).to(dev).eval() for _ in range(9)]
        rb = RF.RefinerBlocks(blocks, torch.nn.Conv2d(c, 3, 1, 1, 0).to(dev).eval())
        d = torch.randn(2 * B, c, G, G, device=dev)
        rb(d)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            rb(d)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / 2
        rows[name] = {"C": c, "G": G, "ms": ms, "tflops": RF.refiner_blocks_flops(2 * B, c, G) / ms * 1e-9}
        tot_ms += ms * NUM_ITR
        tot_flops += RF.refiner_blocks_flops(2 * B, c, G) * NUM_ITR
        launches += rb.launches(2 * B, G) * NUM_ITR
        del rb, d, blocks
    return {"what": "9 blocks of depth-wise 5x5 + batch norm + ReLU + 1x1 convolution, then out_conv, per scale / pass / iteration "
                    "(fp16 activations, fp32 sums); not part of `value`", "ms_per_step": tot_ms, "tflops": tot_flops / tot_ms * 1e-9,
            "launches_per_step": launches, "by_scale": rows}



def full_decoder_run(synth, args, B, dev, hp, batch):
    """Extra key: everything after the backbone for B pairs -- ``decoder.refine`` (coarse match, then per scale and iteration
    refiner input + local correlation + convolution tail + flow update, upsampling; model/network.py:224-287) for the 448 pass and
    the 560 upsample pass on random feature pyramids and random-init refiners of the reference's widths, plus the sparse tail
    (post-process, sampling, kde, homography) of the bench step.  Not part of `value`."""
    import torch
    from gfnet_b200 import decoder
    res, up, _ = synth.WORKLOADS[args.config]
    ddim = {"16": 64, "8": 64, "4": 32, "2": 16, "1": 8}
    chans = {"16": 64, "8": 64, "4": 32, "2": 16, "1": 8}
    radius = {"16": 7, "8": 6, "4": 4, "2": 2, "1": 0}

    class Refiner(torch.nn.Module):                      # the structure and widths of model/network.py:76-155, 444-531
        def __init__(self, s):
            super().__init__()
            c, kk = chans[s], (2 * radius[s] + 1) ** 2 if s != "1" else 0
            dim = 2 * c + ddim[s] + kk
            mk = lambda: torch.nn.Sequential(torch.nn.Conv2d(dim, dim, 5, 1, 2, groups=dim), torch.nn.BatchNorm2d(dim),
                                             torch.nn.ReLU(inplace=True), torch.nn.Conv2d(dim, dim, 1, 1, 0))
            self.block1, self.hidden_blocks = mk(), torch.nn.Sequential(*[mk() for _ in range(8)])
            self.out_conv, self.disp_emb = torch.nn.Conv2d(dim, 3, 1, 1, 0), torch.nn.Conv2d(2, ddim[s], 1, 1, 0)
            self.local_corr_radius, self.corr_in_other, self.amp, self.amp_dtype = radius[s], s != "1", True, torch.float16

    torch.manual_seed(0)
    refiners = torch.nn.ModuleDict({s: Refiner(s) for s in ("16", "8", "4", "2", "1")}).to(dev).eval()
    gen = torch.Generator(device=dev).manual_seed(7)

    def pyramid(r, scales):
        size = {"16": r // 14, "8": r // 8, "4": r // 4, "2": r // 2, "1": r}
        return ({s: torch.randn((2 * B, chans[s], size[s], size[s]), generator=gen, device=dev) for s in scales},
                {s: torch.randn((2 * B, chans[s], size[s], size[s]), generator=gen, device=dev) for s in scales})

    g0 = res // 14
    f0a, f1a = pyramid(res, ("16", "8", "4", "2", "1"))
    grids_a = [g0, g0, 2 * g0, 4 * g0, 8 * g0]
    if up:
        f0b, f1b = pyramid(up, ("8", "4", "2", "1"))
        gu = up // 14
        grids_b = [gu, 2 * gu, 4 * gu, 8 * gu]

    def run():
        c = decoder.refine(f0a, f1a, refiners, grids_a, [NUM_ITR] * 5, res, res)
        if up:
            c = decoder.refine(f0b, f1b, refiners, grids_b, [NUM_ITR] * 4, up, up, scale_factor=up / res, upsample=True,
                               pre_corresps=c["1"][NUM_ITR])
        return c

    run()
    torch.cuda.synchronize(dev)
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    for _ in range(2):
        run()
    e1.record()
    for _ in range(2):
        hp.run_back(batch)
    e2.record()
    torch.cuda.synchronize(dev)
    dec_ms, back_ms = e0.elapsed_time(e1) / 2, e1.elapsed_time(e2) / 2
    return {"what": "decoder.refine for the 448 pass + the 560 upsample pass (coarse match, refiner input + local correlation + "
                    "convolution tail + flow update per scale and iteration, upsampling) + the sparse tail of the bench step; "
                    "random pyramids, random-init refiners; not part of `value`",
            "decoder_ms_per_step": dec_ms, "sparse_tail_ms_per_step": back_ms, "pairs_per_s": B / ((dec_ms + back_ms) / 1e3)}



def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="visir448", choices=sorted(CONFIGS))
    ap.add_argument("--pairs-per-gpu", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the refiner_tail / full_decoder extra keys (N = 1 only)")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = 3 if args.steps is None else min(args.steps, 6)     # ~6 s of CPU work per pair
        args.warmup = 1 if args.warmup is None else min(args.warmup, 2)
        run_reference(args)
    else:
        args.steps = 40 if args.steps is None else args.steps
        args.warmup = 3 if args.warmup is None else args.warmup
        run_ours(args)


if __name__ == "__main__":
    main()
