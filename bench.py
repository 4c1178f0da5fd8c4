#!/usr/bin/env python
"""Benchmark of the PAMR hot path on B200 (BASELINE.json metric: PAMR Mpix/s, 21 classes, 10 iterations).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (N > 1, one rank per GPU)

A step = run_pamr -> _rescale_and_clean -> pseudo_gtmask -> argmax ("PAMR + clean/argmax epilogue as
in stage_net", BASELINE.json configs[1]) over one synthetic batch of B=16 per GPU, 3x321x321 images,
21-class masks at image resolution.  Weak scaling: every rank processes its own B=16 shard
(configs[4]: B=128 over 8 GPUs) and the uint8 label maps are all-gathered with NCCL inside the step.
Rank 0 prints ONE JSON line.  See DESIGN.md "Measurement" for the definitions of each key.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

D6 = [1, 2, 4, 8, 12, 24]
B_PER_GPU, K_IMG, C_CLS, H_IMG, W_IMG, ITERS = 16, 3, 21, 321, 321, 10
P_TAPS = 8 * len(D6)
METRIC = "PAMR Mpix/s (21 cls, 10 iters)"
UNIT = "Mpix/s"
# algorithmic bytes of ONE propagation launch per pixel: affinity 4P + mask in 4C + mask out 4C
BYTES_PER_PIXEL_PROPAGATE = 4 * (P_TAPS + 2 * C_CLS)
HBM_FALLBACK_GBS = 6650.0  # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.thread = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def wait_first_sample(self, timeout=5.0):
        """nvidia-smi takes a moment to attach; do not let its start-up overlap the timed region."""
        t_end = time.time() + timeout
        while self.proc is not None and not self.rows and time.time() < t_end:
            time.sleep(0.02)

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15 and len(r) >= 7] or \
               [r for (_, r) in self.rows if len(r) >= 7]
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        pw = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(rows), "reasons": reasons}


def bind_to_gpu_numa_node(index):
    """Multi-GPU runs: pin this rank's host threads to the CPUs next to its GPU (NVML's ideal affinity), so that
    the pinned staging buffers of the end-to-end path are first-touched on that NUMA node and the N concurrent
    host->device streams do not cross sockets.  Best effort: silently skipped if NVML is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
    except Exception:
        pass


def cpu_baseline_sample(steps, warmup):
    """Times the CPU port of the reference path (oracle/pamr_oracle.c, OpenMP over all host cores) on a
    bounded sample of the workload: one 321x321 image of the batch per step."""
    import numpy as np
    import synth
    from oracle import oracle
    img = synth.image_uniform(1, K_IMG, H_IMG, W_IMG, 0)
    msk = synth.mask_softmax(1, C_CLS, H_IMG, W_IMG, 1)
    lab = synth.labels_bernoulli(1, C_CLS, 2, p=0.3)
    # all host threads, also under torchrun (which exports OMP_NUM_THREADS=1 to every rank)
    try:
        oracle.set_num_threads(len(os.sched_getaffinity(0)))
    except AttributeError:
        oracle.set_num_threads(os.cpu_count() or 1)
    cores = oracle.num_threads()

    def step():
        dec = oracle.run_pamr(img, msk, ITERS, D6)
        cleaned = oracle.rescale_and_clean(dec, (H_IMG, W_IMG), lab)
        return oracle.pseudo_labels(cleaned)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = step()
    dt = (time.perf_counter() - t0) / steps
    assert out.shape == (1, H_IMG, W_IMG) and out.dtype == np.uint8
    return {"value": H_IMG * W_IMG / dt / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "B=1 of the batch (1x3x%dx%d image, %d classes, %d iterations, PAMR + clean/argmax), "
                      "%d timed steps, OpenMP C port of the reference path" % (H_IMG, W_IMG, C_CLS, ITERS, steps),
            "ms_per_image": dt * 1e3}


def config_dict(n_gpus):
    return {"workload": "configs[1] VOC training shape: B=%d per GPU, %dx%dx%d image, %d classes, mask at image "
                        "resolution, PAMR(%d, %s) + clean/argmax epilogue" % (B_PER_GPU, K_IMG, H_IMG, W_IMG, C_CLS,
                                                                              ITERS, D6),
            "global_batch": B_PER_GPU * n_gpus, "per_gpu_batch": B_PER_GPU, "parallelism": "batch-shard dp%d" % n_gpus,
            "collective": "NCCL all-gather of uint8 labels" if n_gpus > 1 else "none",
            "l2_policy": "no flush: inputs+scratch per step (%.0f MB) exceed the 126 MB L2"
                         % (4e-6 * B_PER_GPU * H_IMG * W_IMG * (K_IMG + 3 * C_CLS + P_TAPS))}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (here the oracle port; the
    reference is pure Python and cannot travel to the GPU box), all host threads, rank 0 only."""
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 40)), max(1, min(args.warmup, 3))
    base = cpu_baseline_sample(steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": base["ms_per_image"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args.gpus), "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU leg (profiling runs)")
    ap.add_argument("--only", default="", help="profiling aid: 'step' runs only the timed step loop")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # plain `python bench.py --gpus N`: re-launch one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29400 + os.getpid() % 500), os.path.abspath(__file__)]
        raise SystemExit(subprocess.call(cmd + sys.argv[1:]))

    import torch
    import torch.distributed as dist
    import wseg_b200
    from wseg_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        bind_to_gpu_numa_node(local_rank)
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank if world > 1 else 0)
    torch.cuda.set_device(dev)
    args.warmup = max(args.warmup, 3)
    B, K, C, H, W = B_PER_GPU, K_IMG, C_CLS, H_IMG, W_IMG
    npix = B * H * W

    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    image = torch.rand((B, K, H, W), generator=gen, device=dev)
    mask = torch.softmax(2.0 * torch.randn((B, C, H, W), generator=gen, device=dev), 1)
    labels = (torch.rand((B, C - 1), generator=gen, device=dev) < 0.3).float()
    labels[:, 0] = 1.0
    pamr = wseg_b200.PAMR(ITERS, D6).to(dev)
    gathered = torch.empty((world * B, H, W), dtype=torch.uint8, device=dev) if world > 1 else None

    def step(img, msk, lab):
        out = wseg_b200.refine_and_label(pamr, img, msk, lab)
        if world > 1:
            dist.all_gather_into_tensor(gathered, out)
            return gathered
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _lib.launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, _lib.launch_count() - n0

    # ---- device-resident throughput (`value`)
    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
        sampler.wait_first_sample()
    for _ in range(args.warmup):  # untimed warm-up (clocks ramp, allocator pools fill) before the clock window opens
        step(image, mask, labels)
    barrier()
    t_wall0 = time.time()
    ms_step, launches = timed(lambda: step(image, mask, labels), args.steps, 0)
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    if args.only == "step":
        if rank == 0:
            print(json.dumps({"ms_per_step": ms_step, "launches_per_step": launches / args.steps}))
        return

    # ---- dominant kernel: the propagation launch (ITERS of them per step).  Its average duration is
    # measured live, with CUDA events on the launching stream, as the marginal cost of a launch inside
    # the forward path: (t_forward(2*ITERS) - t_forward(ITERS)) / ITERS.  (Same kernels, same layout,
    # no relayout/repack in the difference; the remainder-strip kernel runs concurrently on a side stream.)
    pamr_2x = wseg_b200.PAMR(2 * ITERS, D6).to(dev)
    nrep = max(5, args.steps // 2)
    ms_fwd, _ = timed(lambda: pamr(image, mask), nrep, 3)
    ms_fwd2, _ = timed(lambda: pamr_2x(image, mask), nrep, 3)
    ms_launch = (ms_fwd2 - ms_fwd) / ITERS
    ms_aff, _ = timed(lambda: wseg_b200.local_affinity(image, D6), nrep, 3)
    dec, cmax = pamr(image, mask, return_class_max=True)
    ms_epi, _ = timed(lambda: wseg_b200.pseudo_labels(dec, labels, None, cmax), max(5, args.steps // 2), 3)
    peak, peak_src = hbm_peak()
    achieved = BYTES_PER_PIXEL_PROPAGATE * npix / (ms_launch * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "propagate_dram_bytes.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "propagate (one of %d launches per step)" % ITERS, "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": BYTES_PER_PIXEL_PROPAGATE * npix,
                "ms_per_launch": ms_launch, "ms_forward": ms_fwd, "ms_affinity_standard_layout": ms_aff, "ms_epilogue": ms_epi,
                "whole_step_GBps": (4 * (K + P_TAPS + ITERS * (P_TAPS + 2 * C)) + 4 * C + 1) * npix / (ms_step * 1e-3) / 1e9}

    # ---- end to end through the public API with HOST buffers (pinned), H2D + D2H inside the timed region
    h_image, h_mask, h_labels = image.cpu().pin_memory(), mask.cpu().pin_memory(), labels.cpu().pin_memory()
    h_out = torch.empty((world * B if world > 1 else B, H, W), dtype=torch.uint8).pin_memory()

    # public host-input API: chunked H2D copies overlapped with the kernels (wseg_b200.HostPipeline)
    pipe = wseg_b200.HostPipeline(pamr, dev, chunks=int(os.environ.get("PAMR_BENCH_CHUNKS", "1")))
    d_local = torch.empty((B, H, W), dtype=torch.uint8, device=dev)

    def e2e_step():
        if world == 1:
            pipe(h_image, h_mask, h_labels, h_out)
        else:  # per-rank shard through the pipeline, then the NCCL gather of the labels, then D2H
            pipe(h_image, h_mask, h_labels, d_out=d_local)
            dist.all_gather_into_tensor(gathered, d_local)
            h_out.copy_(gathered, non_blocking=True)

    ms_e2e, _ = timed(e2e_step, max(3, args.steps // 3), 3)
    h2d = h_image.numel() * 4 + h_mask.numel() * 4 + h_labels.numel() * 4
    d2h = h_out.numel()

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line = {"metric": METRIC, "value": world * npix / (ms_step * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(world),
            "roofline": roofline, "clocks": clocks,
            "e2e": {"value": world * npix / (ms_e2e * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e},
            "gpu_launches": launches}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_sample(20, 2)
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
