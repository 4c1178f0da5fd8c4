#!/usr/bin/env python
"""Benchmark of the PAMR hot path on B200 (BASELINE.json metric: PAMR Mpix/s, 21 classes, 10 iterations).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (N > 1, one rank per GPU)

A step = run_pamr -> _rescale_and_clean -> pseudo_gtmask -> argmax ("PAMR + clean/argmax epilogue as in
stage_net", BASELINE.json configs[1]) over one synthetic batch of B=16 per GPU, 3x321x321 images, 21-class masks at
image resolution.  Weak scaling: every rank processes its own B=16 shard (configs[4]: B=128 over 8 GPUs); the uint8
label maps are all-gathered with NCCL inside the step, on a side stream, so that the gather of step i overlaps the
affinity kernel of step i+1.  Rank 0 prints ONE JSON line; besides the headline it carries
  roofline        the propagation launch against the measured HBM peak (+ the binding on-chip roof, from ncu)
  e2e             the same step from pinned HOST buffers through HostPipeline (H2D + D2H inside the timed region)
  e2e_real_shape  the shape stage_net really calls it with: masks at 81x81 uploaded, up-sampling on the device
  small_map       PAMR.forward at stage_net's mask sizes (the resident kernel: one launch for all iterations) and the
                  whole stage step with masks of that size (ms_stage_step: image down, PAMR, clean + labels at 321x321)
  configs         BASELINE.json configs[2], [3] and [4]-on-one-GPU (N = 1 only)
  strong          configs[4] as stated: B=128 in total, split over the N GPUs
  allgather       the NCCL gather alone (N > 1)
  cpu_baseline    the CPU port of the reference path on the host cores (N = 1 only)
See DESIGN.md "Measurement" for the definitions.
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
B_STRONG = 128  # BASELINE.json configs[4]
P_TAPS = 8 * len(D6)
METRIC = "PAMR Mpix/s (21 cls, 10 iters)"
UNIT = "Mpix/s"
# algorithmic bytes of ONE propagation launch per pixel: affinity 4P + mask in 4C + mask out 4C
BYTES_PER_PIXEL_PROPAGATE = 4 * (P_TAPS + 2 * C_CLS)
HBM_FALLBACK_GBS = 6650.0  # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def step_bytes_per_pixel(K=K_IMG, C=C_CLS, iters=ITERS):
    """Algorithmic HBM bytes of the whole step per pixel: affinity (K + P) + iters * (P + 2C) + epilogue (C + 1/4)."""
    return 4 * (K + P_TAPS + iters * (P_TAPS + 2 * C)) + 4 * C + 1


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def ncu_facts():
    """DRAM traffic and shared-memory pipe utilisation of the propagation kernel from the committed ncu capture of
    this build (profiles/propagate_ncu_facts.json, written by tools/ncu_facts.py from the .ncu-rep)."""
    try:
        with open(os.path.join(ROOT, "profiles", "propagate_ncu_facts.json")) as f:
            return json.load(f)
    except Exception:
        return {}


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
    """Multi-GPU runs: pin this rank's host threads to the CPUs next to its GPU (NVML's ideal affinity) BEFORE any
    pinned buffer is allocated, so that the staging buffers of the end-to-end path are first-touched on that NUMA
    node.  Best effort: silently skipped if NVML is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
    except Exception:
        pass


def cpu_baseline_sample(steps, warmup, images_per_step=1):
    """Times the CPU port of the reference path (oracle/pamr_oracle.c, OpenMP over all host cores) on a bounded
    sample of the workload: `images_per_step` 321x321 images of the batch per step."""
    import numpy as np
    import synth
    from oracle import oracle
    n = images_per_step
    img = synth.image_uniform(n, K_IMG, H_IMG, W_IMG, 0)
    msk = synth.mask_softmax(n, C_CLS, H_IMG, W_IMG, 1)
    lab = synth.labels_bernoulli(n, C_CLS, 2, p=0.3)
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
    assert out.shape == (n, H_IMG, W_IMG) and out.dtype == np.uint8
    return {"value": n * H_IMG * W_IMG / dt / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "B=%d of the batch per step (%dx%dx%dx%d image, %d classes, %d iterations, PAMR + clean/argmax), "
                      "%d timed steps after %d warm-up, OpenMP C port of the reference path"
                      % (n, n, K_IMG, H_IMG, W_IMG, C_CLS, ITERS, steps, warmup),
            "ms_per_step": dt * 1e3}


def config_dict(n_gpus):
    return {"workload": "configs[1] VOC training shape: B=%d per GPU, %dx%dx%d image, %d classes, mask at image "
                        "resolution, PAMR(%d, %s) + clean/argmax epilogue" % (B_PER_GPU, K_IMG, H_IMG, W_IMG, C_CLS,
                                                                              ITERS, D6),
            "global_batch": B_PER_GPU * n_gpus, "per_gpu_batch": B_PER_GPU, "parallelism": "batch-shard dp%d" % n_gpus,
            "collective": "NCCL all-gather of uint8 labels on a side stream (overlaps the next step)" if n_gpus > 1 else "none",
            "l2_policy": "no flush: inputs+scratch per step (%.0f MB) exceed the 126 MB L2"
                         % (4e-6 * B_PER_GPU * H_IMG * W_IMG * (K_IMG + 3 * C_CLS + P_TAPS))}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (here the C port in oracle/; the reference is
    pure Python and cannot travel to the GPU box), all host threads, rank 0 only.  A step is a bounded sample of the
    b200 arm's per-step workload: one 321x321 image per GPU of the run (N images at --gpus N), so that the per-N
    ratio of the two arms compares like with like while the host's core count stays what it is."""
    if rank != 0:
        return
    n = max(1, args.gpus)
    base = cpu_baseline_sample(max(1, args.steps), max(0, args.warmup), images_per_step=n)
    cfg = config_dict(args.gpus)
    cfg["workload"] = ("bounded sample of configs[1]: B=%d per step (1 image of the %d per GPU, x %d GPU(s)), %dx%dx%d, %d classes, "
                       "PAMR(%d, %s) + clean/argmax epilogue on the host CPU" % (n, B_PER_GPU, n, K_IMG, H_IMG, W_IMG, C_CLS, ITERS, D6))
    cfg["sample_batch"] = n
    cfg["collective"] = "none"
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": max(1, args.steps), "warmup": max(0, args.warmup), "ms_per_step": base["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg, "cpu_baseline": base,
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
    ap.add_argument("--no-extras", action="store_true", help="headline + roofline + e2e only (no configs / small maps / strong)")
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

    if world > 1:
        bind_to_gpu_numa_node(local_rank)  # before torch allocates anything pinned
    import torch
    import torch.distributed as dist
    import wseg_b200
    from wseg_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank if world > 1 else 0)
    torch.cuda.set_device(dev)
    args.warmup = max(args.warmup, 3)
    B, K, C, H, W = B_PER_GPU, K_IMG, C_CLS, H_IMG, W_IMG
    npix = B * H * W
    peak, peak_src = hbm_peak()

    def synth_inputs(b, h_img, w_img, h_msk, w_msk, seed):
        gen = torch.Generator(device=dev).manual_seed(seed + rank)
        image = torch.rand((b, K, h_img, w_img), generator=gen, device=dev)
        mask = torch.softmax(2.0 * torch.randn((b, C, h_msk, w_msk), generator=gen, device=dev), 1)
        labels = (torch.rand((b, C - 1), generator=gen, device=dev) < 0.3).float()
        labels[:, 0] = 1.0
        return image, mask, labels

    image, mask, labels = synth_inputs(B, H, W, H, W, 1234)
    pamr = wseg_b200.PAMR(ITERS, D6).to(dev)
    pamr_2x = wseg_b200.PAMR(2 * ITERS, D6).to(dev)
    gather = wseg_b200.OverlappedLabelGather(dev) if world > 1 else None

    def step(img, msk, lab):
        out = wseg_b200.refine_and_label(pamr, img, msk, lab)
        if world > 1:
            return gather.submit(out)  # NCCL all-gather on the side stream; the next step's kernels do not wait for it
        return out

    def barrier():
        if world > 1:
            gather.wait()
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        """Device time of `steps` calls (CUDA events on the launching stream, barrier + synchronize on both sides,
        maximum over the ranks).  Returns (ms per call, kernels launched by libpamr_b200 in the timed region)."""
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _lib.launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        if world > 1:
            gather.wait()  # the last gathers belong to the timed region
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, _lib.launch_count() - n0

    def propagate_roofline(pm, pm2, img, msk, n, b_, h_, w_):
        """Marginal cost of one propagation launch inside the forward path, (t(2*ITERS) - t(ITERS)) / ITERS, and what
        that is of the HBM roofline (360 B per pixel and launch)."""
        t1, _ = timed(lambda: pm(img, msk), n, 2)
        t2, _ = timed(lambda: pm2(img, msk), n, 2)
        ms = (t2 - t1) / ITERS
        gbs = BYTES_PER_PIXEL_PROPAGATE * b_ * h_ * w_ / (ms * 1e-3) / 1e9
        return ms, t1, gbs

    # ---- the propagation launch timed alone in a short run, before the GPU settles at its power cap (reported next to
    # the sustained figure below: the kernel is bound on chip, so it follows the SM clock; the HBM peak does not)
    for _ in range(args.warmup):
        step(image, mask, labels)
    ms_launch_burst, _, achieved_burst = propagate_roofline(pamr, pamr_2x, image, mask, 10, B, H, W)

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

    # ---- dominant kernel: the propagation launch (ITERS of them per step), timed live with CUDA events
    nrep = max(5, args.steps // 2)
    ms_launch, ms_fwd, achieved = propagate_roofline(pamr, pamr_2x, image, mask, nrep, B, H, W)
    ms_aff, _ = timed(lambda: wseg_b200.local_affinity(image, D6), nrep, 3)
    dec, cmax = pamr(image, mask, return_class_max=True)
    ms_epi, _ = timed(lambda: wseg_b200.pseudo_labels(dec, labels, None, cmax), nrep, 3)
    facts = ncu_facts()
    roofline = {"bound": "hbm", "kernel": "propagate (one of %d launches per step)" % ITERS, "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": facts.get("dram_bytes_per_launch"),
                "traffic_source": facts.get("source"),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": BYTES_PER_PIXEL_PROPAGATE * npix,
                "ms_per_launch": ms_launch, "ms_forward": ms_fwd,
                "burst": {"ms_per_launch": ms_launch_burst, "achieved": achieved_burst, "frac": achieved_burst / peak,
                          "note": "same measurement in a 10-repetition run before the timed steps (full SM clock)"},
                "ms_forward_other_than_propagation": ms_fwd - ITERS * ms_launch,
                "ms_affinity_standard_layout": ms_aff, "ms_epilogue": ms_epi,
                "whole_step_GBps": step_bytes_per_pixel() * npix / (ms_step * 1e-3) / 1e9,
                # the on-chip resource that binds this kernel is the shared-memory pipe, not HBM
                "secondary_roof": {"resource": "shared-memory pipe (LDS wavefronts)", "lds_pct": facts.get("lds_pct"),
                                   "ceiling": facts.get("lds_ceiling_frac_of_hbm_roofline"),
                                   "source": facts.get("source")}}

    # ---- end to end through the public API with HOST buffers (pinned), H2D + D2H inside the timed region
    n_e2e = max(20, args.steps // 3)
    pipe = wseg_b200.HostPipeline(pamr, dev, chunks=int(os.environ.get("PAMR_BENCH_CHUNKS", "1")))

    def e2e_record(h_image, h_mask, h_labels):
        # every rank reads back its own shard of the (gathered) label maps: each map reaches the host once per step
        h_out = torch.empty((B, H, W), dtype=torch.uint8).pin_memory()
        d_local = torch.empty((B, H, W), dtype=torch.uint8, device=dev)

        def e2e_step():
            if world == 1:
                pipe(h_image, h_mask, h_labels, h_out, out_size=(H, W))
            else:  # per-rank shard through the pipeline; the NCCL gather of the labels and the D2H of the own slice on the side stream
                pipe(h_image, h_mask, h_labels, d_out=d_local, out_size=(H, W))
                gather.submit(d_local, host_out=h_out)

        ms, _ = timed(e2e_step, n_e2e, 3)
        h2d = h_image.numel() * 4 + h_mask.numel() * 4 + h_labels.numel() * 4
        return {"value": world * npix / (ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": h_out.numel(), "ms_per_step": ms, "steps": n_e2e}

    e2e = e2e_record(image.cpu().pin_memory(), mask.cpu().pin_memory(), labels.cpu().pin_memory())
    # the shape stage_net really calls the path with (SoftMaxAE.py:176-179, 250-259): masks at 81x81 for a 321x321 crop;
    # the image resize, PAMR at 81x81 (resident kernel), up-sampling + clean + labels at 321x321 all run on the device
    _, mask_lo, _ = synth_inputs(B, H, W, 81, 81, 4321)
    e2e_real = e2e_record(image.cpu().pin_memory(), mask_lo.cpu().pin_memory(), labels.cpu().pin_memory())
    e2e_real["workload"] = "B=%d per GPU, image %dx%d and masks 81x81 uploaded, labels %dx%d downloaded" % (B, H, W, H, W)
    ms_real_dev, _ = timed(lambda: step(image, mask_lo, labels), nrep, 3)
    e2e_real["device_resident_ms_per_step"] = ms_real_dev

    extras = {}
    if not args.no_extras:
        # ---- stage_net's own PAMR call shapes: the resident small-map kernel (affinity + all iterations in one launch)
        small = []
        for hw in (41, 81):
            im_s, mk_s, _ = synth_inputs(B, hw, hw, hw, hw, 777 + hw)
            ms_s, l_s = timed(lambda: pamr(im_s, mk_s), max(20, nrep), 5)
            # the whole stage step at this mask size: image down, PAMR, up-sampling + clean + labels at the image size
            ms_st, _ = timed(lambda: step(image, mk_s, labels), max(20, nrep), 5)
            small.append({"shape": "B=%d %dx%d, %d classes, %d iterations" % (B, hw, hw, C, ITERS), "ms_forward": ms_s,
                          "launches_per_forward": l_s / max(20, nrep), "mpix_s": B * hw * hw / (ms_s * 1e-3) / 1e6,
                          "whole_forward_GBps": (step_bytes_per_pixel() - 4 * C - 1) * B * hw * hw / (ms_s * 1e-3) / 1e9,
                          "ms_stage_step": ms_st})
            del im_s, mk_s
        extras["small_map"] = small

        def config_record(name, b_, h_, w_, n):
            img_c, msk_c, lab_c = synth_inputs(b_, h_, w_, h_, w_, 99 + h_ + b_)
            ms_c, _ = timed(lambda: wseg_b200.refine_and_label(pamr, img_c, msk_c, lab_c), n, 2)
            ms_l, ms_f, gbs = propagate_roofline(pamr, pamr_2x, img_c, msk_c, max(3, n // 2), b_, h_, w_)
            rec = {"name": name, "ms_per_step": ms_c, "mpix_s": b_ * h_ * w_ / (ms_c * 1e-3) / 1e6,
                   "ms_per_propagate_launch": ms_l, "propagate_GBps": gbs, "frac": gbs / peak,
                   "whole_step_GBps": step_bytes_per_pixel() * b_ * h_ * w_ / (ms_c * 1e-3) / 1e9}
            del img_c, msk_c, lab_c
            torch.cuda.empty_cache()
            return rec

        # ---- BASELINE.json configs[4] as stated: B=128 in total, split over the GPUs of this run
        b_strong = B_STRONG // world
        img_s, msk_s, lab_s = synth_inputs(b_strong, H, W, H, W, 555)
        ms_strong, _ = timed(lambda: step(img_s, msk_s, lab_s), max(5, min(args.steps, 20)), 2)
        extras["strong"] = {"global_batch": B_STRONG, "per_gpu_batch": b_strong, "ms_per_step": ms_strong,
                            "value": B_STRONG * H * W / (ms_strong * 1e-3) / 1e6, "unit": UNIT,
                            "whole_step_GBps_per_gpu": step_bytes_per_pixel() * b_strong * H * W / (ms_strong * 1e-3) / 1e9}
        del img_s, msk_s, lab_s
        torch.cuda.empty_cache()
        if world == 1:
            recs = [config_record("configs[2] multi-scale inference, B=1 %dx%d" % (s, s), 1, s, s, 10) for s in (256, 512, 768, 1024)]
            recs.append(config_record("configs[3] high resolution, B=8 1024x2048", 8, 1024, 2048, 5))
            recs.append(config_record("configs[4] on one GPU, B=128 321x321", B_STRONG, H, W, 5))
            extras["configs"] = recs
        else:
            # ---- the collective alone: all-gather of this step's uint8 labels (not overlapped with anything)
            lab_u8 = wseg_b200.refine_and_label(pamr, image, mask, labels)
            buf = torch.empty((world * B, H, W), dtype=torch.uint8, device=dev)
            ms_ag, _ = timed(lambda: dist.all_gather_into_tensor(buf, lab_u8), 50, 5)
            extras["allgather"] = {"ms": ms_ag, "bytes_per_rank": lab_u8.numel(), "overlapped_in_step": True,
                                   "note": "issued on a side stream; the next step's kernels start without waiting for it"}

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line = {"metric": METRIC, "value": world * npix / (ms_step * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(world),
            "roofline": roofline, "clocks": clocks, "e2e": e2e, "e2e_real_shape": e2e_real,
            "gpu_launches": launches}
    line.update(extras)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_sample(20, 2)
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
