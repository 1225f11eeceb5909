#!/usr/bin/env python
"""bench.py -- the driver's benchmark contract for the acg_b200 hot path.

Workload (BASELINE.json configs[2], the one the headline metric "GAN train frames/sec" is quoted on): one full
adversarial DNA training iteration, --loss bce --opt adam --dna, = 1 x Trainer.train_d + 1 x Trainer.train_g
(train.py:241-263) at batch 256 PER GPU (weak scaling), 64x64 RGB frames, 10-D action++state, ksize=6 (what
train.py:53-54 passes), random-init weights, synthetic Push-shaped data.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (one rank per GPU under torchrun)
  python bench.py --impl reference [...]                         the reference's CPU path (oracle port)

One JSON line on stdout (rank 0).  `value` = frames/s with the feeds resident in HBM; `e2e` = the same metric
through the public Trainer.train_d / Trainer.train_g calls with HOST (pinned) feeds, H2D copies and the D2H
fetch of the generated frames that train_g returns (train.py:124,130) inside the timed region.
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
import torch  # noqa: E402

BATCH = 256
KSIZE = 6
CPU_BATCH = 16          # BASELINE.json configs[0]: the reference's CPU-runnable case


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  NVML is polled from a thread every 5 ms (the
    timed region of the default run is ~40 ms: `nvidia-smi -lms 100` can miss it entirely); nvidia-smi is the fallback."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml, self.handle = index, [], None, None, None
        self._stop = threading.Event()
        self.sm, self.bits, self.sm_max = [], 0, None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        return pynvml, h

    def start(self):
        try:
            self.nvml, self.handle = self._nvml_handle()
            self.sm_max = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        reasons_fn = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                self.bits |= int(reasons_fn(self.handle))
            except Exception:
                pass
            time.sleep(0.005)

    def sample_now(self):
        """One sample taken from the calling thread (the polling thread can be starved of the GIL while the main thread
        enqueues the timed iterations); called mid-way and at the end of the timed loop."""
        if self.nvml is None:
            return
        try:
            n = self.nvml
            reasons_fn = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons")
            self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
            self.bits |= int(reasons_fn(self.handle))
        except Exception:
            pass

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            if not self.sm:
                time.sleep(0.01)
            self._stop.set()
            self.thread.join(timeout=1)
            reasons = sorted(name for bit, name in self.REASONS.items() if self.bits & bit)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                    "reasons": reasons, "samples": len(self.sm), "source": "nvml, 5 ms"}
        if self.proc is None:
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
            if len(r) < 6:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def synth_batch(B, seed, pinned):
    """Push-shaped synthetic feeds: frames in [-1,1], next = clip(img + 0.1 N(0,1))."""
    rng = np.random.RandomState(seed)
    img = rng.uniform(-1, 1, (B, 64, 64, 3)).astype(np.float32)
    nxt = np.clip(img + 0.1 * rng.randn(B, 64, 64, 3), -1, 1).astype(np.float32)
    act = rng.randn(B, 10).astype(np.float32)
    state = rng.randn(B, 5).astype(np.float32)
    ts = [torch.from_numpy(a) for a in (img, nxt, act, state)]
    if pinned:
        ts = [t.pin_memory() for t in ts]
    return ts


def flops_per_iter(B, K):
    """Nominal 2*M*N*K FLOPs of one train_d + train_g."""
    per = 4.28e9 if K == 6 else 3.99e9
    return per * B


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference's CPU path, all host threads
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, B=CPU_BATCH):
    from oracle import np_ref, torch_ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rng = np.random.RandomState(7)
    params = np_ref.init_params(np_ref.g_dna_spec(KSIZE), rng)
    params.update(np_ref.init_params(np_ref.d_spec(), rng))
    ora = torch_ref.Trainer(params, True, "bce", "adam", True, ksize=KSIZE, dtype=torch.float32)
    img, nxt, act, state = [t.numpy() for t in synth_batch(B, 1, False)]
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        ora.train_d(img, nxt, act)
        ora.train_g(img, nxt, act, state)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    return {"value": B * len(times) / total, "ms_per_step": 1e3 * total / len(times), "cores": cores,
            "sample": "%d iteration(s) of train_d+train_g at batch %d (configs[0] size), fp32 torch-CPU oracle port of "
                      "models.py/ops.py/train.py, %d threads" % (len(times), B, cores)}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    res = cpu_reference_run(steps, min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": "GAN train frames/sec", "value": res["value"], "unit": "frames/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "full adversarial DNA step bce+adam (train_d + train_g), 64x64x3, ksize=6; CPU sample "
                               "at batch %d" % CPU_BATCH, "global_batch": CPU_BATCH},
        "cpu_baseline": {"value": res["value"], "unit": "frames/s", "cores": res["cores"], "kind": "port",
                         "sample": res["sample"]},
        "e2e": {"value": res["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "TF 1.0 cannot be installed here (SURVEY.md section 8(c)); this is the oracle port on host cores",
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
def time_kernel(fn, iters, flush=None):
    """Average device time (ms) of one fn() launch: `iters` back-to-back launches captured in a CUDA graph (no host
    launch cost between them), replayed 3 times between CUDA events on the launching stream; fn rotates its buffer
    sets so that consecutive launches never find their inputs in L2."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    reps = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / (iters * reps)


def dna_microbench(dev, peaks, B=256, K=KSIZE):
    """BASELINE configs[1]: DNA transform fwd / bwd alone.  B=256 makes the working set (176 / 317 MB at K=6)
    larger than the 126 MB L2; B=64 rotates 8 buffer sets for the same reason."""
    from action_conditioned_gans_b200 import kernels as Kn
    out = {}
    for (b, k, nset) in ((B, K, 4), (64, 5, 12)):
        sets = []
        for i in range(nset):
            g = torch.Generator(device=dev).manual_seed(i)
            sets.append((torch.randn(b, 64, 64, k * k, device=dev, generator=g),
                         torch.rand(b, 64, 64, 3, device=dev, generator=g) * 2 - 1,
                         torch.randn(b, 64, 64, 3, device=dev, generator=g),
                         torch.empty(b, 64, 64, 3, device=dev), torch.empty(b, 64, 64, k * k, device=dev)))
        it = [0]

        def fwd():
            lg, im, dy, o, dl = sets[it[0] % nset]
            it[0] += 1
            Kn.dna_fwd(lg, im, o, k)

        def bwd():
            lg, im, dy, o, dl = sets[it[0] % nset]
            it[0] += 1
            Kn.dna_bwd(lg, im, dy, dl, k)

        tf, tb = time_kernel(fwd, 20), time_kernel(bwd, 20)
        bytes_f = b * 4096 * (k * k + 6) * 4
        bytes_b = b * 4096 * (2 * k * k + 6) * 4
        out["B%d_K%d" % (b, k)] = {
            "fwd_us": 1e3 * tf, "fwd_gbs": bytes_f / tf / 1e6, "fwd_frac": bytes_f / tf / 1e6 / peaks["hbm_gbs"],
            "bwd_us": 1e3 * tb, "bwd_gbs": bytes_b / tb / 1e6, "bwd_frac": bytes_b / tb / 1e6 / peaks["hbm_gbs"],
            "buffer_sets": nset}
    return out


def main_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the acg_b200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from action_conditioned_gans_b200 import _lib
    from action_conditioned_gans_b200.trainer import DataParallel, Trainer
    _lib.load()
    dp = None
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner on STDOUT when the communicator is created; the contract is ONE JSON line
        # there, so stdout points at stderr while the communicator comes up (init + first collective).
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
        dp = DataParallel()
    peaks = load_peaks()
    B = args.batch
    trn = Trainer(None, True, "bce", "adam", True, batch_size=B, ksize=KSIZE, device=dev, seed=7, dp=dp,
                  branches=not args.no_branches)

    # feeds: a few distinct host batches (pinned), rotated; device copies for the HBM-resident measurement
    nfeeds = 2
    host = [synth_batch(B, 100 + rank * 10 + i, True) for i in range(nfeeds)]
    resident = [[t.to(dev) for t in hb] for hb in host]

    def iteration_resident(i):
        img, nxt, act, state = resident[i % nfeeds]
        trn.enqueue_train_d(img, nxt, act)
        trn.enqueue_train_g(img, nxt, act, state)

    def barrier():
        if dp is not None:
            dp.dist.barrier()
        torch.cuda.synchronize()

    # W >= 3 warm-up iterations (eager, graph capture, first replay) + 5 more replays so that the timed region starts
    # with captured graphs and settled clocks
    n_warm = max(args.warmup, 3) + 5
    for i in range(n_warm):
        iteration_resident(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count() + trn.replayed_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        iteration_resident(i)
        if rank == 0 and (i == args.steps // 2 or i == args.steps - 1):
            sampler.sample_now()      # the device is busy with the queued iterations at this point
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() + trn.replayed_launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if dp is not None:
        dp.dist.all_reduce(t, op=dp.dist.ReduceOp.MAX)
    ms = float(t.item())
    value = B * world * args.steps / (ms / 1e3)

    # ---- end to end through the public API with host feeds ---------------------------------------------------
    def iteration_e2e(i):
        img, nxt, act, state = host[i % nfeeds]
        trn.train_d(img, nxt, act)
        return trn.train_g(img, nxt, act, state)          # returns the generated frames on the host

    e2e_steps = max(2, min(args.steps, 10))
    iteration_e2e(0)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        frames = iteration_e2e(i)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if dp is not None:
        dp.dist.all_reduce(t, op=dp.dist.ReduceOp.MAX)
    e2e_value = B * world * e2e_steps / float(t.item())
    h2d = sum(x.numel() * 4 for x in host[0][:3]) + sum(x.numel() * 4 for x in host[0])
    d2h = frames.nbytes

    # ---- per-kernel breakdown of one iteration (instrumented pass, right after the timed region; every rank
    # takes part because the step contains collectives) -----------------------------------------------------
    breakdown = kernel_breakdown(trn, resident[0])
    barrier()
    if rank != 0:
        hard_exit()       # no collective is issued after this point; see hard_exit()
    dna = dna_microbench(dev, peaks)
    flops = flops_per_iter(B, KSIZE)
    conv_ms = sum(v["ms"] for k, v in breakdown.items() if k.startswith("acg_conv"))
    top = max(breakdown.items(), key=lambda kv: kv[1]["ms"])
    if top[0].startswith("acg_conv"):
        tens_peak = peaks["bf16_tflops_sustained"]
        achieved = flops / (conv_ms / 1e3) / 1e12
        roofline = {"bound": "tensor", "kernel": "conv family (%s largest)" % top[0], "achieved": achieved,
                    "peak": tens_peak, "unit": "TFLOP/s", "frac": achieved / tens_peak, "traffic": None,
                    "peak_src": peaks["src"] + " (sustained cuBLAS bf16)",
                    "note": "nominal conv FLOPs of one iteration / summed conv-kernel device time"}
    else:
        k = "B%d_K%d" % (256, KSIZE)
        roofline = {"bound": "hbm", "kernel": "dna_bwd", "achieved": dna[k]["bwd_gbs"], "peak": peaks["hbm_gbs"],
                    "unit": "GB/s", "frac": dna[k]["bwd_frac"], "traffic": None, "peak_src": peaks["src"]}
    cpu = cpu_reference_run(1, 1)
    line = {
        "metric": "GAN train frames/sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": n_warm, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": trn.precision, "data": "synthetic",
        "config": {"workload": "full adversarial DNA step bce+adam: 1 x train_d + 1 x train_g per step "
                               "(BASELINE configs[2]), 64x64x3 frames, 10-D action++state, ksize=6, batch 256 per GPU",
                   "global_batch": B * world, "batch_per_gpu": B, "parallelism": "dp%d" % world,
                   "l2": "activations of one step (>1 GB) exceed the 126 MB L2; %d feed sets rotated" % nfeeds},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps,
                "note": "separately timed through Trainer.train_d/train_g with pinned HOST feeds; the H2D copies ride a "
                        "copy stream under the previous call's kernels and the frames leave while the backward pass runs"},
        "roofline": roofline,
        "dna_roofline": dna,
        "kernel_breakdown_ms": {k: round(v["ms"], 4) for k, v in sorted(breakdown.items(), key=lambda kv: -kv[1]["ms"])},
        "cpu_baseline": {"value": cpu["value"], "unit": "frames/s", "cores": cpu["cores"], "kind": "port",
                         "sample": cpu["sample"]},
    }
    print(json.dumps(line), flush=True)
    if dp is not None:
        hard_exit()


def hard_exit():
    """Multi-rank runs leave through os._exit: tearing down an NCCL communicator whose collectives were captured into
    CUDA graphs hung at interpreter exit on this stack (the JSON line was already out); nothing needs the teardown."""
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def kernel_breakdown(trn, feeds):
    """Device time per C-ABI entry point over ONE iteration: CUDA events around every call (instrumented pass)."""
    from action_conditioned_gans_b200 import kernels as Kn
    records = []
    orig = Kn.call

    def timed_call(name, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig(name, *a)
        e1.record()
        records.append((name, e0, e1))

    Kn.call = timed_call
    graphs, trn.use_graphs = trn.use_graphs, False        # the per-kernel view needs eager launches
    from action_conditioned_gans_b200 import engine as En
    En.Branch.enabled = False                             # ... on ONE stream (no overlapping kernels)
    try:
        img, nxt, act, state = feeds
        trn.enqueue_train_d(img, nxt, act)
        trn.enqueue_train_g(img, nxt, act, state)
        torch.cuda.synchronize()
    finally:
        Kn.call = orig
        trn.use_graphs = graphs
        En.Branch.enabled = True
    out = {}
    for name, e0, e1 in records:
        d = out.setdefault(name, {"ms": 0.0, "n": 0})
        d["ms"] += e0.elapsed_time(e1)
        d["n"] += 1
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-branches", action="store_true", help="every chain of a step on one stream (A/B switch)")
    ap.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    a = ap.parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
