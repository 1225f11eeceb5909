#!/usr/bin/env python
"""bench.py -- the driver's benchmark contract for the acg_b200 hot path.

Default workload (BASELINE.json configs[2], the one the headline metric "GAN train frames/sec" is quoted on): one full
adversarial DNA training iteration, --loss bce --opt adam --dna, = 1 x Trainer.train_d + 1 x Trainer.train_g
(train.py:241-263) at batch 256 PER GPU (weak scaling), 64x64 RGB frames, 10-D action++state, ksize=6 (what
train.py:53-54 passes), random-init weights, synthetic Push-shaped data.  --config wass_rmsprop is configs[3]
(5 x train_d + 1 x train_g, train.py:217-218), --config direct_rollout is configs[4] (6-step rollout, batch 512).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C]   our arm (one rank per GPU under torchrun)
  python bench.py --impl reference [...]                             the reference's CPU path (oracle port)

One JSON line on stdout (rank 0).  `value` = frames/s with the feeds resident in HBM (K steps between CUDA events, max
over ranks; `step_ms_median` / `step_ms_max` from one event per step boundary); `e2e` = the same metric through the
public Trainer calls with HOST (pinned) feeds, H2D copies and the D2H fetch of the generated frames that train_g
returns (train.py:124,130) inside the timed region; with N > 1 `strong_scaling` times the same GLOBAL batch split over
the ranks.
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


# ----------------------------------------------------------------------------------------------------------------
# workloads (BASELINE.json configs[2], [3], [4])
# ----------------------------------------------------------------------------------------------------------------
CONFIGS = {
    # name: (dna, loss, opt, default batch per GPU, D steps per iteration, description)
    "dna_bce_adam": (True, "bce", "adam", 256, 1,
                     "full adversarial DNA step bce+adam: 1 x train_d + 1 x train_g per step (BASELINE configs[2])"),
    "wass_rmsprop": (True, "wass", "rmsprop", 256, 5,
                     "Wasserstein + RMSProp variant: 5 x train_d + 1 x train_g per step (BASELINE configs[3])"),
    "direct_rollout": (False, "bce", "adam", 512, 0,
                       "non-DNA direct-pixel generator, 6-step recursive rollout (test_sequence) per step "
                       "(BASELINE configs[4]); frames/s counts generated frames"),
}
# nominal 2*M*N*K FLOPs per sample, ksize 6: train_d 1975.3 MF, train_g 2303.0 MF;
# direct generator forward 650.9 MF per rollout step
TRAIN_D_MF, TRAIN_G_MF, DIRECT_FWD_MF = 1975.3, 2303.0, 650.9


def flops_per_iter(config, B):
    dna, loss, opt, _, d_steps, _ = CONFIGS[config]
    if config == "direct_rollout":
        return DIRECT_FWD_MF * 1e6 * 6 * B
    return (d_steps * TRAIN_D_MF + TRAIN_G_MF) * 1e6 * B


def frames_per_iter(config, B):
    return 6 * B if config == "direct_rollout" else B


def synth_sequences(B, seed, pinned, T=13):
    rng = np.random.RandomState(seed)
    base = rng.uniform(-1, 1, (B, 1, 64, 64, 3)).astype(np.float32)
    seq = np.clip(base + np.cumsum(0.05 * rng.randn(B, T, 64, 64, 3).astype(np.float32), axis=1), -1, 1)
    acts = rng.randn(B, T, 10).astype(np.float32)
    ts = [torch.from_numpy(seq), torch.from_numpy(acts)]
    if pinned:
        ts = [t.pin_memory() for t in ts]
    return ts


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference's CPU path, all host threads
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, config="dna_bce_adam", B=CPU_BATCH):
    from oracle import np_ref, torch_ref
    dna, loss, opt, _, d_steps, _ = CONFIGS[config]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rng = np.random.RandomState(7)
    params = np_ref.init_params(np_ref.g_dna_spec(KSIZE) if dna else np_ref.g_direct_spec(), rng)
    params.update(np_ref.init_params(np_ref.d_spec(), rng))
    ora = torch_ref.Trainer(params, True, loss, opt, dna, ksize=KSIZE, dtype=torch.float32)
    img, nxt, act, state = [t.numpy() for t in synth_batch(B, 1, False)]
    seq, acts = [t.numpy() for t in synth_sequences(B, 2, False)]
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        if config == "direct_rollout":
            ora.test_sequence(seq, seq, acts)
        else:
            for _ in range(d_steps):
                ora.train_d(img, nxt, act)
            ora.train_g(img, nxt, act, state)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    what = "6-step rollout" if config == "direct_rollout" else "%d x train_d + train_g" % d_steps
    return {"value": frames_per_iter(config, B) * len(times) / total, "ms_per_step": 1e3 * total / len(times),
            "cores": cores,
            "sample": "%d iteration(s) of %s at batch %d (configs[0] size), fp32 torch-CPU oracle port of "
                      "models.py/ops.py/train.py, %d threads" % (len(times), what, B, cores)}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    res = cpu_reference_run(steps, warmup, args.config)
    line = {
        "impl": "reference", "metric": "GAN train frames/sec", "value": res["value"], "unit": "frames/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": CONFIGS[args.config][5] + ", 64x64x3, ksize=6; CPU sample at batch %d" % CPU_BATCH,
                   "name": args.config, "global_batch": CPU_BATCH},
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
    larger than the 126 MB L2; B=64 rotates 12 buffer sets for the same reason."""
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

        tf, tb = time_kernel(fwd, 24), time_kernel(bwd, 24)
        bytes_f = b * 4096 * (k * k + 6) * 4
        bytes_b = b * 4096 * (2 * k * k + 6) * 4
        out["B%d_K%d" % (b, k)] = {
            "fwd_us": 1e3 * tf, "fwd_gbs": bytes_f / tf / 1e6, "fwd_frac": bytes_f / tf / 1e6 / peaks["hbm_gbs"],
            "bwd_us": 1e3 * tb, "bwd_gbs": bytes_b / tb / 1e6, "bwd_frac": bytes_b / tb / 1e6 / peaks["hbm_gbs"],
            "buffer_sets": nset}
    return out


def load_traffic():
    """DRAM bytes of the conv-family kernels of one iteration, from the committed ncu capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except ValueError:
            return None
    return None


class Workload:
    """One benchmark configuration: device-resident iteration, end-to-end iteration through the public API."""

    def __init__(self, config, B, dev, dp, rank, branches=True):
        from action_conditioned_gans_b200.trainer import Trainer
        self.config, self.B, self.dev = config, B, dev
        dna, loss, opt, _, self.d_steps, _ = CONFIGS[config]
        self.trn = Trainer(None, True, loss, opt, dna, batch_size=B, ksize=KSIZE, device=dev, seed=7, dp=dp,
                           branches=branches)
        self.nfeeds = 2
        if config == "direct_rollout":
            self.host = [synth_sequences(B, 100 + rank * 10 + i, True) for i in range(self.nfeeds)]
            self.resident = [[t.to(dev) for t in hb] for hb in self.host]
        else:
            self.host = [synth_batch(B, 100 + rank * 10 + i, True) for i in range(self.nfeeds)]
            self.resident = [[t.to(dev) for t in hb] for hb in self.host]
            # uint8 copies of the frames (what a decoded dataset holds): a quarter of the H2D bytes
            q = lambda t: ((t + 1.0) * 127.5).round().clamp(0, 255).to(torch.uint8).pin_memory()
            self.host_u8 = [[q(hb[0]), q(hb[1]), hb[2], hb[3]] for hb in self.host]

    def resident_step(self, i):
        trn = self.trn
        if self.config == "direct_rollout":
            seq, acts = self.resident[i % self.nfeeds]
            trn.rollout(seq[:, 0], acts, steps=6, action_stride=2)
            return
        img, nxt, act, state = self.resident[i % self.nfeeds]
        for _ in range(self.d_steps):
            trn.enqueue_train_d(img, nxt, act)
        trn.enqueue_train_g(img, nxt, act, state)

    def e2e_step(self, i, u8):
        trn = self.trn
        if self.config == "direct_rollout":
            seq, acts = self.host[i % self.nfeeds]
            pred, _ = trn.test_sequence(seq, seq, acts)
            return pred
        img, nxt, act, state = (self.host_u8 if u8 else self.host)[i % self.nfeeds]
        for _ in range(self.d_steps):
            trn.train_d(img, nxt, act)
        return trn.train_g(img, nxt, act, state)          # returns the generated frames on the host

    def e2e_bytes(self, u8):
        if self.config == "direct_rollout":
            seq, acts = self.host[0]
            return seq[:, 0].numel() * 4 + acts.numel() * 4, 6 * self.B * 64 * 64 * 3 * 4
        img, nxt, act, state = (self.host_u8 if u8 else self.host)[0]
        per_d = img.numel() * img.element_size() + nxt.numel() * nxt.element_size() + act.numel() * 4
        per_g = per_d + state.numel() * 4
        return self.d_steps * per_d + per_g, self.B * 64 * 64 * 3 * 4


def timed_steps(step_fn, steps, barrier, sampler=None, rank=0):
    """K steps between CUDA events on the launching stream, one event per step boundary: total, median and max."""
    import gc
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    gc.collect()
    gc.disable()          # a collector pause of the enqueueing thread in the first steps (empty launch queue) is a stall
    try:
        barrier()
        evs[0].record()
        for i in range(steps):
            step_fn(i)
            evs[i + 1].record()
            if sampler is not None and rank == 0 and (i == steps // 2 or i == steps - 1):
                sampler.sample_now()      # the device is busy with the queued iterations at this point
        barrier()
    finally:
        gc.enable()
    per = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
    return evs[0].elapsed_time(evs[steps]), per


def main_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the acg_b200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from action_conditioned_gans_b200 import _lib
    from action_conditioned_gans_b200.trainer import DataParallel
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
    config = args.config
    B = args.batch if args.batch > 0 else CONFIGS[config][3]
    wl = Workload(config, B, dev, dp, rank, branches=not args.no_branches)
    trn = wl.trn

    def barrier():
        if dp is not None:
            dp.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if dp is not None:
            dp.dist.all_reduce(t, op=dp.dist.ReduceOp.MAX)
        return float(t.item())

    # warm-up: W >= 3 (first call eager, second captures the graphs, third is the first replay)
    n_warm = max(args.warmup, 3)
    for i in range(n_warm):
        wl.resident_step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count() + trn.replayed_launches
    ms, per = timed_steps(wl.resident_step, args.steps, barrier, sampler, rank)
    launches = _lib.launch_count() + trn.replayed_launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms = max_over_ranks(ms)
    step_med, step_max = max_over_ranks(float(np.median(per))), max_over_ranks(float(np.max(per)))
    step_argmax = int(np.argmax(per))
    fpi = frames_per_iter(config, B)
    value = fpi * world * args.steps / (ms / 1e3)

    # ---- end to end through the public API with pinned HOST feeds (uint8 frames, then fp32 frames) ---------------
    def e2e(u8):
        wl.e2e_step(0, u8)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            out = wl.e2e_step(i, u8)
        barrier()
        return fpi * world * args.steps / max_over_ranks(time.perf_counter() - t0), out

    e2e_u8, frames = e2e(True)
    h2d_u8, d2h = wl.e2e_bytes(True)
    e2e_f32, h2d_f32 = None, None
    if config != "direct_rollout":
        e2e_f32, _ = e2e(False)
        h2d_f32, _ = wl.e2e_bytes(False)

    # ---- strong scaling: the SAME global batch split over the ranks (BASELINE configs[2]: "batch 256 on 1/2/4/8") ---
    strong = None
    if world > 1 and B % world == 0 and config != "direct_rollout":
        ws = Workload(config, B // world, dev, dp, rank, branches=not args.no_branches)
        for i in range(n_warm):
            ws.resident_step(i)
        barrier()
        ms_s, per_s = timed_steps(ws.resident_step, args.steps, barrier)
        ms_s = max_over_ranks(ms_s)
        strong = {"global_batch": B, "batch_per_gpu": B // world, "ms_per_step": ms_s / args.steps,
                  "value": B * args.steps / (ms_s / 1e3), "unit": "frames/s",
                  "step_ms_median": max_over_ranks(float(np.median(per_s)))}
        del ws

    # ---- per-kernel breakdown of one iteration (instrumented pass, right after the timed region; every rank
    # takes part because the step contains collectives) -----------------------------------------------------
    breakdown = kernel_breakdown(wl)
    barrier()
    if rank != 0:
        hard_exit()       # no collective is issued after this point; see hard_exit()
    dna = dna_microbench(dev, peaks)
    flops = flops_per_iter(config, B)
    conv_ms_eager = sum(v["ms"] for k, v in breakdown.items() if k.startswith("acg_conv"))
    conv_ms = conv_only_graph_ms(config, B, dev, rank)
    traffic = load_traffic()
    tens_peak = peaks["bf16_tflops_sustained"]
    achieved = flops / (conv_ms / 1e3) / 1e12
    roofline = {"bound": "tensor",
                "kernel": "conv family (tcgen05 implicit GEMM: conv_halo2 / conv_px / conv_tc / conv_smallk / conv_wgrad_tc)",
                "achieved": achieved, "peak": tens_peak, "unit": "TFLOP/s", "frac": achieved / tens_peak,
                "conv_ms_per_iteration": conv_ms, "conv_ms_per_iteration_eager": conv_ms_eager,
                "achieved_eager": flops / (conv_ms_eager / 1e3) / 1e12,
                "traffic": traffic["conv_dram_bytes_per_iteration"] if traffic and config == "dna_bce_adam" else None,
                "traffic_note": (traffic or {}).get("note"),
                "peak_src": peaks["src"] + " (sustained cuBLAS bf16)",
                "step_frac": flops / (ms / args.steps / 1e3) / 1e12 / tens_peak,
                "note": "achieved = nominal conv FLOPs of one iteration / device time of ALL conv launches of one iteration, "
                        "replayed back to back on ONE stream as a CUDA graph (no other kernel of ours in it, no overlap; "
                        "CUDA events around 10 replays); achieved_eager = the same FLOPs / the sum of per-launch event "
                        "times of an eager pass (adds ~2-4 us of launch gap per launch); step_frac = the same FLOPs / "
                        "the graph-replayed step time"}
    cpu = cpu_reference_run(20 if config != "direct_rollout" else 5, 1, config)
    line = {
        "metric": "GAN train frames/sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": n_warm, "ms_per_step": ms / args.steps, "step_ms_median": step_med, "step_ms_max": step_max,
        "step_ms_argmax_rank0": step_argmax,
        "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": trn.precision, "data": "synthetic",
        "config": {"workload": CONFIGS[config][5] + ", 64x64x3 frames, 10-D action++state, ksize=6, batch %d per GPU" % B,
                   "name": config, "global_batch": B * world, "batch_per_gpu": B, "parallelism": "dp%d" % world,
                   "l2": "activations of one step (>1 GB) exceed the 126 MB L2; %d feed sets rotated" % wl.nfeeds},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": e2e_u8, "unit": "frames/s", "h2d_bytes_per_step": h2d_u8, "d2h_bytes_per_step": d2h,
                "steps": args.steps, "feed_dtype": "uint8 frames (decoded to [-1,1] on the device), fp32 actions",
                "note": "separately timed through the public Trainer calls with pinned HOST feeds; the H2D copies ride a "
                        "copy stream under the previous call's kernels and the frames leave while the backward pass runs"},
        "e2e_fp32_feeds": ({"value": e2e_f32, "unit": "frames/s", "h2d_bytes_per_step": h2d_f32,
                            "d2h_bytes_per_step": d2h} if e2e_f32 is not None else None),
        "strong_scaling": strong,
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


def conv_only_graph_ms(config, B, dev, rank):
    """Device time of the conv launches of ONE iteration: a second trainer (no data parallelism: no collective is
    involved) whose captured step graphs contain only the acg_conv_* launches -- every other entry point is skipped, the
    chains run on one stream -- replayed 10 times between two CUDA events.  Timing of these kernels does not depend on
    the data, which is garbage here."""
    from action_conditioned_gans_b200 import kernels as Kn
    orig = Kn.call

    def conv_only(name, *a):
        if name.startswith("acg_conv_"):
            orig(name, *a)

    Kn.call = conv_only
    try:
        wl = Workload(config, B, dev, None, rank, branches=False)
        for i in range(4):                    # eager, capture, two replays
            wl.resident_step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for i in range(n):
            wl.resident_step(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    finally:
        Kn.call = orig


def kernel_breakdown(wl):
    """Device time per C-ABI entry point over ONE iteration: CUDA events around every call (instrumented pass)."""
    from action_conditioned_gans_b200 import kernels as Kn
    trn = wl.trn
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
        wl.resident_step(0)
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
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=0, help="batch per GPU (default: the configuration's)")
    ap.add_argument("--config", type=str, default="dna_bce_adam", choices=sorted(CONFIGS))
    ap.add_argument("--no-branches", action="store_true", help="every chain of a step on one stream (A/B switch)")
    ap.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    a = ap.parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
