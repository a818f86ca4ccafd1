#!/usr/bin/env python
"""bench.py -- mel-frames/s of the EfficientSpeech acoustic forward path on B200.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                      (the CPU arm: the reference's own torch path)

One "step" = one pass of the hot path (Phoneme2Mel.forward, layers/networks.py:415) over one synthetic batch.
Headline workload: BASELINE.json configs[1] -- tiny ES, batch 256 per GPU, 128 phonemes, fp32, TEACHER-FORCED with
fixed geometry (durations injected, all 6 -> T = 768 frames / utterance, caller-supplied max_mel_len), which keeps
the stream free of host syncs so the forward replays as one CUDA graph.  Prints ONE JSON line:

  value         whole-job mel-frames/s, inputs already resident in HBM (device timed, max over ranks)
  e2e           same metric through the public module API with HOST (pinned) inputs and outputs: H2D of the batch and
                D2H of the mel inside the timed region (the closing event waits for the last D2H copy)
  roofline      the dominant kernel (decoder layer / block), CUDA-event timed per launch, against MEASURED_PEAKS.json
  cpu_baseline  the reference's own layers.Phoneme2Mel.forward under torch on the host cores (oracle/_ref staged by
                oracle/build_ref.py; kind "reference"), bounded sample; the numpy port is a secondary key
  configs       sub-records measured in the same run: the other BASELINE.json configs (small B=256, base B=64 per GPU
                = 512 over 8, tiny B=1 latency after demo.py:149-167) and a free-running ragged tiny batch
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "mel_frames_per_s"
UNIT = "frames/s"
HOP, SR = 256, 22050        # config/LJSpeech/preprocess.yaml:16,20 -> mel-RTF
BASELINE_CONFIG = {"tiny": "configs[1]", "small": "configs[2]", "base": "configs[3] (per-GPU shard of 512 over 8)"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--variant", default="tiny", choices=["tiny", "small", "base"])
    ap.add_argument("--batch", type=int, default=None, help="utterances per GPU (default 256; base: 64)")
    ap.add_argument("--phonemes", type=int, default=128)
    ap.add_argument("--duration", type=int, default=6, help="frames per phoneme (teacher-forced)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="skip the sub-records of the other BASELINE configs")
    ap.add_argument("--simt", action="store_true", help="force the fp32 SIMT kernels (no tcgen05)")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--long-steps", type=int, default=200, help="secondary, longer device-resident run")
    a = ap.parse_args()
    if a.batch is None:
        a.batch = 64 if a.variant == "base" else 256
    return a


def workload_config(variant, batch, phonemes, duration, n_gpus):
    T = phonemes * duration
    return {
        "workload": f"{variant} ES batch={batch}/GPU phoneme-len={phonemes} fp32 inference, teacher-forced with fixed "
                    f"geometry (durations all {duration} -> T={T}, max_mel_len given; CUDA-graph replay) "
                    f"(BASELINE.json {BASELINE_CONFIG[variant]})",
        "variant": variant, "batch_per_gpu": batch, "global_batch": batch * n_gpus,
        "phonemes": phonemes, "frames_per_utt": T, "parallelism": f"dp{n_gpus}",
        "l2": "per-step working set (>= 2 x B*T*dx2 fp32 activations + mel, >= 260 MB at B=256) exceeds the 126 MB L2; no flush",
    }


# ------------------------------------------------------------------------------------------
# CPU arm: the reference's own torch path (oracle/_ref) on the host cores; numpy port as a secondary figure
# ------------------------------------------------------------------------------------------
def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def _variant_batch(variant, B, N, duration, seed=0):
    from efficientspeech_b200.config import VARIANTS
    from efficientspeech_b200.params import init_state_dict
    from efficientspeech_b200.synthetic import make_batch
    cfg = VARIANTS[variant]
    sd = init_state_dict(cfg, seed=0)
    batch = make_batch(cfg, B, N, seed=seed, ragged=False, fixed_duration=duration)
    return cfg, sd, batch


def time_reference_torch(variant, B, N, duration, steps, warmup, threads=None):
    """The reference's layers.Phoneme2Mel.forward (layers/networks.py:415), unmodified, teacher-forced under no_grad,
    on `threads` host threads.  Returns None when no reference copy is staged."""
    from oracle import ref_shim
    if not ref_shim.reference_available():
        return None
    import torch
    threads = threads or cpu_threads()
    torch.set_num_threads(threads)
    cfg, sd, batch = _variant_batch(variant, B, N, duration)
    model = ref_shim.build_reference_model(cfg, sd)
    keys = ("phoneme", "phoneme_mask", "pitch", "energy", "duration", "mel_len")
    tb = {k: torch.from_numpy(np.ascontiguousarray(batch[k])) for k in keys}
    frames_per_pass = int(batch["mel_len"].sum())
    with torch.no_grad():
        for _ in range(warmup):
            model(tb, train=True)
        t0 = time.perf_counter()
        for _ in range(steps):
            out = model(tb, train=True)
        dt = time.perf_counter() - t0
    assert tuple(out["mel"].shape) == (B, N * duration, cfg.n_mel)
    return {"value": frames_per_pass * steps / dt, "unit": UNIT, "cores": threads, "kind": "reference",
            "sample": f"{steps} passes (+{warmup} warm-up) of {B} utterances x {N} phonemes x {N * duration} frames through the "
                      f"unmodified reference layers.Phoneme2Mel.forward (torch {torch.__version__} CPU, {threads} threads, "
                      f"teacher-forced, no_grad)",
            "ms_per_pass": dt / steps * 1e3}


def time_reference_b1(variant, N, duration, warm=10, timed=100, threads=None):
    """Single-utterance latency of the reference after demo.py:149-167 (10 warm-up, then timed calls), free-running
    (train=False) with the duration head biased so that T ~= N * duration."""
    from oracle import ref_shim
    if not ref_shim.reference_available():
        return None
    import torch
    threads = threads or cpu_threads()
    torch.set_num_threads(threads)
    cfg, sd, batch = _variant_batch(variant, 1, N, duration)
    sd = dict(sd)
    sd["encoder.duration_decoder.linear.bias"] = np.full_like(sd["encoder.duration_decoder.linear.bias"], float(duration))
    model = ref_shim.build_reference_model(cfg, sd)
    tb = {"phoneme": torch.from_numpy(batch["phoneme"]), "phoneme_mask": torch.from_numpy(batch["phoneme_mask"])}
    ts, frames = [], 0
    with torch.no_grad():
        for i in range(warm + timed):
            t0 = time.perf_counter()
            mel, mel_len, _ = model(tb, train=False)
            dt = time.perf_counter() - t0
            if i >= warm:
                ts.append(dt)
                frames += int(mel_len.sum())
    return {"ms_per_utt_mean": float(np.mean(ts)) * 1e3, "ms_per_utt_median": float(np.median(ts)) * 1e3,
            "value": frames / float(np.sum(ts)), "unit": UNIT, "frames_per_utt": frames / timed, "cores": threads,
            "kind": "reference", "protocol": f"{warm} warm-up + {timed} timed single utterances (demo.py:149-167), free-running"}


def train_batches(cfg, B, N, count, seed0, max_dur=12):
    """LJSpeech-shaped synthetic training batches (BASELINE configs[4]): ragged phoneme lengths in [N/2, N], durations
    U{0..max_dur} (so ~6 frames per phoneme, T ~ 800 at N = 128), pitch / energy targets, a random mel target."""
    from efficientspeech_b200.synthetic import make_batch
    out = []
    for i in range(count):
        b = make_batch(cfg, B, N, seed=seed0 + i, ragged=True, fixed_duration=None, max_dur=max_dur)
        T = int(b["mel_len"].max())
        rng = np.random.default_rng(1000 + seed0 + i)
        b["mel"] = rng.standard_normal((B, T, cfg.n_mel)).astype(np.float32)
        out.append(b)
    return out


def time_reference_train(B, N, steps, warmup=1, threads=None):
    """The reference's training step on the host cores: its modules (train=True), its loss source (model.py:167-209),
    torch autograd and torch.optim.AdamW (model.py:280).  Lightning only drives this loop upstream."""
    from oracle import ref_shim
    if not ref_shim.reference_available() or not os.path.isfile(os.path.join(ref_shim.REF_DIR, "model.py")):
        return None
    import torch
    from efficientspeech_b200.config import VARIANTS
    from efficientspeech_b200.params import init_state_dict
    threads = threads or cpu_threads()
    torch.set_num_threads(threads)
    cfg = VARIANTS["tiny"]
    model = ref_shim.build_reference_model(cfg, init_state_dict(cfg, seed=0)).train()
    fns, ns = ref_shim.reference_functions("model.py", ["loss"])
    ns["nn"] = torch.nn
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-6)
    batches = train_batches(cfg, B, N, 2, 0)
    frames, dt = 0, 0.0
    for i in range(warmup + steps):
        b = batches[i % 2]
        x = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in b.items() if k != "mel"}
        T = int(b["mel_len"].max())
        x["mel_mask"] = torch.from_numpy(np.arange(T)[None, :] >= b["mel_len"][:, None])
        t0 = time.perf_counter()
        opt.zero_grad()
        ls = fns["loss"](None, model(x, train=True), {"mel": torch.from_numpy(b["mel"])}, x)
        (10. * ls[0] + 2. * ls[1] + 2. * ls[2] + ls[3]).backward()
        opt.step()
        if i >= warmup:
            dt += time.perf_counter() - t0
            frames += int(b["mel_len"].sum())
    return {"value": frames / dt, "unit": "mel frames trained/s", "ms_per_step": dt / steps * 1e3, "cores": threads, "kind": "reference",
            "sample": f"{steps} optimisation steps (+{warmup} warm-up) of {B} utterances x <= {N} phonemes (T ~ {T}) through the "
                      f"unmodified reference modules + loss + torch autograd + torch.optim.AdamW (torch {torch.__version__} CPU, {threads} threads)"}


def oracle_pass_parallel(batch, sd, workers):
    """One numpy-oracle forward over `batch`, utterances split across a thread pool."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import es_oracle
    B = batch["phoneme"].shape[0]
    workers = max(1, min(workers, B // 2))
    bounds = np.linspace(0, B, workers + 1).astype(int)
    chunks = [{k: v[bounds[i]:bounds[i + 1]] for k, v in batch.items()} for i in range(workers)]

    def run(c):
        return int(es_oracle.phoneme2mel(c, sd, train=True)["mel_len"].sum())

    with ThreadPoolExecutor(workers) as ex:
        return sum(ex.map(run, chunks))


def time_port(variant, N, duration, utts=64, steps=3, warmup=1):
    cfg, sd, batch = _variant_batch(variant, utts, N, duration)
    threads = cpu_threads()
    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=1)
    except Exception:
        limiter = None
    frames = 0
    for _ in range(warmup):
        oracle_pass_parallel(batch, sd, threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        frames += oracle_pass_parallel(batch, sd, threads)
    dt = time.perf_counter() - t0
    if limiter is not None and hasattr(limiter, "unregister"):
        limiter.unregister()
    return {"value": frames / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{steps} passes of {utts} utterances x {N} phonemes x {N * duration} frames (numpy oracle, {threads} worker threads)"}


def cpu_baseline_record(variant, B, N, duration, steps=5, warmup=2):
    cb = time_reference_torch(variant, B, N, duration, steps, warmup)
    if cb is None:
        cb = time_port(variant, N, duration)
        cb["note"] = "oracle/_ref not staged: numpy port timed instead"
    return cb


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, a.steps), min(a.warmup, 2)
    cb = time_reference_torch(a.variant, a.batch, a.phonemes, a.duration, steps, warmup)
    if cb is None:
        cb = time_port(a.variant, a.phonemes, a.duration, steps=max(1, min(steps, 8)))
        cb["ms_per_pass"] = None
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": cb.get("ms_per_pass"),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(a.variant, a.batch, a.phonemes, a.duration, a.gpus),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "mel_rtf": cb["value"] * HOP / SR}
    if not a.no_sub and cb["kind"] == "reference":
        sub = {}
        if a.variant != "small":
            r = time_reference_torch("small", 256, a.phonemes, a.duration, 2, 1)
            sub["small_b256"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "ms_per_pass")}
        if a.variant != "base":
            r = time_reference_torch("base", 64, a.phonemes, a.duration, 3, 1)
            sub["base_b64_per_gpu"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "ms_per_pass")}
        sub["tiny_b1_latency"] = time_reference_b1("tiny", a.phonemes, a.duration)
        sub["tiny_train_b128_per_gpu"] = time_reference_train(32, a.phonemes, 2)
        line["configs"] = sub
        port = time_port(a.variant, a.phonemes, a.duration)
        line["cpu_port"] = port
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for t, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                c, m = float(parts[0]), float(parts[1])
            except ValueError:
                continue
            mx = m
            if t0 <= t <= t1:
                sm.append(c)
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# host placement: bind the rank to its GPU's NUMA node (pinned buffers are then first-touched there)
# ------------------------------------------------------------------------------------------
def bind_to_gpu_numa(local):
    info = {"bound": False}
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        bus = f"{getattr(p, 'pci_domain_id', 0):04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bus}"
        with open(base + "/numa_node") as f:
            info["gpu_numa_node"] = int(f.read().strip())
        with open(base + "/local_cpulist") as f:
            cpulist = f.read().strip()
        info["gpu_local_cpulist"] = cpulist
        local_cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                lo, hi = part.split("-")
                local_cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                local_cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        info["allowed_cpus"] = len(allowed)
        both = allowed & local_cpus
        if both and both != allowed:
            os.sched_setaffinity(0, both)
            info["bound"] = True
            info["bound_cpus"] = len(both)
        elif both:
            info["note"] = "every allowed CPU is already local to the GPU"
        else:
            info["note"] = "no allowed CPU is local to the GPU (container cpuset); not bound"
    except Exception as e:                                  # best effort: never fail the bench over placement
        info["note"] = f"placement probe failed: {type(e).__name__}: {e}"
    return info


# ------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p.get("bf16_tflops", 1590.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


class Ctx:
    """torch / distributed state shared by the measurements of one run."""

    def __init__(self, a):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != a.gpus and self.world > 1:
            raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={self.world}")
        if a.gpus > 1 and self.world == 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.placement = bind_to_gpu_numa(self.local)
        if self.world > 1:
            # stdout carries exactly one JSON line: NCCL's version banner (NCCL_DEBUG=VERSION, set on some boxes)
            # would precede it, so anything below WARN is raised to WARN unless the caller asked for INFO / TRACE
            if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
                os.environ["NCCL_DEBUG"] = "WARN"
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        t = self.torch.tensor([float(v)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps, lib=None, profile=False, finish=None):
        """K calls of fn bracketed by barrier + synchronize; device time (CUDA events on the current stream), max
        over ranks.  `finish` runs after the last call and before the closing event (joins side streams)."""
        torch = self.torch
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        if profile:
            from efficientspeech_b200 import _cabi
            _cabi.check(lib.es_profile_begin(steps * 64))
        ev0.record()
        for _ in range(steps):
            fn()
        if finish is not None:
            finish()
        ev1.record()
        torch.cuda.synchronize()
        if profile:
            lib.es_profile_end()
        self.barrier()
        return self.max_over_ranks(ev0.elapsed_time(ev1))


def build_model_on(ctx, variant, simt=False, dur_bias=None):
    import efficientspeech_b200 as es
    from efficientspeech_b200.params import init_state_dict
    cfg = es.VARIANTS[variant]
    model = es.build_model(variant)
    if ctx.rank == 0:
        sd = init_state_dict(cfg, seed=0)
        if dur_bias is not None:
            sd = dict(sd)
            sd["encoder.duration_decoder.linear.bias"] = np.full_like(sd["encoder.duration_decoder.linear.bias"], float(dur_bias))
        es.load_numpy_state(model, sd)
    model = model.to(ctx.dev).eval()
    if ctx.world > 1:
        # the one collective of the path: a single NCCL broadcast of the weights over NVLink
        from efficientspeech_b200.sharding import broadcast_weights
        broadcast_weights(model, src=0)
    model.return_features = False          # nothing on the mel path reads the expanded features
    if simt:
        model.set_tensor_core(False)
    return cfg, model


def measure_throughput(ctx, a, variant, B, N, duration, steps, warmup, with_roofline=True, with_e2e=True, long_steps=0):
    """Teacher-forced fixed-geometry throughput of one variant: device-resident `value`, per-kernel roofline, `e2e`."""
    torch = ctx.torch
    from efficientspeech_b200 import _cabi
    from efficientspeech_b200.synthetic import make_batch
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    cfg, model = build_model_on(ctx, variant, simt=a.simt)
    T = N * duration
    batch = make_batch(cfg, B, N, seed=1000 + rank, ragged=False, fixed_duration=duration)
    keys = ("phoneme", "phoneme_mask", "pitch", "energy", "duration", "mel_len")
    host = {k: torch.from_numpy(np.ascontiguousarray(batch[k])).pin_memory() for k in keys}
    x = {k: v.to(dev) for k, v in host.items()}
    x["max_mel_len"] = T                   # known bound: keeps the stream free of host syncs
    lib = _cabi.load()
    frames_per_step = int(batch["mel_len"].sum())

    # The forward is a handful of launches of a few microseconds each: it is captured once into a CUDA graph
    # (the library launches on the capturing stream) and every step replays it.  --no-graph times eager launches.
    graphed = None if a.no_graph else model.capture(x, train=True)

    def step_eager():
        with torch.no_grad():
            return model(x, train=True)["mel"]

    def step_resident():
        if graphed is None:
            return step_eager()
        return graphed()["mel"]

    for _ in range(max(warmup, 3)):
        step_resident()
    w0 = time.perf_counter()
    ms = ctx.timed(step_resident, steps)
    w1 = time.perf_counter()
    rec = {"value": frames_per_step * world * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps,
           "steps": steps, "frames_per_step_per_gpu": frames_per_step}
    if long_steps:
        ms_long = ctx.timed(step_resident, long_steps)
        rec["long_run"] = {"steps": long_steps, "ms_per_step": ms_long / long_steps,
                           "value": frames_per_step * world * long_steps / (ms_long * 1e-3)}
    # kernels per step: counted on one eager pass (a graph replay is one host call, same kernels)
    l0 = lib.es_launch_count()
    step_eager()
    rec["launches_per_step"] = int(lib.es_launch_count() - l0)
    rec["window"] = (w0, w1)

    per_kind = {}
    if with_roofline:
        # second timed region: the same K steps launched eagerly with CUDA events around every kernel (per-kernel
        # durations for the roofline; the events cost host time, so `value` is not taken here).  Each step is preceded
        # by a device-side spin long enough for the host to enqueue the whole step, so the kernels run back to back
        # and an event pair brackets device execution only -- not the host's launch latency.
        spin_cycles = int(2.2e6) if variant == "tiny" else int(4e6)

        def step_eager_queued():
            torch.cuda._sleep(spin_cycles)
            step_eager()

        ctx.timed(step_eager_queued, steps, lib=lib, profile=True)
        cap = steps * 64
        kinds = (ctypes.c_int32 * cap)()
        kms = (ctypes.c_float * cap)()
        n = ctypes.c_int(0)
        _cabi.check(lib.es_profile_collect(kinds, kms, cap, ctypes.byref(n)))
        for i in range(n.value):
            per_kind.setdefault(_cabi.KERNEL_KINDS[kinds[i]], []).append(kms[i])
        rec["kernel_ms_per_step"] = {k: float(np.sum(v)) / steps for k, v in per_kind.items()}
        rec["ms_kernels_per_step_profiled"] = float(sum(np.sum(v) for v in per_kind.values())) / steps
        rec["roofline"] = roofline_record(cfg, model, a, B, N, T, steps, per_kind, variant)

    if with_e2e:
        # e2e: host (pinned) batch -> device, forward, mel -> host (pinned), all inside the timed region.
        # Two graphs (two sets of static buffers) alternate so that the D2H of step i runs on a copy stream while
        # step i+1 computes; step i+2 waits for that copy before it overwrites the mel.  The closing event is recorded
        # after the compute stream has joined the copy stream: every D2H copy is inside the timed region.
        mel_hosts = [torch.empty(B, T, cfg.n_mel, dtype=torch.float32).pin_memory() for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=dev)
        graphs = [graphed, None if a.no_graph else model.capture(x, train=True)]
        copy_done = [torch.cuda.Event(), torch.cuda.Event()]
        count = [0]

        def step_e2e():
            k = count[0] & 1
            count[0] += 1
            main = torch.cuda.current_stream(dev)
            main.wait_event(copy_done[k])                 # the mel buffer of graph k has been drained
            if graphs[k] is None:
                xd = {kk: host[kk].to(dev, non_blocking=True) for kk in keys}
                xd["max_mel_len"] = T
                with torch.no_grad():
                    mel = model(xd, train=True)["mel"]
            else:
                mel = graphs[k]({kk: host[kk] for kk in keys})["mel"]     # H2D into the graph's static inputs
            ready = torch.cuda.Event()
            ready.record(main)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ready)
                mel_hosts[k].copy_(mel, non_blocking=True)
                mel.record_stream(copy_stream)
                copy_done[k].record(copy_stream)

        def join_copies():
            torch.cuda.current_stream(dev).wait_stream(copy_stream)

        for _ in range(3):
            step_e2e()
        join_copies()
        ms_e2e = ctx.timed(step_e2e, steps, finish=join_copies)
        # opt-in variant of the same loop: the mel leaves the device as fp16 (es_mel_to_half; NOT the reference's output
        # precision, reported beside the fp32 figure, never instead of it) -- half the PCIe bytes
        from efficientspeech_b200 import mel_to_half
        half_dev = [torch.empty(B, T, cfg.n_mel, dtype=torch.float16, device=dev) for _ in range(2)]
        half_host = [torch.empty(B, T, cfg.n_mel, dtype=torch.float16).pin_memory() for _ in range(2)]

        def step_e2e_half():
            k = count[0] & 1
            count[0] += 1
            main = torch.cuda.current_stream(dev)
            main.wait_event(copy_done[k])
            if graphs[k] is None:
                xd = {kk: host[kk].to(dev, non_blocking=True) for kk in keys}
                xd["max_mel_len"] = T
                with torch.no_grad():
                    mel = model(xd, train=True)["mel"]
            else:
                mel = graphs[k]({kk: host[kk] for kk in keys})["mel"]
            mel_to_half(mel, out=half_dev[k])
            ready = torch.cuda.Event()
            ready.record(main)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ready)
                half_host[k].copy_(half_dev[k], non_blocking=True)
                copy_done[k].record(copy_stream)

        for _ in range(3):
            step_e2e_half()
        join_copies()
        ms_half = ctx.timed(step_e2e_half, steps, finish=join_copies)
        rec["e2e_fp16_d2h"] = {"value": frames_per_step * world * steps / (ms_half * 1e-3), "unit": UNIT,
                               "d2h_bytes_per_step": int(half_host[0].numel() * 2), "ms_per_step": ms_half / steps,
                               "note": "opt-in: mel cast to fp16 on the device before the copy (lossy, ~1e-3 relative)"}
        rec["e2e"] = {"value": frames_per_step * world * steps / (ms_e2e * 1e-3), "unit": UNIT,
                      "h2d_bytes_per_step": int(sum(host[k].numel() * host[k].element_size() for k in keys)),
                      "d2h_bytes_per_step": int(mel_hosts[0].numel() * 4), "ms_per_step": ms_e2e / steps,
                      "d2h_copies_in_region": steps,
                      "closing_event": "recorded after the compute stream joined the copy stream"}
    _cabi.check(lib.es_check_async_errors(torch.cuda.current_stream().cuda_stream))
    rec["_step_resident"] = step_resident
    rec["_model"] = model
    return rec


def roofline_record(cfg, model, a, B, N, T, steps, per_kind, variant):
    """Roofline of the dominant kernel (decoder layer or fused decoder block): algorithmic bytes per launch (DESIGN.md
    section 5) / mean CUDA-event launch time, against the measured STREAM peak."""
    hbm_peak, tf_peak, peak_src = measured_peaks()
    L, nb, C = cfg.n_dec_layers, cfg.n_blocks, cfg.dx2
    row_b = B * T * C * 4
    tab_b = (B * N + 1) * C * 4
    gather_mode = int(model.decoder._backend.gather_mode)
    gather_fused = gather_mode == 2 and C == 128 and not a.simt
    dl = per_kind.get("dec_layer", [])
    db = per_kind.get("dec_block", [])
    if db:
        # one launch per decoder block: reads the block input (= skip) once, writes the block output once; the first
        # block reads the per-phoneme table instead; the last block also runs the mel head (writes mel, not the skip)
        per_blk = []
        for blk in range(nb):
            x_b = tab_b if blk == 0 else row_b
            y_b = B * T * cfg.n_mel * 4 if blk == nb - 1 and len(db) == nb * steps and not per_kind.get("mel") else row_b
            per_blk.append(x_b + y_b)
        bytes_per_launch = float(np.mean(per_blk))
        flops_per_launch = B * T * cfg.block_depth * (2 * C * C + 2 * cfg.decoder_kernel_size * C)
        times, kname = db, "decoder block (block_depth x [dwconv k5 + 1x1 GEMM + bias + tanh + LN] + skip + LN), one launch per block"
    elif dl:
        bytes_per_launch = row_b * (2 * L + nb) / L
        if gather_fused:
            # the first block reads its input rows (first layer) and its skip rows (last layer) from the per-phoneme
            # projection table [B*N+1, dx2] instead of [B,T,dx2] tensors: those two reads are the table, once each
            bytes_per_launch = (row_b * (2 * L + nb - 2) + 2 * tab_b) / L
        flops_per_launch = B * T * (2 * C * C + 2 * cfg.decoder_kernel_size * C)
        times, kname = dl, "decoder layer (dwconv k5 + 1x1 GEMM + bias + tanh + LN [+ skip + LN])" + \
            ("; first block gathers input / skip rows from the per-phoneme projection table" if gather_fused else "")
    else:
        return None
    avg_ms = float(np.mean(times))
    achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        key = f"{variant}_b{B}_n{N}_t{T}_{'block' if db else 'layer'}"
        if key in tj:
            traffic = float(tj[key]["dram_bytes_per_launch"])      # one ncu --set full capture of the same shape
    t_hbm = bytes_per_launch / (hbm_peak * 1e9)
    t_tensor = flops_per_launch / (tf_peak * 1e12)
    roof = {"kernel": kname, "bound": "hbm" if t_hbm >= t_tensor else "tensor",
            "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
            "traffic": traffic, "peak_source": peak_src, "avg_launch_ms": avg_ms, "launches_timed": len(times),
            "algorithmic_bytes_per_launch": bytes_per_launch,
            "tensor_frac_algorithmic": flops_per_launch / (avg_ms * 1e-3) / 1e12 / tf_peak,
            "tensor_frac_issued_3x_split": 3 * flops_per_launch / (avg_ms * 1e-3) / 1e12 / tf_peak,
            "share_of_step": float(np.sum(times)) / float(sum(np.sum(v) for v in per_kind.values())),
            "timed_region": "second pass of the same K steps: eager launches queued behind a device-side spin, "
                            "CUDA events around every kernel on the launching stream"}
    n_pos = nb if db else L
    if len(times) == n_pos * steps:
        per_pos = []
        for l in range(n_pos):
            if db:
                by = (tab_b if l == 0 else row_b) + (B * T * cfg.n_mel * 4 if (l == nb - 1 and not per_kind.get("mel")) else row_b)
                tag = {"block": l}
            else:
                blk, pos = divmod(l, cfg.block_depth)
                last = pos == cfg.block_depth - 1
                x_b = tab_b if (gather_fused and l == 0) else row_b
                skip_b = 0 if not last else (tab_b if (gather_fused and blk == 0) else row_b)
                by = x_b + skip_b + row_b
                tag = {"layer": l, "block_end": bool(last)}
            ms_l = float(np.mean(times[l::n_pos]))
            tag.update({"algorithmic_bytes": by, "ms": ms_l, "frac": by / (ms_l * 1e-3) / 1e9 / hbm_peak})
            per_pos.append(tag)
        roof["per_launch_position"] = per_pos
    return roof


def measure_free_running(ctx, a, variant, B, N, duration, steps):
    """train=False on a ragged batch: predicted durations, one host sync for max(mel_len), eager launches -- the call
    demo.py / model.py:156 makes.  Device timed over `steps` calls (the sync is inside the timed region)."""
    torch = ctx.torch
    from efficientspeech_b200.synthetic import make_batch
    cfg, model = build_model_on(ctx, variant, simt=a.simt, dur_bias=duration)
    batch = make_batch(cfg, B, N, seed=2000 + ctx.rank, ragged=True, fixed_duration=None)
    x = {k: torch.from_numpy(np.ascontiguousarray(batch[k])).to(ctx.dev) for k in ("phoneme", "phoneme_mask")}
    frames = [0]

    def step():
        with torch.no_grad():
            _, mel_len, _ = model(x, train=False)
        frames[0] = mel_len

    for _ in range(3):
        step()
    fps = int(frames[0].sum().item())
    ms = ctx.timed(step, steps)
    return {"value": fps * ctx.world * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps,
            "frames_per_step_per_gpu": fps, "batch_per_gpu": B,
            "workload": f"{variant} B={B}/GPU ragged lengths U{{{N // 2}..{N}}}, train=False (predicted durations ~{duration}/phoneme, "
                        f"host sync for max mel_len, eager launches, padded frames computed)"}


def measure_b1_latency(ctx, a, variant, N, duration, warm=10, timed=100):
    """BASELINE configs[0] shape on the GPU: single utterance, free-running, after demo.py:149-167."""
    torch = ctx.torch
    from efficientspeech_b200.synthetic import make_batch
    cfg, model = build_model_on(ctx, variant, simt=a.simt, dur_bias=duration)
    batch = make_batch(cfg, 1, N, seed=7, ragged=False, fixed_duration=duration)
    x = {k: torch.from_numpy(np.ascontiguousarray(batch[k])).to(ctx.dev) for k in ("phoneme", "phoneme_mask")}
    ts, frames = [], 0
    with torch.no_grad():
        for i in range(warm + timed):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            mel, mel_len, _ = model(x, train=False)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if i >= warm:
                ts.append(dt)
                frames += int(mel.shape[1])
    rec = {"ms_per_utt_mean": float(np.mean(ts)) * 1e3, "ms_per_utt_median": float(np.median(ts)) * 1e3,
           "value": frames / float(np.sum(ts)), "unit": UNIT, "frames_per_utt": frames / timed,
           "protocol": f"{warm} warm-up + {timed} timed single utterances, wall clock with synchronize on both sides "
                       f"(demo.py:149-167), free-running, eager launches"}
    # the same utterance teacher-forced through one CUDA-graph replay (device timed)
    xt = {k: torch.from_numpy(np.ascontiguousarray(v)).to(ctx.dev) for k, v in batch.items() if k != "phoneme_len"}
    xt["max_mel_len"] = N * duration
    g = model.capture(xt, train=True)
    for _ in range(5):
        g()
    ms = ctx.timed(lambda: g(), timed)
    rec["graph_replay_ms_per_utt"] = ms / timed
    return rec


def measure_hifigan(ctx, B, T, steps):
    """The step after the path: HiFi-GAN V2 generator, mel [B,80,T] -> waveform, seeded weights."""
    import importlib.util
    torch = ctx.torch
    import efficientspeech_b200 as es
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "oracle", "make_golden.py"))
    cfgd = {"resblock": "1", "upsample_rates": [8, 8, 2, 2], "upsample_kernel_sizes": [16, 16, 4, 4],
            "upsample_initial_channel": 128, "resblock_kernel_sizes": [3, 7, 11],
            "resblock_dilation_sizes": [[1, 3, 5], [1, 3, 5], [1, 3, 5]]}
    torch.manual_seed(1234)
    G = es.hifigan.Generator(es.hifigan.AttrDict(cfgd)).eval()
    G.remove_weight_norm()
    G = G.to(ctx.dev)
    mel = (torch.randn(B, 80, T, device=ctx.dev) * 1.5 - 4.0)
    with torch.no_grad():
        for _ in range(2):
            wav = G(mel)
        ms = ctx.timed(lambda: G(mel), steps)
    # MACs per output-rate sample: resblocks C^2 * (3+7+11) * 6 per stage, transposed convs Cin*Cout*k/u, conv_pre / post
    macs, C, L = 80 * 128 * 7 * T, 128, T
    for u, k in zip(cfgd["upsample_rates"], cfgd["upsample_kernel_sizes"]):
        L *= u
        macs += C * (C // 2) * (k // u) * L
        C //= 2
        macs += C * C * (3 + 7 + 11) * 6 * L
    macs += C * 7 * L
    flops = 2.0 * macs * B * ctx.world
    sec = ms * 1e-3 / steps
    fp32_peak = 148 * 128 * 2 * 1.965e9
    return {"value": B * ctx.world * wav.shape[-1] / sec, "unit": "samples/s", "ms_per_step": ms / steps, "batch_per_gpu": B,
            "mel_frames_per_utt": T, "audio_rtf": B * ctx.world * wav.shape[-1] / SR / sec,
            "gflop_per_utt": 2.0 * macs / 1e9, "achieved_tflops": flops / sec / 1e12,
            "roofline": {"bound": "fp32 FMA (SIMT kernels)", "achieved": flops / sec / 1e12, "peak": fp32_peak / 1e12,
                         "unit": "TFLOP/s", "frac": flops / sec / fp32_peak,
                         "peak_source": "nominal 148 SMs x 128 FMA lanes x 1.965 GHz (no measured fp32 figure in MEASURED_PEAKS.json)"},
            "workload": f"HiFi-GAN V2 generator (hifigan/models.py:84-127), B={B} utterances of {T} mel frames, fp32, seeded weights"}


def measure_train(ctx, B, N, steps, warmup=3):
    """BASELINE configs[4]: the tiny training step (forward with a tape, the four losses, backward, gradient all-reduce
    across ranks, AdamW), B utterances per GPU, data-parallel.  value = mel frames trained per second over all ranks."""
    torch = ctx.torch
    from efficientspeech_b200 import training
    cfg, model = build_model_on(ctx, "tiny")
    model.train()
    step = training.TrainStep(model, lr=1e-3, weight_decay=1e-6, warmup_steps=50, total_steps=5000, use_graphs=True)
    dev = ctx.dev
    data = []
    for b in train_batches(cfg, B, N, 4, 100 * ctx.rank):
        x = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in b.items() if k != "mel"}
        x["max_mel_len"] = int(b["mel_len"].max())
        data.append((x, {"mel": torch.from_numpy(b["mel"]).to(dev)}, int(b["mel_len"].sum())))
    it = [0]

    def one():
        x, y, _ = data[it[0] % len(data)]
        it[0] += 1
        return step(x, y)

    first = one()
    for _ in range(max(warmup, 2 * len(data)) - 1):     # every batch shape twice: the caching allocator must have seen them all
        one()
    it[0] = 0
    ms = ctx.timed(one, steps)
    last = one()
    # end to end: every step also brings its batch (ids, targets, the mel target) from pinned host memory
    host = [({k: (v.cpu().pin_memory() if torch.is_tensor(v) else v) for k, v in d[0].items()}, d[1]["mel"].cpu().pin_memory())
            for d in data]
    h2d_bytes = sum(sum(v.numel() * v.element_size() for v in hx.values() if torch.is_tensor(v)) + hm.numel() * 4
                    for hx, hm in host) / len(host)

    def one_e2e():
        hx, hm = host[it[0] % len(host)]
        it[0] += 1
        x = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in hx.items()}
        return step(x, {"mel": hm.to(dev, non_blocking=True)})

    it[0] = 0
    for _ in range(2):
        one_e2e()
    it[0] = 0
    ms_e2e = ctx.timed(one_e2e, steps)
    step.use_graphs = False                          # the same steps enqueued kernel by kernel from Python
    for _ in range(2):
        one()
    it[0] = 0
    ms_eager = ctx.timed(one, steps)
    frames = sum(data[i % len(data)][2] for i in range(steps))
    tot = torch.tensor([float(frames)], device=dev)
    if ctx.world > 1:
        ctx.dist.all_reduce(tot)
    n_params = sum(p.numel() for p in model.parameters() if p.requires_grad)
    model.check_async_errors(dev)                    # a tcgen05 GEMM that timed out on an mbarrier raises here
    # data-parallel sanity: after the same averaged gradients every rank must hold bit-identical weights
    chk = step.opt.flat.double().sum().reshape(1)
    lo, hi = chk.clone(), chk.clone()
    if ctx.world > 1:
        ctx.dist.all_reduce(lo, op=ctx.dist.ReduceOp.MIN)
        ctx.dist.all_reduce(hi, op=ctx.dist.ReduceOp.MAX)
    # multiply-accumulates of one forward: every weight matrix times the positions it is applied at (decoder: frames;
    # phoneme side: phonemes, block 1 at half length); a training step is ~3 forwards' worth (forward, dX, dW)
    dec_w = sum(p.numel() for n_, p in model.decoder.named_parameters() if p.dim() >= 2)
    enc_w = sum(p.numel() for n_, p in model.encoder.named_parameters() if p.dim() >= 2 and "embed" not in n_)
    phon = sum(float(d[0]["phoneme_len"].sum()) for d in data) / len(data)
    gflop = 6.0 * (dec_w * frames / steps + enc_w * phon) / 1e9
    rec = {"value": float(tot.item()) / (ms * 1e-3), "unit": "mel frames trained/s", "ms_per_step": ms / steps, "ms_per_step_eager": ms_eager / steps,
           "e2e": {"value": float(tot.item()) / (ms_e2e * 1e-3), "unit": "mel frames trained/s", "ms_per_step": ms_e2e / steps,
                   "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 0,
                   "note": "batch copied from pinned host memory inside the timed region, same stream (not overlapped)"},
           "launch_mode": "CUDA-graph replay of forward + loss + backward (one graph per batch geometry); all-reduce and AdamW eager",
           "batch_per_gpu": B, "global_batch": B * ctx.world, "frames_per_step_per_gpu": frames / steps, "params": n_params,
           "loss_first_step": float(first[0]), "loss_after_timed_steps": float(last[0]),
           "grad_allreduce": "one flat NCCL all-reduce of %d fp32 per step" % n_params if ctx.world > 1 else "none (1 GPU)",
           "gflop_per_step_per_gpu": gflop, "achieved_tflops_per_gpu": gflop / (ms / steps),
           "bound": "passes over the saved activations (unfused operators), not arithmetic: DESIGN.md section 13",
           "replicas_in_sync": bool(float(lo.item()) == float(hi.item())),
           "peak_mem_gib": torch.cuda.max_memory_allocated(dev) / 2 ** 30,
           "workload": f"tiny ES training step, {B} utterances per GPU x <= {N} phonemes (ragged), T ~ {data[0][0]['max_mel_len']} frames, "
                       "fp32, synthetic LJSpeech-shaped batches; generic fp32 kernels + autograd tape (not the fused inference kernels)"}
    del step, model
    return rec


def run_b200_arm(a):
    ctx = Ctx(a)
    torch, world, rank = ctx.torch, ctx.world, ctx.rank
    sampler = ClockSampler(ctx.local) if rank == 0 else None
    main = measure_throughput(ctx, a, a.variant, a.batch, a.phonemes, a.duration, a.steps, a.warmup,
                              long_steps=a.long_steps)
    clocks = None
    if sampler is not None:
        # nvidia-smi samples every ~100 ms and the timed regions are a few ms long, so the clock record is taken over
        # an extra ~0.6 s of the SAME resident step run back to back (outside every timed region)
        c0 = time.perf_counter()
        while time.perf_counter() - c0 < 0.6:
            for _ in range(8):
                main["_step_resident"]()
            torch.cuda.synchronize()
        clocks = sampler.stop(c0 + 0.05, time.perf_counter())
    model = main.pop("_model")
    main.pop("_step_resident")
    main.pop("window")
    gather_mode = int(model.decoder._backend.gather_mode)
    del model
    torch.cuda.empty_cache()

    sub = {}
    if not a.no_sub:
        others = [("small", 256), ("base", 64)]
        for vname, vb in others:
            if vname == a.variant:
                continue
            r = measure_throughput(ctx, a, vname, vb, a.phonemes, a.duration, max(5, a.steps // 2), 3)
            for k in ("_model", "_step_resident", "window"):
                r.pop(k)
            r["config"] = workload_config(vname, vb, a.phonemes, a.duration, world)
            sub[f"{vname}_b{vb}" + ("_per_gpu" if vname == "base" else "")] = r
            torch.cuda.empty_cache()
        sub["tiny_b1_latency"] = measure_b1_latency(ctx, a, "tiny", a.phonemes, a.duration)
        sub["tiny_free_running_ragged"] = measure_free_running(ctx, a, "tiny", 256, a.phonemes, a.duration, max(5, a.steps // 2))
        torch.cuda.empty_cache()
        sub["hifigan_v2_b16"] = measure_hifigan(ctx, 16, a.phonemes * a.duration, 3)
        torch.cuda.empty_cache()
        try:
            sub["tiny_train_b128_per_gpu"] = measure_train(ctx, 128, a.phonemes, 10)
        except Exception as e:                      # a neighbouring row must not take the headline line down with it
            sub["tiny_train_b128_per_gpu"] = {"error": f"{type(e).__name__}: {e}"[:300]}

    if rank != 0:
        if world > 1:
            ctx.dist.destroy_process_group()
        return
    line = {"metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": main["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(a.variant, a.batch, a.phonemes, a.duration, world), "clocks": clocks,
            "e2e": main["e2e"], "e2e_fp16_d2h": main.get("e2e_fp16_d2h"), "gpu_launches": int(main["launches_per_step"] * a.steps), "roofline": main.get("roofline"),
            "mel_rtf": main["value"] * HOP / SR, "kernel_ms_per_step": main.get("kernel_ms_per_step"),
            "long_run": main.get("long_run"),
            "decoder_path": "simt-fp32" if a.simt else "tcgen05-split-fp16", "gather_mode": gather_mode,
            "launch_mode": "eager" if a.no_graph else "cuda-graph replay (one graph per step)",
            "ms_kernels_per_step_profiled": main.get("ms_kernels_per_step_profiled"),
            "host_placement": ctx.placement}
    if sub:
        line["configs"] = sub
    if world == 1 and not a.no_cpu_baseline:
        cb = cpu_baseline_record(a.variant, a.batch, a.phonemes, a.duration)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        port = time_port(a.variant, a.phonemes, a.duration)
        line["cpu_port"] = port
    print(json.dumps(line), flush=True)
    if world > 1:
        ctx.dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_b200_arm(a)


if __name__ == "__main__":
    main()
