#!/usr/bin/env python
"""bench.py -- mel-frames/s of the EfficientSpeech acoustic forward path on B200.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                      (the CPU arm: oracle port on the host cores)

One "step" = one pass of the hot path (Phoneme2Mel.forward, layers/networks.py:415) over one
synthetic batch: BASELINE.json configs[1] -- tiny ES, batch 256, 128 phonemes, fp32, durations
injected (teacher-forced, all 6 -> T = 768 frames / utterance).  Prints ONE JSON line:

  value      whole-job mel-frames/s, inputs already resident in HBM (device timed, max over ranks)
  e2e        same metric through the public module API with HOST (pinned) inputs and outputs:
             H2D of the batch and D2H of the mel inside the timed region
  roofline   the dominant kernel (decoder layer), CUDA-event timed per launch inside the timed
             region, against MEASURED_PEAKS.json
  cpu_baseline  the numpy oracle on the host cores over a bounded sample (rank 0, N = 1 only)
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--variant", default="tiny", choices=["tiny", "small", "base"])
    ap.add_argument("--batch", type=int, default=256, help="utterances per GPU")
    ap.add_argument("--phonemes", type=int, default=128)
    ap.add_argument("--duration", type=int, default=6, help="frames per phoneme (teacher-forced)")
    ap.add_argument("--cpu-utts", type=int, default=64, help="utterances per CPU-baseline pass")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--simt", action="store_true", help="force the fp32 SIMT decoder (no tcgen05)")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    return ap.parse_args()


def workload_config(a, n_gpus):
    T = a.phonemes * a.duration
    return {
        "workload": f"{a.variant} ES batch={a.batch} phoneme-len={a.phonemes} fp32 inference, "
                    f"teacher-forced durations all {a.duration} -> T={T} (BASELINE.json configs[1])",
        "variant": a.variant, "batch_per_gpu": a.batch, "global_batch": a.batch * n_gpus,
        "phonemes": a.phonemes, "frames_per_utt": T, "parallelism": f"dp{n_gpus}",
        "l2": "per-step working set (3 x B*T*dx2 fp32 activations, >= 300 MB) exceeds the 126 MB L2; no flush",
    }


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores (test infrastructure used as the reported baseline)
# ------------------------------------------------------------------------------------------
def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def oracle_pass_parallel(batch, sd, workers):
    """One oracle forward over `batch`, utterances split across a thread pool (numpy releases the
    GIL in BLAS and ufuncs; BLAS itself is pinned to 1 thread per worker)."""
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


def time_cpu(a, steps, warmup):
    from efficientspeech_b200.config import VARIANTS
    from efficientspeech_b200.params import init_state_dict
    from efficientspeech_b200.synthetic import make_batch
    cfg = VARIANTS[a.variant]
    sd = init_state_dict(cfg, seed=0)
    batch = make_batch(cfg, a.cpu_utts, a.phonemes, seed=0, ragged=False, fixed_duration=a.duration)
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
    if limiter is not None:
        limiter.unregister() if hasattr(limiter, "unregister") else None
    return {"value": frames / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{steps} passes of {a.cpu_utts} utterances x {a.phonemes} phonemes x "
                      f"{a.phonemes * a.duration} frames (numpy oracle, {threads} worker threads)",
            "ms_per_pass": dt / steps * 1e3}


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, a.steps)
    cb = time_cpu(a, steps, min(a.warmup, 2))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus,
            "steps": steps, "warmup": min(a.warmup, 2), "ms_per_step": cb["ms_per_pass"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(a, a.gpus),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "mel_rtf": cb["value"] * HOP / SR}
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
# B200 arm
# ------------------------------------------------------------------------------------------
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p.get("bf16_tflops", 1590.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


def run_b200_arm(a):
    import torch
    import torch.distributed as dist
    import efficientspeech_b200 as es
    from efficientspeech_b200 import _cabi
    from efficientspeech_b200.params import init_state_dict
    from efficientspeech_b200.synthetic import make_batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}")
    if a.gpus > 1 and world == 1:
        raise SystemExit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's version banner (NCCL_DEBUG=VERSION, set on some boxes)
        # would precede it, so anything below WARN is raised to WARN unless the caller asked for INFO / TRACE
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    cfg = es.VARIANTS[a.variant]
    model = es.build_model(a.variant)
    if rank == 0:
        es.load_numpy_state(model, init_state_dict(cfg, seed=0))
    model = model.to(dev).eval()
    if world > 1:
        # the one collective of the path: a single NCCL broadcast of the weights over NVLink
        from efficientspeech_b200.sharding import broadcast_weights
        broadcast_weights(model, src=0)
    model.return_features = False          # nothing on the mel path reads the expanded features
    if a.simt:
        model.set_tensor_core(False)

    B, N = a.batch, a.phonemes
    T = N * a.duration
    batch = make_batch(cfg, B, N, seed=1000 + rank, ragged=False, fixed_duration=a.duration)
    keys = ("phoneme", "phoneme_mask", "pitch", "energy", "duration", "mel_len")
    host = {k: torch.from_numpy(np.ascontiguousarray(batch[k])).pin_memory() for k in keys}
    x = {k: v.to(dev) for k, v in host.items()}
    x["max_mel_len"] = T                   # known bound: keeps the stream free of host syncs
    lib = _cabi.load()
    frames_per_step = int(batch["mel_len"].sum())

    # The forward is ~30 launches of a few microseconds each: it is captured once into a CUDA graph
    # (the library launches on the capturing stream) and every step replays it.  --no-graph times
    # the eager launches instead.
    graphed = None if a.no_graph else model.capture(x, train=True)

    def step_eager():
        with torch.no_grad():
            return model(x, train=True)["mel"]

    def step_resident():
        if graphed is None:
            return step_eager()
        return graphed()["mel"]

    mel_host = torch.empty(B, T, cfg.n_mel, dtype=torch.float32).pin_memory()
    h2d_bytes = sum(host[k].numel() * host[k].element_size() for k in keys)
    d2h_bytes = mel_host.numel() * 4

    # e2e: host (pinned) batch -> device, forward, mel -> host (pinned), all inside the timed region.
    # Two graphs (two sets of static buffers) alternate so that the 63 MB D2H of step i runs on a
    # copy stream while step i+1 computes; step i+2 waits for that copy before it overwrites the mel.
    mel_hosts = [mel_host, torch.empty_like(mel_host).pin_memory()]
    copy_stream = torch.cuda.Stream(device=dev)
    graphs = [graphed, None if a.no_graph else model.capture(x, train=True)]
    copy_done = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_count = [0]

    def step_e2e():
        k = e2e_count[0] & 1
        e2e_count[0] += 1
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, profile=False):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        if profile:
            _cabi.check(lib.es_profile_begin(steps * 64))
        l0 = lib.es_launch_count()
        w0 = time.perf_counter()
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        torch.cuda.synchronize()
        w1 = time.perf_counter()
        if profile:
            lib.es_profile_end()
        launches = lib.es_launch_count() - l0
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), launches, (w0, w1)

    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(max(a.warmup, 3)):
        step_resident()
    ms, _, (w0, w1) = timed(step_resident, a.steps)
    # kernels per step: counted on one eager pass (a graph replay is one host call, same kernels)
    l0 = lib.es_launch_count()
    step_eager()
    launches = (lib.es_launch_count() - l0) * a.steps
    # second timed region: the same K steps launched eagerly with CUDA events around every kernel
    # (per-kernel durations for the roofline; the events cost host time, so `value` is not taken here).
    # Each step is preceded by a device-side spin long enough for the host to enqueue the whole step,
    # so the kernels run back to back and an event pair brackets device execution only -- not the
    # host's launch latency.
    spin_cycles = int(2.2e6)

    def step_eager_queued():
        torch.cuda._sleep(spin_cycles)
        step_eager()

    ms_prof, _, _ = timed(step_eager_queued, a.steps, profile=True)
    # per-kernel records of the timed region
    cap = a.steps * 64
    kinds = (ctypes.c_int32 * cap)()
    kms = (ctypes.c_float * cap)()
    n = ctypes.c_int(0)
    _cabi.check(lib.es_profile_collect(kinds, kms, cap, ctypes.byref(n)))
    per_kind = {}
    for i in range(n.value):
        per_kind.setdefault(_cabi.KERNEL_KINDS[kinds[i]], []).append(kms[i])

    for _ in range(3):
        step_e2e()
    ms_e2e, _, _ = timed(step_e2e, a.steps)
    _cabi.check(lib.es_check_async_errors(torch.cuda.current_stream().cuda_stream))

    clocks = None
    if sampler is not None:
        # nvidia-smi samples every ~100 ms and the timed regions are a few ms long, so the clock
        # record is taken over an extra ~0.6 s of the SAME resident step run back to back
        # (outside every timed region): clocks under this workload's load, throttle reasons incl.
        c0 = time.perf_counter()
        while time.perf_counter() - c0 < 0.6:
            for _ in range(8):
                step_resident()
            torch.cuda.synchronize()
        clocks = sampler.stop(c0 + 0.05, time.perf_counter())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    total_frames = frames_per_step * world * a.steps
    value = total_frames / (ms * 1e-3)
    e2e_value = total_frames / (ms_e2e * 1e-3)
    hbm_peak, tf_peak, peak_src = measured_peaks()
    # ---- roofline of the dominant kernel (decoder layer): algorithmic bytes per launch =
    # read x + write y (+ read skip on block-end layers), fp32 [B*T, dx2]   
    L, nb = cfg.n_dec_layers, cfg.n_blocks
    layer_bytes = B * T * cfg.dx2 * 4 * (2 * L + nb) / L
    layer_flops = B * T * (2 * cfg.dx2 * cfg.dx2 + 2 * cfg.decoder_kernel_size * cfg.dx2)
    dl = per_kind.get("dec_layer", [])
    kname = "decoder layer (dwconv k5 + 1x1 GEMM + bias + tanh + LN [+ skip + LN])"
    gather_mode = int(model.decoder._backend.gather_mode)
    gather_fused = gather_mode == 2 and cfg.dx2 == 128 and not a.simt
    if gather_fused:
        # the first block reads its input rows (first layer) and its skip rows (last layer) from the per-phoneme
        # projection table [B*N+1, dx2] instead of [B,T,dx2] tensors: those two reads are the table, once each
        layer_bytes = (B * T * cfg.dx2 * 4 * (2 * L + nb - 2) + 2 * (B * N + 1) * cfg.dx2 * 4) / L
        kname += "; first block gathers input / skip rows from the per-phoneme projection table"
    roof = None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.isfile(tpath) and a.variant == "tiny" and B == 256 and T == 768 and not a.simt:
        with open(tpath) as f:
            tj = json.load(f)
        if int(tj.get("gather_mode", 0)) == gather_mode:
            traffic = float(tj["dram_bytes_per_launch"])      # one ncu --set full capture (same shape)
    if dl:
        avg_ms = float(np.mean(dl))
        achieved = layer_bytes / (avg_ms * 1e-3) / 1e9
        roof = {"kernel": kname,
                "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "avg_launch_ms": avg_ms, "launches_timed": len(dl),
                "algorithmic_bytes_per_launch": layer_bytes,
                "tensor_frac_algorithmic": layer_flops / (avg_ms * 1e-3) / 1e12 / tf_peak,
                "share_of_step": float(np.sum(dl)) / float(sum(np.sum(v) for v in per_kind.values())),
                "timed_region": "second pass of the same K steps: eager launches queued behind a device-side spin, "
                                "CUDA events around every kernel on the launching stream"}
    if roof is not None and len(dl) == L * a.steps:
        # the same numbers per layer position (launch order within a step): the four launches are different
        # instantiations with different algorithmic bytes, and the average above hides which one is how far off
        row_b = B * T * cfg.dx2 * 4
        tab_b = (B * N + 1) * cfg.dx2 * 4
        per_layer = []
        for l in range(L):
            blk, pos = divmod(l, cfg.block_depth)
            last = pos == cfg.block_depth - 1
            x_b = tab_b if (gather_fused and l == 0) else row_b
            skip_b = 0 if not last else (tab_b if (gather_fused and blk == 0) else row_b)
            by = x_b + skip_b + row_b
            ms_l = float(np.mean(dl[l::L]))
            per_layer.append({"layer": l, "block_end": bool(last), "gathered": bool(gather_fused and blk == 0 and (l == 0 or last)),
                              "algorithmic_bytes": by, "ms": ms_l, "frac": by / (ms_l * 1e-3) / 1e9 / hbm_peak})
        roof["per_layer"] = per_layer
    kernel_ms = {k: float(np.sum(v)) / a.steps for k, v in per_kind.items()}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(a, world), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "ms_per_step": ms_e2e / a.steps},
            "gpu_launches": int(launches), "roofline": roof,
            "mel_rtf": value * HOP / SR, "kernel_ms_per_step": kernel_ms,
            "decoder_path": "simt-fp32" if a.simt else "tcgen05-split-fp16", "gather_mode": gather_mode,
            "launch_mode": "eager" if a.no_graph else "cuda-graph replay (one graph per step)",
            "ms_kernels_per_step_profiled": float(sum(np.sum(v) for v in per_kind.values())) / a.steps}
    if world == 1 and not a.no_cpu_baseline:
        cb = time_cpu(a, steps=8, warmup=1)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_b200_arm(a)


if __name__ == "__main__":
    main()
