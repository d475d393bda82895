#!/usr/bin/env python
"""Headline benchmark: 1-s audio frames/s of level-8 wavelet-packet features (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload coif4|sym5|stft|haar] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic frames (randn * 0.1, 22050 samples).  The
default workload is BASELINE.json configs[1] (level-8 coif4 packets, log scale, power 2, batch 4096 per GPU);
sym5 / stft / haar select configs[0]/[2]/[3]'s transforms at the same batch.  Under torchrun every rank runs the
same batch on its own GPU (weak scaling, no collective on the transform path; the haar workload adds its one
NCCL all-reduce per step) and rank 0 prints ONE JSON line.

Keys beyond the base contract:
  roofline      the fused kernel against the roofline that binds it (FP32 FMA for the packet trees, HBM for
                STFT / Haar), `achieved` = algorithmic flops (or bytes) per launch / CUDA-event launch time;
                the HBM view of the same launch is always given under roofline.hbm.
  e2e           same metric through the host-buffer C-ABI call (pinned host in, pinned host out; H2D and D2H
                inside the timed region).
  cpu_baseline  the reference's CPU path (ptwt if importable, else the ptwt-structured torch restatement in
                oracle/ptwt_like.py) timed on this box's host cores on a bounded sample (N=1, rank 0 only).
`--impl reference` times that CPU path alone under the same metric/config.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

N_SAMPLES = 22050
FP32_NOMINAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12     # 74.45: 148 SMs x 128 FMA lanes x 1965 MHz
HBM_FALLBACK_GBS = 6650.0                                # B200_PROFILING.md fallback

WORKLOADS = {
    #  name      kind       wavelet  level  baseline config it is quoted on
    "coif4": ("packets", "coif4", 8, "configs[1]: level-8 coif4 packet features, batch 4096, 1 B200"),
    "sym5": ("packets", "sym5", 8, "configs[0] transform (level-8 sym5) at batch 4096"),
    "stft": ("stft", None, 0, "configs[2]: STFT power spectrogram 511/220, batch 4096"),
    "haar": ("haar", "haar", 14, "configs[3]: Haar level-14 mean|c| fingerprint, 4096 clips per step"),
    "rfft": ("rfft", None, 0, "mean-spectrum fingerprint (fingerprints.py:37-62): clip-sum pass, 4096 clips per step"),
}


# ------------------------------------------------------------------------------------------ helpers
def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": HBM_FALLBACK_GBS}, "fallback"


def algorithmic_work(kind, wavelet, level, sign_channel=False):
    """(flops, hbm_bytes) per frame -- SURVEY.md section 8d / DESIGN.md: direct-form filter bank, 2 flop per FMA,
    only the input frame and the final feature tensor touch HBM."""
    if kind == "packets":
        from oracle.filters import DEC_LO
        F = len(DEC_LO[wavelet])
        n, coeffs = N_SAMPLES, 0
        for l in range(1, level + 1):
            n = (n + F - 1) // 2
            coeffs += (1 << l) * n
        out = (1 << level) * n * 4 * (2 if sign_channel else 1)
        return 2.0 * F * coeffs, N_SAMPLES * 4 + out
    if kind == "stft":
        frames, bins = 1 + N_SAMPLES // 220, 256
        # FFT-class count: 5 n log2 n per complex 512-point-equivalent transform per STFT frame
        return frames * 5.0 * 511 * 9, N_SAMPLES * 4 + frames * bins * 4
    if kind == "haar":
        n, coeffs = N_SAMPLES, 0
        for l in range(1, level + 1):
            n = (n + 1) // 2
            coeffs += (1 << l) * n
        return 2.0 * 2 * coeffs, N_SAMPLES * 4
    if kind == "rfft":
        return float(N_SAMPLES), N_SAMPLES * 4          # one add per sample; the clips are read once
    raise ValueError(kind)


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, device):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            try:     # CUDA_VISIBLE_DEVICES may renumber: resolve through the UUID
                self.h = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + str(torch.cuda.get_device_properties(device).uuid))
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device.index or 0)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.004)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------ CPU reference path
def cpu_reference_step_fn(kind, wavelet, level):
    """Returns (fn(x_cpu) -> features, description, kind) for the reference's CPU implementation of the path."""
    from oracle import ptwt_like

    if kind == "packets":
        if ptwt_like.have_ptwt():
            return (lambda x: ptwt_like.ptwt_packets_forward(x, wavelet, level, True, 2.0),
                    "ptwt.WaveletPacket + stack + log (reference wavelet_math.py:167-220)", "reference")
        return (lambda x: ptwt_like.packets_forward(x, wavelet, level, log_scale=True, power=2.0)[0],
                "oracle/ptwt_like.py: per-node F.pad(reflect)+conv1d(stride 2), 256-leaf loop, stack, log "
                "(ptwt not installable here)", "port")
    if kind == "stft":
        return (lambda x: ptwt_like.stft_layer_forward(x, 511, 220, 2.0, True),
                "torchaudio.transforms.Spectrogram(511, 220, power 2) + log (reference wavelet_math.py:47-66)",
                "reference")
    if kind == "haar":
        return (lambda x: ptwt_like.haar_fingerprint(x, level),
                "oracle/ptwt_like.py Haar level-14 tree + mean|c| (pywt not installable here)", "port")
    raise ValueError(kind)


def time_cpu_reference(kind, wavelet, level, budget_s, frames_per_call):
    torch.set_num_threads(os.cpu_count() or 1)
    fn, desc, rkind = cpu_reference_step_fn(kind, wavelet, level)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(frames_per_call, 1, N_SAMPLES, generator=g) * 0.1
    with torch.no_grad():
        fn(x)                                   # warm-up
        t0 = time.perf_counter()
        calls = 0
        while True:
            fn(x)
            calls += 1
            el = time.perf_counter() - t0
            if (el >= budget_s and calls >= 2) or calls >= 1000:
                break
    return {"value": calls * frames_per_call / el, "unit": "frames/s", "cores": torch.get_num_threads(),
            "kind": rkind, "sample": f"{calls} calls x {frames_per_call} frames of the same workload in {el:.1f} s; {desc}"}


def run_reference_arm(args, kind, wavelet, level, cfg):
    rank, _, world = dist_env()
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    fn, desc, rkind = cpu_reference_step_fn(kind, wavelet, level)
    g = torch.Generator().manual_seed(0)
    probe = torch.randn(8, 1, N_SAMPLES, generator=g) * 0.1
    with torch.no_grad():
        fn(probe)
        t0 = time.perf_counter()
        fn(probe)
        per_frame = (time.perf_counter() - t0) / 8
    total_steps = args.steps + args.warmup
    frames = int(max(4, min(256, 150.0 / max(per_frame * total_steps, 1e-9))))    # whole run <= ~2.5 min
    x = torch.randn(frames, 1, N_SAMPLES, generator=g) * 0.1
    with torch.no_grad():
        for _ in range(args.warmup):
            fn(x)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn(x)
        el = time.perf_counter() - t0
    value = args.steps * frames / el
    sample = f"{frames} frames per step of the same workload; {desc}"
    line = {
        "impl": "reference", "metric": cfg["metric"], "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg["config"], "gpu_launches": 0,
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": rkind,
                         "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="coif4", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=4096, help="frames per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for the cpu_baseline leg")
    ap.add_argument("--e2e-chunk", type=int, default=256, help="frames per pipelined chunk of the host-buffer call")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    kind, wavelet, level, quoted = WORKLOADS[args.workload]
    rank, local_rank, world = dist_env()
    flops, hbm_bytes = algorithmic_work(kind, wavelet, level)
    cfg = {
        "metric": "wpt_level8_frames_per_sec" if kind == "packets" else f"{args.workload}_frames_per_sec",
        "config": {"workload": f"{args.workload}: {quoted}", "transform": kind, "wavelet": wavelet, "level": level,
                   "frame_samples": N_SAMPLES, "batch_per_gpu": args.batch, "global_batch": args.batch * world,
                   "power": 2.0, "log_scale": kind != "haar", "parallelism": f"dp{world} (frames sharded, no collective)"
                   if kind != "haar" else f"dp{world} + one NCCL all-reduce of 16385 fp64 at the end of the timed job",
                   "l2_policy": "inputs (361 MB) and outputs (>=420 MB) per step exceed the 126 MB L2; no flush needed"},
    }
    if args.impl == "reference":
        run_reference_arm(args, kind, wavelet, level, cfg)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    import audiodeepfake_detection_b200 as afd
    from audiodeepfake_detection_b200 import _lib
    from audiodeepfake_detection_b200.wavelets import Wavelet
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner to stdout on first use: route fd 1 to stderr until the communicator exists,
        # so that stdout carries exactly the one JSON line
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    lib = _lib.load()

    B = args.batch
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.randn(B, 1, N_SAMPLES, device=dev, generator=g) * 0.1
    wav = Wavelet(wavelet) if wavelet else None

    launches_per_step = 1
    if kind == "packets":
        mod = afd.Packets(wavelet_str=wavelet, max_lev=level, log_scale=True, power=2.0)
        step = lambda: mod(x)[0]                                                    # noqa: E731
    elif kind == "stft":
        mod = afd.STFTLayer(n_fft=511, hop_length=220, log_scale=True, power=2.0)
        step = lambda: mod(x)[0]                                                    # noqa: E731
    elif kind == "rfft":
        acc = afd.SpectrumFingerprintAccumulator(N_SAMPLES, dev)
        launches_per_step = 2
        args.no_e2e = True          # no host-buffer entry point for this pass
        args.no_cpu_baseline = True
        step = lambda: acc.update(x)                                                # noqa: E731
    else:
        acc = afd.FingerprintAccumulator(level, dev)
        launches_per_step = 2
        step = lambda: acc.update(x)                                                # noqa: E731

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # live FP32-FMA peak (same constant-bank-operand FFMA form the filter-bank kernels use)
    tfl = ctypes.c_double(0.0)
    _lib.check("afd_measure_fp32_fma_tflops",
               lib.afd_measure_fp32_fma_tflops(20, ctypes.byref(tfl), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    fma_peak_live = tfl.value

    for _ in range(args.warmup):
        out = step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(dev) as clk:
        e0.record()
        for _ in range(args.steps):
            out = step()
        if kind == "haar" and world > 1:
            acc.all_reduce()     # configs[3]: the shards' partial sums meet in ONE all-reduce at the end of the job
        e1.record()
        torch.cuda.synchronize()
    barrier()
    ms_local = e0.elapsed_time(e1)
    t = torch.tensor([ms_local], dtype=torch.float64, device=dev)
    per_rank = [t.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, t)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_rank = [float(v.item()) / args.steps for v in per_rank]
    ms_step = ms_total / args.steps
    value = B * world * args.steps / (ms_total * 1e-3)
    del out

    # ---- end to end through the host-buffer C-ABI call (pinned host in / out, copies inside the timed region)
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(3, min(args.steps, 8))
        xh = torch.empty(B, N_SAMPLES, dtype=torch.float32).pin_memory()
        xh.copy_(x[:, 0].cpu())
        if kind == "packets":
            T = afd.wpt_out_len(N_SAMPLES, len(wav.dec_lo), level)
            oh = torch.empty(B, 1, T, 1 << level, dtype=torch.float32).pin_memory()
            taps = (ctypes.c_double * len(wav.dec_lo))(*wav.dec_lo)

            def host_step():
                _lib.check("afd_wpt_forward_host", lib.afd_wpt_forward_host(
                    ctypes.c_void_p(xh.data_ptr()), B, N_SAMPLES, N_SAMPLES, taps, len(wav.dec_lo), level, 0, 2.0, 1,
                    1e-12, 0, ctypes.c_void_p(oh.data_ptr()), None, local_rank, args.e2e_chunk))
            d2h = oh.numel() * 4
        elif kind == "stft":
            frames, bins = afd.stft_out_shape(N_SAMPLES, 511, 220)
            oh = torch.empty(B, 1, frames, bins, dtype=torch.float32).pin_memory()

            def host_step():
                _lib.check("afd_stft_power_host", lib.afd_stft_power_host(
                    ctypes.c_void_p(xh.data_ptr()), B, N_SAMPLES, N_SAMPLES, 511, 220, 2.0, 1, 1e-12,
                    ctypes.c_void_p(oh.data_ptr()), local_rank, args.e2e_chunk))
            d2h = oh.numel() * 4
        else:
            sums = torch.zeros(1 << level, dtype=torch.float64)
            cnt = ctypes.c_int64(0)

            def host_step():
                _lib.check("afd_haar_fingerprint_host", lib.afd_haar_fingerprint_host(
                    ctypes.c_void_p(xh.data_ptr()), B, N_SAMPLES, N_SAMPLES, level, ctypes.c_void_p(sums.data_ptr()),
                    ctypes.byref(cnt), local_rank, args.e2e_chunk))
            d2h = sums.numel() * 8 + 8
        for _ in range(2):
            host_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            host_step()
        el = time.perf_counter() - t0
        barrier()
        t = torch.tensor([el], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": B * world * e2e_steps / float(t.item()), "unit": "frames/s",
               "h2d_bytes_per_step": B * N_SAMPLES * 4, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
               "api": "afd_%s_host (C ABI, pinned host buffers, %d-frame chunks on 4 streams)" %
                      ({"packets": "wpt_forward", "stft": "stft_power", "haar": "haar_fingerprint"}[kind], args.e2e_chunk)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_src = measured_peaks()
    kernel_ms = ms_local / args.steps            # one fused kernel per step on the timed stream
    ach_tflops = flops * B / (kernel_ms * 1e-3) / 1e12
    ach_gbs = hbm_bytes * B / (kernel_ms * 1e-3) / 1e9
    hbm_view = {"achieved": ach_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach_gbs / peaks["hbm_gbs"],
                "peak_source": f"MEASURED_PEAKS.json ({peak_src})", "bytes_per_frame": hbm_bytes}
    fma_bound = kind == "packets"
    if fma_bound:
        roofline = {"bound": "fp32_fma", "achieved": ach_tflops, "peak": fma_peak_live, "unit": "TFLOP/s",
                    "frac": ach_tflops / fma_peak_live, "peak_source": "live FFMA probe (afd_measure_fp32_fma_tflops)",
                    "peak_nominal": FP32_NOMINAL_TFLOPS, "frac_of_nominal": ach_tflops / FP32_NOMINAL_TFLOPS,
                    "flops_per_frame": flops, "traffic": None, "hbm": hbm_view}
    else:
        roofline = dict(hbm_view, bound="hbm", traffic=None,
                        fp32_fma={"achieved": ach_tflops, "peak": fma_peak_live, "unit": "TFLOP/s",
                                  "frac": ach_tflops / fma_peak_live, "flops_per_frame": flops})
    traffic_path = os.path.join(ROOT, "profiles", f"traffic_{args.workload}.json")
    if os.path.exists(traffic_path):
        with open(traffic_path) as fh:
            tr = json.load(fh)
        roofline["traffic"] = tr.get("dram_bytes_per_launch")
        roofline["traffic_source"] = tr.get("source")

    line = {
        "metric": cfg["metric"], "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg["config"],
        "clocks": clk.summary(), "gpu_launches": args.steps * launches_per_step, "roofline": roofline,
        "ms_per_step_by_rank": ms_per_rank,
    }
    if e2e is not None:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu_baseline:
        per_call = {"packets": 128, "stft": 256, "haar": 8}[kind]
        line["cpu_baseline"] = time_cpu_reference(kind, wavelet, level, args.cpu_budget, per_call)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
