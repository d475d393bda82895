#!/usr/bin/env python
"""Headline benchmark: 1-s audio frames/s of level-8 wavelet-packet features (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload coif4|sym5|sym5_b128|stft|haar|rfft]
                    [--impl reference] [--no-workloads] [--frames 1000000]

A "step" is one pass of the hot path over one batch of synthetic frames (randn * 0.1, 22050 samples).  The
headline workload is BASELINE.json configs[1] (level-8 coif4 packets, log scale, power 2, batch 4096 per GPU).
The same process then measures every other workload of the path and reports them under `workloads`:

    sym5        configs[0]'s transform at batch 4096            sym5_b128  configs[0] as specified (batch 128, + DCNN)
    stft        configs[2]: STFT 511/220 power + log, batch 4096
    haar        configs[3]: Haar level-14 mean|c| fingerprint over --frames (1 M) clips, sharded over the ranks,
                streamed in 8192-clip chunks, ONE NCCL all-reduce at the end
    rfft        mean-spectrum fingerprint pass (fingerprints.py:37-62), batch 4096

Under torchrun every rank runs the same per-GPU batch on its own GPU (weak scaling, no collective on the
transform path) and rank 0 prints ONE JSON line.

Keys beyond the base contract:
  roofline      the fused kernel against the roofline that binds it (FP32 FMA for the packet trees, HBM for
                STFT / Haar / rfft).  `achieved` = algorithmic flops (or bytes) per launch / CUDA-event launch time;
                `peak` is a FIXED denominator: the nominal FP32 peak 148 SMs x 128 lanes x 2 x 1.965 GHz = 74.45 TF
                (MEASURED_PEAKS.json has no fp32 figure) or the measured HBM copy bandwidth of MEASURED_PEAKS.json.
                The live FFMA probe is kept as `peak_live` only.
  e2e           same metric through the host-buffer C-ABI call (pinned host in, pinned host out; H2D and D2H
                inside the timed region), K = --steps calls; `copy_ceiling` is the same bytes moved by bare
                concurrent cudaMemcpyAsync on every rank at once (what the host/PCIe side allows at this N).
  cpu_baseline  the reference's CPU path (ptwt if importable, else the ptwt-structured torch restatement in
                oracle/ptwt_like.py) timed on this box's host cores on a bounded sample (N=1, rank 0 only).
  build         sha256 of the CUDA sources compiled into libafd_b200.so and whether it equals the tree's.
`--impl reference` times that CPU path alone under the same metric/config, same batch per step when the
whole run fits ~4 minutes (else the largest batch that does; stated in cpu_baseline.sample).
"""
from __future__ import annotations

import argparse
import ctypes
import importlib.util
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
HAAR_CHUNK = 8192                                        # clips per accumulation launch of the fingerprint job

WORKLOADS = {
    #  name        kind       wavelet  level  batch  baseline config it is quoted on
    "coif4": ("packets", "coif4", 8, 4096, "configs[1]: level-8 coif4 packet features, batch 4096, 1 B200"),
    "sym5": ("packets", "sym5", 8, 4096, "configs[0] transform (level-8 sym5) at batch 4096"),
    "sym5_b128": ("packets", "sym5", 8, 128, "configs[0]: level-8 sym5 packet features, batch 128 (+ DCNN forward)"),
    "stft": ("stft", None, 0, 4096, "configs[2]: STFT power spectrogram 511/220, batch 4096"),
    "haar": ("haar", "haar", 14, HAAR_CHUNK,
             "configs[3]: Haar level-14 mean|c| fingerprint over 1 M clips sharded over the ranks + one all-reduce"),
    "rfft": ("rfft", None, 0, 4096, "mean-spectrum fingerprint (fingerprints.py:37-62): clip-sum pass, 4096 clips per step"),
}
SIDE_WORKLOADS = ["sym5", "sym5_b128", "stft", "haar", "rfft"]


# ------------------------------------------------------------------------------------------ helpers
def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": HBM_FALLBACK_GBS}, "fallback"


def algorithmic_work(kind, n_taps, level, sign_channel=False):
    """(flops, hbm_bytes) per frame -- SURVEY.md section 8d / DESIGN.md: direct-form filter bank, 2 flop per FMA,
    only the input frame and the final feature tensor touch HBM."""
    if kind == "packets":
        n, coeffs = N_SAMPLES, 0
        for l in range(1, level + 1):
            n = (n + n_taps - 1) // 2
            coeffs += (1 << l) * n
        out = (1 << level) * n * 4 * (2 if sign_channel else 1)
        return 2.0 * n_taps * coeffs, N_SAMPLES * 4 + out
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


def n_taps_of(wavelet):
    if not wavelet:
        return 0
    from audiodeepfake_detection_b200.wavelets import Wavelet
    return Wavelet(wavelet).dec_len


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
            self._stop.wait(0.002)

    def __enter__(self):
        if self.nv is not None and not os.environ.get("AFD_BENCH_NO_CLOCKS"):
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


def build_provenance(lib):
    spec = importlib.util.spec_from_file_location(
        "_afd_build", os.path.join(ROOT, "audiodeepfake-detection_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    digest = lib.afd_source_hash().decode()
    return {"source_hash": digest, "matches_tree": digest == mod.source_hash(), "afd_version": lib.afd_version()}


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


def run_reference_arm(args, kind, wavelet, level, batch, cfg):
    rank, _, world = dist_env()
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    fn, desc, rkind = cpu_reference_step_fn(kind, wavelet, level)
    g = torch.Generator().manual_seed(0)
    probe = torch.randn(32, 1, N_SAMPLES, generator=g) * 0.1
    with torch.no_grad():
        fn(probe)
        t0 = time.perf_counter()
        fn(probe)
        per_frame = (time.perf_counter() - t0) / probe.shape[0]
    total_steps = args.steps + args.warmup
    # the arm's own batch per step when the whole run fits the budget, else the largest batch that does
    frames = int(max(4, min(batch, args.reference_budget / max(per_frame * total_steps, 1e-9))))
    x = torch.randn(frames, 1, N_SAMPLES, generator=g) * 0.1
    with torch.no_grad():
        for _ in range(args.warmup):
            fn(x)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn(x)
        el = time.perf_counter() - t0
    value = args.steps * frames / el
    same = frames == batch
    sample = (f"{frames} frames per step ({'the configured batch' if same else f'bounded sample of the {batch}-frame batch'}), "
              f"{args.steps} steps + {args.warmup} warm-up; {desc}")
    line = {
        "impl": "reference", "metric": cfg["metric"], "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg["config"], "gpu_launches": 0, "frames_per_step": frames,
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": rkind,
                         "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
class Bench:
    """Shared state of one bench process: device, distributed group, library, peaks."""

    def __init__(self, args):
        import torch.distributed as dist
        from audiodeepfake_detection_b200 import _lib

        self.args = args
        self.dist = dist
        self.rank, self.local_rank, self.world = dist_env()
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            # NCCL prints its version banner to stdout on first use: route fd 1 to stderr until the communicator
            # exists, so that stdout carries exactly the one JSON line
            sys.stdout.flush()
            saved_fd = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=self.dev)
                dist.barrier()
                torch.cuda.synchronize()
            finally:
                sys.stdout.flush()
                os.dup2(saved_fd, 1)
                os.close(saved_fd)
        self._lib = _lib
        self.lib = _lib.load()
        self.peaks, self.peak_src = measured_peaks()
        tfl = ctypes.c_double(0.0)
        _lib.check("afd_measure_fp32_fma_tflops", self.lib.afd_measure_fp32_fma_tflops(
            20, ctypes.byref(tfl), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        self.fma_peak_live = tfl.value

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, v):
        t = torch.tensor([v], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gather(self, v):
        t = torch.tensor([v], dtype=torch.float64, device=self.dev)
        out = [t.clone() for _ in range(self.world)]
        if self.world > 1:
            self.dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    # ---------------------------------------------------------------- device-resident measurement
    def make_step(self, name):
        """-> (step(i), launches per step, frames per step, state) for one workload; inputs live in HBM."""
        import audiodeepfake_detection_b200 as afd

        kind, wavelet, level, B, _ = WORKLOADS[name]
        if self.args.batch and name == self.args.workload:
            B = self.args.batch
        dev = self.dev
        g = torch.Generator(device=dev).manual_seed(1234 + self.rank)
        # inputs + outputs of one step must exceed the 126 MB L2: small batches rotate through a pool of batches
        per_step_bytes = B * (N_SAMPLES * 4) * 2
        pool = max(1, -(-400_000_000 // per_step_bytes)) if per_step_bytes < 400_000_000 else 1
        xs = [torch.randn(B, 1, N_SAMPLES, device=dev, generator=g) * 0.1 for _ in range(pool)]
        state = {"pool": pool, "B": B, "xs": xs, "keep": [None] * pool}
        if kind == "packets":
            mod = afd.Packets(wavelet_str=wavelet, max_lev=level, log_scale=True, power=2.0)

            def step(i):
                out = mod(xs[i % pool])[0]
                state["keep"][i % pool] = out            # ring of outputs: the allocator cannot hand the same block back
                return out
            return step, 1, B, state
        if kind == "stft":
            mod = afd.STFTLayer(n_fft=511, hop_length=220, log_scale=True, power=2.0)

            def step(i):
                out = mod(xs[i % pool])[0]
                state["keep"][i % pool] = out
                return out
            return step, 1, B, state
        if kind == "rfft":
            acc = afd.SpectrumFingerprintAccumulator(N_SAMPLES, dev)
            return (lambda i: acc.update(xs[i % pool])), 2, B, state
        acc = afd.FingerprintAccumulator(level, dev)
        state["acc"] = acc
        return (lambda i: acc.update(xs[i % pool])), 2, B, state

    def time_steps(self, step, steps, warmup):
        """W warm-up steps, then exactly K steps between CUDA events (barrier + synchronize on both sides);
        returns (local ms, max-over-ranks ms, clocks)."""
        for i in range(warmup):
            step(i)
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(self.dev) as clk:
            e0.record()
            for i in range(steps):
                step(i)
            e1.record()
            torch.cuda.synchronize()
        self.barrier()
        ms_local = e0.elapsed_time(e1)
        return ms_local, self.max_over_ranks(ms_local), clk.summary()

    def roofline(self, name, kind, flops, hbm_bytes, frames, kernel_ms):
        ach_tflops = flops * frames / (kernel_ms * 1e-3) / 1e12
        ach_gbs = hbm_bytes * frames / (kernel_ms * 1e-3) / 1e9
        hbm_peak = self.peaks["hbm_gbs"]
        hbm_view = {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                    "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({self.peak_src})", "bytes_per_frame": hbm_bytes}
        fma_view = {"achieved": ach_tflops, "peak": FP32_NOMINAL_TFLOPS, "unit": "TFLOP/s",
                    "frac": ach_tflops / FP32_NOMINAL_TFLOPS,
                    "peak_source": "nominal FP32: 148 SMs x 128 lanes x 2 flop x 1.965 GHz (no fp32 figure in MEASURED_PEAKS.json)",
                    "peak_live": self.fma_peak_live, "frac_of_live": ach_tflops / self.fma_peak_live,
                    "flops_per_frame": flops}
        if kind == "packets":
            roof = dict(fma_view, bound="fp32_fma", traffic=None, hbm=hbm_view)
        else:
            roof = dict(hbm_view, bound="hbm", traffic=None, fp32_fma=fma_view)
        traffic_path = os.path.join(ROOT, "profiles", f"traffic_{name}.json")
        if os.path.exists(traffic_path):
            with open(traffic_path) as fh:
                tr = json.load(fh)
            roof["traffic"] = tr.get("dram_bytes_per_launch")
            roof["traffic_source"] = tr.get("source")
        return roof

    def measure(self, name, steps, warmup):
        """Device-timed throughput + roofline of one workload (whole job over all ranks)."""
        kind, wavelet, level, _, quoted = WORKLOADS[name]
        if kind == "haar":
            return self.measure_haar_job(name, warmup)
        step, launches, B, state = self.make_step(name)
        # pooled (small-batch) workloads: the warm-up walks the whole ring of output tensors once, so that the timed region
        # re-uses cached blocks instead of calling cudaMalloc (a ~1 ms synchronising call per new block)
        warm = max(warmup, state["pool"] + 2) if state["pool"] > 1 else warmup
        ms_local, ms_max, clocks = self.time_steps(step, steps, warm)
        flops, hbm_bytes = algorithmic_work(kind, n_taps_of(wavelet), level)
        res = {
            "workload": quoted, "value": B * self.world * steps / (ms_max * 1e-3), "unit": "frames/s",
            "batch_per_gpu": B, "steps": steps, "warmup": warm, "ms_per_step": ms_max / steps,
            "ms_per_step_by_rank": [v / steps for v in self.gather(ms_local)],
            "gpu_launches": steps * launches, "clocks": clocks,
            "l2_policy": ("inputs + outputs of one step exceed the 126 MB L2" if state["pool"] == 1 else
                          f"rotating pool of {state['pool']} input batches and output tensors (> 400 MB in flight)"),
            "roofline": self.roofline(name, kind, flops, hbm_bytes, B, ms_local / steps),
        }
        if name == "sym5_b128":
            res.update(self.small_batch_extras(step, state, steps, warmup))
        state.clear()
        torch.cuda.empty_cache()
        return res

    def small_batch_extras(self, step, state, steps, warmup):
        """configs[0] at batch 128: the step is launch-latency bound, so also report (a) the kernel alone replayed from
        a CUDA graph and (b) features -> Normalize -> DCNN forward (random-init DCNN of the sym5 checkpoint's shape)."""
        import audiodeepfake_detection_b200 as afd
        from audiodeepfake_detection_b200.dcnn import DCNN, DCNNConfig

        extra = {}
        pool, B = state["pool"], state["B"]
        try:
            graph = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for i in range(3):
                    step(i)
                with torch.cuda.graph(graph, stream=side):
                    for i in range(pool):
                        step(i)
            torch.cuda.current_stream().wait_stream(side)
            ms_local, ms_max, _ = self.time_steps(lambda i: graph.replay(), max(3, steps // pool), 3)
            per = ms_max / max(3, steps // pool) / pool
            extra["cuda_graph"] = {"ms_per_step": per, "value": B * self.world / (per * 1e-3), "unit": "frames/s",
                                   "note": f"{pool} launches per replay"}
        except Exception as exc:      # noqa: BLE001  (capture is an extra, never the reported value)
            extra["cuda_graph"] = {"error": str(exc)[:200]}
        try:
            torch.manual_seed(0)
            model = DCNN(DCNNConfig(time_len=95, time_dim_add=1)).to(self.dev).eval()
            mod = afd.Packets(wavelet_str="sym5", max_lev=8, log_scale=True, power=2.0)
            mod.fused_norm = afd.wavelet_math.HostNorm(-13.6, 4.9)
            xs = state["xs"]

            def fwd(i):
                with torch.no_grad():
                    return model(mod(xs[i % pool])[0])
            ms_local, ms_max, _ = self.time_steps(fwd, steps, warmup)
            extra["with_dcnn_forward"] = {"ms_per_step": ms_max / steps, "unit": "frames/s",
                                          "value": B * self.world * steps / (ms_max * 1e-3),
                                          "note": "features (fused Normalize) -> DCNN forward, fp32 eager PyTorch/cuDNN, random init"}
        except Exception as exc:      # noqa: BLE001
            extra["with_dcnn_forward"] = {"error": str(exc)[:200]}
        return extra

    def measure_haar_job(self, name, warmup):
        """configs[3]: every rank streams its shard of --frames clips in 8192-clip chunks through the accumulation
        kernel (sums stay on the device), then ONE all-reduce; timed as a whole job (fingerprints.py:85-125)."""
        import audiodeepfake_detection_b200 as afd

        kind, wavelet, level, chunk, quoted = WORKLOADS[name]
        total = int(self.args.frames)
        lo = total * self.rank // self.world
        hi = total * (self.rank + 1) // self.world
        shard = hi - lo
        resident = min(shard, 16 * chunk)              # 131072 clips = 11.6 GB resident; longer shards cycle over it
        g = torch.Generator(device=self.dev).manual_seed(99 + self.rank)
        x = torch.empty(resident, 1, N_SAMPLES, device=self.dev)
        for s in range(0, resident, chunk):
            x[s:s + chunk] = torch.randn(min(chunk, resident - s), 1, N_SAMPLES, device=self.dev, generator=g) * 0.1
        spans = []
        done = 0
        while done < shard:
            n = min(chunk, shard - done)
            off = done % resident
            if off + n > resident:
                n = resident - off
            spans.append((off, n))
            done += n

        def job(acc):
            for off, n in spans:
                acc.update(x[off:off + n])

        for _ in range(max(1, min(warmup, 3))):
            acc = afd.FingerprintAccumulator(level, self.dev)
            job(acc)
            acc.all_reduce()
        acc = afd.FingerprintAccumulator(level, self.dev)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        align = torch.zeros(1, device=self.dev)
        with ClockSampler(self.dev) as clk:
            self.barrier()
            # The ranks leave the host-side barrier milliseconds apart (python, NVML), which a 4 ms job would count as
            # all-reduce time on the early ranks: a one-element all-reduce lines the STREAMS up, e0 follows it on the device.
            if self.world > 1:
                self.dist.all_reduce(align)
            e0.record()
            job(acc)
            e1.record()
            acc.all_reduce()
            e2.record()
            torch.cuda.synchronize()
        self.barrier()
        ms_accum, ms_total = e0.elapsed_time(e1), e0.elapsed_time(e2)
        ms_max = self.max_over_ranks(ms_total)
        count = int(acc.count.item())
        flops, hbm_bytes = algorithmic_work(kind, 2, level)
        res = {
            "workload": quoted, "value": total / (ms_max * 1e-3), "unit": "frames/s", "frames": total,
            "frames_per_rank": shard, "chunk_clips": chunk, "resident_clips_per_rank": resident,
            "ms_job": ms_max, "ms_accumulate_by_rank": self.gather(ms_accum),
            "allreduce_us_by_rank": [v * 1e3 for v in self.gather(ms_total - ms_accum)],
            "allreduce": "one NCCL all-reduce of 16385 fp64 (sums + count), stream-ordered after the last chunk; "
                         "its time on a rank includes waiting for the slowest rank's accumulation",
            "terms_counted": count, "terms_expected": total * 2,
            "gpu_launches": 2 * len(spans), "clocks": clk.summary(),
            "l2_policy": "every chunk (722 MB) exceeds the 126 MB L2",
            "roofline": self.roofline(name, kind, flops, hbm_bytes, shard, ms_accum),
        }
        del x
        torch.cuda.empty_cache()
        return res

    # ---------------------------------------------------------------- end to end through the host-buffer C ABI
    def e2e(self, name, steps):
        import audiodeepfake_detection_b200 as afd
        from audiodeepfake_detection_b200.wavelets import Wavelet

        kind, wavelet, level, B, _ = WORKLOADS[name]
        if self.args.batch and name == self.args.workload:
            B = self.args.batch
        if kind == "rfft":
            return None
        lib, _lib, chunk = self.lib, self._lib, self.args.e2e_chunk
        g = torch.Generator().manual_seed(4321 + self.rank)
        xh = torch.empty(B, N_SAMPLES, dtype=torch.float32).pin_memory()
        xh.copy_(torch.randn(B, N_SAMPLES, generator=g) * 0.1)
        if kind == "packets":
            wav = Wavelet(wavelet)
            T = afd.wpt_out_len(N_SAMPLES, len(wav.dec_lo), level)
            oh = torch.empty(B, 1, T, 1 << level, dtype=torch.float32).pin_memory()
            taps = (ctypes.c_double * len(wav.dec_lo))(*wav.dec_lo)

            def host_step():
                _lib.check("afd_wpt_forward_host", lib.afd_wpt_forward_host(
                    ctypes.c_void_p(xh.data_ptr()), B, N_SAMPLES, N_SAMPLES, taps, len(wav.dec_lo), level, 0, 2.0, 1,
                    1e-12, 0, ctypes.c_void_p(oh.data_ptr()), None, self.local_rank, chunk))
            d2h = oh.numel() * 4
        elif kind == "stft":
            frames, bins = afd.stft_out_shape(N_SAMPLES, 511, 220)
            oh = torch.empty(B, 1, frames, bins, dtype=torch.float32).pin_memory()

            def host_step():
                _lib.check("afd_stft_power_host", lib.afd_stft_power_host(
                    ctypes.c_void_p(xh.data_ptr()), B, N_SAMPLES, N_SAMPLES, 511, 220, 2.0, 1, 1e-12,
                    ctypes.c_void_p(oh.data_ptr()), self.local_rank, chunk))
            d2h = oh.numel() * 4
        else:
            oh = None
            sums = torch.zeros(1 << level, dtype=torch.float64)
            cnt = ctypes.c_int64(0)

            def host_step():
                _lib.check("afd_haar_fingerprint_host", lib.afd_haar_fingerprint_host(
                    ctypes.c_void_p(xh.data_ptr()), B, N_SAMPLES, N_SAMPLES, level, ctypes.c_void_p(sums.data_ptr()),
                    ctypes.byref(cnt), self.local_rank, chunk))
            d2h = sums.numel() * 8 + 8
        for _ in range(2):
            host_step()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            host_step()
        el = time.perf_counter() - t0
        self.barrier()
        el_max = self.max_over_ranks(el)
        h2d = B * N_SAMPLES * 4
        res = {"value": B * self.world * steps / el_max, "unit": "frames/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": steps, "batch_per_gpu": B,
               "api": "afd_%s_host (C ABI, pinned host buffers, %d-frame chunks on 4 streams)" %
                      ({"packets": "wpt_forward", "stft": "stft_power", "haar": "haar_fingerprint"}[kind], chunk)}
        # what the host / PCIe side allows at this rank count: the same bytes as bare concurrent copies on every rank
        d_in = torch.empty(B, N_SAMPLES, dtype=torch.float32, device=self.dev)
        d_out = torch.empty(oh.shape, dtype=torch.float32, device=self.dev) if oh is not None else None
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

        def copies():
            with torch.cuda.stream(s_in):
                d_in.copy_(xh, non_blocking=True)
            if d_out is not None:
                with torch.cuda.stream(s_out):
                    oh.copy_(d_out, non_blocking=True)
        copies()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(max(3, min(steps, 10))):
            copies()
        torch.cuda.synchronize()
        el_copy = self.max_over_ranks(time.perf_counter() - t0) / max(3, min(steps, 10))
        self.barrier()
        ceiling = B * self.world / el_copy
        res["copy_ceiling"] = {
            "value": ceiling, "unit": "frames/s", "h2d_GBs_per_gpu": h2d / el_copy / 1e9,
            "d2h_GBs_per_gpu": d2h / el_copy / 1e9, "host_GBs_all_ranks": (h2d + d2h) * self.world / el_copy / 1e9,
            "how": "the step's H2D and D2H bytes as two bare concurrent cudaMemcpyAsync per rank, all ranks at once"}
        res["frac_of_copy_ceiling"] = res["value"] / ceiling
        res["bound"] = ("host<->device copies (PCIe / host memory at this rank count): "
                        f"{res['value'] / ceiling:.0%} of the bare-copy rate") if res["value"] / ceiling > 0.7 else \
            "host-side pipeline (chunk submission) below the bare-copy rate"
        del xh, oh, d_in, d_out
        torch.cuda.empty_cache()
        return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="coif4", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="frames per GPU per step of --workload (default: its configured batch)")
    ap.add_argument("--frames", type=int, default=1_000_000, help="clips of the Haar fingerprint job (configs[3])")
    ap.add_argument("--no-workloads", action="store_true", help="headline workload only (skip the `workloads` dict)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for the cpu_baseline leg")
    ap.add_argument("--reference-budget", type=float, default=240.0, help="seconds the --impl reference run may take")
    ap.add_argument("--e2e-chunk", type=int, default=256, help="frames per pipelined chunk of the host-buffer call")
    ap.add_argument("--e2e-max-steps", type=int, default=40, help="cap on the e2e leg's K (each step moves ~0.8 GB over PCIe)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    name = args.workload
    kind, wavelet, level, batch, quoted = WORKLOADS[name]
    if args.batch:
        batch = args.batch
    rank, local_rank, world = dist_env()
    cfg = {
        "metric": "wpt_level8_frames_per_sec" if kind == "packets" else f"{name}_frames_per_sec",
        "config": {"workload": f"{name}: {quoted}", "transform": kind, "wavelet": wavelet, "level": level,
                   "frame_samples": N_SAMPLES, "batch_per_gpu": batch, "global_batch": batch * world,
                   "power": 2.0, "log_scale": kind not in ("haar", "rfft"),
                   "parallelism": f"dp{world} (frames sharded, no collective)" if kind != "haar" else
                   f"dp{world} + one NCCL all-reduce of 16385 fp64 at the end of the job",
                   "l2_policy": "inputs and outputs of one step exceed the 126 MB L2 (small batches rotate through a "
                                "pool of batches > 400 MB); no flush needed"},
    }
    if args.impl == "reference":
        run_reference_arm(args, kind, wavelet, level, batch, cfg)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    bench = Bench(args)
    head = bench.measure(name, args.steps, args.warmup)
    e2e = None
    if not args.no_e2e and kind != "rfft":
        e2e = bench.e2e(name, max(3, min(args.steps, args.e2e_max_steps)))
    side = {}
    if not args.no_workloads:
        for other in SIDE_WORKLOADS:
            if other == name:
                continue
            try:
                side[other] = bench.measure(other, args.steps, args.warmup)
                if not args.no_e2e and other in ("sym5", "stft"):
                    side[other]["e2e"] = bench.e2e(other, max(3, min(args.steps, 10)))
            except Exception as exc:      # noqa: BLE001  (a side workload never takes the headline line down)
                side[other] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    if rank != 0:
        if world > 1:
            bench.dist.destroy_process_group()
        return

    line = {
        "metric": cfg["metric"], "value": head["value"], "unit": "frames/s", "n_gpus": world,
        "steps": head.get("steps", args.steps), "warmup": args.warmup,
        "ms_per_step": head.get("ms_per_step", head.get("ms_job")), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg["config"],
        "clocks": head["clocks"], "gpu_launches": head["gpu_launches"], "roofline": head["roofline"],
        "build": build_provenance(bench.lib),
    }
    for key in ("ms_per_step_by_rank", "ms_accumulate_by_rank", "allreduce_us_by_rank", "frames", "frames_per_rank"):
        if key in head:
            line[key] = head[key]
    if e2e is not None:
        line["e2e"] = e2e
    if side:
        line["workloads"] = side
    if world == 1 and not args.no_cpu_baseline and kind != "rfft":
        per_call = {"packets": 256, "stft": 256, "haar": 8}[kind]
        line["cpu_baseline"] = time_cpu_reference(kind, wavelet, level, args.cpu_budget, per_call)
    print(json.dumps(line), flush=True)
    if world > 1:
        bench.dist.destroy_process_group()


if __name__ == "__main__":
    main()
