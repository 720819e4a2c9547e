#!/usr/bin/env python
"""bench.py -- spectrograms/s of the VirtualRadar forward pass (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of BASELINE config 2
(N=256 NTU-shaped sequences, 3x300x25x2 -> 256x19) per GPU, synthetic input.

  value      sequences/s, whole job, inputs already resident in HBM, one fused kernel launch per
             step, CUDA events on the launching stream, barrier + synchronize on both sides, max
             over ranks.  Steps rotate over a pool of input/output buffers larger than the 126 MB L2.
  e2e        the same metric through the public API with HOST buffers: VirtualRadar.forward_host
             (C ABI vr_forward_host_f32): pinned host input -> H2D -> kernel -> D2H inside the timed region.
  roofline   HBM bound: algorithmic bytes per launch (199456 B/spectrogram x 256) over
             the mean launch duration measured here; peak from MEASURED_PEAKS.json (else fallback).
  cpu_baseline  the oracle port (reference algorithm, CPU PyTorch, all host threads) timed on this
             box's host cores on a bounded sample (rank 0, N=1 only).
  --impl reference   times only that CPU port, in the same JSON shape.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH, T, V, M = 256, 300, 25, 2
N_FFT, HOP = 256, 16
F = T // HOP + 1
BYTES_IN = 4 * 3 * T * V * M                      # 180 000
BYTES_OUT = 4 * N_FFT * F                         # 19 456
BYTES_PER_SPEC = BYTES_IN + BYTES_OUT             # 199 456  
FLOPS_PER_SPEC = 55 * T * 24 * M + F * (5 * N_FFT * 8 + 8 * N_FFT)   # 1 025 472
WAVELENGTH = 5e-4
WORKLOAD = "synthetic NTU-60 batch N=256, C=3, T=300, V=25, M=2 (BASELINE configs[1]), wavelength 5e-4, default 24-bone skeleton"


def synth_batch(n, seed):
    """S1 of SURVEY 8d: seeded iid randn * 0.3 (CPU generator)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, 3, T, V, M, generator=g) * 0.3


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram bytes per launch from the committed ncu --set full capture, if any (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get("dram_bytes_per_launch_n256")
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU with NVML while the timed regions run."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.active = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                     nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
            while not self._stop_evt.is_set():
                if self.active.is_set():
                    self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for bit, name in names.items():
                        if r & bit:
                            self.reasons.add(name)
                time.sleep(self.period)
        except Exception as e:  # NVML missing: report nothing rather than fail the bench
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def stop(self):
        self._stop_evt.set()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            return local
    return local


# ----------------------------------------------------------------------------------------------
# CPU reference arm (oracle port == the reference's algorithm on CPU PyTorch)
# ----------------------------------------------------------------------------------------------
def time_cpu_port(n_seq, repeats, warmup=0):
    import torch
    from oracle import virtual_radar_oracle as vro
    torch.set_num_threads(os.cpu_count())
    o = vro.OracleVirtualRadar(wavelength=WAVELENGTH)
    x = synth_batch(n_seq, 0)
    for _ in range(warmup):
        o(x)
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        o(x)
        times.append(time.perf_counter() - t0)
    return times


def run_reference(args, rank, world):
    """--impl reference: only rank 0 works; each step is a bounded sample of the workload sized so
    that warmup+steps finish in about two minutes (the CPU path does ~70-100 spectrograms/s)."""
    if rank != 0:
        return
    import torch
    total = args.steps + args.warmup
    n_seq = int(max(1, min(BATCH, (120.0 * 72.0) // max(1, total))))
    times = time_cpu_port(n_seq, args.steps, warmup=args.warmup)
    elapsed = sum(times)
    value = n_seq * args.steps / elapsed
    cores = torch.get_num_threads()
    sample = "%d of the %d sequences of the batch per step, %d steps" % (n_seq, BATCH, args.steps)
    line = {"impl": "reference", "metric": "spectrograms/sec (NTU 3x300x25x2)", "value": value,
            "unit": "spectrograms/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample_per_step": n_seq,
                       "note": "reference algorithm on CPU PyTorch (oracle port of layers/virtual_radar.py + nnAudio STFT restatement); rank 0 only"},
            "cpu_baseline": {"value": value, "unit": "spectrograms/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "spectrograms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    ge.build_product()
    from skeleton_action_recognition_b200 import VirtualRadar, _cabi

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    layer = VirtualRadar(wavelength=WAVELENGTH, device=dev).to(dev)
    # pool of distinct batches bigger than L2 (126 MB): 6 x (46.1 + 5.0) MB = 306 MB
    pool = 6
    xs = [synth_batch(BATCH, 1000 * rank + i).to(dev) for i in range(pool)]
    outs = [torch.empty(BATCH, N_FFT, F, device=dev) for _ in range(pool)]
    lib = _cabi.lib()
    import ctypes
    stream = torch.cuda.current_stream(dev)
    s_ptr = ctypes.c_void_p(stream.cuda_stream)
    lam_ptr, loc_ptr = layer.wavelength.data_ptr(), layer.radar_location.data_ptr()
    E = len(layer.src)

    # The steps are independent batches in distinct buffers, so the caller's guarantee VR_FLAG_INPUTS_READY holds: the
    # launches are programmatic dependent launches and a step's READS may begin while the previous step's last CTAs
    # are still running (its writes still wait).  Every step is computed in full; `stream_ordered` below is the same
    # loop without the flag (each step waits for the previous one to finish completely).
    def step(i, flags=_cabi.VR_FLAG_INPUTS_READY):
        # the module's forward minus torch.empty: the same C-ABI call on preallocated buffers
        rc = lib.vr_forward_f32(xs[i % pool].data_ptr(), BATCH, T, V, M, layer._src_c, layer._dst_c, E,
                                lam_ptr, loc_ptr, N_FFT, HOP, flags, outs[i % pool].data_ptr(), s_ptr)
        if rc:
            _cabi.check(rc)

    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()

    # ---- device-resident throughput ---------------------------------------------------------
    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.active.set()
    e0.record(stream)
    for i in range(args.steps):
        step(i)
    e1.record(stream)
    barrier()
    sampler.active.clear()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = ms_total / args.steps
    value = world * BATCH * args.steps / (ms_total * 1e-3)
    # correctness guard: the timed launches produced the same bits as the module's forward
    torch.cuda.synchronize(dev)
    last = (args.steps - 1) % pool
    for i in {0, last, (last + 1) % pool}:
        assert torch.equal(outs[i], layer(xs[i])), "timed path != VirtualRadar.forward"
    # the same loop in plain stream order (no overlap between consecutive steps)
    so_steps = max(3, min(args.steps, 500))
    for i in range(3):
        step(i, 0)
    barrier()
    o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    o0.record(stream)
    for i in range(so_steps):
        step(i, 0)
    o1.record(stream)
    barrier()
    so_ms = max_over_ranks(o0.elapsed_time(o1)) / so_steps

    # ---- end to end through the public API with host buffers --------------------------------
    xh = [synth_batch(BATCH, 2000 * rank + i).pin_memory() for i in range(2)]
    oh = torch.empty(BATCH, N_FFT, F).pin_memory()
    e2e_steps = max(3, min(args.steps, 200))
    for i in range(3):
        layer.forward_host(xh[i % 2], out=oh)
    barrier()
    sampler.active.set()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        layer.forward_host(xh[i % 2], out=oh)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    sampler.active.clear()
    dt = max_over_ranks(dt)
    e2e_value = world * BATCH * e2e_steps / dt
    assert torch.equal(oh, layer(xh[(e2e_steps - 1) % 2].to(dev)).cpu()), "e2e path != VirtualRadar.forward"

    # ---- sustained large batch (same kernel, many jobs per CTA), for the roofline picture ----
    big_n = 16384
    xb = synth_batch(256, 77).to(dev).repeat(big_n // 256, 1, 1, 1, 1)
    ob = torch.empty(big_n, N_FFT, F, device=dev)

    def big_step():
        rc = lib.vr_forward_f32(xb.data_ptr(), big_n, T, V, M, layer._src_c, layer._dst_c, E, lam_ptr, loc_ptr,
                                N_FFT, HOP, 0, ob.data_ptr(), s_ptr)
        if rc:
            _cabi.check(rc)
    for _ in range(3):
        big_step()
    barrier()
    sampler.active.set()
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 40
    b0.record(stream)
    for _ in range(reps):
        big_step()
    b1.record(stream)
    barrier()
    sampler.active.clear()
    big_ms = max_over_ranks(b0.elapsed_time(b1)) / reps
    big_value = world * big_n / (big_ms * 1e-3)
    sampler.stop()

    peak, peak_src = measured_peaks()
    achieved = BYTES_PER_SPEC * BATCH / (ms_per_step * 1e-3) / 1e9
    big_achieved = BYTES_PER_SPEC * big_n / (big_ms * 1e-3) / 1e9
    props = torch.cuda.get_device_properties(dev)
    fp32_peak = props.multi_processor_count * 128 * 2 * 1.965e9 / 1e12
    plan = _cabi.plan(BATCH, T, V, M, layer.src, layer.dst, N_FFT, HOP, props.multi_processor_count)

    # ---- the data loader's up-sampling pre-stage, reported beside the headline ----
    pre_stage = None
    if rank == 0 and world == 1:
        from skeleton_action_recognition_b200 import pad_frames
        pn, pk = props.multi_processor_count, 250               # one (sequence, coordinate) plane per CTA, 3 waves
        px = synth_batch(pn, 99).to(dev)
        po = torch.empty(pn, 3, T * pk, V, M, device=dev)
        for _ in range(2):
            pad_frames(px, pk, out=po)
        torch.cuda.synchronize(dev)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(stream)
        for _ in range(5):
            pad_frames(px, pk, out=po)
        p1.record(stream)
        torch.cuda.synchronize(dev)
        pms = p0.elapsed_time(p1) / 5
        pbytes = po.numel() * 4 + px.numel() * 4
        pre_stage = {"op": "pad_frames (Dataset.pad_frames + cast, utils.py:128-140), num_pad_frames=250, sigma=3",
                     "n": pn, "ms": pms, "value": pn / (pms * 1e-3), "unit": "sequences/s",
                     "hbm_achieved_gbs": pbytes / (pms * 1e-3) / 1e9, "hbm_frac": pbytes / (pms * 1e-3) / 1e9 / peak,
                     "bytes_per_sequence": pbytes // pn}
        if not args.no_cpu_baseline:
            from oracle import pad_frames as opf
            t0 = time.perf_counter()
            opf.dataset_getitem(px[0].cpu().numpy(), pk)
            dt1 = time.perf_counter() - t0
            pre_stage["cpu_baseline"] = {"value": 1.0 / dt1, "unit": "sequences/s", "cores": 1, "kind": "port",
                                         "sample": "one sequence through scipy gaussian_filter1d + interp1d (%.2f s)" % dt1}
        del px, po

    # ---- the rows next to the headline path: fused consumer resize, fused up-sampling ----
    next_rows = None
    if rank == 0 and world == 1:
        def timed(fn, reps):
            for _ in range(2):
                fn()
            torch.cuda.synchronize(dev)
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record(stream)
            for _ in range(reps):
                fn()
            t1.record(stream)
            torch.cuda.synchronize(dev)
            return t0.elapsed_time(t1) / reps
        img_n, S = 4096, 256
        xi = xb[:img_n]
        ms_f = timed(lambda: layer.forward_image(xi, S), 20)
        ms_u = timed(lambda: torch.nn.functional.interpolate(layer(xi).unsqueeze(1), S), 20)
        img_bytes = BYTES_IN + 4 * S * S
        next_rows = {"consumer_stage": {
            "op": "VirtualRadar + unsqueeze + nearest interpolate to 256x256 (models/resnet.py:24-26) in one launch",
            "n": img_n, "ms": ms_f, "value": img_n / (ms_f * 1e-3), "unit": "sequences/s",
            "hbm_achieved_gbs": img_bytes * img_n / (ms_f * 1e-3) / 1e9, "hbm_frac": img_bytes * img_n / (ms_f * 1e-3) / 1e9 / peak,
            "bytes_per_sequence": img_bytes, "two_launch_ms": ms_u, "speedup_vs_two_launches": ms_u / ms_f}}
        from skeleton_action_recognition_b200 import pad_frames as _pf
        un, uk = 2 * props.multi_processor_count, 250
        ux = synth_batch(un, 55).to(dev)
        ubuf = torch.empty(un, 3, T * uk, V, M, device=dev)
        ms_f = timed(lambda: layer.forward_upsampled(ux, uk, 3, image_size=S), 3)
        ms_u = timed(lambda: layer.forward_image(_pf(ux, uk, 3, out=ubuf), S), 3)
        next_rows["upsampled_pipeline"] = {
            "op": "Dataset.pad_frames(250, sigma 3) + cast + VirtualRadar + resize (utils.py:128-140, models/resnet.py:23-26): "
                  "spline solve + one fused launch, the 45 MB/sequence up-sampled batch never exists",
            "n": un, "frames_per_sequence": T * uk, "ms": ms_f, "value": un / (ms_f * 1e-3), "unit": "sequences/s",
            "two_launch_ms": ms_u, "two_launch_value": un / (ms_u * 1e-3), "speedup_vs_two_launches": ms_u / ms_f,
            "fp32_tflops": (55 * T * uk * 24 * M) * un / (ms_f * 1e-3) / 1e12}
        # the same stage end to end from the loader's side: RAW pinned host batch -> H2D -> fused launch -> image on
        # the device (where the classifier consumes it); wall clock, copies inside the timed region
        uh = [synth_batch(un, 56 + i).pin_memory() for i in range(2)]
        def from_host(i):
            return layer.forward_upsampled(uh[i % 2].to(dev, non_blocking=True), uk, 3, image_size=S)
        for i in range(2):
            from_host(i)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for i in range(4):
            img = from_host(i)
        torch.cuda.synchronize(dev)
        dt = (time.perf_counter() - t0) / 4
        next_rows["upsampled_pipeline"]["e2e_from_host"] = {
            "value": un / dt, "unit": "sequences/s", "h2d_bytes_per_step": un * BYTES_IN, "d2h_bytes_per_step": 0,
            "note": "raw (N,3,300,25,2) pinned host batch in, (N,1,256,256) image left on the device for the classifier; "
                    "the reference ships the 250x up-sampled batch (45 MB per sequence) over PCIe instead"}
        del ux, ubuf, uh, img

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        times = time_cpu_port(BATCH, 3)
        cpu_baseline = {"value": BATCH / min(times), "unit": "spectrograms/s", "cores": torch.get_num_threads(),
                        "kind": "port", "sample": "the full N=256 batch once per repeat, best of 3 (%.2f s each)" % min(times)}

    if rank == 0:
        line = {
            "metric": "spectrograms/sec (NTU 3x300x25x2)", "value": value, "unit": "spectrograms/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "sharding": "by sequence, no collective on the path",
                       "l2": "steps rotate over %d input/output buffer pairs (%.0f MB > 126 MB L2)" % (pool, pool * BATCH * BYTES_PER_SPEC / 1e6),
                       "launch": plan,
                       "overlap": "steps are independent batches: programmatic dependent launches with VR_FLAG_INPUTS_READY "
                                  "(a step's reads may start while the previous step's last CTAs finish; writes wait)"},
            "stream_ordered": {"value": world * BATCH / (so_ms * 1e-3), "ms_per_step": so_ms, "steps": so_steps,
                               "hbm_frac": BYTES_PER_SPEC * BATCH / (so_ms * 1e-3) / 1e9 / peak,
                               "note": "same loop without the flag: every step waits for the previous kernel to finish"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": BYTES_PER_SPEC * BATCH,
                         "fp32_tflops": FLOPS_PER_SPEC * BATCH / (ms_per_step * 1e-3) / 1e12,
                         "fp32_frac_of_derived_peak": FLOPS_PER_SPEC * BATCH / (ms_per_step * 1e-3) / 1e12 / fp32_peak},
            "large_batch": {"n_per_gpu": big_n, "value": big_value, "ms_per_launch": big_ms,
                            "hbm_achieved_gbs": big_achieved, "hbm_frac": big_achieved / peak,
                            "note": "same kernel, one launch over 16384 sequences per GPU (2.95 GB in, 0.32 GB out)"},
            "e2e": {"value": e2e_value, "unit": "spectrograms/s", "h2d_bytes_per_step": BATCH * BYTES_IN,
                    "d2h_bytes_per_step": BATCH * BYTES_OUT, "steps": e2e_steps,
                    "api": "VirtualRadar.forward_host (C ABI vr_forward_host_f32), pinned host buffers"},
            "gpu_launches": args.steps,
            "clocks": sampler.summary(),
        }
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        if pre_stage:
            line["pre_stage"] = pre_stage
        if next_rows:
            line["next_rows"] = next_rows
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1 and args.impl == "ours":
        # convenience: relaunch under torchrun
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__),
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup)]
        sys.exit(subprocess.call(cmd))
    # exactly ONE line on stdout: libraries (NCCL's version banner, torch warnings) write to fd 1 too, so
    # everything except the final JSON line is sent to stderr
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)
    real_stdout.flush()


if __name__ == "__main__":
    main()
