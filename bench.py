#!/usr/bin/env python
"""bench.py — frame-pairs/s for extract+match @ 640x480, 2048 keypoints (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo's CUDA path
    python bench.py --impl reference [...]                       # CPU path (oracle port), host cores

Workload (BASELINE.json configs[1], "c2"): a TUM-RGB-D-shaped synthetic sequence of 600 frames,
640x480 saliency maps + 30x40x384 NHWC feature maps already past the backbone, K = 2048 keypoints,
D = 256, consecutive-pair matching with matcher M1 (ratio 0.8), fp32 mode (descriptors fp32; the
similarity and the refiner GEMMs run as 3-term fp16 hi/lo splits on tcgen05, error < 3e-6 —
`--mode tf32x3` is the TF32 variant, `--mode f32` the CUDA-core exact kernel, `--mode bf16` the bf16
similarity of config c3).  One *step* is one
pass over the whole sequence: every frame is extracted once (decode -> sample -> refiner MLP ->
L2 norm) and each of the 599 consecutive pairs is matched.  With N GPUs every rank processes its
own 600-frame sequence (weak scaling) and the match lists are gathered on rank 0 over NCCL.

`value`  : pairs / s with inputs resident in HBM, timed with CUDA events, max over ranks.
`e2e`    : the same step through FrontEnd.run_sequence_host — inputs start in pinned HOST memory,
           are streamed over PCIe inside the timed region, match lists are copied back.
`roofline`: the dominant kernel of the step (by summed device time; today the tf32x3 GEMM of the
           refiner MLP): algorithmic flops / CUDA-event duration, against the measured bf16 tensor
           peak.  `kernels` lists every kernel kind the same way (HBM-bound ones against the copy peak).
`cpu_baseline`: the oracle port of the same pipeline on the host cores, on a bounded sample.
"""

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "semantic-slam-master_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

_REAL_STDOUT = None
# dram__bytes_read.sum + dram__bytes_write.sum of the refiner GEMM, per activation row, averaged over the
# six launches of one refiner call in the committed `ncu --set full` capture
# (profiles/r1b_all_kernels_full.txt, 614 400 rows: 1.86 / 1.88 / 2.84 / 1.87 / 2.83 / 1.54 GB); the
# algorithmic figure for the same launches is 3 072 B per row (pair in + pair out) and 4 608 B with
# the residual.
NCU_GEMM_DRAM_BYTES_PER_ROW = 3475.0


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


METRIC = "frame-pairs/sec extract+match @640x480, 2048 kpts"
UNIT = "frame-pairs/s"
H, W, C, D = 480, 640, 384, 256


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=600)
    ap.add_argument("--kpts", type=int, default=2048)
    ap.add_argument("--mode", default="auto", choices=["auto", "f32", "tf32x3", "f16x3", "bf16"])
    ap.add_argument("--chunk", type=int, default=300, help="frames per extraction launch group (device-resident arm)")
    ap.add_argument("--e2e-chunk", type=int, default=50, help="frames per host->device staging buffer (e2e arm)")
    ap.add_argument("--cpu-sample-frames", type=int, default=601)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    return ap.parse_args()


def workload_name(a):
    return (f"c2: TUM RGB-D-shaped synthetic sequence, {a.frames} frames {W}x{H}, consecutive-pair "
            f"matching (M1 ratio 0.8), {a.kpts} kpts x {D}-D, refiner MLP 384-384-384-{D}")


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, \
        "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons every 10 ms while the timed region runs (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz, self.ok = None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:                                            # pragma: no cover
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:                                             # pragma: no cover
                pass
            time.sleep(0.01)

    def summary(self):
        if not self.ok or not self.samples:
            return smi_clocks(self.index)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def smi_clocks(index):
    """Fallback when NVML sampling failed: one nvidia-smi query right after the timed region."""
    import subprocess
    try:
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=20).stdout.strip().split(",")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for n, v in zip(names, out[2:]) if v.strip().lower() == "active"]
        return {"sm_mhz": int(out[0]), "sm_max_mhz": int(out[1]), "reasons": reasons,
                "note": "NVML sampling unavailable; single nvidia-smi sample after the timed region"}
    except Exception as e:                                                # pragma: no cover
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": f"no clock source: {e!r}"}


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            return local
    return local


# --------------------------------------------------------------------------------------- CPU arm
def cpu_pipeline_rate(a, frames, workers):
    """Oracle port of the same pipeline on `frames` frames -> (pairs/s, seconds, pairs)."""
    import torch
    import oracle
    from oracle import pipeline as opipe
    from models.descriptor_refiner import DescriptorRefiner
    from sslam_b200 import synth
    torch.manual_seed(0)
    weights = oracle.RefinerWeights.from_state_dict(DescriptorRefiner(C, 384, D, 4).state_dict())
    sal, feat = synth.make_sequence(frames, seq_id=0, height=H, width=W)
    counts, sec = opipe.run_sequence(sal.numpy()[..., 0], feat.numpy(), weights, a.kpts, 1, workers)
    return (frames - 1) / sec, sec, frames - 1


def run_reference(a):
    """--impl reference: the reference is pure Python and cannot travel to the GPU box, so its CPU
    implementation is represented by the oracle port (oracle/pipeline.py) on every host core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pipeline as opipe
    cores = opipe.host_cores()
    frames = max(3, a.cpu_sample_frames)
    vals = []
    for i in range(a.warmup + a.steps):
        v, sec, pairs = cpu_pipeline_rate(a, frames, cores)
        if i >= a.warmup:
            vals.append((v, sec))
    value = sum(v for v, _ in vals) / len(vals)
    ms = 1e3 * sum(s for _, s in vals) / len(vals)
    sample = f"{frames} frames / {frames - 1} pairs of the c2 workload per step, process pool over frames then pairs"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# --------------------------------------------------------------------------------------- GPU arm
def run_b200(a):
    import torch
    import torch.distributed as dist
    from models.descriptor_refiner import DescriptorRefiner
    from sslam_b200 import dist as sdist
    from sslam_b200 import matchers, ops, synth
    from sslam_b200.pipeline import FrontEnd

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    mode_name = a.mode
    if mode_name == "auto":
        # "fp32 mode" of BASELINE config c2: fp32 in/out, 3-term fp16 hi/lo split on the tensor cores
        mode_name = os.environ.get("SSLAM_BENCH_MODE", "f16x3")
    mode = {"f32": ops.SIM_F32, "tf32x3": ops.SIM_TF32X3, "bf16": ops.SIM_BF16, "f16x3": ops.SIM_F16X3}[mode_name]

    torch.manual_seed(0)
    refiner = DescriptorRefiner(C, 384, D, 4).to(dev)
    fe = FrontEnd(refiner, num_keypoints=a.kpts, grid="pixel", sim_mode=mode)
    T = a.frames
    # per-rank sequence, generated on the device (seeded), sigmoid computed once and shared
    sal, feat = synth.make_sequence(T, seq_id=rank, height=H, width=W, device=dev)
    in_bytes = sal.numel() * 4 + feat.numel() * 4

    replay = None
    if not a.no_graph:
        replay, g_feats, g_pairs, g_pscores, g_counts = fe.capture_sequence(sal, feat, matchers.M1, chunk=a.chunk,
                                                                            ratio_thresh=0.8)

    def step(timers=None, eager=False):
        if replay is not None and not eager:
            replay()                                         # one graph launch = the whole step
            pairs, pscores, counts = g_pairs, g_pscores, g_counts
        else:
            feats, pairs, pscores, counts = fe.run_sequence(sal, feat, matchers.M1, chunk=a.chunk,
                                                            timers=timers, ratio_thresh=0.8)
        gathered = None
        if world > 1:
            gathered = sdist.gather_match_lists(pairs, pscores, counts, dst=0)
        return pairs, pscores, counts, gathered

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        step()
    fence()
    sampler = ClockSampler(physical_gpu_index(local))
    sampler.start()
    launches0 = ops.launch_count()
    launches_per_eager_step = None
    if replay is not None:                                   # graph replays bypass the library's counter
        c0 = ops.launch_count()
        step(eager=True)
        fence()
        launches_per_eager_step = ops.launch_count() - c0
        launches0 = ops.launch_count()
    timer_lists = []
    t_beg, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()          # `ncu --profile-from-start off` sees only the timed steps
    t_beg.record()
    for _ in range(a.steps):
        tl = []
        out = step(tl)
        timer_lists.append(tl)
    t_end.record()
    fence()
    torch.cuda.profiler.stop()
    sampler.stop_flag = True
    launches = ops.launch_count() - launches0
    if launches_per_eager_step is not None:
        launches = launches_per_eager_step * a.steps         # kernels inside the replayed graphs
    ms_total = t_beg.elapsed_time(t_end)
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / a.steps
    pairs_per_step = (T - 1) * world
    value = pairs_per_step / (ms_step * 1e-3)

    # instrumented pass: the same K steps again with a CUDA-event pair around every kernel launch of
    # the library (sslam_profile_*), for the per-kernel roofline numbers
    ops.profile_enable(True)
    for _ in range(a.steps):
        step(eager=True)
    fence()
    kernel_ms = ops.profile_read()
    ops.profile_enable(False)

    # per-stage device time from the event marks (same stream as the kernels)
    stage_ms = {}
    if replay is not None:                                   # stage marks need eager launches
        timer_lists = []
        for _ in range(a.steps):
            tl = []
            step(tl, eager=True)
            timer_lists.append(tl)
        fence()
    for tl in timer_lists:
        for (n0, e0), (n1, e1) in zip(tl[:-1], tl[1:]):
            if n1 != "begin":
                stage_ms[n1] = stage_ms.get(n1, 0.0) + e0.elapsed_time(e1)
    stage_ms = {k: v / a.steps for k, v in stage_ms.items()}

    # ---- e2e: host buffers -> match lists on host, through the public pipeline entry point
    e2e = None
    if not a.no_e2e:
        sal_h = torch.empty(sal.shape, dtype=sal.dtype, pin_memory=True).copy_(sal)
        feat_h = torch.empty(feat.shape, dtype=feat.dtype, pin_memory=True).copy_(feat)
        out_h = None
        for _ in range(max(1, min(2, a.warmup))):
            out_h = fe.run_sequence_host(sal_h, feat_h, matchers.M1, chunk=a.e2e_chunk, out_host=out_h,
                                         ratio_thresh=0.8)
        fence()
        w0 = time.perf_counter()
        e_beg, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e_beg.record()
        for _ in range(a.steps):
            out_h = fe.run_sequence_host(sal_h, feat_h, matchers.M1, chunk=a.e2e_chunk, out_host=out_h,
                                         ratio_thresh=0.8)
        e_end.record()
        fence()
        wall_ms = (time.perf_counter() - w0) * 1e3
        e_ms = max(e_beg.elapsed_time(e_end), wall_ms)       # host-blocking D2H: wall clock bounds it
        if world > 1:
            t = torch.tensor([e_ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = float(t.item())
        d2h = sum(t.numel() * t.element_size() for t in out_h)
        e2e = {"value": pairs_per_step / (e_ms / a.steps * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(in_bytes), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": e_ms / a.steps}
        del sal_h, feat_h

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_src = measured_peaks()
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    hbm = float(peaks["hbm_gbs"])
    rows = T * a.kpts
    blocks = 2
    work = {   # algorithmic work per step of each kernel kind
        "gemm_f16x3": ("tensor", 2.0 * rows * (C * 384 + blocks * 2 * 384 * 384 + 384 * D)),
        "match_tc": ("tensor", 2.0 * a.kpts * a.kpts * D * (T - 1)),
        "match_f32": ("tensor", 2.0 * a.kpts * a.kpts * D * (T - 1)),
        "decode_scan": ("hbm", (4.0 * H * W + 12 * a.kpts) * T),
        "gather": ("hbm", (4.0 * (H // 16) * (W // 16) * C + 8 * a.kpts + 4 * a.kpts * C) * T),
        "l2norm": ("hbm", 8.0 * a.kpts * D * T),
        "layernorm": ("hbm", 8.0 * rows * 384 * 2 * blocks),
    }
    kernels = {}
    for kind, (ms_tot, n) in kernel_ms.items():
        e = {"ms_per_step": ms_tot / a.steps, "launches_per_step": n / a.steps}
        if kind in work:
            bound, w = work[kind]
            if bound == "tensor":
                e["algorithmic_TFLOP/s"] = w / (e["ms_per_step"] * 1e-3) / 1e12
                e["frac_of_bf16_peak"] = e["algorithmic_TFLOP/s"] / peak
            else:
                e["algorithmic_GB/s"] = w / (e["ms_per_step"] * 1e-3) / 1e9
                e["frac_of_hbm_peak"] = e["algorithmic_GB/s"] / hbm
        kernels[kind] = e
    dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
    dk = kernels[dom]
    # dram bytes per launch from the committed `ncu --set full` capture (profiles/), where available
    ncu_traffic = {"gemm_f16x3": NCU_GEMM_DRAM_BYTES_PER_ROW * rows / (dk["launches_per_step"] / 6.0)
                   if dom == "gemm_f16x3" else None}
    tf32_note = ("fp32 accuracy costs three 16-bit MMAs per product (fp16 hi/lo split), so the ceiling of "
                 "this fraction is 1/3 = 0.333 (1/6 for the tf32x3 variant)")
    if work.get(dom, ("", 0))[0] == "tensor":
        ach = dk["algorithmic_TFLOP/s"]
        roofline = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                    "frac": ach / peak, "traffic": ncu_traffic.get(dom),
                    "peak_source": peak_src + ", bf16 dense sustained",
                    "launch_ms": dk["ms_per_step"] / dk["launches_per_step"],
                    "launches_per_step": dk["launches_per_step"],
                    "algorithmic_flops_per_step": work[dom][1],
                    "note": ("achieved = algorithmic flops of all launches of this kernel in a step / their "
                             "summed CUDA-event durations (instrumented pass of the same steps); "
                             + (tf32_note if mode_name != "bf16" or dom == "gemm_f16x3" else ""))}
    else:
        ach = dk.get("algorithmic_GB/s", float("nan"))
        roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": hbm, "unit": "GB/s",
                    "frac": ach / hbm, "traffic": None, "peak_source": peak_src,
                    "launch_ms": dk["ms_per_step"] / dk["launches_per_step"]}
    stages = {k: {"ms_per_step": v} for k, v in stage_ms.items()}

    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        from oracle import pipeline as opipe
        cores = opipe.host_cores()
        frames = max(3, a.cpu_sample_frames)
        v, sec, npairs = cpu_pipeline_rate(a, frames, cores)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{frames} frames / {npairs} pairs of the same workload, oracle port "
                         f"(NumPy) in a {cores}-process pool, {sec:.1f} s"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"f32": "f32", "tf32x3": "tf32x3 (fp32 in/out)", "f16x3": "f16x3 (fp32 in/out)", "bf16": "bf16"}[mode_name],
            "data": "synthetic",
            "config": {"workload": workload_name(a), "frames_per_rank": T, "pairs_per_step": pairs_per_step,
                       "similarity_mode": mode_name, "chunk": a.chunk, "e2e_chunk": a.e2e_chunk,
                       "launch": "CUDA graph replay (one graph per step)" if replay is not None else "eager",
                       "l2_policy": f"inputs larger than L2 ({in_bytes / 1e6:.0f} MB per step per GPU)",
                       "parallelism": f"{world} independent sequence shard(s), final NCCL gather of match lists"
                       if world > 1 else "single GPU"},
            "roofline": roofline, "kernels": kernels, "stages": stages, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": sampler.summary()}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    # stdout carries exactly one JSON line: libraries that print there (NCCL's version banner) are
    # sent to stderr, and the line is written to the saved descriptor at the end.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
