#!/usr/bin/env python3
"""Headline benchmark: 752x576 frames/sec (and p50 per-frame latency) of the fused pix_shuffle
forward on B200, next to the reference's CPU forward.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision bf16|fp32]

One "step" = one pass of the hot path over one batch of 64 synthetic RGB444 framebuffers (mixed
lores / lores-laced / hires / hires-laced pixel modes) per GPU: uint8 RGBA [64,576,752,4] in ->
gamma -> pix_shuffle-lightweight -> gamma -> uint8 RGBA out (BASELINE.json configs[1]; the deployed
contract of convertion_tools/torch2onnx.py).  `value` is timed with the frames already in HBM;
`e2e` goes through the host-buffer C-ABI call (pinned host memory, H2D + D2H inside the timed region).
N>1 (torchrun): every rank runs the same per-GPU batch on its own GPU, no data-path collective.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAME_H, FRAME_W, BATCH = 576, 752, 64
GFLOP_PER_FRAME = 29.472          # 2 * 136080 MAC/px * 108288 px, no halo/padding
BYTES_PER_FRAME_U8 = 2 * FRAME_H * FRAME_W * 4
WORKLOAD = "pix_shuffle-lightweight, 64 synthetic RGB444 752x576 RGBA framebuffers per GPU (16 each lores/lores_laced/hires/hires_laced), u8 in -> u8 out incl. gamma"


def measured_peaks():
    """(burst TFLOP/s, sustained TFLOP/s, HBM GB/s, source).  MEASURED_PEAKS.json is driver-written: the burst figure is
    cuBLAS bf16 timed alone at full clocks, the sustained one back to back for 4 s (power-capped clocks)."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1631.8), d.get("bf16_tflops_sustained", 1379.1), d.get("hbm_gbs", 6549.1), "measured"
    return 1650.0, 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def committed_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this round
    (profiles/r02_dram_traffic.json, written by tools/ncu_summary.py from the .ncu-rep): {kernel label: bytes}."""
    p = os.path.join(ROOT, "profiles", "r02_dram_traffic.json")
    return json.load(open(p)) if os.path.exists(p) else {}


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md) through NVML
    (same counters as `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*`), every ~2 ms."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        self.index, self.sm, self.mask, self.max_mhz = index, [], 0, None
        self._stop = threading.Event()
        self._thr = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as exc:  # noqa: BLE001
            self.err = repr(exc)
            return
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.002)

    def stop(self):
        if self._thr is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "?")]}
        self._stop.set()
        self._thr.join()
        reasons = sorted(name for bit, name in self.REASONS.items() if self.mask & bit)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(self.sm)}


def seeded_model():
    """Config 1 of SURVEY 8d: torch.manual_seed(S) before constructing get_model('lightweight') (product module, CPU)."""
    import torch
    from fs_uae_image_enhancer_project_b200 import model_pix_shuffle
    torch.manual_seed(7)
    return model_pix_shuffle.get_model("lightweight")


def cpu_forward_fps(n_frames: int, batch: int, sd=None, frames=None, repeats: int = 1):
    """The reference's CPU eval forward (oracle port: same torch ops as model_pix_shuffle.py:227-298 +
    the float glue of train.py:57-73) on the host cores, on a bounded sample of the workload: the same weights and
    the same synthetic frames as the GPU arm.  The only place bench.py touches oracle/."""
    import torch
    from oracle import enhancer_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    spec = O.pix_shuffle_preset("lightweight")
    if sd is None:
        sd = {k: v.detach().clone() for k, v in seeded_model().state_dict().items()}
    fb = frames if frames is not None else torch.from_numpy(O.synth_rgb444_frames(n_frames, FRAME_H, FRAME_W, seed=1000))
    with torch.no_grad():
        O.framebuffer_forward(sd, spec, fb[:1])                       # warm-up
        times = []
        for _ in range(repeats):
            t0 = time.perf_counter()
            for i in range(0, n_frames, batch):
                O.framebuffer_forward(sd, spec, fb[i:i + batch])
            times.append(time.perf_counter() - t0)
    return n_frames / statistics.median(times), torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = 8
    vals = []
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_forward_fps(2, 2)
    t_all = time.perf_counter()
    for _ in range(args.steps):
        fps, threads = cpu_forward_fps(n, 4)
        vals.append(fps)
    fps = statistics.median(vals)
    line = {
        "impl": "reference", "metric": "752x576 frames/sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * n / fps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": f"each step = {n}-frame sample of the workload, batch 4"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": f"{n} frames per step, oracle port of the PyTorch eval forward, fp32"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="auto", choices=["auto", "bf16", "fp16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--chunk", type=int, default=64, help="frames per internal engine pass")
    ap.add_argument("--quick", action="store_true", help="skip latency / e2e legs (tuning runs)")
    ap.add_argument("--stream-frames", type=int, default=4096, help="BASELINE config 5: host frames in the sharded stream (0: skip)")
    ap.add_argument("--sustain", type=float, default=2.5, help="seconds of the sustained device-resident leg (0: skip)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from fs_uae_image_enhancer_project_b200 import _lib, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cpus = []
    if world > 1:      # one process per GPU: keep its pinned host buffers on the GPU's NUMA node
        from fs_uae_image_enhancer_project_b200.sharding import bind_to_gpu_numa_node
        numa_cpus = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    model = seeded_model().to(dev)                       # random-init weights of the named architecture, seeded
    model.chunk_frames = args.chunk
    precision = args.precision
    if precision in ("auto", "bf16"):
        try:
            model.set_precision("bf16")
            model.engine_for(dev, FRAME_H, FRAME_W)
            precision = "bf16"
        except _lib.EngineError as exc:
            if args.precision == "bf16" or exc.code != _lib.ERR_UNSUPPORTED:
                raise
            precision = "fp32"
    if precision in ("fp32", "fp16"):
        model.set_precision(precision)
    eng = model.engine_for(dev, FRAME_H, FRAME_W)

    # two rotating batches: 2 x 64 frames x (1.73 MB in + 1.73 MB out) = 443 MB > 126 MB L2
    # synthetic RGB444 framebuffers are generated on the device (fsuae_synth_rgb444_frames: frame g uses pixel mode g & 3,
    # i.e. 16 frames of each mode per batch); pinned host copies feed the end-to-end measurement
    d_in = [synth.synth_rgb444_frames(BATCH, FRAME_H, FRAME_W, seed=1000 + rank, first_frame=i * BATCH, device=dev) for i in range(2)]
    host = [t.cpu().pin_memory() for t in d_in]
    d_out = [torch.empty_like(t) for t in d_in]
    flags = _lib.FLAG_GAMMA_IN | _lib.FLAG_GAMMA_OUT

    def step(i):
        eng.enqueue(d_in[i & 1], d_out[i & 1], BATCH, _lib.FMT_U8_NHWC4, _lib.FMT_U8_NHWC4, flags)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(args.warmup):
        step(i)
    barrier()
    launches_per_step = eng.last_launch_count
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        step(i)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None

    # the bytes just timed, checked outside the timed region: two frames of the last step's output against the oracle
    parity = None
    if rank == 0:
        from oracle import enhancer_oracle as O                      # checker only (never inside a timed region)
        last = (args.steps - 1) & 1
        sd_cpu = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
        worst, exact = 0, 1.0
        for k in (0, BATCH - 1):
            want = O.framebuffer_forward(sd_cpu, O.pix_shuffle_preset("lightweight"), host[last][k:k + 1])
            d = (d_out[last][k:k + 1].cpu().int() - want.int()).abs()
            worst, exact = max(worst, int(d.max())), min(exact, float((d == 0).float().mean()))
        gate = {"bf16": 8, "fp16": 2, "fp32": 1}[precision]     # bf16: 6 LSB for >= 99.99 % of the values, darkest pixels up to 8 (tests/test_parity_gpu.py)
        parity = {"checked": "frames 0 and 63 of the last timed step vs the CPU oracle (u8 RGBA)", "max_lsb": worst,
                  "exact_frac": round(exact, 5), "gate_lsb": gate, "ok": worst <= gate}

    # per-kernel device times (CUDA events around every launch, on the launching stream), outside the timed region
    eng.set_profiling(True)
    per_kernel = {}
    for i in range(3):
        step(i)
        torch.cuda.synchronize(dev)
        for label, t_ms in eng.kernel_times():
            per_kernel.setdefault(label, []).append(t_ms)
    eng.set_profiling(False)
    per_kernel = {k: sum(v) / len(v) for k, v in per_kernel.items()}

    # the other arithmetic path of the same library, same frames, same weights: one kernel per layer (conv3 -> conv4 fused),
    # feature maps through HBM.  Timed like the headline; reported beside it, never instead of it.
    alt = None
    if eng.last_launch_count <= 2 and not args.quick:
        os.environ["FSUAE_NO_MEGA"] = "1"
        try:
            model2 = seeded_model().to(dev)
            model2.chunk_frames = args.chunk
            model2.set_precision(precision)
            eng2 = model2.engine_for(dev, FRAME_H, FRAME_W)
        finally:
            os.environ.pop("FSUAE_NO_MEGA", None)
        for i in range(args.warmup):
            eng2.enqueue(d_in[i & 1], d_out[i & 1], BATCH, _lib.FMT_U8_NHWC4, _lib.FMT_U8_NHWC4, flags)
        barrier()
        ev0.record()
        for i in range(args.steps):
            eng2.enqueue(d_in[i & 1], d_out[i & 1], BATCH, _lib.FMT_U8_NHWC4, _lib.FMT_U8_NHWC4, flags)
        ev1.record()
        barrier()
        alt_ms = ev0.elapsed_time(ev1)
        alt = {"path": "one kernel per layer (conv3 -> conv4 fused), feature maps through HBM", "variant": eng2.variant,
               "launches_per_step": int(eng2.last_launch_count), "ms_per_step": alt_ms / args.steps}
        del eng2, model2
        torch.cuda.empty_cache()

    if args.quick:
        if rank == 0:
            print(json.dumps({"chunk": args.chunk, "fps": BATCH * world * args.steps / (ms / 1000.0),
                              "us_per_frame": 1e3 * ms / args.steps / BATCH}), flush=True)
        return
    def pctl(v, q):
        v = sorted(v)
        return v[min(len(v) - 1, int(len(v) * q))]

    # single-frame latency, device resident (p50 over 200 launches)
    lat = []
    one_in, one_out = d_in[0][:1].contiguous(), d_out[0][:1].contiguous()
    for i in range(20):
        eng.enqueue(one_in, one_out, 1, _lib.FMT_U8_NHWC4, _lib.FMT_U8_NHWC4, flags)
    torch.cuda.synchronize(dev)
    for i in range(200):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.enqueue(one_in, one_out, 1, _lib.FMT_U8_NHWC4, _lib.FMT_U8_NHWC4, flags)
        b.record()
        b.synchronize()
        lat.append(a.elapsed_time(b))

    # end to end through the host-buffer C-ABI call (pinned host memory; H2D and D2H inside)
    h_out = [torch.empty_like(h).pin_memory() for h in host]
    for i in range(2):
        eng.run_host(host[i & 1], h_out[i & 1], BATCH, _lib.FMT_U8_NHWC4, _lib.FMT_U8_NHWC4, flags)
    # the emulator's own call: ONE host frame in, one host frame out, blocking (README.md:21-24)
    lat1 = []
    for i in range(120):
        t0 = time.perf_counter()
        eng.run_host(host[0][i % BATCH:i % BATCH + 1], h_out[0][i % BATCH:i % BATCH + 1], 1, _lib.FMT_U8_NHWC4, _lib.FMT_U8_NHWC4, flags)
        if i >= 20:
            lat1.append(1e3 * (time.perf_counter() - t0))
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    # (a) one synchronous fsuae_engine_run_host call per step (pipeline fill and drain paid on every call)
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        eng.run_host(host[i & 1], h_out[i & 1], BATCH, _lib.FMT_U8_NHWC4, _lib.FMT_U8_NHWC4, flags)
    torch.cuda.synchronize(dev)
    e2e_sync_s = time.perf_counter() - t0
    barrier()
    # (b) the streaming form of the same call: every step is submitted (upload + forward + download of its 64 host
    # frames), consecutive steps overlap, one wait at the end -- all copies of all steps are inside the timed region
    sampler_e2e = ClockSampler(local)        # the legs before this one have warmed the GPU up: is the SM clock still at its maximum?
    if rank == 0:
        sampler_e2e.start()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        eng.submit_host(host[i & 1], h_out[i & 1], BATCH, _lib.FMT_U8_NHWC4, _lib.FMT_U8_NHWC4, flags)
    eng.wait_host()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    e2e_clocks = sampler_e2e.stop() if rank == 0 else None
    barrier()
    # (c) what the box's host<->device path can carry at this rank count: the same bytes, same 16-frame pieces, H2D and D2H
    # on two streams at once, all ranks concurrently, no kernel in between -- the ceiling of any end-to-end number here
    s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    piece = 16
    def copy_pass(steps):
        for i in range(steps):
            for f0 in range(0, BATCH, piece):
                with torch.cuda.stream(s_up):
                    d_in[i & 1][f0:f0 + piece].copy_(host[i & 1][f0:f0 + piece], non_blocking=True)
                with torch.cuda.stream(s_dn):
                    h_out[i & 1][f0:f0 + piece].copy_(d_out[i & 1][f0:f0 + piece], non_blocking=True)
    copy_pass(1)
    barrier()
    t0 = time.perf_counter()
    copy_pass(e2e_steps)
    torch.cuda.synchronize(dev)
    copy_s = time.perf_counter() - t0
    barrier()

    # BASELINE config 5, literally: a stream of 4096 synthetic host frames cut into contiguous per-GPU ranges
    # (sharding.frame_range), submitted in 64-frame pieces through the host-buffer call, one wait at the end
    from fs_uae_image_enhancer_project_b200.sharding import frame_range, max_over_ranks
    stream5 = None
    if args.stream_frames > 0:
        lo, hi = frame_range(args.stream_frames, world, rank)
        n5 = hi - lo
        s_host = s_out = None
        err = ""
        try:
            s_host = torch.empty((n5, FRAME_H, FRAME_W, 4), dtype=torch.uint8, pin_memory=True)
            s_out = torch.empty((n5, FRAME_H, FRAME_W, 4), dtype=torch.uint8, pin_memory=True)
        except RuntimeError as exc:                               # not enough pinnable host memory on this box: say so
            err = str(exc)[:200]
        if max_over_ranks(1.0 if err else 0.0, dev) > 0:          # every rank takes the same branch (collectives below)
            stream5 = {"frames": args.stream_frames, "unavailable": err or "another rank could not pin its range"}
        else:
            for f0 in range(0, n5, 256):                         # the stream is born on the device, frame g = lo + i
                k = min(256, n5 - f0)
                s_host[f0:f0 + k].copy_(synth.synth_rgb444_frames(k, FRAME_H, FRAME_W, seed=5000, first_frame=lo + f0, device=dev))
            torch.cuda.synchronize(dev)
            barrier()
            sampler5 = ClockSampler(local)
            if rank == 0:
                sampler5.start()
            t0 = time.perf_counter()
            for f0 in range(0, n5, BATCH):
                k = min(BATCH, n5 - f0)
                eng.submit_host(s_host[f0:f0 + k], s_out[f0:f0 + k], k, _lib.FMT_U8_NHWC4, _lib.FMT_U8_NHWC4, flags)
            eng.wait_host()
            stream_s = time.perf_counter() - t0
            clocks5 = sampler5.stop() if rank == 0 else None
            barrier()
            # latency of a piece in the stream: blocking 64-frame calls, submit -> all 64 frames back in host memory
            piece_ms = []
            for f0 in range(0, min(n5, 16 * BATCH), BATCH):
                k = min(BATCH, n5 - f0)
                t1 = time.perf_counter()
                eng.run_host(s_host[f0:f0 + k], s_out[f0:f0 + k], k, _lib.FMT_U8_NHWC4, _lib.FMT_U8_NHWC4, flags)
                piece_ms.append(1e3 * (time.perf_counter() - t1))
            stream_s = max_over_ranks(stream_s, dev)
            stream5 = {"frames": args.stream_frames, "frames_per_gpu": n5, "sharding": "contiguous ranges (sharding.frame_range)",
                       "value": args.stream_frames / stream_s, "unit": "frames/s", "seconds": stream_s,
                       "piece_frames": BATCH, "piece_latency_p50_ms": pctl(piece_ms, 0.5), "piece_latency_p99_ms": pctl(piece_ms, 0.99),
                       "api": "fsuae_engine_submit_host per 64-frame piece of the rank's range + one fsuae_engine_wait_host",
                       "clocks": clocks5}
        del s_host, s_out

    # sustained leg: the same device-resident step back to back for >= args.sustain seconds (the burst number above is
    # 20 steps = 50 ms; a long run meets the software power cap), clocks sampled throughout
    sustained = None
    if args.sustain > 0:
        n_sus = max(args.steps, int(args.sustain * 1000.0 / (ms / args.steps)) + 1)
        sampler2 = ClockSampler(local)
        if rank == 0:
            sampler2.start()
        barrier()
        ev0.record()
        for i in range(n_sus):
            step(i)
        ev1.record()
        barrier()
        sus_ms = max_over_ranks(ev0.elapsed_time(ev1), dev)
        sustained = {"steps": n_sus, "seconds": sus_ms / 1000.0, "value": BATCH * world * n_sus / (sus_ms / 1000.0),
                     "unit": "frames/s", "clocks": sampler2.stop() if rank == 0 else None}

    ms, e2e_s, e2e_sync_s, copy_s = (max_over_ranks(v, dev) for v in (ms, e2e_s, e2e_sync_s, copy_s))   # slowest rank
    if alt is not None:
        alt["ms_per_step"] = max_over_ranks(alt["ms_per_step"], dev)
        alt["value"] = BATCH * world / (alt["ms_per_step"] / 1000.0)
        alt["unit"] = "frames/s"

    if rank == 0:
        frames = BATCH * world * args.steps
        fps = frames / (ms / 1000.0)
        tf_burst, tf_sus, hbm_peak, how = measured_peaks()
        # which peak applies: a run at (nearly) maximum SM clock without a power cap is measured against the burst figure
        at_full_clock = bool(clocks and clocks.get("sm_mhz") and clocks["sm_mhz"] >= 0.97 * (clocks.get("sm_max_mhz") or 1e9)
                             and "sw_power_cap" not in clocks.get("reasons", []))
        tf_peak, peak_kind = (tf_burst, "burst") if at_full_clock else (tf_sus, "sustained")
        per_gpu_step_s = ms / 1000.0 / args.steps
        achieved_tf = GFLOP_PER_FRAME * BATCH / per_gpu_step_s / 1000.0
        # dominant kernel: algorithmic FLOPs of the layers it runs / its own event-timed duration
        mac = {"conv1": 3888, "conv2": 11664, "conv3": 23328, "conv4": 46656, "conv5": 23328, "conv6": 23328, "conv7": 3888}
        dom = max(per_kernel, key=per_kernel.get) if per_kernel else None
        traffic = committed_traffic()
        algo_bytes = BYTES_PER_FRAME_U8 * BATCH
        dom_line = None
        if dom:
            layers = list(mac) if dom == "fused_pass" else [l for l in mac if l in dom.replace("+", " ").replace("_", " ").split()]
            gflop = sum(2 * mac[l] * (FRAME_H // 2) * (FRAME_W // 2) for l in layers) * BATCH / 1e9
            dom_tf = gflop / per_kernel[dom]                    # GFLOP / ms = TFLOP/s
            dom_line = {"kernel": dom, "ms": per_kernel[dom], "share_of_step": per_kernel[dom] / sum(per_kernel.values()),
                        "achieved": dom_tf, "traffic": traffic.get(dom)}
        pass_traffic = sum(v for k, v in traffic.items() if k in per_kernel) if traffic and all(k in traffic for k in per_kernel) else None
        line = {
            "metric": "752x576 frames/sec", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"bf16": "bf16", "fp16": "f16", "fp32": "f32"}[precision], "data": "synthetic",
            "config": {"workload": WORKLOAD, "variant": eng.variant, "frames_per_gpu_per_step": BATCH,
                       "launches_per_step": int(launches_per_step),
                       "l2": "inputs rotate over 2 batches (443 MB in+out) > 126 MB L2",
                       "sharding": "frame-wise, one replica per GPU, no collective",
                       "host_numa_binding": f"{len(numa_cpus)} CPUs next to GPU 0" if numa_cpus else "none"},
            "latency_p50_ms": pctl(lat, 0.5), "latency_p99_ms": pctl(lat, 0.99),
            "latency_host_1frame_p50_ms": pctl(lat1, 0.5), "latency_host_1frame_p99_ms": pctl(lat1, 0.99),
            "us_per_frame": 1e6 * per_gpu_step_s / BATCH,
            "e2e": {"value": BATCH * world * e2e_steps / e2e_s, "unit": "frames/s",
                    "h2d_bytes_per_step": BATCH * FRAME_H * FRAME_W * 4, "d2h_bytes_per_step": BATCH * FRAME_H * FRAME_W * 4,
                    "steps": e2e_steps, "api": "fsuae_engine_submit_host per step + one fsuae_engine_wait_host (pinned host buffers)",
                    "sync_call_value": BATCH * world * e2e_steps / e2e_sync_s,
                    "sync_call_api": "one blocking fsuae_engine_run_host per step",
                    "copy_ceiling_fps": BATCH * world * e2e_steps / copy_s,
                    "frac_of_copy_ceiling": (BATCH * world * e2e_steps / e2e_s) / (BATCH * world * e2e_steps / copy_s),
                    "copy_ceiling_note": "the same H2D + D2H bytes in the same 16-frame pieces on two streams, all ranks at once, no kernel",
                    "clocks": e2e_clocks},
            "gpu_launches": int(launches_per_step * args.steps),
            "roofline": {"bound": "tensor", "achieved": dom_line["achieved"] if dom_line else achieved_tf, "peak": tf_peak,
                         "unit": "TFLOP/s", "frac": (dom_line["achieved"] if dom_line else achieved_tf) / tf_peak,
                         "frac_burst": (dom_line["achieved"] if dom_line else achieved_tf) / tf_burst,
                         "frac_sustained": (dom_line["achieved"] if dom_line else achieved_tf) / tf_sus,
                         "peak_kind": peak_kind,
                         "traffic": dom_line["traffic"] if dom_line else None,
                         "algorithmic_bytes": algo_bytes if dom == "fused_pass" else None,
                         "kernel": dom_line["kernel"] if dom_line else None,
                         "kernel_ms": dom_line["ms"] if dom_line else None,
                         "share_of_step": dom_line["share_of_step"] if dom_line else None,
                         "note": f"dominant kernel, CUDA events on its stream; peak = {how} {peak_kind} bf16 (chosen by the clocks record: "
                                 f"burst when the SM clock stayed >= 97 % of max without a power cap); traffic = dram bytes per launch from "
                                 f"the committed ncu --set full capture (profiles/r02_dram_traffic.json)"},
            "roofline_whole_pass": {"bound": "tensor", "achieved": achieved_tf, "peak": tf_peak, "unit": "TFLOP/s",
                                    "frac": achieved_tf / tf_peak, "frac_burst": achieved_tf / tf_burst, "frac_sustained": achieved_tf / tf_sus,
                                    "traffic": pass_traffic, "algorithmic_bytes": algo_bytes,
                                    "traffic_over_algorithmic": (pass_traffic / algo_bytes) if pass_traffic else None,
                                    "note": f"algorithmic 29.472 GFLOP/frame x {BATCH} frames / step device time (all kernels of the "
                                            f"pass); HBM floor: {algo_bytes / 1e9 / (hbm_peak) * 1e3:.3f} ms/step"},
            "kernel_ms": per_kernel,
            "layer_by_layer": alt,
            "parity_check": parity,
            "stream_config5": stream5,
            "sustained": sustained,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            cfps, threads = cpu_forward_fps(16, 4, sd={k: v.detach().float().cpu() for k, v in model.state_dict().items()}, frames=host[0][:16])
            line["cpu_baseline"] = {"value": cfps, "unit": "frames/s", "cores": threads, "kind": "port",
                                    "sample": "16 of the 64 frames, batch 4, oracle port of the PyTorch fp32 eval forward"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
