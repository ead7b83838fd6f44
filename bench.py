#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native hot path.

  python bench.py --gpus N --steps K --warmup W [--config C4|C5|C2]   (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...        (the reference's own CPU implementation of the path, rank 0 only)

Workload (default C4 = BASELINE.json configs[3]): a synthetic 1920x1080 RGB video of 2000 frames, full-frame GeneratorJ
inference, frames sharded across the ranks by the product API (`FrameStylizer.stylize_video` -> `parallel.shard_range`: one
contiguous frame range per GPU, no collective).  STRONG scaling: the video is the same for every N.
One step = one 200-frame segment of the video (a tenth of it; the video is walked cyclically, so --steps 20 = two passes);
every rank stylises its contiguous share of the segment.  The timed region holds exactly K steps between barrier +
synchronize pairs and the slowest rank's device time counts.
  value      frames/s with the uint8 video already resident in HBM (u8 -> generator -> u8 on device)
  e2e        the same through the same call with PINNED HOST segments: H2D of every frame and D2H of every stylised frame
             inside the timed region (copies double-buffered against compute)
  roofline   the dominant kernel (conv11 7x7 implicit GEMM, 50 % of all FLOPs): algorithmic FLOPs per launch / mean
             CUDA-event duration of its launches inside the timed region, against the measured bf16 peak
  cpu_baseline  the reference generator (baseline/_ref copy of the unmodified module; oracle port if absent) on the host
             cores, whole frames, bounded sample (rank 0, N=1)
  cudnn_bar  same-GPU library bar: the reference network as stock torch ops on cuDNN, channels_last, fp16 and bf16,
             cudnn.benchmark, the native arm's frames per pass
  train      config C3: patch training step INCLUDING the sampler (host draws + order tree + gather kernel) on synthetic
             keyframes with two guide directories (Cin 9), batch 80 x 80x80 per GPU, NCCL gradient all-reduce when N>1
--config C5 (3840x2160, 5 channels, 500 frames) and C2 (960x540, 6 channels, 7 frames looped x100) run the same legs on the
other BASELINE.json inference configurations.
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

CONFIGS = {
    # name: frame H, W, input channels, frames of the video, frames per step, fixture weights, description
    "C4": dict(h=1080, w=1920, cin=3, video=2000, seg=200, weights="gen_c3_trained.npz", metric="1080p stylized frames/s",
               what="C4: synthetic 1920x1080 RGB video, 2000 frames, full-frame GeneratorJ inference"),
    "C5": dict(h=2160, w=3840, cin=5, video=500, seg=50, weights="gen_cin5_trained.npz", metric="2160p stylized frames/s",
               what="C5: synthetic 3840x2160 video with 5 input channels (RGB + mask + flow guide), 500 frames, full-frame inference"),
    "C2": dict(h=960, w=540, cin=6, video=700, seg=70, weights="gen_cin6_trained.npz", metric="960x540 stylized frames/s",
               what="C2: 960x540 frames with the tracking guide (6 channels): one real PlatinumChan_x0.5_train frame (fixture) + 6 "
                    "synthetic ones, looped x100 = 700 frames, full-frame inference"),
}
UNIT = "frames/s"


def flops_per_pixel(cin: int) -> int:
    return 2017664 + 9408 * cin  # exact forward FLOPs per output pixel


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms during the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, pw, reasons = [], None, [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(pw) if pw else None}


def synthetic_frames(idx, h, w, c, device, base_seed=1234):
    """frames `idx` (iterable of frame numbers) of the synthetic video: low-frequency noise + fine detail so that
    InstanceNorm statistics are non-degenerate; seeded per frame, so the video does not depend on
    how it is sharded.  Channels beyond RGB: a binary blob mask (channel 3) and smooth guide fields."""
    import torch
    out = []
    for i in idx:
        g = torch.Generator(device=device).manual_seed(base_seed + int(i))
        low = torch.rand((1, c, h // 16 + 1, w // 16 + 1), generator=g, device=device) * 255
        img = torch.nn.functional.interpolate(low, size=(h, w), mode="bilinear", align_corners=False)
        img = img + (torch.rand((1, c, h, w), generator=g, device=device) * 16 - 8)
        if c in (4, 5):
            img[:, 3] = (img[:, 3] > 128).float() * 255
        out.append(img.clamp(0, 255).round().to(torch.uint8).permute(0, 2, 3, 1)[0])
    return torch.stack(out).contiguous()


def fixture_state_dict(name):
    import numpy as np
    import torch
    z = np.load(os.path.join(GOLD, name))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def native_generator(cfg, device, operand):
    """GeneratorJ with the reference-trained fixture weights of this channel count (tests/golden, oracle/make_golden*.py)"""
    from pbt_b200.generator import GeneratorJ
    g = GeneratorJ(input_channels=cfg["cin"], use_bias=True)
    g.load_state_dict(fixture_state_dict(cfg["weights"]), strict=True)
    g.operand_dtype = operand
    return g.to(device)


def reference_generator_module():
    """the UNMODIFIED reference module from baseline/_ref (copied there by __graft_entry__.build(); git-ignored, travels to
    the GPU box), or None"""
    p = os.path.join(ROOT, "baseline", "_ref", "src", "models", "generator.py")
    if not os.path.exists(p):
        return None
    spec = importlib.util.spec_from_file_location("_ref_generator", p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cpu_forward_fn(cfg):
    """(callable x[N,C,H,W] fp32 -> y, kind): the reference's CPU implementation of the generator"""
    import torch
    sd = fixture_state_dict(cfg["weights"])
    ref = reference_generator_module()
    if ref is not None:
        g = ref.GeneratorJ(input_channels=cfg["cin"], use_bias=True)
        g.load_state_dict(sd, strict=True)
        g.eval()
        return (lambda x: g(x)), "reference"
    from oracle import generator_oracle as go
    return (lambda x: go.generator_forward(sd, x)), "port"


def to_u8(y):
    """reference generator.py:643-647"""
    import torch
    return ((y.float().clamp(-1, 1) + 1) * 127.5).clamp(0, 255).permute(0, 2, 3, 1).round().to(torch.uint8)


def normalise(u8):
    return ((u8.permute(0, 3, 1, 2).float() / 255.0) - 0.5) / 0.5


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """the reference's own CPU implementation of the path on the host cores, all threads: WHOLE frames of the configured
    size through the unmodified reference GeneratorJ (baseline/_ref) - the oracle port only if that copy is absent.
    One step = one frame (a bounded sample of the 200-frame step of the native arm)."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import torch
    cfg = CONFIGS[args.config]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fwd, kind = cpu_forward_fn(cfg)
    h, w, c = cfg["h"], cfg["w"], cfg["cin"]
    frames = synthetic_frames(range(2), h, w, c, "cpu")
    warm = max(1, min(args.warmup, 2))
    with torch.no_grad():
        for i in range(warm):
            to_u8(fwd(normalise(frames[i % 2:i % 2 + 1])))
        t0 = time.perf_counter()
        for i in range(args.steps):
            to_u8(fwd(normalise(frames[i % 2:i % 2 + 1])))
        dt = time.perf_counter() - t0
    fps = args.steps / dt
    sample = (f"{args.steps} whole {w}x{h}x{c} frames, one per step (+{warm} warm-up), through "
              + ("the unmodified reference GeneratorJ (baseline/_ref)" if kind == "reference" else "the oracle port") + ", fp32, eval")
    emit({
        "impl": "reference", "metric": cfg["metric"], "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["what"], "frame": [h, w, c], "video_frames": cfg["video"], "frames_per_step": 1,
                   "weights": f"tests/golden/{cfg['weights']} (100 reference training steps)",
                   "sample": "one whole frame per step: a bounded sample of the native arm's 200-frame step"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# ------------------------------------------------------------------------------------------ native arm
def run_native(args):
    import torch
    import torch.distributed as dist
    from pbt_b200 import _native
    from pbt_b200.inference import FrameStylizer
    from pbt_b200.parallel import init_distributed, shard_range

    cfg = CONFIGS[args.config]
    H, W, CIN, VIDEO, SEG = cfg["h"], cfg["w"], cfg["cin"], cfg["video"], cfg["seg"]
    rank, world, local = init_distributed("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    peaks, peak_src = load_peaks()
    operand = os.environ.get("PBT_OPERAND", "fp16")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    def all_ranks(ms):
        """every rank's device time (the slowest one is what counts; the spread shows GPU-to-GPU variation)"""
        if world > 1:
            t = torch.tensor([ms], device=dev)
            every = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(every, t)
            return [float(e.item()) for e in every]
        return [ms]

    gen = native_generator(cfg, dev, operand)
    sty = FrameStylizer(gen)
    if os.environ.get("PBT_FRAMES_PER_PASS"):   # experiment knob; the default is the library's
        sty.frames_per_pass = int(os.environ["PBT_FRAMES_PER_PASS"])
    n_seg = VIDEO // SEG
    # the whole video is indexable on every rank; a rank only materialises the frames of its shares of the segments
    video = torch.empty((VIDEO, H, W, CIN), dtype=torch.uint8, device=dev)
    out = torch.empty((VIDEO, H, W, 3), dtype=torch.uint8, device=dev)
    segs_used = min(n_seg, args.steps + args.warmup)
    real = None
    if args.config == "C2":
        import numpy as np
        real = torch.from_numpy(np.load(os.path.join(GOLD, "gen_cin6_vectors.npz"))["frame_u8"]).to(dev)
    clip = None
    if real is not None:    # the 7-frame clip, looped
        clip = synthetic_frames(range(7), H, W, CIN, dev)
        clip[0] = real
    for s in range(segs_used):
        lo, hi = shard_range(SEG, rank, world)
        a = s * SEG
        for i in range(a + lo, a + hi, 8):
            j = min(i + 8, a + hi)
            video[i:j] = synthetic_frames(range(i, j), H, W, CIN, dev) if clip is None else clip[[k % 7 for k in range(i, j)]]
    torch.cuda.synchronize()

    def step(i):
        a = (i % n_seg) * SEG
        sty.stylize_video(video[a:a + SEG], out[a:a + SEG], rank, world)

    for i in range(args.warmup):
        step(i)
    clocks = ClockSampler(local)
    barrier()
    if rank == 0:
        clocks.start()
    sty.eng.kernel_timer = []
    l0 = _native.LAUNCHES[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record()
    barrier()
    launches = _native.LAUNCHES[0] - l0
    rank_ms = all_ranks(e0.elapsed_time(e1))
    ms_total = max(rank_ms)
    clk = clocks.stop() if rank == 0 else None
    kt = sty.eng.kernel_timer
    sty.eng.kernel_timer = None
    per_pass = sty.pass_size(H, W)
    full = [(a, b) for a, b, nf in kt if nf == per_pass]      # launches carrying a full pass (the ragged last pass of a share is left out)
    k_ms = sum(a.elapsed_time(b) for a, b in full) / max(1, len(full))
    value = SEG * args.steps / (ms_total / 1e3)
    out_checksum = int(out[(args.warmup % n_seg) * SEG + shard_range(SEG, rank, world)[0]].sum())

    # ---- e2e: the same call on PINNED HOST segments: copies of every frame in and out inside the timed region
    host_in = torch.empty((SEG, H, W, CIN), dtype=torch.uint8).pin_memory()
    host_out = torch.empty((SEG, H, W, 3), dtype=torch.uint8).pin_memory()
    lo, hi = shard_range(SEG, rank, world)
    host_in[lo:hi].copy_(video[lo:hi].cpu())
    for _ in range(max(1, min(2, args.warmup))):
        sty.stylize_video(host_in, host_out, rank, world)
    barrier()
    e0.record()
    for _ in range(args.steps):
        sty.stylize_video(host_in, host_out, rank, world)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = SEG * args.steps / (ms_e2e / 1e3)
    checksum = int(host_out[lo:hi].sum()) if hi > lo else 0   # the result is really read on the host
    same = bool(torch.equal(host_out[lo:lo + 1].to(dev), out[lo:lo + 1])) if hi > lo else True
    del video, out
    torch.cuda.empty_cache()

    train = None
    if not args.no_train:
        try:
            train = train_leg(args, rank, world, dev, operand, barrier, max_over_ranks, all_ranks)
        except Exception as e:  # noqa: BLE001 - secondary metric: never lose the headline line over it
            train = {"unavailable": f"{type(e).__name__}: {e}"[:300]}

    if rank != 0:
        return
    cpu = bar = None
    if world == 1 and not args.no_cpu:
        cpu = cpu_leg(cfg, sty, dev)
        bar = cudnn_bar(cfg, per_pass, dev)

    conv11_flops = 2.0 * H * W * (49 * (160 + CIN)) * 64 * per_pass
    peak_tf = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))
    achieved = conv11_flops / (k_ms / 1e3) / 1e12 if k_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "conv11_traffic.json")
    if os.path.exists(tp) and args.config == "C4":
        with open(tp) as f:
            tj = json.load(f)
        traffic = tj.get("dram_bytes_per_launch")   # one ncu --set full capture of a conv11 launch (profiles/)
        if traffic is not None:
            traffic = int(traffic * per_pass / max(1, int(tj.get("frames_per_launch", 1))))
    line = {
        "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": operand + " operands, f32 accumulate", "data": "synthetic",
        "config": {"workload": cfg["what"],
                   "sharding": "FrameStylizer.stylize_video: one contiguous frame range per GPU, no collective",
                   "frames_per_step": SEG, "video_frames": VIDEO, "frames_per_generator_pass": per_pass, "frame": [H, W, CIN],
                   "weights": f"tests/golden/{cfg['weights']} (100 reference training steps)",
                   "l2": f"the {min(segs_used * SEG, VIDEO)} resident frames are distinct and each pass streams ~{3.5 * per_pass * H * W / (1080 * 1920):.0f} GB of activations >> 126 MB L2"},
        "timed_region_s": ms_total / 1e3, "timed_region_s_per_rank": [round(m / 1e3, 4) for m in rank_ms],
        "tflops_algorithmic": flops_per_pixel(CIN) * H * W * value / 1e12,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": SEG * H * W * CIN, "d2h_bytes_per_step": SEG * H * W * 3,
                "ms_per_step": ms_e2e / args.steps, "timed_region_s": ms_e2e / 1e3, "host_checksum": checksum,
                "host_result_equals_device_result": same},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "kernel": f"conv_igemm_kernel (conv11 7x7, {160 + CIN}->64, CTA-pair configuration, norm+ReLU of up1 on load)",
                     "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf if peak_tf else None,
                     "frac_of_burst_peak": achieved / peaks["bf16_tflops"] if peaks.get("bf16_tflops") else None,
                     "traffic": traffic, "traffic_algorithmic": int(per_pass * H * W * 2 * (128 + 32 + CIN + 64)),   # 16-bit reads of the 160+Cin real input channels + the 64-channel store
                     "traffic_source": "profiles/conv11_traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one launch)" if traffic else None,
                     "peak_source": f"{peak_src} bf16_tflops_sustained (the kernel is timed inside a {ms_total / 1e3:.1f}-s step loop)",
                     "launch_ms": k_ms, "launches_timed": len(full), "frames_per_launch": per_pass},
        "clocks": clk, "device_checksum": out_checksum,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if bar is not None:
        line["cudnn_bar"] = bar
    if train is not None:
        line["train"] = train
    emit(line)


def cpu_leg(cfg, sty, dev):
    """the reference generator on the host cores: WHOLE frames of this configuration, about 10-20 s of CPU work; the same
    frame through the native path must agree (parity guard on the benchmark itself)"""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    H, W, CIN = cfg["h"], cfg["w"], cfg["cin"]
    fwd, kind = cpu_forward_fn(cfg)
    u8 = synthetic_frames(range(1), H, W, CIN, "cpu")
    with torch.no_grad():
        t0 = time.perf_counter()
        reps = 0
        while reps < 2 or (time.perf_counter() - t0 < 10.0 and reps < 8):
            yc = to_u8(fwd(normalise(u8)))
            reps += 1
        dtc = time.perf_counter() - t0
    cpu = {"value": reps / dtc, "unit": UNIT, "cores": cores, "kind": kind,
           "sample": f"{reps} whole {W}x{H}x{CIN} frames through " +
                     ("the unmodified reference GeneratorJ (baseline/_ref)" if kind == "reference" else "oracle/generator_oracle.py") +
                     ", fp32, all host threads (no warm-up pass excluded)"}
    yn = sty.stylize_device(u8.to(dev))
    diff = (yn.cpu().int() - yc.int()).abs()
    cpu["max_abs_u8_diff_vs_native"] = int(diff.max())
    cpu["frac_u8_pixels_differing"] = float((diff > 0).float().mean())
    try:
        # the reference training step on the host cores (config C1: batch 40 of 32x32 patches)
        from oracle import generator_oracle as go
        sd1 = fixture_state_dict("gen_c3_trained.npz")
        opt1 = go.AdamState([k for k, v in sd1.items() if v.is_floating_point() and "running_" not in k])
        gcpu = torch.Generator().manual_seed(5)
        x1, t1 = torch.rand(40, 3, 32, 32, generator=gcpu) * 2 - 1, torch.rand(40, 3, 32, 32, generator=gcpu) * 2 - 1
        go.g_only_train_step(sd1, opt1, x1, t1)
        t0, n1 = time.perf_counter(), 0
        while n1 < 2 or time.perf_counter() - t0 < 4.0:
            go.g_only_train_step(sd1, opt1, x1, t1)
            n1 += 1
        cpu["train_c1"] = {"value": 40 * n1 / (time.perf_counter() - t0), "unit": "patches/s", "cores": cores,
                           "sample": f"{n1} G-only steps, batch 40 x 32x32, oracle port (fp32)"}
    except Exception as e:  # noqa: BLE001
        cpu["train_c1"] = {"unavailable": f"{type(e).__name__}: {e}"[:160]}
    return cpu


def cudnn_bar(cfg, per_pass, dev):
    """same-GPU library bar: the reference network as stock torch ops - cuDNN convolutions, unfused norms,
    cat and Upsample - in channels_last, fp16 and bf16 (+ fp32 as trained), cudnn.benchmark on, the native arm's frames
    per pass, >= 20 timed passes"""
    import torch
    from oracle import generator_oracle as go
    H, W, CIN = cfg["h"], cfg["w"], cfg["cin"]
    res = {"frames_per_pass": per_pass, "memory_format": "channels_last", "cudnn_benchmark": True,
           "note": "generator only (no uint8 conversion); functional torch graph of the reference network on this B200"}
    sd = fixture_state_dict(cfg["weights"])
    xg = normalise(synthetic_frames(range(per_pass), H, W, CIN, dev))
    old = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    try:
        for name, dt_, reps in (("fp16", torch.float16, 20), ("bf16", torch.bfloat16, 20), ("fp32", torch.float32, 5)):
            try:
                sdg = {k: (v.to(dev, dt_) if v.is_floating_point() else v.to(dev)) for k, v in sd.items()}
                sdg = {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in sdg.items()}
                xin = xg.to(dt_).contiguous(memory_format=torch.channels_last)
                with torch.no_grad():
                    for _ in range(3):
                        go.generator_forward(sdg, xin)
                    torch.cuda.synchronize(dev)
                    e0.record()
                    for _ in range(reps):
                        go.generator_forward(sdg, xin)
                    e1.record()
                    torch.cuda.synchronize(dev)
                res[name] = per_pass * reps / (e0.elapsed_time(e1) / 1e3)
                del sdg, xin
            except Exception as e:  # noqa: BLE001 - e.g. out of memory at 4K fp32
                res[name] = f"unavailable: {type(e).__name__}"
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.benchmark = old
    res["unit"] = UNIT
    return res


def synthetic_keyframes(n, h, w, seed):
    """n keyframes for the C3 training leg: (pre, post, gauss guide, flow guide) uint8 [h,w,3] and ellipse masks covering
    ~12 % of the frame"""
    import numpy as np
    from PIL import Image
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w]

    def smooth():
        base = rng.randint(0, 256, (h // 24, w // 24, 3)).astype(np.uint8)
        img = np.asarray(Image.fromarray(base).resize((w, h), Image.BILINEAR)).astype(np.int16)
        return np.clip(img + rng.randint(-6, 7, (h, w, 1)), 0, 255).astype(np.uint8)

    pre, post, ga, fl, mask = [], [], [], [], []
    for _ in range(n):
        pre.append(smooth()); post.append(smooth()); ga.append(smooth()); fl.append(smooth())  # noqa: E702
        cy, cx = rng.randint(h // 3, 2 * h // 3), rng.randint(w // 3, 2 * w // 3)
        mask.append(((((yy - cy) / (0.24 * h)) ** 2 + ((xx - cx) / (0.16 * w)) ** 2) <= 1).astype(np.uint8) * 255)
    return pre, post, ga, fl, mask


def train_leg(args, rank, world, dev, operand, barrier, max_over_ranks, all_ranks):
    """config C3: the generator half of the reference training_step (lightning_model.py:211-250,260-292) per rank: sampler
    draw + gather -> forward -> L1*4 -> backward -> (NCCL mean all-reduce) -> clip 0.5 -> Adam, batch 80 x 80x80, Cin 9"""
    import numpy as np
    import torch
    import torch.distributed as dist
    from lightning_model import _IndexLoader
    from pbt_b200.generator import GeneratorJ
    from pbt_b200.graphs import GraphedGeneratorStep
    from pbt_b200.optim import FusedClipAdam
    from pbt_b200.parallel import GradAllReduce, broadcast_module_state, replicas_identical
    from pbt_b200.sampler import StyleTransferDataset

    B, P = 80, 80
    pre, post, ga, fl, mask = synthetic_keyframes(6, 1080, 1920, seed=7)     # same keyframes on every rank
    ds = StyleTransferDataset.from_arrays(pre, post, mask, P, additional={"gauss": ga, "flow": fl}, device=str(dev))
    del pre, post, ga, fl, mask
    np.random.seed(1000 + rank)            # per-rank draw stream (the reference: per-worker numpy seeds)
    torch.manual_seed(0)
    tg = GeneratorJ(input_channels=9, use_bias=True)
    tg.load_state_dict(fixture_state_dict("gen_cin9_trained.npz"), strict=True)
    tg = tg.to(dev).train()
    tg.operand_dtype = operand
    broadcast_module_state(tg)
    opt = FusedClipAdam(tg.parameters(), lr=4e-4, betas=(0.9, 0.999), weight_decay=1e-5, max_grad_norm=0.5)
    loader = iter(_IndexLoader(ds, B, rank, world))
    checks = {}
    ar = None
    if world > 1:
        # (1) the NCCL-reduced gradient equals the mean of the per-rank gradients: one sweep without the exchange,
        # all-gathered and averaged with tensor-library ops, against one sweep through GradAllReduce on the same batch
        b0 = next(loader)
        x0, t0_ = b0["combined_input"], b0["post"]
        bn = tg.smoothers[2]
        keep = (bn.running_mean.clone(), bn.running_var.clone(), bn.num_batches_tracked.clone())
        (torch.nn.functional.l1_loss(tg(x0), t0_) * 4.0).backward()
        mine = tg._engine.grad_bucket().flat.clone()
        every = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        expect = torch.stack(every).mean(0)
        tg.zero_grad(set_to_none=True)
        ar = GradAllReduce(list(tg.named_parameters()), world=world).attach(tg)
        (torch.nn.functional.l1_loss(tg(x0), t0_) * 4.0).backward()
        ar.finish()
        torch.cuda.synchronize()
        err = float((ar.flat - expect).abs().max() / expect.abs().max().clamp_min(1e-30))
        checks["allreduce_max_err_rel_to_peak"] = err       # wgrad accumulates with fp32 atomics: two sweeps differ in the last bits
        checks["allreduce_matches_mean_of_rank_gradients"] = bool(err < 1e-3)
        tg.zero_grad(set_to_none=True)
        with torch.no_grad():
            bn.running_mean.copy_(keep[0]); bn.running_var.copy_(keep[1]); bn.num_batches_tracked.copy_(keep[2])  # noqa: E702

    gstep = GraphedGeneratorStep(tg, opt, (B, 9, P, P), clip=0.5, grad_sync=ar)
    mode = "cuda-graph replay of the whole step (forward, loss, backward, all-reduce, clip, Adam)"

    def pipeline_step():
        b = next(loader)                    # host draws + order-statistic tree + H2D of the positions + gather kernel
        return gstep(b["combined_input"], b["post"])

    tsteps = max(20, args.steps)
    for _ in range(5):
        pipeline_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(tsteps):
        loss = pipeline_step()
    e1.record()
    barrier()
    ms_p = max_over_ranks(e0.elapsed_time(e1))
    # the step alone on a resident batch (comparable with round 1's figure)
    fixed = next(loader)
    xs, ts = fixed["combined_input"].clone(), fixed["post"].clone()
    barrier()
    e0.record()
    for _ in range(tsteps):
        loss = gstep(xs, ts)
    e1.record()
    barrier()
    ms_s = max_over_ranks(e0.elapsed_time(e1))
    local = None
    if world > 1:
        # after all those graph-replayed steps every rank must hold bit-identical weights (BatchNorm running statistics are
        # per-rank by design: DDP would overwrite them with rank 0's before each forward, and rank 0 writes the checkpoint)
        checks["replicas_identical"] = bool(replicas_identical(tg, buffers=False))
        # how much of the step time is the exchange: the same step captured WITHOUT the gradient all-reduce, every rank on its
        # own (run last: the replicas diverge from here on).  Its slowest rank bounds what any synchronous step can reach.
        tg._engine.grad_hook = None
        lstep = GraphedGeneratorStep(tg, opt, (B, 9, P, P), clip=0.5, grad_sync=None)
        for _ in range(3):
            lstep(xs, ts)
        barrier()
        e0.record()
        for _ in range(tsteps):
            lstep(xs, ts)
        e1.record()
        barrier()
        per_rank = [m / tsteps for m in all_ranks(e0.elapsed_time(e1))]
        local = {"ms_per_step_per_rank": [round(m, 4) for m in per_rank], "ms_per_step_slowest_rank": max(per_rank),
                 "allreduce_exposed_ms": ms_s / tsteps - max(per_rank),
                 "note": "the same graph-replayed step without the gradient exchange, ranks independent"}
    lt = loss.detach().clone().reshape(1)
    if world > 1:
        dist.all_reduce(lt)
        lt /= world
    flops = 3 * flops_per_pixel(9) * P * P * B * world
    res = {"metric": "train patches/s", "value": world * B * tsteps / (ms_p / 1e3), "unit": "patches/s",
           "ms_per_step": ms_p / tsteps, "steps": tsteps,
           "config": {"workload": "C3: sampler (numpy draws, order-statistic tree, patch gather kernel) + G-only step (L1*4, clip 0.5, Adam "
                                  "lr 4e-4 wd 1e-5; clip+Adam fused), batch 80 x 80x80 patches per GPU, Cin 9 (RGB + 2 guide dirs), "
                                  "6 synthetic 1080p keyframes resident per GPU, weights tests/golden/gen_cin9_trained.npz",
                      "allreduce_bytes_per_step": (ar.nbytes if ar else 0), "allreduce_groups": (len(ar.group_bounds) if ar else 0),
                      "launch_mode": mode},
           "tflops_algorithmic": flops / (ms_p / tsteps) / 1e9,
           "step_only": {"value": world * B * tsteps / (ms_s / 1e3), "ms_per_step": ms_s / tsteps,
                         "tflops_algorithmic": flops / (ms_s / tsteps) / 1e9, "note": "graph replay on a resident batch, sampler excluded"},
           "final_loss_mean_over_ranks": float(lt), "skipped_steps": opt.skipped_steps, **checks}
    if local is not None:
        res["without_allreduce"] = local
    if world == 1:
        res["gan_step"] = gan_leg(xs, ts, dev, operand, tsteps, barrier)
    return res


def gan_leg(xs, ts, dev, operand, tsteps, barrier):
    """the reference's full step with the adversarial branch on (critic + generator, lightning_model.py:224-250), same
    shape, replayed as one CUDA graph; reported next to the G-only step, never as the headline"""
    import torch
    try:
        from lightning_model import StyleTransferModel
        B = xs.shape[0]
        tcfg = {"batch_size": B, "reconstruction_weight": 4.0, "adversarial_weight": 0.5, "use_image_loss": True,
                "reconstruction_criterion": "L1Loss", "adversarial_criterion": "MSELoss",
                "use_gradient_clipping": True, "gradient_clip_val": 0.5, "cuda_graph": True}
        adam = {"lr": 4e-4, "betas": [0.9, 0.999], "weight_decay": 1e-5}
        torch.manual_seed(0)
        gm = StyleTransferModel({"args": {"input_channels": 9, "use_bias": True}},
                                {"args": {"input_channels": 3, "num_filters": 12, "n_layers": 2, "use_bias": True}},
                                tcfg, {"generator": dict(adam), "discriminator": dict(adam)},
                                {"additional_channels": {}}).to(dev).train()
        gm.generator.load_state_dict(fixture_state_dict("gen_cin9_trained.npz"), strict=True)
        gm.generator.operand_dtype = operand
        gm._optimizers = gm.configure_optimizers()
        gbatch = {"combined_input": xs, "post": ts}
        for i in range(4):
            gm.graphed_training_step(gbatch, i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(tsteps):
            gout = gm.graphed_training_step(gbatch, i)
        e1.record()
        barrier()
        ms_g = e0.elapsed_time(e1) / tsteps
        res = {"value": B / (ms_g / 1e3), "unit": "patches/s", "ms_per_step": ms_g,
               "workload": "C3 shape, critic (DiscriminatorN_IN 12 filters, 2 layers) + generator update, one generator forward "
                           "shared by both halves, one CUDA-graph replay",
               "g_total_loss": float(gout["g_total_loss"]), "d_total_loss": float(gout["d_total_loss"])}
        # the reference's default three-term generator loss (config/model/default.yaml:29-37): + 6.0 x VGG19 taps [0, 3, 5].  The
        # ImageNet checkpoint is not available offline: the taps run on a seeded VGG19 prefix of the same shapes (same cost).
        try:
            from pbt_b200.perceptual import vgg19_prefix
            from src.models.perception import PerceptualVGG19
            gm._graphed = None
            gm.perception_loss_model = PerceptualVGG19.from_features(vgg19_prefix(6), [0, 3, 5], use_normalization=False).to(dev)
            gm.perception_loss_weight = 6.0
            for i in range(4):
                gm.graphed_training_step(gbatch, i)
            barrier()
            e0.record()
            for i in range(tsteps):
                gout = gm.graphed_training_step(gbatch, i)
            e1.record()
            barrier()
            ms_p = e0.elapsed_time(e1) / tsteps
            res["with_perceptual"] = {"ms_per_step": ms_p, "value": B / (ms_p / 1e3), "unit": "patches/s",
                                      "workload": "the same step + the perceptual term (VGG19 features [0, 3, 5], weight 6.0) on the "
                                                  "native kernels; synthetic VGG weights",
                                      "native_taps": gm.perception_loss_model.native_unsupported(ts) is None,
                                      "g_perception_loss": float(gout["g_perception_loss"])}
        except Exception as e:  # noqa: BLE001
            res["with_perceptual"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
        return res
    except Exception as e:  # noqa: BLE001 - secondary figure: never fail the bench line over it
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


def emit(line: dict) -> None:
    """the ONE JSON line goes to the real stdout; everything else this process (or a C library such as NCCL's
    version banner) prints to fd 1 was redirected to stderr in main()"""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="C4", choices=sorted(CONFIGS))
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:  # noqa: BLE001
        pass


if __name__ == "__main__":
    main()
