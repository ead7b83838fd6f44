"""Full-frame video stylisation: the generator applied to whole frames (north-star path; the reference's
per-tile loop generator.py:500-555 is replaced by one GeneratorJ.forward per frame), frames sharded across
ranks with no collective (see parallel.shard_range).

Per frame on the device:  uint8 HWC -> [ToTensor+Normalize fused into the P8 layout conversion] -> generator ->
[clamp, (x+1)*127.5, round fused into the uint8 HWC store]   (reference generator.py:584-616,643-647).
`stylize_host` adds the pinned-host <-> device copies, double-buffered on two streams so that the copies of
frame i+1 / i-1 overlap the convolutions of frame i.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .generator import GeneratorJ, _Engine
from .parallel import dist_env, shard_range


class FrameStylizer:
    #: frames per generator pass; None = chosen from the frame size (`pass_size`).  Several frames per pass fill the GPU
    #: better on the 1/16-resolution layers (their grids are ~1.7 waves per 1080p frame), shorten the last-wave tail of
    #: the persistent conv CTAs and amortise the ~70 launches of a pass; InstanceNorm stays per frame.  Measured
    #: (tools/frames_per_pass_sweep.py): 1080p 190 / 197 / 202 / 202 frames/s at 1 / 2 / 4 / 8 frames per pass,
    #: 960x540 (6 channels) 598 / 698 / 737 / 754.
    frames_per_pass = None

    def pass_size(self, h: int, w: int) -> int:
        """about 8 Mpixel of frames per pass: 8 at 960x540, 4 at 1080p, 2 at 4K (14 GB of activations per 4K frame)"""
        if self.frames_per_pass is not None:
            return max(1, int(self.frames_per_pass))
        floor = 2 if h * w <= 12e6 else 1        # beyond ~4K a single frame holds > 20 GB of activations
        return int(min(8, max(floor, round(8.3e6 / float(h * w)))))

    def __init__(self, gen: GeneratorJ):
        if not next(gen.parameters()).is_cuda:
            raise RuntimeError("FrameStylizer needs the generator on a CUDA device (no CPU path)")
        self.gen = gen.eval()
        gen._check_supported()
        if gen._engine is None:
            gen._engine = _Engine(gen)
        self.eng = gen._engine
        self.device = self.eng.device
        self._streams = None

    @torch.no_grad()
    def stylize_device(self, frames_u8: torch.Tensor, out_u8: Optional[torch.Tensor] = None,
                       masks: Optional[torch.Tensor] = None) -> torch.Tensor:
        """frames_u8: device uint8 [N,H,W,Cin] -> device uint8 [N,H,W,3]; `pass_size` frames per generator pass.
        masks (optional, device fp32 [N,H,W] in [0,1], e.g. from ops.mask_erode7): the frame's own RGB shows through
        where the mask is 0 (reference generator.py:562-563), fused into the uint8 conversion."""
        n, h, w, c = frames_u8.shape
        if c != self.gen.input_channels:
            raise ValueError(f"expected {self.gen.input_channels} channels, got {c}")
        if out_u8 is None:
            out_u8 = torch.empty((n, h, w, 3), dtype=torch.uint8, device=self.device)
        step = self.pass_size(h, w)
        for i in range(0, n, step):
            y = self.eng.forward(frames_u8[i:i + step], save=False, u8_hwc=True)
            ops.composite_to_u8(y, out_u8[i:i + step], frames_u8[i:i + step], None if masks is None else masks[i:i + step])
        return out_u8

    def stylize_video(self, frames_u8: torch.Tensor, out_u8: Optional[torch.Tensor] = None, rank: Optional[int] = None,
                      world: Optional[int] = None, masks: Optional[torch.Tensor] = None):
        """The frame loop of the reference driver (generator.py:674-705) over a whole video, sharded: rank r of R stylises
        the contiguous frame range `shard_range(N, r, R)` and leaves the rest to the other ranks - frames are independent,
        so there is no collective.  rank / world default to the torchrun environment (one process per GPU).
        frames_u8 / out_u8 index the WHOLE video; they may live on this rank's device or in pinned host memory (then the
        copies are overlapped with compute, see stylize_host).  Returns (lo, hi), the range this rank wrote."""
        env_rank, env_world, _ = dist_env()
        rank = env_rank if rank is None else rank
        world = env_world if world is None else world
        lo, hi = shard_range(frames_u8.shape[0], rank, world)
        if hi > lo:
            if frames_u8.is_cuda:
                if out_u8 is None:
                    raise ValueError("stylize_video writes into a caller-owned output video: pass out_u8")
                self.stylize_device(frames_u8[lo:hi], out_u8[lo:hi], None if masks is None else masks[lo:hi])
            else:
                if masks is not None:
                    raise NotImplementedError("host-resident videos: composite on the device path")
                self.stylize_host(frames_u8[lo:hi], out_u8[lo:hi])
        return lo, hi

    @torch.no_grad()
    def stylize_host(self, frames_pinned: torch.Tensor, out_pinned: torch.Tensor) -> None:
        """pinned host uint8 [N,H,W,Cin] -> pinned host uint8 [N,H,W,3], copies overlapped with compute"""
        if not (frames_pinned.is_pinned() and out_pinned.is_pinned()):
            raise ValueError("stylize_host needs pinned host buffers")
        n, h, w, c = frames_pinned.shape
        if self._streams is None:
            self._streams = (torch.cuda.Stream(self.device), torch.cuda.Stream(self.device))
        copy_in, copy_out = self._streams
        main = torch.cuda.current_stream(self.device)
        step = self.pass_size(h, w)
        dev_in = [torch.empty((step, h, w, c), dtype=torch.uint8, device=self.device) for _ in range(2)]
        dev_out = [torch.empty((step, h, w, 3), dtype=torch.uint8, device=self.device) for _ in range(2)]
        in_ready = [torch.cuda.Event() for _ in range(2)]
        in_free = [torch.cuda.Event() for _ in range(2)]
        out_ready = [torch.cuda.Event() for _ in range(2)]
        out_free = [torch.cuda.Event() for _ in range(2)]
        for e in in_free + out_free:
            e.record(main)
        for k, i in enumerate(range(0, n, step)):
            b = k & 1
            m = min(step, n - i)
            with torch.cuda.stream(copy_in):
                copy_in.wait_event(in_free[b])
                dev_in[b][:m].copy_(frames_pinned[i:i + m], non_blocking=True)
                in_ready[b].record(copy_in)
            main.wait_event(in_ready[b])
            main.wait_event(out_free[b])
            y = self.eng.forward(dev_in[b][:m], save=False, u8_hwc=True)
            ops.composite_to_u8(y, dev_out[b][:m])
            in_free[b].record(main)
            out_ready[b].record(main)
            with torch.cuda.stream(copy_out):
                copy_out.wait_event(out_ready[b])
                out_pinned[i:i + m].copy_(dev_out[b][:m], non_blocking=True)
                out_free[b].record(copy_out)
        main.wait_stream(copy_out)
