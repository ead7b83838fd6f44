"""Thin Python wrappers over the C-ABI ops of libpbt.so plus the host-side weight packing.

Every function enqueues native CUDA work on torch's current stream and returns
immediately; PyTorch only owns the buffers.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _native as nv
from ._native import ACT_LEAKY, ACT_NONE, ACT_RELU, BF16, FP16, P8, Act, act_or_null, check, lib, ptr, stream_ptr

__all__ = [
    "pack_conv_weight", "pack_conv_weight_tap_pairs", "s2d_weight", "s2d_weight_grad", "dgrad_weight", "conv_fwd", "conv_wgrad", "conv_num_tiles",
    "norm_finalize", "norm_apply", "upsample2x", "upsample2x_bwd", "norm_bwd", "head_bwd", "channel_sum", "nchw_to_p8",
    "p8_to_nchw", "p8f_to_nchw", "u8hwc_to_p8", "nchw_to_u8hwc", "u8hwc_to_norm_chw", "patch_gather", "mask_dilate7",
    "mask_erode7", "composite_to_u8", "zero_border", "p8s2d_to_nchw",
    "absmax", "make_grad_scale", "grad_scale_feedback",
]


# ----------------------------------------------------------------------------- weights
def pack_conv_weight(w: torch.Tensor, cin_pad: int, blk_c: int, dt: int, pair: bool = False) -> torch.Tensor:
    """[cout, cin, kh, kw] fp32 -> flat 16-bit buffer  w[cb][tap][k/8][cout][k%8]  (see include/pbt.h);
    pair=True: the CTA-pair layout w[cb][half][tap][k/8][cout/2][k%8]."""
    cout, cin, kh, kw = w.shape
    assert cin_pad % 16 == 0 and cin_pad >= cin and cout % 16 == 0
    wp = w.new_zeros((cout, cin_pad, kh, kw))
    wp[:, :cin] = w
    t = wp.permute(2, 3, 1, 0).reshape(kh * kw, cin_pad, cout)  # [tap][ci][co]
    chunks = []
    for c0 in range(0, cin_pad, blk_c):
        c1 = min(c0 + blk_c, cin_pad)
        blk = t[:, c0:c1].reshape(kh * kw, (c1 - c0) // 8, 8, cout).permute(0, 1, 3, 2)  # [tap][k8][co][8]
        if pair:
            blk = blk.reshape(kh * kw, (c1 - c0) // 8, 2, cout // 2, 8).permute(2, 0, 1, 3, 4)  # [half][tap][k8][co/2][8]
        chunks.append(blk.reshape(-1))
    return torch.cat(chunks).to(nv.torch_dtype(dt)).contiguous()


def pack_conv_weight_tap_pairs(w: torch.Tensor, dt: int) -> torch.Tensor:
    """[cout, cin <= 8, kh, kw] -> tap-pair layout [ky * ceil(kw/2) + j][g][cout][c] = w[cout][c][ky][2j + g] (see include/pbt.h)"""
    cout, cin, kh, kw = w.shape
    assert cin <= 8
    kwp = (kw + 1) // 2
    t = w.new_zeros((kh, kwp, 2, cout, 8))
    for j in range(kwp):
        for g in range(2):
            if 2 * j + g < kw:
                t[:, j, g, :, :cin] = w[:, :, :, 2 * j + g].permute(2, 0, 1)
    return t.reshape(-1).to(nv.torch_dtype(dt)).contiguous()


def dgrad_weight(w: torch.Tensor) -> torch.Tensor:
    """weights of the stride-1 conv that maps dY -> dX: channel-transposed, tap-flipped."""
    return w.transpose(0, 1).flip(2, 3).contiguous()


_S2D_TAPS = ((0, 1, 0), (1, 0, 1), (1, 1, 2))  # (s2d tap, phase, original 3x3 tap)


def s2d_weight(w: torch.Tensor) -> torch.Tensor:
    """3x3 stride-2 pad-1 kernel [co, c, 3, 3] -> equivalent 2x2 stride-1 (pad top/left 1) kernel over the
    space-to-depth input [co, 4c, 2, 2]; channel = ((y&1)*2 + (x&1))*c + ci."""
    co, c = w.shape[0], w.shape[1]
    w2 = w.new_zeros((co, 4 * c, 2, 2))
    for ty, py, dy in _S2D_TAPS:
        for tx, px, dx in _S2D_TAPS:
            ph = py * 2 + px
            w2[:, ph * c:(ph + 1) * c, ty, tx] = w[:, :, dy, dx]
    return w2


def s2d_weight_grad(dw2: torch.Tensor, c: int) -> torch.Tensor:
    """inverse of s2d_weight for gradients: [co, 4c, 2, 2] -> [co, c, 3, 3]."""
    co = dw2.shape[0]
    dw = dw2.new_zeros((co, c, 3, 3))
    for ty, py, dy in _S2D_TAPS:
        for tx, px, dx in _S2D_TAPS:
            ph = py * 2 + px
            dw[:, :, dy, dx] = dw2[:, ph * c:(ph + 1) * c, ty, tx]
    return dw


class WeightPacker:
    """device-side job table for pbt_pack_weights: every conv parameter -> packed 16-bit operands in ONE launch"""

    def __init__(self, device):
        self.device = device
        self.jobs = []
        self.out = {}
        self._table = None
        self._max = 0

    def add(self, name: str, param: torch.Tensor, *, s2d=False, dgrad=False, k_pad: int, n_out: int, n_keep: int, blk_c: int,
            dt: int, pair: bool = False, s2d4_cpp: int = 0, tap_pairs: bool = False) -> torch.Tensor:
        """s2d4_cpp > 0 (with s2d): the source is a 4x4 stride-2 pad-1 kernel, packed as the 3x3 stride-1 kernel over the
        space-to-depth input that has s2d4_cpp channels per phase"""
        co, ci, kh, kw = param.shape
        taps = (9 if s2d4_cpp else 4) if s2d else (kh * ((kw + 1) // 2) if tap_pairs else kh * kw)
        if tap_pairs:
            assert ci <= 8 and k_pad == 16 and not (s2d or dgrad or pair)
        dst = torch.empty(taps * k_pad * n_out, dtype=nv.torch_dtype(dt), device=self.device)
        assert param.is_contiguous() and param.dtype in (torch.float32, torch.float16)
        self.jobs.append(nv.PackJob(param.data_ptr(), dst.data_ptr(), co, ci, kh, kw, int(s2d) | (int(dgrad) << 1) | (int(pair) << 2) | (8 if s2d4_cpp else 0) | (16 if tap_pairs else 0),
                                    k_pad, n_out, n_keep, blk_c, dt, int(param.dtype == torch.float16), int(s2d4_cpp)))
        self._src = getattr(self, "_src", []) + [param]   # keep the parameters alive / pointers stable
        self.out[name] = dst
        self._max = max(self._max, dst.numel())
        self._table = None
        return dst

    def run(self) -> None:
        if self._table is None:
            arr = (nv.PackJob * len(self.jobs))(*self.jobs)
            raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
            self._table = raw.to(self.device)
        check(lib().pbt_pack_weights(self._table.data_ptr(), len(self.jobs), self._max, stream_ptr()), "pbt_pack_weights")


# ----------------------------------------------------------------------------- convolution
def conv_num_tiles(h: int, w: int, tiles_per_cta: int) -> int:
    return lib().pbt_conv_num_tiles(h, w, tiles_per_cta)


def conv_fwd(x: P8, wpack: torch.Tensor, cout: int, kh: int, kw: int, pad_t: int, pad_l: int, dt: int, *,
             blk_c: int = 32, tiles_per_cta: int = 2, bias=None, act: int = ACT_NONE, post_scale=None, post_shift=None,
             mask: P8 | None = None, addend32=None, out32=None, out: P8 | None = None, stats_partial=None,
             head_w=None, head_b=None, head_out=None, head_tanh: bool = True, upsample2x: bool = False, debug_flags: int = 0,
             debug_buf=None, pre: P8 | None = None, pre_scale=None, pre_shift=None, pre_act: int = ACT_NONE, ctas_per_sm: int = 0, cta_pair: bool = False, concurrent: bool = False, batch_tiles: bool = False,
             valid_hw=None, tap_pairs: bool = False, up_raw_channels: int = 0) -> None:
    """`pre` (raw output of the previous conv) supplies the first channels, normalised + activated on load; `x` (may be
    None then) the remaining ones."""
    d = nv.ConvDesc()
    d.inp = act_or_null(x)
    d.pre = act_or_null(pre)
    d.pre_scale, d.pre_shift, d.pre_act = ptr(pre_scale), ptr(pre_shift), pre_act
    d.ctas_per_sm = ctas_per_sm
    d.cta_pair = int(cta_pair)
    d.concurrent = int(concurrent)
    d.batch_tiles = int(batch_tiles)
    d.tap_pairs = int(tap_pairs)
    d.up_raw_channels = int(up_raw_channels)
    d.valid_h, d.valid_w = (0, 0) if valid_hw is None else (int(valid_hw[0]), int(valid_hw[1]))
    d.wpack = wpack.data_ptr()
    d.cout, d.kh, d.kw, d.pad_t, d.pad_l = cout, kh, kw, pad_t, pad_l
    d.blk_c, d.tiles_per_cta, d.dtype = blk_c, tiles_per_cta, dt
    d.bias, d.act = ptr(bias), act
    d.post_scale, d.post_shift = ptr(post_scale), ptr(post_shift)
    d.mask = act_or_null(mask)
    d.addend32, d.out32 = ptr(addend32), ptr(out32)
    d.out = act_or_null(out)
    d.stats_partial = ptr(stats_partial)
    d.head_w, d.head_b, d.head_out, d.head_tanh = ptr(head_w), ptr(head_b), ptr(head_out), int(head_tanh)
    d.upsample2x = int(upsample2x)
    d.debug_flags = debug_flags
    d.debug_buf = ptr(debug_buf)
    check(lib().pbt_conv_fwd(C.byref(d), stream_ptr()), "pbt_conv_fwd")


def conv_wgrad(x: P8, dy: P8, kh: int, kw: int, pad_t: int, pad_l: int, dt: int, dw: torch.Tensor, inv_scale=None,
               debug_flags: int = 0) -> None:
    """dw: fp32 [kh*kw, cin, cout], accumulated into."""
    d = nv.WgradDesc()
    d.x, d.dy = x.act(), dy.act()
    d.kh, d.kw, d.pad_t, d.pad_l, d.dtype = kh, kw, pad_t, pad_l, dt
    d.dw, d.inv_scale, d.debug_flags = dw.data_ptr(), ptr(inv_scale), debug_flags
    check(lib().pbt_conv_wgrad(C.byref(d), stream_ptr()), "pbt_conv_wgrad")


# ----------------------------------------------------------------------------- normalisation
def norm_finalize(partial, n, tiles, c, count_per_image, scale, shift, *, eps=1e-5, batch_mode=False, gamma=None,
                  beta=None, running_mean=None, running_var=None, momentum=0.1, mean_out=None, rstd_out=None) -> None:
    check(lib().pbt_norm_finalize(partial.data_ptr(), n, tiles, c, count_per_image, eps, int(batch_mode), ptr(gamma),
                                  ptr(beta), ptr(running_mean), ptr(running_var), momentum, scale.data_ptr(),
                                  shift.data_ptr(), ptr(mean_out), ptr(rstd_out), stream_ptr()), "pbt_norm_finalize")


def norm_apply(x: P8, dt: int, *, scale=None, shift=None, per_channel=False, act=ACT_NONE, residual32=None,
               out: P8 | None = None, out_relu: P8 | None = None, out32=None, out_s2d: P8 | None = None,
               residual16: P8 | None = None, partial=None, tiles: int = 0, count: int = 0, eps: float = 1e-5) -> None:
    """`partial` (with tiles, count, eps): fused InstanceNorm finalize - scale / shift are computed from the conv's per-tile sums
    inside the launch and written to `scale` / `shift`"""
    d = nv.NormApplyDesc()
    d.x = x.act()
    d.scale, d.shift, d.per_channel, d.act = ptr(scale), ptr(shift), int(per_channel), act
    d.residual32 = ptr(residual32)
    d.residual16 = act_or_null(residual16)
    d.partial, d.tiles, d.count, d.eps = ptr(partial), int(tiles), int(count), float(eps)
    d.out, d.out_relu, d.out32, d.out_s2d = act_or_null(out), act_or_null(out_relu), ptr(out32), act_or_null(out_s2d)
    d.dtype = dt
    check(lib().pbt_norm_apply(C.byref(d), stream_ptr()), "pbt_norm_apply")


def upsample2x(x: P8, out: P8, dt: int, scale=None, shift=None, act: int = ACT_NONE) -> None:
    """bilinear x2 (align_corners=True); with scale/shift the taps are normalised + activated on load"""
    a, b = x.act(), out.act()
    check(lib().pbt_upsample2x(C.byref(a), C.byref(b), ptr(scale), ptr(shift), act, dt, stream_ptr()), "pbt_upsample2x")


def upsample2x_bwd(gout: P8, dt: int, gin16: P8 | None = None, gin32=None) -> None:
    a, b = gout.act(), act_or_null(gin16)
    check(lib().pbt_upsample2x_bwd(C.byref(a), C.byref(b), ptr(gin32), dt, stream_ptr()), "pbt_upsample2x_bwd")


def norm_bwd(x: P8, dt: int, *, scale, shift, per_channel=False, act=ACT_NONE, ga: P8 | None = None, ga_is_s2d=False,
             gb16: P8 | None = None, gb32=None, sums, kmul, count: int, batch_mode=False, dx: P8,
             relu_mask_x=False, between=None) -> None:
    """reduce + apply of the (norm -> act) backward; `sums` must be zeroed by the caller.  `between(sums)` (two-launch
    path only) runs after the reduce and may rewrite the sums the apply will read."""
    d = nv.NormBwdDesc()
    d.x = x.act()
    d.scale, d.shift, d.per_channel, d.act = ptr(scale), ptr(shift), int(per_channel), act
    d.ga, d.ga_is_s2d, d.gb16, d.gb32 = act_or_null(ga), int(ga_is_s2d), act_or_null(gb16), ptr(gb32)
    d.sums, d.kmul, d.count, d.batch_mode = sums.data_ptr(), kmul.data_ptr(), count, int(batch_mode)
    d.dx, d.dtype, d.relu_mask_x = dx.act(), dt, int(relu_mask_x)
    if not batch_mode and x.h * x.w <= 16384 and x.n * (x.c // 8) >= 128:
        # patch-sized maps: one launch, a CTA per (image, plane) slice (second pass hits L2)
        check(lib().pbt_norm_bwd_fused(C.byref(d), stream_ptr()), "pbt_norm_bwd_fused")
        return
    check(lib().pbt_norm_bwd_reduce(C.byref(d), stream_ptr()), "pbt_norm_bwd_reduce")
    if between is not None:
        between(sums)
    check(lib().pbt_norm_bwd_apply(C.byref(d), stream_ptr()), "pbt_norm_bwd_apply")


def head_bwd(gy, y, s: P8, head_w, dt: int, *, gscale=None, head_tanh=True, dw, db, gs: P8, dbias_prev=None) -> None:
    a, b = s.act(), gs.act()
    check(lib().pbt_head_bwd(gy.data_ptr(), y.data_ptr(), C.byref(a), head_w.data_ptr(), ptr(gscale), int(head_tanh),
                             dw.data_ptr(), db.data_ptr(), C.byref(b), ptr(dbias_prev), dt, stream_ptr()), "pbt_head_bwd")


def channel_sum(g: P8, out, dt: int, inv_scale=None) -> None:
    a = g.act()
    check(lib().pbt_channel_sum(C.byref(a), out.data_ptr(), ptr(inv_scale), dt, stream_ptr()), "pbt_channel_sum")


# ----------------------------------------------------------------------------- layout / dtype
def nchw_to_p8(x: torch.Tensor, out: P8, dt: int) -> None:
    assert x.is_contiguous() and x.dtype in (torch.float32, torch.float16)
    n, c, h, w = x.shape
    a = out.act()
    check(lib().pbt_nchw_to_p8(x.data_ptr(), int(x.dtype == torch.float16), n, c, h, w, C.byref(a), dt, stream_ptr()),
          "pbt_nchw_to_p8")


def p8_to_nchw(x: P8, c: int, out: torch.Tensor, dt: int, mul: float = 1.0) -> None:
    a = x.act()
    check(lib().pbt_p8_to_nchw_f32(C.byref(a), c, out.data_ptr(), mul, dt, stream_ptr()), "pbt_p8_to_nchw_f32")


def p8f_to_nchw(x32: torch.Tensor, c: int, out: torch.Tensor) -> None:
    n, planes, h, w, _ = x32.shape
    check(lib().pbt_p8f_to_nchw_f32(x32.data_ptr(), n, planes * 8, c, h, w, out.data_ptr(), stream_ptr()),
          "pbt_p8f_to_nchw_f32")


def u8hwc_to_p8(img: torch.Tensor, out: P8, dt: int) -> None:
    n, h, w, c = img.shape
    a = out.act()
    check(lib().pbt_u8hwc_to_p8(img.data_ptr(), n, h, w, c, C.byref(a), dt, stream_ptr()), "pbt_u8hwc_to_p8")


def nchw_to_u8hwc(y: torch.Tensor, out: torch.Tensor) -> None:
    n, c, h, w = y.shape
    check(lib().pbt_nchw_to_u8hwc(y.data_ptr(), n, c, h, w, out.data_ptr(), stream_ptr()), "pbt_nchw_to_u8hwc")


def u8hwc_to_norm_chw(img: torch.Tensor, out: torch.Tensor) -> None:
    h, w, c = img.shape
    check(lib().pbt_u8hwc_to_norm_chw(img.data_ptr(), h, w, c, out.data_ptr(), stream_ptr()), "pbt_u8hwc_to_norm_chw")


def mask_dilate7(mask: torch.Tensor, out: torch.Tensor) -> None:
    h, w = mask.shape
    check(lib().pbt_mask_dilate7(mask.data_ptr(), h, w, out.data_ptr(), stream_ptr()), "pbt_mask_dilate7")


def zero_border(t: P8, valid_h: int, valid_w: int) -> None:
    """zero the pixels outside [0,valid_h) x [0,valid_w) of every plane (critic maps on a fixed grid)"""
    a = t.act()
    check(lib().pbt_zero_border(C.byref(a), valid_h, valid_w, stream_ptr()), "pbt_zero_border")


def p8s2d_to_nchw(x: P8, cpp: int, c: int, out: torch.Tensor, dt: int, mul=None) -> None:
    """space-to-depth P8 [n,4*cpp,h,w] -> NCHW fp32 [n,c,2h,2w] (times the device scalar `mul`)"""
    a = x.act()
    check(lib().pbt_p8s2d_to_nchw_f32(C.byref(a), cpp, c, out.data_ptr(), ptr(mul), dt, stream_ptr()), "pbt_p8s2d_to_nchw_f32")


def res_trunk_supported(channels: int, h: int, w: int) -> bool:
    return bool(lib().pbt_res_trunk_supported(channels, h, w))


def res_trunk_fwd(a, raw_a, hmid, raw_b, w_a, w_b, stats_a, stats_b, residual32: torch.Tensor, last16: P8, dt: int,
                  eps: float = 1e-5) -> None:
    """the whole residual trunk in one launch (csrc/res_trunk.cu).  a / raw_a / hmid / raw_b: lists of P8 [n, 128, h, w], one per
    block (a[0] is the input relu(r_0), a[1:] are written); stats_x: lists of (scale, shift) fp32 [n, 128] tables, written;
    residual32: r_0 on entry (fp32 [n, 16, h, w, 8]); last16: r_nb in 16 bit"""
    nb = len(raw_a)
    assert len(a) == nb and len(hmid) == nb and len(raw_b) == nb and len(w_a) == nb and len(w_b) == nb
    assert residual32.dtype == torch.float32 and residual32.is_contiguous()
    d = nv.ResTrunkDesc()
    d.n_blocks, d.dtype, d.eps = nb, dt, float(eps)
    for b in range(min(nb, nv.TRUNK_MAX_BLOCKS)):
        d.a[b], d.raw_a[b], d.hmid[b], d.raw_b[b] = a[b].act(), raw_a[b].act(), hmid[b].act(), raw_b[b].act()
        d.w_a[b], d.w_b[b] = w_a[b].data_ptr(), w_b[b].data_ptr()
        d.scale_a[b], d.shift_a[b] = stats_a[b][0].data_ptr(), stats_a[b][1].data_ptr()
        d.scale_b[b], d.shift_b[b] = stats_b[b][0].data_ptr(), stats_b[b][1].data_ptr()
    d.residual32 = residual32.data_ptr()
    d.last16 = last16.act()
    check(lib().pbt_res_trunk_fwd(C.byref(d), stream_ptr()), "pbt_res_trunk_fwd")


def feature_mse(f: P8, n_pairs: int, dt: int, *, g: P8 | None = None, grad_mul: float = 0.0, accumulate: bool = False,
                relu: bool = False, tap: bool = True, partial=None, counter=None, loss=None, loss_mul: float = 1.0) -> None:
    """one tensor of the perceptual loss: squared difference of the two halves of `f` into *loss, its gradient (and the
    ReLU mask) into `g`"""
    a, b = f.act(), act_or_null(g)
    flags = int(accumulate) | (int(relu) << 1) | (int(tap) << 2)
    check(lib().pbt_feature_mse(C.byref(a), n_pairs, grad_mul, flags, C.byref(b), ptr(partial), ptr(counter), ptr(loss),
                                loss_mul, dt, stream_ptr()), "pbt_feature_mse")


def maxpool2(x: P8, y: P8, dt: int) -> None:
    a, b = x.act(), y.act()
    check(lib().pbt_maxpool2(C.byref(a), C.byref(b), dt, stream_ptr()), "pbt_maxpool2")


def maxpool2_bwd(x: P8, dy: P8, dx: P8, dt: int) -> None:
    a, b, c = x.act(), dy.act(), dx.act()
    check(lib().pbt_maxpool2_bwd(C.byref(a), C.byref(b), C.byref(c), dt, stream_ptr()), "pbt_maxpool2_bwd")


def mask_erode7(mask_u8: torch.Tensor, out: torch.Tensor) -> None:
    """mask_u8 uint8 [n,h,w] (thresholded) -> out fp32 [n,h,w]: 1 where the whole 7x7 window is set (reference generator.py:327-351)"""
    n, h, w = mask_u8.shape
    assert mask_u8.is_contiguous() and out.is_contiguous() and out.dtype == torch.float32 and out.numel() == mask_u8.numel()
    check(lib().pbt_mask_erode7(mask_u8.data_ptr(), n, h, w, out.data_ptr(), stream_ptr()), "pbt_mask_erode7")


def composite_to_u8(y: torch.Tensor, out: torch.Tensor, frame_u8: torch.Tensor | None = None, mask: torch.Tensor | None = None) -> None:
    """y fp32 [n,3,h,w] -> out uint8 [n,h,w,3]; with `mask` (fp32 [n,h,w]) the frame's RGB shows through where the mask
    is 0: rgb*(1-m) + y*m (reference generator.py:562-563), then clamp / (x+1)*127.5 / round (:643-647)"""
    n, c3, h, w = y.shape
    assert c3 == 3 and y.is_contiguous() and y.dtype == torch.float32 and out.is_contiguous() and out.dtype == torch.uint8
    fc = 0
    if mask is not None:
        assert frame_u8 is not None and frame_u8.is_contiguous() and frame_u8.shape[:3] == (n, h, w)
        assert mask.is_contiguous() and mask.dtype == torch.float32 and mask.numel() == n * h * w
        fc = frame_u8.shape[3]
    check(lib().pbt_composite_to_u8(y.data_ptr(), ptr(frame_u8) if mask is not None else None, fc, ptr(mask), n, h, w,
                                    out.data_ptr(), stream_ptr()), "pbt_composite_to_u8")


def patch_gather(src_table: torch.Tensor, n_src: int, n_images: int, ch: int, img_hw: torch.Tensor, pos: torch.Tensor,
                 patch: int, outs, ch_off, ch_total) -> None:
    """src_table: int64 device tensor [n_src, n_images] of fp32 CHW image pointers."""
    n_patches = pos.shape[0]
    out_ptrs = (C.c_void_p * n_src)(*[o.data_ptr() for o in outs])
    offs = (C.c_int32 * n_src)(*ch_off)
    tots = (C.c_int32 * n_src)(*ch_total)
    check(lib().pbt_patch_gather(src_table.data_ptr(), n_src, n_images, ch, img_hw.data_ptr(), pos.data_ptr(), n_patches,
                                 patch, C.cast(out_ptrs, C.c_void_p), C.cast(offs, C.c_void_p), C.cast(tots, C.c_void_p),
                                 stream_ptr()), "pbt_patch_gather")


def absmax(g: torch.Tensor, out: torch.Tensor) -> None:
    check(lib().pbt_absmax_f32(g.data_ptr(), g.numel(), out.data_ptr(), stream_ptr()), "pbt_absmax_f32")


def make_grad_scale(amax: torch.Tensor, target: float, scale2: torch.Tensor, adjust=None) -> None:
    check(lib().pbt_make_grad_scale(amax.data_ptr(), target, scale2.data_ptr(), ptr(adjust), stream_ptr()), "pbt_make_grad_scale")


def grad_scale_feedback(probe: torch.Tensor, adjust: torch.Tensor) -> None:
    """overflow back-off of the fp16 gradient scale (see include/pbt.h); probe: fp32, contiguous"""
    check(lib().pbt_grad_scale_feedback(probe.data_ptr(), probe.numel(), adjust.data_ptr(), stream_ptr()), "pbt_grad_scale_feedback")
