"""ctypes binding of libpbt.so (C-ABI declared in include/pbt.h).

There is deliberately no fallback: if the library is missing, or a call returns
a non-zero status, a RuntimeError is raised.  Only raw pointers, sizes and the
current CUDA stream handle cross this boundary.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libpbt.so")

BF16, FP16 = 0, 1
ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2

_TORCH_DTYPE = {BF16: torch.bfloat16, FP16: torch.float16}


def torch_dtype(dt: int) -> torch.dtype:
    return _TORCH_DTYPE[dt]


class Act(C.Structure):
    """pbt_act_t — 16-bit P8 activation view."""

    _fields_ = [("ptr", C.c_void_p), ("n", C.c_int32), ("c", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
                ("img_stride", C.c_int64)]


class ConvDesc(C.Structure):
    _fields_ = [
        ("inp", Act), ("wpack", C.c_void_p), ("cout", C.c_int32),
        ("kh", C.c_int32), ("kw", C.c_int32), ("pad_t", C.c_int32), ("pad_l", C.c_int32),
        ("blk_c", C.c_int32), ("tiles_per_cta", C.c_int32), ("dtype", C.c_int32),
        ("bias", C.c_void_p), ("act", C.c_int32), ("post_scale", C.c_void_p), ("post_shift", C.c_void_p),
        ("mask", Act), ("addend32", C.c_void_p), ("out32", C.c_void_p), ("out", Act),
        ("stats_partial", C.c_void_p), ("head_w", C.c_void_p), ("head_b", C.c_void_p), ("head_out", C.c_void_p),
        ("head_tanh", C.c_int32), ("upsample2x", C.c_int32), ("pre", Act), ("pre_scale", C.c_void_p), ("pre_shift", C.c_void_p),
        ("pre_act", C.c_int32), ("ctas_per_sm", C.c_int32), ("cta_pair", C.c_int32), ("concurrent", C.c_int32), ("batch_tiles", C.c_int32), ("up_raw_channels", C.c_int32), ("tap_pairs", C.c_int32), ("valid_h", C.c_int32), ("valid_w", C.c_int32), ("debug_flags", C.c_int32), ("debug_buf", C.c_void_p),
    ]


class WgradDesc(C.Structure):
    _fields_ = [
        ("x", Act), ("dy", Act), ("kh", C.c_int32), ("kw", C.c_int32), ("pad_t", C.c_int32), ("pad_l", C.c_int32),
        ("dtype", C.c_int32), ("dw", C.c_void_p), ("inv_scale", C.c_void_p), ("debug_flags", C.c_int32),
    ]


class NormApplyDesc(C.Structure):
    _fields_ = [
        ("x", Act), ("scale", C.c_void_p), ("shift", C.c_void_p), ("per_channel", C.c_int32), ("act", C.c_int32),
        ("residual32", C.c_void_p), ("out", Act), ("out_relu", Act), ("out32", C.c_void_p), ("out_s2d", Act),
        ("dtype", C.c_int32), ("residual16", Act), ("partial", C.c_void_p), ("tiles", C.c_int32), ("count", C.c_int64), ("eps", C.c_float),
    ]


class PackJob(C.Structure):
    """pbt_pack_job_t (64 bytes)"""

    _fields_ = [("w", C.c_void_p), ("dst", C.c_void_p), ("co", C.c_int32), ("ci", C.c_int32), ("kh", C.c_int32),
                ("kw", C.c_int32), ("mode", C.c_int32), ("k_pad", C.c_int32), ("n_out", C.c_int32), ("n_keep", C.c_int32),
                ("blk_c", C.c_int32), ("dtype", C.c_int32), ("src_is_half", C.c_int32), ("reserved", C.c_int32)]


class OptimJob(C.Structure):
    """pbt_optim_job_t (40 bytes)"""

    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("count", C.c_int64)]


class NormBwdDesc(C.Structure):
    _fields_ = [
        ("x", Act), ("scale", C.c_void_p), ("shift", C.c_void_p), ("per_channel", C.c_int32), ("act", C.c_int32),
        ("ga", Act), ("ga_is_s2d", C.c_int32), ("gb16", Act), ("gb32", C.c_void_p), ("sums", C.c_void_p),
        ("kmul", C.c_void_p), ("count", C.c_int64), ("batch_mode", C.c_int32), ("dx", Act), ("dtype", C.c_int32),
        ("relu_mask_x", C.c_int32),
    ]


TRUNK_MAX_BLOCKS = 16


class ResTrunkDesc(C.Structure):
    """pbt_res_trunk_desc_t"""

    _fields_ = [("n_blocks", C.c_int32), ("dtype", C.c_int32), ("eps", C.c_float), ("reserved", C.c_int32),
                ("a", Act * TRUNK_MAX_BLOCKS), ("raw_a", Act * TRUNK_MAX_BLOCKS), ("hmid", Act * TRUNK_MAX_BLOCKS),
                ("raw_b", Act * TRUNK_MAX_BLOCKS), ("w_a", C.c_void_p * TRUNK_MAX_BLOCKS), ("w_b", C.c_void_p * TRUNK_MAX_BLOCKS),
                ("scale_a", C.c_void_p * TRUNK_MAX_BLOCKS), ("shift_a", C.c_void_p * TRUNK_MAX_BLOCKS),
                ("scale_b", C.c_void_p * TRUNK_MAX_BLOCKS), ("shift_b", C.c_void_p * TRUNK_MAX_BLOCKS),
                ("residual32", C.c_void_p), ("last16", Act)]


_lib = None


def lib() -> C.CDLL:
    """Load libpbt.so once; fail loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"native library {LIB_PATH} is missing — run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU or PyTorch fallback for the hot path)")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    sig = {
        "pbt_abi_version": (C.c_int, []),
        "pbt_error_string": (C.c_char_p, [C.c_int]),
        "pbt_last_cuda_error": (C.c_char_p, []),
        "pbt_conv_num_tiles": (C.c_int, [i32, i32, i32]),
        "pbt_conv_fwd": (C.c_int, [C.POINTER(ConvDesc), vp]),
        "pbt_conv_wgrad": (C.c_int, [C.POINTER(WgradDesc), vp]),
        "pbt_norm_finalize": (C.c_int, [vp, i32, i32, i32, i64, f32, i32, vp, vp, vp, vp, f32, vp, vp, vp, vp, vp]),
        "pbt_norm_apply": (C.c_int, [C.POINTER(NormApplyDesc), vp]),
        "pbt_upsample2x": (C.c_int, [C.POINTER(Act), C.POINTER(Act), vp, vp, i32, i32, vp]),
        "pbt_upsample2x_bwd": (C.c_int, [C.POINTER(Act), C.POINTER(Act), vp, i32, vp]),
        "pbt_norm_bwd_reduce": (C.c_int, [C.POINTER(NormBwdDesc), vp]),
        "pbt_norm_bwd_apply": (C.c_int, [C.POINTER(NormBwdDesc), vp]),
        "pbt_norm_bwd_fused": (C.c_int, [C.POINTER(NormBwdDesc), vp]),
        "pbt_l1_loss_fwd_bwd": (C.c_int, [vp, vp, i64, f32, vp, vp, vp]),
        "pbt_tile_gather": (C.c_int, [vp, i32, i32, i32, vp, i32, i32, vp, vp]),
        "pbt_tile_blend": (C.c_int, [vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp]),
        "pbt_tile_finish": (C.c_int, [vp, vp, vp, vp, i32, i32, vp, vp]),
        "pbt_clip_adam_step": (C.c_int, [vp, i32, C.c_int64, vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                               C.c_double, vp, vp]),
        "pbt_head_bwd": (C.c_int, [vp, vp, C.POINTER(Act), vp, vp, i32, vp, vp, C.POINTER(Act), vp, i32, vp]),
        "pbt_channel_sum": (C.c_int, [C.POINTER(Act), vp, vp, i32, vp]),
        "pbt_nchw_to_p8": (C.c_int, [vp, i32, i32, i32, i32, i32, C.POINTER(Act), i32, vp]),
        "pbt_p8_to_nchw_f32": (C.c_int, [C.POINTER(Act), i32, vp, f32, i32, vp]),
        "pbt_p8f_to_nchw_f32": (C.c_int, [vp, i32, i32, i32, i32, i32, vp, vp]),
        "pbt_u8hwc_to_p8": (C.c_int, [vp, i32, i32, i32, i32, C.POINTER(Act), i32, vp]),
        "pbt_nchw_to_u8hwc": (C.c_int, [vp, i32, i32, i32, i32, vp, vp]),
        "pbt_u8hwc_to_norm_chw": (C.c_int, [vp, i32, i32, i32, vp, vp]),
        "pbt_patch_gather": (C.c_int, [vp, i32, i32, i32, vp, vp, i32, i32, vp, vp, vp, vp]),
        "pbt_mask_dilate7": (C.c_int, [vp, i32, i32, vp, vp]),
        "pbt_mask_erode7": (C.c_int, [vp, i32, i32, i32, vp, vp]),
        "pbt_zero_border": (C.c_int, [C.POINTER(Act), i32, i32, vp]),
        "pbt_p8s2d_to_nchw_f32": (C.c_int, [C.POINTER(Act), i32, i32, vp, vp, i32, vp]),
        "pbt_composite_to_u8": (C.c_int, [vp, vp, i32, vp, i32, i32, i32, vp, vp]),
        "pbt_res_trunk_supported": (C.c_int, [i32, i32, i32]),
        "pbt_res_trunk_fwd": (C.c_int, [C.POINTER(ResTrunkDesc), vp]),
        "pbt_feature_mse": (C.c_int, [C.POINTER(Act), i32, f32, i32, C.POINTER(Act), vp, vp, vp, f32, i32, vp]),
        "pbt_maxpool2": (C.c_int, [C.POINTER(Act), C.POINTER(Act), i32, vp]),
        "pbt_maxpool2_bwd": (C.c_int, [C.POINTER(Act), C.POINTER(Act), C.POINTER(Act), i32, vp]),
        "pbt_absmax_f32": (C.c_int, [vp, i64, vp, vp]),
        "pbt_make_grad_scale": (C.c_int, [vp, f32, vp, vp, vp]),
        "pbt_grad_scale_feedback": (C.c_int, [vp, i64, vp, vp]),
        "pbt_ostree_reset": (None, [vp, i32]),
        "pbt_ostree_take": (i32, [vp, i32, i32]),
        "pbt_pack_weights": (C.c_int, [vp, i32, i64, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if L.pbt_abi_version() != 2:
        raise RuntimeError("libpbt.so ABI version mismatch")
    _lib = L
    return L


EXPORTED_SYMBOLS = [
    "pbt_abi_version", "pbt_error_string", "pbt_last_cuda_error", "pbt_conv_num_tiles", "pbt_conv_fwd",
    "pbt_conv_wgrad", "pbt_norm_finalize", "pbt_norm_apply", "pbt_upsample2x", "pbt_upsample2x_bwd",
    "pbt_norm_bwd_reduce", "pbt_norm_bwd_apply", "pbt_norm_bwd_fused", "pbt_clip_adam_step", "pbt_l1_loss_fwd_bwd", "pbt_tile_gather", "pbt_tile_blend", "pbt_tile_finish", "pbt_head_bwd", "pbt_channel_sum", "pbt_nchw_to_p8",
    "pbt_p8_to_nchw_f32", "pbt_p8f_to_nchw_f32", "pbt_u8hwc_to_p8", "pbt_nchw_to_u8hwc", "pbt_u8hwc_to_norm_chw",
    "pbt_patch_gather", "pbt_mask_dilate7", "pbt_mask_erode7", "pbt_composite_to_u8", "pbt_zero_border", "pbt_p8s2d_to_nchw_f32", "pbt_feature_mse", "pbt_maxpool2", "pbt_maxpool2_bwd", "pbt_res_trunk_supported", "pbt_res_trunk_fwd", "pbt_absmax_f32", "pbt_make_grad_scale", "pbt_grad_scale_feedback", "pbt_ostree_reset",
    "pbt_ostree_take", "pbt_pack_weights",
]


LAUNCHES = [0]  # native kernel launches issued through this binding (bench.py reports it)


def check(status: int, what: str) -> None:
    LAUNCHES[0] += 1
    if status != 0:
        L = lib()
        raise RuntimeError(f"{what} failed: {L.pbt_error_string(status).decode()} — {L.pbt_last_cuda_error().decode()}")


def stream_ptr() -> int:
    """Handle of torch's current CUDA stream (all native work is enqueued there)."""
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


class P8:
    """A 16-bit activation tensor in the P8 layout [n][c/8][h][w][8] plus channel-range views."""

    __slots__ = ("t", "n", "c", "h", "w", "plane0", "planes_total")

    def __init__(self, t: torch.Tensor, c: int | None = None, plane0: int = 0):
        assert t.dim() == 5 and t.shape[4] == 8 and t.is_contiguous()
        self.t = t
        self.n, self.planes_total, self.h, self.w = t.shape[0], t.shape[1], t.shape[2], t.shape[3]
        self.plane0 = plane0
        self.c = (self.planes_total - plane0) * 8 if c is None else c
        assert self.c % 8 == 0 and plane0 + self.c // 8 <= self.planes_total

    @staticmethod
    def empty(n, c, h, w, dt, device="cuda", zero=False) -> "P8":
        if GUARD_ELEMS:
            return P8(_guarded((n, c // 8, h, w, 8), torch_dtype(dt), device, zero))
        f = torch.zeros if zero else torch.empty
        return P8(f((n, c // 8, h, w, 8), dtype=torch_dtype(dt), device=device))

    def view(self, c0: int, c: int) -> "P8":
        """channels [c0, c0+c) of this tensor (multiples of 8)"""
        assert c0 % 8 == 0 and c % 8 == 0
        return P8(self.t, c, self.plane0 + c0 // 8)

    def act(self) -> Act:
        off = self.plane0 * self.h * self.w * 8 * self.t.element_size()
        return Act(self.t.data_ptr() + off, self.n, self.c, self.h, self.w, self.planes_total * self.h * self.w * 8)

    def to_nchw(self) -> torch.Tensor:
        """fp32 NCHW copy via plain torch ops (tests / debugging only)."""
        v = self.t[:, self.plane0:self.plane0 + self.c // 8]
        return v.permute(0, 1, 4, 2, 3).reshape(self.n, self.c, self.h, self.w).float()

    @staticmethod
    def from_nchw(x: torch.Tensor, dt: int, c_pad: int | None = None) -> "P8":
        """torch-op construction (tests / debugging only)."""
        n, c, h, w = x.shape
        cp = c_pad or ((c + 7) // 8 * 8)
        xp = torch.zeros((n, cp, h, w), dtype=torch_dtype(dt), device=x.device)
        xp[:, :c] = x.to(torch_dtype(dt))
        return P8(xp.reshape(n, cp // 8, 8, h, w).permute(0, 1, 3, 4, 2).contiguous())


# ---- debug aid (PBT_GUARD=<elements>, the GPU test-suite runs once with it): every P8.empty allocation sits between two bands of a
# sentinel bit pattern.  An out-of-bounds WRITE of any kernel changes a band (check_guards raises); an out-of-bounds READ pulls
# 0x5A5A = 203.25 (fp16) / 1.5e16 (bf16) into the arithmetic and fails the parity assertions.  compute-sanitizer is not available
# on the GPU pool, this is the bounds check that is.
GUARD_ELEMS = int(os.environ.get("PBT_GUARD", "0")) // 8 * 8
_SENTINEL = 0x5A5A
_guards: list = []


def _guarded(shape, dtype, device, zero):
    import weakref
    numel = 1
    for d in shape:
        numel *= d
    flat = torch.empty(numel + 2 * GUARD_ELEMS, dtype=dtype, device=device)
    bits = flat.view(torch.int16)
    bits[:GUARD_ELEMS] = _SENTINEL
    bits[GUARD_ELEMS + numel:] = _SENTINEL
    body = flat[GUARD_ELEMS:GUARD_ELEMS + numel]
    if zero:
        body.zero_()
    else:
        body.view(torch.int16).fill_(_SENTINEL)           # "uninitialised" memory is poison too
    t = body.view(shape)
    _guards[:] = [e for e in _guards if e[0]() is not None]
    _guards.append((weakref.ref(t), flat, numel, GUARD_ELEMS))
    return t


def check_guards() -> int:
    """verify the sentinel bands of every live guarded allocation; returns how many were checked"""
    _guards[:] = [e for e in _guards if e[0]() is not None]
    bad = []
    for _, flat, numel, band in _guards:
        bits = flat.view(torch.int16)
        ok = (bits[:band] == _SENTINEL).all() & (bits[band + numel:] == _SENTINEL).all()
        bad.append(~ok)
    if bad and bool(torch.stack(bad).any()):
        which = [i for i, b in enumerate(bad) if bool(b)]
        raise RuntimeError(f"PBT_GUARD: {len(which)} of {len(bad)} guarded P8 allocations were written out of bounds "
                           f"(sizes {[_guards[i][2] for i in which[:8]]})")
    return len(bad)


NULL_ACT = Act(None, 0, 0, 0, 0, 0)


def act_or_null(p: "P8 | None") -> Act:
    return NULL_ACT if p is None else p.act()
