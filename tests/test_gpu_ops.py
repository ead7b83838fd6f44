"""GPU parity of every native op (through the C-ABI) against float64 restatements — see tests/gpu_checks.py."""
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import gpu_checks as gc
    from pbt_b200._native import ACT_LEAKY, ACT_RELU, BF16, FP16

CONV_CASES = [
    dict(cin=16, cout=16),
    dict(cin=64, cout=16, blk_c=32),
    dict(cin=48, cout=16, blk_c=32),
    dict(cin=16, cout=160),
    dict(cin=16, cout=16, dt=1),
    dict(cin=16, cout=32, kh=7, kw=7, pad_t=3, pad_l=3),
    dict(cin=32, cout=16, kh=2, kw=2, pad_t=1, pad_l=1, blk_c=32),
    dict(cin=32, cout=16, kh=2, kw=2, pad_t=0, pad_l=0, blk_c=32),
    dict(n=2, cin=32, cout=64, h=48, w=72, kh=7, kw=7, pad_t=3, pad_l=3, T=3, blk_c=32),
    dict(n=3, cin=128, cout=128, h=20, w=20, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32),
    dict(n=5, cin=128, cout=128, h=8, w=8, kh=3, kw=3, pad_t=1, pad_l=1, T=1, blk_c=64),
    dict(n=2, cin=32, cout=32, h=16, w=16, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32, in_off=16, in_extra=8, out_off=8,
         out_extra=16),
    dict(cin=176, cout=64, h=32, w=32, kh=7, kw=7, pad_t=3, pad_l=3, T=3, blk_c=32),
    dict(cin=256, cout=128, h=16, w=16, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32),
    dict(cin=32, cout=32, h=16, w=16, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32, bias=True, act=1),
    dict(cin=32, cout=32, h=16, w=16, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32, bias=True, act=2, affine=True,
         integer=False),
    dict(cin=32, cout=32, h=16, w=24, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32, mask=True, addend=True, out32=True),
    dict(cin=32, cout=32, h=16, w=24, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32, out32=True, store16=False),
    dict(n=2, cin=32, cout=64, h=24, w=40, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32, stats=True),
    dict(n=2, cin=32, cout=128, h=20, w=20, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32, stats=True, integer=False),
    dict(n=2, cin=64, cout=64, h=16, w=24, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32, bias=True, act=1, head=True,
         integer=False),
    dict(cin=128, cout=128, h=32, w=32, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32, integer=False, dt=1),
    # small-footprint configuration (four co-resident CTAs per SM, one epilogue warp per TMEM quadrant)
    dict(n=2, cin=64, cout=64, h=40, w=56, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32, bias=True, act=1, head=True,
         integer=False, cps=4),
    dict(n=2, cin=64, cout=64, h=24, w=40, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32, bias=True, act=1, affine=True, cps=4),
    dict(n=2, cin=16, cout=32, h=33, w=50, kh=7, kw=7, pad_t=3, pad_l=3, T=3, blk_c=16, stats=True, cps=4),
    dict(n=3, cin=128, cout=128, h=20, w=20, kh=3, kw=3, pad_t=1, pad_l=1, T=1, blk_c=32, stats=True, integer=False, cps=4),
    dict(n=2, cin=32, cout=16, h=16, w=24, kh=2, kw=2, pad_t=1, pad_l=1, T=3, blk_c=32, mask=True, addend=True, out32=True, cps=4),
    # CTA-pair configuration (cta_group::2, M = 256 across two CTAs, each staging half of the weight columns)
    dict(n=2, cin=64, cout=64, h=40, w=56, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32, pair=True),
    dict(n=1, cin=32, cout=32, h=16, w=40, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32, bias=True, act=1, pair=True),  # odd grid
    dict(n=2, cin=176, cout=64, h=48, w=72, kh=7, kw=7, pad_t=3, pad_l=3, T=3, blk_c=32, bias=True, act=1, integer=False, pair=True),
    dict(n=3, cin=128, cout=128, h=36, w=36, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32, stats=True, integer=False, pair=True),
    dict(n=2, cin=64, cout=64, h=24, w=48, kh=3, kw=3, pad_t=1, pad_l=1, T=3, blk_c=32, bias=True, act=1, head=True,
         integer=False, pair=True),
    dict(n=2, cin=64, cout=64, h=40, w=56, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=16, bias=True, act=1, pair=True, cps=4),
    dict(n=1, cin=48, cout=64, h=33, w=40, kh=7, kw=7, pad_t=3, pad_l=3, T=2, blk_c=16, stats=True, integer=False, pair=True,
         cps=4),
    # batch tiles: the T tiles of a CTA are the same spatial tile of T consecutive images (patch-sized maps)
    dict(n=5, cin=128, cout=128, h=20, w=20, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32, stats=True, integer=False, bt=True),
    dict(n=4, cin=64, cout=32, h=20, w=20, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32, stats=True, bt=True),
    dict(n=7, cin=32, cout=64, h=40, w=40, kh=3, kw=3, pad_t=1, pad_l=1, T=3, blk_c=32, bias=True, act=1, bt=True),
    dict(n=3, cin=128, cout=128, h=20, w=20, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32, mask=True, addend=True, out32=True,
         store16=False, bt=True),
    dict(n=6, cin=128, cout=256, h=20, w=20, kh=2, kw=2, pad_t=0, pad_l=0, T=2, blk_c=32, bt=True),
    dict(n=2, cin=64, cout=64, h=24, w=40, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32, bias=True, act=1, head=True, integer=False,
         bt=True),
    dict(n=170, cin=64, cout=128, h=20, w=20, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32, stats=True, integer=False, bt=True),
    # tap pairs (first layer, <= 8 real input channels): one K = 16 MMA = one channel plane at two adjacent taps
    dict(n=2, cin=16, cout=32, h=33, w=50, kh=7, kw=7, pad_t=3, pad_l=3, T=3, blk_c=16, stats=True, cps=4, tp=3),
    dict(n=1, cin=16, cout=32, h=40, w=24, kh=7, kw=7, pad_t=3, pad_l=3, T=2, blk_c=16, bias=True, act=2, tp=8, integer=False),
    dict(n=3, cin=16, cout=16, h=16, w=17, kh=3, kw=3, pad_t=1, pad_l=1, T=1, blk_c=16, tp=5),
    dict(n=2, cin=16, cout=32, h=300, w=500, kh=7, kw=7, pad_t=3, pad_l=3, T=3, blk_c=16, stats=True, cps=4, tp=6),
    # more units than one wave of CTAs: every CTA walks several units (persistent loop, barrier phases carried over)
    dict(n=2, cin=64, cout=128, h=272, w=480, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32, stats=True, integer=False),
    dict(n=3, cin=32, cout=64, h=200, w=330, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=16, bias=True, act=1, head=True,
         integer=False, cps=4),
    dict(n=2, cin=16, cout=32, h=300, w=500, kh=7, kw=7, pad_t=3, pad_l=3, T=3, blk_c=16, stats=True, cps=4),
    dict(n=2, cin=48, cout=64, h=250, w=410, kh=7, kw=7, pad_t=3, pad_l=3, T=3, blk_c=32, bias=True, act=1, integer=False,
         pair=True),
    dict(n=2, cin=32, cout=32, h=260, w=400, kh=3, kw=3, pad_t=1, pad_l=1, T=2, blk_c=32, mask=True, addend=True, out32=True,
         store16=False),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_conv_fwd(case):
    ok, err, msg = gc.check_conv(**case)
    assert ok, f"err={err} {msg}"


@pytest.mark.parametrize("kw", [dict(), dict(cin=192, cout=128, h=16, w=16), dict(n=3, cin=256, cout=128, h=10, w=12),
                                dict(cin=64, cout=64, h=24, w=24, T=3, dt=0), dict(cin=32, cout=16, h=5, w=7, T=1),
                                dict(n=2, cin=64, cout=64, h=136, w=200, T=2),
                                # raw leading channels normalised + activated inside the staged low-res tile
                                dict(cin=192, cout=128, h=16, w=16, raw_c=128), dict(n=3, cin=64, cout=32, h=10, w=12, T=1, raw_c=64),
                                dict(n=2, cin=96, cout=64, h=136, w=200, T=2, raw_c=32),
                                # upsample-on-load in the CTA-pair configuration (odd and even unit counts, persistent CTAs)
                                dict(pair=True), dict(n=1, cin=192, cout=128, h=18, w=20, pair=True),
                                dict(n=3, cin=256, cout=128, h=10, w=12, T=1, pair=True),
                                dict(n=2, cin=64, cout=128, h=136, w=200, T=2, pair=True)],
                         ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()) or "default")
def test_conv_with_upsample_on_load(kw):
    ok, err, msg = gc.check_conv_upsample(**kw)
    assert ok, msg


@pytest.mark.parametrize("kw", [dict(), dict(cpre=128, cin=48, cout=64, h=40, w=50), dict(cpre=256, cin=0, cout=256, h=20, w=20, blk_c=64),
                                dict(cpre=64, cin=0, cout=16, h=9, w=70, T=3, dt=0, act="leaky"),
                                dict(n=3, cpre=16, cin=16, cout=16, h=5, w=7, T=1, blk_c=16, act="none"),
                                dict(cpre=128, cin=48, cout=64, h=40, w=50, T=3, pair=True), dict(cpre=64, cin=0, cout=128, h=33, w=40, pair=True),
                                dict(cpre=64, cin=32, cout=64, h=40, w=50, blk_c=16, cps=4),
                                dict(cpre=128, cin=48, cout=64, h=40, w=50, blk_c=16, cps=4, pair=True),
                                dict(n=2, cpre=64, cin=32, cout=64, h=270, w=470, T=3, pair=True), dict(n=3, cpre=64, cin=0, cout=64, h=240, w=400)],
                         ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()) or "default")
def test_conv_with_norm_on_load(kw):
    ok, err, msg = gc.check_conv_norm_on_load(**kw)
    assert ok, msg


WGRAD_CASES = [
    dict(cin=16, cout=16),
    dict(cin=128, cout=16),
    dict(cin=16, cout=128),
    dict(cin=16, cout=16, kh=3, kw=3, pad_t=1, pad_l=1),
    dict(n=2, cin=128, cout=128, h=32, w=32, kh=3, kw=3, pad_t=1, pad_l=1),
    dict(cin=176, cout=64, h=32, w=32, kh=7, kw=7, pad_t=3, pad_l=3),
    dict(n=3, cin=256, cout=128, h=20, w=20, kh=2, kw=2, pad_t=1, pad_l=1),
    dict(cin=192, cout=128, h=24, w=24, kh=3, kw=3, pad_t=1, pad_l=1, dt=1, integer=False, inv_scale=0.25),
    # few input channels + wide kernel: the taps-in-M kernel (M = 8 horizontal taps x 8 channels)
    dict(n=2, cin=16, cout=32, h=40, w=56, kh=7, kw=7, pad_t=3, pad_l=3),
    dict(n=3, cin=8, cout=64, h=21, w=30, kh=7, kw=7, pad_t=3, pad_l=3, dt=1, integer=False),
    dict(n=2, cin=32, cout=32, h=24, w=24, kh=5, kw=5, pad_t=2, pad_l=2),
    dict(n=1, cin=16, cout=128, h=16, w=24, kh=4, kw=4, pad_t=1, pad_l=2),
]


@pytest.mark.parametrize("case", WGRAD_CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_conv_wgrad(case):
    ok, err, msg = gc.check_wgrad(**case)
    assert ok, f"err={err} {msg}"


@pytest.mark.parametrize("dt", [0, 1])
def test_layout_roundtrip(dt):
    assert gc.check_layout_roundtrip(dt=dt)[0]


def test_u8_and_mask_kernels():
    ok, _, msg = gc.check_u8()
    assert ok, msg


@pytest.mark.parametrize("kw", [dict(dt=0), dict(dt=1), dict(dt=1, batch_mode=True)])
def test_norm_finalize_apply(kw):
    ok, err, msg = gc.check_norm(**kw)
    assert ok, msg


@pytest.mark.parametrize("dt", [0, 1])
def test_upsample_and_transpose(dt):
    ok, err, msg = gc.check_upsample(dt=dt)
    assert ok, msg


@pytest.mark.parametrize("kw", [dict(dt=1), dict(dt=1, s2d=True), dict(dt=1, batch_mode=True), dict(dt=0),
                                dict(dt=1, n=40, c=32), dict(dt=1, n=40, c=32, s2d=True), dict(dt=0, n=33, c=32)],
                         ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_norm_backward(kw):
    ok, err, msg = gc.check_norm_bwd(**kw)
    assert ok, msg


@pytest.mark.parametrize("dt", [0, 1])
def test_head_backward(dt):
    ok, err, msg = gc.check_head_bwd(dt=dt)
    assert ok, msg


def test_grad_scale():
    ok, _, msg = gc.check_grad_scale()
    assert ok, msg


@pytest.mark.parametrize("kw", [dict(patch=32), dict(patch=80, n_patches=9), dict(patch=7, n_patches=9)])
def test_patch_gather_bit_exact(kw):
    assert gc.check_gather(**kw)[0]


def test_patch_gather_empty_batch():
    from pbt_b200 import ops
    t = torch.zeros((1, 1), dtype=torch.int64, device="cuda")
    hw = torch.zeros((1, 2), dtype=torch.int32, device="cuda")
    pos = torch.zeros((0, 3), dtype=torch.int32, device="cuda")
    out = torch.zeros((0, 3, 8, 8), device="cuda")
    ops.patch_gather(t, 1, 1, 3, hw, pos, 8, [out], [0], [3])  # must be a no-op, not an error


def test_bad_arguments_raise():
    from pbt_b200 import ops
    from pbt_b200._native import P8
    x = P8.empty(1, 24, 16, 16, 0)  # cin not a multiple of 16
    w = torch.zeros(16 * 24 * 2, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="cin must be a multiple of 16"):
        ops.conv_fwd(x, w, 16, 1, 1, 0, 0, 0, blk_c=16, tiles_per_cta=1, out=P8.empty(1, 16, 16, 16, 0))


def test_bad_arguments_of_the_round2_entry_points_raise():
    """shape / mode violations come back as an error code that the wrapper raises - never a silent wrong answer"""
    import ctypes as C
    from pbt_b200 import ops
    from pbt_b200._native import P8, lib, stream_ptr
    x = P8.empty(2, 16, 20, 20, 1)
    w = torch.zeros(49 * 16 * 32, device="cuda", dtype=torch.float16)
    out = P8.empty(2, 32, 20, 20, 1)
    with pytest.raises(RuntimeError, match="batch_tiles"):          # batch tiles need >= 2 tiles per CTA
        ops.conv_fwd(x, w, 32, 3, 3, 1, 1, 1, blk_c=16, tiles_per_cta=1, out=out, batch_tiles=True)
    with pytest.raises(RuntimeError, match="tap_pairs"):            # tap pairs and batch tiles exclude each other
        ops.conv_fwd(x, w, 32, 7, 7, 3, 3, 1, blk_c=16, tiles_per_cta=2, out=out, batch_tiles=True, tap_pairs=True)
    with pytest.raises(RuntimeError, match="valid window"):
        ops.conv_fwd(x, w, 32, 3, 3, 1, 1, 1, blk_c=16, tiles_per_cta=2, out=out, valid_hw=(21, 20))
    with pytest.raises(RuntimeError, match="up_raw_channels"):      # raw leading channels need their scale / shift tables
        ops.conv_fwd(P8.empty(2, 32, 10, 10, 1), torch.zeros(9 * 32 * 32, device="cuda", dtype=torch.float16), 32, 3, 3, 1, 1, 1,
                     blk_c=32, tiles_per_cta=2, out=out, upsample2x=True, up_raw_channels=32)
    a = x.act()
    assert lib().pbt_zero_border(C.byref(a), 21, 20, stream_ptr()) != 0
    r = P8.empty(2, 16, 20, 20, 1)
    with pytest.raises(RuntimeError, match="residual16"):           # both residual kinds at once
        ops.norm_apply(x, 1, residual32=torch.zeros((2, 2, 20, 20, 8), device="cuda"), residual16=r, out=P8.empty(2, 16, 20, 20, 1))
    with pytest.raises(RuntimeError, match="fused finalize"):       # > 64 tiles per image cannot be finalised inside the launch
        ops.norm_apply(x, 1, scale=torch.zeros((2, 16), device="cuda"), shift=torch.zeros((2, 16), device="cuda"),
                       partial=torch.zeros((2, 100, 2, 16), device="cuda"), tiles=100, count=400, out=P8.empty(2, 16, 20, 20, 1))
    y = torch.zeros((1, 3, 8, 8), device="cuda")
    with pytest.raises(AssertionError):                             # a mask without the uint8 frame
        ops.composite_to_u8(y, torch.zeros((1, 8, 8, 3), dtype=torch.uint8, device="cuda"), None, torch.ones((1, 8, 8), device="cuda"))
    torch.cuda.synchronize()


@pytest.mark.parametrize("kw", [dict(), dict(clip=None), dict(wd=0.0, steps=3)], ids=["clip", "noclip", "nowd"])
def test_fused_clip_adam_matches_torch(kw):
    ok, err, msg = gc.check_fused_clip_adam(**kw)
    assert ok, msg


def test_fused_l1_loss_matches_torch():
    ok, err, msg = gc.check_fused_l1()
    assert ok, msg


def test_guard_bands_catch_an_out_of_bounds_write(monkeypatch):
    """the PBT_GUARD debug mode itself: a write one element past a guarded allocation is reported"""
    import pbt_b200._native as nv
    monkeypatch.setattr(nv, "GUARD_ELEMS", 64)
    t = nv.P8.empty(1, 8, 4, 4, nv.FP16)
    assert nv.check_guards() >= 1
    flat = [e[1] for e in nv._guards if e[0]() is t.t][0]
    flat[64 + t.t.numel()] = 1.0
    with pytest.raises(RuntimeError, match="out of bounds"):
        nv.check_guards()
    flat.view(torch.int16)[64 + t.t.numel()] = nv._SENTINEL
    assert nv.check_guards() >= 1
