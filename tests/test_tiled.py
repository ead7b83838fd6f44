"""Tiled inference mode (reference generator.py:327-565): oracle vs fixtures produced by the unmodified reference methods
(CPU), and the native path vs the same fixtures (GPU)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "tiled_golden.npz"))


@pytest.fixture(scope="module")
def trained_sd():
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLD, "gen_c3_trained.npz")).items()}


def test_oracle_mask_and_windows_match_reference(gold):
    from oracle import tiled_oracle as to
    mask = to.process_mask(torch.from_numpy(gold["mask_raw"])[None])
    assert torch.equal(mask, torch.from_numpy(gold["mask"]))
    for patch, overlap in ((32, 30.0), (80, 30.0), (48, 50.0)):
        pos = to.valid_patch_positions(mask.unsqueeze(0), patch, overlap)
        assert np.array_equal(np.asarray(pos, dtype=np.int32), gold[f"pos_p{patch}"]), patch


@pytest.mark.parametrize("patch,overlap,key", [(32, 30.0, "y_p32"), (80, 30.0, "y_p80"), (48, 50.0, "y_p48"), (32, 30.0, "y_p32_nomask")])
def test_oracle_tiled_output_matches_reference(gold, trained_sd, patch, overlap, key):
    from oracle import generator_oracle as go
    from oracle import tiled_oracle as to
    sd = {k: v.float() for k, v in trained_sd.items()}
    frame = torch.from_numpy(gold["frame"])
    mask = None if key.endswith("nomask") else torch.from_numpy(gold["mask"]).unsqueeze(0)
    y, _ = to.process_large_image(lambda p: go.generator_forward(sd, p), frame, mask, patch, overlap)
    err = (y - torch.from_numpy(gold[key])).abs().max().item()
    assert err < 5e-5, err


@pytest.mark.gpu
@pytest.mark.parametrize("patch,overlap,key", [(32, 30.0, "y_p32"), (80, 30.0, "y_p80"), (48, 50.0, "y_p48"), (32, 30.0, "y_p32_nomask")])
def test_native_tiled_inference_matches_reference(gold, trained_sd, patch, overlap, key):
    from pbt_b200 import tiled
    from pbt_b200.generator import GeneratorJ
    g = GeneratorJ(input_channels=3, use_bias=True)
    g.load_state_dict(trained_sd, strict=True)
    g = g.cuda().eval()
    frame = torch.from_numpy(gold["frame"]).cuda()
    mask = None
    if not key.endswith("nomask"):
        mask = tiled.process_mask(torch.from_numpy(gold["mask_raw"])[None].cuda())
        assert torch.equal(mask.cpu(), torch.from_numpy(gold["mask"]))
        mask = mask.unsqueeze(0)
    y, boxes = tiled.process_large_image(g, frame, mask, patch, overlap, tile_batch=16, return_windows=True)
    if mask is not None:
        assert np.array_equal(np.asarray(boxes, dtype=np.int32), gold[f"pos_p{patch}"])
    ref = torch.from_numpy(gold[key])
    err = (y.cpu() - ref).abs().max().item()
    mse = ((y.cpu() - ref) ** 2).mean().item()
    psnr = 10 * np.log10(4.0 / max(mse, 1e-20))
    print(f"tiled {key}: max_abs={err:.5f} psnr={psnr:.1f} dB windows={len(boxes)}")
    assert err <= 2e-2 and psnr >= 40.0, (err, psnr)    # north-star tolerance for generator outputs
