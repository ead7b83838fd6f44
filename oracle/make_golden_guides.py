"""Reference-trained fixtures for the guide-channel configurations and the sampler's augmentation branch
(test infrastructure; needs /root/reference, run in the build container only):

    python oracle/make_golden_guides.py            # everything
    python oracle/make_golden_guides.py cin9 aug   # a subset

Every file below is an output of the UNMODIFIED reference modules (`src/models/generator.py::GeneratorJ`,
`src/data/dataset.py::StyleTransferDataset`, torch DataLoader / Adam / clip_grad_norm_), like oracle/make_golden.py:

  gen_cin9_trained.npz / gen_cin9_vectors.npz   config C3: RGB + two RGB-converted guide directories (the sample data has one
        guide directory, `tracking/`; it is used twice, as the survey's C3 recipe says).  100 reference G-only steps
        (batch 40 x 32x32) on miku_train_sorce, then (a) a C1-sized batch and (b) a batch of sixteen 80x80 patches (the C3
        patch size): train/eval outputs, loss, one-step gradients.
  gen_cin6_trained.npz / gen_cin6_vectors.npz   config C2: RGB + `tracking` on PlatinumChan_x0.5_train, 100 steps, then one
        whole 960x540 frame (input kept as uint8, output as float16 - 1e-3 of resolution on a 2e-2 tolerance) plus a
        patch batch with gradients.
  gen_cin5_trained.npz / gen_cin5_vectors.npz   config C5's channel count: RGB + 1-channel mask + 1-channel guide (first band
        of `tracking`).  The reference loader cannot produce 1-channel guides (it RGB-converts every directory), so the
        batches are assembled here from the reference dataset's own tensors; model, loss and optimiser are the reference's.
  sampler_aug_golden.npz                        StyleTransferDataset(augmentation_factor=2) on tests/golden/mini_dataset:
        both draws of every item (reference dataset.py:254,278), `already` / `channel_guide_aug` patches of two batches.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, REF)
_oc = types.ModuleType("omegaconf")
_oc.DictConfig = type("DictConfig", (dict,), {})
sys.modules.setdefault("omegaconf", _oc)

from torch.utils.data import DataLoader, Dataset  # noqa: E402

from src.data.dataset import StyleTransferDataset  # noqa: E402
from src.models.generator import GeneratorJ  # noqa: E402

STEPS = 100


def _cat(batch, names):
    return torch.cat([batch["pre"]] + [batch[f"channel_{n}"] for n in names], dim=1)   # lightning_model.py:211-221


def _train(G, batches, tag):
    """reference G-only step (lightning_model.py:239-250,260-292): L1*4, clip 0.5, Adam(4e-4, wd 1e-5)"""
    opt = torch.optim.Adam(G.parameters(), lr=4e-4, betas=(0.9, 0.999), weight_decay=1e-5)
    l1 = torch.nn.L1Loss()
    G.train()
    losses = []
    for it, (x, t) in enumerate(batches):
        opt.zero_grad()
        loss = l1(G(x), t) * 4.0
        loss.backward()
        torch.nn.utils.clip_grad_norm_(G.parameters(), 0.5)
        opt.step()
        losses.append(float(loss))
        if it % 20 == 0:
            print(f"[{tag}] reference step {it}: loss {float(loss):.4f}", flush=True)
        if it + 1 >= STEPS:
            break
    return losses


def _one_step(G, x, t, prefix, out):
    """train-mode forward, loss, one-step gradients, eval-mode forward of the same batch"""
    sd = {k: v.detach().clone() for k, v in G.state_dict().items()}
    G.train()
    G.zero_grad()
    y = G(x)
    loss = torch.nn.functional.l1_loss(y, t) * 4.0
    loss.backward()
    after = {k: v.detach().clone() for k, v in G.state_dict().items()}
    out[prefix + "x"], out[prefix + "target"] = x.numpy(), t.numpy()
    out[prefix + "y_train"], out[prefix + "loss"] = y.detach().numpy(), float(loss)
    out[prefix + "bn_rm_after"] = after["smoothers.2.running_mean"].numpy()
    out[prefix + "bn_rv_after"] = after["smoothers.2.running_var"].numpy()
    for k, p in G.named_parameters():
        g = p.grad.detach().reshape(-1)
        out[f"{prefix}gnorm_{k}"] = float(g.double().norm())
        out[f"{prefix}gmax_{k}"] = float(g.abs().max())
        if g.numel() <= 40000:
            out[f"{prefix}g_{k}"] = p.grad.detach().numpy().copy()
        else:                                    # strided sample of the big tensors (the oracle supplies the rest)
            out[f"{prefix}gs_{k}"] = g[:: max(1, g.numel() // 4096)].numpy().copy()
    G.load_state_dict(sd)
    G.eval()
    with torch.no_grad():
        out[prefix + "y_eval"] = G(x).numpy()


def _loader(ds, bs, seed):
    torch.manual_seed(seed)
    np.random.seed(seed)
    return DataLoader(ds, batch_size=bs, shuffle=True, num_workers=0)


def cin9():
    miku = os.path.join(REF, "test_dataset", "miku_train_sorce")
    guides = {"gauss": {"path": os.path.join(miku, "tracking"), "depth": 3},
              "flow": {"path": os.path.join(miku, "tracking"), "depth": 3}}
    ds = StyleTransferDataset(os.path.join(miku, "input"), os.path.join(miku, "output"), os.path.join(miku, "mask"), 32,
                              additional_channels=guides)
    torch.manual_seed(0)
    G = GeneratorJ(input_channels=9, use_bias=True)
    losses = _train(G, ((_cat(b, guides), b["post"]) for b in _loader(ds, 40, 0)), "cin9")
    np.savez(os.path.join(GOLD, "gen_cin9_trained.npz"), **{k: v.numpy() for k, v in G.state_dict().items()})
    out = {"train_losses": np.array(losses)}
    b = next(iter(_loader(ds, 40, 11)))
    _one_step(G, _cat(b, guides).clone(), b["post"].clone(), "", out)
    ds.patch_size = 80                                     # the C3 patch size (config/data/default.yaml)
    b = next(iter(_loader(ds, 16, 12)))
    _one_step(G, _cat(b, guides).clone(), b["post"].clone(), "p80_", out)
    np.savez_compressed(os.path.join(GOLD, "gen_cin9_vectors.npz"), **out)
    print("cin9 written; final training loss", losses[-1])


def cin6():
    pc = os.path.join(REF, "test_dataset", "PlatinumChan_x0.5_train")
    guides = {"tracking": {"path": os.path.join(pc, "tracking"), "depth": 3}}
    ds = StyleTransferDataset(os.path.join(pc, "input"), os.path.join(pc, "output"), os.path.join(pc, "mask"), 32,
                              additional_channels=guides)
    torch.manual_seed(1)
    G = GeneratorJ(input_channels=6, use_bias=True)
    losses = _train(G, ((_cat(b, guides), b["post"]) for b in _loader(ds, 40, 1)), "cin6")
    np.savez(os.path.join(GOLD, "gen_cin6_trained.npz"), **{k: v.numpy() for k, v in G.state_dict().items()})
    out = {"train_losses": np.array(losses)}
    b = next(iter(_loader(ds, 40, 13)))
    _one_step(G, _cat(b, guides).clone(), b["post"].clone(), "", out)
    # one whole frame (config C2: 960 rows x 540 columns): the resident tensors are (u8/255 - 0.5)/0.5, so the uint8 frame
    # is recovered exactly and the test rebuilds the fp32 input from it
    x = torch.cat([ds.images_pre[2], ds.additional_channel_data["tracking"][2]], 0).unsqueeze(0)
    u8 = torch.round((x[0] * 0.5 + 0.5) * 255.0).to(torch.uint8)
    assert torch.equal(((u8.float() / 255.0) - 0.5) / 0.5, x[0])
    G.eval()
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        y = G(x)
    out["frame_u8"] = u8.permute(1, 2, 0).contiguous().numpy()       # [H, W, 6]
    out["y_frame_f16"] = y[0].to(torch.float16).numpy()
    np.savez_compressed(os.path.join(GOLD, "gen_cin6_vectors.npz"), **out)
    print("cin6 written; final training loss", losses[-1], "frame", tuple(x.shape))


class _Five(Dataset):
    """reference dataset items re-assembled as 5-channel inputs: RGB + mask band + first band of the guide"""

    def __init__(self, ds):
        self.ds = ds

    def __len__(self):
        return len(self.ds)

    def __getitem__(self, idx):
        it = self.ds[idx]
        return torch.cat([it["pre"], it["channel_mask"][:1], it["channel_tracking"][:1]], 0), it["post"]


def cin5():
    miku = os.path.join(REF, "test_dataset", "miku_train_sorce")
    guides = {"mask": {"path": os.path.join(miku, "mask"), "depth": 3},
              "tracking": {"path": os.path.join(miku, "tracking"), "depth": 3}}
    ds = StyleTransferDataset(os.path.join(miku, "input"), os.path.join(miku, "output"), os.path.join(miku, "mask"), 32,
                              additional_channels=guides)
    five = _Five(ds)
    torch.manual_seed(2)
    G = GeneratorJ(input_channels=5, use_bias=True)
    losses = _train(G, _loader(five, 40, 2), "cin5")
    np.savez(os.path.join(GOLD, "gen_cin5_trained.npz"), **{k: v.numpy() for k, v in G.state_dict().items()})
    out = {"train_losses": np.array(losses)}
    x, t = next(iter(_loader(five, 40, 14)))
    _one_step(G, x.clone(), t.clone(), "", out)
    np.savez_compressed(os.path.join(GOLD, "gen_cin5_vectors.npz"), **out)
    print("cin5 written; final training loss", losses[-1])


class _Recorder(Dataset):
    def __init__(self, ds):
        self.ds, self.log = ds, []

    def __len__(self):
        return len(self.ds)

    def __getitem__(self, idx):
        item = self.ds[idx]
        (y, x), (yr, xr) = self.ds.last_patch_positions
        self.log.append((idx, idx % len(self.ds.images_pre), y, x, yr, xr))
        return item


def aug():
    m = lambda s: os.path.join(GOLD, "mini_dataset", s)  # noqa: E731
    ds = StyleTransferDataset(m("input"), m("output"), m("mask"), 32, augmentation_factor=2,
                              additional_channels={"guide": {"path": m("guide"), "depth": 3}})
    rec = _Recorder(ds)
    torch.manual_seed(321)
    np.random.seed(321)
    batches = []
    for bi, batch in enumerate(DataLoader(rec, batch_size=8, shuffle=True, num_workers=0)):
        if bi < 2:
            batches.append({k: v.numpy() for k, v in batch.items()})
        if bi >= 59:
            break
    np.savez_compressed(os.path.join(GOLD, "sampler_aug_golden.npz"), log=np.array(rec.log, dtype=np.int64), length=len(ds),
                        n_valid=np.array([len(v) for v in ds.valid_indices]),
                        **{f"b{bi}_{k}": v for bi, b in enumerate(batches) for k, v in b.items()})
    print("augmentation golden: draws", len(rec.log), "len", len(ds), "keys", sorted(batches[0]))


if __name__ == "__main__":
    todo = sys.argv[1:] or ["aug", "cin9", "cin6", "cin5"]
    torch.set_num_threads(os.cpu_count())
    for name in todo:
        {"aug": aug, "cin9": cin9, "cin6": cin6, "cin5": cin5}[name]()
