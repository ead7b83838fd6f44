"""Generate tests/golden/* by running the UNMODIFIED reference modules (needs /root/reference; run in the build
container only):   python oracle/make_golden.py

The reference has no tests or golden vectors of its own, so every fixture is an output
of the reference itself:
  mini_dataset/            synthetic keyframes written by this script (inputs, not reference outputs)
  sampler_golden.npz       (idx, img, y, x) stream + first patches from reference StyleTransferDataset + DataLoader
  cutpatch_golden.npz      reference _cut_patch at the edge cases SURVEY.md section 8(c) lists
  gen_c3_trained.npz       reference GeneratorJ(3, use_bias=True) state_dict after 100 reference G-only steps
  gen_c3_vectors.npz       inputs, train/eval outputs, loss and one-step gradients of that model
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
# the reference dataset imports omegaconf only for one isinstance check; a one-class stub makes the
# unmodified file importable (omegaconf is not installed in this image)
_oc = types.ModuleType("omegaconf")
_oc.DictConfig = type("DictConfig", (dict,), {})
sys.modules.setdefault("omegaconf", _oc)

from PIL import Image  # noqa: E402
from torch.utils.data import DataLoader, Dataset  # noqa: E402

from src.data.dataset import StyleTransferDataset  # noqa: E402
from src.models.generator import GeneratorJ  # noqa: E402


def make_mini_dataset(root):
    rng = np.random.RandomState(7)
    H, W = 96, 128
    for sub in ("input", "output", "mask", "guide"):
        os.makedirs(os.path.join(root, sub), exist_ok=True)
    yy, xx = np.mgrid[0:H, 0:W]
    for i in range(5):
        base = rng.randint(0, 256, (H // 8, W // 8, 3)).astype(np.uint8)
        img = np.asarray(Image.fromarray(base).resize((W, H), Image.BILINEAR)).astype(np.int32)
        img = np.clip(img + rng.randint(-8, 9, img.shape), 0, 255).astype(np.uint8)
        Image.fromarray(img).save(os.path.join(root, "input", f"{i:03d}.png"))
        Image.fromarray(255 - img[:, ::-1].copy()).save(os.path.join(root, "output", f"{i:03d}.png"))
        g = np.clip(img[::-1].astype(np.int32) // 2 + 60, 0, 255).astype(np.uint8)
        Image.fromarray(g.copy()).save(os.path.join(root, "guide", f"{i:03d}.png"))
        if i == 3:  # tiny blob: the without-replacement list runs empty and is refilled within the recording
            m = ((yy - 40) ** 2 + (xx - 50) ** 2 <= 2).astype(np.uint8) * 255
        elif i == 4:  # touches the border: exercises _cut_patch clamping
            m = ((yy < 6) | (xx > W - 5)).astype(np.uint8) * 255
        else:
            cy, cx, ry, rx = rng.randint(30, 66), rng.randint(30, 98), rng.randint(10, 30), rng.randint(10, 40)
            m = ((((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2) <= 1).astype(np.uint8) * 200
        Image.fromarray(m, mode="L").save(os.path.join(root, "mask", f"{i:03d}.png"))


class _Recorder(Dataset):
    def __init__(self, ds):
        self.ds, self.log = ds, []

    def __len__(self):
        return len(self.ds)

    def __getitem__(self, idx):
        item = self.ds[idx]
        y, x = self.ds.last_patch_positions[0]
        self.log.append((idx, idx % len(self.ds.images_pre), y, x))
        return item


def sampler_golden():
    root = os.path.join(GOLD, "mini_dataset")
    make_mini_dataset(root)
    ds = StyleTransferDataset(os.path.join(root, "input"), os.path.join(root, "output"), os.path.join(root, "mask"), 32,
                              additional_channels={"guide": {"path": os.path.join(root, "guide"), "depth": 3}})
    rec = _Recorder(ds)
    torch.manual_seed(123)
    np.random.seed(123)
    dl = DataLoader(rec, batch_size=8, shuffle=True, num_workers=0)
    batches = []
    for bi, batch in enumerate(dl):
        if bi < 2:
            batches.append({k: v.numpy() for k, v in batch.items()})
        if bi >= 99:
            break
    log = np.array(rec.log, dtype=np.int64)
    np.savez_compressed(os.path.join(GOLD, "sampler_golden.npz"), log=log, n_valid=np.array([len(v) for v in ds.valid_indices]),
                        length=len(ds), **{f"b{bi}_{k}": v for bi, b in enumerate(batches) for k, v in b.items()})
    # _cut_patch edge cases on image 0
    t = ds.images_pre[0]
    H, W = t.shape[1], t.shape[2]
    pts = [(0, 0), (5, 5), (H - 1, W - 1), (H - 10, W - 20), (48, 64), (16, 16), (15, 17)]
    out = {}
    for P in (32, 80, 7):
        ds.patch_size = P
        for (y, x) in pts:
            out[f"P{P}_{y}_{x}"] = ds._cut_patch(t, torch.tensor([y, x])).numpy()
    np.savez_compressed(os.path.join(GOLD, "cutpatch_golden.npz"), image=t.numpy(), **out)
    print("sampler golden: draws", len(log), "valid per image", [len(v) for v in ds.valid_indices])
    # real sample data (only checkable where /root/reference exists)
    miku = os.path.join(REF, "test_dataset", "miku_train_sorce")
    ds2 = StyleTransferDataset(os.path.join(miku, "input"), os.path.join(miku, "output"), os.path.join(miku, "mask"), 32)
    rec2 = _Recorder(ds2)
    torch.manual_seed(0)
    np.random.seed(0)
    for bi, _ in enumerate(DataLoader(rec2, batch_size=40, shuffle=True, num_workers=0)):
        if bi >= 4:
            break
    np.savez_compressed(os.path.join(GOLD, "sampler_miku_golden.npz"), log=np.array(rec2.log, dtype=np.int64),
                        n_valid=np.array([len(v) for v in ds2.valid_indices]), length=len(ds2))
    return ds2


def generator_golden(ds_miku):
    torch.manual_seed(0)
    np.random.seed(0)
    torch.set_num_threads(os.cpu_count())
    G = GeneratorJ(input_channels=3, use_bias=True)
    G.train()
    opt = torch.optim.Adam(G.parameters(), lr=4e-4, betas=(0.9, 0.999), weight_decay=1e-5)
    dl = DataLoader(ds_miku, batch_size=40, shuffle=True, num_workers=0)
    l1 = torch.nn.L1Loss()
    losses = []
    trained = os.path.join(GOLD, "gen_c3_trained.npz")
    if os.path.exists(trained) and "--retrain" not in sys.argv:
        z = np.load(trained)
        G.load_state_dict({k: torch.from_numpy(z[k]) for k in z.files})
        losses = list(np.load(os.path.join(GOLD, "gen_c3_vectors.npz"))["train_losses"])
        dl_train = []
    else:
        dl_train = dl
    for it, batch in enumerate(dl_train):
        opt.zero_grad()
        loss = l1(G(batch["pre"]), batch["post"]) * 4.0
        loss.backward()
        torch.nn.utils.clip_grad_norm_(G.parameters(), 0.5)
        opt.step()
        losses.append(float(loss))
        if it % 20 == 0:
            print("ref train step", it, float(loss), flush=True)
        if it >= 99:
            break
    sd = {k: v.detach().clone() for k, v in G.state_dict().items()}
    if dl_train:
        np.savez(trained, **{k: v.numpy() for k, v in sd.items()})
    # vectors: a fresh batch of real patches at the C1 batch size (40 x 32x32)
    torch.manual_seed(11)
    np.random.seed(11)
    batch = next(iter(dl))
    x, tgt = batch["pre"].clone(), batch["post"].clone()
    G.train()
    G.zero_grad()
    y_train = G(x)
    loss = l1(y_train, tgt) * 4.0
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in G.named_parameters()}
    sd_after = {k: v.detach().clone() for k, v in G.state_dict().items()}
    G.load_state_dict(sd)  # undo the BatchNorm running-stat update of that forward
    G.eval()
    with torch.no_grad():
        y_eval = G(x)
        img = ds_miku.images_pre[0][:, 300:556, 700:1084].unsqueeze(0).clone()  # 256 x 384 crop of a real keyframe
        y_frame = G(img)
    out = dict(x=x.numpy(), target=tgt.numpy(), y_train=y_train.detach().numpy(), y_eval=y_eval.numpy(), loss=float(loss),
               frame=img.numpy(), y_frame=y_frame.numpy(), train_losses=np.array(losses),
               bn_rm_after=sd_after["smoothers.2.running_mean"].numpy(), bn_rv_after=sd_after["smoothers.2.running_var"].numpy())
    for k, g in grads.items():
        flat = g.reshape(-1)
        out["gsum_" + k] = float(flat.double().sum())
        out["gnorm_" + k] = float(flat.double().norm())
        out["gmax_" + k] = float(flat.abs().max())
        if flat.numel() <= 40000:
            out["g_" + k] = g.numpy()
        else:
            out["gs_" + k] = flat[:: max(1, flat.numel() // 4096)].numpy()
    np.savez_compressed(os.path.join(GOLD, "gen_c3_vectors.npz"), **out)
    print("generator golden written; final loss", losses[-1])


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    ds = sampler_golden()
    generator_golden(ds)
