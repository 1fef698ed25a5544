"""f-4 on the device: zsg_resize_rgb8 / zsg_embed_gather (csrc/data.cu) through gpu_data.GpuImageStage / GpuBatchStage, bit
for bit against what the reference's CPU worker computes with Pillow and torch (dat_loader.py:98-146)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
SIZES = [(480, 640), (375, 500), (300, 300), (123, 457), (1024, 768), (333, 300), (300, 411), (64, 48)]


def pil_path(a, resample=None):
    """dat_loader.py:121,136: img.resize((300, 300)) -> pil2tensor(...).float().div_(255)"""
    from PIL import Image
    im = Image.fromarray(a)
    im = im.resize((300, 300)) if resample is None else im.resize((300, 300), resample)
    return torch.from_numpy(np.asarray(im, dtype=np.uint8).copy()).permute(2, 0, 1).float().div_(255)


@pytest.mark.parametrize("resample", ["bicubic", "nearest"])
def test_batch_resize_equals_pillow_bit_for_bit(resample):
    from PIL import Image
    import zsg_b200  # noqa: F401
    from zsg_b200 import gpu_data
    rng = np.random.RandomState(1)
    imgs = [rng.randint(0, 256, (h, w, 3), dtype=np.uint8) for h, w in SIZES]
    imgs[0][:4, :9] = 255
    imgs[0][-3:, -5:] = 0
    stage = gpu_data.GpuImageStage("cuda", (300, 300), resample)
    for _ in range(2):                                              # second call: tables cached and already on the device
        out = stage(imgs)
        torch.cuda.synchronize()
        assert out.shape == (len(imgs), 3, 300, 300) and out.dtype == torch.float32
        for i, a in enumerate(imgs):
            want = pil_path(a, Image.NEAREST if resample == "nearest" else None)
            assert torch.equal(out[i].cpu(), want), (resample, SIZES[i], float((out[i].cpu() - want).abs().max()))
    more = [rng.randint(0, 256, (77, 91, 3), dtype=np.uint8)]       # a new source size after the first upload
    assert torch.equal(stage(more)[0].cpu(), pil_path(more[0], Image.NEAREST if resample == "nearest" else None))


def test_embed_gather_and_batch_stage_equal_the_cpu_loader(tmp_path):
    import PIL.Image
    import zsg_b200  # noqa: F401
    from zsg_b200 import dat_loader, gpu_data
    img_dir = tmp_path / "imgs"
    img_dir.mkdir()
    rng = np.random.RandomState(0)
    rows = ["img_id,bbox,query"]
    for i, (w, h) in enumerate([(64, 48), (400, 380), (333, 233), (640, 480)]):
        PIL.Image.fromarray(rng.randint(0, 255, (h, w, 3), dtype=np.uint8)).save(img_dir / f"im{i}.png")
        rows.append(f'im{i}.png,"[{w // 4}, {h // 4}, {w // 2}, {3 * h // 4}]","the red thing{" now" * i}"')
    csv = tmp_path / "a.csv"
    csv.write_text("\n".join(rows) + "\n")
    cfg = {"resize_img": [300, 300], "ds_info": {"refclef": {"img_dir": str(img_dir)}}, "bs": 4, "nw": 0}
    vocab = {"the": 0, "red": 1, "thing": 2, "now": 3, "PD": 4}
    table = torch.randn(5, 300, generator=torch.Generator().manual_seed(0))
    tok = lambda text: [vocab[t] for t in str(text).split()]
    emb = lambda text: (table[[vocab[t] for t in str(text).split()]].numpy() if str(text).split() else np.zeros((0, 300), np.float32))
    raw = dat_loader.ImgQuDataset(cfg, csv, "refclef", raw=True, tokenize=tok)
    ref = dat_loader.ImgQuDataset(cfg, csv, "refclef", embed=emb)
    want = dat_loader.collater([ref[i] for i in range(4)])          # the reference's batch (dat_loader.py:187-196)
    stage = gpu_data.GpuBatchStage("cuda", (300, 300), "bicubic", embed_table=table)
    got = stage(dat_loader.raw_collater([raw[i] for i in range(4)]))
    torch.cuda.synchronize()
    assert set(want) <= set(got)
    for k, v in want.items():
        assert got[k].dtype == torch.float32 and torch.equal(got[k].cpu(), v), k
    # padding ids give zero vectors
    t = torch.tensor([[0, -1, 4]], dtype=torch.int32, device="cuda")
    g = gpu_data.embed_gather(t, table.cuda())
    assert torch.equal(g[0, 0].cpu(), table[0]) and float(g[0, 1].abs().sum()) == 0.0 and torch.equal(g[0, 2].cpu(), table[4])
