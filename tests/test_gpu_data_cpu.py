"""f-4, host side of the device data path: the Pillow-resize oracle (oracle/pil_resize.py) is pinned against PIL itself, the
product's coefficient / index tables (zsg_b200/gpu_data.py) against the oracle's, and the raw dataset mode keeps the
reference's item contract for everything but the image and the word vectors."""
import numpy as np
import pytest
import torch

SIZES = [(480, 640), (375, 500), (300, 300), (123, 457), (1024, 768), (333, 300), (300, 411), (64, 48), (301, 299)]


@pytest.mark.parametrize("hw", SIZES)
def test_resize_oracle_equals_pillow(hw):
    """Image.resize((300, 300)) (dat_loader.py:121) bit for bit: BICUBIC = the default of the container's Pillow (>= 7.0),
    NEAREST = the default of the reference's pinned pillow 6.1.0."""
    from PIL import Image
    from oracle import pil_resize as pr
    h, w = hw
    a = np.random.RandomState(h * 1000 + w).randint(0, 256, (h, w, 3), dtype=np.uint8)
    a[:3, :5] = 255
    a[-2:, -7:] = 0                                                   # saturated edges: the negative bicubic lobes must clip
    im = Image.fromarray(a)
    assert np.array_equal(pr.resize_bicubic(a, 300, 300), np.asarray(im.resize((300, 300))))
    assert np.array_equal(pr.resize_bicubic(a, 300, 300), np.asarray(im.resize((300, 300), Image.BICUBIC)))
    assert np.array_equal(pr.resize_nearest(a, 300, 300), np.asarray(im.resize((300, 300), Image.NEAREST)))
    t = torch.from_numpy(np.asarray(im.resize((300, 300))).copy()).permute(2, 0, 1).float().div_(255)      # dat_loader.py:136
    assert np.array_equal(pr.to_unit_float(pr.resize_bicubic(a, 300, 300)), t.numpy())


@pytest.mark.parametrize("n_in", [640, 480, 300, 123, 1024, 48, 301, 299, 457, 2000, 3])
def test_product_tables_equal_oracle_tables(n_in):
    from zsg_b200 import gpu_data
    from oracle import pil_resize as pr
    b, k = gpu_data.bicubic_table(n_in, 300)
    b2, k2 = pr.precompute_coeffs(n_in, 300)
    assert np.array_equal(b, b2) and np.array_equal(k, k2) and k.dtype == np.int32
    assert np.array_equal(gpu_data.nearest_table(n_in, 300), pr.nearest_tables(n_in, n_in, 300, 300)[0])
    # every row of coefficients sums to 1.0 in fixed point up to the rounding of its entries
    assert np.abs(k.sum(1) - (1 << gpu_data.PRECISION_BITS)).max() <= k.shape[1]
    plan = gpu_data.ResizePlan(n_in, n_in, 300, 300, "bicubic")
    assert plan.y_first >= 0 and plan.y_first + plan.n_rows <= n_in and plan.htab.size == 300 * (2 + plan.hksize)
    assert gpu_data.DESC.itemsize == 48


def test_raw_dataset_mode_keeps_the_item_contract(tmp_path):
    import PIL.Image
    from zsg_b200 import dat_loader
    img_dir = tmp_path / "imgs"
    img_dir.mkdir()
    rng = np.random.RandomState(0)
    rows = ["img_id,bbox,query"]
    for i, (w, h) in enumerate([(64, 48), (40, 80), (33, 33)]):
        PIL.Image.fromarray(rng.randint(0, 255, (h, w, 3), dtype=np.uint8)).save(img_dir / f"im{i}.png")
        rows.append(f'im{i}.png,"[{w // 4}, {h // 4}, {w // 2}, {3 * h // 4}]","the red thing{" now" * i}"')
    csv = tmp_path / "a.csv"
    csv.write_text("\n".join(rows) + "\n")
    cfg = {"resize_img": [300, 300], "ds_info": {"refclef": {"img_dir": str(img_dir)}}, "bs": 3, "nw": 0}
    vocab = {"the": 0, "red": 1, "thing": 2, "now": 3, "PD": 4}
    tok = lambda text: [vocab[t] for t in str(text).split()]
    emb = lambda text: np.stack([np.full(300, vocab[t], np.float32) for t in str(text).split()]) if str(text).split() else np.zeros((0, 300))
    raw = dat_loader.ImgQuDataset(cfg, csv, "refclef", raw=True, tokenize=tok)
    ref = dat_loader.ImgQuDataset(cfg, csv, "refclef", embed=emb)
    for i in range(3):
        a, b = raw[i], ref[i]
        assert set(a) == {"idxs", "annot", "orig_annot", "img_size", "tokens", "qlens", "img_raw"}
        for k in ("idxs", "annot", "orig_annot", "img_size", "qlens"):
            assert torch.equal(a[k], b[k]), k
        assert a["img_raw"].dtype == torch.uint8 and tuple(a["img_raw"].shape) == (int(b["img_size"][0]), int(b["img_size"][1]), 3)
        assert a["tokens"].dtype == torch.int32 and a["tokens"].shape == (50,)
        assert a["tokens"].tolist() == [int(v) for v in b["qvec"][:, 0].tolist()]       # ids <-> the vectors of the CPU path
    batch = dat_loader.raw_collater([raw[0], raw[1], raw[2]])
    assert isinstance(batch["img_raw"], list) and len(batch["img_raw"]) == 3 and batch["tokens"].shape == (3, 50)
    dl = dat_loader.get_dataloader(cfg, raw, is_train=False)
    assert next(iter(dl))["tokens"].shape == (3, 50)
