"""f-4: the reference's on-disk formats either side of the hot path — annotation CSV in (dat_loader.py:66-196),
prediction pickle out (utils.py:377-381, 500-509) — on a tiny generated dataset."""
import pickle

import numpy as np
import pytest
import torch


def _embed(text):
    """Deterministic stand-in for spaCy vectors: one 300-d vector per whitespace token."""
    out = []
    for tok in str(text).split():
        g = np.random.RandomState(abs(hash(tok)) % (2 ** 31))
        out.append(g.randn(300).astype(np.float32))
    return np.array(out)


@pytest.fixture()
def tiny_ds(tmp_path):
    import PIL.Image
    img_dir = tmp_path / "imgs"
    img_dir.mkdir()
    rng = np.random.RandomState(0)
    rows = ["img_id,bbox,query"]
    for i, (w, h) in enumerate([(64, 48), (40, 80), (33, 33), (120, 60), (50, 50)]):
        PIL.Image.fromarray(rng.randint(0, 255, (h, w, 3), dtype=np.uint8)).save(img_dir / f"im{i}.png")
        q = "a_red thing" if i % 2 else "['the dog', 'dog on the left']"
        rows.append(f'im{i}.png,"[{w // 4}, {h // 4}, {w // 2}, {3 * h // 4}]","{q}"')
    strs = tmp_path / "single.csv"
    strs.write_text("\n".join([rows[0]] + [r for i, r in enumerate(rows[1:]) if i % 2]) + "\n")
    lists = tmp_path / "lists.csv"
    lists.write_text("\n".join([rows[0]] + [r for i, r in enumerate(rows[1:]) if not i % 2]) + "\n")
    return img_dir, strs, lists


def test_csv_dataset_items_follow_the_reference_contract(tiny_ds):
    from zsg_b200 import dat_loader
    img_dir, strs, lists = tiny_ds
    cfg = {"resize_img": [300, 300], "ds_info": {"refclef": {"img_dir": str(img_dir)}}}
    ds = dat_loader.ImgQuDataset(cfg, strs, "refclef", embed=_embed)
    assert len(ds) == 2
    it = ds[0]                                                    # im1.png: 40 x 80, query "a_red thing" -> "a red thing"
    assert set(it) == {"img", "idxs", "qvec", "qlens", "annot", "orig_annot", "img_size"}
    assert it["img"].shape == (3, 300, 300) and it["img"].dtype == torch.float32
    assert 0.0 <= float(it["img"].min()) and float(it["img"].max()) <= 1.0
    assert it["qvec"].shape == (50, 300) and int(it["qlens"]) == 3
    assert torch.equal(it["qvec"][3], it["qvec"][49])             # ' PD' padding tokens (dat_loader.py:107)
    assert it["img_size"].tolist() == [80, 40]
    assert it["orig_annot"].tolist() == [10.0, 20.0, 20.0, 60.0]  # pixel x1 y1 x2 y2
    # y1 x1 y2 x2, scaled by (h, w), mapped to [-1, 1] (dat_loader.py:119-128)
    np.testing.assert_allclose(it["annot"].numpy(), [2 * 20 / 80 - 1, 2 * 10 / 40 - 1, 2 * 60 / 80 - 1, 2 * 20 / 40 - 1], rtol=1e-6)
    # list-valued queries: one is drawn per access (dat_loader.py:150-151)
    dl = dat_loader.ImgQuDataset(cfg, lists, "refclef", embed=_embed)
    assert {int(dl[0]["qlens"]) for _ in range(20)} <= {2, 4}
    batch = dat_loader.collater([ds[0], ds[1]])
    assert batch["qvec"].shape == (2, 3, 300) and batch["img"].shape == (2, 3, 300, 300)
    assert all(v.dtype == torch.float32 for v in batch.values())  # the collater casts everything (dat_loader.py:193)


def test_get_data_over_csv_files_and_missing_vectors(tiny_ds):
    from zsg_b200 import dat_loader
    img_dir, strs, lists = tiny_ds
    cfg = {"resize_img": [300, 300], "bs": 2, "nw": 0, "num_gpus": 1, "do_dist": False, "tmp_path": "./tmp",
           "ds_to_use": "refclef",
           "ds_info": {"refclef": {"img_dir": str(img_dir), "trn_csv_file": str(strs), "val_csv_file": str(strs),
                                   "test_csv_file": str(lists)}}}
    data = dat_loader.get_data(cfg, embed=_embed)
    b = next(iter(data.valid_dl))
    assert b["img"].shape == (2, 3, 300, 300) and b["annot"].shape == (2, 4) and set(data.test_dl) == {"test0"}
    try:
        import spacy  # noqa: F401
    except ImportError:
        with pytest.raises(RuntimeError, match="word vectors"):
            dat_loader.ImgQuDataset(cfg, strs, "refclef")


def test_prediction_file_format(tmp_path):
    from zsg_b200 import dat_loader
    metric = {"idxs": torch.tensor([3.0, 7.0]), "pred_boxes": torch.tensor([[1.0, 2.0, 3.0, 4.0], [5.0, 6.0, 7.0, 8.0]]),
              "pred_scores": torch.tensor([0.25, 0.5])}
    recs = dat_loader.prediction_records(metric)
    assert recs == [{"id": 3.0, "pred_boxes": [1.0, 2.0, 3.0, 4.0], "pred_scores": 0.25},
                    {"id": 7.0, "pred_boxes": [5.0, 6.0, 7.0, 8.0], "pred_scores": 0.5}]
    f = dat_loader.write_prediction_file(recs, tmp_path / "valid_preds.pkl")
    assert pickle.load(open(f, "rb")) == recs
    f1 = dat_loader.write_prediction_file(recs, tmp_path / "valid_preds.pkl", rank=1, distributed=True)
    assert f1.name == "1_valid_preds.pkl" and pickle.load(open(f1, "rb")) == recs
