"""Batch contract of the reference's code/dat_loader.py (the boundary on the input side).

What the hot path needs from the loader is the batch dict (dat_loader.py:136-144, 187-196) and
`get_data(cfg) -> DataWrap`.  This module supplies both
  * over a seeded synthetic dataset of the BASELINE shape (300x300 images, 300-d query vectors), and
  * over the reference's on-disk format: the annotation CSV `img_id,bbox,query`
    (DATA_PREP_README.md:10-11) read like ImgQuDataset (dat_loader.py:66-185), images decoded and resized with PIL
    as there.  The word vectors are the one thing the reference takes from a package that is not in this image
    (spaCy en_core_web_md, dat_loader.py:23): ImgQuDataset takes an `embed(text) -> [n_tokens, 300]` callable
    and falls back to spaCy only when it is importable.
The prediction file of Learner.validate / update_prediction_file (utils.py:377-381, 500-509) is written by
`prediction_records` / `write_prediction_file`."""
import ast
import pickle
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, Optional, Union

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset
from torch.utils.data.distributed import DistributedSampler


@dataclass
class DataWrap:                                     # utils.py:115-120
    path: str
    train_dl: DataLoader
    valid_dl: DataLoader
    test_dl: Optional[Union[DataLoader, Dict]] = None


class NewDistributedSampler(DistributedSampler):
    """dat_loader.py:36-65: epoch-seeded permutation, padded to a multiple of the world size, rank slice."""

    def __init__(self, dataset, num_replicas=None, rank=None, shuffle=True):
        super().__init__(dataset, num_replicas=num_replicas, rank=rank)
        self.shuffle = shuffle

    def __iter__(self):
        if self.shuffle:
            g = torch.Generator()
            g.manual_seed(self.epoch)
            indices = torch.randperm(len(self.dataset), generator=g).tolist()
        else:
            indices = torch.arange(len(self.dataset)).tolist()
        indices += indices[: (self.total_size - len(indices))]
        off = self.num_samples * self.rank
        return iter(indices[off: off + self.num_samples])


class SyntheticImgQuDataset(Dataset):
    """Items with the keys/dtypes of ImgQuDataset.simple_item_getter (dat_loader.py:136-144)."""

    def __init__(self, n=256, phrase_len=50, img_hw=300, seed=0, min_len=1, max_len=20):
        self.n, self.phrase_len, self.hw, self.seed, self.min_len, self.max_len = n, phrase_len, img_hw, seed, min_len, max_len

    def __len__(self):
        return self.n

    def __getitem__(self, idx):
        g = torch.Generator().manual_seed(self.seed * 1000003 + idx)
        qlen = int(torch.randint(self.min_len, self.max_len + 1, (1,), generator=g))
        qvec = torch.zeros(self.phrase_len, 300)
        qvec[:qlen] = torch.randn(qlen, 300, generator=g)
        c = torch.rand(2, generator=g) * 1.2 - 0.6
        s = torch.rand(2, generator=g) * 0.7 + 0.1
        annot = torch.cat([c - s / 2, c + s / 2]).clamp_(-1, 1)
        h, w = 480.0, 640.0
        orig = torch.tensor([(annot[1] + 1) / 2 * w, (annot[0] + 1) / 2 * h, (annot[3] + 1) / 2 * w, (annot[2] + 1) / 2 * h])
        return {"img": torch.rand(3, self.hw, self.hw, generator=g), "idxs": torch.tensor(idx),
                "qvec": qvec, "qlens": torch.tensor(qlen), "annot": annot, "orig_annot": orig,
                "img_size": torch.tensor([h, w])}


class ImgQuDataset(Dataset):
    """dat_loader.py:66-185.  One item per CSV row: the image resized to cfg.resize_img and scaled to [0,1] (no mean /
    std normalisation, 89-90), the phrase padded with ' PD' tokens to 50 vectors (107-114), the box turned from pixel
    x1y1x2y2 into y1x1y2x2 in [-1,1] (119-128).  Keys and dtypes are the reference's (136-144)."""

    def __init__(self, cfg, csv_file, ds_name, split_type="train", embed=None, raw=False, tokenize=None):
        """raw=True (device-side data path, gpu_data.GpuBatchStage): an item carries the DECODED image as it is (`img_raw`,
        uint8 [h, w, 3]) instead of the resized float tensor -- the resize and /255 run on the GPU -- and, when `tokenize`
        (text -> list of vocabulary ids) is given, `tokens` (int32, -1 padded to 50) instead of `qvec`: the embedding
        lookup runs on the GPU as well.  Workers then only decode and tokenise."""
        self.cfg, self.ann_file, self.ds_name, self.split_type = cfg, csv_file, ds_name, split_type
        self.image_data = self._read_annotations(csv_file)
        self.img_dir = Path(cfg["ds_info"][ds_name]["img_dir"])
        self.phrase_len = 50
        self.raw, self.tokenize = raw, tokenize
        self.embed = embed if embed is not None else (None if (raw and tokenize is not None) else _spacy_embedder())

    def __len__(self):
        return len(self.image_data)

    def _read_annotations(self, csv_file):
        """dat_loader.py:163-185: bbox is a python-literal list [x1,y1,x2,y2]; query is a string or a python-literal
        list of strings (decided on the first row); flickr30k image names get a .jpg suffix."""
        import pandas as pd
        df = pd.read_csv(csv_file)
        df["bbox"] = df.bbox.apply(ast.literal_eval)
        if str(df["query"].iloc[0])[0] == "[":
            df["query"] = df["query"].apply(ast.literal_eval)
        names = df.img_id.apply(lambda x: f"{x}.jpg") if self.ds_name == "flickr30k" else df.img_id
        return [(n, b[0], b[1], b[2], b[3], q) for n, b, q in zip(names, df.bbox, df["query"])]

    def load_annotations(self, idx):
        img_file, x1, y1, x2, y2, queries = self.image_data[idx]
        q = str(np.random.choice(queries)) if isinstance(queries, list) else queries
        assert isinstance(q, str)
        return self.img_dir / f"{img_file}", np.array([x1, y1, x2, y2]), q.replace("_", " ")

    def __getitem__(self, idx):
        import PIL.Image
        img_file, annot, q = self.load_annotations(idx)
        img = PIL.Image.open(img_file).convert("RGB")
        h, w = img.height, img.width
        q = q.strip()
        t = np.array([annot[1] / h, annot[0] / w, annot[3] / h, annot[2] / w])
        item = {"idxs": torch.tensor(idx).long(), "annot": torch.from_numpy(2 * t - 1).float(),
                "orig_annot": torch.tensor(annot).float(), "img_size": torch.tensor([h, w])}
        if self.raw and self.tokenize is not None:
            ids = list(self.tokenize(q))
            qlen = len(ids)
            ids = ids + list(self.tokenize(" PD" * (self.phrase_len - qlen)))       # the ' PD' padding tokens (dat_loader.py:110)
            tok = torch.full((self.phrase_len,), -1, dtype=torch.int32)
            tok[:min(len(ids), self.phrase_len)] = torch.tensor(ids[: self.phrase_len], dtype=torch.int32)
            item["tokens"] = tok
        else:
            qlen = len(self.embed(q))
            vecs = np.asarray(self.embed(q + " PD" * (self.phrase_len - qlen)), dtype=np.float32)[: self.phrase_len]
            item["qvec"] = torch.from_numpy(vecs)
        if qlen == 0:
            raise NotImplementedError("empty query")                 # dat_loader.py:103-105
        item["qlens"] = torch.tensor(qlen)
        if self.raw:
            item["img_raw"] = torch.from_numpy(np.asarray(img, dtype=np.uint8).copy())
            return item
        img = img.resize((self.cfg["resize_img"][0], self.cfg["resize_img"][1]))
        item["img"] = torch.from_numpy(np.asarray(img, dtype=np.uint8).copy()).permute(2, 0, 1).float().div_(255)
        return item


def _spacy_embedder():
    try:
        import spacy
        nlp = spacy.load("en_core_web_md")                         # dat_loader.py:23
    except Exception as e:                                           # not in this image
        raise RuntimeError("ImgQuDataset needs word vectors: pass embed=callable(text) -> [n_tokens, 300] "
                           "(spaCy en_core_web_md, the reference's source, is not importable here)") from e
    return lambda text: np.array([t.vector for t in nlp(str(text))])


def prediction_records(metric):
    """The per-sample prediction dicts Learner.validate collects (utils.py:377-383): metric = Evaluator output."""
    ids, boxes, scores = metric["idxs"].tolist(), metric["pred_boxes"].tolist(), metric["pred_scores"].tolist()
    return [{"id": i, "pred_boxes": b, "pred_scores": s} for i, b, s in zip(ids, boxes, scores)]


def write_prediction_file(predictions, pred_file, rank=0, distributed=False):
    """utils.py:500-509: a pickled list of {'id','pred_boxes','pred_scores'}; one file per rank under DDP
    ('<rank>_<name>'), which eval_script.py:4-8 reads back."""
    pred_file = Path(pred_file)
    target = pred_file.parent / f"{rank}_{pred_file.name}" if distributed else pred_file
    with target.open("wb") as f:
        pickle.dump(predictions, f)
    return target


def synthetic_batch(B, seed=0, T=20, img_hw=300, pin=False):
    """A whole batch at once with the collater's keys/dtypes (BASELINE shape: 300x300, qlen = T)."""
    g = torch.Generator().manual_seed(seed)
    c = torch.rand(B, 2, generator=g) * 1.2 - 0.6
    s = torch.rand(B, 2, generator=g) * 0.7 + 0.1
    annot = torch.cat([c - s / 2, c + s / 2], dim=1).clamp_(-1, 1)
    hw = torch.tensor([[480.0, 640.0]]).repeat(B, 1)
    orig = torch.stack([(annot[:, 1] + 1) / 2 * 640, (annot[:, 0] + 1) / 2 * 480, (annot[:, 3] + 1) / 2 * 640,
                        (annot[:, 2] + 1) / 2 * 480], dim=1)
    out = {"img": torch.rand(B, 3, img_hw, img_hw, generator=g), "idxs": torch.arange(B).float(),
           "qvec": torch.randn(B, T, 300, generator=g), "qlens": torch.full((B,), float(T)), "annot": annot,
           "orig_annot": orig, "img_size": hw}
    return {k: v.pin_memory() for k, v in out.items()} if pin else out


class DevicePrefetcher:
    """Wraps an iterable of host batches (pinned memory) and yields device batches, copying batch i + 1 on a side
    stream while batch i is being used -- what `Learner.train_epoch`'s `batch[k].to(device)` (utils.py:405-406) does,
    taken off the critical path.  The consumer's stream waits on the copy's event, and the tensors are recorded on it
    so the allocator does not recycle them early.

    The image tensor (69 MB at bs = 64) is copied in chunks of CHUNK_BYTES: the host-to-device DMA engine serves copies
    in order, and the step that is starting has a few tiny host-to-device copies of its own (query lengths, LSTM initial
    states) -- behind one monolithic image copy of the NEXT batch they, and the whole forward pass with them, waited
    for it (measured: the end-to-end step was slower than the device-resident one by exactly the image copy time).
    `qlens_cpu` (the host copy of the query lengths) rides along so that the model does not read them back.

    lstm_state=True also stages what the model would otherwise copy at the start of its forward pass: the int32 query
    lengths and the two torch.randn(2, B, 128) draws of mdl.py:279-294 (made here, when the batch is fetched, from the
    same global CPU RNG and in batch order -- the values a forward pass would draw if nothing else consumed the RNG in
    between).  The step then starts without any host-to-device copy queued behind the next batch's image."""
    CHUNK_BYTES = 4 << 20
    _streams = {}            # one copy stream per device for every prefetcher: the caching allocator pools memory per stream,
                             # so a fresh stream per epoch would start each epoch with cudaMalloc stalls (measured: 100-200 ms)

    def __init__(self, batches, device, lstm_state=False):
        self.it, self.device, self.lstm_state = iter(batches), torch.device(device), lstm_state
        key = (self.device.type, self.device.index if self.device.index is not None else torch.cuda.current_device())
        if key not in DevicePrefetcher._streams:
            DevicePrefetcher._streams[key] = torch.cuda.Stream(device=self.device)
        self.stream = DevicePrefetcher._streams[key]
        self.next = self._fetch()

    def _to_device(self, v):
        n = v.numel() * v.element_size()
        if n <= self.CHUNK_BYTES or v.dim() == 0 or not v.is_contiguous():
            return v.to(self.device, non_blocking=True)
        out = torch.empty(v.shape, dtype=v.dtype, device=self.device)
        flat_s, flat_d = v.view(-1), out.view(-1)
        step = max(1, self.CHUNK_BYTES // v.element_size())
        for o in range(0, flat_s.numel(), step):
            flat_d[o:o + step].copy_(flat_s[o:o + step], non_blocking=True)
        return out

    def _fetch(self):
        try:
            host = next(self.it)
        except StopIteration:
            return None
        extra = {}
        if self.lstm_state and "qlens" in host:
            from .engine import Engine
            from .mdl import draw_lstm_state
            h0, c0, inv = draw_lstm_state(host["qlens"])
            extra["_zsg_h0c0"] = Engine.lstm_state_per_sample(h0, c0, inv).pin_memory()
            extra["_zsg_lens"] = host["qlens"].to(torch.int32).pin_memory()
        with torch.cuda.stream(self.stream):
            dev = {k: self._to_device(v) for k, v in list(host.items()) + list(extra.items())}
            ev = torch.cuda.Event()
            ev.record(self.stream)
        if "qlens" in host and "qlens_cpu" not in host:
            dev["qlens_cpu"] = host["qlens"]
        return dev, ev

    def __iter__(self):
        return self

    def __next__(self):
        if self.next is None:
            raise StopIteration
        dev, ev = self.next
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for v in dev.values():
            if v.is_cuda:
                v.record_stream(cur)
        self.next = self._fetch()
        return dev


def collater(batch):
    """dat_loader.py:187-196: stack, cast everything to float, trim qvec to the batch's longest phrase."""
    qlens = torch.Tensor([i["qlens"] for i in batch])
    max_qlen = int(qlens.max().item())
    out = {k: torch.stack([b[k] for b in batch]).float() for k in batch[0]}
    out["qvec"] = out["qvec"][:, :max_qlen]
    return out


def raw_collater(batch):
    """Collate for the device-side data path: images stay a list (their sizes differ), everything else is stacked as in
    collater(); gpu_data.GpuBatchStage turns the result into the reference's batch dict on the device."""
    out = {k: torch.stack([b[k] for b in batch]) for k in batch[0] if k != "img_raw"}
    out["img_raw"] = [b["img_raw"] for b in batch]
    return out


def get_dataloader(cfg, dataset, is_train):
    """dat_loader.py:209-227: per-GPU batch under DDP, `cfg.nw` worker processes (the synthetic datasets need none)."""
    dist = bool(cfg["do_dist"]) if "do_dist" in cfg else False
    bs = cfg["bs"] if dist else cfg["bs"] * max(1, int(cfg["num_gpus"]) if "num_gpus" in cfg else 1)
    if dist:
        sampler = NewDistributedSampler(dataset, shuffle=True)
    else:
        sampler = (torch.utils.data.RandomSampler if is_train else torch.utils.data.SequentialSampler)(dataset)
    nw = int(cfg["nw"]) if ("nw" in cfg and isinstance(dataset, ImgQuDataset)) else 0
    collate = raw_collater if getattr(dataset, "raw", False) else collater
    return DataLoader(dataset, batch_size=bs, sampler=sampler, drop_last=is_train, num_workers=nw, collate_fn=collate,
                      pin_memory=not getattr(dataset, "raw", False) and torch.cuda.is_available())


def get_data(cfg, embed=None, raw=False, tokenize=None):
    """Same signature as dat_loader.py:230-253.  With cfg.ds_to_use / cfg.ds_info set, the reference's CSV datasets
    (train / valid / test); otherwise synthetic data of the BASELINE shape.  raw / tokenize: the device-side data path
    (ImgQuDataset raw mode; wrap the loaders' batches with gpu_data.GpuBatchStage)."""
    if "ds_to_use" in cfg and "ds_info" in cfg and cfg["ds_to_use"] in cfg["ds_info"]:
        name = cfg["ds_to_use"]
        info = cfg["ds_info"][name]
        ds = {k: ImgQuDataset(cfg, info[f"{k}_csv_file"], name, split_type="train" if k == "trn" else "valid", embed=embed,
                              raw=raw, tokenize=tokenize)
              for k in ("trn", "val", "test")}
        return DataWrap(path=cfg["tmp_path"] if "tmp_path" in cfg else "./tmp",
                        train_dl=get_dataloader(cfg, ds["trn"], True), valid_dl=get_dataloader(cfg, ds["val"], False),
                        test_dl={"test0": get_dataloader(cfg, ds["test"], False)})
    n = cfg["synthetic_len"] if "synthetic_len" in cfg else 256
    trn, val = SyntheticImgQuDataset(n, seed=0), SyntheticImgQuDataset(max(n // 4, 1), seed=1)
    return DataWrap(path=cfg["tmp_path"] if "tmp_path" in cfg else "./tmp", train_dl=get_dataloader(cfg, trn, True),
                    valid_dl=get_dataloader(cfg, val, False), test_dl={"test0": get_dataloader(cfg, val, False)})
