"""Batch contract of the reference's code/dat_loader.py (the boundary on the input side).

The reference's loader is CPU work outside the hot path (PIL decode, spaCy vectors; SURVEY.md
section 2 #8) and is not rebuilt.  What the hot path needs from it is the batch dict
(dat_loader.py:136-144, 187-196) and `get_data(cfg) -> DataWrap`; this module supplies both over
a seeded synthetic dataset of the BASELINE shape (300x300 images, 300-d query vectors)."""
from dataclasses import dataclass
from typing import Dict, Optional, Union

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
    so the allocator does not recycle them early."""

    def __init__(self, batches, device):
        self.it, self.device = iter(batches), torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.next = self._fetch()

    def _fetch(self):
        try:
            host = next(self.it)
        except StopIteration:
            return None
        with torch.cuda.stream(self.stream):
            dev = {k: v.to(self.device, non_blocking=True) for k, v in host.items()}
            ev = torch.cuda.Event()
            ev.record(self.stream)
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


def get_dataloader(cfg, dataset, is_train):
    dist = bool(cfg["do_dist"]) if "do_dist" in cfg else False
    bs = cfg["bs"] if dist else cfg["bs"] * max(1, int(cfg["num_gpus"]) if "num_gpus" in cfg else 1)
    if dist:
        sampler = NewDistributedSampler(dataset, shuffle=True)
    else:
        sampler = (torch.utils.data.RandomSampler if is_train else torch.utils.data.SequentialSampler)(dataset)
    return DataLoader(dataset, batch_size=bs, sampler=sampler, drop_last=is_train, num_workers=0, collate_fn=collater)


def get_data(cfg):
    """Same signature as dat_loader.py:230-253, over synthetic data."""
    n = cfg["synthetic_len"] if "synthetic_len" in cfg else 256
    trn, val = SyntheticImgQuDataset(n, seed=0), SyntheticImgQuDataset(max(n // 4, 1), seed=1)
    return DataWrap(path=cfg["tmp_path"] if "tmp_path" in cfg else "./tmp", train_dl=get_dataloader(cfg, trn, True),
                    valid_dl=get_dataloader(cfg, val, False), test_dl={"test0": get_dataloader(cfg, val, False)})
