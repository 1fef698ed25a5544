"""Host-side anchor tables (init-time only; mirrors code/anchors.py:47-87 of the reference).

The table is built once with the same torch expressions and dtypes as the reference (float64
aspects from numpy scales, the per-level 2/h factor rounded to float32 first, float32 cell
centres promoted by the concatenation) and handed to the CUDA kernels as a float64 [A,4] array."""
import numpy as np
import torch


def cell_centres(n):
    return torch.linspace(-1 + 1 / n, 1 - 1 / n, n) if n > 1 else torch.tensor([0.0])


def cell_grid(H, W):
    """[H,W,2] float32: channel 0 = row centre (y), channel 1 = column centre (x) (anchors.py:47-63)."""
    ys = cell_centres(H).view(H, 1).expand(H, W)
    xs = cell_centres(W).view(1, W).expand(H, W)
    return torch.stack([ys, xs], dim=2).contiguous()


def create_anchors(sizes, ratios, scales, flatten=True, device=None):
    """Same call signature as the reference's create_anchors (anchors.py:66); returns tlbr float64."""
    aspects = torch.tensor([[[s * np.sqrt(r), s * np.sqrt(1 / r)] for s in scales] for r in ratios],
                           dtype=torch.float64).view(-1, 2)
    out = []
    for h, w in sizes:
        h, w = int(h), int(w)
        hw = aspects * torch.tensor([2 / h, 2 / w], dtype=torch.float32).double()
        ctr = cell_grid(h, w).view(-1, 2).double()
        n, a = ctr.shape[0], hw.shape[0]
        out.append(torch.cat([ctr.unsqueeze(1).expand(n, a, 2), hw.unsqueeze(0).expand(n, a, 2)], dim=2).reshape(-1, 4))
    cthw = torch.cat(out, dim=0)
    tlbr = torch.cat([cthw[:, :2] - cthw[:, 2:] / 2, cthw[:, :2] + cthw[:, 2:] / 2], dim=1).contiguous()
    return tlbr.to(device) if device is not None else tlbr
