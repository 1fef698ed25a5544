"""Seeded synthetic inputs and deterministic weights for the ZSGNet hot path.

TEST INFRASTRUCTURE (part of oracle/): imported only by tests/, bench.py's
cpu_baseline / --impl reference legs, __graft_entry__.smoke() and the golden
generator.  The product package never imports this module.

Input contract follows the reference's batch dict (dat_loader.py:136-144,
187-196): every tensor float32; `annot` is (y1,x1,y2,x2) in [-1,1]; `qvec`
B x T x 300; `qlens` float.  Distributions follow SURVEY.md section 8(d).
"""
import math
import torch

RESNET_LAYERS = ((3, 64, 1), (4, 128, 2), (6, 256, 2), (3, 512, 2))  # blocks, width, stride
LEVEL_SIZES = (38, 19, 10, 5, 3, 1)          # P3..P8 for a 300x300 image
NUM_ANCHORS_PER_CELL = 9
NUM_ANCHORS = NUM_ANCHORS_PER_CELL * sum(s * s for s in LEVEL_SIZES)  # 17460


def make_batch(B, seed=1234, T=20, var_len=False, img_hw=300, adversarial=False):
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(B, 3, img_hw, img_hw, generator=g)
    qvec = torch.randn(B, T, 300, generator=g)
    if var_len:
        qlens = torch.randint(1, T + 1, (B,), generator=g).float()
        qlens[0] = float(T)
    else:
        qlens = torch.full((B,), float(T))
    c = torch.rand(B, 2, generator=g) * 1.2 - 0.6
    s = torch.rand(B, 2, generator=g) * 0.7 + 0.1
    annot = torch.cat([c - s / 2, c + s / 2], dim=1).clamp_(-1.0, 1.0)
    if adversarial:
        annot = adversarial_boxes(B, annot)
    img_size = torch.tensor([[480.0, 640.0]]).repeat(B, 1)
    hw = img_size
    orig = torch.stack([(annot[:, 1] + 1) / 2 * hw[:, 1], (annot[:, 0] + 1) / 2 * hw[:, 0],
                        (annot[:, 3] + 1) / 2 * hw[:, 1], (annot[:, 2] + 1) / 2 * hw[:, 0]], dim=1)
    return {"img": img, "qvec": qvec, "qlens": qlens, "annot": annot,
            "idxs": torch.arange(B).float(), "img_size": img_size, "orig_annot": orig}


def adversarial_boxes(B, base):
    """Tie storms and edge cases for the matcher: a box equal to an
    anchor, a box inside many same-shape anchors, the full image, a tiny box."""
    out = base.clone()
    cases = [
        [-1.0, -1.0, 1.0, 1.0],                       # full image
        [-0.01, -0.01, 0.01, 0.01],                   # tiny: only the top-1 is positive
        [-0.2, -0.2, 0.2, 0.2],                       # symmetric about the centre: ties
        [0.0, 0.0, 0.5, 0.5],
        [-1.0, -1.0, -0.5, -0.25],                    # touching the border
        [-4 / 38, -4 / 38, 4 / 38, 4 / 38],           # equals a ratio-1 scale-4 P3 anchor size
    ]
    for i, cbox in enumerate(cases):
        if i < B:
            out[i] = torch.tensor(cbox)
    return out


def param_specs():
    """(key, shape, kind) in the reference's state_dict order:
    backbone.encoder.* (torchvision resnet50 names), backbone.fpn.*, att_reg_box.*, lstm.*"""
    specs = []
    enc = "backbone.encoder."

    def bn(prefix, c):
        specs.append((prefix + ".weight", (c,), "bn_w"))
        specs.append((prefix + ".bias", (c,), "bn_b"))
        specs.append((prefix + ".running_mean", (c,), "bn_rm"))
        specs.append((prefix + ".running_var", (c,), "bn_rv"))
        specs.append((prefix + ".num_batches_tracked", (), "bn_n"))

    specs.append((enc + "conv1.weight", (64, 3, 7, 7), "conv"))
    bn(enc + "bn1", 64)
    inplanes = 64
    for li, (nblk, width, stride) in enumerate(RESNET_LAYERS, start=1):
        for b in range(nblk):
            p = f"{enc}layer{li}.{b}."
            specs.append((p + "conv1.weight", (width, inplanes, 1, 1), "conv"))
            bn(p + "bn1", width)
            specs.append((p + "conv2.weight", (width, width, 3, 3), "conv"))
            bn(p + "bn2", width)
            specs.append((p + "conv3.weight", (width * 4, width, 1, 1), "conv"))
            bn(p + "bn3", width * 4)
            if b == 0:
                specs.append((p + "downsample.0.weight", (width * 4, inplanes, 1, 1), "conv"))
                bn(p + "downsample.1", width * 4)
            inplanes = width * 4
    specs.append((enc + "fc.weight", (1000, 2048), "lin"))      # present, unused (mdl.py:149-156)
    specs.append((enc + "fc.bias", (1000,), "bias"))
    f = "backbone.fpn."
    for name, cin, k in (("P7_2", 256, 3), ("P6", 2048, 3), ("P5_1", 2048, 1), ("P5_2", 256, 3),
                         ("P4_1", 1024, 1), ("P4_2", 256, 3), ("P3_1", 512, 1), ("P3_2", 256, 3)):
        specs.append((f + name + ".weight", (256, cin, k, k), "conv"))
        specs.append((f + name + ".bias", (256,), "bias"))
    specs.append(("att_reg_box.0.0.weight", (256, 514, 3, 3), "conv"))
    specs.append(("att_reg_box.0.0.bias", (256,), "bias"))
    for i in range(1, 5):
        specs.append((f"att_reg_box.{i}.0.weight", (256, 256, 3, 3), "conv"))
        specs.append((f"att_reg_box.{i}.0.bias", (256,), "bias"))
    specs.append(("att_reg_box.5.weight", (45, 256, 3, 3), "conv"))
    specs.append(("att_reg_box.5.bias", (45,), "final_bias"))
    for sfx in ("", "_reverse"):
        specs.append((f"lstm.weight_ih_l0{sfx}", (512, 300), "lstm"))
        specs.append((f"lstm.weight_hh_l0{sfx}", (512, 128), "lstm"))
        specs.append((f"lstm.bias_ih_l0{sfx}", (512,), "lstm"))
        specs.append((f"lstm.bias_hh_l0{sfx}", (512,), "lstm"))
    return specs


VGG_CFG = (64, 64, "M", 128, 128, "M", 256, 256, 256, "C", 512, 512, 512, "M", 512, 512, 512)   # ssd_vgg.py:174-177
# extras (ssd_vgg.py:140-154, cfg at 179-182): (cin, cout, kernel, stride, pad); ReLU after each (ssd_vgg.py:92-93)
VGG_EXTRAS = ((1024, 256, 1, 1, 0), (256, 512, 3, 2, 1), (512, 128, 1, 1, 0), (128, 256, 3, 2, 1),
              (256, 128, 1, 1, 0), (128, 256, 3, 1, 0), (256, 128, 1, 1, 0), (128, 256, 3, 1, 0))
VGG_MBOX = (4, 6, 6, 6, 4, 4)                                                                    # ssd_vgg.py:183-186


def vgg_layers():
    """ssd_vgg.py:111-133 as a flat list aligned with the `vgg` ModuleList indices:
    ("conv", cin, cout, k, pad, dil) | ("relu",) | ("pool", k, stride, pad, ceil)."""
    out, cin = [], 3
    for v in VGG_CFG:
        if v == "M":
            out.append(("pool", 2, 2, 0, False))
        elif v == "C":
            out.append(("pool", 2, 2, 0, True))
        else:
            out += [("conv", cin, v, 3, 1, 1), ("relu",)]
            cin = v
    out += [("pool", 3, 1, 1, False), ("conv", 512, 1024, 3, 6, 6), ("relu",), ("conv", 1024, 1024, 1, 0, 1), ("relu",)]
    return out


def vgg_param_specs():
    """(key, shape, kind) of ZSGNet over SSDBackBone (mdl.py:162-168, ssd_vgg.py:31-52): vgg.*, fproj1-3,
    extras.*, the unused loc.* / conf.* multibox heads (built at ssd_vgg.py:157-171, never called), then the
    shared head and the LSTM as in param_specs()."""
    specs = []
    e = "backbone.encoder."

    def conv(name, cin, cout, k):
        specs.append((name + ".weight", (cout, cin, k, k), "conv"))
        specs.append((name + ".bias", (cout,), "bias"))
    for i, L in enumerate(vgg_layers()):
        if L[0] == "conv":
            conv(f"{e}vgg.{i}", L[1], L[2], L[3])
    conv(e + "fproj1", 512, 256, 1)
    conv(e + "fproj2", 1024, 256, 1)
    conv(e + "fproj3", 512, 256, 1)
    for i, (cin, cout, k, _, _) in enumerate(VGG_EXTRAS):
        conv(f"{e}extras.{i}", cin, cout, k)
    src_ch = (512, 1024, 512, 256, 256, 256)
    for i, (c, nb) in enumerate(zip(src_ch, VGG_MBOX)):
        conv(f"{e}loc.{i}", c, nb * 4, 3)
    for i, (c, nb) in enumerate(zip(src_ch, VGG_MBOX)):
        conv(f"{e}conf.{i}", c, nb * 21, 3)
    specs += [s for s in param_specs() if s[0].startswith(("att_reg_box.", "lstm."))]
    return specs


def make_state_dict(seed=0, model="retina"):
    """Deterministic random-init weights with the reference's names and shapes.
    Distributions mimic the PyTorch defaults in scale; BN affine parameters are
    randomised (not 1/0) so that parity tests exercise them."""
    sd = {}
    specs, base = (param_specs(), 0) if model == "retina" else (vgg_param_specs(), 50000)
    for i, (key, shape, kind) in enumerate(specs):
        g = torch.Generator().manual_seed(seed * 100003 + base + i)
        if kind == "conv":
            fan_in = shape[1] * shape[2] * shape[3]
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in) * 0.8
        elif kind == "lin":
            t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(shape[1])
        elif kind == "bias":
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
        elif kind == "final_bias":
            t = torch.zeros(shape)
            t[torch.arange(4, shape[0], 5)] = -4.0           # mdl.py:214-215
        elif kind == "bn_w":
            t = torch.rand(shape, generator=g) * 0.6 + 0.7
        elif kind == "bn_b":
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.1
        elif kind == "bn_rm":
            t = torch.zeros(shape)
        elif kind == "bn_rv":
            t = torch.ones(shape)
        elif kind == "bn_n":
            t = torch.zeros(shape, dtype=torch.long)
        elif kind == "lstm":
            t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(128.0)
        else:
            raise KeyError(kind)
        sd[key] = t
    return sd


def default_cfg(model="retina"):
    """The hot-path keys of configs/cfg.json:1-43 (only those the path reads)."""
    return {"do_norm": False, "use_same_atb": True, "mdl_to_use": model,
            "resize_img": [300, 300], "use_multi": True, "use_focal": True,
            "use_softmax": False, "alpha": 0.25, "gamma": 2, "emb_dim": 300,
            "matching_threshold": 0.6, "use_bidirectional": True, "lstm_dim": 128,
            "lamb_reg": 1, "acc_iou_threshold": 0.5, "use_lang": True, "use_img": True,
            "ratios": "[1/2, 1, 2]", "scales": "[1, 2**(1/3), 2**(2/3)]", "scale_factor": 4,
            "device": "cpu"}


def ratios_scales(cfg=None):
    """main_dist.py:24-31: ratios as python floats, scales as numpy float64."""
    import numpy as np
    cfg = cfg or default_cfg()
    ratios = eval(cfg["ratios"], {}) if not isinstance(cfg["ratios"], list) else cfg["ratios"]
    sc = eval(cfg["scales"], {}) if not isinstance(cfg["scales"], list) else cfg["scales"]
    scales = cfg["scale_factor"] * np.array(sc)
    return ratios, scales
