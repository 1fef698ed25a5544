"""Static description of the ZSGNet parameter sets: names and shapes are the reference's state_dict
contract (mdl.py:171-229, fpn_resnet.py:108-152, torchvision resnet50 for
mdl_to_use='retina'; ssd_vgg.py:31-52,111-171 for 'ssd_vgg'), listed in FORWARD execution order."""

RESNET_LAYERS = ((3, 64, 1), (4, 128, 2), (6, 256, 2), (3, 512, 2))     # blocks, width, stride
LEVEL_SIZES = (38, 19, 10, 5, 3, 1)                                     # P3..P8 at 300x300
CELLS = tuple(s * s for s in LEVEL_SIZES)
TOTAL_CELLS = sum(CELLS)                                                # 1940
ANCHORS_PER_CELL = 9
NUM_ANCHORS = TOTAL_CELLS * ANCHORS_PER_CELL                            # 17460
FUSED_C = 514
FUSED_CP = 520                                                          # padded to a multiple of 8


def _bn(prefix, c):
    return [(prefix + ".weight", (c,), "bn_w"), (prefix + ".bias", (c,), "bn_b")]


def bn_buffers(prefix, c):
    return [(prefix + ".running_mean", (c,)), (prefix + ".running_var", (c,)), (prefix + ".num_batches_tracked", ())]


MODELS = ("retina", "ssd_vgg")

# ssd_vgg.py:174-177 (base['300']), 179-182 (extras['300']), 183-186 (mbox['300'])
VGG_CFG = (64, 64, "M", 128, 128, "M", 256, 256, 256, "C", 512, 512, 512, "M", 512, 512, 512)
VGG_EXTRAS = ((1024, 256, 1, 1, 0), (256, 512, 3, 2, 1), (512, 128, 1, 1, 0), (128, 256, 3, 2, 1),     # cin, cout, k, stride, pad
              (256, 128, 1, 1, 0), (128, 256, 3, 1, 0), (256, 128, 1, 1, 0), (128, 256, 3, 1, 0))
VGG_MBOX = (4, 6, 6, 6, 4, 4)


def vgg_layers():
    """The `vgg` ModuleList (ssd_vgg.py:111-133), index-aligned: ("conv", cin, cout, k, pad, dil) | ("relu",) |
    ("pool", k, stride, pad, ceil_mode)."""
    out, cin = [], 3
    for v in VGG_CFG:
        if v in ("M", "C"):
            out.append(("pool", 2, 2, 0, v == "C"))
        else:
            out += [("conv", cin, v, 3, 1, 1), ("relu",)]
            cin = v
    out += [("pool", 3, 1, 1, False), ("conv", 512, 1024, 3, 6, 6), ("relu",), ("conv", 1024, 1024, 1, 0, 1), ("relu",)]
    return out


def _head_lstm_specs():
    out = []
    for sfx in ("", "_reverse"):
        out.append((f"lstm.weight_ih_l0{sfx}", (512, 300), "lstm"))
        out.append((f"lstm.weight_hh_l0{sfx}", (512, 128), "lstm"))
        out.append((f"lstm.bias_ih_l0{sfx}", (512,), "lstm"))
        out.append((f"lstm.bias_hh_l0{sfx}", (512,), "lstm"))
    out.append(("att_reg_box.0.0.weight", (256, FUSED_C, 3, 3), "conv"))
    out.append(("att_reg_box.0.0.bias", (256,), "bias"))
    for i in range(1, 5):
        out.append((f"att_reg_box.{i}.0.weight", (256, 256, 3, 3), "conv"))
        out.append((f"att_reg_box.{i}.0.bias", (256,), "bias"))
    out.append(("att_reg_box.5.weight", (45, 256, 3, 3), "conv"))
    out.append(("att_reg_box.5.bias", (45,), "final_bias"))
    return out


def _conv_bias(name, cin, cout, k):
    return [(name + ".weight", (cout, cin, k, k), "conv"), (name + ".bias", (cout,), "bias")]


def vgg_trainable_specs():
    """SSD-VGG model in forward order: vgg.*, extras.*, fproj1..3 (their gradients are the first of the trunk to
    complete in the backward), LSTM, head."""
    e = "backbone.encoder."
    out = []
    for i, L in enumerate(vgg_layers()):
        if L[0] == "conv":
            out += _conv_bias(f"{e}vgg.{i}", L[1], L[2], L[3])
    for i, (cin, cout, k, _, _) in enumerate(VGG_EXTRAS):
        out += _conv_bias(f"{e}extras.{i}", cin, cout, k)
    for j, cin in enumerate((512, 1024, 512), start=1):
        out += _conv_bias(f"{e}fproj{j}", cin, 256, 1)
    return out + _head_lstm_specs()


def vgg_unused_specs():
    """loc.* / conf.*: the multibox heads SSD builds (ssd_vgg.py:157-171, 21 classes) and ZSGNet never calls."""
    e = "backbone.encoder."
    out = []
    for kind, mult in (("loc", 4), ("conf", 21)):
        for i, (c, nb) in enumerate(zip((512, 1024, 512, 256, 256, 256), VGG_MBOX)):
            out += _conv_bias(f"{e}{kind}.{i}", c, nb * mult, 3)
    return out


def trainable_specs(model="retina"):
    """[(name, shape, kind)] in forward order; `fc` (unused by the path, mdl.py:149-156) comes last."""
    if model == "ssd_vgg":
        return vgg_trainable_specs()
    out = []
    e = "backbone.encoder."
    out.append((e + "conv1.weight", (64, 3, 7, 7), "conv"))
    out += _bn(e + "bn1", 64)
    inpl = 64
    for li, (nblk, width, _) in enumerate(RESNET_LAYERS, start=1):
        for b in range(nblk):
            p = f"{e}layer{li}.{b}."
            out.append((p + "conv1.weight", (width, inpl, 1, 1), "conv"))
            out += _bn(p + "bn1", width)
            out.append((p + "conv2.weight", (width, width, 3, 3), "conv"))
            out += _bn(p + "bn2", width)
            out.append((p + "conv3.weight", (width * 4, width, 1, 1), "conv"))
            out += _bn(p + "bn3", width * 4)
            if b == 0:
                out.append((p + "downsample.0.weight", (width * 4, inpl, 1, 1), "conv"))
                out += _bn(p + "downsample.1", width * 4)
            inpl = width * 4
    f = "backbone.fpn."
    for name, cin, k in (("P5_1", 2048, 1), ("P5_2", 256, 3), ("P4_1", 1024, 1), ("P4_2", 256, 3),
                         ("P3_1", 512, 1), ("P3_2", 256, 3), ("P6", 2048, 3), ("P7_2", 256, 3)):
        out.append((f + name + ".weight", (256, cin, k, k), "conv"))
        out.append((f + name + ".bias", (256,), "bias"))
    for sfx in ("", "_reverse"):
        out.append((f"lstm.weight_ih_l0{sfx}", (512, 300), "lstm"))
        out.append((f"lstm.weight_hh_l0{sfx}", (512, 128), "lstm"))
        out.append((f"lstm.bias_ih_l0{sfx}", (512,), "lstm"))
        out.append((f"lstm.bias_hh_l0{sfx}", (512,), "lstm"))
    out.append(("att_reg_box.0.0.weight", (256, FUSED_C, 3, 3), "conv"))
    out.append(("att_reg_box.0.0.bias", (256,), "bias"))
    for i in range(1, 5):
        out.append((f"att_reg_box.{i}.0.weight", (256, 256, 3, 3), "conv"))
        out.append((f"att_reg_box.{i}.0.bias", (256,), "bias"))
    out.append(("att_reg_box.5.weight", (45, 256, 3, 3), "conv"))
    out.append(("att_reg_box.5.bias", (45,), "final_bias"))
    return out


UNUSED_SPECS = [("backbone.encoder.fc.weight", (1000, 2048), "lin"), ("backbone.encoder.fc.bias", (1000,), "bias")]


def unused_specs(model="retina"):
    return vgg_unused_specs() if model == "ssd_vgg" else list(UNUSED_SPECS)


def buffer_specs(model="retina"):
    out = []
    if model == "ssd_vgg":
        return out                                   # no BatchNorm anywhere (vgg(..., batch_norm=False))
    e = "backbone.encoder."
    out += bn_buffers(e + "bn1", 64)
    for li, (nblk, width, _) in enumerate(RESNET_LAYERS, start=1):
        for b in range(nblk):
            p = f"{e}layer{li}.{b}."
            out += bn_buffers(p + "bn1", width) + bn_buffers(p + "bn2", width) + bn_buffers(p + "bn3", width * 4)
            if b == 0:
                out += bn_buffers(p + "downsample.1", width * 4)
    return out


def reference_param_order(model="retina"):
    """Parameter names in the REFERENCE module's registration order (named_parameters()): torch.optim.Adam's state_dict
    indexes parameters by it (utils.py:479-497 saves optimizer.state_dict()), so the drop-in registers its parameters in
    this order.  retina: encoder incl. the unused fc (torchvision resnet50), fpn as P7_2, P6, P5_1 .. P3_2
    (fpn_resnet.py:113-152), att_reg_box, lstm (mdl.py:208-228).  ssd_vgg: vgg, fproj1-3, extras, loc, conf
    (ssd_vgg.py:31-52), att_reg_box, lstm.  Pinned against the real reference by tests/golden/param_order.json."""
    names = [n for n, _, _ in trainable_specs(model) + unused_specs(model)]
    e = "backbone.encoder."

    def take(prefixes):
        return [n for pre in prefixes for n in names if n.startswith(pre)]
    if model == "ssd_vgg":
        order = take([e + "vgg.", e + "fproj1.", e + "fproj2.", e + "fproj3.", e + "extras.", e + "loc.", e + "conf."])
    else:
        order = [n for n in names if n.startswith(e) and not n.startswith(e + "fc.")] + take([e + "fc."])
        order += take(["backbone.fpn." + k + "." for k in ("P7_2", "P6", "P5_1", "P5_2", "P4_1", "P4_2", "P3_1", "P3_2")])
    order += take(["att_reg_box.", "lstm."])
    assert sorted(order) == sorted(names) and len(set(order)) == len(order)
    return order
