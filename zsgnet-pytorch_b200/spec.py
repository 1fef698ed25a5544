"""Static description of the ZSGNet (ResNet-50 + FPN) parameter set: names and shapes are the
reference's state_dict contract (mdl.py:171-229, fpn_resnet.py:108-152,
torchvision resnet50), listed in FORWARD execution order."""

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


def trainable_specs():
    """[(name, shape, kind)] in forward order; `fc` (unused by the path, mdl.py:149-156) comes last."""
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


def buffer_specs():
    out = []
    e = "backbone.encoder."
    out += bn_buffers(e + "bn1", 64)
    for li, (nblk, width, _) in enumerate(RESNET_LAYERS, start=1):
        for b in range(nblk):
            p = f"{e}layer{li}.{b}."
            out += bn_buffers(p + "bn1", width) + bn_buffers(p + "bn2", width) + bn_buffers(p + "bn3", width * 4)
            if b == 0:
                out += bn_buffers(p + "downsample.1", width * 4)
    return out
