"""Host logic of the engine without a GPU: build the launch program on CPU tensors with the C-ABI call
replaced by a recorder, and check the launch sequence, buffer bookkeeping and gradient-bucket marks."""
import collections

import pytest
import torch


@pytest.fixture()
def recorded(monkeypatch):
    import zsg_b200
    from zsg_b200 import _lib, ops, engine, spec
    calls = []

    def fake_call(name, *args):
        calls.append(name)
        return 0
    monkeypatch.setattr(ops, "call", fake_call)
    monkeypatch.setattr(ops, "stream", lambda: 0)
    monkeypatch.setattr(ops, "match_loss_workspace", lambda b, d: torch.empty(64, dtype=torch.float64))
    return calls, engine, spec, ops


def test_engine_program_builds_and_runs_on_host(recorded):
    calls, engine, spec, ops = recorded
    B, T = 2, 5
    store = engine.ParamStore(torch.device("cpu"))
    assert store.used % 64 == 0 and store.total > store.used
    bufs = {}
    for name, shp in spec.buffer_specs():
        bufs[name] = torch.zeros(shp if len(shp) else (), dtype=torch.float32 if len(shp) else torch.long)
    eng = engine.Engine(store, bufs, B, T, torch.device("cpu"))
    assert len(eng.bns) == 53                                         # torchvision resnet50 has 53 BatchNorm2d
    img = torch.rand(B, 3, 300, 300)
    qvec = torch.randn(B, 4, 300)
    lens = torch.tensor([4.0, 2.0])
    inv = torch.tensor([0, 1])
    eng.set_inputs(img, qvec, lens, inv, torch.randn(2, B, 128), torch.randn(2, B, 128))
    del calls[:]
    out = eng.forward(training=True)
    assert out.shape == (B, spec.NUM_ANCHORS, 5)
    fwd = collections.Counter(calls)
    # 53 trunk convs + 8 FPN + 6 head + 1 LSTM projection + the language GEMM of the split first head conv (lang x W_l)
    assert fwd["zsg_conv_fwd"] == 53 + 8 + 6 + 1 + 1
    # BatchNorm statistics come out of the conv epilogues (per-row-group partials), not from a pass over the activations
    assert fwd["zsg_bn_stats"] == 0 and fwd["zsg_bn_finalize_partials"] == 53
    assert fwd["zsg_bn_finalize"] == 0 and fwd["zsg_bn_apply"] == 16
    # a-6: the [feat | lang | grid] tensor is never built: three channel slices of the 514-channel weight, border-class sums
    # of lang x W_l and the grid term, added per row in the conv epilogue (zsg_conv_params.row_add)
    assert fwd["zsg_split_tf32"] == 2 and fwd["zsg_pad_channels"] == 1 and fwd["zsg_copy_cols"] == 3
    assert fwd["zsg_fuse_lang_grid"] == 0 and fwd["zsg_head0_lang_grid_terms"] == 1
    assert fwd["zsg_lstm_fwd_dir"] == 1 and fwd["zsg_lstm_rev_step"] == 1
    del calls[:]
    eng.forward(training=False)
    ev = collections.Counter(calls)
    assert ev["zsg_bn_finalize_partials"] == 0 and ev["zsg_bn_eval_affine"] == 53
    del calls[:]
    seen = []
    eng.backward(torch.zeros(B, spec.NUM_ANCHORS, 5), on_bucket=lambda lo, hi: seen.append((lo, hi)))
    bwd = collections.Counter(calls)
    # every conv has a wgrad (+5 LSTM weight gradients... 4 LSTM matrices), every conv but the stem a dgrad
    assert bwd["zsg_conv_wgrad"] == 53 + 8 + 6 + 4 + 1                 # + dW_l of the split first head conv
    assert bwd["zsg_unfuse_lang_grid"] == 0 and bwd["zsg_head0_backward_sums"] == 1 and bwd["zsg_copy_cols"] == 2
    # data gradients run through the forward kernel; the five 3x3 / stride-2 convs (layer2-4.0.conv2, P6, P7_2) take
    # one launch per input-pixel parity class (4) instead of one zero-stuffed launch
    assert bwd["zsg_conv_fwd"] == 52 + 8 + 6 + 5 * 3 + 1               # + d lang = tap sums x W_l
    # BatchNorm backward: the reduce pass of the 32 in-block BatchNorm+ReLU pairs (bn1 / bn2) comes out of the epilogue of the
    # data gradient that writes their dy (zsg_conv_params.bnb_*), except behind the three stride-2 conv2 (parity-class launches)
    assert bwd["zsg_bn_bwd_reduce"] == 53 - 29 and bwd["zsg_bn_bwd_apply"] == 53
    assert bwd["zsg_bn_bwd_center_sums"] == 29 and bwd["zsg_bn_stats_partials"] == 29
    assert sum(1 for op in eng.bwd if isinstance(op, ops.ConvOp) and op.p.bnb_partials) == 29
    # 64 of the transposed-flipped weight copies (all that go arena -> pool) are one batched launch; the padded last head
    # weight and the two slices of the first one (W_f for d feat, W_l for d lang) keep their own
    assert bwd["zsg_weight_transpose_flip_batched32"] == 1 and bwd["zsg_weight_transpose_flip"] == 3   # tiled: all dims % 32 == 0
    assert len(eng._wtf) == 52 + 8 + 4 and bwd["zsg_split_tf32"] == 1
    # buckets: contiguous, ordered, covering the used arena exactly once
    assert seen[0][0] == 0 and seen[-1][1] == store.used
    for (a, b), (c, d) in zip(seen, seen[1:]):
        assert b == c and a < b
    # 20 backward stages (head, lstm, fpn, 16 blocks, stem) coalesced into buckets of >= 4 Mi elements: one graph
    # segment + one all-reduce each
    assert len(eng.bucket_marks) == 16 + 4 and len(seen) == len(eng.segments) == 6
    assert all(b - a >= eng.bucket_elems for a, b in seen[:-1])
    assert [s[0] for s in eng.segments] == [0] + [s[1] for s in eng.segments[:-1]] and eng.segments[-1][1] == len(eng.bwd)


def test_bf16_engine_program_on_host(recorded):
    """dtype='bf16' (BASELINE configs[2..4]): every convolution reads bf16 images (the stem through an 8-channel bf16 input
    image); only the 300-wide LSTM projection keeps the fp32 (3xTF32) path; BatchNorm passes write bf16 images directly."""
    calls, engine, spec, ops = recorded
    B, T = 2, 5
    store = engine.ParamStore(torch.device("cpu"))
    bufs = {}
    for name, shp in spec.buffer_specs():
        bufs[name] = torch.zeros(shp if len(shp) else (), dtype=torch.float32 if len(shp) else torch.long)
    eng = engine.Engine(store, bufs, B, T, torch.device("cpu"), dtype="bf16")
    f32 = engine.Engine(store, bufs, B, T, torch.device("cpu"))
    kinds = collections.Counter(it[1].kernel for it in eng.fwd if it[0] == "op")
    assert kinds == {"conv_bf16_kernel": 53 + 8 + 6 + 1, "conv_tc_async_kernel": 1}  # + lang x W_l; the LSTM projection stays fp32
    bk = collections.Counter(op.kernel for op in eng.bwd if isinstance(op, (ops.ConvOp, ops.WgradOp)))
    assert bk["wgrad_bf16_kernel"] == 53 + 8 + 6 and bk["wgrad_tc_async_kernel"] == 0 and bk["wgrad_tc_kernel"] == 4 + 1   # + dW_l
    assert bk["conv_bf16_kernel"] == 52 + 8 + 6 + 5 * 3 + 1 and "conv_tc_async_kernel" not in bk                          # + d lang
    # no materialised BN-ReLU tensors and half-size images: the activation-side buffers shrink (at B = 2 the weight
    # images dominate both engines, so compare without them)
    wimg = lambda e: 4 * (2 * store.total + 3 * e.pool_n) + (2 * (store.total + e.pool_n) if e.bf16 else 0)
    assert eng.nbytes - wimg(eng) < 0.85 * (f32.nbytes - wimg(f32))
    eng.set_inputs(torch.rand(B, 3, 300, 300), torch.randn(B, 4, 300), torch.tensor([4.0, 2.0]), torch.tensor([0, 1]),
                   torch.randn(2, B, 128), torch.randn(2, B, 128))
    del calls[:]
    eng.forward(training=True)
    fwd = collections.Counter(calls)
    # bf16 storage of the trunk: conv outputs / block outputs are bfloat16 tensors (their own operand images): the 32
    # BatchNorm+ReLU-on-load passes are bfloat16 -> bfloat16, the stem pool output needs no cast at all; fp32 tensors that
    # feed a GEMM (2 weight arenas, 4 FPN inner maps incl. relu(P6), 2 head inputs: feat, lang) are still cast; the images of
    # the 5 hidden maps of the head come out of the producing conv's epilogue (zsg_conv_params.y_img_bf16)
    assert eng.b16act and eng.dbg["c5"].dtype == torch.bfloat16 and eng.dbg["blocks"][0]["r1"].dtype == torch.bfloat16
    assert fwd["zsg_act_b16"] == 32 and fwd["zsg_cast_bf16"] == 2 + 4 + 2 and fwd["zsg_split_act"] == 1
    assert sum(1 for it in eng.fwd if it[0] == "op" and it[1].p.y_img_bf16) == 5
    assert fwd["zsg_bn_apply_b16"] == 16 and fwd["zsg_bn_apply_bf16"] == 0 and fwd["zsg_maxpool_bn_relu_fwd_b16"] == 1
    del calls[:]
    eng.backward(torch.zeros(B, spec.NUM_ANCHORS, 5))
    bwd = collections.Counter(calls)
    # (the bf16 engine keeps the separate reduce passes by default: fusing them into the data gradients does not move its step)
    assert bwd["zsg_bn_bwd_apply_b16"] == 53 and bwd["zsg_bn_bwd_reduce_b16"] == 53 and bwd["zsg_bn_bwd_center_sums"] == 0
    assert bwd["zsg_bn_bwd_apply"] == 0 and bwd["zsg_bn_bwd_apply_bf16"] == 0 and bwd["zsg_split_tf32"] == 0


def test_ssd_vgg_program_builds_and_runs_on_host(recorded):
    """a-8 (config 5): launch program of the SSD-VGG trunk (ssd_vgg.py:54-102) + shared head."""
    calls, engine, spec, ops = recorded
    B, T = 1, 4
    store = engine.ParamStore(torch.device("cpu"), "ssd_vgg")
    assert store.offsets["backbone.encoder.loc.0.weight"] >= store.used          # unused multibox heads: no all-reduce / Adam
    eng = engine.Engine(store, {}, B, T, torch.device("cpu"))
    assert len(eng.bns) == 0
    eng.set_inputs(torch.rand(B, 3, 300, 300), torch.randn(B, 4, 300), torch.tensor([4.0]), torch.tensor([0]),
                   torch.randn(2, B, 128), torch.randn(2, B, 128))
    del calls[:]
    out = eng.forward(training=True)
    assert out.shape == (B, spec.NUM_ANCHORS, 5)
    fwd = collections.Counter(calls)
    # 15 VGG convs + 8 extras + 3 fproj + 6 head + 1 LSTM projection + lang x W_l of the split first head conv
    assert fwd["zsg_conv_fwd"] == 15 + 8 + 3 + 6 + 1 + 1
    assert fwd["zsg_maxpool_fwd"] == 5 and fwd["zsg_l2norm_fwd"] == 1 and fwd["zsg_bn_stats"] == 0
    del calls[:]
    seen = []
    eng.backward(torch.zeros(B, spec.NUM_ANCHORS, 5), on_bucket=lambda lo, hi: seen.append((lo, hi)))
    bwd = collections.Counter(calls)
    assert bwd["zsg_conv_wgrad"] == 15 + 8 + 3 + 6 + 4 + 1
    # every conv but vgg.0 has a data gradient; extras.1 / extras.3 (3x3, stride 2) take 4 parity-class launches each
    assert bwd["zsg_conv_fwd"] == 14 + 8 + 3 + 6 + 2 * 3 + 1
    assert bwd["zsg_maxpool_bwd"] == 5 and bwd["zsg_l2norm_bwd"] == 1 and bwd["zsg_relu_bwd"] == 3
    assert seen[0][0] == 0 and seen[-1][1] == store.used
    for (a, b), (c, d) in zip(seen, seen[1:]):
        assert b == c and a < b
    assert len(eng.bucket_marks) == 6 + 2 + 2                                   # VGG segments, extras, fproj, lstm, head
    assert 2 <= len(seen) == len(eng.segments) <= 6                             # coalesced to >= 4 Mi elements each


def test_param_store_views_follow_reference_shapes(recorded):
    _, engine, spec, _ = recorded
    store = engine.ParamStore(torch.device("cpu"))
    v = store.view("backbone.encoder.layer1.0.conv2.weight")
    assert v.shape == (64, 64, 3, 3) and v.stride() == (576, 1, 192, 64)      # OIHW view of [O][H][W][I] storage
    assert store.view("att_reg_box.0.0.weight").shape == (256, 514, 3, 3)
    g = store.grad_view("lstm.weight_ih_l0")
    assert g.shape == (512, 300) and g.is_contiguous()
    # the unused fc sits behind the all-reduce / Adam range
    assert store.offsets["backbone.encoder.fc.weight"] >= store.used
    # head parameters come first (their gradients are final first)
    assert store.offsets["att_reg_box.5.bias"] == 0
