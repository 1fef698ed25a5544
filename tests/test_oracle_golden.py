"""Pin the oracle restatement (oracle/zsg_oracle.py) to outputs of the real reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py from /root/reference)."""
import numpy as np
import pytest
import torch

from oracle import synth, zsg_oracle as zo
from conftest import load_npz

RTOL = 1e-4     # BASELINE.json north_star: fp32 box/loss outputs within 1e-4 relative


def test_anchor_table_bit_exact():
    ref = load_npz("anchors")["anchs"]
    mine = zo.default_anchors().numpy()
    assert mine.dtype == np.float64 and mine.shape == (17460, 4)
    assert np.array_equal(mine, ref)


@pytest.mark.parametrize("name", ["rand8", "adv8", "rand3"])
def test_loss_and_metric_vs_reference(name, golden_meta):
    c = golden_meta["loss_cases"][name]
    z = load_npz("loss_" + name)
    B, seed = c["B"], c["seed"]
    g = torch.Generator().manual_seed(seed)
    batch = synth.make_batch(B, seed=seed, adversarial=c["adv"])
    assert np.array_equal(batch["annot"].numpy(), z["annot"])
    att = (torch.randn(B, synth.NUM_ANCHORS, 1, generator=g) * 1.5 - 3.0).requires_grad_(True)
    bbx = (torch.randn(B, synth.NUM_ANCHORS, 4, generator=g) * 0.7).requires_grad_(True)
    anchs = zo.default_anchors()
    ls = zo.zsg_loss(att, bbx, batch["annot"], anchs)
    # bit-exact integer work
    assert np.array_equal(ls["top1"].numpy(), z["top1"])
    assert np.array_equal(ls["pos"].nonzero().numpy().astype(np.int32), z["pos_idx"])
    for k in ("loss", "cls_ls", "box_ls"):
        assert ls[k].item() == pytest.approx(c[k], rel=1e-6)
    ls["loss"].mean().backward()
    pos = ls["pos"]
    np.testing.assert_allclose(att.grad.squeeze(-1)[pos].numpy(), z["datt_pos"], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(bbx.grad[pos].numpy(), z["dbbx_pos"], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(att.grad.squeeze(-1)[:, ::97].numpy(), z["datt_stride"], rtol=1e-5, atol=1e-12)
    met = zo.evaluate(att.detach(), bbx.detach(), batch["annot"], batch["img_size"], anchs)
    assert met["Acc"].item() == c["Acc"] and met["MaxPos"].item() == c["MaxPos"]
    assert np.array_equal(met["idxs_best"].numpy(), z["best_ids"])
    np.testing.assert_allclose(met["pred_boxes"].numpy(), z["pred_boxes"], rtol=1e-12)
    np.testing.assert_allclose(met["pred_scores"].numpy(), z["best"], rtol=1e-7)


@pytest.mark.parametrize("name", ["net2", "net3v"])
def test_full_network_vs_reference(name, golden_meta):
    c = golden_meta["net_cases"][name]
    z = load_npz(name)
    torch.set_num_threads(max(1, torch.get_num_threads()))
    sd = synth.make_state_dict(0)
    batch = synth.make_batch(c["B"], seed=c["seed"], var_len=c["var_len"])
    ls, met, grads, out, _ = zo.train_step(sd, batch, seed=c["seed"], do_adam=False)
    for k in ("loss", "cls_ls", "box_ls"):
        assert ls[k].item() == pytest.approx(c[k], rel=RTOL)
    assert met["Acc"].item() == c["Acc"] and met["MaxPos"].item() == c["MaxPos"]
    att = out["att_out"].detach().squeeze(-1)
    np.testing.assert_allclose(att[:, ::53].numpy(), z["att_stride"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(out["bbx_out"].detach()[:, ::53].numpy(), z["bbx_stride"], rtol=1e-3, atol=1e-4)
    assert np.array_equal(met["idxs_best"].numpy(), z["best_ids"])
    for k, ref in c["gnorm"].items():
        if ref is None:
            assert grads[k] is None or float(grads[k].abs().sum()) == 0.0, k
        else:
            assert float(grads[k].double().norm()) == pytest.approx(ref, rel=2e-3, abs=1e-7), k
    for key in z.files:
        if key.startswith("g:"):
            np.testing.assert_allclose(grads[key[2:]].numpy(), z[key], rtol=2e-3, atol=1e-5)
        elif key.startswith("gs:"):
            np.testing.assert_allclose(grads[key[3:]].flatten()[::101].numpy(), z[key], rtol=2e-3, atol=1e-5)
    np.testing.assert_allclose(sd["backbone.encoder.bn1.running_mean"].numpy(), z["bn1_rm"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(sd["backbone.encoder.layer4.2.bn3.running_var"].numpy(), z["l4_rv"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("name", ["vgg2", "vgg3v"])
def test_ssd_vgg_network_vs_reference(name, golden_meta):
    """a-8: ZSGNet over the SSD-VGG trunk (ssd_vgg.py:54-102), restatement vs the reference's own modules."""
    c = golden_meta["vgg_cases"][name]
    z = load_npz(name)
    sd = synth.make_state_dict(0, "ssd_vgg")
    batch = synth.make_batch(c["B"], seed=c["seed"], var_len=c["var_len"])
    ls, met, grads, out, _ = zo.train_step(sd, batch, seed=c["seed"], do_adam=False)
    for k in ("loss", "cls_ls", "box_ls"):
        assert ls[k].item() == pytest.approx(c[k], rel=RTOL)
    assert met["Acc"].item() == c["Acc"] and met["MaxPos"].item() == c["MaxPos"]
    att = out["att_out"].detach().squeeze(-1)
    np.testing.assert_allclose(att[:, ::53].numpy(), z["att_stride"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(out["bbx_out"].detach()[:, ::53].numpy(), z["bbx_stride"], rtol=1e-3, atol=1e-4)
    assert np.array_equal(met["idxs_best"].numpy(), z["best_ids"])
    for k, ref in c["gnorm"].items():
        if ref is None:                                        # loc.* / conf.*: built, never called
            assert grads[k] is None or float(grads[k].abs().sum()) == 0.0, k
        else:
            assert float(grads[k].double().norm()) == pytest.approx(ref, rel=2e-3, abs=1e-7), k
    for key in z.files:
        if key.startswith("g:"):
            np.testing.assert_allclose(grads[key[2:]].numpy(), z[key], rtol=2e-3, atol=1e-6)
        elif key.startswith("gs:"):
            np.testing.assert_allclose(grads[key[3:]].flatten()[::101].numpy(), z[key], rtol=2e-3, atol=1e-6)


def test_loop_and_batched_lstm_agree():
    sd = synth.make_state_dict(0)
    batch = synth.make_batch(5, seed=3, var_len=True)
    torch.manual_seed(1)
    h0, c0 = zo.draw_h0c0(5)
    a = zo.lstm_query(sd, batch["qvec"], batch["qlens"], h0, c0)
    b = zo.lstm_query_batched(sd, batch["qvec"], batch["qlens"], h0, c0)
    np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-5, atol=1e-6)


def test_upsample_index_table(golden_meta):
    """fpn_resnet.py:161,166: nearest source index = floor(dst * in/out)."""
    for key, ref in golden_meta["upsample_idx"].items():
        i, o = (int(v) for v in key.split("->"))
        mine = [min(int(np.floor(np.float32(d) * np.float32(i / o))), i - 1) for d in range(o)]
        assert mine == ref


def test_oracle_nan_guard_matches_reference_on_zero_area_box():
    """loss.py:128-133 through loss.py:92 (inf * 0): the dump of the real reference (tests/golden/make_golden_extra.py)."""
    import numpy as np
    from oracle import synth, zsg_oracle as zo
    z = load_npz("loss_zero_area")
    B, A = 2, synth.NUM_ANCHORS
    batch = synth.make_batch(B, seed=4)
    batch["annot"][0] = torch.tensor([0.1, 0.1, 0.1, 0.4])
    g = torch.Generator().manual_seed(4)
    att = (torch.randn(B, A, 1, generator=g) - 3.0).requires_grad_(True)
    bbx = (torch.randn(B, A, 4, generator=g) * 0.3).requires_grad_(True)
    ls = zo.zsg_loss(att, bbx, batch["annot"], zo.default_anchors())
    assert [ls["loss"].item(), ls["cls_ls"].item(), ls["box_ls"].item()] == z["losses"].tolist() == [1.01, 1.0, 0.01]
    assert np.array_equal(ls["top1"].numpy(), z["top1"])
    ls["loss"].backward()
    assert att.grad is None and bbx.grad is None and float(z["datt_abs_sum"]) == 0.0


def test_reference_parameter_registration_order():
    """spec.reference_param_order (the order ZSGNet registers its parameters in, hence torch.optim.Adam's state indices)
    against named_parameters() of the real reference modules (tests/golden/make_param_order.py)."""
    import json
    import os
    from zsg_b200 import spec
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "param_order.json")) as f:
        gold = json.load(f)
    for model in ("retina", "ssd_vgg"):
        assert spec.reference_param_order(model) == [n for n, _ in gold[model]]
        shapes = {n: tuple(s) for n, s, _ in spec.trainable_specs(model) + spec.unused_specs(model)}
        assert all(shapes[n] == tuple(s) for n, s in gold[model])
