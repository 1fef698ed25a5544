"""End-to-end parity of the CUDA hot path through the reference-facing API (mdl / loss / evaluator):
against the golden dumps of the real reference (tests/golden, B=2 and ragged B=3) and against the
CPU oracle executed live.

Tolerances
  * indices / assignments (positives, IoU argmax, predicted anchor id, Acc): bit-exact;
  * fp32 losses: 1e-4 relative (BASELINE.json north_star);
  * individual head outputs and end-to-end gradients: judged against the MEASURED fp32 noise floor of
    this 50-layer train-mode-BatchNorm network (tools/noise_floor.py, tools/diag_grads.py; numbers in
    DESIGN.md): PyTorch's own CUDA fp32 path differs from the CPU fp32 reference by 1.0e-4 rms at C5 and by
    a median 3.6 % per parameter gradient, so the gradient check asserts "as close to the reference as
    torch's CUDA fp32 autograd is", computed live in the same test.  Every backward kernel is checked
    on its own at 1e-4..1e-5 in tests/test_kernels_gpu.py."""
import numpy as np
import pytest
import torch

from conftest import load_npz

pytestmark = pytest.mark.gpu
RTOL = 1e-4


@pytest.fixture(scope="module")
def stack():
    assert torch.cuda.is_available()
    import zsg_b200
    from zsg_b200 import mdl, loss, evaluator
    from oracle import synth
    cfg = synth.default_cfg()
    cfg["device"] = "cuda"
    ratios, scales = synth.ratios_scales(cfg)
    net = mdl.get_default_net(num_anchors=9, cfg=cfg)
    crit = loss.get_default_loss(ratios, scales, cfg)
    ev = evaluator.get_default_eval(ratios, scales, cfg)
    return net, crit, ev, synth


def to_dev(batch):
    return {k: v.cuda() for k, v in batch.items()}


def run_step(net, crit, ev, synth, B, seed, var_len, train=True):
    net.load_state_dict(synth.make_state_dict(0), strict=True)
    net.train(train)
    net.zero_grad()
    batch = to_dev(synth.make_batch(B, seed=seed, var_len=var_len))
    torch.manual_seed(seed)                          # h0/c0 come from the global CPU RNG (mdl.py:279-294)
    out = net(batch)
    ls = crit(out, batch)
    if train:
        ls["loss"].mean().backward()
    met = ev(out, batch)
    torch.cuda.synchronize()
    return batch, out, ls, met


def test_state_dict_contract(stack):
    net, _, _, synth = stack
    sd = net.state_dict()
    ref = synth.make_state_dict(0)
    assert set(sd) == set(ref) and len(sd) == 356
    for k in ref:
        assert tuple(sd[k].shape) == tuple(ref[k].shape), k
    net.load_state_dict(ref, strict=True)
    back = net.state_dict()
    for k in ("backbone.encoder.conv1.weight", "att_reg_box.0.0.weight", "lstm.weight_hh_l0_reverse",
              "backbone.encoder.layer3.4.bn2.running_var", "backbone.encoder.fc.bias"):
        assert torch.equal(back[k].cpu(), ref[k]), k


@pytest.mark.parametrize("name", ["net2", "net3v"])
def test_train_step_vs_reference_golden(stack, golden_meta, name):
    net, crit, ev, synth = stack
    c = golden_meta["net_cases"][name]
    z = load_npz(name)
    batch, out, ls, met = run_step(net, crit, ev, synth, c["B"], c["seed"], c["var_len"])
    assert out["att_out"].shape == (c["B"], 17460, 1) and out["bbx_out"].shape == (c["B"], 17460, 4)
    assert out["feat_sizes"].tolist() == [[38, 38], [19, 19], [10, 10], [5, 5], [3, 3], [1, 1]]
    for k in ("loss", "cls_ls", "box_ls"):
        assert ls[k].item() == pytest.approx(c[k], rel=RTOL), k
    assert ls["loss"].dtype == torch.float64 and ls["cls_ls"].dtype == torch.float32
    assert met["Acc"].item() == c["Acc"] and met["MaxPos"].item() == c["MaxPos"]
    assert np.array_equal(met["best_ids"].cpu().numpy(), z["best_ids"])
    att = out["att_out"].detach().squeeze(-1).cpu()
    # head outputs: rms error relative to the tensor's rms, 3e-4 = 3x the fp32 noise floor measured at C5
    def rms_rel(a, b):
        a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
        return np.sqrt(((a - b) ** 2).mean() / (b ** 2).mean())
    assert rms_rel(att[:, ::53].numpy(), z["att_stride"]) < 3e-4
    assert rms_rel(out["bbx_out"].detach()[:, ::53].cpu().numpy(), z["bbx_stride"]) < 3e-4
    np.testing.assert_allclose(att[:, ::53].numpy(), z["att_stride"], rtol=2e-3, atol=1e-3)
    np.testing.assert_allclose(out["bbx_out"].detach()[:, ::53].cpu().numpy(), z["bbx_stride"], rtol=2e-3, atol=1e-3)
    np.testing.assert_allclose(met["pred_boxes"].cpu().numpy(), z["pred_boxes"], rtol=1e-3, atol=5e-2)
    # gradients: unused parameters stay None; norms within the chaotic-fp32 band (see module docstring)
    grads = {k: p.grad for k, p in net.named_parameters()}
    for k, ref in c["gnorm"].items():
        if ref is None:
            assert grads[k] is None, k
        else:
            assert float(grads[k].double().norm()) == pytest.approx(ref, rel=0.1, abs=1e-7), k
    for key in z.files:
        if key.startswith("g:"):
            assert rms_rel(grads[key[2:]].cpu().numpy(), z[key]) < 0.15, key
        elif key.startswith("gs:"):
            assert rms_rel(grads[key[3:]].cpu().flatten()[::101].numpy(), z[key]) < 0.15, key
    sd = net.state_dict()
    np.testing.assert_allclose(sd["backbone.encoder.bn1.running_mean"].cpu().numpy(), z["bn1_rm"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(sd["backbone.encoder.layer4.2.bn3.running_var"].cpu().numpy(), z["l4_rv"], rtol=1e-4,
                               atol=1e-6)
    assert int(sd["backbone.encoder.bn1.num_batches_tracked"]) == 1


def test_train_step_vs_live_oracle_b4(stack):
    net, crit, ev, synth = stack
    from oracle import zsg_oracle as zo
    B, seed = 4, 31
    batch, out, ls, met = run_step(net, crit, ev, synth, B, seed, True)
    sd = synth.make_state_dict(0)
    ols, omet, ograds, oout, _ = zo.train_step(sd, synth.make_batch(B, seed=seed, var_len=True), seed=seed, do_adam=False)
    for k in ("loss", "cls_ls", "box_ls"):
        assert ls[k].item() == pytest.approx(ols[k].item(), rel=RTOL), k
    assert torch.equal(crit.last_top1.cpu(), ols["top1"])
    assert torch.equal(crit.last_pos.cpu().bool(), ols["pos"])
    assert torch.equal(met["best_ids"].cpu(), omet["idxs_best"])
    assert met["Acc"].item() == omet["Acc"].item()
    # noise floor: the same oracle functions executed by PyTorch on CUDA in fp32 (TF32 off)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sdg = {k: v.cuda() for k, v in synth.make_state_dict(0).items()}
    _, _, cgrads, _, _ = zo.train_step(sdg, batch, seed=seed, do_adam=False)
    mine_e, cuda_e = [], []
    for k, g in ograds.items():
        if g is None:
            assert net.get_parameter(k).grad is None, k
            continue
        r = g.double()
        n = r.norm().clamp_min(1e-30)
        mine_e.append(float((net.get_parameter(k).grad.cpu().double() - r).norm() / n))
        cuda_e.append(float((cgrads[k].cpu().double() - r).norm() / n))
    mine_e, cuda_e = np.array(mine_e), np.array(cuda_e)
    print(f"gradient error vs CPU fp32 reference: zsg_b200 median {np.median(mine_e):.3e} max {mine_e.max():.3e}; "
          f"torch CUDA fp32 median {np.median(cuda_e):.3e} max {cuda_e.max():.3e}")
    assert np.median(mine_e) < 2.0 * np.median(cuda_e) + 1e-3
    assert mine_e.max() < 3.0 * cuda_e.max() + 1e-3


def test_eval_mode_uses_running_stats(stack):
    net, crit, ev, synth = stack
    from oracle import zsg_oracle as zo
    B, seed = 2, 5
    batch, out, ls, met = run_step(net, crit, ev, synth, B, seed, False, train=False)
    sd = synth.make_state_dict(0)
    torch.manual_seed(seed)
    oout = zo.zsgnet_forward(sd, synth.make_batch(B, seed=seed), training=False)
    a, b = out["att_out"].cpu().flatten(), oout["att_out"].flatten()
    assert float((a - b).abs().max()) < 2e-4 * float(b.abs().max())
    assert int(net.state_dict()["backbone.encoder.bn1.num_batches_tracked"]) == 0


def test_full_size_bs64_properties(stack):
    """BASELINE configs[1] size (bs = 64, 300x300, qlen 20), where the CPU oracle of the whole network is too slow:
      * eval-mode outputs of a sample do not depend on what else is in the batch (BatchNorm uses running stats), so rows
        0..3 of the bs=64 launch geometry must reproduce a bs=4 forward of the same samples;
      * the forward pass is bit-reproducible run to run (per-issuer TMEM accumulators, fixed promotion order);
      * matching / loss / metric on the bs=64 head output agree with the CPU oracle of that stage (indices bit-exact,
        losses 1e-4), which is cheap at any batch size."""
    net, crit, ev, synth = stack
    from oracle import zsg_oracle as zo
    B, seed = 64, 77
    net.load_state_dict(synth.make_state_dict(0), strict=True)
    net.train(False)
    big = synth.make_batch(B, seed=seed)
    # the LSTM initial states are drawn per forward from the CPU RNG in sorted-row order (mdl.py:279-294); give every
    # sample the same state so that the comparison below does not depend on batch size or tie order of the sort
    g = torch.Generator().manual_seed(5)
    base = [torch.randn(2, 1, 128, generator=g), torch.randn(2, 1, 128, generator=g)]
    orig, calls = torch.randn, [0]

    def fake_randn(*a, **k):
        if len(a) == 3 and a[0] == 2 and a[2] == 128 and not k:
            calls[0] += 1
            return base[(calls[0] - 1) % 2].expand(2, a[1], 128).clone()
        return orig(*a, **k)

    torch.randn = fake_randn
    try:
        outs = []
        for _ in range(2):
            out = net(to_dev(big))
            outs.append((out["att_out"].clone(), out["bbx_out"].clone()))
        small = {k: v[:4].clone() for k, v in big.items()}
        out4 = net(to_dev(small))
        out4 = (out4["att_out"].clone(), out4["bbx_out"].clone())
        torch.cuda.synchronize()
        assert calls[0] == 6
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
        for i in (0, 1):
            a, b = outs[0][i][:4].cpu(), out4[i].cpu()
            assert float((a - b).abs().max()) <= 2e-5 * float(b.abs().max()), i
        # loss + metric stage against the oracle at full size (train-mode forward: batch statistics)
        net.train(True)
        dbatch = to_dev(big)
        out = net(dbatch)
        ls = crit(out, dbatch)
        met = ev(out, dbatch)
        torch.cuda.synchronize()
    finally:
        torch.randn = orig
    anchs = zo.default_anchors()
    att, bbx = out["att_out"].detach().cpu(), out["bbx_out"].detach().cpu()
    ols = zo.zsg_loss(att, bbx, big["annot"], anchs)
    omet = zo.evaluate(att, bbx, big["annot"], big["img_size"], anchs)
    for k in ("loss", "cls_ls", "box_ls"):
        assert ls[k].item() == pytest.approx(ols[k].item(), rel=RTOL), k
    assert torch.equal(crit.last_top1.cpu(), ols["top1"]) and torch.equal(crit.last_pos.cpu().bool(), ols["pos"])
    assert torch.equal(met["best_ids"].cpu(), omet["idxs_best"]) and met["Acc"].item() == omet["Acc"].item()
    assert int(crit.last_pos.sum(1).min()) >= 1                          # every row keeps its top-1 anchor (loss.py:80-87)


# ------------------------------------------------------------------------------- a-8: SSD-VGG trunk (config 5)
@pytest.fixture(scope="module")
def vgg_stack():
    assert torch.cuda.is_available()
    import zsg_b200
    from zsg_b200 import mdl, loss, evaluator
    from oracle import synth
    cfg = synth.default_cfg("ssd_vgg")
    cfg["device"] = "cuda"
    ratios, scales = synth.ratios_scales(cfg)
    net = mdl.get_default_net(num_anchors=9, cfg=cfg)
    return net, loss.get_default_loss(ratios, scales, cfg), evaluator.get_default_eval(ratios, scales, cfg), synth


def rms_rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.sqrt(((a - b) ** 2).mean() / max((b ** 2).mean(), 1e-60))


def run_vgg_step(net, crit, ev, synth, B, seed, var_len):
    net.load_state_dict(synth.make_state_dict(0, "ssd_vgg"), strict=True)
    net.train()
    net.zero_grad()
    batch = to_dev(synth.make_batch(B, seed=seed, var_len=var_len))
    torch.manual_seed(seed)
    out = net(batch)
    ls = crit(out, batch)
    ls["loss"].mean().backward()
    met = ev(out, batch)
    torch.cuda.synchronize()
    return batch, out, ls, met


def test_ssd_vgg_state_dict_contract(vgg_stack):
    net, _, _, synth = vgg_stack
    sd, ref = net.state_dict(), synth.make_state_dict(0, "ssd_vgg")
    assert set(sd) == set(ref)                      # key names verified against the reference's modules by make_golden.py
    for k in ref:
        assert tuple(sd[k].shape) == tuple(ref[k].shape), k
    net.load_state_dict(ref, strict=True)
    back = net.state_dict()
    for k in ("backbone.encoder.vgg.0.weight", "backbone.encoder.vgg.31.bias", "backbone.encoder.extras.5.weight",
              "backbone.encoder.fproj2.weight", "backbone.encoder.conf.3.weight"):
        assert torch.equal(back[k].cpu(), ref[k]), k


@pytest.mark.parametrize("name", ["vgg2", "vgg3v"])
def test_ssd_vgg_train_step_vs_reference_golden(vgg_stack, golden_meta, name):
    """ZSGNet over the SSD-VGG trunk against the dump of the reference's own modules (tests/golden/make_golden.py vgg).
    No BatchNorm in this model, so gradients are held to 1e-3 rms per tensor (max-pool / ReLU ties are the noise)."""
    net, crit, ev, synth = vgg_stack
    c = golden_meta["vgg_cases"][name]
    z = load_npz(name)
    batch, out, ls, met = run_vgg_step(net, crit, ev, synth, c["B"], c["seed"], c["var_len"])
    assert out["att_out"].shape == (c["B"], 17460, 1) and out["bbx_out"].shape == (c["B"], 17460, 4)
    for k in ("loss", "cls_ls", "box_ls"):
        assert ls[k].item() == pytest.approx(c[k], rel=RTOL), k
    assert met["Acc"].item() == c["Acc"] and met["MaxPos"].item() == c["MaxPos"]
    assert np.array_equal(met["best_ids"].cpu().numpy(), z["best_ids"])
    att = out["att_out"].detach().squeeze(-1).cpu()
    assert rms_rel(att[:, ::53].numpy(), z["att_stride"]) < 1e-5
    assert rms_rel(out["bbx_out"].detach()[:, ::53].cpu().numpy(), z["bbx_stride"]) < 1e-4
    np.testing.assert_allclose(att[:, ::53].numpy(), z["att_stride"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(out["bbx_out"].detach()[:, ::53].cpu().numpy(), z["bbx_stride"], rtol=1e-4, atol=1e-5)
    # Gradients.  The model has no BatchNorm, but five max-pools and 26 ReLUs: where the two largest candidates of a
    # pooling window (or a pre-activation and zero) are closer than the fp32 rounding noise of the forward pass
    # (~1e-5 relative after 13 convs), the winner differs from the CPU reference's and the gradient of that window
    # moves to the neighbouring pixel.  Measured (tools/vgg_debug2.py, B=2): 3 of 369,664 pool5 windows flip, which
    # alone is 1e-2 of ||d conv5_3|| although pool5's own backward equals torch's bit for bit on the same inputs.
    # A ReLU flip in conv6/7 or the extras does the same on a smaller scale (vgg3v: 2.5e-4 on vgg.31/33).
    # So: head and LSTM at 1e-4, the trunk at the flip floor with a 1e-3 median; every backward kernel is checked on
    # its own at 1e-5 in tests/test_kernels_gpu.py.
    grads = {k: p.grad for k, p in net.named_parameters()}
    tight = ("att_reg_box.", "lstm.")
    for k, ref in c["gnorm"].items():
        if ref is None:
            assert grads[k] is None, k               # loc.* / conf.*: never called
        else:
            assert float(grads[k].double().norm()) == pytest.approx(ref, rel=1e-4 if k.startswith(tight) else 2e-2, abs=1e-12), k
    worst = {}
    for key in z.files:
        if key.startswith("g:"):
            worst[key[2:]] = rms_rel(grads[key[2:]].cpu().numpy(), z[key])
        elif key.startswith("gs:"):
            worst[key[3:]] = rms_rel(grads[key[3:]].cpu().flatten()[::101].numpy(), z[key])
    print({k: f"{v:.2e}" for k, v in worst.items()})
    for k, v in worst.items():
        assert v < (1e-4 if k.startswith(tight) else 5e-2), (k, v)
    assert float(np.median(list(worst.values()))) < 1e-3, worst


# ------------------------------------------------------------------------------- f-3: checkpoints in the reference's format
def test_checkpoint_round_trip_reference_format(stack, tmp_path):
    """utils.py:440-497: {model_state_dict, optimizer_state_dict, scheduler_state_dict, num_it, num_epoch, cfgtxt,
    best_met}.  The optimizer state must load into a stock torch.optim.Adam (the reference's optimiser) and resuming
    from the file must continue like the original (restored state bit for bit; the next step up to the fp32 atomics of
    the split-K weight gradients, whose summation order varies from run to run)."""
    net, crit, ev, synth = stack
    from zsg_b200 import checkpoint, mdl, optim
    net.load_state_dict(synth.make_state_dict(0), strict=True)
    net.train()
    opt = optim.FusedAdam(net.parameters(), lr=1e-4, net=net)

    def one_step(n, o, seed):
        batch = to_dev(synth.make_batch(2, seed=seed))
        torch.manual_seed(seed)
        o.zero_grad()
        crit(n(batch), batch)["loss"].mean().backward()
        o.step()
    one_step(net, opt, 1)
    one_step(net, opt, 2)
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, factor=0.1, patience=2)
    path = tmp_path / "ck.pth"
    checkpoint.save_model_dict(path, net, opt, sched, num_it=2, num_epoch=1, best_met=0.5, cfg=synth.default_cfg(),
                               ddp_prefix=True)
    ck = torch.load(path, weights_only=False)
    assert set(ck) == {"model_state_dict", "optimizer_state_dict", "scheduler_state_dict", "num_it", "num_epoch", "cfgtxt",
                       "best_met"}
    assert all(k.startswith("module.") for k in ck["model_state_dict"])
    w = ck["model_state_dict"]["module.backbone.encoder.layer1.0.conv2.weight"]
    assert w.shape == (64, 64, 3, 3) and w.is_contiguous()
    # reference side: a stock Adam over same-shaped parameters accepts the optimizer state
    ref_params = [torch.nn.Parameter(p.detach().cpu().clone()) for p in net.parameters()]
    adam = torch.optim.Adam(ref_params, lr=1e-4, betas=(0.9, 0.99))
    adam.load_state_dict(ck["optimizer_state_dict"])
    assert int(adam.state[ref_params[0]]["step"]) == 2
    assert adam.state[ref_params[0]]["exp_avg"].shape == ref_params[0].shape
    # resume into a fresh net + optimizer, then one more identical step on both
    cfg = synth.default_cfg()
    cfg["device"] = "cuda"
    net2 = mdl.get_default_net(num_anchors=9, cfg=cfg)
    net2.train()
    opt2 = optim.FusedAdam(net2.parameters(), lr=3e-3, net=net2)
    info = checkpoint.load_model_dict(path, net2, opt2)
    assert info == {"num_it": 2, "num_epoch": 1, "best_met": 0.5}
    used = net.store.used
    assert opt2.t == 2 and opt2.param_groups[0]["lr"] == 1e-4
    assert torch.equal(net2.store.param_arena, net.store.param_arena)
    assert torch.equal(opt2.m, opt.m) and torch.equal(opt2.v, opt.v)
    for k, v in net.state_dict().items():
        assert torch.equal(net2.state_dict()[k], v), k
    one_step(net, opt, 3)
    one_step(net2, opt2, 3)
    torch.cuda.synchronize()
    d = (net2.store.param_arena[:used] - net.store.param_arena[:used]).abs()
    # an Adam update is at most lr = 1e-4 per element; a wrong moment / step count / lr would move every element by O(lr)
    assert float(d.mean()) < 1e-7 and float(d.max()) <= 2.5e-4, (float(d.mean()), float(d.max()))


# ------------------------------------------------------------------------------- edge cases of the input contract
@pytest.mark.parametrize("B,lens", [(1, [1.0]), (2, [50.0, 7.0]), (3, [1.0, 20.0, 1.0])])
def test_extreme_batch_and_phrase_lengths_vs_live_oracle(stack, B, lens):
    """Smallest batch (one pair: BatchNorm statistics over a single image), one-token phrases (the reverse LSTM direction
    and the forward one see the same single token) and the loader's maximum phrase length of 50 (dat_loader.py:86)."""
    net, crit, ev, synth = stack
    from oracle import zsg_oracle as zo
    T = int(max(lens))
    batch = synth.make_batch(B, seed=91, T=T)
    batch["qlens"] = torch.tensor(lens)
    net.load_state_dict(synth.make_state_dict(0), strict=True)
    net.train()
    net.zero_grad()
    torch.manual_seed(5)
    out = net(to_dev(batch))
    ls = crit(out, to_dev(batch))
    ls["loss"].mean().backward()
    met = ev(out, to_dev(batch))
    torch.cuda.synchronize()
    sd = synth.make_state_dict(0)
    ols, omet, ograds, oout, _ = zo.train_step(sd, batch, seed=5, do_adam=False)
    assert out["att_out"].shape == (B, 17460, 1)
    for k in ("loss", "cls_ls", "box_ls"):
        assert ls[k].item() == pytest.approx(ols[k].item(), rel=RTOL), k
    assert torch.equal(crit.last_top1.cpu(), ols["top1"]) and torch.equal(crit.last_pos.cpu().bool(), ols["pos"])
    assert torch.equal(met["best_ids"].cpu(), omet["idxs_best"]) and met["Acc"].item() == omet["Acc"].item()
    # the language path is exact arithmetic-wise (no BatchNorm in it): its gradients must agree tightly
    for k in ("lstm.weight_ih_l0_reverse", "lstm.bias_hh_l0"):
        g, r = net.get_parameter(k).grad.cpu().double(), ograds[k].double()
        assert float((g - r).norm() / r.norm().clamp_min(1e-30)) < 5e-2, k


def test_inputs_are_validated_not_silently_accepted(stack):
    net, crit, ev, synth = stack
    batch = synth.make_batch(2, seed=3)
    with pytest.raises(RuntimeError):
        net(batch)                                        # host tensors: there is no CPU path
    bad = to_dev(batch)
    bad["img"] = bad["img"][:, :, :256, :256].contiguous()
    with pytest.raises((AssertionError, RuntimeError, NotImplementedError)):
        net(bad)                                          # only 300x300 is built (resize_img)


def test_pretrained_trunk_weights_load_like_the_reference(tmp_path):
    """mdl.py:411 starts from torchvision's resnet50 weights: a torchvision state_dict (random here: no network) loads into
    backbone.encoder through cfg['zsg_pretrained']; without it get_default_net warns that the trunk is randomly initialised."""
    import torchvision
    import zsg_b200  # noqa: F401
    from zsg_b200 import mdl
    from oracle import synth
    torch.manual_seed(3)
    tv = torchvision.models.resnet50(weights=None)
    path = tmp_path / "resnet50.pth"
    torch.save(tv.state_dict(), path)
    cfg = synth.default_cfg()
    cfg["device"] = "cuda"
    with pytest.warns(UserWarning, match="randomly initialised"):
        mdl.get_default_net(num_anchors=9, cfg=cfg)
    cfg["zsg_pretrained"] = str(path)
    net = mdl.get_default_net(num_anchors=9, cfg=cfg)
    sd = net.state_dict()
    for k, v in tv.state_dict().items():
        assert torch.equal(sd["backbone.encoder." + k].cpu(), v), k


# ------------------------------------------------------------------------------- ablation config: do_norm (mdl.py:118-130)
@pytest.mark.parametrize("split", ["1", "0"])
def test_do_norm_train_step_vs_live_oracle(split, monkeypatch):
    """cfg do_norm: every feature pixel and the language vector are L2-normalised in front of the fusion.  The oracle's
    do_norm is pinned against the unmodified reference in tests/test_ablation_cfg_cpu.py.  Both formulations of the first
    head conv (split / materialised)."""
    monkeypatch.setenv("ZSG_SPLIT_HEAD0", split)
    import zsg_b200  # noqa: F401
    from zsg_b200 import mdl, loss, evaluator
    from oracle import synth, zsg_oracle as zo
    cfg = synth.default_cfg()
    cfg["device"], cfg["do_norm"], cfg["zsg_quiet"] = "cuda", True, True
    ratios, scales = synth.ratios_scales(cfg)
    net = mdl.get_default_net(num_anchors=9, cfg=cfg)
    crit, ev = loss.get_default_loss(ratios, scales, cfg), evaluator.get_default_eval(ratios, scales, cfg)
    B, seed = 3, 41
    batch, out, ls, met = run_step(net, crit, ev, synth, B, seed, True)
    assert net.engine_for(B, 20).do_norm
    ols, omet, ograds, oout, _ = zo.train_step(synth.make_state_dict(0), synth.make_batch(B, seed=seed, var_len=True), seed=seed,
                                               do_adam=False, do_norm=True)
    plain, _, _, _, _ = zo.train_step(synth.make_state_dict(0), synth.make_batch(B, seed=seed, var_len=True), seed=seed, do_adam=False)
    assert abs(plain["loss"].item() - ols["loss"].item()) > 1e-3 * abs(ols["loss"].item())      # the flag changes the function
    for k in ("loss", "cls_ls", "box_ls"):
        assert ls[k].item() == pytest.approx(ols[k].item(), rel=RTOL), k
    assert torch.equal(crit.last_top1.cpu(), ols["top1"]) and torch.equal(crit.last_pos.cpu().bool(), ols["pos"])
    assert rms_rel(out["att_out"].detach().cpu().numpy(), oout["att_out"].detach().numpy()) < 3e-4
    # gradients: against the noise floor of this network (the same oracle functions executed by PyTorch on CUDA in fp32, TF32
    # off), like test_train_step_vs_live_oracle_b4
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sdg = {k: v.cuda() for k, v in synth.make_state_dict(0).items()}
    _, _, cgrads, _, _ = zo.train_step(sdg, batch, seed=seed, do_adam=False, do_norm=True)
    mine_e, cuda_e = [], []
    for k, g in ograds.items():
        if g is None:
            continue
        r = g.double()
        n = r.norm().clamp_min(1e-30)
        mine_e.append(float((net.get_parameter(k).grad.cpu().double() - r).norm() / n))
        cuda_e.append(float((cgrads[k].cpu().double() - r).norm() / n))
    mine_e, cuda_e = np.array(mine_e), np.array(cuda_e)
    print(f"do_norm gradient error vs CPU fp32: zsg_b200 median {np.median(mine_e):.3e} max {mine_e.max():.3e}; torch CUDA fp32 "
          f"median {np.median(cuda_e):.3e} max {cuda_e.max():.3e}")
    assert np.median(mine_e) < 2.0 * np.median(cuda_e) + 1e-3
    assert mine_e.max() < 3.0 * cuda_e.max() + 1e-3
