"""One small training step of the hot path on cuda:0, checked against the CPU oracle
(used by __graft_entry__.smoke)."""
import torch


def run(B=2, seed=9):
    import zsg_b200  # noqa: F401
    from zsg_b200 import mdl, loss, evaluator
    from oracle import synth, zsg_oracle as zo
    assert torch.cuda.is_available(), "smoke() needs cuda:0"
    torch.cuda.set_device(0)
    cfg = synth.default_cfg()
    cfg["device"] = "cuda"
    ratios, scales = synth.ratios_scales(cfg)
    net = mdl.get_default_net(num_anchors=9, cfg=cfg)
    crit = loss.get_default_loss(ratios, scales, cfg)
    ev = evaluator.get_default_eval(ratios, scales, cfg)
    net.load_state_dict(synth.make_state_dict(0), strict=True)
    net.train()
    cpu_batch = synth.make_batch(B, seed=seed, var_len=True)
    batch = {k: v.cuda() for k, v in cpu_batch.items()}
    torch.manual_seed(seed)
    out = net(batch)
    ls = crit(out, batch)
    ls["loss"].mean().backward()
    met = ev(out, batch)
    torch.cuda.synchronize()
    sd = synth.make_state_dict(0)
    ols, omet, ograds, _, _ = zo.train_step(sd, cpu_batch, seed=seed, do_adam=False)
    for k in ("loss", "cls_ls", "box_ls"):
        a, b = ls[k].item(), ols[k].item()
        assert abs(a - b) <= 1e-4 * abs(b), (k, a, b)
    assert torch.equal(crit.last_top1.cpu(), ols["top1"]) and torch.equal(crit.last_pos.cpu().bool(), ols["pos"])
    assert met["Acc"].item() == omet["Acc"].item()
    g = net.get_parameter("backbone.encoder.conv1.weight").grad.cpu().double()
    r = ograds["backbone.encoder.conv1.weight"].double()
    assert float((g - r).norm() / r.norm()) < 0.25      # end-to-end fp32 trunk gradients are chaotic here
    g = net.get_parameter("att_reg_box.5.bias").grad.cpu().double()
    r = ograds["att_reg_box.5.bias"].double()
    assert float((g - r).norm() / r.norm()) < 1e-3      # the last head layer's gradient is not
    print(f"smoke: loss {ls['loss'].item():.6f} (oracle {ols['loss'].item():.6f}), Acc {met['Acc'].item()}")
    # the bf16 operand path (BASELINE configs 3-5) on the same batch, against the oracle in the same arithmetic; the second
    # and third forward of the engine are the CUDA-graph capture and its replay
    net.set_compute_dtype("bf16")
    net.load_state_dict(synth.make_state_dict(0), strict=True)
    with zo.conv_mode("bf16"):
        bls, _, _, _, _ = zo.train_step(synth.make_state_dict(0), cpu_batch, seed=seed, do_adam=False)
    first = None
    for rep in range(3):
        net.load_state_dict(synth.make_state_dict(0), strict=True)
        net.zero_grad()
        torch.manual_seed(seed)
        ls = crit(net(batch), batch)
        ls["loss"].mean().backward()
        torch.cuda.synchronize()
        a, b = ls["loss"].item(), bls["loss"].item()
        assert abs(a - b) <= 3e-2 * abs(b), ("bf16", rep, a, b)      # end to end the random-weight network is chaotic
        assert rep == 0 or abs(a - first) <= 1e-9 * abs(first), ("bf16 graph replay", rep, a, first)   # capture / replay = eager
        first = a
    assert torch.equal(crit.last_top1.cpu(), ols["top1"])
    print(f"smoke: bf16 loss {a:.6f} (bf16 oracle {b:.6f}); graph replay ok")
