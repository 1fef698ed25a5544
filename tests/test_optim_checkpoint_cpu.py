"""Host logic of FusedAdam's checkpoint format (torch.optim.Adam's state_dict layout, utils.py:479-497) and of
checkpoint.py, on a CPU stand-in for the net (parameters as views of a CPU ParamStore; no kernels are launched)."""
import pytest
import torch
import torch.nn as nn


class _HostNet(nn.Module):
    def __init__(self, model):
        super().__init__()
        from zsg_b200 import engine, mdl, spec
        self.store = engine.ParamStore(torch.device("cpu"), model)
        self.param_names = [n for n, _, _ in spec.trainable_specs(model)]
        for n in self.param_names + [u[0] for u in spec.unused_specs(model)]:
            mdl._attach(self, n, nn.Parameter(self.store.view(n)), True)


@pytest.mark.parametrize("model", ["retina", "ssd_vgg"])
def test_fused_adam_state_dict_has_torch_adam_layout(model):
    from zsg_b200 import optim
    net = _HostNet(model)
    opt = optim.FusedAdam(net.parameters(), lr=3e-4, net=net)
    assert opt.state_dict()["state"] == {}                         # no step yet: like a fresh torch Adam
    g = torch.Generator().manual_seed(1)
    opt.m.copy_(torch.randn(opt.m.shape, generator=g))
    opt.v.copy_(torch.rand(opt.v.shape, generator=g))
    opt.t = 7
    sd = opt.state_dict()
    names = [n for n, _ in net.named_parameters()]
    used = [i for i, n in enumerate(names) if net.store.offsets[n] < net.store.used]
    assert sorted(sd["state"]) == used and len(used) == len(net.param_names)
    assert sd["param_groups"][0]["params"] == list(range(len(names))) and sd["param_groups"][0]["lr"] == 3e-4
    i = names.index("att_reg_box.0.0.weight")
    st = sd["state"][i]
    assert st["exp_avg"].shape == (256, 514, 3, 3) and st["exp_avg"].is_contiguous() and int(st["step"]) == 7
    assert torch.equal(st["exp_avg"], net.store.view("att_reg_box.0.0.weight", opt.m))
    # a stock torch Adam over same-shaped parameters accepts it (the reference's optimiser, main_dist.py:50)
    ref = [nn.Parameter(p.detach().clone().contiguous()) for p in net.parameters()]
    adam = torch.optim.Adam(ref, lr=1e-4, betas=(0.9, 0.99))
    adam.load_state_dict(sd)
    assert torch.equal(adam.state[ref[i]]["exp_avg_sq"], st["exp_avg_sq"]) and adam.param_groups[0]["lr"] == 3e-4
    # and back: the state a torch Adam writes resumes here
    opt2 = optim.FusedAdam(net.parameters(), lr=1.0, net=net)
    opt2.load_state_dict(adam.state_dict())
    assert opt2.t == 7 and opt2.param_groups[0]["lr"] == 3e-4
    for n in net.param_names:                                       # (the arenas also hold alignment padding between tensors)
        assert torch.equal(net.store.view(n, opt2.m), net.store.view(n, opt.m)), n
        assert torch.equal(net.store.view(n, opt2.v), net.store.view(n, opt.v)), n
    bad = adam.state_dict()
    bad["param_groups"][0]["weight_decay"] = 0.1
    with pytest.raises(ValueError):
        opt2.load_state_dict(bad)


def test_checkpoint_file_round_trip_on_host(tmp_path):
    from zsg_b200 import checkpoint, optim
    net = _HostNet("retina")
    with torch.no_grad():
        net.store.param_arena.copy_(torch.randn(net.store.param_arena.shape, generator=torch.Generator().manual_seed(2)))
    opt = optim.FusedAdam(net.parameters(), lr=1e-4, net=net)
    opt.t = 3
    opt.m.fill_(0.5)
    path = tmp_path / "ck.pth"
    checkpoint.save_model_dict(path, net, opt, num_it=11, num_epoch=2, best_met=0.25, cfg={"bs": 64}, ddp_prefix=True)
    net2 = _HostNet("retina")
    opt2 = optim.FusedAdam(net2.parameters(), lr=1e-2, net=net2)
    info = checkpoint.load_model_dict(path, net2, opt2)
    assert info == {"num_it": 11, "num_epoch": 2, "best_met": 0.25}
    a, b = net.state_dict(), net2.state_dict()
    assert set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in a)      # (arena padding between tensors is not state)
    assert opt2.t == 3 and opt2.param_groups[0]["lr"] == 1e-4
    assert all(torch.equal(net.store.view(n, opt2.m), net.store.view(n, opt.m)) for n in net.param_names)
    ck = torch.load(path, weights_only=False)
    assert all(k.startswith("module.") for k in ck["model_state_dict"]) and ck["cfgtxt"] == '{"bs": 64}'
