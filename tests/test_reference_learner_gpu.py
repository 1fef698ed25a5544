"""The drop-in under its real caller: the UNMODIFIED reference `Learner` (code/utils.py:393-438 train_epoch, 353-391
validate; imported from oracle/_ref -- or /root/reference -- through oracle/ref_harness.py) drives zsg_b200's ZSGNet /
ZSGLoss / Evaluator / DataWrap exactly as code/main_dist.py:18-54 wires them: `mdl.to(device)`, a stock
`torch.optim.Adam(betas=(0.9, 0.99))` over `mdl.parameters()`, once plain and once wrapped in
torch.nn.parallel.DistributedDataParallel(find_unused_parameters=True, broadcast_buffers=True) on one rank.
INTEGRATION.md's claims rest on this test."""
import os
import socket
import types
from functools import partial

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_harness
    if ref_harness.find_reference() is None:
        pytest.skip("no reference tree (oracle/_ref is made by oracle/build_ref.py in the build container)")
    return ref_harness.import_reference()


def make_learner(ref, tmp_path, wrap_ddp):
    import zsg_b200  # noqa: F401
    from zsg_b200 import dat_loader, evaluator, loss, mdl
    from oracle import ref_harness
    cfg = ref["cfg"]
    cfg.device = "cuda"
    cfg.mdl_to_use = "retina"
    cfg.bs, cfg.nw, cfg.num_gpus, cfg.do_dist, cfg.local_rank = 4, 0, 1, bool(wrap_ddp), 0
    cfg.tmp_path, cfg.synthetic_len, cfg.resume = str(tmp_path), 12, False
    ratios, scales = ref_harness.ratios_scales(cfg)
    device = torch.device("cuda")
    data = dat_loader.get_data(cfg)                                   # main_dist.py:20
    net = mdl.get_default_net(num_anchors=len(ratios) * len(scales), cfg=cfg)
    net.to(device)                                                    # main_dist.py:35
    if wrap_ddp:
        net = torch.nn.parallel.DistributedDataParallel(net, device_ids=[0], output_device=0, broadcast_buffers=True,
                                                        find_unused_parameters=True)      # main_dist.py:37-40
    loss_fn = loss.get_default_loss(ratios, scales, cfg)
    loss_fn.to(device)
    eval_fn = evaluator.get_default_eval(ratios, scales, cfg)
    opt_fn = partial(torch.optim.Adam, betas=(0.9, 0.99))             # main_dist.py:50
    return ref["utils"].Learner(uid="zsg_dropin", data=data, mdl=net, loss_fn=loss_fn, opt_fn=opt_fn, eval_fn=eval_fn,
                                device=device, cfg=cfg)


@pytest.mark.parametrize("wrap_ddp", [False, True])
def test_reference_learner_trains_and_validates_the_dropin(ref, tmp_path, wrap_ddp):
    import torch.distributed as dist
    if wrap_ddp:
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{_free_port()}", rank=0, world_size=1,
                                device_id=torch.device("cuda", 0))
    try:
        learn = make_learner(ref, tmp_path, wrap_ddp)
        assert learn.loss_keys == ["loss", "cls_ls", "box_ls"] and learn.met_keys == ["Acc", "MaxPos"]
        learn.optimizer = learn.prepare_optimizer()                  # utils.py:667-672: torch.optim.Adam(mdl.parameters(), lr)
        assert type(learn.optimizer) is torch.optim.Adam
        core = learn.mdl.module if wrap_ddp else learn.mdl
        before = core.store.param_arena.clone()
        mb = types.SimpleNamespace(child=types.SimpleNamespace(comment=""))
        trn_loss, trn_met = learn.train_epoch(mb)                     # 12 samples / bs 4 = 3 iterations (utils.py:393-438)
        assert learn.num_it == 3
        assert set(trn_loss) == {"loss", "cls_ls", "box_ls"} and set(trn_met) == {"Acc", "MaxPos"}
        assert all(torch.isfinite(torch.as_tensor(v)).all() for v in trn_loss.values())
        moved = (core.store.param_arena[: core.store.used] - before[: core.store.used]).abs()
        assert float(moved.max()) > 0 and float(moved.max()) <= 3.5e-4      # three Adam steps of lr 1e-4 moved the ARENA in place
        assert torch.equal(core.store.param_arena[core.store.used:], before[core.store.used:])   # unused fc: no gradient, no update
        assert int(core.state_dict()["backbone.encoder.bn1.num_batches_tracked"]) == 3
        val_loss, val_met, preds = learn.validate()                   # utils.py:353-391, eval mode (running statistics)
        assert set(val_loss) == {"loss", "cls_ls", "box_ls"} and 0.0 <= float(val_met["Acc"]) <= 1.0
        assert len(preds) == 3 and set(preds[0]) == {"id", "pred_boxes", "pred_scores"} and len(preds[0]["pred_boxes"]) == 4
        assert int(core.state_dict()["backbone.encoder.bn1.num_batches_tracked"]) == 3      # validate() did not train
        # the checkpoint the Learner writes (utils.py:479-497) loads back into a fresh drop-in AND keeps Adam's per-parameter
        # state attached to the right tensors (registration order = the reference's)
        learn.lr_scheduler = learn.prepare_scheduler(learn.optimizer)
        learn.save_model_dict()
        ck = torch.load(learn.model_file, weights_only=False)
        names = [n for n, _ in core.named_parameters()]
        st = ck["optimizer_state_dict"]["state"]
        for idx, s in st.items():
            assert tuple(s["exp_avg"].shape) == tuple(core.get_parameter(names[idx]).shape), names[idx]
        assert len(st) >= len(core.param_names)                       # every trained parameter has moments
    finally:
        if wrap_ddp:
            dist.destroy_process_group()
