"""cfg do_norm (mdl.py:118-130): the oracle's restatement is pinned against the UNMODIFIED reference run here on CPU
(/root/reference in the build container, oracle/_ref on the GPU box); skipped where neither exists."""
import pytest
import torch

from oracle import ref_harness, synth, zsg_oracle as zo


@pytest.mark.skipif(ref_harness.find_reference() is None, reason="no reference tree (python oracle/build_ref.py)")
def test_oracle_do_norm_equals_the_reference_module():
    ref = ref_harness.import_reference()
    cfg = ref["cfg"]
    cfg.mdl_to_use = "retina"
    cfg.do_norm = True
    try:
        torch.manual_seed(3)
        net = ref["mdl"].get_default_net(num_anchors=9, cfg=cfg)
        net.train()
        sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
        batch = synth.make_batch(2, seed=5)
        torch.manual_seed(11)
        out = net(batch)
        torch.manual_seed(11)
        oout = zo.zsgnet_forward(dict(sd), batch, training=True, do_norm=True)
        for k in ("att_out", "bbx_out"):
            a, b = out[k].detach(), oout[k].detach()
            assert float((a - b).abs().max()) <= 2e-5 * float(b.abs().max()), k
        # and it is a different function from the default one
        torch.manual_seed(11)
        plain = zo.zsgnet_forward(dict(sd), batch, training=True)
        assert float((plain["att_out"] - oout["att_out"]).abs().max()) > 1e-3
    finally:
        cfg.do_norm = False
