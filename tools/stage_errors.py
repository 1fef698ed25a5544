"""Prints the stage-wise errors of tests/test_stages_gpu.py without asserting (tolerance calibration / debugging)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import test_stages_gpu as T

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
for dtype in sys.argv[1:] or ["fp32", "bf16"]:
    st = T.Stage(dtype)
    zo, eng = st.zo, st.eng
    print(f"=== {dtype}")
    for label in ["layer1.0", "layer1.2", "layer2.0", "layer3.1", "layer4.0"]:
        blk = next(b for b in eng.dbg["blocks"] if b["label"] == label)
        p = f"backbone.encoder.{label}."
        keys = [k for k in st.sd if k.startswith(p) and "running" not in k and "num_batches" not in k]
        sd = st.sdg(keys)
        x = st.acts[label].clone().requires_grad_(True)
        stride = 2 if (label.endswith(".0") and not label.startswith("layer1")) else 1
        with zo.conv_mode(st.dtype):
            y = zo.bottleneck(sd, x, p, stride, zo.BNState(sd, True))
            G = torch.randn(y.shape, generator=torch.Generator().manual_seed(2))
            (y * G).sum().backward()
        st.put(blk["inp"], x)
        st.fwd(label)
        e_f = T.err(blk["out"], T.nhwc(y))

        def fill():
            st.put(blk["g_out"], G)
            if blk["acc_in"]:
                blk["g_in"].zero_()
        st.bwd(label, fill)
        e_x = T.err(blk["g_in"], T.nhwc(x.grad))
        worst = {k[len(p):]: T.err(st.grad(k), sd[k].grad) for k in keys}
        print(f"{label}: fwd {e_f:.2e} dx {e_x:.2e} params " + " ".join(f"{k}={v:.1e}" for k, v in worst.items()))
    # head
    keys = [k for k in st.sd if k.startswith("att_reg_box.")]
    sd = st.sdg(keys)
    feats = [f.clone().requires_grad_(True) for f in st.inter["feats"]]
    lang = st.inter["lang"].clone().requires_grad_(True)
    with zo.conv_mode(st.dtype):
        att, bbx = zo.fuse_and_head(sd, feats, lang)
        packed = torch.cat([bbx, att], dim=2)
        G = torch.randn(packed.shape, generator=torch.Generator().manual_seed(4))
        (packed * G).sum().backward()
    for i, f in enumerate(feats):
        st.put(eng.dbg["fl"][i], f)
    eng.lang.copy_(lang.detach().cuda())
    st.fwd("head")
    e_f = T.err(eng.out, packed)
    st.bwd("head", lambda: eng.d_out.copy_(G.cuda()))
    print(f"head: fwd {e_f:.2e} dfeat " + " ".join(f"{T.err(eng.dbg['dfl'][i], T.nhwc(f.grad)):.1e}" for i, f in enumerate(feats))
          + f" dlang {T.err(eng.dbg['dlang'], lang.grad):.1e} params "
          + " ".join(f"{k[12:]}={T.err(st.grad(k), sd[k].grad):.1e}" for k in keys))
