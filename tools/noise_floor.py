"""How far apart are legitimate fp32 implementations of this network?  Ground truth = the oracle in float64.
Compared against it: the CPU fp32 oracle (= the reference's arithmetic), torch CUDA fp32 (TF32 off) and zsg_b200."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zsg_b200
from zsg_b200 import mdl
from oracle import synth, zsg_oracle as zo

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def rel(a, r):
    a, r = a.double().flatten().cpu(), r.double().flatten().cpu()
    return float((a - r).pow(2).mean().sqrt() / r.pow(2).mean().sqrt())


def trunk(sd, img, dev, dtype):
    sd = {k: (v.to(dev).to(dtype) if v.dtype.is_floating_point else v.to(dev)) for k, v in sd.items()}
    bn = zo.BNState(sd, True)
    c3, c4, c5 = zo.resnet50_c3c4c5(sd, img.to(dev).to(dtype), bn)
    return [c3, c4, c5] + zo.fpn(sd, c3, c4, c5)


def main(B, seed=21):
    cb = synth.make_batch(B, seed=seed)
    sd0 = synth.make_state_dict(0)
    with torch.no_grad():
        ref = trunk(sd0, cb["img"], "cpu", torch.float64)
        cpu32 = trunk(sd0, cb["img"], "cpu", torch.float32)
        gpu32 = trunk(sd0, cb["img"], "cuda", torch.float32)
    cfg = synth.default_cfg(); cfg["device"] = "cuda"
    net = mdl.get_default_net(9, cfg)
    net.load_state_dict(synth.make_state_dict(0), strict=True)
    net.train()
    torch.manual_seed(seed)
    net({k: v.cuda() for k, v in cb.items()})
    torch.cuda.synchronize()
    d = net.engine_for(B, 20).dbg
    lo = d["lvl_off"]
    mine = [d["c3"], d["c4"], d["c5"]] + [d["feat"][lo[i]:lo[i + 1]] for i in range(6)]
    nh = lambda t: t.permute(0, 2, 3, 1).reshape(-1, t.shape[1])
    print(f"B={B}: relative rms error vs float64 ground truth")
    print(f"{'stage':6s} {'cpu fp32 (reference)':>22s} {'torch cuda fp32':>18s} {'zsg_b200':>12s}")
    for i, name in enumerate(["c3", "c4", "c5", "P3", "P4", "P5", "P6", "P7", "P8"]):
        r = nh(ref[i])
        print(f"{name:6s} {rel(nh(cpu32[i]), r):22.3e} {rel(nh(gpu32[i]), r):18.3e} {rel(mine[i], r):12.3e}")


if __name__ == "__main__":
    main(2)
    main(8)
