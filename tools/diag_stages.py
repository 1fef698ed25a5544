"""Diagnostics (GPU box): where does numerical error enter?  (1) single-conv accuracy and bias of the 3xTF32
tensor-core kernel vs an fp64 reference, next to torch's fp32 conv; (2) stage-by-stage error vs the CPU oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import zsg_b200
from zsg_b200 import ops, geometry, mdl, loss, evaluator, spec
from oracle import synth, zsg_oracle as zo

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def stats(name, a, ref64):
    a, r = a.double().flatten(), ref64.flatten()
    d = a - r
    rms = float(d.pow(2).mean().sqrt() / r.pow(2).mean().sqrt())
    bias = float((d * r.sign()).mean() / r.abs().mean())
    print(f"{name:34s} rel-rms {rms:.3e}  signed-bias(toward +|x|) {bias:+.3e}  max {float(d.abs().max()/r.abs().max()):.3e}")


def conv_accuracy():
    g = torch.Generator().manual_seed(0)
    for (B, cin, H, cout, k, pos) in ((4, 256, 38, 256, 3, False), (4, 256, 38, 256, 3, True), (4, 64, 75, 64, 1, True), (2, 2048, 10, 512, 1, True)):
        x = torch.randn(B, cin, H, H, generator=g).cuda()
        if pos:
            x = x.abs()                                  # post-ReLU-like, same-sign products accumulate
        w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).cuda()
        ref64 = F.conv2d(x.double(), w.double(), padding=k // 2)
        t32 = F.conv2d(x, w, padding=k // 2)
        rows = geometry.conv_rows(B, H, H, cin, H, H, cout, 1, k // 2).cuda()
        y = torch.empty(B, H, H, cout, device="cuda")
        ops.ConvOp(x.permute(0, 2, 3, 1).contiguous(), w.permute(0, 2, 3, 1).contiguous(), y, rows, B * H * H, cin, cout, k, k)()
        torch.cuda.synchronize()
        tag = f"cin{cin} k{k} {'pos' if pos else 'rnd'}"
        stats("torch fp32 conv  " + tag, t32.permute(0, 2, 3, 1), ref64.permute(0, 2, 3, 1))
        stats("zsg 3xTF32 conv  " + tag, y, ref64.permute(0, 2, 3, 1))


def stages(B=2, seed=21):
    cfg = synth.default_cfg(); cfg["device"] = "cuda"
    net = mdl.get_default_net(9, cfg)
    net.load_state_dict(synth.make_state_dict(0), strict=True)
    net.train()
    cb = synth.make_batch(B, seed=seed)
    torch.manual_seed(seed)
    out = net({k: v.cuda() for k, v in cb.items()})
    torch.cuda.synchronize()
    eng = net.engine_for(B, 20)
    sd = synth.make_state_dict(0)
    torch.manual_seed(seed)
    o = zo.zsgnet_forward(sd, cb, training=True, return_inter=True)
    it = o["_inter"]
    nh = lambda t: t.permute(0, 2, 3, 1).reshape(-1, t.shape[1])
    d = eng.dbg
    stats("lang (LSTM)", d["lang"].cpu(), it["lang"].double())
    stats("c3", d["c3"].cpu(), nh(it["c3"]).double())
    stats("c4", d["c4"].cpu(), nh(it["c4"]).double())
    stats("c5", d["c5"].cpu(), nh(it["c5"]).double())
    lo = d["lvl_off"]
    for i in range(6):
        stats(f"P{i+3}", d["feat"][lo[i]:lo[i + 1]].cpu(), nh(it["feats"][i]).double())
    stats("att_out", out["att_out"].detach().cpu(), o["att_out"].double())
    stats("bbx_out", out["bbx_out"].detach().cpu(), o["bbx_out"].double())


def quick_time(B=64, steps=5):
    from zsg_b200.trainer import FusedStep
    cfg = synth.default_cfg(); cfg["device"] = "cuda"
    net = mdl.get_default_net(9, cfg)
    net.load_state_dict(synth.make_state_dict(0), strict=True)
    net.train()
    ratios, scales = synth.ratios_scales(cfg)
    fs = FusedStep(net, ratios, scales, cfg)
    batch = {k: v.cuda() for k, v in synth.make_batch(B, seed=1).items()}
    batch["qlens_cpu"] = batch["qlens"].cpu()
    for _ in range(2):
        r = fs.step(batch)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    t0 = time.time()
    for i in range(steps):
        r = fs.step(batch)
        ev[i + 1].record()
    host = time.time() - t0
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
    print(f"B={B}: step ms {['%.1f' % m for m in ms]}  host-issue {host/steps*1e3:.1f} ms/step  pairs/s {B/ (sum(ms)/steps/1e3):.1f}  loss {r['loss'].item():.4f}")
    print(f"engine buffers {net.engine_for(B, 20).nbytes/2**30:.2f} GiB, peak alloc {torch.cuda.max_memory_allocated()/2**30:.2f} GiB")
    # coarse breakdown: forward only / loss / backward
    eng = net.engine_for(B, 20)
    def timeit(fn, n=3):
        torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n): fn()
        b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
    print("forward  ms", timeit(lambda: eng.forward(True)))
    print("backward ms", timeit(lambda: eng.backward(None)))


if __name__ == "__main__":
    conv_accuracy()
    stages()
    quick_time()
