#!/usr/bin/env python
"""bench.py — image-query pairs/s of the ZSGNet training hot path on B200 (contract: see README / DESIGN.md).

  python bench.py --gpus N --steps K --warmup W            zsg_b200 (this repo's CUDA path)
  python bench.py --impl reference --gpus N ...            the reference's own CPU implementation (oracle/_ref)

One "step" = the five calls of the reference's hot loop (utils.py:405-414): model forward, loss, backward
(+ NCCL gradient all-reduce when N > 1), Adam, metric — over one batch of synthetic 300x300 images and
length-20 queries.  Headline workload: BASELINE.json configs[1] (bs=64 per GPU, ResNet-50+FPN, fp32 = 3xTF32 on the
tensor cores); the per-GPU batch is kept for N>1 (configs[3] is 512 over 8 GPUs), i.e. weak scaling.  The same line
carries a `bf16` block: the bf16 operand path (configs[2]: bs=128 at N=1; configs[3]: bs=64 per GPU at N>1) measured in the
same run.  `--dtype bf16` makes that arithmetic the top-level measurement instead.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "image-query-pairs/sec (300x300, qlen=20)"
UNIT = "pairs/s"
FLOP_PER_PAIR = {"retina": 97.7e9,      # conv fwd 32.57 GFLOP x 3 (fwd + dgrad + wgrad)
                 "ssd_vgg": 225.0e9}    # SSD-VGG model 75.0 GFLOP forward, x3 for training
LOSS_BYTES_PER_ANCHOR = 40              # att 4 + reg 16 read, d_att 4 + d_reg 16 written


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(tflops=p["bf16_tflops_sustained"], tflops_burst=p["bf16_tflops"], hbm=p["hbm_gbs"], src="measured")
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0, src="fallback")


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload_name(model, batch, dtype):
    arith = "fp32 (3xTF32 on the tensor cores)" if dtype == "fp32" else "bf16 tensor-core convs (fp32 accumulate)"
    if model == "ssd_vgg":
        return (f"vg_split config, bs={batch} per GPU, SSD-VGG backbone (ssd_vgg.py), qlen=20, 300x300 synthetic "
                f"images+queries, {arith} (trunk of BASELINE configs[4])")
    tag = {("fp32", 64): "BASELINE configs[1]", ("bf16", 128): "BASELINE configs[2]", ("bf16", 64): "BASELINE configs[3] per GPU"}
    return (f"refclef config, bs={batch} per GPU, ResNet-50+FPN, qlen=20, 300x300 synthetic images+queries, {arith}"
            + (f" ({tag[(dtype, batch)]})" if (dtype, batch) in tag else ""))


def config_dict(args, world):
    """The workload description; identical in both arms (the reference arm measures the same workload on host cores)."""
    return {"workload": workload_name(args.model, args.batch, args.dtype), "global_batch": args.batch * world,
            "per_gpu_batch": args.batch, "qlen": 20, "parallelism": f"dp{world}",
            "step": "forward + loss + backward (+ all-reduce) + Adam + metric",
            "l2": "no flush needed: one step streams tens of GiB of activations, far beyond the 126 MB L2; 4 resident input "
                  "batches are rotated"}


# ---------------------------------------------------------------------------------------------------------------------
# CPU side: the reference's own implementation (oracle/_ref, copied from /root/reference by oracle/build_ref.py) or, when
# that copy does not exist, the oracle port (the reference's arithmetic restated, without its 1.22 GB torch.eye)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_step_fn(model):
    import torch
    from oracle import ref_harness, synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if ref_harness.find_reference() is not None:
        step = ref_harness.build_reference_step(model)
        return "reference", cores, (lambda bs, i: step(synth.make_batch(bs, seed=1234 + i)))
    from oracle import zsg_oracle as zo
    sd, state = synth.make_state_dict(0, model), {}
    return "port", cores, (lambda bs, i: zo.train_step(sd, synth.make_batch(bs, seed=1234 + i), opt_state=state, seed=i))


def cpu_rate(step, bs, steps, warmup):
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step(bs, i)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return bs / sec, sec


def run_reference(args, rank, world):
    """The reference arm: the reference's CPU implementation of the path on this box's host cores, all threads."""
    if rank != 0:
        return
    kind, cores, step = cpu_step_fn(args.model)
    total = args.steps + args.warmup
    bs = args.ref_batch
    if bs <= 0:
        # each step = the largest batch of the workload's per-GPU batch for which the whole run stays inside the budget;
        # probed with one step of 8 pairs (per-pair cost only falls with the batch size, so the estimate is safe)
        t0 = time.perf_counter()
        step(8, 10 ** 6)
        per_pair = (time.perf_counter() - t0) / 8
        bs = next((b for b in (args.batch, 32, 16, 8) if b <= args.batch and total * b * per_pair <= args.ref_budget_s), 4)
    rate, sec = cpu_rate(step, bs, args.steps, args.warmup)
    what = ("the unmodified reference (code/mdl.py ZSGNet + loss.py ZSGLoss + torch.optim.Adam + evaluator.py Evaluator, "
            "utils.py:405-414)" if kind == "reference" else "the oracle port of the reference's arithmetic (oracle/zsg_oracle.py)")
    sample = (f"each step = {bs} pairs" + ("" if bs == args.batch else f" (a bounded sample of the bs={args.batch} step)")
              + f": forward, loss, backward, Adam, metric of {what}, PyTorch CPU on all {cores} host threads, {sec:.2f} s/step")
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(args, max(world, 1)),
            "sample_batch": bs,
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------------------------------------------------
def net_cfg(model, dtype, local_rank):
    return {"do_norm": False, "use_same_atb": True, "mdl_to_use": model, "resize_img": [300, 300], "use_multi": True,
            "use_focal": True, "use_softmax": False, "alpha": 0.25, "gamma": 2, "emb_dim": 300,
            "matching_threshold": 0.6, "use_bidirectional": True, "lstm_dim": 128, "lamb_reg": 1,
            "acc_iou_threshold": 0.5, "use_lang": True, "use_img": True, "device": f"cuda:{local_rank}",
            "zsg_dtype": dtype, "zsg_direct_grads": True, "zsg_quiet": True}


def measure(model, dtype, B, steps, warmup, rank, world, local_rank, want_e2e=True, sampler=None):
    """Build the net at (dtype, B) and time it: device-resident `value`, per-kernel roofline pass, the loss pass alone,
    and the end-to-end module-API loop with host batches.  Returns a dict of results (no printing)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from zsg_b200 import _lib, dat_loader, ddp, evaluator, loss, mdl, ops, optim, spec
    from zsg_b200.trainer import FusedStep

    cfg = net_cfg(model, dtype, local_rank)
    ratios, scales = [1 / 2, 1, 2], 4 * np.array([1, 2 ** (1 / 3), 2 ** (2 / 3)])      # main_dist.py:24-31
    torch.manual_seed(0)
    net = mdl.get_default_net(num_anchors=9, cfg=cfg)                 # random init (no checkpoints offline)
    net.train()
    reducer = ddp.GradReducer(net.store)
    reducer.broadcast_state(net)
    fused = FusedStep(net, ratios, scales, cfg, lr=1e-4, reducer=reducer)
    dev = net.store.device

    # ---- resident inputs: each rank owns its shard of the global batch (no data-path collective) ----
    nres = 4
    host = [dat_loader.synthetic_batch(B, seed=1000 * rank + i, pin=True) for i in range(nres)]
    resident = []
    for hb in host:
        d = {k: v.to(dev) for k, v in hb.items()}
        d["qlens_cpu"] = hb["qlens"]
        resident.append(d)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------------------------------------------------------- value: inputs resident in HBM
    for i in range(warmup):
        torch.manual_seed(i)
        fused.step(resident[i % nres])
    barrier()
    if sampler is not None:
        sampler.start()
    launches0 = _lib.LAUNCH_COUNT[0]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(steps):
        fused.step(resident[i % nres])
    ev1.record()
    barrier()
    ms_step = max_over_ranks(ev0.elapsed_time(ev1)) / steps
    value = B * world / (ms_step / 1e3)
    eng = net.engine_for(B, 20)
    # kernels launched per step: the step replays CUDA graphs, so the count is that of the captured launch program (the
    # ctypes counter only moves in eager / capture passes): one eager step after the timed region counts it
    graphs0 = eng.use_graphs
    eng.use_graphs = False
    c0 = _lib.LAUNCH_COUNT[0]
    fused.step(resident[0])
    per_step_launches = _lib.LAUNCH_COUNT[0] - c0
    eng.use_graphs = graphs0
    torch.cuda.synchronize()
    launches = per_step_launches * steps
    del launches0

    # per-kernel durations for the roofline: CUDA events around every implicit-GEMM launch, on the launching stream, in
    # `psteps` further steps of the same loop, launched eagerly.  In the product step weight gradients run on a side
    # stream next to the data gradients (their intervals overlap and cannot be attributed), so for this pass they are put
    # back in line.
    overlap0, eng.overlap_wgrad = eng.overlap_wgrad, False
    prof = ops.LaunchProfiler()
    ops.PROFILER = prof
    psteps = min(steps, 5)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for i in range(psteps):
        fused.step(resident[i % nres])
    p1.record()
    barrier()
    ops.PROFILER = None
    eng.overlap_wgrad = overlap0
    ms_step_inline = p0.elapsed_time(p1) / psteps
    clocks = sampler.stop() if sampler is not None else None

    # ---------------------------------------------------------------- roofline of the dominant kernel
    pk = peaks()
    ksum = prof.summary()
    dom = max(ksum, key=lambda k: ksum[k]["ms"])
    d = ksum[dom]
    achieved = d["flops"] / (d["ms"] / 1e3) / 1e12
    traffic = None                # DRAM bytes per launch of the dominant kernel from the committed ncu capture of this round
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(f"{dom}@{dtype}@bs{B}", {}).get("dram_bytes_per_launch")
    note = ("fp32 path = 3xTF32: 3 kind::tf32 MMAs per algorithmic product, 128-column tiles (the fp32 running sums of the "
            "chunked promotion fill the drain warps' registers at 128 columns).  Measured (tools/micro/mma_bench.cu): a "
            "128x128x8 kind::tf32 MMA issues every 75.6 cycles = 3468 FLOP/cycle/SM (a 128x256x8 one every 128.0 = 4096), so "
            "the ceiling of this arithmetic and tile is 148 x 3468 x clock / 3 = 316 TFLOP/s at 1.845 GHz = 0.23 of the bf16 "
            "peak used here (373 = 0.27 with 256-column tiles)" if dtype == "fp32" else
            "bf16 path: one kind::f16 MMA per product, fp32 accumulation in TMEM; 256-column tiles (8192 FLOP/cycle/SM, one "
            "issuing thread, whole-K accumulation) where the layer has a multiple of 256 output channels, 128-column tiles "
            "(6936 FLOP/cycle/SM) elsewhere")
    roof = {"kernel": dom, "bound": "tensor", "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s",
            "frac": achieved / pk["tflops"], "traffic": traffic, "peak_source": f"{pk['src']} bf16 sustained (MEASURED_PEAKS.json)",
            "launches_per_step": d["launches"] / psteps, "avg_launch_ms": d["ms"] / d["launches"],
            "share_of_step": d["ms"] / psteps / ms_step_inline, "timed_in": f"{psteps} eager steps after the timed region, weight "
            f"gradients in line ({ms_step_inline:.2f} ms/step; the product step replays CUDA graphs and overlaps them with the "
            "data gradients)", "note": note}
    kernels = {k: {"tflops": v["flops"] / (v["ms"] / 1e3) / 1e12, "ms_per_step": v["ms"] / psteps,
                   "launches_per_step": v["launches"] / psteps} for k, v in ksum.items()}
    gemm_flops = sum(v["flops"] for v in ksum.values()) / psteps
    gemm_ms = sum(v["ms"] for v in ksum.values()) / psteps
    step_tflops = value / world * FLOP_PER_PAIR[model] / 1e12

    # HBM-bound side: the fused match + loss + gradient pass, timed alone with CUDA events
    bufs = fused._bufs(B)
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    A = spec.NUM_ANCHORS
    flat, dflat = eng.out.view(-1), eng.d_out.view(-1)
    reps = 20

    def match_pass():
        ops.match(resident[0]["annot"], fused.anchs, B, A, 0.6, True, bufs["top1"], bufs["pos"], bufs["ws"])

    def loss_pass():
        ops.loss_grad(flat[4:], 5, eng.out, 5, resident[0]["annot"], fused.anchs, bufs["pos"], B, A, 0.25, 2.0, 1.0,
                      bufs["losses"], dflat[4:], 5, eng.d_out, 5, bufs["ws"])
    match_pass()
    loss_pass()
    torch.cuda.synchronize()
    mg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(mg, capture_error_mode="thread_local"):
        for _ in range(reps):
            match_pass()
    mg.replay()
    torch.cuda.synchronize()
    a0.record()
    mg.replay()
    a1.record()
    torch.cuda.synchronize()
    match_ms = a0.elapsed_time(a1) / reps
    # `reps` back-to-back calls captured in one CUDA graph and replayed: launched one by one from Python the two ~10 us
    # kernels of a call are issued more slowly (~50 us of ctypes + launch overhead per call) than they execute, and the
    # events would time the host
    lg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(lg, capture_error_mode="thread_local"):
        for _ in range(reps):
            loss_pass()
    lg.replay()
    torch.cuda.synchronize()
    a0.record()
    lg.replay()
    a1.record()
    torch.cuda.synchronize()
    loss_ms = a0.elapsed_time(a1) / reps
    loss_gbs = B * A * LOSS_BYTES_PER_ANCHOR / (loss_ms / 1e3) / 1e9
    roof_hbm = {"kernel": "zsg_loss_grad (loss_grad_packed_kernel + loss_finalize_kernel)", "bound": "hbm", "achieved": loss_gbs,
                "peak": pk["hbm"], "unit": "GB/s", "frac": loss_gbs / pk["hbm"], "traffic": None, "ms": loss_ms,
                "match_ms": match_ms,
                "note": f"{B * A * LOSS_BYTES_PER_ANCHOR / 1e6:.0f} MB algorithmic per call (40 B per anchor: scores + boxes read, both "
                        f"gradients written), two launches per call; {reps} back-to-back calls on the same buffers replayed as one CUDA "
                        "graph (L2-resident at this size, as in the step, where the head has just written the scores).  The anchor "
                        "match (zsg_match, match_ms: fp64 IoU of every (sample, anchor) pair, ~1 B per anchor of traffic) is ALU "
                        "work with no HBM roofline; in the step it runs on a side stream under the forward pass"}

    # ---------------------------------------------------------------- e2e: public module API, host buffers
    e2e = None
    if want_e2e:
        crit = loss.get_default_loss(ratios, scales, cfg)
        evalr = evaluator.get_default_eval(ratios, scales, cfg)
        opt = optim.FusedAdam(net.parameters(), lr=1e-4, betas=(0.9, 0.99), net=net, reducer=reducer)
        net._on_bucket = reducer.on_bucket if world > 1 else None

        def e2e_step(batch):
            opt.zero_grad()
            out = net(batch)
            ls = crit(out, batch)
            ls["loss"].mean().backward()
            opt.step()
            met = evalr(out, batch)
            return ls["loss"], met["Acc"]

        def host_batches(n):                                                        # pinned host batches, like a DataLoader
            for i in range(n):
                yield host[i % nres]

        for batch in dat_loader.DevicePrefetcher(host_batches(5), dev, lstm_state=True):             # warm-up (allocator caches)
            e2e_step(batch)
        barrier()
        t0 = time.perf_counter()
        # every step's batch is copied host -> device inside the timed region (utils.py:405-406), one step ahead on a
        # copy stream; every step ends with the device -> host read of its loss and metric (utils.py:426 formats the loss)
        marks = [t0]
        for batch in dat_loader.DevicePrefetcher(host_batches(steps), dev, lstm_state=True):
            lt, at = e2e_step(batch)
            lv, av = float(lt.item()), float(at.item())
            marks.append(time.perf_counter())
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        step_ms = [round((b - a) * 1e3, 2) for a, b in zip(marks, marks[1:])]
        h2d = sum(v.numel() * v.element_size() for v in host[0].values())
        e2e = {"value": B * world * steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12,
               "ms_per_step": dt / steps * 1e3, "api": "mdl.get_default_net(cfg)(batch) -> loss.get_default_loss -> "
               ".backward() -> optim.FusedAdam.step -> evaluator.get_default_eval (cfg zsg_direct_grads: param.grad are views of the "
               "gradient arena); dat_loader.DevicePrefetcher copies every step's pinned host batch to the device one step ahead "
               "on a copy stream", "last_loss": lv, "last_acc": av, "step_ms": step_ms}
        net._on_bucket = None

    nsteps_total = warmup + steps + 1 + psteps + ((steps + 5) if want_e2e else 0)
    res = {"value": value, "ms_per_step": ms_step, "gpu_launches": launches, "gpu_launches_per_step": per_step_launches,
           "cuda_graphs": bool(eng.use_graphs), "graph_launches_per_step": (1 + len(eng.segments) if world > 1 else 2) if eng.use_graphs else 0,
           "roofline": roof, "roofline_hbm": roof_hbm, "kernels": kernels,
           "gemm_tflops_all_kernels": gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms else None,
           "gemm_ms_per_step_inline": gemm_ms, "ms_per_step_eager_inline": ms_step_inline,
           "step_tensor_tflops_per_gpu": step_tflops, "step_tensor_frac_of_peak": step_tflops / pk["tflops"],
           "engine_gib": eng.nbytes / 2 ** 30, "e2e": e2e, "clocks": clocks,
           "allreduce": {"bytes_per_step": reducer.bytes_reduced / nsteps_total, "buckets_per_step": reducer.calls / nsteps_total}}
    # release this arm's buffers before the next one is built
    del fused, net, eng, resident
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="zsg", choices=["zsg", "reference"])
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default 64)")
    ap.add_argument("--model", default="retina", choices=["retina", "ssd_vgg"],
                    help="image trunk (cfg mdl_to_use); the headline workload is retina")
    ap.add_argument("--dtype", default="fp32", choices=["fp32", "bf16"],
                    help="arithmetic of the dense contractions of the top-level measurement (headline: fp32 = configs[1])")
    ap.add_argument("--no-bf16-block", action="store_true", help="skip the secondary bf16 measurement (configs[2]/[3])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--ref-batch", type=int, default=0, help="reference arm: pairs per step (0 = choose within --ref-budget-s)")
    ap.add_argument("--ref-budget-s", type=float, default=240.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.batch is None:
        args.batch = 64
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"              # the version banner goes to stdout, where the ONE JSON line belongs
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import zsg_b200  # noqa: F401

    sampler = ClockSampler(local_rank) if rank == 0 else None
    main_res = measure(args.model, args.dtype, args.batch, args.steps, args.warmup, rank, world, local_rank,
                       want_e2e=not args.no_e2e, sampler=sampler)

    # ---------------------------------------------------------------- bf16 block: configs[2] (N=1) / configs[3] (N>1)
    bf16 = None
    if args.dtype == "fp32" and args.model == "retina" and not args.no_bf16_block:
        b2 = 128 if world == 1 else 64
        r = measure(args.model, "bf16", b2, min(args.steps, 10), 3, rank, world, local_rank, want_e2e=not args.no_e2e)
        bf16 = {"workload": workload_name(args.model, b2, "bf16"), "per_gpu_batch": b2, "dtype": "bf16",
                "steps": min(args.steps, 10), "warmup": 3,
                **{k: r[k] for k in ("value", "ms_per_step", "e2e", "roofline", "roofline_hbm", "kernels", "gemm_tflops_all_kernels",
                                     "step_tensor_tflops_per_gpu", "step_tensor_frac_of_peak", "engine_gib", "gpu_launches_per_step")}}

    # ---------------------------------------------------------------- CPU baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        kind, cores, step = cpu_step_fn(args.model)
        rate, sec = cpu_rate(step, 16, 2, 1)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"2 timed steps (+1 warm-up) of bs=16 from the same synthetic distribution, {sec:.1f} s/step: forward, "
                         "loss, backward, Adam, metric of " + ("the unmodified reference (oracle/_ref)" if kind == "reference"
                                                              else "the oracle port") + ", PyTorch CPU on all host threads"}

    if rank == 0:
        line = {"metric": METRIC, "value": main_res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32" if args.dtype == "fp32" else "bf16", "data": "synthetic",
                "config": config_dict(args, world), "clocks": main_res["clocks"], "e2e": main_res["e2e"],
                "gpu_launches": main_res["gpu_launches"], "roofline": main_res["roofline"],
                "roofline_hbm": main_res["roofline_hbm"], "cpu_baseline": cpu, "bf16": bf16,
                **{k: main_res[k] for k in ("kernels", "gemm_tflops_all_kernels", "gemm_ms_per_step_inline", "ms_per_step_eager_inline",
                                            "step_tensor_tflops_per_gpu", "step_tensor_frac_of_peak", "engine_gib", "cuda_graphs",
                                            "gpu_launches_per_step", "graph_launches_per_step", "allreduce")}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
