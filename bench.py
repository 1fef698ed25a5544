#!/usr/bin/env python
"""bench.py — image-query pairs/s of the ZSGNet training hot path on B200 (contract: see README / DESIGN.md).

  python bench.py --gpus N --steps K --warmup W            zsg_b200 (this repo's CUDA path)
  python bench.py --impl reference --gpus N ...            the reference's CPU implementation (oracle port)

One "step" = the five calls of the reference's hot loop (utils.py:405-414): model forward, loss, backward
(+ NCCL gradient all-reduce when N > 1), Adam, metric — over one batch of synthetic 300x300 images and
length-20 queries.  Workload at N=1: BASELINE.json configs[1] (bs=64, ResNet-50+FPN, fp32); per-GPU batch is
kept at 64 for N>1 (configs[3] is 512 over 8 GPUs), i.e. weak scaling.
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
FLOP_PER_PAIR_TRAIN = 97.7e9        # conv fwd 32.57 GFLOP x 3 (fwd + dgrad + wgrad)
LOSS_BYTES_PER_ANCHOR = 40          # att 4 + reg 16 read, d_att 4 + d_reg 16 written


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
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
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


def cpu_oracle_rate(bs, steps, warmup, seed=1234, model="retina"):
    """pairs/s of the oracle port (the reference's arithmetic, PyTorch CPU fp32/fp64) on all host threads."""
    import torch
    from oracle import synth, zsg_oracle as zo
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synth.make_state_dict(0, model)
    state = {}
    times = []
    for i in range(warmup + steps):
        batch = synth.make_batch(bs, seed=seed + i)
        t0 = time.perf_counter()
        zo.train_step(sd, batch, opt_state=state, seed=i)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return bs / (sum(times) / len(times)), cores, sum(times) / len(times)


def run_reference(args, rank, world):
    """The reference arm: its CPU implementation of the path on this box's host cores (oracle port)."""
    if rank != 0:
        return
    sample_bs = 8
    set_model(args.model)
    rate, cores, sec = cpu_oracle_rate(sample_bs, args.steps, args.warmup, model=args.model)
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": args.batch * args.gpus, "per_gpu_batch": args.batch,
                       "qlen": 20, "parallelism": "cpu"},
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"each step = {sample_bs} pairs of the bs={args.batch} workload (fwd, loss, bwd, Adam, "
                                       f"metric), PyTorch CPU on all {cores} host threads"},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


WORKLOAD = "refclef config, bs=64 per GPU, ResNet-50+FPN, qlen=20, 300x300 synthetic images+queries, fp32 (BASELINE configs[1])"


def set_model(model):
    """--model ssd_vgg: the trunk of BASELINE configs[4] (SSD-VGG16) in fp32 on this arm's batch; not the headline."""
    global WORKLOAD, FLOP_PER_PAIR_TRAIN
    if model == "ssd_vgg":
        WORKLOAD = ("vg_split config, SSD-VGG backbone (ssd_vgg.py), qlen=20, 300x300 synthetic images+queries, fp32 "
                    "(trunk of BASELINE configs[4]; per-GPU batch as given)")
        FLOP_PER_PAIR_TRAIN = 225.0e9   # SSD-VGG model 75.0 GFLOP forward, x3 for training


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="zsg", choices=["zsg", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch")
    ap.add_argument("--model", default="retina", choices=["retina", "ssd_vgg"],
                    help="image trunk (cfg mdl_to_use); the headline workload is retina")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import zsg_b200  # noqa: F401
    from zsg_b200 import _lib, dat_loader, ddp, evaluator, loss, mdl, ops, optim, spec
    from zsg_b200.trainer import FusedStep
    import numpy as np

    B = args.batch
    set_model(args.model)
    cfg = {"do_norm": False, "use_same_atb": True, "mdl_to_use": args.model, "resize_img": [300, 300], "use_multi": True,
           "use_focal": True, "use_softmax": False, "alpha": 0.25, "gamma": 2, "emb_dim": 300,
           "matching_threshold": 0.6, "use_bidirectional": True, "lstm_dim": 128, "lamb_reg": 1,
           "acc_iou_threshold": 0.5, "use_lang": True, "use_img": True, "device": f"cuda:{local_rank}"}
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
    for i in range(args.warmup):
        torch.manual_seed(i)
        fused.step(resident[i % nres])
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.LAUNCH_COUNT[0]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        fused.step(resident[i % nres])
    ev1.record()
    barrier()
    launches = _lib.LAUNCH_COUNT[0] - launches0
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    ms_step = ms_total / args.steps
    value = B * world / (ms_step / 1e3)

    # per-kernel durations for the roofline: CUDA events around every implicit-GEMM launch, on the launching stream, in
    # `psteps` further steps of the same loop.  In the product step weight gradients run on a side stream next to the
    # data gradients (their intervals overlap and cannot be attributed), so for this pass they are put back in line.
    eng0 = net.engine_for(B, 20)
    overlap0, eng0.overlap_wgrad = eng0.overlap_wgrad, False
    prof = ops.LaunchProfiler()
    ops.PROFILER = prof
    psteps = min(args.steps, 5)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for i in range(psteps):
        fused.step(resident[i % nres])
    p1.record()
    barrier()
    ops.PROFILER = None
    eng0.overlap_wgrad = overlap0
    ms_step_inline = p0.elapsed_time(p1) / psteps
    clocks = sampler.stop() if rank == 0 else None

    # ---------------------------------------------------------------- roofline of the dominant kernel
    pk = peaks()
    ksum = prof.summary()
    dom = max(ksum, key=lambda k: ksum[k]["ms"])
    d = ksum[dom]
    achieved = d["flops"] / (d["ms"] / 1e3) / 1e12
    traffic = None                                    # DRAM bytes per launch of the dominant kernel, from the committed ncu capture
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(dom, {}).get("dram_bytes_per_launch")
    roof = {"kernel": dom, "bound": "tensor", "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s",
            "frac": achieved / pk["tflops"], "traffic": traffic, "peak_source": f"{pk['src']} bf16 sustained (MEASURED_PEAKS.json)",
            "launches_per_step": d["launches"] / psteps, "avg_launch_ms": d["ms"] / d["launches"],
            "share_of_step": d["ms"] / psteps / ms_step_inline, "timed_in": f"{psteps} steps after the timed region, weight "
            f"gradients in line ({ms_step_inline:.2f} ms/step; the product step overlaps them with the data gradients)",
            "note": "fp32 path = 3xTF32: 3 kind::tf32 MMAs per algorithmic product; measured on this B200 a 128x128x8 "
                    "kind::tf32 MMA takes 64 cycles = 4096 FLOP/cycle/SM (tools/micro/mma_bench.cu), so the ceiling of this "
                    "arithmetic is 148 SMs x 4096 x clock / 3 = 373 TFLOP/s at 1.845 GHz, i.e. 0.27 of the bf16 peak used here"}
    kernels = {k: {"tflops": v["flops"] / (v["ms"] / 1e3) / 1e12, "ms_per_step": v["ms"] / psteps,
                   "launches_per_step": v["launches"] / psteps} for k, v in ksum.items()}
    step_tflops = value / world * FLOP_PER_PAIR_TRAIN / 1e12

    # HBM-bound side: the fused match + loss + gradient pass, timed alone with CUDA events
    eng = net.engine_for(B, 20)
    bufs = fused._bufs(B)
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    A = spec.NUM_ANCHORS
    flat, dflat = eng.out.view(-1), eng.d_out.view(-1)
    reps = 20
    torch.cuda.synchronize()
    a0.record()
    for _ in range(reps):
        ops.match_loss(flat[4:], 5, eng.out, 5, resident[0]["annot"], fused.anchs, B, A, 0.6, 0.25, 2.0, 1.0, True,
                       bufs["losses"], dflat[4:], 5, eng.d_out, 5, bufs["top1"], bufs["pos"], bufs["ws"])
    a1.record()
    torch.cuda.synchronize()
    loss_ms = a0.elapsed_time(a1) / reps
    loss_gbs = B * A * LOSS_BYTES_PER_ANCHOR / (loss_ms / 1e3) / 1e9
    roof_hbm = {"kernel": "zsg_match_loss (4 launches)", "bound": "hbm", "achieved": loss_gbs, "peak": pk["hbm"],
                "unit": "GB/s", "frac": loss_gbs / pk["hbm"], "traffic": None, "ms": loss_ms,
                "note": "1.1 M anchors x 40 B = 45 MB per launch: latency-bound at this size, L2-resident"}

    # ---------------------------------------------------------------- e2e: public module API, host buffers
    e2e = None
    if not args.no_e2e:
        crit = loss.get_default_loss(ratios, scales, cfg)
        evalr = evaluator.get_default_eval(ratios, scales, cfg)
        opt = optim.FusedAdam(net.parameters(), lr=1e-4, betas=(0.9, 0.99), net=net, reducer=reducer)
        net._on_bucket = reducer.on_bucket

        def e2e_step(batch):
            opt.zero_grad()
            out = net(batch)
            ls = crit(out, batch)
            ls["loss"].mean().backward()
            opt.step()
            met = evalr(out, batch)
            return float(ls["loss"].item()), float(met["Acc"].item())               # utils.py:426 formats the loss

        def host_batches(n):                                                        # pinned host batches, like a DataLoader
            for i in range(n):
                yield host[i % nres]

        for batch in dat_loader.DevicePrefetcher(host_batches(3), dev):             # warm-up
            e2e_step(batch)
        barrier()
        t0 = time.perf_counter()
        # every step's batch is copied host -> device inside the timed region (utils.py:405-406), one step ahead on a
        # copy stream; every step ends with the device -> host read of its loss and metric
        for batch in dat_loader.DevicePrefetcher(host_batches(args.steps), dev):
            lv, av = e2e_step(batch)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        h2d = sum(v.numel() * v.element_size() for v in host[0].values())
        e2e = {"value": B * world * args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12,
               "ms_per_step": dt / args.steps * 1e3, "api": "mdl.get_default_net(...)(batch) -> loss.ZSGLoss -> backward -> "
               "optim.FusedAdam.step -> evaluator.Evaluator; dat_loader.DevicePrefetcher copies every step's pinned host batch to the device "
               "(one step ahead, on a copy stream)", "last_loss": lv}
        net._on_bucket = None

    # ---------------------------------------------------------------- CPU baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, cores, sec = cpu_oracle_rate(16, 2, 1, model=args.model)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"2 timed steps (+1 warm-up) of bs=16 from the same synthetic distribution, {sec:.1f} s/step: forward, "
                         "loss, backward, Adam, metric in PyTorch CPU (the reference's arithmetic) on all host threads"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "global_batch": B * world, "per_gpu_batch": B, "qlen": 20,
                           "parallelism": f"dp{world}", "step": "forward + loss + backward (+ all-reduce) + Adam + metric",
                           "l2": f"no flush needed: the step streams {eng.nbytes / 2**30:.1f} GiB of activations, far beyond the 126 MB L2; "
                                 f"{nres} resident input batches are rotated"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
                "roofline": roof, "roofline_hbm": roof_hbm, "kernels": kernels,
                "step_tensor_tflops_per_gpu": step_tflops, "step_tensor_frac_of_peak": step_tflops / pk["tflops"],
                "cpu_baseline": cpu,
                "allreduce": {"bytes_per_step": reducer.bytes_reduced / max(1, args.steps + args.warmup + psteps + (0 if args.no_e2e else args.steps + 3)),
                              "buckets_per_step": reducer.calls / max(1, args.steps + args.warmup + psteps + (0 if args.no_e2e else args.steps + 3))}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
