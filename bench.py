#!/usr/bin/env python
"""bench.py -- mel-frames/s of the full reverse-diffusion loop (BASELINE.json metric).

One "step" = one complete pass of the hot path over one batch: the infer branch of
GaussianDiffusion.forward after the conditioner (gaussian start, K DDPM steps through DiffNet, denorm
epilogue) for B utterances of T frames -- BASELINE.json configs[1]: DiffSpeech B=16, T=1024, K=100 on one
B200.  Synthetic cond / x_T (seeded), random-init weights of the real architecture.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--precision fp16x3|fp16|fp32] [--impl reference]

N > 1: launched by torchrun, one rank per GPU, utterance-sharded (weak scaling: B per GPU fixed), one
all-gather of the finished mels per step inside the timed region.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import subprocess
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

HP = dict(hidden_size=256, residual_layers=20, residual_channels=256, dilation_cycle_length=1,
          audio_num_mel_bins=80, keep_bins=80)
# algorithmic FLOPs per mel frame: whole DiffNet evaluation, conditioner projection hoisted
_JSON_OUT = sys.stdout
FLOP_PER_FRAME_EVAL = 21184512
# one residual-layer kernel launch, per frame: 2*(3*256*512 + 256*512)  (dilated conv + output projection)
FLOP_PER_FRAME_LAYER = 2 * (3 * 256 * 512 + 256 * 512)


def lj_spec_minmax():
    # value range of usr/configs/lj_ds_beta6.yaml (only its scale matters: it is a per-bin affine map)
    return torch.linspace(-4.6, -5.3, 80), torch.linspace(0.7, -0.2, 80)


def make_net(dsx, dev=None):
    torch.manual_seed(0)
    net = dsx.DiffNet(80, hparams=HP)
    torch.nn.init.normal_(net.output_projection.weight, std=0.02)
    return net.eval() if dev is None else net.to(dev).eval()


def make_inputs(B, T, rank):
    g = torch.Generator().manual_seed(1234 + 7919 * rank)
    cond = torch.randn(B, T, 256, generator=g)            # frames-major, handed over as the transposed view
    g2 = torch.Generator().manual_seed(1235 + 7919 * rank)
    xT = torch.randn(B, 1, 80, T, generator=g2)
    return cond, xT


class ClockSampler:
    """Samples SM clock / throttle reasons DURING the timed region with an nvidia-smi subprocess (the recipe's
    clocks line); a separate process so the sampling never contends with the launching thread."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.path = f"/tmp/dsx_clocks_{os.getpid()}_{index}.csv"
        self.proc = None
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def start(self):
        time.sleep(0.3)          # let the first samples land

    def finish(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        self.f.close()
        clk, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                clk.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        try:
            os.remove(self.path)
        except OSError:
            pass
        # samples under load only (idle samples sit at the floor clock)
        load = [c for c in clk if c > 0.5 * (mx or 1)] or clk
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(clk)}


def pick_cpu_threads():
    """MKL-DNN does not always scale to every core of a big host: time a tiny p_sample at a few thread counts
    and keep the fastest (so the CPU baseline is the best the host can do, not an oversubscribed one)."""
    from oracle import diffnet_oracle as O
    n = os.cpu_count() or 1
    cands = sorted({n, min(n, 64), min(n, 32), min(n, 16)}, reverse=True)
    sd = O.build_state_dict(0)
    S = O.make_schedule(O.linear_beta_schedule(100, 0.06))
    cond, x = make_inputs(4, 256, 0)
    cond = cond.transpose(1, 2)
    best = (None, 1e30)
    for c in cands:
        torch.set_num_threads(c)
        with torch.no_grad():
            O.p_sample(sd, S, x, 50, cond, x)
            t0 = time.perf_counter()
            for _ in range(2):
                O.p_sample(sd, S, x, 50, cond, x)
            dt = time.perf_counter() - t0
        if dt < best[1]:
            best = (c, dt)
    return best[0]


def cpu_baseline(B, T, K, n_evals, threads=None):
    """The reference's CPU algorithm (oracle port, torch CPU fp32 == the ATen kernels the reference runs) on
    this box's host cores: n_evals DDPM steps of the same workload, extrapolated linearly to K (the cost of a
    p_sample step does not depend on t)."""
    from oracle import diffnet_oracle as O
    cores = threads or pick_cpu_threads()
    torch.set_num_threads(cores)
    sd = O.build_state_dict(0)
    S = O.make_schedule(O.linear_beta_schedule(100, 0.06))
    cond, x = make_inputs(B, T, 0)
    cond = cond.transpose(1, 2)
    g = torch.Generator().manual_seed(1236)
    with torch.no_grad():
        noise = torch.randn(x.shape, generator=g)
        x = O.p_sample(sd, S, x, K - 1, cond, noise)          # warm-up (thread pool, allocator)
        t0 = time.perf_counter()
        for j in range(n_evals):
            x = O.p_sample(sd, S, x, K - 2 - j, cond, noise)
        dt = (time.perf_counter() - t0) / n_evals
    return dt, cores


def run_reference_arm(args, rank, world):
    """`--impl reference`: the reference's own CPU path (oracle port) timed on the host cores."""
    if rank != 0:
        return
    B, T, K = args.B, args.T, args.K
    n_evals = args.ref_evals
    per = []
    threads = pick_cpu_threads()
    for i in range(min(args.warmup, 1) + args.steps):
        dt, cores = cpu_baseline(B, T, K, n_evals, threads)
        if i >= min(args.warmup, 1):
            per.append(dt)
    step_s = float(np.mean(per)) * K
    value = B * T / step_s
    line = {
        "impl": "reference", "metric": "mel-frames/s", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"DiffSpeech B={B} T_frames={T} K={K} DDPM, full reverse loop (configs[1])",
                   "B": B, "T_frames": T, "K": K, "layers": 20, "channels": 256},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{n_evals} p_sample steps at B={B},T={T} per bench step, x{K}/{n_evals} extrapolated "
                                   "(per-step cost is independent of t)"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="dsx", choices=["dsx", "reference"])
    ap.add_argument("--precision", default=os.environ.get("DSX_BENCH_PRECISION", "fp16x2"))
    ap.add_argument("--cta-group", type=int, default=0)
    ap.add_argument("--B", type=int, default=16)
    ap.add_argument("--T", type=int, default=1024)
    ap.add_argument("--K", type=int, default=100)
    ap.add_argument("--ref-evals", type=int, default=1)
    ap.add_argument("--cpu-evals", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    args = ap.parse_args()

    # stdout carries exactly ONE JSON line: everything libraries print there (NCCL's version banner, ...) goes to stderr
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import __graft_entry__
    __graft_entry__.build()
    import diffsinger_b200 as dsx
    from diffsinger_b200 import _capi
    from diffsinger_b200.parallel import all_gather_batch
    from oracle import diffnet_oracle as O

    B, T, K = args.B, args.T, args.K
    net = make_net(dsx, dev)
    S = O.make_schedule(O.linear_beta_schedule(100, 0.06))       # LJ DiffSpeech schedule (lj_ds_beta6.yaml)
    smin, smax = lj_spec_minmax()
    cond_h, xT_h = make_inputs(B, T, rank)

    def build_sampler(prec):
        s = dsx.DsxSampler(net, prec, 1)
        s.ensure_weights(dev)
        if args.cta_group and prec != "fp32":
            s.set_option(_capi.OPT_TC_CTA_GROUP, args.cta_group)
        s.set_schedule(S)
        return s

    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)       # > 126 MB L2

    def measure(prec, steps, warmup, gather=True, clocks=False):
        s = build_sampler(prec)
        cond = cond_h.to(dev).transpose(1, 2)
        xT = xT_h.to(dev)
        smin_d, smax_d = smin.to(dev), smax.to(dev)

        def step(i):
            mel = s.infer(cond, K, smin_d, smax_d, x_start=xT, seed=1236 + i)
            if world > 1 and gather:
                mel = all_gather_batch(mel, B * world)
            return mel

        for i in range(warmup):
            step(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        sampler_thread = ClockSampler(local_rank) if clocks else None
        if sampler_thread:
            sampler_thread.start()
        l0 = s.info(_capi.INFO_KERNEL_LAUNCHES)
        times = []
        for i in range(steps):
            flush.zero_()                                          # L2 flush between timed iterations
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            mel = step(warmup + i)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        l1 = s.info(_capi.INFO_KERNEL_LAUNCHES)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        clk = sampler_thread.finish() if sampler_thread else None
        total_ms = torch.tensor([sum(times)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        assert torch.isfinite(mel).all()
        return s, float(total_ms.item()), (l1 - l0), clk, mel

    s, total_ms, launches, clk, mel = measure(args.precision, args.steps, args.warmup, clocks=True)
    ms_per_step = total_ms / args.steps
    value = world * B * T / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (the fused residual-layer kernel), measured live --------------------
    roof = None
    if args.precision != "fp32":
        s.set_option(_capi.OPT_PROFILE, 1)
        cond = cond_h.to(dev).transpose(1, 2)
        s.infer(cond, K, smin.to(dev), smax.to(dev), x_start=xT_h.to(dev), seed=99)
        ns, n = s.info(_capi.INFO_LAYER_KERNEL_NS), s.info(_capi.INFO_LAYER_KERNEL_LAUNCHES)
        s.set_option(_capi.OPT_PROFILE, 0)
        avg_s = ns * 1e-9 / max(n, 1) / 20          # brackets are per evaluation (all 20 residual layers)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        traffic = None
        try:     # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel, per launch
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_layer_traffic.json")))[args.precision]
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        except Exception:
            pass
        ach = FLOP_PER_FRAME_LAYER * B * T / avg_s / 1e12
        # executed MMA passes: the conditioner projection is hoisted out of the loop (k_tc_condproj, once per call), so
        # the kernel executes exactly the algorithmic FLOPs times the number of hi/lo passes
        passes = {"fp16x3": 3.0, "fp16x2": 2.0, "fp16": 1.0}[args.precision]
        roof = {"bound": "tensor", "kernel": "k_tc_layer (fused residual-layer stack, tcgen05; time per layer = stack time / 20)", "achieved": ach, "peak": peak,
                "unit": "TFLOP/s", "frac": ach / peak,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained",
                "traffic": traffic,
                "traffic_note": "DRAM bytes per launch (= 20 residual layers) from profiles/r01_layer_traffic.json (ncu --set full); "
                                "671 MB of it is the hoisted conditioner projection streamed once per evaluation",
                "flops_per_launch": FLOP_PER_FRAME_LAYER * B * T * 20, "avg_launch_us": avg_s * 20e6,
                "avg_layer_us": avg_s * 1e6, "evaluations_profiled": n,
                "layer_kernels_share_of_step": ns * 1e-6 / ms_per_step,
                "mma_passes": passes,
                "executed_tflops": FLOP_PER_FRAME_LAYER * passes * B * T / avg_s / 1e12,
                "whole_step_algorithmic_tflops": FLOP_PER_FRAME_EVAL * B * T * K / (ms_per_step * 1e-3) / 1e12}
    s.close()

    # ---- e2e: same metric through the C ABI with HOST buffers (H2D of cond + x_T, D2H of mel inside) ------
    s2 = build_sampler(args.precision)
    cond_p = cond_h.pin_memory()
    xT_p = xT_h.pin_memory()
    out_p = torch.empty(B, T, 80).pin_memory()
    cond_view = cond_p.transpose(1, 2)
    e2e_steps = max(2, min(args.steps, 3))
    for i in range(2):
        s2.infer_host(cond_view, K, smin, smax, x_start=xT_p, seed=5 + i, out=out_p, device=dev)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        s2.infer_host(cond_view, K, smin, smax, x_start=xT_p, seed=50 + i, out=out_p, device=dev)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([(time.perf_counter() - t0) / e2e_steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e = {"value": world * B * T / float(e2e_s.item()), "unit": "frames/s",
           "h2d_bytes_per_step": int(cond_p.numel() * 4 + xT_p.numel() * 4 + 2 * 80 * 4),
           "d2h_bytes_per_step": int(out_p.numel() * 4), "ms_per_step": float(e2e_s.item()) * 1e3,
           "api": "dsx_infer_host (C ABI, pinned host buffers)"}
    s2.close()

    extra = {}
    notes = {"fp16": "single MMA pass, fp16 operands (conditioner projection exact): mel MAE 7e-5, max |d| 1.6e-3 after 100 steps (tests)",
             "fp16x2": "weights hi/lo split, 2 MMA passes, conditioner projection exact: max |d| 1.5e-4 after 100 steps (tests)",
             "fp16x3": "hi/lo split of both operands, 3 MMA passes: max |d| 1.4e-5 after 100 steps (tests)"}
    if not args.no_extra and world == 1:
        for other in ("fp16x3", "fp16x2", "fp16"):
            if other == args.precision:
                continue
            so, tms, _, _, _ = measure(other, 2, 2, clocks=False)
            so.close()
            extra[other] = {"value": B * T / (tms / 2 * 1e-3), "unit": "frames/s", "ms_per_step": tms / 2, "note": notes[other]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        dt, cores = cpu_baseline(B, T, K, args.cpu_evals)
        cpu = {"value": B * T / (dt * K), "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": f"{args.cpu_evals} p_sample steps at B={B},T={T} (oracle port of the reference's PyTorch-CPU path), "
                         f"x{K}/{args.cpu_evals} extrapolated", "ms_per_diffnet_step": dt * 1e3}

    if rank == 0:
        line = {
            "metric": "mel-frames/s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"fp16x2": "f16 operands, weights hi+lo split (2 MMA passes), conditioner projection hoisted (f32), f32 accumulate and state",
                                           "fp16x3": "f16 hi+lo split x3 MMA, f32 accumulate (fp32-equivalent)",
                                           "fp16": "f16 operands, f32 accumulate", "fp32": "f32"}[args.precision],
            "data": "synthetic",
            "config": {"workload": f"DiffSpeech B={B} T_frames={T} K={K} DDPM gaussian start, full reverse loop + denorm "
                                   "(BASELINE.json configs[1])", "B_per_gpu": B, "T_frames": T, "K": K, "layers": 20,
                       "channels": 256, "precision": args.precision, "precision_note": notes[args.precision], "noise": "in-kernel Philox4x32-10",
                       "l2": "256 MB buffer written between timed iterations (L2 flush)",
                       "parallelism": f"utterance-sharded x{world}, one all-gather per step" if world > 1 else "single GPU"},
            "diffnet_step_ms": ms_per_step / K,
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
            "extra": extra,
        }
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
