"""Development checks of the register-resident stack kernel (dsx_stack.cu) on a GPU box.

    python tools/dev_stack.py parity      # golden / oracle parity of every mode, old kernel vs new kernel
    python tools/dev_stack.py timing      # config-2 loop timings per mode and kernel
    python tools/dev_stack.py trace [B T] # clock64 timeline of one evaluation (CTA 0 / 1); default 16 x 1024
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import diffsinger_b200 as dsx  # noqa: E402
from diffsinger_b200 import _capi  # noqa: E402
from oracle import diffnet_oracle as O  # noqa: E402  (checker only)
from conftest import HP, golden, rs_normal  # noqa: E402

DEV = torch.device("cuda", 0)


def make(cycle, prec, stack_kernel=1, schedule=None, sets=None, opts=()):
    hp = dict(HP, dilation_cycle_length=cycle)
    torch.manual_seed(0)
    net = dsx.DiffNet(80, hparams=hp)
    torch.nn.init.normal_(net.output_projection.weight, std=0.02)
    net = net.to(DEV).eval()
    s = dsx.DsxSampler(net, prec, cycle)
    s._handle(DEV)
    if sets is not None:
        s.set_option(_capi.OPT_SR_SETS, sets)
    s.ensure_weights(DEV)
    s.set_option(_capi.OPT_STACK_KERNEL, stack_kernel)
    for k, v in opts:
        s.set_option(k, v)
    if schedule is not None:
        s.set_schedule(schedule)
    return s


def quick():
    """first contact: one evaluation of the new kernel per mode against the golden vectors; aborts on the first failure"""
    g = golden("diffnet_fwd_cycle4.npz")
    spec, cond, t = (torch.from_numpy(g[k]).to(DEV) for k in ("spec", "cond", "t"))
    B, _, M, T = g["spec"].shape
    for prec, rows in (("fp16x2", 128), ("fp16s", 128), ("fp16", 128), ("fp16x2", 64), ("fp16s", 64)):
        s = make(4, prec, 1, opts=((_capi.OPT_STACK_ROWS, rows),))
        eps = s.diffnet_forward(spec, t, cond).cpu().numpy()
        x_last = s.debug_read(0, B, T).cpu().numpy()
        skip = s.debug_read(1, B, T).cpu().numpy()
        print(f"quick {prec} rows {rows}/{s.info(_capi.INFO_STACK_ROWS)}: eps {np.abs(eps - g['eps']).max():.3e} x20 {np.abs(x_last[1].T - g['x20_b1']).max():.3e} "
              f"skip {np.abs(skip[0].T - g['skip_sum_b0']).max():.3e} stack launches {s.info(_capi.INFO_STACK_KERNEL_LAUNCHES)}",
              flush=True)
        s.close()


def parity():
    rc, rep = dsx.selftest(0)
    print("selftest", rc, flush=True)
    res = {}
    for cycle in (1, 4):
        g = golden(f"diffnet_fwd_cycle{cycle}.npz")
        spec, cond, t = (torch.from_numpy(g[k]).to(DEV) for k in ("spec", "cond", "t"))
        B, _, M, T = g["spec"].shape
        for prec, sk in (("fp16x2", 0), ("fp16x2", 1), ("fp16s", 1), ("fp16", 1), ("fp16", 0)):
            s = make(cycle, prec, sk)
            try:
                eps = s.diffnet_forward(spec, t, cond).cpu().numpy()
                x_last = s.debug_read(0, B, T).cpu().numpy()
                skip = s.debug_read(1, B, T).cpu().numpy()
                s.set_layer_limit(1)
                s.diffnet_forward(spec, t, cond)
                x1 = s.debug_read(0, B, T).cpu().numpy()
                s.set_layer_limit(-1)
                r = dict(eps=float(np.abs(eps - g["eps"]).max()), x20=float(np.abs(x_last[1].T - g["x20_b1"]).max()),
                         skip=float(np.abs(skip[0].T - g["skip_sum_b0"]).max()), x1=float(np.abs(x1[0].T - g["x1_b0"]).max()),
                         launches=s.info(1))
            except Exception as e:  # noqa: BLE001
                r = dict(error=str(e))
            res[f"fwd_c{cycle}_{prec}_k{sk}"] = r
            print(f"fwd cycle{cycle} {prec} stack_kernel={sk}: {r}", flush=True)
            s.close()
    # ragged shapes: new kernel vs old kernel (fp16x2: same operands, different accumulation order)
    for B, T in ((1, 96), (3, 333), (1, 300), (2, 1000), (40, 520), (5, 128), (2, 129)):
        outs = {}
        for sk in (0, 1):
            s = make(4, "fp16x2", sk)
            x, cond = rs_normal(40 + B, (B, 1, 80, T)).to(DEV), rs_normal(50 + T, (B, 256, T)).to(DEV)
            t = torch.full((B,), 7, dtype=torch.long, device=DEV)
            try:
                a = s.diffnet_forward(x, t, cond).cpu()
                b = s.diffnet_forward(x, t + 1, cond).cpu()
                a2 = s.diffnet_forward(x, t, cond).cpu()
                outs[sk] = (a, b, bool(torch.equal(a, a2)))
            except Exception as e:  # noqa: BLE001
                outs[sk] = str(e)
            s.close()
        if isinstance(outs[0], tuple) and isinstance(outs[1], tuple):
            d = max((outs[0][0] - outs[1][0]).abs().max().item(), (outs[0][1] - outs[1][1]).abs().max().item())
            res[f"ragged_{B}x{T}"] = dict(diff=d, repeatable=outs[1][2], scale=outs[0][0].abs().max().item())
        else:
            res[f"ragged_{B}x{T}"] = dict(error=str(outs))
        print(f"ragged {B}x{T}: {res[f'ragged_{B}x{T}']}", flush=True)
    # K = 100 DDPM golden loop
    g = golden("ddpm_lj_K100.npz")
    S = O.make_schedule(O.linear_beta_schedule(100, 0.06))
    cond, xT = torch.from_numpy(g["cond"]).to(DEV), torch.from_numpy(g["xT"]).to(DEV)
    noise = rs_normal(int(g["noise_seed"]), (100,) + tuple(g["xT"].shape)).to(DEV)
    for prec, sk, sets in (("fp16x2", 0, None), ("fp16x2", 1, None), ("fp16s", 1, 64), ("fp16s", 1, 16), ("fp16s", 1, 1),
                           ("fp16", 1, None), ("fp16", 0, None)):
        s = make(1, prec, sk, S, sets)
        try:
            x0 = s.sample_ddpm(xT, cond, 100, 100, noise=noise).cpu().numpy()
            d = np.abs(x0 - g["x0"])
            r = dict(max=float(d.max()), mae=float(d.mean()))
        except Exception as e:  # noqa: BLE001
            r = dict(error=str(e))
        res[f"ddpm100_{prec}_k{sk}_sets{sets}"] = r
        print(f"ddpm K=100 {prec} stack_kernel={sk} sets={sets}: {r}", flush=True)
        s.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "dev_parity.json"), "w"), indent=1)


def experiments():
    """tuning knobs: publish mode (fences of the halo hand-over) and the approximate gate, timing + K=100 golden error"""
    S = O.make_schedule(O.linear_beta_schedule(100, 0.06))
    g = golden("ddpm_lj_K100.npz")
    condg, xTg = torch.from_numpy(g["cond"]).to(DEV), torch.from_numpy(g["xT"]).to(DEV)
    noise = rs_normal(int(g["noise_seed"]), (100,) + tuple(g["xT"].shape)).to(DEV)
    B, T = 16, 1024
    gen = torch.Generator().manual_seed(1)
    cond = torch.randn(B, T, 256, generator=gen).transpose(1, 2).to(DEV)
    x = torch.randn(B, 1, 80, T, generator=gen).to(DEV)
    res = {}
    for prec in ("fp16s", "fp16x2"):
        for pm in (0,):
            for ga in (0, 1):
                s = make(1, prec, 1, S, opts=((_capi.OPT_GATE_APPROX, ga),))
                x0 = s.sample_ddpm(xTg, condg, 100, 100, noise=noise).cpu().numpy()
                d = np.abs(x0 - g["x0"])
                s.sample_ddpm(x, cond, 100, 4, seed=1)
                torch.cuda.synchronize()
                s.set_option(_capi.OPT_PROFILE, 1)
                ref = None
                for rep in range(3):
                    out = s.sample_ddpm(x, cond, 100, 10, seed=1)
                    ref = out if ref is None else ref
                    assert torch.equal(out, ref), "non-deterministic result (race?)"
                torch.cuda.synchronize()
                layer_ms = s.info(_capi.INFO_LAYER_KERNEL_NS) / 1e6 / max(s.info(_capi.INFO_LAYER_KERNEL_LAUNCHES), 1)
                s.set_option(_capi.OPT_PROFILE, 0)
                r = dict(max=float(d.max()), mae=float(d.mean()), stack_ms=layer_ms,
                         stack_tflops=B * T * 20 * 1048576 / (layer_ms * 1e-3) / 1e12)
                res[f"{prec}_pm{pm}_ga{ga}"] = r
                print(f"exp {prec} publish_mode={pm} gate_approx={ga}: {r}", flush=True)
                s.close()
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "dev_experiments.json"), "w"), indent=1)


def small():
    """small batches: 64 vs 128 frames per CTA (BASELINE config 1 and the per-GPU shard of the strong-scaled config 4)"""
    S = O.make_schedule(O.linear_beta_schedule(100, 0.06))
    res = {}
    for (B, T) in ((1, 512), (4, 1024), (8, 1024), (2, 2048)):
        gen = torch.Generator().manual_seed(1)
        cond = torch.randn(B, T, 256, generator=gen).transpose(1, 2).to(DEV)
        x = torch.randn(B, 1, 80, T, generator=gen).to(DEV)
        for prec in ("fp16s", "fp16x2"):
            outs = {}
            for rows in (128, 64):
                s = make(1, prec, 1, S, opts=((_capi.OPT_STACK_ROWS, rows),))
                K = 20
                outs[rows] = s.sample_ddpm(x, cond, 100, 5, seed=1)
                torch.cuda.synchronize()
                s.set_option(_capi.OPT_PROFILE, 1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                s.sample_ddpm(x, cond, 100, K, seed=1)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / K
                layer_ms = s.info(_capi.INFO_LAYER_KERNEL_NS) / 1e6 / max(s.info(_capi.INFO_LAYER_KERNEL_LAUNCHES), 1)
                r = dict(ms_per_step=ms, stack_ms=layer_ms, frames_per_s_K100=B * T / (ms * 100 / 1e3), rows=s.info(_capi.INFO_STACK_ROWS))
                res[f"{B}x{T}_{prec}_rows{rows}"] = r
                print(f"small {B}x{T} {prec} rows={rows}: {r}", flush=True)
                s.close()
            print(f"   64 vs 128 rows, 5 DDPM steps (Philox noise): max diff {(outs[64] - outs[128]).abs().max().item():.3e}", flush=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "dev_small.json"), "w"), indent=1)


def k100():
    """the K = 100 DDPM golden loop (injected noise) in the stack-kernel modes: max / mean |dx| against the live-reference golden"""
    g = golden("ddpm_lj_K100.npz")
    S = O.make_schedule(O.linear_beta_schedule(100, 0.06))
    cond, xT = torch.from_numpy(g["cond"]).to(DEV), torch.from_numpy(g["xT"]).to(DEV)
    noise = rs_normal(int(g["noise_seed"]), (100,) + tuple(g["xT"].shape)).to(DEV)
    for prec, sets in (("fp16s", 64), ("fp16x2", None), ("fp16", None)):
        s = make(1, prec, 1, S, sets)
        x0 = s.sample_ddpm(xT, cond, 100, 100, noise=noise).cpu().numpy()
        d = np.abs(x0 - g["x0"])
        print(f"ddpm K=100 {prec} sets={sets}: max {d.max():.3e} mae {d.mean():.3e}", flush=True)
        s.close()


def timing():
    S = O.make_schedule(O.linear_beta_schedule(100, 0.06))
    res = {}
    for (B, T) in ((16, 1024), (1, 512), (4, 1024)):
        gen = torch.Generator().manual_seed(1)
        cond = torch.randn(B, T, 256, generator=gen).transpose(1, 2).to(DEV)
        x = torch.randn(B, 1, 80, T, generator=gen).to(DEV)
        for prec, sk in (("fp16x2", 0), ("fp16x2", 1), ("fp16s", 1), ("fp16", 1), ("fp16", 0)):
            s = make(1, prec, sk, S)
            K = 20
            try:
                s.sample_ddpm(x, cond, 100, 5, seed=1)
                torch.cuda.synchronize()
                s.set_option(_capi.OPT_PROFILE, 1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                s.sample_ddpm(x, cond, 100, K, seed=1)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / K
                layer_ms = s.info(_capi.INFO_LAYER_KERNEL_NS) / 1e6 / max(s.info(_capi.INFO_LAYER_KERNEL_LAUNCHES), 1)
                s.set_option(_capi.OPT_PROFILE, 0)
                flops = B * T * 20 * 1048576
                r = dict(ms_per_step=ms, stack_ms=layer_ms, frames_per_s_K100=B * T / (ms * 100 / 1e3),
                         stack_tflops=flops / (layer_ms * 1e-3) / 1e12)
            except Exception as e:  # noqa: BLE001
                r = dict(error=str(e))
            res[f"{B}x{T}_{prec}_k{sk}"] = r
            print(f"timing {B}x{T} {prec} stack_kernel={sk}: {r}", flush=True)
            s.close()
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "dev_timing.json"), "w"), indent=1)


def trace():
    import ctypes
    B, T = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (16, 1024)     # e.g. `trace 1 512`: 64-frame tiles
    gen = torch.Generator().manual_seed(1)
    cond = torch.randn(B, T, 256, generator=gen).transpose(1, 2).to(DEV)
    x = torch.randn(B, 1, 80, T, generator=gen).to(DEV)
    t = torch.full((B,), 57, dtype=torch.long, device=DEV)
    out = {}
    for prec in ("fp16s", "fp16x2"):
        s = make(1, prec, 1, O.make_schedule(O.linear_beta_schedule(100, 0.06)))
        xs = torch.randn(B, 1, 80, T, generator=gen).to(DEV)
        s.sample_ddpm(xs, cond, 100, 2, seed=1)          # the loop entry point: no debug taps, as in production
        buf = (ctypes.c_int64 * (10 * 256))()
        _capi.check(_capi.lib.dsx_debug_trace(s._h, 1, None), "trace on")
        s.sample_ddpm(xs, cond, 100, 2, seed=1)
        _capi.check(_capi.lib.dsx_debug_trace(s._h, 2, ctypes.cast(buf, ctypes.c_void_p)), "trace read")
        full = np.array(buf[:], dtype=np.int64)
        a = full[:6 * 256].reshape(2, 3, 256)
        cta = full[6 * 256:].reshape(256, 4)
        cta = cta[cta[:, 0] != 0]
        if len(cta):
            t_in, t_out = cta[:, 0], cta[:, 2]
            cyc = cta[:, 3] - cta[:, 1]
            ns = t_out - t_in
            print(f"--- {prec}: {len(cta)} CTAs; wall ns: first entry -> last exit {int(t_out.max() - t_in.min())}, entry spread "
                  f"{int(t_in.max() - t_in.min())}, exit spread {int(t_out.max() - t_out.min())}; per CTA ns median {int(np.median(ns))} "
                  f"max {int(ns.max())}; cycles median {int(np.median(cyc))}; MHz inside the kernel {float(np.median(cyc / ns)) * 1e3:.0f}; "
                  f"CTA 0: entry -> first G1 {int(a[0, 1, 0] - cta[0, 1])} cycles, x0 stored -> exit {int(cta[0, 3] - a[0, 2, 253])} cycles")
        out[prec] = a.tolist()
        base = a[0, 1, 0]
        print(f"--- {prec}: MMA thread (CTA 0) per layer: G1 start, centre taps issued, halo landed, G1 issued, G2 kb0 start, G2 kb2 start, G2 issued")
        for l in range(0, 6):
            print(l, [int(a[0, 1, l * 8 + k] - base) for k in range(7)])
        print("epilogue (warp 4): per layer e1c0 [enter, acc ready, done], e1c1 [...], e2 [enter, ready, done]")
        for l in range(0, 6):
            print(l, [int(a[0, 2, l * 12 + k] - base) for k in range(9)])
        print("producer 0 (relative to the kernel's entry): y0 loads issued from", int(a[0, 0, 0] - cta[0, 1]), "to", int(a[0, 0, 1] - cta[0, 1]),
              "| entry prefetches issued", int(a[0, 0, 2] - cta[0, 1]), "| first G1", int(a[0, 1, 0] - cta[0, 1]))
        print("producer 0: per layer [-, -, -, z stored]")
        for l in range(1, 6):
            print(l, [int(a[0, 0, l * 4 + k] - base) for k in range(4)])
        print("end of skip GEMM issue (MMA thread)", int(a[0, 1, 250] - base), "| exit epilogue: enter", int(a[0, 2, 248] - base),
              "skip sum ready", int(a[0, 2, 249] - base), "done", int(a[0, 2, 250] - base), "| layer 19 G2 issued", int(a[0, 1, 19 * 8 + 6] - base))
        print("fused head: MMA issue end", int(a[0, 1, 251] - base), "| epilogue: h tiles", int(a[0, 2, 251] - base), "mel phase",
              int(a[0, 2, 252] - base), "x0 / y0 stored", int(a[0, 2, 253] - base))
        per_layer = (a[0, 1, 19 * 8] - a[0, 1, 1 * 8]) / 18.0
        print(f"{prec}: cycles per layer (MMA thread, layers 1..19): {per_layer:.0f}")
        s.close()
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "dev_trace.json"), "w"))


if __name__ == "__main__":
    {"quick": quick, "k100": k100, "small": small, "experiments": experiments, "parity": parity, "timing": timing, "trace": trace}[sys.argv[1]]()
