"""Per-layer time of the residual stack for each precision and tuning-knob setting (DSX_OPT_PROFILE events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import diffsinger_b200 as dsx
from diffsinger_b200 import _capi

dev = torch.device("cuda", 0)
net = bench.make_net(dsx, dev)
cond, xT = bench.make_inputs(16, 1024, 0)
cond, xT = cond.to(dev).transpose(1, 2), xT.to(dev)
t = torch.full((16,), 50, dtype=torch.long, device=dev)
for prec in ("fp16x2", "fp16", "fp16x3"):
    s = dsx.DsxSampler(net, prec, 1)
    s.ensure_weights(dev)
    for mode in (0, 1):
        s.set_option(_capi.OPT_CP_PREFETCH, mode)
        for _ in range(3):
            s.diffnet_forward(xT, t, cond)
        s.set_option(_capi.OPT_PROFILE, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            s.diffnet_forward(xT, t, cond)
        e1.record()
        torch.cuda.synchronize()
        ns, n = s.info(_capi.INFO_LAYER_KERNEL_NS), s.info(_capi.INFO_LAYER_KERNEL_LAUNCHES)
        s.set_option(_capi.OPT_PROFILE, 0)
        print(f"{prec} cp_prefetch={mode}: stack {ns / max(n, 1) / 1e3:.1f} us per launch -> {ns / max(n, 1) / 20e3:.2f} us/layer; "
              f"whole forward (cond pack + condproj + in-proj + stack + head) {e0.elapsed_time(e1) / 20 * 1e3:.0f} us", flush=True)
    s.close()
