"""Where does a DDPM step go?  Whole loop (events) vs the sum of the stack launches vs the sum of the head launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import diffsinger_b200 as dsx
from diffsinger_b200 import _capi
from oracle import diffnet_oracle as O
dev = torch.device("cuda", 0)
net = bench.make_net(dsx, dev)
for prec in sys.argv[1:] or ("fp16x2",):
    s = dsx.DsxSampler(net, prec, 1)
    s.ensure_weights(dev)
    s.set_schedule(O.make_schedule(O.linear_beta_schedule(100, 0.06)))
    cond, xT = bench.make_inputs(16, 1024, 0)
    cond, xT = cond.to(dev).transpose(1, 2), xT.to(dev)
    K = 100
    for rep in range(2):
        s.sample_ddpm(xT, cond, 100, K, seed=1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); s.sample_ddpm(xT, cond, 100, K, seed=1); e1.record(); torch.cuda.synchronize()
    total = e0.elapsed_time(e1)
    out = [f"{prec}: {K} DDPM steps {total:.2f} ms = {total / K * 1e3:.0f} us/step"]
    for mode, name in ((1, "layer stack"), (2, "head")):
        s.set_option(_capi.OPT_PROFILE, mode)
        s.sample_ddpm(xT, cond, 100, K, seed=1)
        ns, n = s.info(_capi.INFO_LAYER_KERNEL_NS), s.info(_capi.INFO_LAYER_KERNEL_LAUNCHES)
        s.set_option(_capi.OPT_PROFILE, 0)
        out.append(f"{name} {ns / 1e3 / max(n, 1):.1f} us x {n}")
    print("; ".join(out), flush=True)
    s.close()
