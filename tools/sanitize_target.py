"""Tiny workload for compute-sanitizer: ragged shapes (partial tiles, padding CTAs), every tcgen05 precision."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import diffsinger_b200 as dsx
from oracle import diffnet_oracle as O

dev = torch.device("cuda", 0)
net = bench.make_net(dsx, dev)
S = O.make_schedule(O.linear_beta_schedule(100, 0.06))
smin, smax = bench.lj_spec_minmax()
for prec in sys.argv[1:] or ("fp16x2", "fp16", "fp16x3"):
    s = dsx.DsxSampler(net, prec, 4)
    s.ensure_weights(dev)
    s.set_schedule(S)
    for B, T in ((1, 96), (3, 200), (2, 129)):
        g = torch.Generator().manual_seed(B * 1000 + T)
        cond = torch.randn(B, T, 256, generator=g).transpose(1, 2).to(dev)
        x = torch.randn(B, 1, 80, T, generator=g).to(dev)
        mel = s.infer(cond, 3, smin.to(dev), smax.to(dev), x_start=x, seed=1)
        t = torch.full((B,), 5, dtype=torch.long, device=dev)
        eps = s.diffnet_forward(x, t, cond)
        torch.cuda.synchronize()
        print(prec, B, T, float(mel.abs().mean()), float(eps.abs().mean()), flush=True)
    s.close()
print("done")
