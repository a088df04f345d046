"""Small fixed workload for ncu: the config-2 shape (B=16, T=1024), a few DDPM steps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import diffsinger_b200 as dsx
from oracle import diffnet_oracle as O

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16x2"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
net = bench.make_net(dsx, dev)
s = dsx.DsxSampler(net, prec, 1)
s.ensure_weights(dev)
s.set_schedule(O.make_schedule(O.linear_beta_schedule(100, 0.06)))
cond, xT = bench.make_inputs(16, 1024, 0)
smin, smax = bench.lj_spec_minmax()
mel = s.infer(cond.to(dev).transpose(1, 2), K, smin.to(dev), smax.to(dev), x_start=xT.to(dev), seed=1)
torch.cuda.synchronize()
print("ok", float(mel.abs().mean()))
