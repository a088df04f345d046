import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
import diffsinger_b200 as dsx
from diffsinger_b200._capi import lib, check
from oracle import diffnet_oracle as O
prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
dev = torch.device("cuda", 0)
net = bench.make_net(dsx, dev)
s = dsx.DsxSampler(net, prec, 1); s.ensure_weights(dev)
s.set_schedule(O.make_schedule(O.linear_beta_schedule(100, 0.06)))
cond, xT = bench.make_inputs(16, 1024, 0)
cond, xT = cond.to(dev).transpose(1, 2), xT.to(dev)
s.sample_ddpm(xT, cond, 100, 3, seed=1)
check(lib.dsx_debug_trace(s._h, 1, None))
s.sample_ddpm(xT, cond, 100, 3, seed=1)     # last head launch: HEAD|UPDATE (no in-proj); the one before: full
buf = np.zeros(6 * 256, dtype=np.int64)
check(lib.dsx_debug_trace(s._h, 0, buf.ctypes.data_as(ctypes.c_void_p)))
st = buf.reshape(2, 3, 256)[0, 2, 100:107]
print(prec, "head stamps (start, H1 acc ready, epi-H1 done, H2 acc ready, mel done, I acc ready, end):", [int(v - st[0]) for v in st])
