"""Wall-clock (globaltimer) entry / exit of the last stack and head launches of a short DDPM loop: who waits for whom."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
import diffsinger_b200 as dsx
from diffsinger_b200._capi import lib, check
from oracle import diffnet_oracle as O
prec = sys.argv[1] if len(sys.argv) > 1 else "fp16x2"
dev = torch.device("cuda", 0)
net = bench.make_net(dsx, dev)
s = dsx.DsxSampler(net, prec, 1); s.ensure_weights(dev)
s.set_schedule(O.make_schedule(O.linear_beta_schedule(100, 0.06)))
cond, xT = bench.make_inputs(16, 1024, 0)
cond, xT = cond.to(dev).transpose(1, 2), xT.to(dev)
s.sample_ddpm(xT, cond, 100, 20, seed=1)
check(lib.dsx_debug_trace(s._h, 1, None))
s.sample_ddpm(xT, cond, 80, 20, seed=1)        # t_start 80, 20 steps: the last head launch still has the in-projection
buf = np.zeros(6 * 256, dtype=np.int64)
check(lib.dsx_debug_trace(s._h, 0, buf.ctypes.data_as(ctypes.c_void_p)))
ev = []
for i in range(8):
    a, b = buf[220 + 2 * i], buf[221 + 2 * i]
    if a and b: ev.append((int(a), int(b), "stack"))
    a, b = buf[5 * 256 + 220 + 2 * i], buf[5 * 256 + 221 + 2 * i]
    if a and b: ev.append((int(a), int(b), "head "))
ev.sort()
t0 = ev[0][0]
prev_end = None
for a, b, name in ev:
    gap = "" if prev_end is None else f"  gap since previous exit {(a - prev_end) / 1e3:6.1f} us"
    print(f"{name}: entry {(a - t0) / 1e3:8.1f} us  exit {(b - t0) / 1e3:8.1f} us  duration {(b - a) / 1e3:6.1f} us{gap}")
    prev_end = b
st = buf.reshape(2, 3, 256)[0, 2]
print("last head launch, cycles since kernel entry: setup done, H1 acc ready, epi-H1 done, H2 acc ready, mel done, I acc ready, end, exit:",
      [int(v - st[98]) for v in st[100:107]], int(st[99] - st[98]))
