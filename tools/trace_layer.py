"""Timeline of one residual-layer kernel launch (CTA 0/1): clock64 stamps of the producer, MMA issuer and epilogue."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import diffsinger_b200 as dsx
from diffsinger_b200._capi import lib, check

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
group = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
net = bench.make_net(dsx, dev)
s = dsx.DsxSampler(net, prec, 1)
s.ensure_weights(dev)
s.set_option(0, group)
cond, xT = bench.make_inputs(16, 1024, 0)
cond, xT = cond.to(dev).transpose(1, 2), xT.to(dev)
t = torch.full((16,), 50, dtype=torch.long, device=dev)
for _ in range(2):
    s.diffnet_forward(xT, t, cond)
check(lib.dsx_debug_trace(s._h, 1, None))
s.set_layer_limit(6)
s.diffnet_forward(xT, t, cond)
buf = np.zeros(6 * 256, dtype=np.int64)
check(lib.dsx_debug_trace(s._h, 0, buf.ctypes.data_as(ctypes.c_void_p)))
tr = buf.reshape(2, 3, 256)
P = {"fp16": 1, "fp16x2": 2, "fp16x3": 3}[prec]
U1, U2 = 1, (1 if P == 1 else 2)      # MMA stamps are per k-block now (slot = order index + 16*chunk)
for cta in (0, 1):
    t0 = tr[cta, 0, 254]
    rel = lambda v: int(v - t0) if v else -1
    print(f"--- CTA {cta} ({prec}); cycles since setup done; all roles done at {rel(tr[cta,0,255])}; "
          f"kernel entry at {rel(tr[cta,0,250])}, exit at {rel(tr[cta,0,251])}; entry->exit {(tr[cta,1,251]-tr[cta,1,250])/1e3:.1f} us (globaltimer)")
    print("producer ring1 issue (first units):", [rel(v) for v in tr[cta, 0, 0:40]])
    print("producer ring2 issue:", [rel(v) for v in tr[cta, 0, 128:128 + 8 * U2]])
    print("mma k-block start (GEMM1, chunk 0 then chunk 1, order cond|centre|halo):", [rel(v) for v in tr[cta, 1, 0:32]])
    print("mma k-block start (GEMM2):", [rel(v) for v in tr[cta, 1, 128:128 + 8 * U2:U2]])
    print("mma: before zfull wait, after zfull, tempty q0, q1:", [rel(v) for v in tr[cta, 1, 200:204]])
    e = [rel(v) for v in tr[cta, 2, :16]]
    print("epilogue: [c0 wait-start, wait-done, done] [c1 ...] res wait-start, wait-done, skip wait-start, wait-done, end:")
    print("  ", e[0:3], e[4:7], e[8:13])
    if cta == 0:
        print("per layer (li = 0..5): producer past g2done", [rel(v) for v in tr[0, 0, 200:206]])
        print("                       producer flags seen  ", [rel(v) for v in tr[0, 0, 210:216]])
        print("                       mma TMEM buf0 free   ", [rel(v) for v in tr[0, 1, 210:216]])
        print("                       mma GEMM1 first MMA  ", [rel(v) for v in tr[0, 1, 230:236]])
        print("                       mma GEMM2 issued     ", [rel(v) for v in tr[0, 1, 220:226]])
        print("                       epi2 residual done   ", [rel(v) for v in tr[0, 2, 100:106]])
        print("                       epi2 skip done       ", [rel(v) for v in tr[0, 2, 110:116]])
