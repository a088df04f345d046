"""Small fixed workload for ncu: the config-2 shape (B=16, T=1024), a few DDPM steps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import diffsinger_b200 as dsx

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16s"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = dict(bench.CONFIGS[sys.argv[3] if len(sys.argv) > 3 else "2"])
dev = torch.device("cuda", 0)
arm = bench.Arm(dsx, dict(cfg, K=K) if cfg["sampler"] == "ddpm" else cfg, prec, dev, 0)
mel = arm.step(0)
torch.cuda.synchronize()
print("ok", float(mel.abs().mean()), "stack launches", arm.s.info(9))
