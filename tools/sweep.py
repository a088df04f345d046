"""Throughput over the other BASELINE.json configs (device-resident, CUDA events, 2 warm-up + 2 timed loops)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import diffsinger_b200 as dsx
from oracle import diffnet_oracle as O

dev = torch.device("cuda", 0)
CFG = [  # name, B, T, schedule, K, pndm interval, dilation cycle
    ("cfg1 DiffSpeech LJ B=1 T=512 K=100 DDPM", 1, 512, (100, 0.06), 100, 0, 1),
    ("cfg2 DiffSpeech B=16 T=1024 K=100 DDPM", 16, 1024, (100, 0.06), 100, 0, 1),
    ("cfg3 PopCS B=8 T=2048 T=K=1000 DDPM (200 of 1000 steps timed)", 8, 2048, (1000, 0.02), 200, 0, 1),
    ("cfg4 OpenCpop PLMS B=32 T=1024 K=1000 interval 40", 32, 1024, (1000, 0.02), 1000, 40, 4),
    ("sweep B=64 T=256 K=100", 64, 256, (100, 0.06), 100, 0, 1),
    ("sweep B=4 T=4096 K=100", 4, 4096, (100, 0.06), 100, 0, 1),
]
out = []
for name, B, T, (Ts, mb), K, interval, cycle in CFG:
    hp = dict(bench.HP, dilation_cycle_length=cycle)
    torch.manual_seed(0)
    net = dsx.DiffNet(80, hparams=hp)
    torch.nn.init.normal_(net.output_projection.weight, std=0.02)
    net = net.to(dev).eval()
    S = O.make_schedule(O.linear_beta_schedule(Ts, mb))
    cond, xT = bench.make_inputs(B, T, 0)
    cond, xT = cond.to(dev).transpose(1, 2), xT.to(dev)
    for prec in ("fp16x3", "fp16x2", "fp16"):
        s = dsx.DsxSampler(net, prec, cycle)
        s.ensure_weights(dev)
        s.set_schedule(S)
        run = (lambda: s.sample_plms(xT, cond, K, interval)) if interval else (lambda: s.sample_ddpm(xT, cond, Ts, K, seed=1))
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        evals = (K // interval + (1 if K % interval else 0) + 1) if interval else K
        scale = (1000 / K) if "200 of 1000" in name else 1.0
        rec = {"config": name, "precision": prec, "ms_loop": ms * scale, "evals": int(evals * scale), "ms_per_eval": ms / evals,
               "frames_per_s": B * T / (ms * scale * 1e-3), "stack_launches_per_eval": -(-B // max(148 // ((T + 127) // 128), 1))}
        print(json.dumps(rec))
        out.append(rec)
        s.close()
json.dump(out, open("gpurun_out/sweep.json", "w"), indent=1)
