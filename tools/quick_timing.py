"""Quick device-side timing of one DiffNet evaluation per precision / cta_group (CUDA events)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffsinger_b200 as dsx

dev = torch.device("cuda", 0)
hp = dict(hidden_size=256, residual_layers=20, residual_channels=256, dilation_cycle_length=1)
torch.manual_seed(0)
net = dsx.DiffNet(80, hparams=hp)
torch.nn.init.normal_(net.output_projection.weight, std=0.02)
net = net.to(dev).eval()
B, T = 16, 1024
g = torch.Generator().manual_seed(1)
cond = torch.randn(B, T, 256, generator=g).transpose(1, 2).to(dev)
x = torch.randn(B, 1, 80, T, generator=g).to(dev)
t = torch.full((B,), 50, dtype=torch.long, device=dev)
FLOP = 21184512 * B * T
for prec, group in (("fp16", 2), ("fp16x3", 2)):
    s = dsx.DsxSampler(net, prec, 1)
    s.ensure_weights(dev)
    if group:
        s.set_option(0, group)
    try:
        for _ in range(3):
            s.diffnet_forward(x, t, cond)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 3 if prec == "fp32" else 20
        e0.record()
        for _ in range(n):
            s.diffnet_forward(x, t, cond)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"{prec} group={group}: {ms:.3f} ms/eval  ->  {FLOP / ms / 1e9:.1f} TFLOP/s algorithmic (incl. cond pack + SIMT in/out proj)")
    except Exception as ex:
        print(f"{prec} group={group}: FAILED {ex}")
    s.close()
