"""TEST INFRASTRUCTURE ONLY -- CPU emulation of the GPU operand formats, to choose the
tensor-core precision mode before spending GPU time.

Emulates "round GEMM operands to fp16/bf16 (optionally hi+lo split), accumulate in fp32" for
the three contractions of each residual layer + the in/skip/out projections, and runs the
full K-step DDPM loop of tests/golden/ddpm_lj_K100.npz against the fp32 reference output.

    python -m oracle.precision_study          # bf16 / fp16, 1 or 3 passes
    python -m oracle.precision_study mixed    # per-operand schemes (weights / activations / conditioner / head)
"""
import math, os, sys
import numpy as np, torch, torch.nn.functional as F
from . import diffnet_oracle as O
from .gen_golden import rs_normal, OUT


def rnd(x, fmt, passes):
    """Operand as the tensor core sees it: sum of `passes`-term split in `fmt`."""
    dt = torch.float16 if fmt == "fp16" else torch.bfloat16
    hi = x.to(dt).float()
    if passes == 1:
        return hi
    lo = (x - hi).to(dt).float()
    return hi + lo            # 3-pass product drops only lo*lo (~2^-22 relative)


def forward(P, Pq, spec, t, cond, cycle, fmt, passes, approx_act):
    L = O.num_layers(P)
    q = lambda a: rnd(a, fmt, passes)
    x = F.relu(F.conv1d(q(spec[:, 0]), Pq["input_projection.weight"], P["input_projection.bias"]))
    e = O.step_embedding(P, t)
    condq = q(cond)
    skip = 0
    for i in range(L):
        p = f"residual_layers.{i}."
        d = F.linear(e, P[p + "diffusion_projection.weight"], P[p + "diffusion_projection.bias"]).unsqueeze(-1)
        dil = 2 ** (i % cycle)
        y = F.conv1d(q(x + d), Pq[p + "dilated_conv.weight"], P[p + "dilated_conv.bias"], padding=dil, dilation=dil) \
            + F.conv1d(condq, Pq[p + "conditioner_projection.weight"], P[p + "conditioner_projection.bias"])
        g, f = torch.chunk(y, 2, dim=1)
        z = torch.sigmoid(g) * torch.tanh(f)
        if approx_act:      # tanh.approx.f32: ~2^-11 relative error
            z = z * (1 + (torch.rand_like(z) - 0.5) * 2 ** -10)
        o = F.conv1d(q(z), Pq[p + "output_projection.weight"], P[p + "output_projection.bias"])
        r, s = torch.chunk(o, 2, dim=1)
        x = (x + r) / math.sqrt(2.0)
        skip = skip + s
    x = skip / math.sqrt(L)
    x = F.relu(F.conv1d(q(x), Pq["skip_projection.weight"], P["skip_projection.bias"]))
    x = F.conv1d(q(x), Pq["output_projection.weight"], P["output_projection.bias"])
    return x[:, None]


def main():
    torch.set_num_threads(8)
    g = np.load(os.path.join(OUT, "ddpm_lj_K100.npz"))
    cond, xT, x0 = (torch.from_numpy(g[k]) for k in ("cond", "xT", "x0"))
    K = 100
    noise = rs_normal(int(g["noise_seed"]), (K,) + tuple(xT.shape))
    P = O.build_state_dict(0)
    S = O.make_schedule(O.linear_beta_schedule(100, 0.06))
    for fmt, passes, approx in (("bf16", 1, False), ("fp16", 1, False), ("fp16", 1, True), ("bf16", 3, False), ("fp16", 3, False)):
        Pq = {k: (rnd(v, fmt, passes) if k.endswith("weight") and v.dim() == 3 else v) for k, v in P.items()}
        x = xT
        b = x.shape[0]
        single = None
        with torch.no_grad():
            for j, t in enumerate(reversed(range(K))):
                eps = forward(P, Pq, x, torch.full((b,), t), cond, 1, fmt, passes, approx)
                if j == 0:
                    single = (eps - O.diffnet_forward(P, x, torch.full((b,), t), cond, 1)).abs().max().item()
                xr = (S["sqrt_recip_alphas_cumprod"][t] * x - S["sqrt_recipm1_alphas_cumprod"][t] * eps).clamp(-1, 1)
                mean = S["posterior_mean_coef1"][t] * xr + S["posterior_mean_coef2"][t] * x
                x = mean + (0.0 if t == 0 else 1.0) * (0.5 * S["posterior_log_variance_clipped"][t]).exp() * noise[j]
        d = (x - x0).abs()
        print(f"{fmt} x{passes} approx_act={approx}: single-eval max|d eps| {single:.2e}; after K=100: max|dx| {d.max():.2e} "
              f"MAE {d.mean():.2e} frac>1e-3 {(d > 1e-3).float().mean():.2e}")


if __name__ == "__main__" and not (len(sys.argv) > 1 and sys.argv[1] == "mixed"):
    main()


# --------------------------------------------------------------------------------------------------------------
# Mixed schemes (python -m oracle.precision_study mixed): which roundings matter?  Each GEMM operand is either
# one fp16 value (1) or a hi+lo fp16 pair (2); the conditioner projection is exact (0: hoisted out of the loop and
# computed once), fp16 (1) or hi+lo (2).  Weight and conditioner roundings are the same at every step (coherent over
# the 100 steps); the roundings of the running activations are fresh noise each step.
# --------------------------------------------------------------------------------------------------------------
def _split(x, passes):
    hi = x.half().float()
    return hi if passes == 1 else hi + (x - hi).half().float()


def forward_mixed(P, Pw, Ph, spec, t, cond, cycle, act_p, cond_p, head_p):
    L = O.num_layers(P)
    x = F.relu(F.conv1d(_split(spec[:, 0], head_p), Ph["input_projection.weight"], P["input_projection.bias"]))
    e = O.step_embedding(P, t)
    skip = 0
    for i in range(L):
        p = f"residual_layers.{i}."
        d = F.linear(e, P[p + "diffusion_projection.weight"], P[p + "diffusion_projection.bias"]).unsqueeze(-1)
        dil = 2 ** (i % cycle)
        if cond_p == 0:
            cp = F.conv1d(cond, P[p + "conditioner_projection.weight"], P[p + "conditioner_projection.bias"])
        else:
            cp = F.conv1d(_split(cond, cond_p), Pw[p + "conditioner_projection.weight"], P[p + "conditioner_projection.bias"])
        y = F.conv1d(_split(x + d, act_p), Pw[p + "dilated_conv.weight"], P[p + "dilated_conv.bias"], padding=dil,
                     dilation=dil) + cp
        g, f = torch.chunk(y, 2, dim=1)
        z = torch.sigmoid(g) * torch.tanh(f)
        o = F.conv1d(_split(z, act_p), Pw[p + "output_projection.weight"], P[p + "output_projection.bias"])
        r, s = torch.chunk(o, 2, dim=1)
        x = (x + r) / math.sqrt(2.0)
        skip = skip + s
    x = skip / math.sqrt(L)
    x = F.relu(F.conv1d(_split(x, head_p), Ph["skip_projection.weight"], P["skip_projection.bias"]))
    x = F.conv1d(_split(x, head_p), Ph["output_projection.weight"], P["output_projection.bias"])
    return x[:, None]


def main_mixed():
    torch.set_num_threads(16)
    g = np.load(os.path.join(OUT, "ddpm_lj_K100.npz"))
    cond, xT, x0 = (torch.from_numpy(g[k]) for k in ("cond", "xT", "x0"))
    K = 100
    noise = rs_normal(int(g["noise_seed"]), (K,) + tuple(xT.shape))
    P = O.build_state_dict(0)
    S = O.make_schedule(O.linear_beta_schedule(100, 0.06))
    # name, (W1 passes, W2 passes), activation passes, conditioner mode, head activation passes, head weight passes
    schemes = (
        ("fp16 before the hoist (all single pass)", (1, 1), 1, 1, 1, 1),
        ("fp16 mode now (conditioner exact)", (1, 1), 1, 0, 1, 1),
        ("fp16 + 3-pass head", (1, 1), 1, 0, 2, 2),
        ("fp16x2 before the hoist (W hi/lo, conditioner hi/lo)", (2, 2), 1, 2, 2, 2),
        ("fp16x2 now (W hi/lo, conditioner exact, 3-pass head)", (2, 2), 1, 0, 2, 2),
        ("W hi/lo, conditioner fp16", (2, 2), 1, 1, 2, 2),
        ("fp16x2 with single-pass head activations", (2, 2), 1, 0, 1, 2),
        ("W1 single, W2 hi/lo, conditioner exact, 3-pass head", (1, 2), 1, 0, 2, 2),
        ("W1 hi/lo, W2 single, conditioner exact, 3-pass head", (2, 1), 1, 0, 2, 2),
    )
    for name, wp, act_p, cond_p, head_p, hw_p in schemes:
        Pw = {k: (_split(v, wp[1] if "output_projection" in k else wp[0]) if k.endswith("weight") and v.dim() == 3 else v)
              for k, v in P.items()}
        Ph = {k: (_split(v, hw_p) if k.endswith("weight") and v.dim() == 3 else v) for k, v in P.items()}
        x, b = xT.clone(), xT.shape[0]
        with torch.no_grad():
            for j, t in enumerate(reversed(range(K))):
                eps = forward_mixed(P, Pw, Ph, x, torch.full((b,), t), cond, 1, act_p, cond_p, head_p)
                xr = (S["sqrt_recip_alphas_cumprod"][t] * x - S["sqrt_recipm1_alphas_cumprod"][t] * eps).clamp(-1, 1)
                mean = S["posterior_mean_coef1"][t] * xr + S["posterior_mean_coef2"][t] * x
                x = mean + (0.0 if t == 0 else 1.0) * (0.5 * S["posterior_log_variance_clipped"][t]).exp() * noise[j]
        d = (x - x0).abs()
        print(f"{name}: max {d.max():.2e} MAE {d.mean():.2e} frac>1e-3 {(d > 1e-3).float().mean():.2e}", flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "mixed":
    main_mixed()
