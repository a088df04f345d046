"""Host-side logic that needs no GPU: the C-ABI library loads and exports every declared symbol, the
module mirror keeps the reference's names / buffers, failures are loud, sharding is consistent."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import HP, ROOT, golden
from oracle import diffnet_oracle as O


def test_library_exports_every_declared_symbol(lib_built):
    from diffsinger_b200 import _capi
    header = open(os.path.join(ROOT, "include", "dsx.h")).read()
    declared = set(re.findall(r"\b(dsx_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_capi.SYMBOLS), declared ^ set(_capi.SYMBOLS)
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _capi.lib.dsx_version() == 100


def test_header_enums_match_the_ctypes_constants(lib_built):
    """Precision / info / option codes of include/dsx.h and diffsinger_b200/_capi.py must agree."""
    from diffsinger_b200 import _capi
    header = open(os.path.join(ROOT, "include", "dsx.h")).read()
    enums = {k: int(v) for k, v in re.findall(r"\b(DSX_[A-Z0-9_]+)\s*=\s*(-?\d+)", header)}
    for name, value in enums.items():
        for prefix in ("PREC_", "INFO_", "OPT_"):
            if name.startswith("DSX_" + prefix):
                assert getattr(_capi, name[4:]) == value, name
    assert {v for k, v in enums.items() if k.startswith("DSX_PREC_")} == set(_capi.PRECISIONS.values())


def test_no_cpu_fallback_is_loud(lib_built):
    import diffsinger_b200 as dsx
    torch.manual_seed(0)
    net = dsx.DiffNet(80, hparams=HP).eval()
    with torch.no_grad(), pytest.raises(dsx.DsxError):
        net(torch.zeros(1, 1, 80, 16), torch.zeros(1, dtype=torch.long), torch.zeros(1, 256, 16))
    if not torch.cuda.is_available():
        from diffsinger_b200 import _capi
        h = ctypes.c_void_p()
        rc = _capi.lib.dsx_create(0, ctypes.byref(h))
        assert rc == -2 and b"no CPU fallback" in _capi.lib.dsx_last_error()


def test_missing_library_fails_import(lib_built):
    code = "import os; os.environ['DSX_LIB']='/nonexistent/libdsx.so'; import diffsinger_b200"
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


@pytest.mark.parametrize("cycle", [1, 4])
def test_module_mirror_names_shapes_and_init(lib_built, cycle):
    import diffsinger_b200 as dsx
    hp = dict(HP, dilation_cycle_length=cycle)
    torch.manual_seed(0)
    net = dsx.DiffNet(80, hparams=hp)
    ref_sd = O.build_state_dict(0, dilation_cycle_length=cycle, out_std=None)   # the reference ctor's order
    sd = net.state_dict()
    assert list(sd.keys()) == list(ref_sd.keys())
    for k in sd:
        assert torch.equal(sd[k], ref_sd[k]), k
    assert [int(l.dilated_conv.dilation[0]) for l in net.residual_layers] == [2 ** (i % cycle) for i in range(20)]
    assert dsx.DsxSampler(net)._cycle_len() == cycle
    # training branch (autograd) agrees with the oracle's forward
    torch.nn.init.normal_(net.output_projection.weight, std=0.02)
    g = golden(f"diffnet_fwd_cycle{cycle}.npz")
    spec = torch.from_numpy(g["spec"])[:1, :, :, :40]
    cond = torch.from_numpy(g["cond"])[:1, :, :40]
    t = torch.from_numpy(g["t"])[:1]
    out = net(spec, t, cond)
    with torch.no_grad():
        ref = O.diffnet_forward({k: v.detach() for k, v in net.state_dict().items()}, spec, t, cond, cycle)
    assert out.requires_grad and torch.allclose(out, ref, atol=1e-6)


def test_gaussian_diffusion_mirror_buffers(lib_built):
    import diffsinger_b200 as dsx
    g = golden("schedules.npz")
    torch.manual_seed(0)
    net = dsx.DiffNet(80, hparams=HP)
    smin, smax = [-5.0] * 80, [0.5] * 80
    gd = dsx.GaussianDiffusion(None, 80, net, timesteps=100, K_step=71, loss_type="l1", spec_min=smin, spec_max=smax,
                               fs2=torch.nn.Identity(), hparams=HP)
    for b in O.SCHEDULE_BUFFERS:
        assert np.array_equal(getattr(gd, b).numpy(), g[f"linear006_T100.{b}"]), b
    assert gd.spec_min.shape == (1, 1, 80) and gd.K_step == 71 and gd.num_timesteps == 100
    keys = set(gd.state_dict().keys())
    assert "denoise_fn.residual_layers.19.output_projection.bias" in keys and "posterior_mean_coef2" in keys
    with pytest.raises(dsx.DsxError):
        dsx.GaussianDiffusion(None, 80, net, spec_min=smin, spec_max=smax, hparams=HP)


def test_shard_bounds_cover_batch():
    from diffsinger_b200.parallel import shard_bounds
    for n in (1, 7, 8, 32, 33):
        for w in (1, 2, 4, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, tmp):
    import torch.distributed as dist
    from diffsinger_b200.parallel import all_gather_batch, shard_batch
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    full = torch.arange(5 * 3 * 4, dtype=torch.float32).reshape(5, 3, 4)
    local = shard_batch(full) * 1.0
    out = all_gather_batch(local, 5)
    ok = torch.equal(out, full)
    open(os.path.join(tmp, f"r{rank}"), "w").write("ok" if ok else "bad")
    dist.destroy_process_group()


def test_all_gather_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "r0").read() == "ok" and open(tmp_path / "r1").read() == "ok"
