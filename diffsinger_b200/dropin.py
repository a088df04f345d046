"""Drop-in installation inside the reference tree.

    import diffsinger_b200.dropin as dropin; dropin.install()      # after utils.hparams.set_hparams(...)

replaces, without editing any reference file,
  * ``usr.diff.shallow_diffusion_tts.GaussianDiffusion`` (and the names bound from it in
    ``usr.diffspeech_task`` / ``usr.diffsinger_task`` / ``inference.svs.*`` if already imported) by a
    subclass whose ``forward(infer=True)`` runs the sm_100a sampler, and
  * the ``'wavenet'`` entry of every ``DIFF_DECODERS`` registry by ``diffsinger_b200.DiffNet``.
Construction arguments, parameter / buffer names, ``p_losses`` and the returned ``ret`` dict are the
reference's own, so ``usr/diffspeech_task.py:23-38`` and ``usr/diffsinger_task.py:40-64`` run unchanged.
"""
import importlib
import sys

import torch

from .modules import DiffNet, DsxInferMixin

_installed = {}


def make_subclass(ref_cls):
    """Subclass of the reference's GaussianDiffusion with the infer branch routed to dsx."""

    class DsxGaussianDiffusion(DsxInferMixin, ref_cls):
        def forward(self, txt_tokens, mel2ph=None, spk_embed=None, ref_mels=None, f0=None, uv=None, energy=None,
                    infer=False, **kwargs):
            if not infer:
                return ref_cls.forward(self, txt_tokens, mel2ph, spk_embed, ref_mels, f0, uv, energy, infer, **kwargs)
            dsx_kw = {k: kwargs.pop(k) for k in list(kwargs) if k.startswith('dsx_')}
            # same call as usr/diff/shallow_diffusion_tts.py:236-238
            ret = self.fs2(txt_tokens, mel2ph, spk_embed, ref_mels, f0, uv, energy, skip_decoder=False, infer=True,
                           **kwargs)
            cond = ret['decoder_inp'].transpose(1, 2)
            with torch.no_grad():
                return self.dsx_infer(ret, cond, mel2ph, step_noise=dsx_kw.get('dsx_step_noise'),
                                      start_noise=dsx_kw.get('dsx_start_noise'), seed=dsx_kw.get('dsx_seed'))

    DsxGaussianDiffusion.__name__ = ref_cls.__name__
    DsxGaussianDiffusion.__qualname__ = ref_cls.__qualname__
    return DsxGaussianDiffusion


def install():
    sdt = importlib.import_module("usr.diff.shallow_diffusion_tts")
    ref_cls = _installed.get("ref_cls") or sdt.GaussianDiffusion
    new_cls = _installed.get("new_cls") or make_subclass(ref_cls)
    _installed.update(ref_cls=ref_cls, new_cls=new_cls)
    sdt.GaussianDiffusion = new_cls
    net_mod = importlib.import_module("usr.diff.net")
    _installed.setdefault("ref_net", net_mod.DiffNet)
    net_mod.DiffNet = DiffNet
    wavenet = lambda hp: DiffNet(hp['audio_num_mel_bins'])
    for name, mod in list(sys.modules.items()):
        if mod is None or not (name.startswith("usr.") or name.startswith("inference.") or name.startswith("tasks.")):
            continue
        if getattr(mod, "GaussianDiffusion", None) is ref_cls:
            mod.GaussianDiffusion = new_cls
        reg = getattr(mod, "DIFF_DECODERS", None)
        if isinstance(reg, dict) and "wavenet" in reg:
            reg["wavenet"] = wavenet
    return new_cls


def uninstall():
    if not _installed:
        return
    sdt = importlib.import_module("usr.diff.shallow_diffusion_tts")
    sdt.GaussianDiffusion = _installed["ref_cls"]
    importlib.import_module("usr.diff.net").DiffNet = _installed["ref_net"]
    for name, mod in list(sys.modules.items()):
        if mod is not None and getattr(mod, "GaussianDiffusion", None) is _installed["new_cls"]:
            mod.GaussianDiffusion = _installed["ref_cls"]
