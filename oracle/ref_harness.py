"""Import the UNMODIFIED reference from /root/reference (authoring container only).

TEST INFRASTRUCTURE.  Used by oracle/gen_golden.py and by the `not gpu` tests that
pin the oracle against the reference when /root/reference is present.  Nothing
that runs on the GPU box may import this module.

Stubs: `alias_free_torch`, `k_diffusion`, `einops_exts` are pip
dependencies that are not installed here; none of them is on the hot path.
"""
import importlib.util
import os
import sys
import types

REF_ROOT = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "stable_audio_tools"))


def _install_stubs():
    if "alias_free_torch" not in sys.modules:
        m = types.ModuleType("alias_free_torch")

        class Activation1d:  # only used when antialias_activation=True (autoencoders.py:43-44)
            def __init__(self, *a, **k):
                raise NotImplementedError("alias_free_torch stub")

        m.Activation1d = Activation1d
        sys.modules["alias_free_torch"] = m
    if "k_diffusion" not in sys.modules:
        sys.modules["k_diffusion"] = types.ModuleType("k_diffusion")
    if "einops_exts" not in sys.modules:
        m = types.ModuleType("einops_exts")
        m.rearrange_many = lambda *a, **k: None
        sys.modules["einops_exts"] = m


def load():
    """Returns a namespace with the reference modules on the hot path."""
    assert available(), "reference not mounted"
    _install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import stable_audio_tools.models.transformer as transformer
    transformer.flash_attn_func = None  # CPU: force the SDPA path (transformer.py:429-440)
    import stable_audio_tools.models.dit as dit
    import stable_audio_tools.models.blocks as blocks
    import stable_audio_tools.models.bottleneck as bottleneck
    import stable_audio_tools.models.autoencoders as autoencoders
    import stable_audio_tools.inference.sampling as sampling
    spec = importlib.util.spec_from_file_location(
        "ref_auraloss", os.path.join(REF_ROOT, "stable_audio_tools/training/losses/auraloss.py"))
    auraloss = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(auraloss)
    return types.SimpleNamespace(transformer=transformer, dit=dit, blocks=blocks, bottleneck=bottleneck,
                                 autoencoders=autoencoders, sampling=sampling, auraloss=auraloss)
