"""Zero-edit activation of `b200sat.install()` for the reference's entry points (train.py, run_gradio.py, user scripts).

`b200sat.pth` (next to this package; copy or symlink it into site-packages, or `site.addsitedir()` its directory) imports this
module at interpreter start.  Nothing happens unless SAT_B200=1.  With SAT_B200=1 a meta-path hook waits for the reference to
import `stable_audio_tools.models.factory` and wraps `create_model_from_config` (models/factory.py:3-24) so that the first
model construction calls `b200sat.install()` — by then every module install() patches is importable, and the patch is in place
before the model's first forward.  SAT_B200_STRICT=1 -> install(strict=True);  SAT_B200_FP32_MODELS=1 -> bf16 compute for
fp32 models called outside autocast (see install.py).
"""
import importlib.abc
import importlib.util
import os
import sys

_TARGET = "stable_audio_tools.models.factory"
_state = {"armed": False, "installed": False}


def _install_now():
    if _state["installed"]:
        return
    from .install import install
    install(strict=os.environ.get("SAT_B200_STRICT", "0") == "1")
    _state["installed"] = True


def _wrap_factory(mod):
    orig = getattr(mod, "create_model_from_config", None)
    if orig is None or getattr(orig, "__b200sat_wrapped__", False):
        return

    def create_model_from_config(*a, **k):
        _install_now()
        return orig(*a, **k)

    create_model_from_config.__b200sat_wrapped__ = True
    create_model_from_config.__wrapped__ = orig
    create_model_from_config.__doc__ = orig.__doc__
    mod.create_model_from_config = create_model_from_config
    # `from .factory import create_model_from_config` in stable_audio_tools/__init__.py and models/__init__.py binds the name at import
    # time: rebind the copies that already exist
    for name in ("stable_audio_tools", "stable_audio_tools.models"):
        pkg = sys.modules.get(name)
        if pkg is not None and getattr(pkg, "create_model_from_config", None) is orig:
            pkg.create_model_from_config = create_model_from_config


class _Loader(importlib.abc.Loader):
    def __init__(self, inner):
        self.inner = inner

    def create_module(self, spec):
        return self.inner.create_module(spec)

    def exec_module(self, module):
        self.inner.exec_module(module)
        _wrap_factory(module)


class _Finder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path, target=None):
        if fullname != _TARGET:
            return None
        for f in sys.meta_path:
            if f is self or not hasattr(f, "find_spec"):
                continue
            spec = f.find_spec(fullname, path, target)
            if spec is not None and spec.loader is not None:
                spec.loader = _Loader(spec.loader)
                return spec
        return None


def arm():
    """Idempotent; returns True when the hook is (already) in place."""
    if _state["armed"]:
        return True
    if _TARGET in sys.modules:
        _wrap_factory(sys.modules[_TARGET])
    else:
        sys.meta_path.insert(0, _Finder())
    _state["armed"] = True
    return True


if os.environ.get("SAT_B200", "0") == "1":
    arm()
