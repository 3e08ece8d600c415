"""ConfigMixin / register_to_config plumbing (diffusers 0.10.2 restated): records the
__init__ kwargs into a `.config` namespace.  No arithmetic."""
import functools
import inspect
from types import SimpleNamespace


class _Config(SimpleNamespace):
    def __getitem__(self, k):
        return getattr(self, k)


class ConfigMixin:
    config_name = "config.json"

    @property
    def config(self):
        return self._internal_dict


def register_to_config(init):
    @functools.wraps(init)
    def inner(self, *args, **kwargs):
        sig = inspect.signature(init)
        bound = sig.bind(self, *args, **kwargs)
        bound.apply_defaults()
        cfg = {k: v for k, v in bound.arguments.items() if k != "self"}
        self._internal_dict = _Config(**cfg)
        init(self, *args, **kwargs)

    return inner
