"""BaseOutput / logging plumbing (diffusers 0.10.2 restated)."""
import logging as _pylogging
from collections import OrderedDict
from dataclasses import fields


class BaseOutput(OrderedDict):
    def __post_init__(self):
        for f in fields(self):
            v = getattr(self, f.name)
            if v is not None:
                self[f.name] = v

    def __getitem__(self, k):
        if isinstance(k, str):
            return dict(self.items())[k]
        return tuple(self.values())[k]


class logging:  # noqa: N801  (mirrors `from diffusers.utils import logging`)
    @staticmethod
    def get_logger(name):
        return _pylogging.getLogger(name)
