import logging as _logging
from collections import OrderedDict
from dataclasses import fields

import torch
from packaging import version


class _Logging:
    @staticmethod
    def get_logger(name):
        return _logging.getLogger(name)


logging = _Logging()


def deprecate(*args, **kwargs):
    pass


def maybe_allow_in_graph(cls):
    return cls


def is_torch_version(op, ver):
    import operator
    ops = {">": operator.gt, ">=": operator.ge, "==": operator.eq, "<": operator.lt, "<=": operator.le}
    return ops[op](version.parse(version.parse(torch.__version__).base_version), version.parse(ver))


class BaseOutput(OrderedDict):
    """dataclass-style output container with attribute and key access"""

    def __post_init__(self):
        for f in fields(self):
            v = getattr(self, f.name)
            if v is not None:
                self[f.name] = v

    def __getitem__(self, k):
        if isinstance(k, str):
            return dict(self.items())[k]
        return tuple(self.values())[k]

    def to_tuple(self):
        return tuple(self[k] for k in self.keys())
