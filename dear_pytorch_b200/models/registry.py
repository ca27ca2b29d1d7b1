"""Name -> constructor registry used by the benchmark drivers (``--model``)."""
from __future__ import annotations

import importlib

_REGISTRY = {
    "resnet18": ("resnet", "resnet18"), "resnet34": ("resnet", "resnet34"), "resnet50": ("resnet", "resnet50"),
    "resnet101": ("resnet", "resnet101"), "resnet152": ("resnet", "resnet152"),
    "vgg11": ("vgg", "vgg11"), "vgg16": ("vgg", "vgg16"), "vgg19": ("vgg", "vgg19"),
    "densenet121": ("densenet", "densenet121"), "densenet169": ("densenet", "densenet169"),
    "densenet201": ("densenet", "densenet201"),
    "inceptionv4": ("inceptionv4", "inceptionv4"),
    "mnist": ("mnist", "Net"),
    "bert_base": ("bert", "bert_base"), "bert": ("bert", "bert_large"), "bert_large": ("bert", "bert_large"),
}


def available():
    return sorted(_REGISTRY)


def create(name: str, **kwargs):
    if name not in _REGISTRY:
        raise KeyError("unknown model %r; available: %s" % (name, ", ".join(available())))
    mod, fn = _REGISTRY[name]
    return getattr(importlib.import_module("dear_pytorch_b200.models." + mod), fn)(**kwargs)


def input_size(name: str) -> int:
    return 299 if name in ("inceptionv4", "inception_v3") else 224
