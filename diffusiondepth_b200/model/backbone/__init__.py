"""Backbone registry: `get(args)` resolves `args.backbone_module` / `args.backbone_name` to a zero-argument
factory, as reference src/model/backbone/__init__.py:5-11 does."""
from importlib import import_module


def get(args):
    module = import_module(f"{__name__}.{args.backbone_module.lower()}")
    return getattr(module, args.backbone_name)
