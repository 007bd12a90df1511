"""Swap the fused modules into the reference's model files.

The reference model shells bind the hot-path classes by name at import time
(``utae.py:8-9``, ``wtae.py:9-10``, ``timeunet.py:4-5``), so replacing those names is enough for
``get_model`` (learning/utils.py:50-136), ``train.py`` and ``src/webapp/prediction.py`` to construct
the fused modules unchanged.
"""
from __future__ import annotations

import importlib
import sys

from . import modules

_TARGETS = {
    "src.backbones.tae": ("LTAE", "LTAE4WTAE"),
    "src.backbones.temporal_aggregator": ("TemporalAggregator",),
    "src.backbones.utae": ("LTAE", "TemporalAggregator"),
    "src.backbones.wtae": ("LTAE4WTAE", "TemporalAggregator"),
    "src.backbones.timeunet": ("LTAE", "TemporalAggregator"),
    "src.backbones.recunet": ("TemporalAggregator",),
}
_saved = {}


def install(import_missing: bool = True) -> list:
    """Replace ``LTAE`` / ``LTAE4WTAE`` / ``TemporalAggregator`` in the reference's ``src.backbones``
    modules with the fused classes.  Returns the list of ``module.name`` bindings that were swapped."""
    swapped = []
    for mod_name, names in _TARGETS.items():
        mod = sys.modules.get(mod_name)
        if mod is None and import_missing:
            try:
                mod = importlib.import_module(mod_name)
            except Exception:  # the reference is not on sys.path, or an optional dependency is absent
                mod = None
        if mod is None:
            continue
        for name in names:
            if not hasattr(mod, name):
                continue
            _saved.setdefault((mod_name, name), getattr(mod, name))
            setattr(mod, name, getattr(modules, name))
            swapped.append(f"{mod_name}.{name}")
    return swapped


def uninstall() -> None:
    """Restore the reference classes."""
    for (mod_name, name), obj in list(_saved.items()):
        mod = sys.modules.get(mod_name)
        if mod is not None:
            setattr(mod, name, obj)
        del _saved[(mod_name, name)]
