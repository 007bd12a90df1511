"""Swap the fused modules into the reference's model files.

The reference model shells bind the hot-path classes by name at import time
(``utae.py:8-9``, ``wtae.py:9-10``, ``timeunet.py:4-5``), so replacing those names is enough for
``get_model`` (learning/utils.py:50-136), ``train.py`` and ``src/webapp/prediction.py`` to construct
the fused modules unchanged.
"""
from __future__ import annotations

import functools
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
_saved_forward = {}
_saved_defaults = None


def _wrap_timeunet_forward(mod) -> bool:
    """``TimeUNet_v1.forward`` (timeunet.py:169-210) hands the attention back only when its own ``return_att`` is set
    (timeunet.py:205) but always asks the encoder for it (timeunet.py:178-180): 64 MB of fp32 per patch written for
    nothing.  The wrapper tells the encoder before every forward whether the model-level caller wants the attention."""
    cls = getattr(mod, "TimeUNet_v1", None)
    if cls is None or (cls, "forward") in _saved_forward:
        return False
    original = cls.forward

    @functools.wraps(original)
    def forward(self, input, batch_positions=None, return_att=False):
        enc = getattr(self, "temporal_encoder", None)
        if isinstance(enc, modules.LTAE):
            enc.return_attention = bool(return_att)
        return original(self, input, batch_positions=batch_positions, return_att=return_att)

    _saved_forward[(cls, "forward")] = original
    cls.forward = forward
    return True


def install(import_missing: bool = True, assume_zero_padded: bool = False) -> list:
    """Replace ``LTAE`` / ``LTAE4WTAE`` / ``TemporalAggregator`` in the reference's ``src.backbones``
    modules with the fused classes.  Returns the list of ``module.name`` bindings that were swapped.

    ``assume_zero_padded``: encoders constructed from now on skip padded frames (correct for the shipped models with
    ``pad_value=0``: ``smart_forward`` leaves exactly ``pad_value`` there, temp_shared_block.py:30-40).
    ``TimeUNet_v1.forward`` is wrapped so that its encoder only stores the attention when the model-level
    ``return_att`` asks for it."""
    global _saved_defaults
    # import every target FIRST (a model file imports src.backbones.tae itself: saving while importing would record
    # an already swapped class as "the original"), then save the originals, then swap
    mods = {}
    for mod_name in _TARGETS:
        mod = sys.modules.get(mod_name)
        if mod is None and import_missing:
            try:
                mod = importlib.import_module(mod_name)
            except Exception:  # the reference is not on sys.path, or an optional dependency is absent
                mod = None
        if mod is not None:
            mods[mod_name] = mod
    fused = {getattr(modules, n) for names in _TARGETS.values() for n in names}
    for mod_name, mod in mods.items():
        for name in _TARGETS[mod_name]:
            if hasattr(mod, name) and getattr(mod, name) not in fused:
                _saved.setdefault((mod_name, name), getattr(mod, name))
    swapped = []
    for mod_name, mod in mods.items():
        for name in _TARGETS[mod_name]:
            if not hasattr(mod, name):
                continue
            setattr(mod, name, getattr(modules, name))
            swapped.append(f"{mod_name}.{name}")
    if "src.backbones.timeunet" in mods and _wrap_timeunet_forward(mods["src.backbones.timeunet"]):
        swapped.append("src.backbones.timeunet.TimeUNet_v1.forward (return_att -> encoder)")
    if _saved_defaults is None:
        _saved_defaults = dict(modules._DEFAULTS)
    modules._DEFAULTS["assume_zero_padded"] = bool(assume_zero_padded)
    return swapped


def uninstall() -> None:
    """Restore the reference classes, ``TimeUNet_v1.forward`` and the construction defaults."""
    global _saved_defaults
    by_name = {name: obj for (_, name), obj in _saved.items()}
    for (mod_name, name), obj in list(_saved.items()):
        mod = sys.modules.get(mod_name)
        if mod is not None:
            setattr(mod, name, obj)
        del _saved[(mod_name, name)]
    for mod_name, names in _TARGETS.items():  # model files imported after install() picked up the fused names
        mod = sys.modules.get(mod_name)
        for name in names:
            if mod is not None and name in by_name and getattr(mod, name, None) is getattr(modules, name):
                setattr(mod, name, by_name[name])
    for (cls, attr), fn in list(_saved_forward.items()):
        setattr(cls, attr, fn)
        del _saved_forward[(cls, attr)]
    if _saved_defaults is not None:
        modules._DEFAULTS.update(_saved_defaults)
        _saved_defaults = None
