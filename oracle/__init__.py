"""CPU oracle for the L-TAE + TemporalAggregator hot path.

TEST INFRASTRUCTURE ONLY.  This package is a plain-numpy restatement of the
reference algorithm (Many98/Crop2Seg, ``src/backbones/tae.py``,
``positional_encoding.py``, ``temporal_aggregator.py``).  It exists so that the
CUDA path can be checked against an independent implementation and so that
``bench.py`` can time a CPU baseline.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it; nothing under ``crop2seg_b200/`` does.

Pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md section 8c), so the oracle is pinned against outputs of the
reference itself, generated in the build container by importing
``/root/reference`` (``tests/golden/make_golden.py``) and committed as
``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks every fixture.
"""
from .ltae_oracle import (  # noqa: F401
    LtaeConfig,
    absolute_positional_encoding,
    group_norm_rows,
    lightweight_attention,
    ltae4wtae_forward,
    ltae_forward,
    sinusoid_positional_encoding,
)
from .aggregator_oracle import (  # noqa: F401
    avg_pool2d,
    bilinear_upsample,
    pad_mask_from_input,
    temporal_aggregator,
)
from .skipconv_oracle import aggregate_skip_conv, skip_conv  # noqa: F401
