#!/usr/bin/env python
"""Golden vector for ``smart_forward``: the reference ``TemporallySharedBlock`` wrapped
around a seeded strided convolution, applied to a padded batch.  Runs ONLY in the build container.

    python tests/golden/make_smart_forward_golden.py
"""
import json
import os
import sys

import numpy as np
import torch

REF = os.environ.get("CROP2SEG_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.path.insert(0, REF)
    from src.backbones.temp_shared_block import TemporallySharedBlock

    class Block(TemporallySharedBlock):
        def __init__(self):
            super().__init__(pad_value=0)
            self.conv = torch.nn.Conv2d(3, 5, kernel_size=4, stride=2, padding=1)

        def forward(self, x):
            return torch.relu(self.conv(x))

    torch.manual_seed(3)
    rng = np.random.RandomState(3)
    blk = Block().eval()
    x = rng.standard_normal((3, 6, 3, 8, 8)).astype(np.float32)
    x[0, 4:] = 0
    x[2, 2:] = 0
    with torch.no_grad():
        out = blk.smart_forward(torch.from_numpy(x))
    arrays = {"cfg": np.array(json.dumps({"pad_value": 0})), "x": x, "out::out": out.numpy()}
    for k, v in blk.state_dict().items():
        arrays["param::" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "smart_forward.npz"), **arrays)
    print("smart_forward.npz", os.path.getsize(os.path.join(HERE, "smart_forward.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
