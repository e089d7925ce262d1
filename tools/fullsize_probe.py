#!/usr/bin/env python
"""Full-size BASELINE configs[2] / configs[3] on one GPU, per-pass times (the `configs` block of bench.py, on its own).

usage: fullsize_probe.py [3|4] [passes]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import rbffd_b200 as rb  # noqa: E402


def main():
    which = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    passes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    ctx = rb.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    peak = ctx.measure_fp64_peak()
    g = 3162 if which == 3 else 271
    out = bench.full_size_config(rb, ctx, torch, dev, bench.CONFIGS[which], g, peak, 6548.2, passes=passes)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
