"""DNA transform microbenchmark (BASELINE configs[1]) -- also the command profiled with ncu."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

if __name__ == "__main__":
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    from action_conditioned_gans_b200 import _lib
    _lib.load()
    print(json.dumps(bench.dna_microbench(dev, bench.load_peaks()), indent=1))
