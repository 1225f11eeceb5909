"""Rerun spread of the generator forward (bf16 path): the batch-norm moments are summed with fp64 atomics in the conv
epilogues, so reruns agree to rounding only.  Prints max |frame_i - frame_0| with and without stream branches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from action_conditioned_gans_b200 import engine as E

dev = torch.device("cuda:0")
for B in (3, 16):
    for br in (False, True):
        rng = np.random.RandomState(7)
        store = E.ParamStore(E.g_dna_spec(6), dev, E.xavier_init(E.g_dna_spec(6), rng))
        run = E.GeneratorRun(store, B, dev, True, 6, branches=br)
        store.refresh_packs()
        img = torch.rand(B, 64, 64, 3, device=dev) * 2 - 1
        act = torch.randn(B, 10, device=dev)
        outs = []
        for i in range(8):
            g, s = run.forward(img, act)
            torch.cuda.synchronize()
            outs.append((g.clone(), s.clone()))
        d = [float((o[0] - outs[0][0]).abs().max()) for o in outs[1:]]
        ds = [float((o[1] - outs[0][1]).abs().max()) for o in outs[1:]]
        print("B=%d branches=%s frame spread %s state spread %s" % (B, br, ["%.1e" % x for x in d], ["%.1e" % x for x in ds]))
