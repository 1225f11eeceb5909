"""First layers (g/conv1, d/conv1) forward at B=256: small-K gather kernel vs the halo kernel's pixel-pair mode, graph-timed."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from action_conditioned_gans_b200 import _lib  # noqa: E402
from action_conditioned_gans_b200 import kernels as K  # noqa: E402

PROBE = len(sys.argv) > 2 and sys.argv[2] == "probe"
if PROBE:       # stage knock-out (ACG_DBG_SKIP: 1 halo TMA, 2 weight TMA, 4 MMA, 16 epilogue stores, 32 epilogue)
    _lib.use_probe_library()

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for name, Cin, Cout in (("g/conv1", 3, 32), ("d/conv1", 6, 64)):
    shape = K.conv_shape(B, 64, 64, Cin, Cout, 5, 2, "SAME")
    x = torch.zeros(B, 64, 64, 8, dtype=torch.bfloat16, device=dev)
    x[..., :Cin] = torch.randn(B, 64, 64, Cin, device=dev)
    w = torch.randn(5, 5, Cin, Cout, device=dev) / 10
    pf = torch.empty(K.pack_size(shape, 0, 8), dtype=torch.bfloat16, device=dev)
    pp = torch.empty(K.pack_size(shape, 2, 16), dtype=torch.bfloat16, device=dev)
    K.pack_weights(shape, w, 0, 8, pf)
    K.pack_weights(shape, w, 2, 16, pp)
    y = torch.empty(B, 32, 32, Cout, dtype=torch.bfloat16, device=dev)
    stats = torch.zeros(2 * Cout, dtype=torch.float64, device=dev)
    fix = K.stats_accumulators(Cout, dev)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    beta = torch.zeros(Cout, device=dev)
    m, r, sc, sh = (torch.zeros(Cout, device=dev) for _ in range(4))
    for pair, skip in ([(False, 0), (True, 0)] if not PROBE else [(True, m) for m in (0, 1, 2, 4, 16, 32, 36, 7)]):
        os.environ["ACG_DBG_SKIP"] = str(skip)

        def run():
            K.conv_fprop_tc(shape, x, pp if pair else pf, y, 8, Cout, stats=stats,
                            bn=(cnt, beta, m, r, sc, sh, B * 1024, 1e-3), stats_fix=fix, pair_x=pair)
        run()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                for _ in range(10):
                    run()
        for _ in range(2):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        print("%s  %-10s skip %2d  %.1f us" % (name, "pixel-pair" if pair else "small-K", skip, e0.elapsed_time(e1) / 50 * 1e3))
