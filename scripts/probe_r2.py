"""Round-2 probes (probe library: -DACG_PROBES).

1. Graph-timed per-layer table: every conv launch of the bench step (B=256) timed as 10 back-to-back launches inside a
   CUDA graph (no host launch cost, which dominates the eager numbers of the 10-us kernels).
2. Pipeline-stage knock-out of the persistent halo kernel (g/tconv3 fwd, g/tconv4 fwd, d/conv2 dgrad): the same launch
   with the halo TMA / weight TMA / MMAs / epilogue switched off tells which stage bounds the tile time.

    python scripts/probe_r2.py [batch]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from action_conditioned_gans_b200 import _lib  # noqa: E402

_lib.use_probe_library()
from action_conditioned_gans_b200 import engine as E  # noqa: E402
from action_conditioned_gans_b200 import kernels as K  # noqa: E402


def graph_time(fn, iters=10, reps=3):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / (iters * reps)


def make_launches(L, s, st, x, z, dz, dx, dw, pk, ldz):
    if L.kind == "conv":
        f = lambda: K.conv_fprop_tc(s, x, pk[3], z, st.ld_in, ldz, splitk=st.splitk_f)
        d = lambda: K.conv_dgrad_tc(s, dz, pk[6], dx, st.ldz, st.ld_in, splitk=st.splitk_b)
        w = lambda: K.conv_wgrad_tc(s, x, dz, dw, st.ld_in, st.ldz)
    else:
        f = lambda: K.conv_dgrad_tc(s, x, pk[3], z, st.ld_in, ldz, splitk=st.splitk_f)
        d = lambda: K.conv_fprop_tc(s, dz, pk[6], dx, st.ldz, st.ld_in, splitk=st.splitk_b)
        w = lambda: K.conv_wgrad_tc(s, dz, x, dw, st.ldz, st.ld_in)
    return f, d, w


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    dev = torch.device("cuda:0")
    rows = []
    launches = {}
    for spec, runcls in ((E.g_dna_spec(6), "g"), (E.d_spec(), "d")):
        store = E.ParamStore(spec, dev)
        store.flat.normal_(0, 0.05)
        run = E.GeneratorRun(store, B, dev, True, 6) if runcls == "g" else E.DiscriminatorRun(store, B, dev)
        store.refresh_packs()
        for L in spec:
            st = run.layers[L.name]
            s = st.shape
            flops = 2.0 * s.B * s.OH * s.OW * s.Cout * s.KH * s.KW * s.Cin
            x = torch.randn(B, st.in_hw[0], st.in_hw[1], st.ld_in, device=dev).to(torch.bfloat16)
            out_dt = torch.float32 if L.name == "g/tconv4" else torch.bfloat16
            ldz = 36 if L.name == "g/tconv4" else st.ldz
            z = torch.empty(B, st.out_hw[0], st.out_hw[1], ldz, device=dev, dtype=out_dt)
            dz = torch.randn(B, st.out_hw[0], st.out_hw[1], st.ldz, device=dev).to(torch.bfloat16)
            dx = torch.empty(B, st.in_hw[0], st.in_hw[1], st.ld_in, device=dev, dtype=torch.bfloat16)
            dw = store.gviews[L.name + "/weights"]
            pk = store.packs[L.name]
            f, d, w = make_launches(L, s, st, x, z, dz, dx, dw, pk, ldz)
            t_f, t_d, t_w = graph_time(f), graph_time(d), graph_time(w)
            rows.append((L.name, L.kind, flops / 1e9, t_f, t_d, t_w))
            launches[L.name] = (f, d, w)
        if runcls == "g":
            keep_g = (store, run)
        else:
            keep_d = (store, run)
    print("GRAPH-TIMED per-layer table, B=%d (10 launches per graph, 3 replays)" % B)
    print("%-10s %-6s %8s | %8s %7s | %8s %7s | %8s %7s" % ("layer", "kind", "GFLOP", "fwd ms", "TF/s", "bwdD ms",
                                                         "TF/s", "wgrad ms", "TF/s"))
    tot = [0, 0, 0, 0]
    for n, k, gf, tf, td, tw in rows:
        print("%-10s %-6s %8.2f | %8.4f %7.1f | %8.4f %7.1f | %8.4f %7.1f" % (n, k, gf, tf, gf / tf, td, gf / td, tw,
                                                                              gf / tw))
        tot[0] += gf; tot[1] += tf; tot[2] += td; tot[3] += tw
    print("%-10s %-6s %8.2f | %8.4f %7.1f | %8.4f %7.1f | %8.4f %7.1f" % (
        "total", "", tot[0], tot[1], tot[0] / tot[1], tot[2], tot[0] / tot[2], tot[3], tot[0] / tot[3]))

    # ---- knock-out of pipeline stages of the persistent halo kernel --------------------------------------------
    print("\nPERSISTENT HALO KERNEL, stage knock-out (ACG_DBG_SKIP bits: 1 halo TMA, 2 weight TMA, 4 MMA, 16 epilogue "
          "stores, 32 epilogue)")
    cases = [("g/tconv3 fwd", launches["g/tconv3"][0]), ("g/tconv4 fwd", launches["g/tconv4"][0]),
             ("d/conv2 dgrad", launches["d/conv2"][1]), ("d/conv1 dgrad", launches["d/conv1"][1]),
             ("g/tconv4 dgrad", launches["g/tconv4"][1]), ("g/tconv3 dgrad", launches["g/tconv3"][1]),
             ("d/conv2 fwd", launches["d/conv2"][0]), ("g/conv2 fwd", launches["g/conv2"][0])]
    masks = [0, 1, 2, 3, 4, 16, 32, 1 | 2 | 32, 4 | 32, 1 | 2 | 4, 1 | 4 | 32, 2 | 4 | 32]
    print("%-14s " % "skip mask" + " ".join("%7d" % m for m in masks))
    for name, fn in cases:
        ts = []
        for m in masks:
            os.environ["ACG_DBG_SKIP"] = str(m)
            ts.append(graph_time(fn) * 1e3)
        os.environ["ACG_DBG_SKIP"] = "0"
        print("%-14s " % name + " ".join("%7.1f" % t for t in ts) + "   us")


if __name__ == "__main__":
    main()
