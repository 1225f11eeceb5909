"""A/B of an environment switch on single conv launches of the bench step (graph-timed, probe_r2's method).

    python scripts/first_layer_ab.py ACG_EPI_DIRECT g/conv1,d/conv1 [batch]

Prints forward / data-gradient / weight-gradient time of each layer with the switch unset and set, and checks that the
forward outputs of the two modes are bit-identical."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from action_conditioned_gans_b200 import engine as E  # noqa: E402
from action_conditioned_gans_b200 import kernels as K  # noqa: E402


# the PRODUCT library on purpose (the probe library's knock-out branches sit inside the small-K kernel's inner loops)
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
    var = sys.argv[1]
    names = sys.argv[2].split(",")
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    dev = torch.device("cuda:0")
    print("%-10s | %-22s | %-22s | %-22s" % ("layer", "fwd us (unset / set)", "dgrad us", "wgrad us"))
    for spec, runcls in ((E.g_dna_spec(6), "g"), (E.d_spec(), "d")):
        store = E.ParamStore(spec, dev)
        store.flat.normal_(0, 0.05)
        run = E.GeneratorRun(store, B, dev, True, 6) if runcls == "g" else E.DiscriminatorRun(store, B, dev)
        store.refresh_packs()
        for L in spec:
            if L.name not in names:
                continue
            st = run.layers[L.name]
            s = st.shape
            x = torch.randn(B, st.in_hw[0], st.in_hw[1], st.ld_in, device=dev).to(torch.bfloat16)
            out_dt = torch.float32 if L.name == "g/tconv4" else torch.bfloat16
            ldz = 36 if L.name == "g/tconv4" else st.ldz
            z = torch.empty(B, st.out_hw[0], st.out_hw[1], ldz, device=dev, dtype=out_dt)
            dz = torch.randn(B, st.out_hw[0], st.out_hw[1], st.ldz, device=dev).to(torch.bfloat16)
            dx = torch.empty(B, st.in_hw[0], st.in_hw[1], st.ld_in, device=dev, dtype=torch.bfloat16)
            dw = store.gviews[L.name + "/weights"]
            f, d, w = make_launches(L, s, st, x, z, dz, dx, dw, store.packs[L.name], ldz)
            res, outs = {}, {}
            for mode in ("unset", "set"):
                if mode == "set":
                    os.environ[var] = "1"
                else:
                    os.environ.pop(var, None)
                z.fill_(float("nan"))
                dx.fill_(float("nan"))
                f()
                d()
                torch.cuda.synchronize()
                outs[mode] = (z.clone(), dx.clone())
                res[mode] = [1e3 * graph_time(fn) for fn in (f, d, w)]
            os.environ.pop(var, None)
            same = all(torch.equal(a.view(torch.int16) if a.dtype == torch.bfloat16 else a,
                                   b.view(torch.int16) if b.dtype == torch.bfloat16 else b)
                       for a, b in zip(outs["unset"], outs["set"]))
            print("%-10s | %9.2f / %9.2f | %9.2f / %9.2f | %9.2f / %9.2f | outputs identical: %s" % (
                L.name, res["unset"][0], res["set"][0], res["unset"][1], res["set"][1], res["unset"][2], res["set"][2], same),
                flush=True)


if __name__ == "__main__":
    main()
