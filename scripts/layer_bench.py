"""Per-layer timing of the tcgen05 conv kernels at the bench batch (B=256): fprop / dgrad / wgrad of every layer of
the DNA generator and the discriminator, with achieved TFLOP/s (2*M*N*K of the real, unpadded problem)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from action_conditioned_gans_b200 import engine as E  # noqa: E402
from action_conditioned_gans_b200 import kernels as K  # noqa: E402


def timeit(fn, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    dev = torch.device("cuda:0")
    rows = []
    for spec, runcls in ((E.g_dna_spec(6), "g"), (E.d_spec(), "d")):
        store = E.ParamStore(spec, dev)
        store.flat.normal_(0, 0.05)
        run = E.GeneratorRun(store, B, dev, True, 6) if runcls == "g" else E.DiscriminatorRun(store, B, dev)
        store.refresh_packs()
        for L in spec:
            st = run.layers[L.name]
            s = st.shape
            flops = 2.0 * s.B * s.OH * s.OW * s.Cout * s.KH * s.KW * s.Cin
            if s.stride == 2 and L.kind == "deconv":
                pass
            x = torch.randn(B, st.in_hw[0], st.in_hw[1], st.ld_in, device=dev).to(torch.bfloat16)
            z = torch.empty(B, st.out_hw[0], st.out_hw[1], st.ldz, device=dev, dtype=torch.bfloat16)
            dz = torch.randn(B, st.out_hw[0], st.out_hw[1], st.ldz, device=dev).to(torch.bfloat16)
            dx = torch.empty(B, st.in_hw[0], st.in_hw[1], st.ld_in, device=dev, dtype=torch.bfloat16)
            dw = store.gviews[L.name + "/weights"]
            pk = store.packs[L.name]
            if L.kind == "conv":
                t_f = timeit(lambda: K.conv_fprop_tc(s, x, pk[3], z, st.ld_in, st.ldz, splitk=st.splitk_f))
                t_d = timeit(lambda: K.conv_dgrad_tc(s, dz, pk[6], dx, st.ldz, st.ld_in, splitk=st.splitk_b))
                t_w = timeit(lambda: K.conv_wgrad_tc(s, x, dz, dw, st.ld_in, st.ldz))
            else:
                t_f = timeit(lambda: K.conv_dgrad_tc(s, x, pk[3], z, st.ld_in, st.ldz, splitk=st.splitk_f))
                t_d = timeit(lambda: K.conv_fprop_tc(s, dz, pk[6], dx, st.ldz, st.ld_in, splitk=st.splitk_b))
                t_w = timeit(lambda: K.conv_wgrad_tc(s, dz, x, dw, st.ldz, st.ld_in))
            # stride-2 taps: a dgrad-form op only multiplies the taps that hit (1/4 of kh*kw*...) -> same FLOPs
            rows.append((L.name, L.kind, flops / 1e9, t_f, t_d, t_w))
    print("%-10s %-6s %8s | %8s %7s | %8s %7s | %8s %7s" % ("layer", "kind", "GFLOP", "fwd ms", "TF/s", "bwdD ms",
                                                         "TF/s", "wgrad ms", "TF/s"))
    tot = [0, 0, 0, 0]
    for n, k, gf, tf, td, tw in rows:
        print("%-10s %-6s %8.2f | %8.3f %7.1f | %8.3f %7.1f | %8.3f %7.1f" % (
            n, k, gf, tf, gf / tf, td, gf / td, tw, gf / tw))
        tot[0] += gf; tot[1] += tf; tot[2] += td; tot[3] += tw
    print("%-10s %-6s %8.2f | %8.3f %7.1f | %8.3f %7.1f | %8.3f %7.1f" % (
        "total", "", tot[0], tot[1], tot[0] / tot[1], tot[2], tot[0] / tot[2], tot[3], tot[0] / tot[3]))


if __name__ == "__main__":
    main()
