"""What the fused batch-norm moments cost per launch: forward of every BN layer at B=256, graph-timed on the PRODUCT
library, (a) plain, (b) with moments + in-kernel finalize as the engine launches it (integer-limb accumulators),
(c) the same with fp64 atomics (no limb accumulators)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from action_conditioned_gans_b200 import engine as E  # noqa: E402
from first_layer_ab import graph_time  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
print("%-10s | %9s %9s %9s %9s | %s" % ("layer", "plain us", "limbs us", "fp64 us", "no ticket", "moments cost (limbs)"))
tot = [0.0, 0.0, 0.0, 0.0]
for spec, runcls in ((E.g_dna_spec(6), "g"), (E.d_spec(), "d")):
    store = E.ParamStore(spec, dev)
    store.flat.normal_(0, 0.05)
    run = E.GeneratorRun(store, B, dev, True, 6) if runcls == "g" else E.DiscriminatorRun(store, B, dev)
    store.refresh_packs()
    for L in spec:
        if not L.bn:
            continue
        st = run.layers[L.name]
        x = torch.randn(B, st.in_hw[0], st.in_hw[1], st.ld_in, device=dev).to(torch.bfloat16)
        beta = store.views[L.name + "/BatchNorm/beta"]
        bn = (st.counter, beta, st.mean, st.rstd, st.scale, st.shift, st.rows, E.BN_EPS)
        fix = st.stats_fix

        def plain():
            run._conv_fwd(st, x, st.z, st.ldz)

        def limbs():
            st.stats_fix = fix
            run._conv_fwd(st, x, st.z, st.ldz, stats=st.stats, bn=bn)

        def fp64():
            st.stats_fix = None
            run._conv_fwd(st, x, st.z, st.ldz, stats=st.stats, bn=bn)

        def atomics_only():          # fp64 atomics, no ticket / last-CTA finalize
            st.stats_fix = None
            run._conv_fwd(st, x, st.z, st.ldz, stats=st.stats, bn=None)

        ts = [1e3 * graph_time(f) for f in (plain, limbs, fp64, atomics_only)]
        st.stats_fix = fix
        for i in range(4):
            tot[i] += ts[i]
        print("%-10s | %9.2f %9.2f %9.2f %9.2f | %+6.2f" % (L.name, ts[0], ts[1], ts[2], ts[3], ts[1] - ts[0]), flush=True)
print("%-10s | %9.2f %9.2f %9.2f %9.2f | %+6.2f" % ("sum", tot[0], tot[1], tot[2], tot[3], tot[1] - tot[0]))
