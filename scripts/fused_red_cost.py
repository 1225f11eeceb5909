"""Data-gradient launches with and without the fused batch-norm backward reduction of the consumer layer
(acg_tc_args.red_*), graph-timed on the product library, plus the separate acg_bn_act_bwd_reduce pass it would replace."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from action_conditioned_gans_b200 import engine as E  # noqa: E402
from action_conditioned_gans_b200 import kernels as K  # noqa: E402
from first_layer_ab import graph_time  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
PAIRS = {"g": [("g/tconv4", "g/tconv3"), ("g/tconv3", "g/tconv2"), ("g/conv3", "g/conv2"), ("g/conv2", "g/conv1")],
         "d": [("d/conv3", "d/conv2"), ("d/conv2", "d/conv1")]}
print("%-22s | %9s %9s | %9s | kind" % ("producer -> consumer", "plain us", "fused us", "reduce us"))
for spec, runcls in ((E.g_dna_spec(6), "g"), (E.d_spec(), "d")):
    store = E.ParamStore(spec, dev)
    store.flat.normal_(0, 0.05)
    run = E.GeneratorRun(store, B, dev, True, 6) if runcls == "g" else E.DiscriminatorRun(store, B, dev)
    store.refresh_packs()
    for prod, cons in PAIRS[runcls]:
        st, sc = run.layers[prod], run.layers[cons]
        L, Lc = st.spec, sc.spec
        st.dz.copy_(torch.randn_like(st.dz.float()).to(st.dz.dtype))
        sc.z.copy_(torch.randn_like(sc.z.float()).to(sc.z.dtype))
        sc.rstd.fill_(1.0)
        pk = store.packs[prod]
        fn = K.conv_dgrad_tc if L.kind == "conv" else K.conv_fprop_tc
        nl = getattr(st, "dx_channels", 0)
        which = 1 if L.kind == "conv" else 0
        kind = K.kernel_kind(st.shape, which, st.ldz, nl)
        red = (sc.red, sc.z, sc.ldz, Lc.cout, Lc.act, sc.mean, sc.rstd, sc.shift)

        def plain():
            fn(st.shape, st.dz, pk[6], st.dx, st.ldz, st.ld_in, splitk=st.splitk_b, n_limit=nl)

        def fused():
            fn(st.shape, st.dz, pk[6], st.dx, st.ldz, st.ld_in, red=red, splitk=st.splitk_b, n_limit=nl)

        def reduce():
            K.bn_act_bwd_reduce(st.dx, None, st.dx.shape[3], sc.z, sc.ldz, sc.rows, Lc.cout, 1, sc.mean, sc.rstd, sc.shift,
                                Lc.act, sc.red)

        ts = [1e3 * graph_time(f) for f in (plain, fused, reduce)]
        print("%-22s | %9.2f %9.2f | %9.2f | %d" % (prod + " -> " + cons, ts[0], ts[1], ts[2], kind), flush=True)
