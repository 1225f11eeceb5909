"""Weight-ring depth of the halo kernel (probe library, ACG_H2_NBS = cap): graph-timed forward / data gradient of the
halo layers at B=256 for caps 8 (default) .. 4."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from action_conditioned_gans_b200 import engine as E  # noqa: E402
from probe_r2 import graph_time, make_launches  # noqa: E402  (loads the probe library)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
caps = [8, 7, 6, 5, 4]
names = ("g/conv2", "g/tconv2", "g/tconv3", "g/tconv4", "d/conv2", "d/conv3")
print("%-14s " % "ring cap" + " ".join("%7d" % m for m in caps))
for spec, runcls in ((E.g_dna_spec(6), "g"), (E.d_spec(), "d")):
    store = E.ParamStore(spec, dev)
    store.flat.normal_(0, 0.05)
    run = E.GeneratorRun(store, B, dev, True, 6) if runcls == "g" else E.DiscriminatorRun(store, B, dev)
    store.refresh_packs()
    for L in spec:
        if L.name not in names:
            continue
        st = run.layers[L.name]
        x = torch.randn(B, st.in_hw[0], st.in_hw[1], st.ld_in, device=dev).to(torch.bfloat16)
        out_dt = torch.float32 if L.name == "g/tconv4" else torch.bfloat16
        ldz = 36 if L.name == "g/tconv4" else st.ldz
        z = torch.empty(B, st.out_hw[0], st.out_hw[1], ldz, device=dev, dtype=out_dt)
        dz = torch.randn(B, st.out_hw[0], st.out_hw[1], st.ldz, device=dev).to(torch.bfloat16)
        dx = torch.empty(B, st.in_hw[0], st.in_hw[1], st.ld_in, device=dev, dtype=torch.bfloat16)
        f, d, w = make_launches(L, st.shape, st, x, z, dz, dx, store.gviews[L.name + "/weights"], store.packs[L.name], ldz)
        for tag, fn in (("fwd", f), ("dgrad", d)):
            ts = []
            for m in caps:
                os.environ["ACG_H2_NBS"] = str(m)
                ts.append(graph_time(fn) * 1e3)
            os.environ.pop("ACG_H2_NBS", None)
            print("%-14s " % (L.name + " " + tag) + " ".join("%7.1f" % t for t in ts) + "   us", flush=True)
