"""Stage knock-out of the persistent small-K kernel (first layers), probe library: ACG_DBG_SKIP bits 1 (no gather),
4 (no MMAs), 16 (epilogue: TMEM loads only), 32 (no epilogue work).  Graph-timed forward of g/conv1 and d/conv1."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from action_conditioned_gans_b200 import engine as E  # noqa: E402
from probe_r2 import graph_time, make_launches  # noqa: E402  (loads the probe library)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
masks = [0, 1, 4, 32, 1 | 4, 1 | 32, 4 | 32, 1 | 4 | 32, 1 | 4 | 32 | 64, 1 | 4 | 32 | 128]
print("%-12s " % "skip mask" + " ".join("%7d" % m for m in masks))
for spec, runcls in ((E.g_dna_spec(6), "g"), (E.d_spec(), "d")):
    store = E.ParamStore(spec, dev)
    store.flat.normal_(0, 0.05)
    run = E.GeneratorRun(store, B, dev, True, 6) if runcls == "g" else E.DiscriminatorRun(store, B, dev)
    store.refresh_packs()
    L = spec[0]
    st = run.layers[L.name]
    x = torch.randn(B, st.in_hw[0], st.in_hw[1], st.ld_in, device=dev).to(torch.bfloat16)
    z = torch.empty(B, st.out_hw[0], st.out_hw[1], st.ldz, device=dev, dtype=torch.bfloat16)
    dz = torch.randn(B, st.out_hw[0], st.out_hw[1], st.ldz, device=dev).to(torch.bfloat16)
    dx = torch.empty(B, st.in_hw[0], st.in_hw[1], st.ld_in, device=dev, dtype=torch.bfloat16)
    f, d, w = make_launches(L, st.shape, st, x, z, dz, dx, store.gviews[L.name + "/weights"], store.packs[L.name], st.ldz)
    ts = []
    for m in masks:
        os.environ["ACG_DBG_SKIP"] = str(m)
        ts.append(graph_time(f) * 1e3)
    os.environ["ACG_DBG_SKIP"] = "0"
    print("%-12s " % (L.name + " fwd") + " ".join("%7.1f" % t for t in ts) + "   us", flush=True)
    if runcls == "g":       # calibration across boxes: a kernel that did not change (g/tconv3 forward, ~55 us)
        L3 = [l for l in spec if l.name == "g/tconv3"][0]
        st = run.layers[L3.name]
        x = torch.randn(B, st.in_hw[0], st.in_hw[1], st.ld_in, device=dev).to(torch.bfloat16)
        z = torch.empty(B, st.out_hw[0], st.out_hw[1], st.ldz, device=dev, dtype=torch.bfloat16)
        dz = torch.randn(B, st.out_hw[0], st.out_hw[1], st.ldz, device=dev).to(torch.bfloat16)
        dx = torch.empty(B, st.in_hw[0], st.in_hw[1], st.ld_in, device=dev, dtype=torch.bfloat16)
        f3, _, _ = make_launches(L3, st.shape, st, x, z, dz, dx, store.gviews[L3.name + "/weights"], store.packs[L3.name], st.ldz)
        print("calibration: g/tconv3 fwd %.1f us (55.0 on the reference box)" % (graph_time(f3) * 1e3), flush=True)
