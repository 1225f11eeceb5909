import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from action_conditioned_gans_b200 import _lib
_lib.use_probe_library()
from action_conditioned_gans_b200 import engine as E, kernels as K
dev = torch.device("cuda:0"); B = 256
buf = (ctypes.c_ulonglong * 8)()
def phases(tag, fn):
    for _ in range(2): fn()
    torch.cuda.synchronize(); _lib.call("acg_debug_phase_times", buf)
    fn(); torch.cuda.synchronize(); _lib.call("acg_debug_phase_times", buf)
    n = max(buf[3], 1)
    print("%-22s CTAs %5d  kblocks/CTA %5.1f | setup %6.2f us  mainloop %6.2f us (%.3f us/kblock)  epilogue %6.2f us" % (
        tag, buf[3], buf[4] / n, buf[0] / n / 1e3, buf[1] / n / 1e3, buf[1] / max(buf[4], 1) / 1e3, buf[2] / n / 1e3),
        " | t0 epilogue loop %.2f us, mma waits %.2f us, warp1 barrier wait %.2f us" % (buf[5] / n / 1e3, buf[6] / n / 1e3, buf[7] / n / 1e3))
for spec, kind in ((E.g_dna_spec(6), "g"), (E.d_spec(), "d")):
    store = E.ParamStore(spec, dev); store.flat.normal_(0, 0.05)
    run = E.GeneratorRun(store, B, dev, True, 6) if kind == "g" else E.DiscriminatorRun(store, B, dev)
    store.refresh_packs()
    for L in spec:
        if L.name not in ("g/conv1", "g/conv2", "g/tconv4", "d/conv1", "d/conv2", "d/conv3", "d/conv5"): continue
        st = run.layers[L.name]; s = st.shape; pk = store.packs[L.name]
        x = torch.randn(B, st.in_hw[0], st.in_hw[1], st.ld_in, device=dev).to(torch.bfloat16)
        z = torch.empty(B, st.out_hw[0], st.out_hw[1], st.ldz, device=dev, dtype=torch.bfloat16)
        dz = torch.randn(B, st.out_hw[0], st.out_hw[1], st.ldz, device=dev).to(torch.bfloat16)
        dx = torch.empty(B, st.in_hw[0], st.in_hw[1], st.ld_in, device=dev, dtype=torch.bfloat16)
        if L.kind == "conv":
            phases(L.name + " fwd(CONV)", lambda: K.conv_fprop_tc(s, x, pk[3], z, st.ld_in, st.ldz))
            phases(L.name + " bwd(ADJ)", lambda: K.conv_dgrad_tc(s, dz, pk[6], dx, st.ldz, st.ld_in))
        else:
            phases(L.name + " fwd(ADJ)", lambda: K.conv_dgrad_tc(s, x, pk[3], z, st.ld_in, st.ldz))
            phases(L.name + " bwd(CONV)", lambda: K.conv_fprop_tc(s, dz, pk[6], dx, st.ldz, st.ld_in))
