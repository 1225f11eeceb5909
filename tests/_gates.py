"""Test support: the relu / lrelu branch every element took on the device, for oracle.torch_ref.GATES."""


def device_gates(*runs):
    """runs: NetRun objects in the order the oracle applies the networks (generator, D(gen), D(real)).
    Key = '<layer>#<k>' with k counting applications of the same layer name."""
    out, count = {}, {}
    for run in runs:
        for name, st in run.layers.items():
            if st.spec.act not in ("relu", "lrelu") or not hasattr(st, "z"):
                continue
            k = count.get(name, 0)
            count[name] = k + 1
            u = st.z[..., :st.spec.cout].float() * st.scale + st.shift
            out["%s#%d" % (name, k)] = (u > 0).cpu().numpy()
    return out
