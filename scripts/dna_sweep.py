"""DNA transform: band height (ACG_DNA_ROWS = 2 | 4) x ring depth (ACG_DNA_STAGES) x grid (ACG_DNA_GRID; the last two
exist in the probe library only) at several batch sizes, graph-timed like bench.py's microbench (24 launches per
graph, rotating buffer sets larger than L2)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def measure(dev, b, k, nset):
    from action_conditioned_gans_b200 import kernels as Kn
    sets = []
    for i in range(nset):
        g = torch.Generator(device=dev).manual_seed(i)
        sets.append((torch.randn(b, 64, 64, k * k, device=dev, generator=g),
                     torch.rand(b, 64, 64, 3, device=dev, generator=g) * 2 - 1,
                     torch.randn(b, 64, 64, 3, device=dev, generator=g),
                     torch.empty(b, 64, 64, 3, device=dev), torch.empty(b, 64, 64, k * k, device=dev),
                     torch.empty(b, 64, 64, (k * k + 15) // 16 * 16, device=dev, dtype=torch.bfloat16)))
    it = [0]

    def fwd():
        lg, im, dy, o, dl, dp = sets[it[0] % nset]
        it[0] += 1
        Kn.dna_fwd(lg, im, o, k)

    def bwd():
        lg, im, dy, o, dl, dp = sets[it[0] % nset]
        it[0] += 1
        Kn.dna_bwd(lg, im, dy, dl, k)

    def bwd_pad():
        lg, im, dy, o, dl, dp = sets[it[0] % nset]
        it[0] += 1
        Kn.dna_bwd(lg, im, dy, dp, k)

    return [1e3 * bench.time_kernel(f, 24) for f in (fwd, bwd, bwd_pad)]


if __name__ == "__main__":
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    from action_conditioned_gans_b200 import _lib
    _lib.use_probe_library()
    _lib.load()
    peak = bench.load_peaks()["hbm_gbs"]
    full = len(sys.argv) > 1 and sys.argv[1] == "full"
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"      # default ring depth and grid only
    print("%-28s %4s %2s | %8s %6s | %8s %6s | %8s" % ("variant", "B", "K", "fwd us", "frac", "bwd us", "frac", "bwd16 us"))
    for (b, k, nset) in ((64, 5, 12), (256, 6, 4)) + (((64, 6, 10), (16, 5, 40), (128, 6, 6)) if full else ()):
        combos = [(r, st, g) for r in ("2", "4") for st in (None, "2", "3", "4", "5") for g in (None,)]
        if quick:
            combos = [("2", None, None), ("4", None, None)]
        elif b == 64 and k == 5:
            combos += [(r, st, str(148 * m)) for r in ("2",) for st in ("2", "3") for m in (3, 4, 5, 6, 8)]
            combos += [("4", "3", str(148 * m)) for m in (1, 2)] + [("4", "2", str(148 * m)) for m in (2, 3)]
        for rows, st, grid in combos:
            if rows == "4" and st == "5":
                continue
            os.environ["ACG_DNA_ROWS"] = rows
            for key, v in (("ACG_DNA_STAGES", st), ("ACG_DNA_GRID", grid)):
                if v:
                    os.environ[key] = v
                else:
                    os.environ.pop(key, None)
            try:
                tf, tb, tp = measure(dev, b, k, nset)
            except RuntimeError as e:
                print("rows=%s stages=%s grid=%s: %s" % (rows, st, grid, str(e)[:80]))
                continue
            bf = b * 4096 * (k * k + 6) * 4
            bb = b * 4096 * (2 * k * k + 6) * 4
            print("rows=%s stages=%-4s grid=%-5s %4d %2d | %8.2f %6.3f | %8.2f %6.3f | %8.2f" % (
                rows, st or "def", grid or "def", b, k, tf, bf / tf / 1e3 / peak, tb, bb / tb / 1e3 / peak, tp), flush=True)
