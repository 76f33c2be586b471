import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import stress_overlap as so
from archnemesis_dist_b200 import ops
from oracle import oracle as orc
rng = np.random.default_rng(7)
for case in range(16):
    k, dkdT, amount, dg, desc = so.make_case(rng, case)
print(desc)
orc.set_sort_mode(orc.NUMBA_ORDER)
otab = ops.OverlapTables(dg)
print("seq flag", otab.seq)
rt = orc.k_overlap(dg, k, amount)
tau = ops.koverlap(ops.to_dev(k), ops.to_dev(amount), otab).cpu().numpy()
bad = np.argwhere(np.abs(tau - rt) > 1e-12 * np.maximum(np.abs(rt), 1e-300))
print("bad entries", len(bad), "of", tau.size)
cells = sorted(set((int(b[0]), int(b[2])) for b in bad))
print("bad cells (iw,l):", cells[:10])
iw, l = cells[0]
np.set_printoptions(precision=4, linewidth=200)
print("k gas0", k[iw, :, l, 0]); print("k gas1", k[iw, :, l, 1]); print("amount", amount[:, l])
print("ref tau", rt[iw, :, l]); print("gpu tau", tau[iw, :, l])
print("a=k0*am0", k[iw, :, l, 0] * amount[0, l]); print("b=k1*am1", k[iw, :, l, 1] * amount[1, l])
