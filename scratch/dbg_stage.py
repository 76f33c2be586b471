import sys
sys.path.insert(0,'/root/repo')
import numpy as np
from tests.golden_util import load, stage_table
from tests.util import relerr, colerr, cpu
from archnemesis_dist_b200 import ops, plan
g = load("stages.npz")
tab = stage_table()["tab"]
otab = ops.OverlapTables(tab["DELG"])
am = ops.to_dev(g["ko_amount"])
t = cpu(ops.koverlap(ops.to_dev(g["ko_k"]), am, otab))
print('tau nograd equal', np.array_equal(t, g['ko_tau']), relerr(t, g['ko_tau']))
o64 = ops.OverlapTables(tab["DELG"].astype(np.float64))
print('seq flags', otab.seq, o64.seq)
t64 = cpu(ops.koverlap(ops.to_dev(g["ko_k"]), am, o64))
print('tau f64 equal', np.array_equal(t64, g['ko_tau_f64delg']), relerr(t64, g['ko_tau_f64delg']))
bad = np.argwhere(t64 != g['ko_tau_f64delg'])
print(len(bad), bad[:10])
for b in bad[:5]:
    print(repr(t64[tuple(b)]), repr(g['ko_tau_f64delg'][tuple(b)]))
