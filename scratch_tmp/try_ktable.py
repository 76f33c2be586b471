import sys, numpy as np, time
sys.path.insert(0, "/root/repo")
from tests.test_lbl_dropin import _reference_objects, _line_set_spec, _line_data, _spectroscopy
ans, ld, LineSetData = _reference_objects()
sp_mod = sys.modules["archnemesis.Spectroscopy_0"]
lss = _line_set_spec(ans, ld, LineSetData, 5, 1, 300, 1990.0, 2110.0, 3)
obj = _line_data(ans, ld, [lss], 5, 1)
obj._params_fetched_lines_last = True
obj._params_fetched_partition_last = True
obj.set_params = lambda **k: obj
wave_lbl = np.linspace(2040.0, 2060.0, 11)
S_LBL = _spectroscopy(ans, ld, wave_lbl.copy(), [(5, 1, obj)])
# the k-table spectroscopy: bins
S = ans.Spectroscopy_0(ILBL=ans.enum.SpectralCalculationModeEnum.K_TABLES)
S.ISPACE = ans.enum.WaveUnitEnum.Wavenumber_cm
S.WAVE = np.linspace(2045.0, 2055.0, 21); S.NWAVE = 21
S.NG = 10
from numpy.polynomial.legendre import leggauss
x, w = leggauss(10)
S.G_ORD = 0.5 * (x + 1); S.DELG = 0.5 * w
S.NP, S.NT, S.PRESS, S.TEMP = 2, 2, np.array([1e-2, 1.0]), np.array([180.0, 260.0])
t = time.time()
k = sp_mod.calc_ktable_chunk(np.arange(3, 12), S, S_LBL, 0.3, None)
print(k.shape, time.time() - t, k[0, :, 0, 0])
