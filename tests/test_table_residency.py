"""CPU: the memoised Spectroscopy_0.read_tables of forward_model.install() (table residency across evaluations,
SURVEY.md 8f-3) on a stand-in class with the reference method's contract: hits re-install the same read-only arrays,
a touched file or a different range is read again, run-time line-by-line and on-line tables bypass the cache."""
import os

import numpy as np

from archnemesis_dist_b200.forward_model import make_cached_read_tables


class FakeSpectroscopy:
    reads = 0

    def __init__(self, files, ilbl=0, online=False):
        self.ILBL, self.LOCATION, self.ONLINE = ilbl, files, online
        self.NG, self.NP, self.NT, self.NGAS = 4, 3, 2, len(files)
        self.WAVE, self.NWAVE, self.K = None, 0, None

    def read_header(self):
        self.WAVE = np.linspace(10.0, 20.0, 101)
        self.NWAVE = 101

    def ref_read_tables(self, wavemin=0., wavemax=1.0e10, wavedelta=1.0):
        type(self).reads += 1
        if self.WAVE is None:
            self.read_header()
        keep = (self.WAVE >= wavemin) & (self.WAVE <= wavemax)
        self.WAVE = self.WAVE[keep].copy()
        self.NWAVE = len(self.WAVE)
        self.K = np.full((self.NWAVE, self.NG, self.NP, self.NT, self.NGAS), float(type(self).reads))


def test_memoised_read_tables(tmp_path):
    files = []
    for i in range(2):
        f = tmp_path / ("gas%d.kta" % i)
        f.write_bytes(b"x" * (10 + i))
        files.append(str(f)[:-4])                       # LOCATION entries come without the extension
    FakeSpectroscopy.reads = 0
    cached = make_cached_read_tables(FakeSpectroscopy.ref_read_tables)
    a = FakeSpectroscopy(files)
    cached(a, wavemin=12.0, wavemax=15.0)
    assert FakeSpectroscopy.reads == 1 and cached.hits == 0 and not a.K.flags.writeable and not a.WAVE.flags.writeable
    b = FakeSpectroscopy(files)                        # the deep copy of the next evaluation
    cached(b, wavemin=12.0, wavemax=15.0)
    assert FakeSpectroscopy.reads == 1 and cached.hits == 1
    assert b.K is a.K and b.WAVE is a.WAVE and b.NWAVE == a.NWAVE == 31
    c = FakeSpectroscopy(files)                        # another range: read
    cached(c, wavemin=12.0, wavemax=16.0)
    assert FakeSpectroscopy.reads == 2 and c.K is not a.K
    st = os.stat(files[0] + ".kta")                    # a rewritten table file: read again
    os.utime(files[0] + ".kta", ns=(st.st_atime_ns, st.st_mtime_ns + 10**9))
    d = FakeSpectroscopy(files)
    cached(d, wavemin=12.0, wavemax=15.0)
    assert FakeSpectroscopy.reads == 3 and d.K is not a.K
    for kw in (dict(ilbl=1), dict(online=True)):       # run-time LBL / on-line HDF5: straight to the reference method
        e = FakeSpectroscopy(files, **kw)
        n = FakeSpectroscopy.reads
        cached(e, wavemin=12.0, wavemax=15.0)
        cached(e, wavemin=12.0, wavemax=15.0)
        assert FakeSpectroscopy.reads == n + 2 and e.K.flags.writeable
    # the cache is bounded
    for lo in range(8):
        cached(FakeSpectroscopy(files), wavemin=10.0 + lo, wavemax=19.0)
    assert len(cached.cache) <= 4
