"""TEST INFRASTRUCTURE: an oracle-backed stand-in for ``archnemesis_dist_b200.engine`` so that the
host logic of the drop-in (forward_model.py: continuum assembly, surface terms, projection folding,
layout conversions, the nemesisfmg override) can be exercised in the CPU container against the live
reference.  It implements the same ``HotPath`` surface with numpy arrays instead of device tensors.
Never imported by the product."""
import numpy as np

from archnemesis_dist_b200 import plan as _plan
from archnemesis_dist_b200.engine import Evaluation, THERMAL, TRANSMISSION  # noqa: F401  (re-exported)
from oracle import oracle as orc


class Staged:
    pass


class HotPath:
    def __init__(self, K, PRESS, TEMP, DELG, WAVE, ops=None):
        self.K = np.asarray(K, dtype=np.float64)
        self.lbl_table = self.K.ndim == 4
        self.PRESS, self.TEMP, self.DELG = np.asarray(PRESS), np.asarray(TEMP), np.asarray(DELG)
        self.WAVE = np.asarray(WAVE, dtype=np.float64)
        self.launches = 0

    def stage(self, ev, return_grad, M=None):
        s = Staged()
        s.ev, s.grad, s.M = ev, return_grad, M
        if getattr(ev, "continuum", None) is not None:
            # the device makes the dense continuum arrays from the plan (ansb200_continuum); here the oracle does
            import copy
            tables, cplan = ev.continuum
            ev = copy.copy(ev)
            ev.taucia, ev.taudust, ev.tauray, ev.dtaucon = orc.continuum_eval(tables, cplan, return_grad)
            s.ev = ev
            HotPath.continuum_plans += 1
        return s

    continuum_plans = 0

    _cont_lock = __import__("threading").Lock()

    def continuum_tables(self, key, factory):
        with HotPath._cont_lock:
            cache = self.__dict__.setdefault("_cont_tables", {})
            if key not in cache:
                cache.clear()
                cache[key] = factory()
            return cache[key]

    def gas_opacity(self, s, timers=None):
        ev = s.ev
        if self.lbl_table:
            return orc.lbl_table_opacity(self.K, self.PRESS, self.TEMP, ev.press_atm, ev.temp, ev.amount, s.grad)
        if s.grad:
            k, d = orc.calc_k(self.K, self.PRESS, self.TEMP, ev.press_atm, ev.temp, want_grad=True)
            return orc.k_overlap(self.DELG, k, ev.amount, dkdT=d)
        k = orc.calc_k(self.K, self.PRESS, self.TEMP, ev.press_atm, ev.temp)
        return orc.k_overlap(self.DELG, k, ev.amount)

    def run(self, s, timers=None):
        ev = s.ev
        go = self.gas_opacity(s)
        tau, dk = go if s.grad else (go, None)
        nw, nlay = tau.shape[0], tau.shape[2]
        z = np.zeros((nw, nlay))
        # the reference adds TAUGAS + TAUCIA + TAUDUST + TAURAY in that order (ForwardModel_0.py:3989)
        t = tau + (ev.taucia if ev.taucia is not None else z)[:, None, :]
        t = t + (ev.taudust if ev.taudust is not None else z)[:, None, :]
        t = t + (ev.tauray if ev.tauray is not None else z)[:, None, :]
        dtc = ev.dtaucon if ev.dtaucon is not None else np.zeros((nw, ev.NPAR, nlay))
        tl, tp, dtl = orc.assemble_opacity(t, dk, ev.gas_slot, ev.NVMR, ev.NPAR, z, dtc, ev.LAYINC, ev.SCALE)
        xfac = ev.xfac if ev.xfac is not None else np.ones(nw)
        if ev.mode == THERMAL:
            zz = np.zeros(nw)
            S, dS, dT = orc.thermal_paths(ev.ISPACE, self.WAVE, tl, dtl, ev.NVMR, ev.NLAYIN, ev.EMTEMP, ev.LAYPRESS,
                                          ev.LAYINC, ev.TSURF, ev.EMISSIVITY if ev.EMISSIVITY is not None else zz, xfac,
                                          ev.SOLFLUX if ev.SOLFLUX is not None else zz,
                                          ev.REFLECTANCE if ev.REFLECTANCE is not None else zz,
                                          ev.SOL_ANG, ev.EMISS_ANG)
        else:
            S, dS = orc.transmission(tp, dtl, xfac)
            dT = np.zeros_like(S) if s.grad else None
        if not s.grad:
            return orc.g_integrate(S, None, None, self.DELG)
        spec, dspec, dts = orc.g_integrate(S, dS, dT, self.DELG)
        dspec_dev = np.ascontiguousarray(np.transpose(dspec, (0, 3, 1, 2)))     # device layout [NWAVE,NPATH,NPAR,NLAYMAX]
        if s.M is None:
            return spec, dspec_dev, dts
        nwv, npath, npar, nlm = dspec_dev.shape
        dx = np.einsum("wpe,pex->wpx", dspec_dev.reshape(nwv, npath, npar * nlm), s.M)
        return spec, dx, dts

    def cirsrad(self, ev, return_grad=False):
        return self.run(self.stage(ev, return_grad))

    def forward_jacobian(self, ev, M):
        return self.run(self.stage(ev, True, M))

    def forward_jacobian_conv(self, ev, M, conv_op, jsurf=-1, wgeom=1.0):
        spec, dx, dts = self.forward_jacobian(ev, M)
        block = np.concatenate([spec[:, :1], dx[:, 0, :]], axis=1)
        if jsurf >= 0:
            block[:, 1 + jsurf] = dts[:, 0]
        if wgeom != 1.0:
            block = block * float(wgeom)
        return np.concatenate([orc.apply_conv(conv_op, block[:, 0])[:, None], orc.apply_conv(conv_op, block[:, 1:])], axis=1)

    def forward_jacobian_mix_conv(self, ev, M, mix, conv_op, Mlay=None):
        spec, dx, _ = self.forward_jacobian(ev, M)
        full = np.concatenate([spec[:, :, None], dx], axis=2)                  # [NWAVE, NPATH, 1+NX]
        block = np.zeros((full.shape[0], len(mix["lo"]), full.shape[2]))
        for i, (lo, hi, a, b) in enumerate(zip(mix["lo"], mix["hi"], mix["wlo"], mix["whi"])):
            block[:, i] = full[:, lo] if hi < 0 else full[:, lo] * a + full[:, hi] * b
        nw, ngeom, nc = block.shape
        out = orc.apply_conv(conv_op, block.reshape(nw, ngeom * nc))
        return block[:, :, 0], out.reshape(out.shape[0], ngeom, nc)

    def project(self, dspec, M):
        nwv, npath, npar, nlm = dspec.shape
        return np.einsum("wpe,pex->wpx", np.asarray(dspec).reshape(nwv, npath, npar * nlm), M)

    def conv_operator(self, op):
        return op

    @staticmethod
    def to_host(t):
        return np.asarray(t)

    def close(self):
        pass
