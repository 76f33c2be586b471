"""Optimal-estimation linear algebra on the device: the consumer of the all-gathered KK.

Mirrors ``OptimalEstimation_0.calc_gain_matrix`` / ``calc_phiret`` / ``calc_next_xn`` / ``calc_serr``
(archnemesis/OptimalEstimation_0.py:545-720) on device tensors, so that KK[NY,NX] -- assembled on every rank by the
NCCL all-gather of ``dist.py`` -- never has to visit the host before the state update.  These are plain dense
products and one NY x NY solve: library work (cuBLAS / cuSOLVER through ``torch.linalg``), the same Rodgers form
and the same association order as the reference, no hand-written kernel.  No CPU fallback: every function raises
without a CUDA device.

``OptimalEstimationDeviceMixin`` puts the four methods over an ``OptimalEstimation_0`` instance (attributes KK,
SA, SE, Y, YN, XA, XN, NX, NY in; DD, AA, CHISQ, PHI, SM, SN, ST out, as numpy arrays like the reference).
"""
import numpy as np
import torch

from .ops import _require_cuda, to_dev


def _t(a):
    return a if isinstance(a, torch.Tensor) else to_dev(np.asarray(a, dtype=np.float64))


def calc_gain_matrix(KK, SA, SE):
    """dd = sa kk^T (kk sa kk^T + se)^-1 by a linear solve, aa = dd kk  (:545-559).  Returns DD[NX,NY], AA[NX,NX]."""
    _require_cuda()
    KK, SA, SE = _t(KK), _t(SA), _t(SE)
    sa_kt = SA @ KK.T                       # (NX, NY)
    M = KK @ sa_kt + SE                     # (NY, NY); a (1,1) SE broadcasts like the reference
    X_T = torch.linalg.solve(M.T, sa_kt.T)  # (NY, NX)
    DD = X_T.T.contiguous()
    return DD, DD @ KK


def calc_phiret(Y, YN, XN, XA, SE, SA):
    """Cost function (:560-600): CHISQ = (yn-y)^T se^-1 (yn-y) / NY, PHI = that + (xn-xa)^T sa^-1 (xn-xa).
    SE may be (1,1), diagonal or full, as in the reference.  Returns python floats (CHISQ, PHI)."""
    _require_cuda()
    Y, YN, XN, XA, SE, SA = (_t(a) for a in (Y, YN, XN, XA, SE, SA))
    b = YN - Y
    d = XN - XA
    if tuple(SE.shape) == (1, 1):
        meas = float(torch.dot(b, b) / SE[0, 0])
    elif bool(torch.all(SE == torch.diag(torch.diagonal(SE)))):
        meas = float(torch.dot(b / torch.diagonal(SE), b))
    else:
        meas = float(torch.dot(b, torch.linalg.solve(SE, b)))
    apri = float(torch.dot(d, torch.linalg.solve(SA, d)))
    if np.isnan(meas + apri):
        raise AssertionError("PHI cannot be NAN")
    return meas / b.numel(), meas + apri


def calc_next_xn(XA, XN, Y, YN, DD, AA):
    """xn+1 = xa + dd (y - yn) - aa (xa - xn)  (:632-652)."""
    _require_cuda()
    XA, XN, Y, YN, DD, AA = (_t(a) for a in (XA, XN, Y, YN, DD, AA))
    return XA + DD @ (Y - YN) - AA @ (XA - XN)


def calc_serr(DD, AA, SA, SE, simple=False):
    """sm = dd se dd^T, sn = (aa-I) sa (aa-I)^T, st = sn + sm  (:654-690)."""
    _require_cuda()
    DD, AA, SA, SE = (_t(a) for a in (DD, AA, SA, SE))
    a = DD * SE[0, 0] if simple else DD @ SE
    SM = a @ DD.T
    b = AA - torch.eye(AA.shape[0], dtype=AA.dtype, device=AA.device)
    SN = (b @ SA) @ b.T
    return SM, SN, SN + SM


def _np(a):
    return a.cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


class OptimalEstimationDeviceMixin:
    """The four methods over an OptimalEstimation_0-like object; results land on the attributes the reference
    sets, as numpy arrays.  ``class OE_B200(OptimalEstimationDeviceMixin, archnemesis.OptimalEstimation_0)`` is what
    ``forward_model.install()`` puts under ``coreretOE`` (OptimalEstimation_0.py:1263).  `b200_oe` is the module
    whose functions do the algebra (this one; tests substitute an oracle-backed namespace on CPU-only machines)."""

    b200_oe = None

    def _oe(self):
        import sys
        return self.b200_oe if self.b200_oe is not None else sys.modules[__name__]

    def calc_gain_matrix(self):
        DD, AA = self._oe().calc_gain_matrix(self.KK, self.SA, self.SE)
        self.DD, self.AA = _np(DD), _np(AA)

    def calc_phiret(self):
        self.CHISQ, self.PHI = self._oe().calc_phiret(self.Y[:self.NY], self.YN[:self.NY], self.XN[:self.NX],
                                                      self.XA[:self.NX], self.SE, self.SA)
        assert not np.isnan(self.PHI), "PHI cannot be NAN"
        assert not np.isnan(self.CHISQ), "CHISQ cannot be NAN"

    def calc_next_xn(self):
        return _np(self._oe().calc_next_xn(self.XA, self.XN, self.Y, self.YN, self.DD, self.AA))

    def calc_serr(self, simple=False):
        SM, SN, ST = self._oe().calc_serr(self.DD, self.AA, self.SA, self.SE, simple)
        self.SM, self.SN, self.ST = _np(SM), _np(SN), _np(ST)


def make_oe_class(reference_cls):
    """OE_B200: the reference's OptimalEstimation_0 with the four algebra methods on the device."""
    cls = type("OE_B200", (OptimalEstimationDeviceMixin, reference_cls),
               {"__doc__": (reference_cls.__doc__ or "") + "\n\n(B200: gain matrix, cost function, state update and error "
                                                              "covariances on the device -- archnemesis_dist_b200.oe)"})
    return cls
