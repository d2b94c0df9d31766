"""Config 3 inputs and metric: the 5-body outer solar system of examples/outer_ss_example.jl:16-33
(Hairer, Lubich & Wanner 2006, Sept 5 1994; AU, AU/day, solar masses), shifted to the centre of mass as :36-48,
and the energy / angular-momentum definition of compute_energy :83-99.  Data, not code, from the reference."""
import numpy as np

GNEWT = 39.4845 / (365.242 * 365.242)

_X = np.array([[-2.079997415328555E-04, 7.127853194812450E-03, -1.352450694676177E-05],
               [-3.502576700516146E+00, -4.111754741095586E+00, 9.546978009906396E-02],
               [9.075323061767737E+00, -3.443060862268533E+00, -3.008002403885198E-01],
               [8.309900066449559E+00, -1.782348877489204E+01, -1.738826162402036E-01],
               [1.147049510166812E+01, -2.790203169301273E+01, 3.102324955757055E-01]])
_V = np.array([[-6.227982601533108E-06, 2.641634501527718E-06, 1.564697381040213E-07],
               [5.647185656190083E-03, -4.540768041260330E-03, -1.077099720398784E-04],
               [1.677252499111402E-03, 5.205044577942047E-03, -1.577215030049337E-04],
               [3.535508197097127E-03, 1.479452678720917E-03, -4.019422185567764E-05],
               [2.882592399188369E-03, 1.211095412047072E-03, -9.118527716949448E-05]])
_M = np.array([1.00000597682, 0.000954786104043, 0.000285583733151, 0.0000437273164546, 0.0000517759138449])


def outer_ss_cartesian():
    """Returns m (5,), x (5,3), v (5,3) in the centre-of-mass frame."""
    m, x, v = _M.copy(), _X.copy(), _V.copy()
    xcm = (m[:, None] * x).sum(0) / m.sum()
    vcm = (m[:, None] * v).sum(0) / m.sum()
    return m, x - xcm, v - vcm


def energy_angmom(m, x, v):
    """KE + PE and the angular-momentum vector."""
    ke = 0.5 * np.sum(m * np.sum(v * v, axis=1))
    pe = 0.0
    n = len(m)
    for j in range(n - 1):
        for k in range(j + 1, n):
            pe += -GNEWT * m[j] * m[k] / np.linalg.norm(x[j] - x[k])
    L = np.sum(m[:, None] * np.cross(x, v), axis=0)
    return ke + pe, L
