"""Host-side initial-conditions layer, batched with numpy (the reference keeps this layer on the host too).

Mirrors the reference types and functions it replaces for this path:
  Elements / ElementsIC / CartesianIC          src/ics/InitialConditions.jl:21-44, 64-107, 142-282
  init_nbody, kepcalc, d_dm, amatrix           src/ics/init_nbody.jl:13-229
  kepler_init (with 7x7 Jacobian)              src/ics/kepler_init.jl:66-210
  ekepler                                      src/ics/kepler.jl:1-41
  hierarchy([N,1,...,1]) (fully nested)        src/ics/setup_hierarchy.jl:9-29
  get_default_ICs / TRAPPIST-1, Kepler-36      src/ics/defaults.jl:3-85

All functions take a leading batch axis B (B = 1 for the reference's single-system calls).  Array index
order is numpy's [sys, body, ...]; the C ABI packs to the reference's column-major layouts.
"""
import numpy as np

YEAR = 365.242
GNEWT = 39.4845 / (YEAR * YEAR)  # src/NbodyGradient.jl:14-15
THIRD = 1.0 / 3.0


def nested_hierarchy(n):
    """Epsilon matrix of ElementsIC(t0, N::Int, ...) = hierarchy([N, 1, ..., 1])."""
    e = np.zeros((n, n))
    for i in range(n - 1):
        e[i, : i + 1] = -1.0
        e[i, i + 1] = 1.0
    e[n - 1, :] = -1.0
    return e


def hierarchy(H):
    """hierarchy(::Vector{Int}) for fully nested vectors [N,1,...,1]; pass an explicit matrix for anything else."""
    H = list(H)
    n = H[0]
    if H[1:] != [1] * (n - 1):
        raise NotImplementedError("only fully nested hierarchies [N,1,...,1] are generated; pass the epsilon matrix instead")
    return nested_hierarchy(n)


def trappist1_elements():
    """src/ics/defaults.jl:3-16 (== test/elements.txt)."""
    return np.array([
        [1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0],
        [2.5901135977661885e-5, 1.510880055106516, 7257.547487248826, 0.02436651768325364, 0.018169884000968452, 1.5707963267948966, 0.0],
        [5.7871255112412840e-5, 2.4218013609356652, 7258.592163817471, 0.020060810686211832, 0.011189705094395375, 1.5707963267948966, 0.0],
        [1.4602772830539989e-6, 4.0503542353950355, 7257.023855669221, 0.007411490159357976, -0.02016424872931776, 1.5707963267948966, 0.0],
        [1.9235328222249013e-5, 6.099281590191818, 7257.816770447013, 0.0011801938769616127, 0.000731913417670215, 1.5707963267948966, 0.0],
        [2.7302687390082730e-5, 9.20618480814173, 7257.1228936246725, -699952921060827e-19, 0.0002252519365921506, 1.5707963267948966, 0.0],
        [3.5331017018761430e-5, 12.353988709624156, 7257.667328639113, -0.0009722026578612578, 0.001276000403979281, 1.5707963267948966, 0.0],
        [1.6410627049780406e-6, 18.733535095576702, 7250.524231929195, -0.010402303111464135, -0.014289870200773339, 1.5707963267948966, 0.0],
    ])


def kepler36_elements():
    """src/ics/defaults.jl:18-28."""
    return np.array([
        [1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0],
        [1.2479484222582662e-5, 13.83989, 0.0, 0.002, -0.004, 1.5707963267948966, 0.0],
        [2.2659378094037732e-5, 16.23855, 5.062100000213832, -0.01, 0.007, 1.5707963267948966, 0.0],
    ])


_SYSTEMS = {"trappist-1": trappist1_elements, "trappist 1": trappist1_elements, "kepler-36": kepler36_elements, "kepler 36": kepler36_elements}


def available_systems():
    return ("trappist-1", "kepler-36")


class Elements:
    """Elements(m, P, t0, ecosω, esinω, I, Ω) — InitialConditions.jl:21-44 (keyword form, P-based subset)."""

    def __init__(self, m, P=0.0, t0=0.0, ecosw=0.0, esinw=0.0, I=0.0, Omega=0.0):
        e = np.hypot(ecosw, esinw)
        if not (0.0 <= e < 1.0):
            raise ValueError("Eccentricity must be in [0,1), e=%r" % e)
        if P < 0:
            raise ValueError("Period must be positive")
        self.m, self.P, self.t0, self.ecosw, self.esinw, self.I, self.Omega = map(float, (m, P, t0, ecosw, esinw, I, Omega))
        self.e = float(e)
        self.w = float(np.arctan2(esinw, ecosw))

    def row(self):
        return [self.m, self.P, self.t0, self.ecosw, self.esinw, self.I, self.Omega]


def amatrix(eps, m):
    """init_nbody.jl:176-188.  eps (n,n), m (B,n) -> A (B,n,n)."""
    n = eps.shape[0]
    same = (eps[:, :, None] == eps[:, None, :]).astype(float)  # same[i,j,l] = eps[i,j]==eps[i,l]
    summ = np.einsum("ijl,bl->bij", same, m)
    with np.errstate(invalid="ignore", divide="ignore"):
        return eps[None] * m[:, None, :] / summ


def _ekepler(m, ecc):
    """kepler.jl:1-41, vectorised; repeat-terminated fixed point on E - M."""
    pi2 = 2.0 * np.pi
    ms = np.mod(m, pi2)
    de0 = ecc * 0.85 * np.sign(ms)
    de1 = 2 * de0
    de2 = 3 * de0
    active = np.ones_like(m, dtype=bool)
    for _ in range(20):
        d2 = de1.copy()
        d1 = de0.copy()
        f3 = ecc * np.cos(de0 + ms)
        f2 = ecc * np.sin(de0 + ms)
        new = (f2 - d1 * f3) / (1 - f3)
        de0 = np.where(active, new, de0)
        de1 = np.where(active, d1, de1)
        de2 = np.where(active, d2, de2)
        active = active & ~((de0 == de1) | (de0 == de2))
        if not active.any():
            break
    return np.where(m != 0.0, de0 + m, 0.0)


def kepler_init(time, mass, el):
    """kepler_init.jl:66-210.  mass (B,), el (B,6)=(P,t0,ecosw,esinw,I,Omega) -> x (B,3), v (B,3), jac (B,7,7)."""
    B = mass.shape[0]
    period, t0, ecosom, esinom, inc, capom = (el[:, k] for k in range(6))
    n = 2 * np.pi / period
    semi = np.cbrt(GNEWT * mass * period ** 2 / 4 / np.pi ** 2)
    dsemidp = 2 * THIRD * semi / period
    dsemidm = THIRD * semi / mass
    ecc = np.sqrt(esinom ** 2 + ecosom ** 2)
    nz = ecc != 0.0
    safe = np.where(nz, ecc, 1.0)
    deccdecos = np.where(nz, ecosom / safe, 0.0)
    deccdesin = np.where(nz, esinom / safe, 0.0)
    s1 = np.sqrt(1.0 - ecc ** 2)
    den1 = esinom - ecosom - ecc
    tp_e = t0 - s1 / n * ecosom / (1.0 - esinom) - 2 / n * np.arctan2(np.sqrt(1.0 - ecc) * (esinom + ecosom + ecc), np.sqrt(1.0 + ecc) * den1)
    tp = np.where(nz, tp_e, t0 - 3 * period / 4)
    dtpdp = (tp - t0) / period
    fac = np.sqrt((1.0 - ecc) / (1.0 + ecc))
    with np.errstate(divide="ignore", invalid="ignore"):
        den2 = 1.0 / den1 ** 2
        theta = fac * (esinom + ecosom + ecc) / den1
        dthetadecc = ((ecc + ecosom) ** 2 + 2 * (1.0 - ecc ** 2) * esinom - esinom ** 2) / (s1 * (1.0 + ecc)) * den2
        dthetadecos = 2 * fac * esinom * den2
        dthetadesin = -2 * fac * (ecosom + ecc) * den2
        omes = 1.0 - esinom
        dtpdecc = ecc / s1 / n * ecosom / omes - 2 / n / (1.0 + theta ** 2) * dthetadecc
        dtpdecos = dtpdecc * deccdecos - s1 / n / omes - 2 / n / (1.0 + theta ** 2) * dthetadecos
        dtpdesin = dtpdecc * deccdesin - s1 / n * ecosom / omes ** 2 - 2 / n / (1.0 + theta ** 2) * dthetadesin
    m = n * (time - tp)
    dmdp = -m / period
    dmdtp = -n
    ekep = _ekepler(m, ecc)
    ce, se = np.cos(ekep), np.sin(ekep)
    r = semi * (1.0 - ecc * ce)
    denom = semi / r
    dekepdecos = se * denom * deccdecos
    dekepdesin = se * denom * deccdesin
    dekepdm = denom
    cO, sO = np.cos(capom), np.sin(capom)
    cw = np.where(nz, ecosom / safe, 1.0)
    sw = np.where(nz, esinom / safe, 0.0)
    ci, si = np.cos(inc), np.sin(inc)
    Z, O = np.zeros(B), np.ones(B)

    def mat(rows):
        return np.stack([np.stack(rw, axis=-1) for rw in rows], axis=-2)  # (B,3,3)

    P1 = mat([[cw, -sw, Z], [sw, cw, Z], [Z, Z, O]])
    P2 = mat([[O, Z, Z], [Z, ci, -si], [Z, si, ci]])
    P3 = mat([[cO, -sO, Z], [sO, cO, Z], [Z, Z, O]])
    Mi = mat([[Z, Z, Z], [Z, -si, -ci], [Z, ci, -si]])
    Mc = mat([[-sO, -cO, Z], [cO, -sO, Z], [Z, Z, Z]])
    P32 = P3 @ P2
    P321 = P32 @ P1
    vec = lambda a, b_, c: np.stack([a, b_, c], axis=-1)
    mv = lambda A, x: np.einsum("bij,bj->bi", A, x)
    col = lambda s: s[:, None]
    xplane = col(semi) * vec(ce - ecc, s1 * se, Z)
    vplane = vec(-se, s1 * ce, Z)
    x = mv(P321, xplane)
    dxda = x / col(semi)
    dxdekep = mv(P321 * semi[:, None, None], vplane)
    with np.errstate(divide="ignore", invalid="ignore"):
        dxdecc = mv(-P321 * semi[:, None, None] / ecc[:, None, None], vec(ce, se / s1, Z))
        dxdecos = dxdecc * col(deccdecos) + mv(P32 / ecc[:, None, None], xplane)
        dxdesin = dxdecc * col(deccdesin) + mv(P32 / ecc[:, None, None], vec(-xplane[:, 1], xplane[:, 0], Z))
    dxdinc = mv(P3 @ Mi @ P1, xplane)
    dxdcom = mv(Mc @ P2 @ P1, xplane)
    nsd = (n * semi * denom)[:, None, None]
    v = mv(P321 * nsd, vplane)
    dvda = v / col(semi)
    dvdp = -v / col(period)
    dvdekep = -v * col(ecc * se * denom) + mv(P321 * nsd, vec(-ce, -s1 * se, Z))
    with np.errstate(divide="ignore", invalid="ignore"):
        dvdecc = -v / col(ecc) + v * col(ce * denom) + mv(P321 * nsd, vec(Z, -ecc / s1 * ce, Z))
        dvdecos = dvdecc * col(deccdecos) + mv(P32 * nsd / ecc[:, None, None], vplane)
        dvdesin = dvdecc * col(deccdesin) + mv(P32 * nsd / ecc[:, None, None], vec(-vplane[:, 1], vplane[:, 0], Z))
    dvdinc = mv((P3 @ Mi @ P1) * nsd, vplane)
    dvdcom = mv((Mc @ P2 @ P1) * nsd, vplane)
    jac = np.zeros((B, 7, 7))
    c1 = col(dekepdm * (dmdp + dmdtp * dtpdp))
    c2 = col(dekepdm * dmdtp)
    c3 = col(dekepdm * dmdtp * dtpdecos + dekepdecos)
    c4 = col(dekepdm * dmdtp * dtpdesin + dekepdesin)
    nzc = col(nz)
    jac[:, 0:3, 0] = dxda * col(dsemidp) + dxdekep * c1
    jac[:, 0:3, 1] = dxdekep * c2
    jac[:, 0:3, 2] = np.where(nzc, dxdecos + dxdekep * c3, 0.0)
    jac[:, 0:3, 3] = np.where(nzc, dxdesin + dxdekep * c4, 0.0)
    jac[:, 0:3, 4] = dxdinc
    jac[:, 0:3, 5] = dxdcom
    jac[:, 0:3, 6] = dxda * col(dsemidm)
    jac[:, 3:6, 0] = dvdp + dvda * col(dsemidp) + dvdekep * c1
    jac[:, 3:6, 1] = dvdekep * c2
    jac[:, 3:6, 2] = np.where(nzc, dvdecos + dvdekep * c3, 0.0)
    jac[:, 3:6, 3] = np.where(nzc, dvdesin + dvdekep * c4, 0.0)
    jac[:, 3:6, 4] = dvdinc
    jac[:, 3:6, 5] = dvdcom
    jac[:, 3:6, 6] = dvda * col(dsemidm)
    jac[:, 6, 6] = 1.0
    return x, v, jac


def init_nbody_elements(elements, t0, eps=None, der=True):
    """init_nbody(ic::ElementsIC) — init_nbody.jl:13-27 with kepcalc :50-105 and d_dm :120-162.

    elements (B,n,7) -> x (B,n,3), v (B,n,3), jac_init (B,M,M) [row, col] (None when der is False)."""
    elements = np.asarray(elements, dtype=np.float64)
    if elements.ndim == 2:
        elements = elements[None]
    B, n, _ = elements.shape
    M = 7 * n
    eps = nested_hierarchy(n) if eps is None else np.asarray(eps, dtype=np.float64)
    m = elements[:, :, 0].copy()
    A = amatrix(eps, m)
    rk = np.zeros((B, n, 3))
    rdk = np.zeros((B, n, 3))
    jk = np.zeros((B, 6 * n, M))
    i, b = 1, 0
    while i < n:
        ind = eps[i - 1] != 0
        mu = m[:, ind].sum(axis=1)
        if not ind[0]:
            b += 1
        xk, vk, j21 = kepler_init(t0, mu, elements[:, i + b, 1:7])
        rk[:, i - 1] = xk
        rdk[:, i - 1] = vk
        if der:
            jk[:, (i - 1) * 6:(i - 1) * 6 + 6, i * 7:i * 7 + 6] = j21[:, 0:6, 0:6]
            for j in range(n):
                if eps[i - 1, j] != 0:
                    jk[:, (i - 1) * 6:(i - 1) * 6 + 6, j * 7 + 6] = j21[:, 0:6, 6]
        if b > 0:
            b -= 2
        elif b < 0:
            b = 0
        i += 1
    Ainv = np.linalg.inv(A)
    x = np.einsum("bij,bjk->bik", Ainv, rk)
    v = np.einsum("bij,bjk->bik", Ainv, rdk)
    if not der:
        return x, v, None
    # d_dm
    same = (eps[:, :, None] == eps[:, None, :]).astype(float)
    summ = np.einsum("ijl,bl->bij", same, m)  # Sigma m(i,j)
    kd = np.eye(n)
    with np.errstate(invalid="ignore", divide="ignore"):
        # dAdm[b,i,j,k] = delta(k,j) eps[i,j]/Sm[i,j] - delta(eps[i,j],eps[i,k]) eps[i,j] m[j]/Sm[i,j]^2   (init_nbody.jl:131-134)
        t1 = kd[None, None, :, :] * (eps[None] / summ)[:, :, :, None]
        t2 = same[None] * (eps[None] * m[:, None, :] / summ ** 2)[:, :, :, None]
        dAdm = t1 - t2
    dAinvdm = -np.einsum("bij,bjlk,blm->bimk", Ainv, dAdm, Ainv)  # [b,i,m,k]
    jac_init = np.zeros((B, M, M))
    J4 = jk.reshape(B, n, 6, M)
    blk = np.einsum("bik,bkrl->birl", Ainv, J4)  # (B,n,6,M)
    Jv = jac_init.reshape(B, n, 7, M)
    Jv[:, :, 0:6, :] = blk
    dxdm = np.einsum("bilk,blc->bikc", dAinvdm, rk)   # [b,i,k,c]
    dvdm = np.einsum("bilk,blc->bikc", dAinvdm, rdk)
    for k in range(n):
        Jv[:, :, 0:3, 7 * k + 6] += dxdm[:, :, k, :]
        Jv[:, :, 3:6, 7 * k + 6] += dvdm[:, :, k, :]
    for i in range(n):
        jac_init[:, 7 * i + 6, 7 * i + 6] = 1.0
    return x, v, jac_init


class ElementsIC:
    """ElementsIC(t0, H, elements) — InitialConditions.jl:142-241.  H: int (fully nested), hierarchy vector, or epsilon matrix.
    `elements` may be (n,7) or batched (B,n,7); rows beyond nbody are ignored as in the reference."""

    def __init__(self, t0, H, elements, der=True):
        if isinstance(elements, (list, tuple)) and elements and isinstance(elements[0], Elements):
            elements = np.array([e.row() for e in elements])
        elements = np.asarray(elements, dtype=np.float64)
        if isinstance(H, (int, np.integer)):
            eps = nested_hierarchy(int(H))
        else:
            H = np.asarray(H)
            eps = hierarchy(H.tolist()) if H.ndim == 1 else H.astype(np.float64)
        n = eps.shape[0]
        self.batched = elements.ndim == 3
        el = elements if self.batched else elements[None]
        self.elements = np.ascontiguousarray(el[:, :n, :]).copy()
        self.eps = eps
        self.nbody = n
        self.t0 = float(t0)
        self.der = der
        self.m = self.elements[:, :, 0].copy()

    @property
    def amat(self):
        return amatrix(self.eps, self.m)

    def init_nbody(self):
        # masses live in ic.m (test code perturbs them separately: test_transit_timing.jl:44-46)
        el = self.elements.copy()
        el[:, :, 0] = self.m
        return init_nbody_elements(el, self.t0, self.eps, self.der)


class CartesianIC:
    """CartesianIC(t0, N, coords) — InitialConditions.jl:259-282; coords rows = m, x, y, z, vx, vy, vz."""

    def __init__(self, t0, N, coords):
        coords = np.asarray(coords, dtype=np.float64)
        self.batched = coords.ndim == 3
        c = coords if self.batched else coords[None]
        self.nbody = int(N)
        self.m = c[:, :N, 0].copy()
        self.x = c[:, :N, 1:4].copy()
        self.v = c[:, :N, 4:7].copy()
        self.t0 = float(t0)

    def init_nbody(self):
        B, n = self.m.shape
        return self.x.copy(), self.v.copy(), np.broadcast_to(np.eye(7 * n), (B, 7 * n, 7 * n)).copy()


def get_default_ICs(system_name, t0=0.0, n=0):
    """src/ics/defaults.jl:62-85."""
    key = system_name.lower()
    if key not in _SYSTEMS:
        raise ValueError("%s not an available system name." % system_name)
    el = _SYSTEMS[key]()
    nmax = el.shape[0]
    if n == 0 or n > nmax:
        n = nmax
    return ElementsIC(t0, n, el)
