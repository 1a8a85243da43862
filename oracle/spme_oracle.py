"""CPU restatement (numpy) of the reference's SPME reciprocal-space Coulomb path for ONE domain -- TEST INFRASTRUCTURE ONLY.

Only tests/ may import this module; the product never does.  PARITY UNPINNED: the reference holds no vectors for this path
and cannot be built here (Fortran), so this restatement is pinned only (a) against the exact Ewald reciprocal sum, to the
accuracy SPME has at the chosen order / grid, and (b) by the consistency of its forces with the finite-difference derivative
of its own energy (tests/test_spme_oracle.py).

What is restated, routine by routine (mxnode = 1: the domain is the whole grid, exchange_grid is the periodic wrap):
  control.F90:1707-1713            alpha and the k-grid from spme_precision      -> spme_grid
  parallel_fft.F90:1999-2054       adjust_kmax / pfft_length_ok                  -> adjust_kmax
  bspline.F90:73-190               bspline_coeffs_gen (|b(m)|^2 per dimension)   -> bspline_norm2
  bspline.F90:192-306              bspline_splines_gen (values + 1st derivative) -> bspline_splines
  ewald_general.F90:517-576        spme_construct_charge_array                   -> charge_grid
  ewald_spole.F90:1257-1386        spme_construct_potential_grid_coul            -> potential_grid
  ewald_general.F90:717-869        spme_calc_force_energy                        -> force_energy
  spme.F90:159-231                 spme_self_interaction                         -> self_interaction
  ewald_spole.F90:244-477          ewald_spme_forces_coul (the driver)           -> ewald_spme_forces_coul
"""
import math

import numpy as np

SQRPI = 1.7724538509055160273        # constants.F90:56
ZERO_PLUS = np.finfo(np.float64).tiny


def pfft_length_ok(n):
    """parallel_fft.F90:2031-2054: n = 2^a 3^b 5^c."""
    for p in (2, 3, 5):
        while n % p == 0 and n > 1:
            n //= p
    return n == 1


def adjust_kmax(kmax, P=1):
    """parallel_fft.F90:1999-2029."""
    if kmax % P != 0:
        kmax = (kmax // P + 1) * P
    while not pfft_length_ok(kmax // P):
        kmax += P
    return kmax


def dcell_widths(cell):
    """numerics.F90 dcell(7:9): perpendicular widths of the cell (rows of `cell` are the lattice vectors)."""
    a, b, c = np.asarray(cell, dtype=np.float64).reshape(3, 3)
    vol = abs(np.dot(a, np.cross(b, c)))
    return np.array([vol / np.linalg.norm(np.cross(b, c)), vol / np.linalg.norm(np.cross(c, a)), vol / np.linalg.norm(np.cross(a, b))])


def spme_grid(precision, rcut, cell, dims=(1, 1, 1)):
    """control.F90:1707-1713 + ewald.F90:1210-1212: (alpha, k_vec_dim) from spme_precision."""
    tol = math.sqrt(abs(math.log(precision * rcut)))
    alpha = math.sqrt(abs(math.log(precision * rcut * tol))) / rcut
    tol1 = math.sqrt(-math.log(precision * rcut * (2.0 * tol * alpha) ** 2))
    w = dcell_widths(cell)
    k = [2 * int(math.floor(0.25 + w[d] * alpha * tol1 / math.pi + 0.5)) for d in range(3)]   # Nint
    return alpha, tuple(adjust_kmax(k[d], dims[d]) for d in range(3))


def _cardinal(n, x):
    """Cardinal B-spline M_n(x) of order n (support [0, n]) by the Cox-de Boor recursion; x any array."""
    x = np.asarray(x, dtype=np.float64)
    if n == 2:
        return np.where((x >= 0.0) & (x <= 2.0), 1.0 - np.abs(x - 1.0), 0.0)
    return (x * _cardinal(n - 1, x) + (n - x) * _cardinal(n - 1, x - 1.0)) / (n - 1.0)


def bspline_norm2(kdim, n):
    """bspline.F90:73-190: norm2(d, i) = |b_d(i)|^2, b(i) = w^{i(n-1)} / sum_{k=0}^{n-2} M_n(k+1) w^{ik}, w = exp(2 pi i / K)."""
    out = []
    csp = _cardinal(n, np.arange(1, n, dtype=np.float64))          # cspline(k+2) = M_n(k+1), k = 0..n-2 (:128-134)
    for K in kdim:
        i = np.arange(K)
        den = np.zeros(K, dtype=np.complex128)
        for k in range(n - 1):
            den += csp[k] * np.exp(2j * np.pi * ((i * k) % K) / K)
        b = np.exp(2j * np.pi * ((i * (n - 1)) % K) / K) / den
        out.append((b * np.conj(b)).real)
    return out


def bspline_splines(u, n):
    """bspline.F90:192-306 for num_deriv >= 1: for scaled coordinates u (natoms, 3) returns (idx, d0, d1) with
    idx = Int(u), d0[:, :, l-1] = derivs(:, 0, l, i) = M_n(w + n - l), d1 = the first derivative, l = 1..n: spline l sits on grid
    point (1-based) idx + 1 - n + l (ewald_general.F90:788-797)."""
    idx = np.trunc(u).astype(np.int64)                              # recip_indices = Int(recip_coords)
    w = u - np.trunc(u)                                             # recip_coords - Aint(recip_coords)
    l = np.arange(1, n + 1, dtype=np.float64)
    arg = w[:, :, None] + (n - l)[None, None, :]
    d0 = _cardinal(n, arg)
    d1 = _cardinal(n - 1, arg) - _cardinal(n - 1, arg - 1.0)
    return idx, d0, d1


def charge_grid(kdim, idx, d0, q):
    """ewald_general.F90:517-576 with the +side images wrapped back (mxnode = 1): Q(j,k,l) += q d0x d0y d0z."""
    n = d0.shape[2]
    Q = np.zeros(kdim)
    live = np.abs(q) > ZERO_PLUS
    off = np.arange(1, n + 1) - n                                   # 0-based grid point of spline l: idx - n + l
    for a in np.nonzero(live)[0]:
        jx = (idx[a, 0] + off) % kdim[0]; jy = (idx[a, 1] + off) % kdim[1]; jz = (idx[a, 2] + off) % kdim[2]
        np.add.at(Q, (jx[:, None, None], jy[None, :, None], jz[None, None, :]),
                  q[a] * d0[a, 0][:, None, None] * d0[a, 1][None, :, None] * d0[a, 2][None, None, :])
    return Q


def potential_grid(Q, rcell, alpha, norm2):
    """ewald_spole.F90:1257-1386: forward FFT, B(m) exp(-x^2) / (sqrt(pi) x^2) with x = pi |m| / alpha inside the spherical
    k cutoff, the stress kernel, backward FFT (both transforms unnormalised).  rcell(9) is the Fortran-order inverse cell:
    recip_pos = jj rcell(1:9:3) + kk rcell(2:9:3) + ll rcell(3:9:3).  Returns (real potential grid, stress_contrib(9))."""
    K = Q.shape
    conv = math.pi / alpha
    test_fac = (1.0e-6 / conv) ** 2
    rc = np.asarray(rcell, dtype=np.float64)
    widths = dcell_widths(rc.reshape(3, 3))                         # dcell(recip_cell)(7:9)
    cut2 = (0.5 * 1.05 * np.min(np.array(K, dtype=np.float64) * widths)) ** 2
    S = np.fft.ifftn(Q) * Q.size                                    # direction 1, exp(+i ...), unnormalised (Q is real: the sign is immaterial)
    fr = [np.where(2 * np.arange(k) > k, np.arange(k) - k, np.arange(k)).astype(np.float64) for k in K]
    jj, kk, ll = np.meshgrid(fr[0], fr[1], fr[2], indexing="ij")
    m = np.stack([jj * rc[0] + kk * rc[1] + ll * rc[2], jj * rc[3] + kk * rc[4] + ll * rc[5], jj * rc[6] + kk * rc[7] + ll * rc[8]], 0)
    k2 = (m * m).sum(0)
    ok = (k2 <= cut2) & (k2 > test_fac)
    bb = norm2[0][:, None, None] * norm2[1][None, :, None] * norm2[2][None, None, :]
    x2 = np.where(ok, k2, 1.0) * conv * conv
    comp = np.where(ok, bb * S * np.exp(-x2) / (SQRPI * x2), 0.0)
    pv = np.where(ok, (comp * (-2.0 * ((1.0 + x2) / np.where(ok, k2, 1.0))) * np.conj(S)).real, 0.0)
    stress = np.zeros((3, 3))
    for a in range(3):
        for b in range(3):
            stress[b, a] = (m[a] * m[b] * pv).sum()
    phi = np.fft.fftn(comp)                                         # direction -1, unnormalised
    return phi.real, stress.reshape(9, order="F")


def force_energy(phi, rcell, kdim, idx, d0, d1, q, megatm):
    """ewald_general.F90:717-869: per-atom gather; returns (energy_sum, forces(natms, 3)) before the `scale` factors."""
    n = d0.shape[2]
    rmat = np.asarray(rcell, dtype=np.float64).reshape(3, 3, order="F")
    recip_kmax = rmat @ np.array(kdim, dtype=np.float64)
    off = np.arange(1, n + 1) - n
    nat = len(q)
    f = np.zeros((nat, 3))
    e_tot = 0.0
    f_tot = np.zeros(3)
    for a in range(nat):
        if abs(q[a]) <= ZERO_PLUS:
            continue
        jx = (idx[a, 0] + off) % kdim[0]; jy = (idx[a, 1] + off) % kdim[1]; jz = (idx[a, 2] + off) % kdim[2]
        g = phi[jx[:, None, None], jy[None, :, None], jz[None, None, :]]
        x0, y0, z0 = d0[a, 0][:, None, None], d0[a, 1][None, :, None], d0[a, 2][None, None, :]
        x1, y1, z1 = d1[a, 0][:, None, None], d1[a, 1][None, :, None], d1[a, 2][None, None, :]
        e_tot += q[a] * (x0 * y0 * z0 * g).sum()
        cur = q[a] * np.array([(x1 * y0 * z0 * g).sum() * recip_kmax[0], (x0 * y1 * z0 * g).sum() * recip_kmax[1],
                               (x0 * y0 * z1 * g).sum() * recip_kmax[2]])
        f_tot -= cur
        f[a] -= cur
    f -= f_tot / float(megatm)                                      # :862-866, every atom of the domain
    return e_tot, f


def self_interaction(q, alpha, scaling):
    """spme.F90:159-231, pot_order 1: -sum q^2 scaling alpha / Gamma(1/2)."""
    return -float((q * q).sum()) * scaling * alpha / SQRPI


def ewald_spme_forces_coul(cell, xyz, q, alpha, kdim, nspl, scaling):
    """ewald_spole.F90:244-477 for one domain holding every atom.  cell: 9 doubles (rows = lattice vectors), xyz (natms, 3),
    q charges.  Returns dict(engcpe_rc, vircpe_rc, stress(9), forces(natms, 3))."""
    cell = np.asarray(cell, dtype=np.float64).reshape(3, 3)
    xyz = np.asarray(xyz, dtype=np.float64); q = np.asarray(q, dtype=np.float64)
    inv = np.linalg.inv(cell)                                        # invert(cell, rcell): s_dim = rcell(dim) x + rcell(dim+3) y + rcell(dim+6) z,
    rcell = inv.reshape(9)                                           # r = s_a a + s_b b + s_c c  =>  rcell(dim + 3 j) = inv[j][dim - 1]
    volm = abs(np.linalg.det(cell))
    scale = math.pi * SQRPI * alpha ** (1 - 3) * (0.5 / volm) * scaling
    kr = np.array(kdim, dtype=np.float64)
    u = kr[None, :] * (xyz @ inv + 0.5)                              # :311-318
    idx, d0, d1 = bspline_splines(u, nspl)
    Q = charge_grid(tuple(kdim), idx, d0, q)
    phi, s0 = potential_grid(Q, rcell, alpha, bspline_norm2(kdim, nspl))
    e0, f = force_energy(phi, rcell, tuple(kdim), idx, d0, d1, q, len(q))
    eng = e0 * scale
    f = f * scale * 2.0
    st = s0 * scale
    st[0::4] += eng
    return {"engcpe_rc": eng + self_interaction(q, alpha, scaling), "vircpe_rc": -float(st[0::4].sum()), "stress": st, "forces": f,
            "eng_recip": eng, "kdim": tuple(kdim)}


def ewald_recip_exact(cell, xyz, q, alpha, scaling, mmax=None, tol=1e-16):
    """The exact Ewald reciprocal energy scaling / (2 pi V) sum_{m != 0} exp(-pi^2 m^2 / alpha^2) / m^2 |S(m)|^2 by brute force
    (the pin of this restatement; O(N n_k), small systems only)."""
    cell = np.asarray(cell, dtype=np.float64).reshape(3, 3)
    rec = np.linalg.inv(cell).T                                      # rows = reciprocal lattice vectors (no 2 pi)
    volm = abs(np.linalg.det(cell))
    if mmax is None:
        gmin = np.min(np.linalg.norm(rec, axis=1))
        mmax = int(math.ceil(math.sqrt(-math.log(tol)) * alpha / math.pi / gmin)) + 1
    r = np.arange(-mmax, mmax + 1)
    h, k, l = np.meshgrid(r, r, r, indexing="ij")
    hk = np.stack([h.ravel(), k.ravel(), l.ravel()], 1)
    hk = hk[(hk != 0).any(1)]
    m = hk @ rec
    m2 = (m * m).sum(1)
    w = np.exp(-math.pi ** 2 * m2 / alpha ** 2) / m2
    keep = w > tol * w.max()
    m, w = m[keep], w[keep]
    e = 0.0
    for s in range(0, len(m), 4096):
        ph = np.exp(2j * np.pi * (xyz @ m[s:s + 4096].T))
        S = (q[:, None] * ph).sum(0)
        e += float((w[s:s + 4096] * (S * np.conj(S)).real).sum())
    return scaling * e / (2.0 * math.pi * volm)
