"""Setup-time table generation for the native (no Fortran host) mode.

Under the real DL_POLY host these arrays arrive through the C ABI (dlpgpu_set_vdw / dlpgpu_set_ewald) exactly as
the reference built them.  When the library is driven natively (bench.py, tests) somebody has to build them: this
module mirrors, on the host with numpy,

* ``vdw_generate``           vdw.F90:1397-1576   (analytic form sampled on i*dlrpot, i=1..max_grid, tab(0)=Huge)
* ``vdw_direct_fs_generate`` vdw.F90:969-1049    (force-shift constants afs/bfs)
* ``vdw_table_read``         vdw.F90:1051-1370   (TABLE file parsing + 3-point re-gridding)
* ``erfcgen``                electrostatic.F90:88-127 + numerics.F90:215-248,3647-3683 (A&S erfc polynomial)
* ``vdw_lrc``                vdw.F90:617-967     (long-range corrections elrc / vlrc)
* sizes/alpha                bounds.F90:811,820,907 ; control.F90:1709-1710

The same arrays (same bits) are handed to the GPU library and, in tests/bench, to the CPU oracle.
"""
import math

import numpy as np

# vdw.F90:64-117
VDW_NULL, VDW_TAB, VDW_12_6, VDW_LJ, VDW_BUCK, VDW_BHM = -1, 0, 1, 2, 4, 5
KEYPOT = {"tab": VDW_TAB, "12-6": VDW_12_6, "lj": VDW_LJ, "buck": VDW_BUCK, "bhm": VDW_BHM}

R4PIE0 = 138935.4835          # constants.F90:100
DELR_MAX = 0.01               # constants.F90:139
ZERO_PLUS = np.finfo(np.float64).tiny
HUGE = np.finfo(np.float64).max
PI = 4.0 * math.atan(1.0)
RSQRPI = 1.0 / math.sqrt(PI)


def f_nint(x):
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


def max_grid(rcut):
    """bounds.F90:811,820  Max(1004, Nint(rcut/delr_max)+4)."""
    return max(1004, f_nint(rcut / DELR_MAX) + 4)


def max_list(density, rx, densvar_factor=1.0):
    """bounds.F90:907  Nint(fdens*(7.5/3)*pi*rx**3)."""
    return f_nint(density * densvar_factor * (7.5 / 3.0) * PI * rx ** 3)


def ewald_alpha(precision, rcut):
    """control.F90:1709-1710."""
    tol = math.sqrt(abs(math.log(precision * rcut)))
    return math.sqrt(abs(math.log(precision * rcut * tol))) / rcut


def adjust_kmax(kmax, nproc=1):
    """parallel_fft.F90:1999-2054: the next multiple of nproc whose per-domain length is 2^a 3^b 5^c."""
    def ok(n):
        for p in (2, 3, 5):
            while n > 1 and n % p == 0:
                n //= p
        return n == 1
    if kmax % nproc:
        kmax = (kmax // nproc + 1) * nproc
    while not ok(kmax // nproc):
        kmax += nproc
    return kmax


def spme_grid(precision, rcut, cell, dims=(1, 1, 1)):
    """control.F90:1707-1713 + ewald.F90:1210-1212: (alpha, k_vec_dim) the reference derives from spme_precision; cell = 9
    doubles, rows = lattice vectors; dims = domains per direction."""
    a, b, c = np.asarray(cell, dtype=np.float64).reshape(3, 3)
    vol = abs(float(np.dot(a, np.cross(b, c))))
    widths = [vol / float(np.linalg.norm(np.cross(b, c))), vol / float(np.linalg.norm(np.cross(c, a))),
              vol / float(np.linalg.norm(np.cross(a, b)))]                      # dcell(7:9)
    tol = math.sqrt(abs(math.log(precision * rcut)))
    alpha = math.sqrt(abs(math.log(precision * rcut * tol))) / rcut
    tol1 = math.sqrt(-math.log(precision * rcut * (2.0 * tol * alpha) ** 2))
    k = [2 * f_nint(0.25 + w * alpha * tol1 / math.pi) for w in widths]
    return alpha, tuple(adjust_kmax(int(k[d]), int(dims[d])) for d in range(3))


def _powi(x, n):
    """libgcc __powidf2 ordering (what real**integer compiles to)."""
    m = abs(n)
    y = x.copy() if (m & 1) else np.ones_like(x)
    x = x.copy()
    m >>= 1
    while m:
        x = x * x
        if m & 1:
            y = y * x
        m >>= 1
    return 1.0 / y if n < 0 else y


def pot_energy(keypot, p, r):
    """(energy, gamma=-r dU/dr) -- two_body_potentials.F90:260-270,307-317,471-485,499-514.  p is 0-based."""
    r = np.asarray(r, dtype=np.float64)
    if keypot == VDW_12_6:
        r6 = _powi(1.0 / r, 6)
        return (p[0] * r6 - p[1]) * r6, 6.0 * r6 * (2.0 * p[0] * r6 - p[1])
    if keypot == VDW_LJ:
        s6 = _powi(p[1] / r, 6)
        return 4.0 * p[0] * s6 * (s6 - 1.0), 24.0 * p[0] * s6 * (2.0 * s6 - 1.0)
    if keypot == VDW_BUCK:
        b = r / p[1]
        t1 = p[0] * np.exp(-b)
        t2 = -p[2] / _powi(r, 6)
        return t1 + t2, t1 * b + 6.0 * t2
    if keypot == VDW_BHM:
        ri2 = _powi(r, -2)
        t1 = p[0] * np.exp(p[1] * (p[2] - r))
        t2 = -p[3] * _powi(ri2, 3)
        t3 = -p[4] * _powi(ri2, 4)
        return t1 + t2 + t3, (t1 * r * p[1] + 6.0 * t2 + 8.0 * t3)
    raise ValueError("analytic form %r not available in the native table generator" % (keypot,))


def vdw_generate(keypot, param, rvdw, mxgrid):
    """vdw.F90:1437-1457,1570-1572 for one potential -> (tab_potential(0:mxgrid), tab_force(0:mxgrid))."""
    dlrpot = rvdw / float(mxgrid - 4)
    r = np.arange(1, mxgrid + 1, dtype=np.float64) * dlrpot
    e, g = pot_energy(keypot, param, r)
    tp = np.empty(mxgrid + 1)
    tf = np.empty(mxgrid + 1)
    tp[1:] = e
    tf[1:] = g
    tp[0] = HUGE
    tf[0] = HUGE
    return tp, tf


def vdw_direct_fs(keypot, param, rvdw):
    """vdw.F90:1003-1046  afs = dz/rvdw, bfs = -z - dz."""
    z, dz = pot_energy(keypot, param, np.array([rvdw]))
    return float(dz[0]) / rvdw, -float(z[0]) - float(dz[0])


def vdw_lrc(ff, num_type, numfrz=None, imcon=1, volm=1.0):
    """vdw.F90:617-967: long-range corrections (elrc, vlrc) of a finalized ForceField; num_type / numfrz = atoms / frozen
    atoms per type over the whole system.  TABLE potentials carry theirs in param[0:2] (vdw.F90:1166-1170)."""
    if ff.force_shift or imcon in (0, 6):
        return 0.0, 0.0
    nt = np.asarray(num_type, dtype=np.float64)
    nf = np.zeros_like(nt) if numfrz is None else np.asarray(numfrz, dtype=np.float64)

    def pw(x, n):
        return float(_powi(np.array([float(x)]), n)[0])

    r = ff.rvdw
    r3, r5, r9 = pw(r, 3), pw(r, 5), pw(r, 9)
    elrc = plrc = 0.0
    ivdw = 0
    for i in range(1, ff.ntypes + 1):
        for j in range(1, i + 1):
            k = int(ff.vdw_list_c[ivdw]) - 1
            ivdw += 1
            p, key = ff.param[k], int(ff.ltp[k])
            eadd = padd = 0.0
            if key == VDW_TAB:
                eadd, padd = p[0], -p[1]
            elif key == VDW_12_6:
                eadd = p[0] / (9.0 * r9) - p[1] / (3.0 * r3)
                padd = 12.0 * p[0] / (9.0 * r9) - 6.0 * p[1] / (3.0 * r3)
            elif key == VDW_LJ:
                eadd = 4.0 * p[0] * (pw(p[1], 12) / (9.0 * r9) - pw(p[1], 6) / (3.0 * r3))
                padd = 8.0 * p[0] * (6.0 * pw(p[1], 12) / (9.0 * r9) - pw(p[1], 6) / r3)
            elif key == VDW_BUCK:
                eadd = -p[2] / (3.0 * r3)
                padd = -2.0 * p[2] / r3
            elif key == VDW_BHM:
                eadd = -p[3] / (3.0 * r3) - p[4] / (5.0 * r5)
                padd = -2.0 * p[3] / r3 - 8.0 * p[4] / (5.0 * r5)
            if i != j:
                eadd, padd = eadd * 2.0, padd * 2.0
            denprd = 2.0 * PI * (nt[i - 1] * nt[j - 1] - nf[i - 1] * nf[j - 1]) / pw(volm, 2)
            elrc = elrc + volm * denprd * eadd
            plrc = plrc + denprd * padd / 3.0
    return float(elrc), float(plrc * (-3.0 * volm))


def two_body_totals(out, elrc, vlrc, mxnode=1, spme=False, sumchg=0.0, alpha=0.0, eps=1.0, volm=1.0, engcpe_rc=0.0, vircpe_rc=0.0):
    """The end of two_body_forces (two_body.F90:672-790) for this path: ``out`` = the 16 partial sums of the C ABI AFTER the
    caller's gsum.  Returns (engcpe, vircpe, engsrp, virsrp, stress9): Fuchs' net-charge term (SPME, |sumchg| > 1e-6), the
    long-range corrections, and this rank's share of both on the stress diagonal."""
    engcpe_nz = vircpe_nz = 0.0
    if spme and abs(sumchg) > 1.0e-6:
        q = sumchg / alpha
        engcpe_nz = (-0.5 * (PI * R4PIE0 / eps) * (q * q)) / volm
        vircpe_nz = -3.0 * engcpe_nz
    engcpe = 0.0 + engcpe_rc + out[2] + 0.0 + out[4] + 0.0 + engcpe_nz
    vircpe = 0.0 + vircpe_rc + out[3] + 0.0 + out[5] + 0.0 + vircpe_nz + 0.0
    engsrp = 0.0 + (out[0] + elrc)
    virsrp = 0.0 + (out[1] + vlrc)
    stress = np.array(out[6:15], dtype=np.float64)
    for corr in (-vircpe_nz / (3.0 * float(mxnode)), -(vlrc + 0.0) / (3.0 * float(mxnode))):
        stress[0] += corr
        stress[4] += corr
        stress[8] += corr
    return engcpe, vircpe, engsrp, virsrp, stress


def erfc_as(x):
    """numerics.F90:3659-3665 Abramowitz-Stegun 5-term erfc."""
    a1, a2, a3, a4, a5, pp = 0.254829592, -0.284496736, 1.421413741, -1.453152027, 1.061405429, 0.3275911
    tt = 1.0 / (1.0 + pp * x)
    return tt * (a1 + tt * (a2 + tt * (a3 + tt * (a4 + tt * a5)))) * np.exp(-(x * x))


def erfcgen(rcut, alpha, nsamples=None):
    """electrostatic.F90:88-127.  Returns (erfc_tab, erfc_deriv_tab, recip_spacing); arrays have nsamples+1 entries
    with index i == Fortran table(i); index 0 is the reference's out-of-bounds slot (numerics.F90:235), set to 0."""
    n = max_grid(rcut) if nsamples is None else nsamples
    spacing = rcut / float(n - 4)
    recip = 1.0 / spacing
    r = np.arange(1, n + 1, dtype=np.float64) * spacing
    e = erfc_as(alpha * r) / r
    rsq = r * r
    d = (e + alpha * (2.0 * np.exp(-((alpha * r) * (alpha * r))) * RSQRPI)) / rsq
    et = np.zeros(n + 1)
    dt = np.zeros(n + 1)
    et[1:] = e
    dt[1:] = d
    return et, dt, recip


def regrid_table(buf, delpot, rvdw, mxgrid, is_force, engunit=1.0):
    """vdw.F90:1196-1341 for one array read from TABLE (``buf`` = the ngrid file values, 0-based)."""
    ngrid = len(buf)
    b = np.zeros(ngrid + 1)
    b[1:] = buf
    dlrpot = rvdw / float(mxgrid - 4)
    if abs(delpot - dlrpot) <= 1.0e-8:
        delpot = dlrpot
    remake = abs(1.0 - (delpot / dlrpot)) > 1.0e-8
    tab = np.zeros(mxgrid + 1)
    tab[0] = (2.0 * b[1] - 0.5 * b[2]) / delpot if is_force else 2.0 * b[1] - b[2]
    if remake:
        rdr = 1.0 / delpot
        for i in range(1, mxgrid - 3):
            rrr = float(i) * dlrpot
            l = int(rrr * rdr)
            ppp = rrr * rdr - float(l)
            vk = b[l]
            if l + 2 > ngrid:
                if l + 1 > ngrid:
                    vk1 = 2.0 * b[l] - b[l - 1]
                    vk2 = 2.0 * vk1 - b[l]
                else:
                    vk1 = b[l + 1]
                    vk2 = 2.0 * b[l + 1] - b[l]
            else:
                vk1 = b[l + 1]
                vk2 = b[l + 2]
            t1 = vk + (vk1 - vk) * ppp
            t2 = vk1 + (vk2 - vk1) * (ppp - 1.0)
            tab[i] = t1 + (t2 - t1) * ppp * 0.5
    else:
        tab[1:mxgrid - 3] = b[1:mxgrid - 3]
        tab[mxgrid - 3] = 2.0 * tab[mxgrid - 4] - tab[mxgrid - 5]
    tab[mxgrid - 2] = 2.0 * tab[mxgrid - 3] - tab[mxgrid - 4]
    if (not is_force) and abs(tab[0]) <= ZERO_PLUS:
        tab[0] = math.copysign(ZERO_PLUS, tab[0])
    return tab * engunit


def write_table_file(path, pairs, delpot, cutpot, ngrid, title="TABLE generated by dl-poly_b200"):
    """DL_POLY TABLE format (vdw.F90:1092-1265): header, 'delpot cutpot ngrid', then per pair a label record
    'atom1 atom2 elrc vlrc' followed by ngrid potential values and ngrid force values, four per line."""
    with open(path, "w") as f:
        f.write(title + "\n")
        f.write("%20.12e %20.12e %10d\n" % (delpot, cutpot, ngrid))
        for (a1, a2, elrc, vlrc, pot, frc) in pairs:
            f.write("%-8s%-8s %20.12e %20.12e\n" % (a1, a2, elrc, vlrc))
            for arr in (pot, frc):
                for i in range(0, ngrid, 4):
                    f.write(" ".join("%24.16e" % v for v in arr[i:i + 4]) + "\n")


def read_table_file(path):
    """Parses what write_table_file wrote -> (delpot, cutpot, ngrid, [(a1,a2,elrc,vlrc,pot[ngrid],frc[ngrid])])."""
    with open(path) as f:
        f.readline()
        w = f.readline().split()
        delpot, cutpot, ngrid = float(w[0]), float(w[1]), f_nint(float(w[2]))
        out = []
        while True:
            line = f.readline()
            if not line.strip():
                break
            w = line.split()
            a1, a2, elrc, vlrc = w[0], w[1], float(w[2]), float(w[3])
            arrs = []
            for _ in range(2):
                vals = []
                while len(vals) < ngrid:
                    vals.extend(float(t) for t in f.readline().split())
                arrs.append(np.array(vals[:ngrid]))
            out.append((a1, a2, elrc, vlrc, arrs[0], arrs[1]))
    return delpot, cutpot, ngrid, out


class ForceField:
    """What read_field leaves behind for the pair path (ffield.F90:3620-3960, 4317-4318): pair -> potential map,
    tables, force-shift constants, Ewald tables.  Arrays are laid out exactly as the C ABI expects them."""

    def __init__(self, ntypes, rvdw, rcut, force_shift=False, direct=False):
        self.ntypes = ntypes
        self.rvdw = float(rvdw)
        self.rcut = float(rcut)
        self.force_shift = bool(force_shift)
        self.direct = bool(direct)
        self.ntab = ntypes * (ntypes + 1) // 2
        self.vdw_list = np.zeros(self.ntab, dtype=np.int32)        # vdws%list(key), 1-based k, 0 = unset
        self.pots = []                                              # [(keypot, param7)]
        self.mxgrid = max_grid(self.rvdw)
        self.ew_active = False
        self.coul_kind = 0
        self.coul_damp = False
        self.alpha = 0.0
        self.scaling = 0.0
        self.eps = 1.0
        self.table_arrays = {}                                      # k -> (tp, tf) for VDW_TAB

    @staticmethod
    def key(ai, aj):
        """vdw.F90:1875-1879 (1-based types)."""
        return (max(ai, aj) * (max(ai, aj) - 1)) // 2 + min(ai, aj)

    def add(self, ai, aj, form, params):
        keypot = KEYPOT[form] if isinstance(form, str) else int(form)
        p = np.zeros(7)
        p[:len(params)] = params
        self.pots.append((keypot, p))
        k = len(self.pots)
        kk = self.key(ai, aj)
        if self.vdw_list[kk - 1] != 0:
            raise ValueError("error 15: duplicate vdw pair")     # ffield.F90:3868
        self.vdw_list[kk - 1] = k
        return k

    def add_table(self, ai, aj, tp, tf):
        k = self.add(ai, aj, VDW_TAB, [])
        self.table_arrays[k] = (np.asarray(tp, dtype=np.float64), np.asarray(tf, dtype=np.float64))
        return k

    def set_ewald(self, precision=None, alpha=None, eps=1.0):
        self.ew_active = True
        self.eps = eps
        self.alpha = ewald_alpha(precision, self.rcut) if alpha is None else float(alpha)
        self.scaling = R4PIE0 / eps                                # two_body.F90:188

    COUL_KINDS = {"coul": 1, "dddp": 2, "fscp": 3, "rfp": 4}   # coul_cp / coul_dddp / coul_fscp / coul_rfp_forces

    def set_coulomb(self, kind, eps=1.0, damping=0.0):
        """The direct-space Coulomb variants of coul_spole.F90 (two_body.F90:480-514).  damping > 0 (fscp, rfp only) is the
        Fennell-Gezelter damped form through erfc tables generated with alpha = damping (coul_spole.F90:186-202, :407-417)."""
        self.ew_active = False
        self.coul_kind = self.COUL_KINDS[kind]
        self.coul_damp = damping > 0.0 and self.coul_kind in (3, 4)
        self.eps = float(eps)
        self.scaling = R4PIE0 / eps
        self.alpha = float(damping)
        rc = self.rcut
        self.coul_force_shift, self.coul_energy_shift = 0.0, 0.0
        self.coul_rf = np.zeros(3)
        if self.coul_kind == 4:
            b0 = 2.0 * (eps - 1.0) / (2.0 * eps + 1.0)
            self.coul_rf[0] = b0 / rc ** 3
            self.coul_rf[1] = (1.0 + 0.5 * b0) / rc
            self.coul_rf[2] = 0.5 * self.coul_rf[0]
        if self.coul_damp:
            self.erfc, self.erfc_deriv, self.ew_recip = erfcgen(rc, self.alpha)
            self.ew_n = len(self.erfc) - 1
            self.coul_force_shift = self.erfc_deriv[self.ew_n - 4] * rc            # end_sample = table(nsamples - 4), numerics.F90:245
            self.coul_energy_shift = -(self.erfc[self.ew_n - 4] + self.coul_force_shift * rc)
        elif self.coul_kind == 3:
            self.coul_force_shift = 1.0 / rc ** 2
            self.coul_energy_shift = -2.0 / rc

    def finalize(self):
        n_vdw = len(self.pots)
        self.n_vdw = n_vdw
        # ffield.F90:3939-3954: undefined pairs point past the defined range with ltp = VDW_NULL
        self.max_vdw = n_vdw + 1 if n_vdw < self.ntab else max(n_vdw, 1)
        lst = self.vdw_list.copy()
        lst[lst == 0] = n_vdw + 1
        self.vdw_list_c = np.ascontiguousarray(lst, dtype=np.int32)
        self.ltp = np.full(self.max_vdw, VDW_NULL, dtype=np.int32)
        self.param = np.zeros((self.max_vdw, 7))
        self.afs = np.zeros(self.max_vdw)
        self.bfs = np.zeros(self.max_vdw)
        g = self.mxgrid
        self.tab_potential = np.zeros((self.max_vdw, g + 1))      # C order [k][i] == Fortran (0:g, 1:max_vdw)
        self.tab_force = np.zeros((self.max_vdw, g + 1))
        for k, (keypot, p) in enumerate(self.pots):
            self.ltp[k] = keypot
            self.param[k] = p
            if keypot == VDW_TAB:
                tp, tf = self.table_arrays[k + 1]
                if self.force_shift:                               # vdw.F90:1343-1352
                    tp = tp.copy(); tf = tf.copy()
                    tp[g - 3] = tp[g - 2] = 0.0
                    tf[g - 3] = tf[g - 2] = 0.0
            elif self.direct and keypot not in (VDW_12_6, VDW_LJ, VDW_BUCK, VDW_BHM):
                # vdw_method direct never reads the tables; the forms this host generator does not tabulate are evaluated
                # per pair on the device (forces.cu::pot_direct_any)
                if self.force_shift:
                    raise ValueError("force-shifted direct evaluation of form %d is not available in the native host" % keypot)
                tp, tf = np.zeros(g + 1), np.zeros(g + 1)
                tp[0] = tf[0] = HUGE
            else:
                tp, tf = vdw_generate(keypot, p, self.rvdw, g)
                if self.force_shift and self.direct:
                    self.afs[k], self.bfs[k] = vdw_direct_fs(keypot, p, self.rvdw)
            if abs(tp[0]) <= ZERO_PLUS:
                tp[0] = math.copysign(ZERO_PLUS, tp[0])
            self.tab_potential[k] = tp
            self.tab_force[k] = tf
        if self.ew_active:
            self.erfc, self.erfc_deriv, self.ew_recip = erfcgen(self.rcut, self.alpha)
            self.ew_n = len(self.erfc) - 1
        return self
