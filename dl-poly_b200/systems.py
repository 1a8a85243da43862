"""Synthetic CONFIG/FIELD/CONTROL-equivalent inputs for the BASELINE.json configs.

Each generator returns a :class:`System`: the cell (DL_POLY row-major lattice vectors, origin-centred box as
input_files.rst:866-868 requires), positions in CONFIG order, per-atom site index, the site tables
(site.F90: type_site, charge_site, freeze_site, weight_site), a :class:`tables.ForceField`, cutoffs and -- for
molecular systems -- the exclusion table build_excl_intra would produce (sorted global ids per atom,
build_excl.F90:1181-1184).
"""
import numpy as np

from . import tables


class System:
    def __init__(self, name, cell, xyz, lsite, type_site, charge_site, weight_site, ff, rcut, padding,
                 freeze_site=None, excl=None, temperature=0.0, seed=0):
        self.name = name
        self.cell = np.ascontiguousarray(cell, dtype=np.float64).reshape(9)
        offdiag = np.abs(self.cell[[1, 2, 3, 5, 6, 7]]).max()
        # imcon of the CONFIG header: 1 cubic, 2 orthorhombic, 3 parallelepiped
        self.imcon = 3 if offdiag > 0.0 else (1 if (self.cell[0] == self.cell[4] == self.cell[8]) else 2)
        self.xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        self.megatm = self.xyz.shape[0]
        self.lsite = np.ascontiguousarray(lsite, dtype=np.int32)            # 1-based site index per atom
        self.type_site = np.ascontiguousarray(type_site, dtype=np.int32)    # 1-based type per site
        self.charge_site = np.ascontiguousarray(charge_site, dtype=np.float64)
        self.weight_site = np.ascontiguousarray(weight_site, dtype=np.float64)
        self.freeze_site = (np.zeros(len(type_site), dtype=np.int32) if freeze_site is None
                            else np.ascontiguousarray(freeze_site, dtype=np.int32))
        self.ff = ff
        self.rcut = float(rcut)
        self.padding = float(padding)
        self.rx = self.rcut + self.padding
        self.pdplnc = 50.0                                                   # neighbours.F90:84
        self.excl = excl                                                     # (megatm, max_exclude+1) int32 or None
        self.max_exclude = 0 if excl is None else excl.shape[1] - 1
        self.volume = float(abs(np.linalg.det(self.cell.reshape(3, 3))))
        self.density = self.megatm / self.volume
        self.max_list = min(tables.max_list(self.density, self.rx), self.megatm - 1)   # bounds.F90:907-908
        self.weight_by_type = np.ones(int(self.type_site.max()))
        for s, t in enumerate(self.type_site):
            self.weight_by_type[t - 1] = self.weight_site[s]
        self.vel = None
        if temperature > 0.0:
            rng = np.random.default_rng(seed + 7)
            boltz = 8.31446261815324e-1                                      # kB in 10 J/mol/K (constants.F90)
            m = self.weight_site[self.lsite - 1]
            self.vel = rng.standard_normal((self.megatm, 3)) * np.sqrt(boltz * temperature / m)[:, None]
            self.vel -= (self.vel * m[:, None]).sum(0) / m.sum()
        self.megfrz = int((self.freeze_site[self.lsite - 1] > 0).sum())

    @property
    def lbook(self):
        return self.excl is not None


def _n3(ncell):
    """ncell may be an int (cubic box) or a 3-tuple of lattice repeats (orthorhombic box, used for weak scaling)."""
    if np.isscalar(ncell):
        return (int(ncell),) * 3
    return tuple(int(v) for v in ncell)


def _fcc(ncell, a):
    base = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]])
    nx, ny, nz = _n3(ncell)
    g = np.stack(np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij"), -1).reshape(-1, 3)
    return ((g[:, None, :] + base[None, :, :]).reshape(-1, 3)) * a


def _wrap(xyz, L):
    s = xyz / L
    s -= np.rint(s)
    s[s >= 0.5] = -0.5
    return s * L


def argon(ncell=20, seed=1001, rcut=8.5, padding=0.3, form="12-6", direct=False, force_shift=False, jitter=0.25,
          temperature=85.0):
    """C1: Ar 12-6 fluid, 4*ncell^3 atoms (32,000 at ncell=20), rho=0.02138 A^-3."""
    nc = np.array(_n3(ncell), dtype=np.float64)
    n = int(4 * nc.prod())
    a = (4.0 / 0.02138) ** (1.0 / 3.0)
    L = nc * a
    rng = np.random.default_rng(seed)
    xyz = _fcc(ncell, a) - 0.5 * L + 0.25 * a
    xyz = _wrap(xyz + rng.uniform(-jitter, jitter, xyz.shape), L)
    eps, sig = 99.61, 3.405                                   # 0.9961 kJ/mol in 10 J/mol units
    ff = tables.ForceField(1, rcut, rcut, force_shift=force_shift, direct=direct)
    if form == "12-6":
        ff.add(1, 1, "12-6", [4 * eps * sig ** 12, 4 * eps * sig ** 6])
    else:
        ff.add(1, 1, "lj", [eps, sig])
    ff.finalize()
    cell = np.diag(L)
    return System("C1-argon-%d" % n, cell, xyz, np.ones(n, dtype=np.int32), [1], [0.0], [39.948], ff, rcut, padding,
                  temperature=temperature, seed=seed)


# One sane parameter set per analytic form of two_body_potentials.F90 (keys: vdw.F90:71-113) for an argon-like fluid whose
# nearest neighbours sit at ~3.3 A and whose cutoff is 8.5 A: used by the parity tests of vdw_method direct.
_S12, _S6 = 3.4 ** 12, 3.4 ** 6
VDW_DIRECT_PARAMS = {
    1: [4.0 * _S12, 4.0 * _S6], 2: [1.0, 3.4], 3: [1.0, 12.0, 6.0, 3.8], 4: [1.0e5, 0.3, 1.0e3], 5: [1.0e3, 3.0, 3.0, 5.0e2, 1.0e3],
    6: [1.0e6, 1.0e5], 7: [1.0, 12.0, 6.0, 3.8, 8.0], 8: [1.0, 3.8, 1.5], 9: [1.0, 3.4, 0.5, 2.0 ** (1.0 / 6.0) * 3.4 + 0.5],
    10: [25.0, 6.0], 11: [1.0, 3.8], 12: [1.0, 3.4, 0.8], 13: [1.0, 3.8, 1.5, 1.0e5], 14: [100.0, -20.0, 1.2], 15: [18.0, 18.0],
    16: [18.0, 18.0, 3.5, 0.25, 1.0, 3.8, 1.5], 17: [18.0, 18.0, 3.5, 0.25, 1.0e5, 0.3, 1.0e3], 18: [1.0, 3.4, 6.0, 8.0],
    19: [1.0e5, 0.3, 1.0e3, 6.0, 8.0], 20: [4.0 * _S12, 4.0 * _S6, 6.0, 8.0], 21: [1.0, 3.4 ** 2, 64.0], 22: [1.0, 4.0, 1.5],
    23: [25.0, 2.0, 3.0, 6.0], 24: [1.0, 7.05, 0.602, 3.4, 4.0, 0.0, 1.8],
}


def vdw_direct_fluid(keypot, ncell=6, seed=1012, rcut=8.5, padding=0.3, jitter=0.25, temperature=85.0):
    """An argon-density fluid whose single vdW potential is analytic form `keypot` (1..24) evaluated with vdw_method direct."""
    nc = np.array(_n3(ncell), dtype=np.float64)
    n = int(4 * nc.prod())
    a = (4.0 / 0.02138) ** (1.0 / 3.0)
    L = nc * a
    rng = np.random.default_rng(seed)
    xyz = _wrap(_fcc(ncell, a) - 0.5 * L + 0.25 * a + rng.uniform(-jitter, jitter, (n, 3)), L)
    ff = tables.ForceField(1, rcut, rcut, direct=True)
    ff.add(1, 1, int(keypot), VDW_DIRECT_PARAMS[int(keypot)])
    ff.finalize()
    return System("direct-key%d-%d" % (keypot, n), np.diag(L), xyz, np.ones(n, dtype=np.int32), [1], [0.0], [39.948], ff, rcut,
                  padding, temperature=temperature, seed=seed)


def argon_triclinic(ncell=6, seed=1011, rcut=8.5, padding=0.3, shear=(0.20, 0.10, 0.15), jitter=0.25, temperature=85.0):
    """The C1 fluid in a parallelepiped cell (imcon = 3): the cubic lattice sheared by b += shear[0] a, c += shear[1] a +
    shear[2] b.  Positions are cell . s with s the reduced coordinates of the jittered fcc sites."""
    nc = np.array(_n3(ncell), dtype=np.float64)
    n = int(4 * nc.prod())
    a = (4.0 / 0.02138) ** (1.0 / 3.0)
    L = nc * a
    rng = np.random.default_rng(seed)
    xyz = _fcc(ncell, a) - 0.5 * L + 0.25 * a
    sred = _wrap(xyz + rng.uniform(-jitter, jitter, xyz.shape), L) / L            # reduced coordinates in [-0.5, 0.5)
    cell = np.array([[L[0], 0.0, 0.0], [shear[0] * L[0], L[1], 0.0], [shear[1] * L[0], shear[2] * L[1], L[2]]])
    xyz = sred @ cell                                                             # r = s_a a + s_b b + s_c c  (rows = lattice vectors)
    eps, sig = 99.61, 3.405
    ff = tables.ForceField(1, rcut, rcut)
    ff.add(1, 1, "12-6", [4 * eps * sig ** 12, 4 * eps * sig ** 6])
    ff.finalize()
    return System("argon-triclinic-%d" % n, cell.reshape(9), xyz, np.ones(n, dtype=np.int32), [1], [0.0], [39.948], ff, rcut, padding,
                  temperature=temperature, seed=seed)


# Fumi-Tosi NaCl, internal units (10 J/mol): A, B(1/A), sigma(A), C, D
_BHM = {(1, 1): [2544.35, 3.1545, 2.3400, 1.0117e4, 4.8177e3],
        (1, 2): [2035.48, 3.1545, 2.7550, 6.7448e4, 8.3708e4],
        (2, 2): [1526.61, 3.1545, 3.1700, 6.9857e5, 1.4032e6]}


def _rocksalt(ncell, a, seed, jitter):
    nx, ny, nz = _n3(ncell)
    L = np.array([nx, ny, nz], dtype=np.float64) * a
    g = np.stack(np.meshgrid(np.arange(2 * nx), np.arange(2 * ny), np.arange(2 * nz), indexing="ij"),
                 -1).reshape(-1, 3)
    species = (g.sum(1) % 2).astype(np.int32) + 1             # 1 = Na+, 2 = Cl-
    rng = np.random.default_rng(seed)
    xyz = g * (0.5 * a) - 0.5 * L + 0.125 * a
    xyz = _wrap(xyz + rng.uniform(-jitter, jitter, xyz.shape), L)
    return xyz, species


def nacl(ncell=15, seed=1002, rcut=12.0, padding=0.24, tabfile=False, direct=False, force_shift=False, ewald=True,
         jitter=0.3, temperature=1200.0, spme_precision=1.0e-6, rvdw=None, vdw_pairs=((1, 1), (1, 2), (2, 2)),
         coulomb=None, eps=1.0, damping=0.0):
    """C2/C4/C5-ionic: molten NaCl, 8*ncell^3 ions at the TEST01 density (V=963,882.2 A^3 for 27,000 ions).
    tabfile=True builds the three pair tables through a TABLE-format round trip (C4)."""
    nc = np.array(_n3(ncell), dtype=np.float64)
    n = int(8 * nc.prod())
    a = (963882.2 * 8.0 / 27000.0) ** (1.0 / 3.0)             # rock-salt cell edge at the TEST01 density
    L = nc * a
    xyz, species = _rocksalt(ncell, a, seed, jitter)
    # rvdw < rcut (control.F90:1499-1543) gives the vdW and Ewald tables different grids; vdw_pairs selects which type
    # pairs carry a vdW potential at all (the others interact through the Ewald term only)
    ff = tables.ForceField(2, rcut if rvdw is None else rvdw, rcut, force_shift=force_shift, direct=direct)
    if tabfile:
        g = ff.mxgrid
        ngrid = g                                             # = max(1004, nint(rcut/0.01)+4): no re-gridding
        delpot = rcut / float(g - 4)
        r = np.arange(1, ngrid + 1, dtype=np.float64) * delpot
        for (ai, aj), p in _BHM.items():
            if (ai, aj) not in vdw_pairs:
                continue
            e, gm = tables.pot_energy(tables.VDW_BHM, p, r)
            tp = tables.regrid_table(e, delpot, rcut, g, is_force=False)
            tf = tables.regrid_table(gm, delpot, rcut, g, is_force=True)
            ff.add_table(ai, aj, tp, tf)
    else:
        for (ai, aj), p in _BHM.items():
            if (ai, aj) in vdw_pairs:
                ff.add(ai, aj, "bhm", p)
    if coulomb is not None:       # "coul" | "dddp" | "fscp" | "rfp": the direct-space variants of coul_spole.F90 instead of Ewald
        ff.set_coulomb(coulomb, eps=eps, damping=damping)
    elif ewald:
        ff.set_ewald(precision=spme_precision)
    ff.finalize()
    cell = np.diag(L)
    return System("NaCl-%d%s" % (n, "-TABLE" if tabfile else ""), cell, xyz, species, [1, 2], [1.0, -1.0],
                  [22.9898, 35.453], ff, rcut, padding, temperature=temperature, seed=seed)


def ionic_mixture(ncell=4, ntypes=4, seed=1010, rcut=8.0, padding=0.2, jitter=0.3, temperature=800.0, spme_precision=1.0e-6):
    """A rock-salt lattice with ``ntypes`` species (charges alternate +1 / -1) and a tabulated 12-6 potential for EVERY type
    pair (ntypes (ntypes + 1) / 2 tables) next to real-space Ewald: the force field with many potentials."""
    nc = np.array(_n3(ncell), dtype=np.float64)
    n = int(8 * nc.prod())
    a = 6.2
    L = nc * a
    xyz, parity = _rocksalt(ncell, a, seed, jitter)
    rng = np.random.default_rng(seed + 1)
    half = ntypes // 2
    species = np.where(parity == 1, 1 + 2 * rng.integers(0, ntypes - half, n), 2 + 2 * rng.integers(0, half, n)).astype(np.int32)
    ff = tables.ForceField(ntypes, rcut, rcut)
    for i in range(1, ntypes + 1):
        for j in range(i, ntypes + 1):
            sig = 2.2 + 0.15 * (i + j)
            eps = 40.0 + 7.0 * i * j
            ff.add(i, j, "12-6", [4 * eps * sig ** 12, 4 * eps * sig ** 6])
    ff.set_ewald(precision=spme_precision)
    ff.finalize()
    q = [1.0 if t % 2 == 1 else -1.0 for t in range(1, ntypes + 1)]
    w = [20.0 + 3.0 * t for t in range(1, ntypes + 1)]
    return System("mixture-%d-%dtypes" % (n, ntypes), np.diag(L), xyz, species, list(range(1, ntypes + 1)), q, w, ff, rcut, padding,
                  temperature=temperature, seed=seed)


def spce_water(nmol=72000, seed=1003, rcut=9.0, padding=0.18, spme_precision=1.0e-6, temperature=0.0, coulomb=None, eps=1.0,
               damping=0.0, rvdw=None):
    """C3: SPC/E water, 3*nmol atoms, rigid geometry 1.0 A / 109.47 deg, random orientations, O-O LJ, exclusions =
    the two other atoms of the molecule."""
    L = (nmol / 0.0334) ** (1.0 / 3.0)
    nside = int(np.ceil(nmol ** (1.0 / 3.0)))
    rng = np.random.default_rng(seed)
    g = np.stack(np.meshgrid(np.arange(nside), np.arange(nside), np.arange(nside), indexing="ij"), -1).reshape(-1, 3)
    g = g[rng.permutation(len(g))[:nmol]]
    g = g[np.lexsort((g[:, 2], g[:, 1], g[:, 0]))]
    com = (g + 0.5) * (L / nside) - 0.5 * L
    q = rng.standard_normal((nmol, 4))
    q /= np.linalg.norm(q, axis=1)[:, None]
    w, x, y, z = q.T
    R = np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
                  np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
                  np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], 1)
    th = np.deg2rad(109.47) / 2
    body = np.array([[0.0, 0.0, 0.0], [np.sin(th), 0.0, np.cos(th)], [-np.sin(th), 0.0, np.cos(th)]])
    xyz = (com[:, None, :] + np.einsum("mij,aj->mai", R, body)).reshape(-1, 3)
    xyz = _wrap(xyz, L)
    n = 3 * nmol
    lsite = np.tile(np.array([1, 2, 3], dtype=np.int32), nmol)
    ff = tables.ForceField(2, rcut if rvdw is None else rvdw, rcut)      # rvdw < rcut: control.F90:1499-1543
    ff.add(1, 1, "lj", [65.0, 3.166])
    if coulomb is not None:
        ff.set_coulomb(coulomb, eps=eps, damping=damping)
    else:
        ff.set_ewald(precision=spme_precision)
    ff.finalize()
    gid = np.arange(1, n + 1, dtype=np.int32).reshape(nmol, 3)
    excl = np.zeros((n, 3), dtype=np.int32)
    excl[:, 0] = 2
    for a in range(3):
        others = np.sort(np.delete(gid, a, axis=1), axis=1)
        excl[a::3, 1:] = others
    cell = np.diag([L, L, L])
    return System("SPCE-%d" % n, cell, xyz, lsite, [1, 2, 2], [-0.8476, 0.4238, 0.4238], [15.9994, 1.008, 1.008],
                  ff, rcut, padding, excl=excl, temperature=temperature, seed=seed)


def by_name(name, **kw):
    """BASELINE.json configs by short name."""
    if name == "c1":
        return argon(20, **kw)
    if name == "c2":
        return nacl(15, **kw)
    if name == "c3":
        return spce_water(72000, **kw)
    if name == "c4":
        return nacl(50, seed=1004, tabfile=True, **kw)
    if name == "c5_ionic":
        return nacl(100, seed=1005, **kw)
    if name == "c5_lj":
        return argon(126, seed=1005, **kw)
    if name == "ionic_1m":
        return nacl(50, seed=1005, **kw)
    if name == "lj_1m":
        return argon(63, seed=1005, **kw)
    raise KeyError(name)
