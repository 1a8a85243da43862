"""Pins oracle/spme_oracle.py (numpy restatement of the reference's SPME reciprocal-space Coulomb path, CPU only): against the
exact Ewald reciprocal sum to the accuracy SPME has, and its forces against the finite-difference derivative of its own energy."""
import numpy as np
import pytest

from dl_poly_b200 import systems, tables
from oracle import spme_oracle as so


def _nacl(n=3, jitter=0.3, triclinic=False):
    s = systems.nacl(n, rcut=8.0, padding=0.2, jitter=jitter)
    q = s.charge_site[s.lsite - 1]
    cell = s.cell.copy()
    return s, cell, s.xyz.copy(), q


def test_grid_selection_follows_control_and_adjust_kmax():
    assert [so.adjust_kmax(k) for k in (7, 14, 22, 97, 128)] == [8, 15, 24, 100, 128]
    assert so.adjust_kmax(22, 4) == 24 and so.adjust_kmax(26, 4) == 32      # kmax / P has to be 2^a 3^b 5^c
    alpha, kd = so.spme_grid(1.0e-6, 12.0, np.diag([100.0, 100.0, 100.0]).reshape(9))
    assert alpha == tables.ewald_alpha(1.0e-6, 12.0) and kd == (54, 54, 54)


def test_bsplines_partition_of_unity_and_derivative():
    rng = np.random.default_rng(3)
    u = rng.uniform(0.0, 30.0, (50, 3))
    for n in (4, 6, 8):
        idx, d0, d1 = so.bspline_splines(u, n)
        assert np.allclose(d0.sum(2), 1.0, atol=1e-14) and np.allclose(d1.sum(2), 0.0, atol=1e-13)
        h = 1e-6
        _, dp, _ = so.bspline_splines(u + h, n); _, dm, _ = so.bspline_splines(u - h, n)
        same = (np.trunc(u + h) == np.trunc(u - h))
        assert np.allclose(((dp - dm) / (2 * h))[same], d1[same], atol=1e-7)


@pytest.mark.parametrize("nspl,prec,tol", [(8, 1.0e-6, 2.0e-5), (10, 1.0e-8, 2.0e-7)])
def test_spme_energy_against_the_exact_reciprocal_sum(nspl, prec, tol):
    s, cell, xyz, q = _nacl(3)
    alpha, kdim = so.spme_grid(prec, s.rcut, cell)
    kdim = tuple(so.adjust_kmax(2 * k) for k in kdim)             # a finer grid than the default: the pin is about the algebra
    r = so.ewald_spme_forces_coul(cell, xyz, q, alpha, kdim, nspl, s.ff.scaling)
    exact = so.ewald_recip_exact(cell, xyz, q, alpha, s.ff.scaling)
    assert abs(r["eng_recip"] - exact) <= tol * abs(exact), (r["eng_recip"], exact)
    # virial theorem of the reciprocal sum: -trace(stress) = vircpe_rc, and for a cubic box dE/dV by finite differences
    assert abs(r["vircpe_rc"] + r["stress"][0::4].sum()) <= 1e-12 * abs(r["vircpe_rc"])


def test_spme_forces_are_the_gradient_of_the_spme_energy():
    s, cell, xyz, q = _nacl(2)
    alpha, kdim = so.spme_grid(1.0e-6, s.rcut, cell)
    kdim = tuple(so.adjust_kmax(2 * k) for k in kdim)
    r = so.ewald_spme_forces_coul(cell, xyz, q, alpha, kdim, 8, s.ff.scaling)
    f = r["forces"]
    assert np.abs(f.sum(0)).max() <= 1e-9 * np.abs(f).max()        # the net force is removed (ewald_general.F90:862-866)
    h = 1.0e-5
    for a, d in ((0, 0), (5, 1), (11, 2)):
        xp, xm = xyz.copy(), xyz.copy()
        xp[a, d] += h; xm[a, d] -= h
        ep = so.ewald_spme_forces_coul(cell, xp, q, alpha, kdim, 8, s.ff.scaling)["eng_recip"]
        em = so.ewald_spme_forces_coul(cell, xm, q, alpha, kdim, 8, s.ff.scaling)["eng_recip"]
        fd = -(ep - em) / (2 * h)
        # the analytic force misses only the (tiny) net-force correction
        assert abs(fd - f[a, d]) <= 2e-6 * np.abs(f).max(), (fd, f[a, d])


def test_spme_stress_against_the_volume_derivative():
    s, cell, xyz, q = _nacl(2)
    alpha, kdim = so.spme_grid(1.0e-6, s.rcut, cell)
    kdim = tuple(so.adjust_kmax(2 * k) for k in kdim)
    r = so.ewald_spme_forces_coul(cell, xyz, q, alpha, kdim, 8, s.ff.scaling)
    # isotropic scaling of cell and coordinates: dE/d(ln V) = -(trace of the stress) / 3 ... with alpha and the grid fixed
    h = 1.0e-5
    ep = so.ewald_spme_forces_coul(cell * (1 + h), xyz * (1 + h), q, alpha, kdim, 8, s.ff.scaling)["eng_recip"]
    em = so.ewald_spme_forces_coul(cell * (1 - h), xyz * (1 - h), q, alpha, kdim, 8, s.ff.scaling)["eng_recip"]
    dE_dlnL = (ep - em) / (2 * h)
    assert abs(dE_dlnL + r["stress"][0::4].sum()) <= 5e-5 * abs(r["stress"][0::4].sum()), (dE_dlnL, r["stress"][0::4].sum())


def test_host_grid_selection_agrees_with_the_oracle():
    rng = np.random.default_rng(11)
    for _ in range(50):
        L = rng.uniform(20.0, 400.0, 3)
        cell = np.diag(L).reshape(9)
        if rng.random() < 0.3:
            cell[3] = 0.2 * L[0]; cell[7] = 0.1 * L[1]
        prec = 10.0 ** rng.uniform(-8, -4); rcut = rng.uniform(6.0, 14.0)
        dims = tuple(int(x) for x in rng.choice([1, 2, 3, 4], 3))
        assert tables.spme_grid(prec, rcut, cell, dims) == so.spme_grid(prec, rcut, cell, dims)


@pytest.mark.parametrize("name", ["spme_nacl_512_order8", "spme_water_1536_order6"])
def test_golden_spme_fixture(name):
    """tests/golden/spme_*.json (tests/golden/make_golden_spme.py) freeze the restatement's answers."""
    import json, os
    from dl_poly_b200 import dd
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", name + ".json")))
    s = getattr(systems, g["generator"])(**g["kwargs"])
    xyz = dd.read_config_fold(s.xyz, s.cell)[0]
    q = s.charge_site[s.lsite - 1]
    alpha, kdim = so.spme_grid(1.0e-6, s.rcut, s.cell)
    assert alpha == g["alpha"] and list(kdim) == g["kdim"]
    r = so.ewald_spme_forces_coul(s.cell, xyz, q, alpha, kdim, g["nspl"], s.ff.scaling)
    for k in ("engcpe_rc", "vircpe_rc", "eng_recip"):
        assert abs(r[k] - g[k]) <= 1e-12 * abs(g[k]), k
    assert np.abs(r["stress"] - np.array(g["stress"])).max() <= 1e-12 * np.abs(g["stress"]).max()
    assert abs(np.abs(r["forces"]).sum() - g["force_l1"]) <= 1e-11 * g["force_l1"]
    assert np.abs(r["forces"][0] - np.array(g["force_first"])).max() <= 1e-10 * np.abs(r["forces"]).max()
    assert np.abs(r["forces"][-1] - np.array(g["force_last"])).max() <= 1e-10 * np.abs(r["forces"]).max()
