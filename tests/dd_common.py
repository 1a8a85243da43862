"""Multi-rank parity check shared by tests/dd_check.py (one PROCESS per GPU under torch.distributed.run), by
tests/test_gpu_threads.py (ranks = threads of one process, several domains on ONE GPU) and by bench.py's pre-flight: the
dd.Domain engine against the CPU oracle's P-domain world on the same seeded system --

* resident atoms after relocation + halo build: same atoms, same order, same bits (positions, charges, ltg, lsite, ltype, ixyz);
* forces per atom within 1e-9 (north-star bar, see util.per_atom_force_error), the six energy / virial sums within 1e-10 after
  the gsum over the ranks;
* the one-kernel peer-memory halo refresh against the staged exchange (bit for bit) and the oracle;
* a trajectory with padding-driven rebuilds during which atoms MIGRATE between domains: same rebuild decisions, same atom
  counts per rank, energies tracked;
* the library-enqueued step (dlpgpu_dev_md_step): its sums arrive already reduced over the ranks (the gsum rides on the gmax
  message) and equal the oracle's gsum.

TEST INFRASTRUCTURE: imports the oracle.
"""
import numpy as np

from dl_poly_b200 import dd, systems
from oracle import oracle as ora

FORCE_TOL, ENERGY_TOL = 1.0e-9, 1.0e-10


def make_system(which):
    if which == "nacl":
        return systems.nacl((8, 8, 8), rcut=8.0, padding=0.3, temperature=1200.0)
    if which == "nacl_hot":      # many atoms cross the faces between rebuilds: several holes / tail movers per migration stage
        return systems.nacl((12, 12, 12), rcut=6.0, padding=0.5, temperature=6000.0)
    if which == "nacl_small":
        return systems.nacl((6, 6, 6), rcut=6.0, padding=0.3, temperature=1500.0)
    if which == "water":
        return systems.spce_water(4096, rcut=8.0, padding=0.3, temperature=300.0)
    return systems.argon(12, temperature=200.0)


def _sig_err(fg, fo):
    d = np.linalg.norm(fg - fo, axis=1)
    n = np.linalg.norm(fo, axis=1)
    return d, n


def check_rank(t, device, which="nacl", stream_ctx=None, lazy_steps=10, log=None):
    """Runs the whole check for the calling rank; every rank of the transport ``t`` must call it.  Returns a report dict
    (identical global numbers on every rank) or raises AssertionError."""
    import torch
    s = make_system(which)
    rank, world = t.rank, t.world
    ora.build()
    w = ora.World.from_system(s, P=world)
    dom = dd.Domain(s, device=device, transport=t)
    assert tuple(int(v) for v in w.dd(rank)[0][:3]) == tuple(dom.dims)
    w.relocate(); w.set_halo(); assert w.link_cell_pairs() == 0
    oo = w.two_body()
    dom.rebuild()
    out = dom.forces()
    natms, nlast = dom.sr.dev_counts()
    c = w.counts(rank)
    assert (natms, nlast) == (c["natms"], c["nlast"]), ((natms, nlast), (c["natms"], c["nlast"]))
    po, pg = w.parts(rank), dom.sr.dev_get_parts()
    io, ig = w.ints(rank), dom.sr.dev_get_ints()
    for k in ("xxx", "yyy", "zzz", "chge"):
        assert np.array_equal(po[k], pg[k]), k
    for k in ("ltg", "lsite", "ltype", "ixyz"):
        assert np.array_equal(io[k], ig[k]), k
    fo = np.stack([po["fxx"], po["fyy"], po["fzz"]], 1)[:natms]
    fg = np.stack([pg["fxx"], pg["fyy"], pg["fzz"]], 1)[:natms]
    d, n = _sig_err(fg, fo)
    fmax = t.allreduce_max(float(n.max()))
    big = n >= 1.0e-3 * fmax
    per_atom = t.allreduce_max(float((d[big] / n[big]).max()) if big.any() else 0.0)
    max_norm = t.allreduce_max(float(d.max())) / fmax
    assert per_atom <= FORCE_TOL and max_norm <= FORCE_TOL, (per_atom, max_norm)
    tot = dom.gsum(out)
    scale = np.abs(oo[:6]).max()
    erel = max(abs(tot[k] - oo[k]) / max(abs(oo[k]), 1e-6 * scale) for k in range(6))
    assert erel <= ENERGY_TOL, (erel, tot[:6], oo[:6])
    srel = float(np.abs(tot[6:15] - oo[6:15]).max() / np.abs(oo[6:15]).max())
    assert srel <= ENERGY_TOL, srel
    # the one-kernel peer-memory refresh must reproduce the staged exchange bit for bit
    pull_checked = False
    if dom.p2p and not isinstance(t, dd.ThreadTransport):
        with torch.cuda.stream(dom.stream):
            dom.sr.dev_vv(1, 0.001)
        dom.publish()
        t.barrier()
        dom.refresh_halo(staged=True)
        p_staged = dom.sr.dev_get_parts()
        torch.cuda.synchronize(); t.barrier()
        dom.refresh_halo()
        p_pull = dom.sr.dev_get_parts()
        for k in ("xxx", "yyy", "zzz", "chge"):
            assert np.array_equal(p_staged[k], p_pull[k]), ("pull vs staged", k)
        w.vv(1, 0.001, s.weight_by_type); assert w.refresh_halo() == 0
        po2 = w.parts(rank)
        for k in ("xxx", "yyy", "zzz"):
            assert np.abs(po2[k] - p_pull[k]).max() < 1e-9, ("pull vs oracle", k)
        oo = w.two_body(); w.vv(2, 0.001, s.weight_by_type)
        dom.forces()
        with torch.cuda.stream(dom.stream):
            dom.sr.dev_vv(2, 0.001)
        pull_checked = True
    # trajectory, eager driver: decisions and counts per rank; atoms migrate between the domains on rebuilds
    dt, nsteps = (0.0005, 8) if which == "water" else (0.002, 25)
    reb = 0
    gid0 = set(dom.sr.dev_get_ints()["ltg"][:natms].tolist())
    for step in range(nsteps):
        w.vv(1, dt, s.weight_by_type)
        upd, tol = w.vnl_check()
        if upd:
            w.relocate(); w.set_halo(); assert w.link_cell_pairs() == 0
        else:
            assert w.refresh_halo() == 0
        oo = w.two_body()
        w.vv(2, dt, s.weight_by_type)
        r0 = dom.rebuilds
        out = dom.step(dt)
        assert (dom.rebuilds != r0) == upd, (step, upd)
        reb += int(upd)
        tot = dom.gsum(out)
        assert abs(tot[0] + tot[2] - oo[0] - oo[2]) <= 1e-8 * abs(oo[0] + oo[2]), (step, tot[:4], oo[:4])
        n2, l2 = dom.sr.dev_counts()
        c = w.counts(rank)
        assert (n2, l2) == (c["natms"], c["nlast"]), (step, (n2, l2), (c["natms"], c["nlast"]))
    assert reb >= 1 or which == "water"
    n2, _ = dom.sr.dev_counts()
    gid1 = set(dom.sr.dev_get_ints()["ltg"][:n2].tolist())
    migrated = int(t.allreduce_sum([len(gid1 - gid0)])[0])
    # after the trajectory the resident atoms still are the oracle's (same set per rank; order follows the same rules)
    assert gid1 == set(w.ints(rank)["ltg"][:w.counts(rank)["natms"]].tolist())
    # ... and in the oracle's ORDER, halo included: deport_atomic_data's restack (the r-th hole takes the r-th staying atom counted
    # from the end, deport_data.F90:822-925) and the append order of the received atoms are reproduced, not just the sets
    io2, ig2 = w.ints(rank), dom.sr.dev_get_ints()
    c2 = w.counts(rank)
    assert np.array_equal(io2["ltg"][:c2["nlast"]], ig2["ltg"][:c2["nlast"]]), "resident order after the trajectory"
    # library-enqueued steps: the sums of step n come back with step n+1, already summed over the ranks (mailbox gsum)
    lazy_checked = 0
    if dom.xchg and dom.p2p and which != "water":
        ref_tot = []
        prevs = []
        for step in range(lazy_steps):
            w.vv(1, dt, s.weight_by_type)
            upd, tol = w.vnl_check()
            if upd:
                w.relocate(); w.set_halo(); assert w.link_cell_pairs() == 0
            else:
                assert w.refresh_halo() == 0
            ref_tot.append(w.two_body())
            w.vv(2, dt, s.weight_by_type)
            r0 = dom.rebuilds
            prev = dom.step(dt, lazy=True)
            assert (dom.rebuilds != r0) == upd, ("lazy", step, upd)
            if step > 0:
                prevs.append(prev)
        last = dom.gsum(dom.collect())            # the last step's sums are this rank's partials: reduce them the classic way
        prevs.append(last)
        for k, (got, want) in enumerate(zip(prevs, ref_tot)):
            # trajectories of the two engines drift apart at the 1e-13 level per step (summation order), energies follow
            assert abs(got[0] + got[2] - want[0] - want[2]) <= 1e-8 * abs(want[0] + want[2]), ("lazy gsum", k, got[:4], want[:4])
            assert np.abs(got[6:15] - want[6:15]).max() <= 1e-7 * np.abs(want[6:15]).max(), ("lazy gsum stress", k)
        # every rank holds the SAME reduced bits (rank-ordered sum): compare through a max/min reduction
        chk = np.array(prevs[0][:6])
        hi = np.array([t.allreduce_max(v) for v in chk]); lo = -np.array([t.allreduce_max(-v) for v in chk])
        assert np.array_equal(hi, lo), "mailbox gsum differs between ranks"
        lazy_checked = len(prevs)
    t.barrier()
    rep = {"ok": True, "ranks": world, "domains": list(dom.dims), "system": s.name, "atoms": int(s.megatm),
           "per_atom_force_rel": per_atom, "max_normalised_force": max_norm, "energy_rel": float(erel), "stress_rel": srel,
           "rebuilds": reb, "migrated_atoms": migrated, "pull_refresh_checked": pull_checked, "mailbox_gsum_steps": lazy_checked,
           "exchange": "peer-memory (fused)" if dom.xchg else ("peer-memory refresh + staged" if dom.p2p else "staged messages")}
    if log is not None and rank == 0:
        log(rep)
    dom.close()
    return rep


def spme_rank(t, device, which):
    """One rank of the several-domain SPME: its own atoms, the replicated grid summed over the ranks (dd.Domain.spme_forces)."""
    from oracle import spme_oracle as so
    s = make_system(which)
    dom = dd.Domain(s, device=device, transport=t)
    _, kdim = so.spme_grid(1.0e-6, s.rcut, s.cell)
    dom.set_spme(kdim, 8)
    out = dom.spme_forces()
    tot = dom.gsum(out)
    natms, _ = dom.sr.dev_counts()
    pg = dom.sr.dev_get_parts()
    f = np.stack([pg["fxx"], pg["fyy"], pg["fzz"]], 1)[:natms]
    ltg = dom.sr.dev_get_ints(natms)["ltg"][:natms]
    # a second call ADDS the same forces again
    dom.spme_forces()
    pg2 = dom.sr.dev_get_parts()
    f2 = np.stack([pg2["fxx"], pg2["fyy"], pg2["fzz"]], 1)[:natms]
    assert np.abs(f2 - 2.0 * f).max() <= 1e-12 * max(np.abs(f).max(), 1.0)
    t.barrier()
    dom.close()
    return {"tot": tot, "ltg": ltg, "f": f, "kdim": kdim}


def spme_compare(reps, which):
    """Per-atom forces by global id and the gsum-med sums of the ranks' reports against the one-domain numpy oracle."""
    from oracle import spme_oracle as so
    from util import per_atom_force_error
    s = make_system(which)
    xyz = dd.read_config_fold(s.xyz, s.cell)[0]
    q = s.charge_site[s.lsite - 1]
    ref = so.ewald_spme_forces_coul(s.cell, xyz, q, s.ff.alpha, reps[0]["kdim"], 8, s.ff.scaling)
    f = np.zeros((s.megatm, 3))
    seen = np.zeros(s.megatm, dtype=int)
    for r in reps:
        f[r["ltg"] - 1] = r["f"]
        seen[r["ltg"] - 1] += 1
    assert (seen == 1).all()
    rep = per_atom_force_error(f, ref["forces"])
    assert rep["max_normalised"] <= FORCE_TOL and rep["per_atom_significant"] <= FORCE_TOL, rep
    tot = reps[0]["tot"]
    for r in reps[1:]:
        assert np.array_equal(r["tot"], tot)
    assert abs(tot[11] - ref["eng_recip"]) <= ENERGY_TOL * abs(ref["eng_recip"])
    assert abs(tot[0] - ref["engcpe_rc"]) <= ENERGY_TOL * abs(ref["engcpe_rc"])
    assert abs(tot[1] - ref["vircpe_rc"]) <= ENERGY_TOL * abs(ref["vircpe_rc"])
    assert np.abs(tot[2:11] - ref["stress"]).max() <= ENERGY_TOL * np.abs(ref["stress"]).max()
    return rep
