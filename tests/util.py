"""Shared helpers of the parity tests (tests only)."""
import numpy as np

from oracle import oracle as ora


def world_for(sysm, P=1, with_halo=True, with_list=True):
    w = ora.World.from_system(sysm, P=P)
    if with_halo:
        w.set_halo()
    if with_list:
        rc = w.link_cell_pairs()
        assert rc == 0, "oracle link_cell_pairs rc=%d" % rc
    return w


def domain_inputs(w, rank):
    """The arrays a DL_POLY rank would pass through the C ABI for its domain."""
    c = w.counts(rank)
    ints = w.ints(rank)
    parts = w.parts(rank)
    d = dict(natms=c["natms"], nlast=c["nlast"], parts=parts, ltype=ints["ltype"], ltg=ints["ltg"], lfrzn=ints["lfrzn"],
             max_list=c["max_list"], dd=w.dd(rank)[0])
    d["list_excl"] = w.list_excl(rank) if c["max_exclude"] > 0 else None
    return d


def rel_err(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def force_errors(f, fref):
    """(max_i |dF_i| / max_j |F_j| , max_i |dF_i| / |F_i|)"""
    d = np.linalg.norm(f - fref, axis=1)
    n = np.linalg.norm(fref, axis=1)
    return float(d.max() / n.max()), float((d / np.maximum(n, 1e-300)).max())


SIGNIFICANT = 1.0e-3      # atoms whose net force is at least this fraction of the largest force carry the per-atom bar


def per_atom_force_error(f, fref):
    """North-star bar "per-atom forces within 1e-9 relative": the worst |dF_i| / |F_i| over the atoms with
    |F_i| >= SIGNIFICANT * max|F| (an atom whose pair terms cancel to a net force a thousand times smaller than the typical
    one loses those digits to the summation order in ANY implementation, the reference's own MPI decomposition included),
    next to the worst ratio over all atoms and the max-normalised error.  Returns a dict of the three numbers."""
    d = np.linalg.norm(f - fref, axis=1)
    n = np.linalg.norm(fref, axis=1)
    big = n >= SIGNIFICANT * n.max()
    return {"per_atom_significant": float((d[big] / n[big]).max()) if big.any() else 0.0,
            "per_atom_all": float((d / np.maximum(n, 1e-300)).max()),
            "max_normalised": float(d.max() / max(n.max(), 1e-300)),
            "significant_atoms": int(big.sum()), "atoms": int(len(n))}


def parts_forces(parts, n):
    return np.stack([parts["fxx"][:n], parts["fyy"][:n], parts["fzz"][:n]], axis=1)
