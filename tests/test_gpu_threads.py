"""Multi-rank parity on ONE GPU: the ranks of the domain decomposition run as threads of this process, one CUDA stream and
one library context each, and exchange migrating atoms, halo atoms, refreshed positions, the gmax and the gsum through the
library's peer-memory kernels exactly as they do between GPUs (the peers' buffers are plain device pointers instead of
CUDA-IPC mappings).  Checked against the oracle's P-domain world: see tests/dd_common.py.  This is what puts the N > 1 path
on the record of a 1-GPU box; tests/test_gpu_multi.py runs the same check with one process per GPU when there are several."""
import gc
import threading

import numpy as np
import pytest

from dl_poly_b200 import dd
import dd_common

pytestmark = pytest.mark.gpu


def run_threads(world, which, fn=None, need_ok=True):
    grp = dd.ThreadGroup(world)
    reps, errs = [None] * world, [None] * world

    def body(rank):
        import torch
        try:
            torch.cuda.set_device(0)
            reps[rank] = (fn or dd_common.check_rank)(grp.transport(rank), 0, which)
        except BaseException as e:          # a failing rank must not leave the others waiting at a barrier for ever
            errs[rank] = e
            grp.bar.abort()

    # dlpgpu_destroy frees pinned host memory and peer regions, which waits for the whole device; a context of an EARLIER test
    # collected by the cyclic GC from inside a rank thread would wait for a peer's spinning receive kernel, which waits for
    # that thread's next message.  Collect before the ranks start and keep the collector quiet while they run.  (The hang this
    # suite did hit was a different host-side wait: see the page-locked counts of dlpgpu_dev_xchg_rebuild.)
    gc.collect()
    gc.disable()
    try:
        th = [threading.Thread(target=body, args=(r,)) for r in range(world)]
        for x in th:
            x.start()
        for x in th:
            x.join(timeout=900)
    finally:
        gc.enable()
    real = [e for e in errs if e is not None and not isinstance(e, threading.BrokenBarrierError)]
    if real:
        raise AssertionError("rank errors: " + " || ".join("rank %d: %r" % (r, e) for r, e in enumerate(errs) if e is not None)) from real[0]
    assert all(e is None for e in errs), errs
    if not need_ok:
        return reps
    assert all(r is not None and r["ok"] for r in reps)
    return reps[0]


@pytest.mark.parametrize("world,which", [(2, "nacl"), (4, "nacl"), (8, "nacl"), (2, "water"), (8, "argon"), (8, "nacl_hot"), (2, "nacl_hot")])
def test_domains_as_threads_against_oracle(world, which):
    rep = run_threads(world, which)
    print("THREAD-RANKS %s" % rep)
    assert rep["ranks"] == world and rep["rebuilds"] >= (0 if which == "water" else 1)
    if which != "water":
        assert rep["mailbox_gsum_steps"] > 0
    if which in ("nacl", "nacl_hot"):          # the hot melt moves atoms across domain faces within the checked steps; cold argon need not
        assert rep["migrated_atoms"] > 0


def test_scanning_migration_stages_against_oracle(monkeypatch):
    """The first implementation of the migration stages (every stage scans all atoms) is kept as a diagnostic
    (dlpgpu_dev_xchg_set_migration); the default compacts the movers once and runs one single-block kernel per stage.  Both are
    held to the oracle's resident atoms, bit for bit and in order, by the same check."""
    monkeypatch.setenv("DLP_DD_SCAN_MIGRATION", "1")
    rep = run_threads(4, "nacl")
    assert rep["ranks"] == 4 and rep["migrated_atoms"] > 0
    rep = run_threads(8, "nacl_hot")
    assert rep["migrated_atoms"] > 50


@pytest.mark.parametrize("world", [2, 8])
def test_spme_over_domains_against_oracle(world):
    """f4 over several domains (replicated grid: spread per rank, sum of the grids over the ranks, whole-grid transform and gather
    per rank, net force of all ranks removed) against the one-domain numpy oracle of the whole system: per-atom forces by global
    id, energy / virial / stress after gsum."""
    reps = run_threads(world, "nacl", fn=dd_common.spme_rank, need_ok=False)
    rep = dd_common.spme_compare(reps, "nacl")
    print("SPME over %d domains: %s" % (world, rep))
