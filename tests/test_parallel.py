"""Partitioning + halo exchange (host logic), on the CPU: the CPU oracle stands in for the GPU kernels so the
N-rank plumbing (partition sets, local meshes, send/recv lists, exchange schedule) is checked without a GPU.
N-rank results on owned entities must be BIT-IDENTICAL to the single-partition run."""
import os

import numpy as np
import pytest

from mpas_regent_b200 import _abi, dynamics, init_jw, parallel, partition
from mpas_regent_b200 import mesh as M
from mpas_regent_b200._abi import CELL, EDGE, FIELD_ENTITY, VERTEX, VERTICAL

L = 6
DT = 600.0


def test_partition_regions_reproduces_survey_numbers(grid2562):
    """SURVEY.md 8c cross-checks on the bundled mesh + 16-way METIS colouring, colours 0/1/2/15 and totals."""
    want = {
        M.LITERAL: dict(ghost_1=[267, 282, 266, 334], ghost_2=[475, 488, 438, 559], shared_1=[102, 88, 114, 108],
                        private_1=[59, 67, 47, 56], shared_2=[143, 119, 155, 143], private_2=[18, 36, 6, 21],
                        totals=dict(ghost_1=4405, ghost_2=7374, shared_1=1586)),
        M.CORRECTED: dict(ghost_1=[54, 46, 49, 53], ghost_2=[114, 97, 103, 111], shared_1=[48, 41, 44, 48],
                          private_1=[113, 114, 117, 116], shared_2=[88, 77, 83, 89], private_2=[73, 78, 78, 75],
                          totals=dict(ghost_1=790, ghost_2=1662, shared_1=706)),
    }
    for pol, w in want.items():
        part = partition.partition_regions(grid2562, grid2562.partition, 16, pol)
        for k, vals in w.items():
            if k == "totals":
                for kk, tot in vals.items():
                    assert sum(len(a) for a in getattr(part, kk)) == tot, (pol, kk)
            else:
                assert [len(getattr(part, k)[c]) for c in (0, 1, 2, 15)] == vals, (pol, k)
        for c in range(16):     # the algebra of mesh_loading.rg:448-471
            p = set(part.p[c])
            assert not (set(part.ghost_1[c]) & p) and not (set(part.ghost_2[c]) & p)
            assert set(part.shared_1[c]) | set(part.private_1[c]) == p
            assert set(part.private_2[c]) <= set(part.private_1[c])
            if pol == M.CORRECTED:
                assert set(part.ghost_1[c]) <= set(part.ghost_2[c])


def test_sfc_colouring_is_balanced_and_compact(grid2562):
    col = partition.sfc_colouring(grid2562, 8)
    sizes = np.bincount(col, minlength=8)
    assert sizes.max() - sizes.min() <= 1
    part = partition.partition_regions(grid2562, col, 8, M.CORRECTED)
    assert max(len(g) for g in part.ghost_1) < 120        # compact chunks: ring ~ O(sqrt(n))


def _shards(mesh, world):
    st = init_jw.make_state(mesh, L, _abi.INDEX_CORRECTED)
    return st, parallel.make_shards(st, world)


def test_halo_lists_are_symmetric(grid642):
    st, shards = _shards(grid642, 4)
    locs = [s["lm"] for s in shards]
    for ent, attr in (("cell", "cells"), ("edge", "edges"), ("vertex", "vertices")):
        owned_total = 0
        for h, lm in enumerate(locs):
            n_own = lm.n_owned[("cell", "edge", "vertex").index(ent)]
            owned_total += n_own
            assert np.all(lm.owner[ent][:n_own] == h) and np.all(lm.owner[ent][n_own:] != h)
            for o, ridx in lm.recv[ent].items():
                sidx = locs[o].send[ent][h]
                assert np.array_equal(getattr(lm, attr)[ridx], getattr(locs[o], attr)[sidx])     # same global entities, same order
                assert np.all(ridx >= n_own) and np.all(sidx < locs[o].n_owned[("cell", "edge", "vertex").index(ent)])
            covered = np.concatenate([v for v in lm.recv[ent].values()]) if lm.recv[ent] else np.zeros(0, int)
            assert len(covered) == len(getattr(lm, attr)) - n_own                               # every ghost has exactly one source
        assert owned_total == {"cell": grid642.nCells, "edge": grid642.nEdges, "vertex": grid642.nVertices}[ent]


def _single(mesh, st, steps, physics=_abi.PHYSICS_LITERAL):
    from oracle.oracle import Oracle
    o = Oracle(dynamics.dims_of(mesh, L), _abi.default_config(rkarg_policy=_abi.RKARG_STAGE_INDEX, physics_mode=physics))
    o.upload_mesh(st.static); o.upload_state(st.f, st.vert)
    o.atm_compute_solve_diagnostics(False, -1)
    for _ in range(steps):
        o.atm_srk3(DT)
    return o


def _rank_backend(sh, physics=_abi.PHYSICS_LITERAL):
    from oracle.oracle import Oracle
    lm = sh["lm"]
    o = Oracle(_abi.make_dims(len(lm.cells), len(lm.edges), len(lm.vertices), L),
               _abi.default_config(rkarg_policy=_abi.RKARG_STAGE_INDEX, physics_mode=physics))
    o.upload_mesh(sh["static"]); o.upload_state(sh["f"], sh["vert"])
    return o


def _assert_owned_equal(single, backend, lm, names=None):
    rows = {CELL: lm.cells, EDGE: lm.edges, VERTEX: lm.vertices}
    nown = {CELL: lm.n_owned[0], EDGE: lm.n_owned[1], VERTEX: lm.n_owned[2]}
    for name, ent, _ in _abi.FIELDS:
        if names and name not in names:
            continue
        a = backend.download_field(name)
        ref = single.download_field(name)
        if ent == VERTICAL:
            assert np.array_equal(a, ref, equal_nan=True), name
        else:
            n = nown[ent]
            assert np.array_equal(a[:n], ref[rows[ent][:n]], equal_nan=True), (name, lm.rank)


def _task_schedule(cfg):
    """the (task, args) list of one atm_srk3, recorded from the shared host replay (rk_timestep.rg:378-481)."""
    seq = []

    class Rec(dynamics.TaskAPI):
        def __init__(self):
            self.cfg, self.dims = cfg, None

        def _call(self, name, *args):
            seq.append((name, args))

    Rec().atm_srk3_by_tasks(DT)
    return seq


@pytest.mark.parametrize("physics", [_abi.PHYSICS_LITERAL, _abi.PHYSICS_CORRECTED], ids=["literal", "corrected_physics"])
@pytest.mark.parametrize("world", [2, 4])
def test_n_ranks_equal_single_partition_bitwise(grid642, world, physics):
    st, shards = _shards(grid642, world)
    single = _single(grid642, st, 2, physics)
    backs = [_rank_backend(s, physics) for s in shards]
    ex = parallel.InProcessExchanger(backs, [s["lm"] for s in shards])
    exchanges = parallel.exchanges_for(backs[0].cfg)
    # lock-step: every rank runs the same task, then ONE exchange serves all of them
    for b in backs:
        b.atm_compute_solve_diagnostics(False, -1)
    ex.exchange(exchanges["compute_solve_diagnostics"])
    seq = _task_schedule(backs[0].cfg)
    assert ("recover_large_step_variables" in [n for n, _ in seq]) == (physics == _abi.PHYSICS_CORRECTED)
    for _ in range(2):
        for name, args in seq:
            for b in backs:
                b._call(name, *args)
            spec = exchanges.get(parallel.exchange_key(name, args))
            if spec:
                ex.exchange(spec)
    for b, s in zip(backs, shards):
        _assert_owned_equal(single, b, s["lm"])
    # the step's sanity scan (summarize_timestep, rk_timestep.rg:29-359; bench.py `check`): per-rank summaries over the
    # OWNED entities merge to exactly the single-partition summary -- min / max with place, counts, bit checksum
    import bench
    want = bench.run_check(single)
    parts = [bench.run_check(b, s["lm"]) for b, s in zip(backs, shards)]
    got = bench.merge_checks([p["fields"] for p in parts])
    assert got["fields"] == want["fields"] and got["combined_checksum"] == want["combined_checksum"]


def test_summarize_np_definition():
    """the numpy restatement of MpasFieldSummary (include/mpas_b200.h) on a hand-checkable case"""
    from oracle.oracle import summarize_np
    a = np.array([[1.0, -0.0, 0.0], [np.nan, -3.5, np.inf]])
    s = summarize_np(a)
    assert (s["min"], s["min_at"], s["max"], s["max_at"], s["n_nan"], s["n_inf"], s["count"]) == (-3.5, [1, 1], np.inf, [1, 2], 1, 1, 6)
    # checksum: order-independent (permuting rows together with their ids changes nothing), sensitive to a swap of values
    g = np.array([7, 2])
    assert summarize_np(a, g)["checksum"] == summarize_np(a[::-1], g[::-1])["checksum"]
    assert summarize_np(a, g)["checksum"] != summarize_np(a[::-1], g)["checksum"]
    assert summarize_np(a, g)["min_at"] == [2, 1]
    # first place in (id, level) order on ties
    b = np.array([[2.0, 2.0], [2.0, 1.0], [1.0, 5.0]])
    assert summarize_np(b)["min_at"] == [1, 1] and summarize_np(b)["max_at"] == [2, 1]
    assert summarize_np(np.full((2, 2), np.nan))["min_at"] == [-1, -1]


def test_overlapped_schedule_in_process(grid642):
    """the interior / boundary split of DistributedDynamics._acoustic_pair (sent cells, exchange, interior cells, interior
    edges, edges next to ghosts; ghost cells never advanced) leaves owned entities bit-identical to the single partition."""
    world = 4
    st, shards = _shards(grid642, world)
    single = _single(grid642, st, 2)
    backs = [_rank_backend(s) for s in shards]
    ex = parallel.InProcessExchanger(backs, [s["lm"] for s in shards])
    exchanges = parallel.exchanges_for(backs[0].cfg)
    for b in backs:
        b.atm_compute_solve_diagnostics(False, -1)
    ex.exchange(exchanges["compute_solve_diagnostics"])
    seq = _task_schedule(backs[0].cfg)
    for _ in range(2):
        i = 0
        while i < len(seq):
            name, args = seq[i]
            if name == "advance_acoustic_step":
                assert seq[i + 1][0] == "divergence_damping_3d"
                for b in backs:
                    b.restrict(CELL, 1); b._call(name, *args)
                ex.exchange(exchanges[parallel.exchange_key(name, args)])
                for b in backs:
                    b.restrict(CELL, 0); b._call(name, *args)
                    b.restrict(EDGE, 0); b._call(*seq[i + 1][:1], *seq[i + 1][1])
                    b.restrict(EDGE, 1); b._call(*seq[i + 1][:1], *seq[i + 1][1])
                    b.restrict(CELL); b.restrict(EDGE)
                i += 2
                continue
            for b in backs:
                b._call(name, *args)
            spec = exchanges.get(parallel.exchange_key(name, args))
            if spec:
                ex.exchange(spec)
            i += 1
    for b, s in zip(backs, shards):
        _assert_owned_equal(single, b, s["lm"])
    # ghost cells were never advanced: their rw_p is still what the upload left (zero), while owned cells moved
    lm = shards[0]["lm"]
    rw = backs[0].download_field("rw_p")
    assert not rw[lm.n_owned[0]:].any() and rw[:lm.n_owned[0]].any()


def _gloo_worker(rank, world, port, tmp, overlap=False):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mpas_regent_b200 import icosa
        mesh = icosa.make_icosahedral_mesh(642)
        st, shards = _shards(mesh, world)
        sh = shards[rank]
        b = _rank_backend(sh)
        run = parallel.DistributedDynamics(b, parallel.HostDistExchanger(b, sh["lm"]))
        if overlap:
            assert run.enable_overlap()
        run.init_diagnostics()
        for _ in range(2):
            run.step(DT)
        run.flush()
        single = _single(mesh, st, 2)
        _assert_owned_equal(single, b, sh["lm"])
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [False, True], ids=["plain", "overlapped_schedule"])
def test_gloo_world_size_2(tmp_path, overlap):
    """the torch.distributed path (batch_isend_irecv), world_size 2, gloo on 127.0.0.1; with ``overlap`` the schedule that
    bench.py --gpus N runs (interior / boundary split, deferred diagnostics exchange), the exchange blocking in start()."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path), overlap), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


def test_shard_roundtrip(grid642, tmp_path):
    st, shards = _shards(grid642, 2)
    parallel.save_shard(str(tmp_path / "r0"), shards[0])
    sh = parallel.load_shard(str(tmp_path / "r0"))
    lm0, lm1 = shards[0]["lm"], sh["lm"]
    assert np.array_equal(lm0.cells, lm1.cells) and lm0.n_owned == lm1.n_owned
    for ent in ("cell", "edge", "vertex"):
        assert lm0.send[ent].keys() == lm1.send[ent].keys()
        for p in lm0.recv[ent]:
            assert np.array_equal(lm0.recv[ent][p], lm1.recv[ent][p])
    for k in shards[0]["f"]:
        assert np.array_equal(shards[0]["f"][k], sh["f"][k])


def test_launch_classes_for_overlap(grid642):
    """cells: 0 = owned & never sent, 1 = owned & sent, 2 = ghost; edges: 0 = both cells owned.  A class-0 cell has only
    owned neighbours, so advancing it needs nothing that an exchange in flight delivers."""
    st, shards = _shards(grid642, 4)
    for sh in shards:
        lm, loc = sh["lm"], sh["static"]
        cc, ec = loc["cellClass"], loc["edgeClass"]
        n_own = lm.n_owned[0]
        assert cc.dtype == np.uint8 and np.all(cc[n_own:] == 2) and np.all(cc[:n_own] < 2)
        sent = np.unique(np.concatenate(list(lm.send["cell"].values())))
        assert np.array_equal(np.nonzero(cc == 1)[0], sent)
        coe = loc["cellsOnEdge"].reshape(-1, 2).astype(np.int64) - 1
        both_owned = ((coe >= 0) & (coe < n_own)).all(axis=1)
        assert np.array_equal(ec == 0, both_owned)
        # every edge of a class-0 cell is a class-0 edge (its other cell is owned too)
        eoc = loc["edgesOnCell"].astype(np.int64) - 1
        for c in np.nonzero(cc == 0)[0][:200]:
            es = eoc[c, :loc["nEdgesOnCell"][c]]
            assert np.all(ec[es] == 0)
        assert (cc == 0).sum() > 0 and (cc == 1).sum() > 0


@pytest.mark.parametrize("kind", ["random", "stripes", "one_rank_tiny"])
def test_irregular_colourings_still_match_single_partition(grid642, kind):
    """colourings a METIS file could contain: scattered cells (every cell next to another rank), latitude stripes, a rank
    that owns three cells.  Owned entities stay bit-identical to the single partition, with and without the split schedule."""
    rng = np.random.default_rng(1)
    n = grid642.nCells
    if kind == "random":
        col = rng.integers(0, 3, n).astype(np.int32)
    elif kind == "stripes":
        col = np.minimum((np.argsort(np.argsort(grid642.v["latCell"])) * 4) // n, 3).astype(np.int32)
    else:
        col = np.zeros(n, np.int32); col[n // 2:] = 1; col[:3] = 2
    world = int(col.max()) + 1
    st = init_jw.make_state(grid642, L, _abi.INDEX_CORRECTED)
    shards = parallel.make_shards(st, world, colours=col)
    single = _single(grid642, st, 1)
    backs = [_rank_backend(s) for s in shards]
    ex = parallel.InProcessExchanger(backs, [s["lm"] for s in shards])
    exchanges = parallel.exchanges_for(backs[0].cfg)
    for b in backs:
        b.atm_compute_solve_diagnostics(False, -1)
    ex.exchange(exchanges["compute_solve_diagnostics"])
    for name, args in _task_schedule(backs[0].cfg):
        for b in backs:
            b._call(name, *args)
        spec = exchanges.get(parallel.exchange_key(name, args))
        if spec:
            ex.exchange(spec)
    for b, s in zip(backs, shards):
        _assert_owned_equal(single, b, s["lm"])
        cc = s["static"]["cellClass"]
        assert (cc == 2).sum() == len(s["lm"].cells) - s["lm"].n_owned[0]


def test_n_ranks_equal_single_partition_with_scalar_transport(grid642):
    """config_scalar_advection: atm_advance_scalars reads `scalars` two rings out, so it is exchanged after every stage
    (parallel.EXCHANGES["advance_scalars"], rk_timestep.rg:469); the array-typed field travels slot by slot."""
    from oracle.oracle import Oracle
    world = 3
    st = init_jw.make_state(grid642, L, _abi.INDEX_CORRECTED)
    st.f["scalars"] = 1e-3 * (1.0 + np.random.default_rng(2).random((grid642.nCells, L + 1, 8)))
    shards = parallel.make_shards(st, world)
    cfg = dict(rkarg_policy=_abi.RKARG_STAGE_INDEX, config_scalar_advection=1, physics_mode=_abi.PHYSICS_CORRECTED)
    single = Oracle(dynamics.dims_of(grid642, L), _abi.default_config(**cfg))
    single.upload_mesh(st.static); single.upload_state(st.f, st.vert)
    single.atm_compute_solve_diagnostics(False, -1)
    single.atm_srk3(DT)
    backs = []
    for sh in shards:
        lm = sh["lm"]
        o = Oracle(_abi.make_dims(len(lm.cells), len(lm.edges), len(lm.vertices), L), _abi.default_config(**cfg))
        o.upload_mesh(sh["static"]); o.upload_state(sh["f"], sh["vert"])
        backs.append(o)
    ex = parallel.InProcessExchanger(backs, [s["lm"] for s in shards])
    exchanges = parallel.exchanges_for(backs[0].cfg)
    for b in backs:
        b.atm_compute_solve_diagnostics(False, -1)
    ex.exchange(exchanges["compute_solve_diagnostics"])
    seq = _task_schedule(backs[0].cfg)
    assert [n for n, _ in seq].count("advance_scalars") == 3
    for name, args in seq:
        for b in backs:
            b._call(name, *args)
        spec = exchanges.get(parallel.exchange_key(name, args))
        if spec:
            ex.exchange(spec)
    assert not np.array_equal(single.download_field("scalars"), st.f["scalars"])
    for b, s in zip(backs, shards):
        _assert_owned_equal(single, b, s["lm"], names=("scalars", "scalars_old", "w", "theta_m", "u"))
