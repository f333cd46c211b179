"""ScanContext search over a database sharded across ranks and exchanged through peer-memory windows (csrc/sc_shard.cuh,
liorf_sc_shard_*; SURVEY §8e, BASELINE config 5).  Here the ranks are several contexts on ONE GPU that map each other's
windows by pointer (across processes the same windows are mapped through cudaIpc handles — bench.py --gpus N checks that path against
the unsharded search on every rank); the kernels, the flags and the protocol are the ones the multi-GPU run uses.  Loop ids, shifts,
fp64 distances and the candidate triples must equal the ORACLE (include/Scancontext.cpp:253-344 semantics) and the unsharded GPU
search bit for bit, on every rank, batch after batch."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _keys(o, descs):
    return np.stack([o.sc_keys_from_desc(d)[0] for d in descs])


def _eq_dist(a, b):
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])


def _run(synth, K, Q_list, world, path, bounds=None, oracle=None):
    import torch
    import liorf_b200
    from liorf_b200.sc_sharded import PeerShardedSearch
    db = synth.sc_descriptors(K, seed=71 + K)
    if K > 300:
        db[211] = db[5]; db[K - 3] = db[5]                             # duplicate keys on different shards: ties resolved by global index
    ref = liorf_b200.Context(); ref.scAddDescriptors(db); ref.scSetSearchPath(path)
    bounds = bounds or [K * g // world for g in range(world + 1)]
    ctxs = [liorf_b200.Context() for _ in range(world)]
    for g, c in enumerate(ctxs):
        if bounds[g + 1] > bounds[g]:
            c.scAddDescriptors(db[bounds[g]:bounds[g + 1]])
        c.scSetSearchPath(path)
    S = [PeerShardedSearch(c, g, world, bounds, max(Q_list), torch) for g, c in enumerate(ctxs)]
    for s in S:
        s.connect_local(S)
    PeerShardedSearch.sync_keys_local(S)                               # the replicated index: every rank's ring keys in every rank's key array
    dev = S[0].dev
    okeys = _keys(oracle, db) if oracle is not None else None
    for b, Q in enumerate(Q_list):
        qd, src, shift = synth.sc_queries(db, Q, seed=72 + b)
        if K > 300:
            qd[0] = db[5]
        r_loop, r_sh, r_dist, r_cand = ref.scQueryBatch(qd)
        if oracle is not None:                                         # the unsharded GPU search is itself checked against the oracle here
            o_loop, o_sh, o_dist, o_cand = oracle.sc_query_batch(okeys, db, _keys(oracle, qd), qd)
            assert np.array_equal(r_cand, o_cand) and np.array_equal(r_loop, o_loop) and np.array_equal(r_sh, o_sh) and _eq_dist(r_dist, o_dist), b
        order = list(range(world)) if b % 2 == 0 else list(range(world - 1, -1, -1))      # enqueue order must not matter
        res, dq = {}, {}
        for g in order:
            with torch.cuda.stream(S[g].stream):
                dq[g] = torch.from_numpy(qd).to(dev)
        # the ranks share ONE device here: enqueue step by step over the ranks, so that a kernel that waits for a peer's push is
        # never queued in front of the kernel that pushes (separate GPUs run the three steps as one call)
        for step in ((1, 2, 4) if world > 1 else (7,)):
            for g in order:
                res[g] = S[g].query(dq[g], phases=step)
        for c in ctxs:
            c.sync()
        torch.cuda.synchronize()
        for g in range(world):
            loop, sh, dd, cand = [t.cpu().numpy() for t in res[g]]
            assert np.array_equal(cand, r_cand), (b, g)
            assert np.array_equal(loop, r_loop) and np.array_equal(sh, r_sh), (b, g)
            assert _eq_dist(dd, r_dist), (b, g)
        planted = src >= 0
        planted[0] = False                                             # query 0 was replaced above
        assert (planted.sum() > 0 or Q < 16) and np.array_equal(r_loop[planted], src[planted])
    stats = [c.scTensorStats() for c in ctxs]
    for c in ctxs:
        c.close()
    ref.close()
    return stats


def test_three_shards_tensor_path(synth, oracle):
    """tcgen05 filter + exact re-rank of each rank's query slice against the replicated index; four batches of different sizes through
    the same windows (flags carry the batch number, the arrays are reused); results == oracle == unsharded GPU search"""
    stats = _run(synth, 13000, [333, 1000, 64, 777], 3, 2, oracle=oracle)
    print("candidate chunks per rank of the last batch:", [s["candidates"] for s in stats])


def test_two_uneven_shards_brute_path(synth, oracle):
    """CUDA-core exact search of the slices (phase C pushed by k_scsh_push_c), shards of very different size (one holds 5 rows)"""
    _run(synth, 2005, [50, 129], 2, 1, bounds=[0, 2000, 2005], oracle=oracle)


def test_uneven_shards_straddling_the_auto_threshold(synth):
    """K = 32767 over 8 ranks gives shards of 4095 and 4096 rows — with the path chosen from the LOCAL extent the ranks
    disagreed and stalled.  The path now depends on (slice size, TOTAL rows) and every path ends in the same push, so any mix works."""
    _run(synth, 32767, [1024, 70], 8, 0)


def test_empty_shard_and_tiny_batches(synth):
    """a rank without rows still searches its slice and raises its flags; Q < world leaves some slices empty"""
    _run(synth, 900, [40, 2, 1], 3, 0, bounds=[0, 450, 450, 900])


def test_single_rank_degenerate(synth, oracle):
    _run(synth, 5000, [200], 1, 2, oracle=oracle)


def test_stage1_work_shards_with_the_rank_count(synth):
    """stage 1 is split by query: the candidate chunks the ranks re-rank add up to what one unsharded search re-ranks (each query is
    filtered exactly once, against the same keys)"""
    import liorf_b200
    K, Q = 40000, 1000
    db = synth.sc_descriptors(K, seed=71 + K)
    db[211] = db[5]; db[K - 3] = db[5]
    qd, _, _ = synth.sc_queries(db, Q, seed=72)
    qd[0] = db[5]
    one = liorf_b200.Context(); one.scAddDescriptors(db); one.scSetSearchPath(2); one.scQueryBatch(qd)
    base = one.scTensorStats()["candidates"]; one.close()
    stats = _run(synth, K, [Q], 4, 2)
    total = sum(s["candidates"] for s in stats)
    print(f"candidate chunks: unsharded {base}, 4 shards together {total}")
    assert total == base


def test_wrong_ownership_is_refused(synth):
    """a context whose rows are not the ones announced at connect time must be refused, not silently read stale slots"""
    import torch
    import liorf_b200
    from liorf_b200.sc_sharded import PeerShardedSearch
    db = synth.sc_descriptors(400, seed=5)
    c = liorf_b200.Context(); c.scAddDescriptors(db[:100])
    s = PeerShardedSearch(c, 0, 2, [0, 200, 400], 16, torch)
    with pytest.raises(RuntimeError):                               # holds 100 rows, announced 200
        s.connect_local([s, s])
    c.close()
    c0, c1 = liorf_b200.Context(), liorf_b200.Context()
    c0.scAddDescriptors(db[:200]); c1.scAddDescriptors(db[200:])
    S = [PeerShardedSearch(c0, 0, 2, [0, 200, 400], 16, torch), PeerShardedSearch(c1, 1, 2, [0, 200, 400], 16, torch)]
    for s in S:
        s.connect_local(S)
    d_q = torch.from_numpy(db[:8].copy()).to(S[0].dev)
    with pytest.raises(RuntimeError):                               # keys not replicated yet
        S[0].query(d_q, phases=1)
    PeerShardedSearch.sync_keys_local(S)
    c0.scAddDescriptors(db[:3])                                     # the database grew after connect: ownership no longer matches
    with pytest.raises(RuntimeError):
        S[0].query(d_q, phases=1)
    c0.close(); c1.close()


def test_whole_batch_call_replays_from_a_graph(synth):
    """liorf_sc_shard_query_dev (the one-call batch a real rank issues) on a single rank: the third identical request (same buffers,
    same size) is captured as a CUDA graph and replayed from then on.  New query CONTENTS in the same buffers must give the new
    answers — nothing about a batch may be baked into the graph except the kernel sequence."""
    import torch
    import liorf_b200
    from liorf_b200.sc_sharded import PeerShardedSearch
    K, Q = 9000, 500
    db = synth.sc_descriptors(K, seed=81)
    ref = liorf_b200.Context(); ref.scAddDescriptors(db)
    ctx = liorf_b200.Context(); ctx.scAddDescriptors(db); ctx.scSetSearchPath(2)
    S = PeerShardedSearch(ctx, 0, 1, [0, K], Q, torch)
    S.connect_local([S])
    d_q = torch.empty((Q, 1200), dtype=torch.float64, device=S.dev)
    for b in range(6):
        qd, src, shift = synth.sc_queries(db, Q, seed=90 + b)
        with torch.cuda.stream(S.stream):
            d_q.copy_(torch.from_numpy(qd).to(S.dev))
        loop, sh, dd, cand = S.query(d_q)
        ctx.sync(); torch.cuda.synchronize()
        r_loop, r_sh, r_dist, r_cand = ref.scQueryBatch(qd)
        assert np.array_equal(cand.cpu().numpy(), r_cand), b
        assert np.array_equal(loop.cpu().numpy(), r_loop) and np.array_equal(sh.cpu().numpy(), r_sh), b
        assert np.array_equal(np.nan_to_num(dd.cpu().numpy(), nan=-7.0), np.nan_to_num(r_dist, nan=-7.0)), b
    ctx.close(); ref.close()


def test_borrowed_database_two_batches_in_flight(synth):
    """liorf_sc_borrow_database: a second context on the same device searches the owner's database without copying it, so two query batches
    can be in flight on one GPU (own stream, own scratch, own peer windows each).  Both lanes must give the reference answers while running
    concurrently, and the borrower must refuse to modify the database."""
    import torch
    import liorf_b200
    from liorf_b200.sc_sharded import PeerShardedSearch
    K, Q = 9000, 600
    db = synth.sc_descriptors(K, seed=95)
    owner = liorf_b200.Context(); owner.scAddDescriptors(db); owner.scSetSearchPath(2)
    guest = liorf_b200.Context(); guest.scBorrowDatabase(owner); guest.scSetSearchPath(2)
    with pytest.raises(liorf_b200.api.LiorfError):
        guest.scAddDescriptors(db[:3])
    lanes = []
    for c in (owner, guest):
        s = PeerShardedSearch(c, 0, 1, [0, K], Q, torch); s.connect_local([s]); lanes.append(s)
    ref = liorf_b200.Context(); ref.scAddDescriptors(db)
    qs = [synth.sc_queries(db, Q, seed=96 + b)[0] for b in range(2)]
    d_q = [torch.from_numpy(q).to(lanes[0].dev) for q in qs]
    torch.cuda.synchronize()
    for rep in range(5):                                          # the third round replays both lanes from their CUDA graphs
        outs = [lanes[k].query(d_q[k]) for k in range(2)]          # both enqueued before either is waited for
        owner.sync(); guest.sync(); torch.cuda.synchronize()
        for k in range(2):
            r_loop, r_sh, r_dist, r_cand = ref.scQueryBatch(qs[k])
            loop, sh, dd, cand = [t.cpu().numpy() for t in outs[k]]
            assert np.array_equal(cand, r_cand) and np.array_equal(loop, r_loop) and np.array_equal(sh, r_sh), (rep, k)
            assert np.array_equal(np.nan_to_num(dd, nan=-7.0), np.nan_to_num(r_dist, nan=-7.0)), (rep, k)
    guest.close(); owner.close(); ref.close()


def test_two_ranks_two_lanes_borrowed_index(synth):
    """what bench.py --gpus N runs per GPU: an owner context (database shard + replicated index) and a second lane that borrows both
    (k_total_max = 0: no key area of its own) with its own windows; the lanes answer different batches concurrently"""
    import torch
    import liorf_b200
    from liorf_b200.sc_sharded import PeerShardedSearch
    K, Q, world = 9000, 500, 2
    db = synth.sc_descriptors(K, seed=101)
    bounds = [0, 4000, 9000]
    ref = liorf_b200.Context(); ref.scAddDescriptors(db)
    owners = [liorf_b200.Context() for _ in range(world)]
    for g, c in enumerate(owners):
        c.scAddDescriptors(db[bounds[g]:bounds[g + 1]]); c.scSetSearchPath(2)
    S0 = [PeerShardedSearch(c, g, world, bounds, Q, torch) for g, c in enumerate(owners)]
    for s in S0:
        s.connect_local(S0)
    PeerShardedSearch.sync_keys_local(S0)
    for c in owners:
        c.sync()
    guests = [liorf_b200.Context() for _ in range(world)]
    for g, c in enumerate(guests):
        c.scBorrowDatabase(owners[g]); c.scSetSearchPath(2)
    S1 = [PeerShardedSearch(c, g, world, bounds, Q, torch, k_total_max=0) for g, c in enumerate(guests)]
    for s in S1:
        s.connect_local(S1)
    qs = [synth.sc_queries(db, Q, seed=102 + b)[0] for b in range(2)]
    d_q = [torch.from_numpy(q).to(S0[0].dev) for q in qs]
    torch.cuda.synchronize()
    for rep in range(2):
        res = {}
        for step in (1, 2, 4):
            for lane, S in enumerate((S0, S1)):
                for g in range(world):
                    res[(lane, g)] = S[g].query(d_q[lane], phases=step)
        for c in owners + guests:
            c.sync()
        torch.cuda.synchronize()
        for lane in range(2):
            r_loop, r_sh, r_dist, r_cand = ref.scQueryBatch(qs[lane])
            for g in range(world):
                loop, sh, dd, cand = [t.cpu().numpy() for t in res[(lane, g)]]
                assert np.array_equal(cand, r_cand) and np.array_equal(loop, r_loop) and np.array_equal(sh, r_sh) and _eq_dist(dd, r_dist), (rep, lane, g)
    for c in guests + owners:
        c.close()
    ref.close()


def test_full_scale_100k_four_shards_vs_oracle(synth, oracle):
    """BASELINE config 5 at its full database size, 4 ranks, tensor path, two batches: == oracle == unsharded, every query"""
    _run(synth, 100000, [4096, 1500], 4, 0, oracle=oracle)
