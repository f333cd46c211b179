"""ScanContext search over a database sharded across ranks and exchanged through peer-memory windows (csrc/sc_shard.cuh,
liorf_sc_shard_*; SURVEY §8e, BASELINE config 5).  Here the ranks are several contexts on ONE GPU that map each other's
windows by pointer (across processes the same windows are mapped through cudaIpc handles — bench.py --gpus N); the kernels,
the flags and the protocol are the ones the multi-GPU run uses.  Loop ids, shifts, fp64 distances and the candidate triples
must equal the unsharded search (include/Scancontext.cpp:253-344 semantics) bit for bit, on every rank, batch after batch."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(synth, K, Q_list, world, path, bounds=None):
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
        c.scAddDescriptors(db[bounds[g]:bounds[g + 1]]); c.scSetSearchPath(path)
    S = [PeerShardedSearch(c, g, world, bounds, max(Q_list), torch) for g, c in enumerate(ctxs)]
    for s in S:
        s.connect_local(S)
    dev = S[0].dev
    for b, Q in enumerate(Q_list):
        qd, src, shift = synth.sc_queries(db, Q, seed=72 + b)
        if K > 300:
            qd[0] = db[5]
        r_loop, r_sh, r_dist, r_cand = ref.scQueryBatch(qd)
        outs = []
        order = list(range(world)) if b % 2 == 0 else list(range(world - 1, -1, -1))      # enqueue order must not matter
        res, dq = {}, {}
        for g in order:
            with torch.cuda.stream(S[g].stream):
                dq[g] = torch.from_numpy(qd).to(dev)
        # the ranks share ONE device here: enqueue step by step over the ranks, so that a kernel that waits for a peer's push is
        # never queued in front of the kernel that pushes (separate GPUs run the four steps as one call)
        for step in (16, 1, 2, 4, 8):
            for g in order:
                res[g] = S[g].query(dq[g], phases=step)
        for c in ctxs:
            c.sync()
        torch.cuda.synchronize()
        for g in range(world):
            loop, sh, dd, cand = [t.cpu().numpy() for t in res[g]]
            assert np.array_equal(cand, r_cand), (b, g)
            assert np.array_equal(loop, r_loop) and np.array_equal(sh, r_sh), (b, g)
            assert np.array_equal(dd, r_dist) or np.array_equal(np.isnan(dd), np.isnan(r_dist)) and np.array_equal(dd[~np.isnan(dd)], r_dist[~np.isnan(r_dist)]), (b, g)
        planted = src >= 0
        planted[0] = False                                             # query 0 was replaced above
        assert planted.sum() > 0 and np.array_equal(r_loop[planted], src[planted])
    stats = [c.scTensorStats() for c in ctxs]
    for c in ctxs:
        c.close()
    ref.close()
    return stats


def test_three_shards_tensor_path(synth):
    """tcgen05 filter + exact re-rank on every shard, global candidate threshold from the phase-T exchange; four batches of
    different sizes through the same windows (flags carry the batch number, slots are reused)."""
    stats = _run(synth, 13000, [333, 1000, 64, 777], 3, 2)
    print("candidate chunks per rank of the last batch:", [s["candidates"] for s in stats])


def test_two_uneven_shards_brute_path(synth):
    """CUDA-core exact search on the shards (no phase T), shards of very different size (one holds 5 rows)"""
    _run(synth, 2005, [50, 129], 2, 1, bounds=[0, 2000, 2005])


def test_single_rank_degenerate(synth):
    _run(synth, 5000, [200], 1, 2)


def test_global_threshold_shards_the_rerank(synth):
    """the point of phase T: with the GLOBAL threshold the candidates a rank re-ranks shrink with its share of the database —
    4 shards together emit about as many candidate chunks as one unsharded search, not 4 times as many"""
    import liorf_b200
    K, Q = 40000, 1000
    db = synth.sc_descriptors(K, seed=71 + K)
    qd, _, _ = synth.sc_queries(db, Q, seed=72)
    one = liorf_b200.Context(); one.scAddDescriptors(db); one.scSetSearchPath(2); one.scQueryBatch(qd)
    base = one.scTensorStats()["candidates"]; one.close()
    stats = _run(synth, K, [Q], 4, 2)
    total = sum(s["candidates"] for s in stats)
    print(f"candidate chunks: unsharded {base}, 4 shards together {total}")
    assert total < 1.6 * base


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
