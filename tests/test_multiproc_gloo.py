"""N>1 path on CPU: world_size-2 gloo run of the sharded ScanContext search protocol (liorf_b200/sc_sharded.py).
The four local steps are stood in by the oracle (test-only); what is under test is the orchestration the GPU arm uses:
global-index bookkeeping, the top-3 merge rule, owner-computes stage 2 and the final decision — the sharded result
must equal the unsharded one bit for bit (SURVEY §4 "multi-GPU without a cluster")."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from liorf_b200.sc_sharded import TorchPacking  # noqa: E402


class OracleOps(TorchPacking):
    """CPU stand-in for GpuOps (tests only)."""

    def __init__(self, o, keys, descs, off):
        self.o, self.keys, self.descs, self.off = o, keys, descs, off

    def knn(self, q):
        idx, d = self.o.ringkey_top3(self.keys, q["keys"])
        n = len(self.keys)
        idx = idx.astype(np.int64) + self.off
        if n < 3:                                                  # unfilled slots
            idx[:, n:] = 0x7fffffff; d[:, n:] = np.inf
        return torch.from_numpy(d.copy()), torch.from_numpy(idx.astype(np.int32))

    def merge(self, gd, gi):
        from liorf_b200.sc_sharded import merge_top3_numpy
        d, i = merge_top3_numpy(gd.numpy(), gi.numpy())
        return torch.from_numpy(d), torch.from_numpy(i)

    def distance(self, q, cand):
        c = cand.numpy(); Q = len(c)
        pd = np.full((Q, 3), np.inf); ps = np.zeros((Q, 3), np.int32)
        for qi in range(Q):
            for j in range(3):
                lc = int(c[qi, j]) - self.off
                if 0 <= lc < len(self.descs):
                    pd[qi, j], ps[qi, j] = self.o.sc_distance(q["desc"][qi], self.descs[lc])
        return torch.from_numpy(pd), torch.from_numpy(ps)

    def decide(self, pd, ps, cand):
        pd, ps, c = pd.numpy(), ps.numpy(), cand.numpy()
        Q = len(pd); loop = np.full(Q, -1, np.int32); sh = np.zeros(Q, np.int32); dd = np.zeros(Q)
        for qi in range(Q):
            mn, al, nn = 10000000.0, 0, 0
            for j in range(3):
                if pd[qi, j] < mn:
                    mn, al, nn = pd[qi, j], ps[qi, j], c[qi, j]
            loop[qi] = nn if mn < 0.3 else -1; sh[qi] = al; dd[qi] = mn
        return torch.from_numpy(loop), torch.from_numpy(sh), torch.from_numpy(dd)


def _worker(rank, world, port, K, Q, out):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle as o
    from tools import synth
    from liorf_b200.sc_sharded import ShardedScanContextSearch
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    db = synth.sc_descriptors(K, seed=11)
    db[5] = db[K - 3]                                              # duplicate keys across shards → tie on (dist, idx)
    qd, src, shift = synth.sc_queries(db, Q, seed=12)
    keys = np.stack([o.sc_keys_from_desc(d)[0] for d in db]); qkeys = np.stack([o.sc_keys_from_desc(d)[0] for d in qd])
    kloc = K // world; off = rank * kloc
    ops = OracleOps(o, keys[off:off + kloc], db[off:off + kloc], off)
    s = ShardedScanContextSearch(ops, rank, world, dist)
    loop, sh, dd, cand = s.query(dict(keys=qkeys, desc=qd, Q=Q))
    if rank == 0:
        ref = o.sc_query_batch(keys, db, qkeys, qd)
        ok = (np.array_equal(loop.numpy(), ref[0]) and np.array_equal(sh.numpy(), ref[1]) and np.array_equal(cand.numpy(), ref[3])
              and np.allclose(dd.numpy(), ref[2], atol=0, rtol=0, equal_nan=True))
        out.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_scancontext_world2_gloo():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0)); port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 600, 48, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) is True


def test_merge_rule_matches_unsharded():
    from liorf_b200.sc_sharded import merge_top3_numpy
    rng = np.random.default_rng(0)
    for G in (1, 2, 4, 8):
        Q = 64
        d = rng.integers(0, 6, size=(Q, 3 * G)).astype(np.float32)           # many ties
        idx = np.stack([rng.permutation(1000)[:3 * G] for _ in range(Q)]).astype(np.int32)
        order = np.lexsort((idx, d), axis=1); rows = np.arange(Q)[:, None]
        want_d, want_i = d[rows, order][:, :3], idx[rows, order][:, :3]
        # shard lists: each shard holds its own 3 best in sorted order
        gd = np.empty((G, Q, 3), np.float32); gi = np.empty((G, Q, 3), np.int32)
        for g in range(G):
            sd, si = d[:, 3 * g:3 * g + 3], idx[:, 3 * g:3 * g + 3]
            o2 = np.lexsort((si, sd), axis=1)
            gd[g], gi[g] = sd[rows, o2], si[rows, o2]
        md, mi = merge_top3_numpy(gd, gi)
        assert np.array_equal(md, want_d) and np.array_equal(mi, want_i)
