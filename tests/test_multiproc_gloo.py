"""N>1 path on CPU: world_size-2 gloo run of the sharded ScanContext search's decomposition (tests/sc_protocol_model.py mirrors
csrc/sc_shard.cuh: replicated ring-key index, stage 1 by query slice, stage 2 by owner, two exchanges).  The local steps are stood in by
the oracle (test-only); what is under test is the bookkeeping the GPU arm relies on — slices, global indices, ownership, uneven and
empty shards, the final decision: the sharded result must equal the unsharded one bit for bit (SURVEY §4 "multi-GPU without a cluster")."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _worker(rank, world, port, K, Q, bounds, out):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import pyoracle as o
    from tools import synth
    from sc_protocol_model import ShardedSearchModel
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    db = synth.sc_descriptors(K, seed=11)
    db[5] = db[K - 3]                                              # duplicate keys across shards → tie on (dist, idx)
    qd, src, shift = synth.sc_queries(db, Q, seed=12)
    keys = np.stack([o.sc_keys_from_desc(d)[0] for d in db]); qkeys = np.stack([o.sc_keys_from_desc(d)[0] for d in qd])
    lo, hi = bounds[rank], bounds[rank + 1]
    s = ShardedSearchModel(o, rank, world, bounds, db[lo:hi], keys[lo:hi], dist)
    loop, sh, dd, cand = s.query(qd, qkeys)
    ref = o.sc_query_batch(keys, db, qkeys, qd)
    ok = (np.array_equal(loop, ref[0]) and np.array_equal(sh, ref[1]) and np.array_equal(cand, ref[3]) and np.allclose(dd, ref[2], atol=0, rtol=0, equal_nan=True))
    out.put((rank, bool(ok)))                                      # every rank holds the complete, identical answer
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("K,Q,bounds", [(600, 48, [0, 300, 600]), (601, 37, [0, 596, 601]), (300, 1, [0, 300, 300])])
def test_sharded_scancontext_world2_gloo(K, Q, bounds):
    """even shards; uneven shards with an odd query count; one EMPTY shard and fewer queries than ranks (an empty slice)"""
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0)); port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, K, Q, bounds, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = dict(out.get(timeout=5) for _ in range(2))
    assert got == {0: True, 1: True}
