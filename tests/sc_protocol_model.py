"""CPU model of the sharded ScanContext search (csrc/sc_shard.cuh) for the world_size-2 gloo test: the same decomposition — replicated
ring-key index, stage 1 split by query slice, stage 2 split by the owner of the candidate row, two exchanges (phase C, phase D) — with the
exchanges done by torch.distributed all_gather and the local steps by the oracle.  Test infrastructure only."""
import numpy as np
import torch


class ShardedSearchModel:
    def __init__(self, o, rank, world, row_begin, descs_local, keys_local, dist):
        """descs_local / keys_local: rows [row_begin[rank], row_begin[rank+1]) of the database"""
        self.o, self.rank, self.world, self.rows, self.dist = o, rank, world, [int(v) for v in row_begin], dist
        self.descs = descs_local
        # the replicated index: every rank's ring keys, gathered once (k_scsh_push_keys on the GPU)
        parts = [None] * world
        dist.all_gather_object(parts, np.ascontiguousarray(keys_local, np.float32))
        self.keys_all = np.concatenate(parts, 0)
        assert len(self.keys_all) == self.rows[-1]

    def _gather(self, t):
        out = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return out

    def query(self, qdesc, qkeys):
        o, G, r = self.o, self.world, self.rank
        Q = len(qdesc)
        q0, q1 = Q * r // G, Q * (r + 1) // G
        # stage 1 for this rank's query slice against ALL keys: the global top-3 directly
        cd = np.full((Q, 3), np.inf, np.float32); ci = np.full((Q, 3), 0x7fffffff, np.int32)
        if q1 > q0:
            idx, d = o.ringkey_top3(self.keys_all, qkeys[q0:q1])
            n = len(self.keys_all)
            idx = idx.astype(np.int32)
            if n < 3:
                idx[:, n:] = 0x7fffffff; d[:, n:] = np.inf
            cd[q0:q1], ci[q0:q1] = d, idx
        # phase C: every rank's slice rows → everyone (disjoint rows of ONE array: the element-wise minimum over ranks reassembles it)
        gd = torch.stack(self._gather(torch.from_numpy(cd))).numpy(); gi = torch.stack(self._gather(torch.from_numpy(ci))).numpy()
        for g in range(G):
            a, b = Q * g // G, Q * (g + 1) // G
            cd[a:b], ci[a:b] = gd[g, a:b], gi[g, a:b]
        # stage 2: owner computes
        lo, hi = self.rows[r], self.rows[r + 1]
        pd = np.full((Q, 3), np.inf); ps = np.zeros((Q, 3), np.int32); own = np.zeros((Q, 3), np.int32)
        for qi in range(Q):
            for j in range(3):
                c = int(ci[qi, j])
                if c != 0x7fffffff and lo <= c < hi:
                    pd[qi, j], ps[qi, j] = o.sc_distance(qdesc[qi], self.descs[c - lo]); own[qi, j] = 1
        # phase D: entry [pair] is written by exactly one rank
        gpd = torch.stack(self._gather(torch.from_numpy(pd))).numpy(); gps = torch.stack(self._gather(torch.from_numpy(ps))).numpy()
        gown = torch.stack(self._gather(torch.from_numpy(own))).numpy()
        assert (gown.sum(0) <= 1).all()
        who = gown.argmax(0)
        rows, cols = np.indices((Q, 3))
        pd = np.where(gown.sum(0) == 1, gpd[who, rows, cols], np.inf); ps = np.where(gown.sum(0) == 1, gps[who, rows, cols], 0)
        # decision (include/Scancontext.cpp:302-340)
        loop = np.full(Q, -1, np.int32); sh = np.zeros(Q, np.int32); dd = np.zeros(Q)
        for qi in range(Q):
            mn, al, nn = 10000000.0, 0, 0
            for j in range(3):
                if pd[qi, j] < mn:
                    mn, al, nn = pd[qi, j], ps[qi, j], ci[qi, j]
            loop[qi] = nn if mn < 0.3 else -1; sh[qi] = al; dd[qi] = mn
        return loop, sh, dd, ci
