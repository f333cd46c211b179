"""ScanContext loop-closure search over a keyframe database sharded across ranks (BASELINE config 5).

Partitioning: database rows [g*K/G, (g+1)*K/G) live on rank g (ring keys + descriptors); queries are replicated.
Protocol per query batch (the only data-path collectives of the whole library):
  1. every rank: exact local top-3 ring-key candidates (global indices)                      [local kernel]
  2. all_gather of Q x 3 x (f32 dist, i32 idx) = 24 B per query per rank                       [NCCL / gloo]
  3. every rank: merge the G lists by (dist, idx) -> the GLOBAL top-3 (identical everywhere)   [local kernel]
  4. owner-computes: distanceBtnScanContext for the candidates this rank owns, +inf otherwise  [local kernel]
  5. all_gather of Q x 3 x (f64 dist, i32 shift); every pair is finite on exactly one rank      [NCCL / gloo]
  6. every rank: strict-< argmin in kNN order + SC_DIST_THRES                                  [local kernel]
Evaluating exactly the global top-3 (never extra local candidates) keeps results identical to the reference.

`ops` supplies the four local steps, so the same orchestration runs on the GPU library (GpuOps) and, in the CPU tests,
on a stand-in that calls the oracle (tests/test_multiproc_gloo.py)."""
import ctypes as C

import numpy as np


def merge_top3_numpy(gd, gi):
    """(G,Q,3) dist / idx -> global top-3 by (dist, idx); unfilled slots carry idx INT_MAX."""
    G, Q, _ = gd.shape
    d = np.transpose(gd, (1, 0, 2)).reshape(Q, G * 3)
    i = np.transpose(gi, (1, 0, 2)).reshape(Q, G * 3).astype(np.int64)
    d = np.where(i == 0x7fffffff, np.inf, d)
    order = np.lexsort((i, d), axis=1)[:, :3]
    rows = np.arange(Q)[:, None]
    return d[rows, order].astype(np.float32), i[rows, order].astype(np.int32)


class ShardedScanContextSearch:
    def __init__(self, ops, rank, world, dist=None):
        self.ops, self.rank, self.world, self.dist = ops, rank, world, dist

    def _all_gather(self, t):
        if self.world == 1:
            return t.unsqueeze(0)
        out = t.new_empty((self.world * t.shape[0],) + tuple(t.shape[1:]))     # concatenated layout works for NCCL and gloo
        self.dist.all_gather_into_tensor(out, t.contiguous())
        return out.view((self.world,) + tuple(t.shape))

    def query(self, q):
        """q: prepared queries (ops-specific handle).  Returns (loop_id, shift, dist, cand) as tensors of the ops' device."""
        ld, li = self.ops.knn(q)                                   # 1
        if self.world > 1:
            gd, gi = self._all_gather(ld), self._all_gather(li)    # 2
            cd, ci = self.ops.merge(gd, gi)                        # 3
        else:
            cd, ci = ld, li
        pd, ps = self.ops.distance(q, ci)                          # 4
        if self.world > 1:
            gpd, gps = self._all_gather(pd), self._all_gather(ps)  # 5
            best = gpd.argmin(dim=0, keepdim=True)
            pd, ps = gpd.gather(0, best)[0].contiguous(), gps.gather(0, best)[0].contiguous()
        return self.ops.decide(pd, ps, ci) + (ci,)                 # 6


class GpuOps:
    """The four local steps on the CUDA library; all tensors live on the context's device and stream."""

    def __init__(self, ctx, global_offset, torch):
        self.ctx, self.off, self.torch = ctx, int(global_offset), torch
        self.dev = torch.device(f"cuda:{ctx.params.device}")
        self.stream = torch.cuda.ExternalStream(ctx.stream(), device=self.dev)

    @staticmethod
    def _vp(t):
        return C.c_void_p(t.data_ptr())

    def prepare(self, qdesc_host):
        t = self.torch
        with t.cuda.stream(self.stream):
            d_q = t.from_numpy(np.ascontiguousarray(qdesc_host, np.float64).reshape(-1, 1200)).to(self.dev)
        return self.prepare_dev(d_q)

    def prepare_dev(self, d_q):
        """ring keys (a11), sector keys and column norms of query descriptors already on the device."""
        t = self.torch
        with t.cuda.stream(self.stream):
            Q = d_q.shape[0]
            keys = t.empty((Q, 20), dtype=t.float32, device=self.dev)
            sk = t.empty((Q, 60), dtype=t.float64, device=self.dev); cn = t.empty_like(sk)
            self.ctx.lib.liorf_sc_prepare_queries_dev(self.ctx.h, self._vp(d_q), Q, self._vp(keys), self._vp(sk), self._vp(cn))
        return dict(desc=d_q, keys=keys, sk=sk, cn=cn, Q=Q)

    def knn(self, q):
        t = self.torch
        with t.cuda.stream(self.stream):
            d = t.empty((q["Q"], 3), dtype=t.float32, device=self.dev); i = t.empty((q["Q"], 3), dtype=t.int32, device=self.dev)
            self.ctx.lib.liorf_sc_knn_batch_dev(self.ctx.h, self._vp(q["keys"]), q["Q"], self.off, self._vp(d), self._vp(i))
        return d, i

    def merge(self, gd, gi):
        t = self.torch
        G, Q, _ = gd.shape
        with t.cuda.stream(self.stream):
            d = t.empty((Q, 3), dtype=t.float32, device=self.dev); i = t.empty((Q, 3), dtype=t.int32, device=self.dev)
            self.ctx.lib.liorf_sc_merge_top3_dev(self.ctx.h, self._vp(gd), self._vp(gi), G, Q, self._vp(d), self._vp(i))
        return d, i

    def distance(self, q, cand):
        t = self.torch
        with t.cuda.stream(self.stream):
            pd = t.full((q["Q"], 3), float("inf"), dtype=t.float64, device=self.dev); ps = t.zeros((q["Q"], 3), dtype=t.int32, device=self.dev)
            self.ctx.lib.liorf_sc_distance_batch_dev(self.ctx.h, self._vp(q["desc"]), self._vp(q["sk"]), self._vp(q["cn"]), self._vp(cand), q["Q"], self.off,
                                                     self._vp(pd), self._vp(ps))
        return pd, ps

    def decide(self, pd, ps, cand):
        t = self.torch
        Q = pd.shape[0]
        with t.cuda.stream(self.stream):
            loop = t.empty(Q, dtype=t.int32, device=self.dev); sh = t.empty(Q, dtype=t.int32, device=self.dev); dd = t.empty(Q, dtype=t.float64, device=self.dev)
            self.ctx.lib.liorf_sc_decide_dev(self.ctx.h, self._vp(pd), self._vp(ps), self._vp(cand), Q, self._vp(loop), self._vp(sh), self._vp(dd))
        return loop, sh, dd
