"""ScanContext loop-closure search over a keyframe database sharded across ranks (BASELINE config 5) — the Python side of
liorf_sc_shard_* (csrc/sc_shard.cuh).

"Replicated index, sharded payload": rank g holds the descriptors / sector keys / column norms of rows [row_begin[g], row_begin[g+1]);
the 80-byte ring key of every row is replicated (pushed over NVLink once, `sync_keys`).  Per batch every rank runs stage 1 (ring-key
top-3, include/Scancontext.cpp:289-295) for ITS SLICE of the queries against all keys and stage 2 (distanceBtnScanContext, :302-317) for
the candidates IT OWNS; the two exchanges are done by the library's own kernels through peer-memory windows (plain stores over NVLink +
system-scope flags), one library call per batch, no collective launch, no host round trip.  torch.distributed is used once, to hand the
64-byte cudaIpc handles around."""
import ctypes as C


class PeerShardedSearch:
    def __init__(self, ctx, rank, world, row_begin, q_max, torch, k_total_max=None):
        """row_begin: world + 1 global row indices — rank g holds database rows [row_begin[g], row_begin[g + 1]).
        k_total_max: capacity of the replicated ring-key index (None = row_begin[-1]; 0 = this context borrowed the database AND the
        index of a synced context on the same device, see Context.scBorrowDatabase)."""
        self.ctx, self.rank, self.world, self.q_max, self.torch = ctx, rank, world, int(q_max), torch
        self.row_begin = [int(v) for v in row_begin]
        assert len(self.row_begin) == world + 1
        self.off = self.row_begin[rank]
        self.k_cap = self.row_begin[-1] if k_total_max is None else int(k_total_max)
        self.dev = torch.device(f"cuda:{ctx.params.device}")
        self.stream = torch.cuda.ExternalStream(ctx.stream(), device=self.dev)
        self.handle = (C.c_ubyte * 64)()
        self.window = C.c_void_p()
        rc = ctx.lib.liorf_sc_shard_init(ctx.h, C.c_int(rank), C.c_int(world), C.c_int(self.q_max), C.c_int(self.k_cap), self.handle, C.byref(self.window))
        if rc < 0:
            raise RuntimeError(f"liorf_sc_shard_init failed with code {rc}")
        self._out = {}

    def _rows(self):
        return (C.c_int * (self.world + 1))(*self.row_begin)

    def connect_processes(self, dist):
        """peers live in other processes (one per GPU): exchange the cudaIpc handles, map the windows, replicate the ring keys"""
        if self.world > 1:
            hs = [None] * self.world
            dist.all_gather_object(hs, bytes(self.handle))
            buf = (C.c_ubyte * (64 * self.world)).from_buffer_copy(b"".join(hs))
            rc = self.ctx.lib.liorf_sc_shard_connect(self.ctx.h, buf, None, self._rows())
        else:
            rc = self.ctx.lib.liorf_sc_shard_connect(self.ctx.h, self.handle, None, self._rows())
        if rc < 0:
            raise RuntimeError(f"liorf_sc_shard_connect failed with code {rc}")
        if self.world > 1:
            dist.barrier()                                         # every window is mapped everywhere before anybody pushes into one
            if self.k_cap > 0:
                self.sync_keys()
            dist.barrier()

    def connect_local(self, searches):
        """peers are other contexts of THIS process (tests: several shards on one GPU); call sync_keys_local(searches) afterwards"""
        ptrs = (C.c_void_p * self.world)(*[s.window.value for s in searches])
        rc = self.ctx.lib.liorf_sc_shard_connect(self.ctx.h, None, ptrs, self._rows())
        if rc < 0:
            raise RuntimeError(f"liorf_sc_shard_connect failed with code {rc}")

    def sync_keys(self, phases=3):
        rc = self.ctx.lib.liorf_sc_shard_sync_keys_phases(self.ctx.h, C.c_int(phases))
        if rc < 0:
            raise RuntimeError(f"liorf_sc_shard_sync_keys failed with code {rc}")

    @staticmethod
    def sync_keys_local(searches):
        """ranks sharing one device: every push is enqueued before any wait"""
        if searches[0].world == 1:
            return
        for s in searches:
            s.sync_keys(1)
        for s in searches:
            s.sync_keys(2)

    def wait_stats(self):
        """(ns waited for the peers per phase {C, D, KEYS} since the last call, batch counter)"""
        w = (C.c_ulonglong * 4)(); b = C.c_uint(0)
        self.ctx.lib.liorf_sc_shard_wait_stats(self.ctx.h, w, C.byref(b))
        return dict(C=w[0], D=w[1], KEYS=w[2]), b.value

    def query(self, d_q, phases=7):
        """d_q: (Q, 1200) float64 query descriptors on this rank's device (the same on every rank).  Asynchronous on the context's
        stream.  Returns (loop_id, shift, dist, cand) device tensors.  phases: bit mask of the batch's three steps (tests that put
        several ranks on one device enqueue step by step over the ranks; a real rank passes 7)."""
        t = self.torch
        Q = int(d_q.shape[0])
        o = self._out.get(Q)
        if o is None:
            o = self._out[Q] = (t.empty(Q, dtype=t.int32, device=self.dev), t.empty(Q, dtype=t.int32, device=self.dev),
                                t.empty(Q, dtype=t.float64, device=self.dev), t.empty((Q, 3), dtype=t.int32, device=self.dev))
        args = (self.ctx.h, C.c_void_p(d_q.data_ptr()), C.c_int(Q), C.c_int(self.off), C.c_void_p(o[0].data_ptr()), C.c_void_p(o[1].data_ptr()),
                C.c_void_p(o[2].data_ptr()), C.c_void_p(o[3].data_ptr()))
        if phases == 7:
            rc = self.ctx.lib.liorf_sc_shard_query_dev(*args)           # the whole batch (a CUDA graph replay from the third identical request on)
        else:
            rc = self.ctx.lib.liorf_sc_shard_query_phases_dev(*args, C.c_int(phases))
        if rc < 0:
            raise RuntimeError(f"liorf_sc_shard_query_dev failed with code {rc}")
        return o

    def query_host(self, pin_q, host_out):
        """the batch from HOST memory (liorf_sc_shard_query_async): pin_q = pinned (Q, 1200) float64 tensor, host_out = pinned tensors
        (loop int32[Q], shift int32[Q], dist float64[Q], cand int32[Q, 3]); asynchronous, results valid after ctx.sync()"""
        Q = int(pin_q.shape[0])
        rc = self.ctx.lib.liorf_sc_shard_query_async(self.ctx.h, C.c_void_p(pin_q.data_ptr()), C.c_int(Q), C.c_int(self.off), C.c_void_p(host_out[0].data_ptr()),
                                                     C.c_void_p(host_out[1].data_ptr()), C.c_void_p(host_out[2].data_ptr()), C.c_void_p(host_out[3].data_ptr()))
        if rc < 0:
            raise RuntimeError(f"liorf_sc_shard_query_async failed with code {rc}")
        return host_out

    def debug_state(self):
        """counters and exchange flags of this rank's window (debugging)"""
        o = (C.c_uint * 68)()
        self.ctx.lib.liorf_sc_shard_debug_state(self.ctx.h, o)
        return dict(batch=o[0], raise_counter=o[1], owned=o[2], err=o[3], flags={g: tuple(o[4 + 4 * g + p] for p in range(3)) for g in range(self.world)})
