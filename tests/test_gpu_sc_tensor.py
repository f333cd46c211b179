"""ScanContext ring-key search on the tensor cores (csrc/sc_tensor.cuh): the tcgen05 coarse filter + exact re-rank must
return exactly what the CUDA-core brute force and the oracle (nanoflann arithmetic, include/nanoflann.hpp:383-408,
ties by (dist, idx)) return — candidate ids AND fp32 distances bit for bit — and the raw tensor-core distances must stay
inside the error bound the completeness proof uses."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _keys(oracle, descs):
    return np.stack([oracle.sc_keys_from_desc(d)[0] for d in descs])


def _knn_dev(ctx, qkeys, offset=0):
    import ctypes as C
    import torch
    dev = torch.device(f"cuda:{ctx.params.device}")
    st = torch.cuda.ExternalStream(ctx.stream(), device=dev)
    with torch.cuda.stream(st):
        q = torch.from_numpy(np.ascontiguousarray(qkeys, np.float32)).to(dev)
        d = torch.empty((len(qkeys), 3), dtype=torch.float32, device=dev); i = torch.empty((len(qkeys), 3), dtype=torch.int32, device=dev)
        rc = ctx.lib.liorf_sc_knn_batch_dev(ctx.h, C.c_void_p(q.data_ptr()), len(qkeys), offset, C.c_void_p(d.data_ptr()), C.c_void_p(i.data_ptr()))
        assert rc == 0
    ctx.sync()
    return d.cpu().numpy(), i.cpu().numpy()


def test_tensor_distances_within_bound(ctx, oracle, synth):
    K, Q = 3000, 300                                               # neither a multiple of the 128 / 256 tiles
    db = synth.sc_descriptors(K, seed=21)
    q, src, shift = synth.sc_queries(db, Q, seed=22)
    ctx.scAddDescriptors(db)
    keys, qkeys = _keys(oracle, db), _keys(oracle, q)
    dt, center = ctx.scTensorDump(qkeys)
    assert np.allclose(center, keys.astype(np.float64).mean(0), rtol=0, atol=1e-5)
    x = qkeys.astype(np.float64) - center.astype(np.float64); y = keys.astype(np.float64) - center.astype(np.float64)
    d = ((x[:, None, :] - y[None, :, :]) ** 2).sum(-1)
    scale = (x ** 2).sum(1)[:, None] + (y ** 2).sum(1)[None, :]
    rel = np.abs(dt.astype(np.float64) - d) / scale
    print("max |d~ - d| / (|x|^2+|y|^2) = 2^%.2f" % np.log2(rel.max()))
    assert rel.max() < 2.0 ** -15                                   # the filter assumes 2^-13: at least a 4x margin


@pytest.mark.parametrize("K,Q", [(6000, 300), (129, 70), (40000, 1000), (2, 5)])
def test_tensor_topk_equals_brute_force_and_oracle(oracle, synth, K, Q):
    import liorf_b200
    ctx = liorf_b200.Context()
    db = synth.sc_descriptors(K, seed=31 + K)
    if K > 200:
        db[123] = db[77]; db[150] = db[77]                          # exact duplicate keys → ties resolved by index
    q, src, shift = synth.sc_queries(db, Q, seed=32 + K)
    if K > 200:
        q[0] = db[77]                                               # a query at distance exactly 0 from three keys
    ctx.scAddDescriptors(db)
    keys, qkeys = _keys(oracle, db), _keys(oracle, q)
    ctx.scSetSearchPath(1); bd, bi = _knn_dev(ctx, qkeys, offset=1000)
    ctx.scSetSearchPath(2); td, ti = _knn_dev(ctx, qkeys, offset=1000)
    st = ctx.scTensorStats()
    print("K=%d Q=%d candidates/query=%.1f overflow=%d" % (K, Q, st["candidates"] / Q, st["overflow"]))
    assert np.array_equal(bi, ti) and np.array_equal(bd.view(np.uint32), td.view(np.uint32))
    oi, od = oracle.ringkey_top3(keys, qkeys)
    nfill = min(K, 3)
    assert np.array_equal(ti[:, :nfill] - 1000, oi[:, :nfill]) and np.array_equal(td[:, :nfill].view(np.uint32), od[:, :nfill].view(np.uint32))
    if K > 200:
        assert list(ti[0] - 1000) == [77, 123, 150] and not td[0].any()
    assert st["overflow"] == 0
    ctx.close()


def test_tensor_overflow_falls_back_to_exact(oracle, synth):
    """more than SCT_CAP candidate chunks inside the filter band of a query (here: 3000 identical keys spread over ~160
    chunks of 32) → that query is answered by the brute-force kernel; results stay exact."""
    import liorf_b200
    ctx = liorf_b200.Context()
    db = synth.sc_descriptors(5000, seed=41)
    db[1000:4000] = db[17]
    q, _, _ = synth.sc_queries(db, 128, seed=42)
    q[5] = db[17]
    ctx.scAddDescriptors(db)
    keys, qkeys = _keys(oracle, db), _keys(oracle, q)
    ctx.scSetSearchPath(1); bd, bi = _knn_dev(ctx, qkeys)
    ctx.scSetSearchPath(2); td, ti = _knn_dev(ctx, qkeys)
    st = ctx.scTensorStats()
    assert st["overflow"] >= 1
    assert np.array_equal(bi, ti) and np.array_equal(bd.view(np.uint32), td.view(np.uint32))
    assert list(ti[5]) == [17, 1000, 1001]
    ctx.close()


def test_query_batch_uses_tensor_path_and_matches_oracle(oracle, synth):
    """liorf_sc_query_batch end to end (auto path: Q >= 64, K >= 4096 → tensor cores): loop ids, shifts, candidates."""
    import liorf_b200
    ctx = liorf_b200.Context()
    K, Q = 8000, 256
    db = synth.sc_descriptors(K, seed=51)
    q, src, shift = synth.sc_queries(db, Q, seed=52)
    ctx.scAddDescriptors(db)
    loop, sh, dist, cand = ctx.scQueryBatch(q)
    assert ctx.scTensorStats()["candidates"] >= 3 * Q               # the tensor path ran
    keys, qkeys = _keys(oracle, db), _keys(oracle, q)
    o_loop, o_sh, o_dist, o_cand = oracle.sc_query_batch(keys, db, qkeys, q)
    assert np.array_equal(cand, o_cand) and np.array_equal(loop, o_loop) and np.array_equal(sh, o_sh)
    assert np.allclose(dist, o_dist, rtol=0, atol=1e-12, equal_nan=True)
    ctx.close()


@pytest.mark.parametrize("case", ["random_walk", "all_identical", "huge_heights", "two_clusters"])
def test_tensor_filter_adversarial_databases(oracle, synth, case):
    """key distributions chosen to stress the filter's completeness argument: a drive-like random walk (neighbouring keys nearly
    equal → many near-ties), a database of identical keys (every distance ties → overflow → brute-force fallback), heights of
    ~1000 m (large norms → wide eps band) and two far-apart clusters (the database mean is far from every key).  In every case
    the tensor path must return the brute-force kernel's ids and distances bit for bit."""
    import liorf_b200
    rng = np.random.default_rng(77)
    K, Q = 7000, 200
    base = synth.sc_descriptors(K, seed=91)
    if case == "random_walk":
        db = np.empty_like(base); cur = base[0].copy()
        for i in range(K):
            cur = np.clip(cur + rng.normal(scale=0.02, size=1200) * (cur > 0), 0, 12); db[i] = cur
        q = db[rng.integers(0, K, Q)] + rng.normal(scale=0.01, size=(Q, 1200)) * (db[0] > 0)
    elif case == "all_identical":
        db = np.repeat(base[:1], K, 0); q = np.concatenate([base[:1], base[1:Q]], 0)
    elif case == "huge_heights":
        db = base * 90.0; q = db[rng.integers(0, K, Q)] + rng.normal(scale=0.5, size=(Q, 1200))
    else:
        db = base.copy(); db[K // 2:] += 500.0; q = np.concatenate([db[:Q // 2] + 0.01, db[K // 2:K // 2 + Q // 2] - 0.01], 0)
    ctx = liorf_b200.Context()
    ctx.scAddDescriptors(db)
    qkeys = _keys(oracle, q)
    ctx.scSetSearchPath(1); bd, bi = _knn_dev(ctx, qkeys)
    ctx.scSetSearchPath(2); td, ti = _knn_dev(ctx, qkeys)
    st = ctx.scTensorStats()
    print(case, "candidate chunks/query %.1f overflow %d" % (st["candidates"] / Q, st["overflow"]))
    assert np.array_equal(bi, ti) and np.array_equal(bd.view(np.uint32), td.view(np.uint32))
    if case == "all_identical":
        assert st["overflow"] >= 1 and list(ti[0]) == [0, 1, 2]
    ctx.close()
