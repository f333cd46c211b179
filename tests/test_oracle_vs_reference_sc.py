"""Pins oracle/liorf_oracle.hpp's ScanContext restatement (a10-a14) against the REFERENCE'S OWN SOURCE: /root/reference/include/Scancontext.cpp
compiled unchanged into oracle/_ref/libliorf_ref_sc.so against the header stand-ins of oracle/shim (Eigen / PCL / OpenCV are not in the image;
recipe: oracle/Makefile).  Bit for bit: descriptors, ring keys, sector keys, distanceBtnScanContext (distance bits and shift), and the
whole detectLoopClosureID sequence incl. the stale tree.  Reduction order inside mean / norm / dot is the shim's (sequential) — stated in
DESIGN.md §5; everything else (control flow, indexing, thresholds, float/double mixes, kd-tree, candidate order) is the reference's."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def refsc(oracle):
    if oracle.refsc() is None:
        pytest.skip("oracle/_ref/libliorf_ref_sc.so was not built (the reference tree is absent and no prebuilt copy travelled)")
    return oracle


def _cloud(rng, n, kind):
    if kind == 0:                                    # street-like: ground + walls, some returns beyond 80 m, some below the sensor
        r = rng.uniform(0.5, 95.0, n); th = rng.uniform(-np.pi, np.pi, n)
        z = np.where(rng.random(n) < 0.6, -1.73 + rng.normal(0, 0.02, n), rng.uniform(-2.5, 9.0, n))
    elif kind == 1:                                  # points ON sector / ring boundaries (multiples of 6 degrees, multiples of 4 m)
        th = np.deg2rad(rng.integers(0, 60, n) * 6.0) + rng.choice([0.0, 1e-7, -1e-7], n)
        r = rng.integers(1, 21, n) * 4.0 + rng.choice([0.0, 1e-5, -1e-5], n); z = rng.uniform(-3, 5, n)
    else:                                            # sparse: most bins empty, axis-aligned points (x = 0 or y = 0)
        r = rng.uniform(1, 70, n); th = rng.choice([0, np.pi / 2, np.pi, -np.pi / 2, 0.3, 2.0], n); z = rng.uniform(-1999, 30, n) * (rng.random(n) < 0.9)
    p = np.zeros((n, 4), np.float32)
    p[:, 0] = r * np.cos(th); p[:, 1] = r * np.sin(th); p[:, 2] = z
    return p


def test_make_scancontext_and_keys_bit_exact_1200_clouds(refsc):
    o = refsc
    ref = o.RefSCManager()
    rng = np.random.default_rng(2024)
    for i in range(1200):
        pts = _cloud(rng, int(rng.integers(1, 3000)), i % 3)
        d0, rk0, sk0 = ref.make(pts)
        d1, rk1, sk1 = o.sc_make(pts)
        assert np.array_equal(d0.view(np.int64), np.asarray(d1).view(np.int64)), i
        assert np.array_equal(rk0.view(np.int32), rk1.view(np.int32)) and np.array_equal(sk0.view(np.int64), sk1.view(np.int64)), i
    # empty cloud: all bins NO_POINT → 0
    d0, rk0, sk0 = ref.make(np.zeros((0, 4), np.float32)); d1, rk1, sk1 = o.sc_make(np.zeros((0, 4), np.float32))
    assert np.array_equal(d0, d1) and not d0.any()


def test_xy2theta_boundaries(refsc):
    """the angle that decides the sector: the oracle and the device evaluate atan on the float quotient in double — what THIS compile of the
    reference source does too (only <cmath> in scope: unqualified atan is ::atan(double)); see DESIGN.md §4 for the <math.h> caveat"""
    o = refsc
    rng = np.random.default_rng(5)
    xs = np.concatenate([rng.normal(0, 30, 4000), [0, 0, 1, -1, 1e-30, -1e-30, 3.0, -3.0]]).astype(np.float32)
    ys = np.concatenate([rng.normal(0, 30, 4000), [1, -1, 0, 0, 1, 1, 3.0 * np.tan(np.deg2rad(6.0)), 3.0 * np.tan(np.deg2rad(6.0))]]).astype(np.float32)
    one = np.zeros((1, 4), np.float32)
    for x, y in zip(xs, ys):
        t_ref = o.refsc().refsc_xy2theta(float(x), float(y))
        one[0, :3] = (x, y, 1.0)
        d, _, _ = o.sc_make(one)
        r = np.float32(np.sqrt(np.float32(x * x + y * y)))
        if r > 80 or np.isnan(t_ref):
            continue
        sector = max(min(60, int(np.ceil((np.float64(t_ref) / 360.0) * 60))), 1) - 1
        assert d[:, sector].max() == 3.0, (x, y, t_ref)               # the oracle put the point into the sector the reference's angle names


def test_distance_btn_scancontext_bit_exact(refsc, synth):
    o = refsc
    ref = o.RefSCManager()
    rng = np.random.default_rng(9)
    db = synth.sc_descriptors(300, seed=33)
    for i in range(600):
        a = db[rng.integers(0, 300)].reshape(20, 60).copy()
        kind = i % 4
        if kind == 0:
            b = np.roll(a, int(rng.integers(0, 60)), axis=1) + rng.normal(0, 0.05, a.shape) * (a != 0)
        elif kind == 1:
            b = db[rng.integers(0, 300)].reshape(20, 60).copy()
        elif kind == 2:
            b = a.copy(); b[:, rng.integers(0, 60, 25)] = 0.0; a[:, rng.integers(0, 60, 25)] = 0.0     # empty columns on both sides
        else:
            b = np.zeros_like(a) if i % 8 == 3 else a.copy()                                               # all-zero overlap (NaN) / identical
        d0, s0 = ref.distance(a, b); d1, s1 = o.sc_distance(a, b)
        assert s0 == s1, (i, s0, s1)
        assert (np.isnan(d0) and np.isnan(d1)) or np.float64(d0).view(np.int64) == np.float64(d1).view(np.int64), (i, d0, d1)


def test_detect_loop_closure_sequence_with_stale_tree(refsc, synth):
    """descriptors arrive one by one, detectLoopClosureID after every one (the tree is rebuilt on calls 0, 10, 20 ... past the 31-entry
    early-out and is stale in between): loop id and yaw of EVERY call equal the reference's, revisits planted along the way"""
    o = refsc
    ref, orc = o.RefSCManager(), o.SCManager()
    rng = np.random.default_rng(77)
    base = synth.sc_descriptors(400, seed=55)
    found = 0
    for i in range(400):
        if i > 60 and i % 7 == 0:                                      # a revisit of an older place, rotated and noisy
            j = int(rng.integers(0, i - 40))
            d = np.roll(base[j].reshape(20, 60), int(rng.integers(0, 60)), axis=1); d = d + rng.normal(0, 0.05, d.shape) * (d != 0)
        elif i > 100 and i % 11 == 0:
            d = base[i - 1].reshape(20, 60)                           # a near-duplicate of a RECENT entry (inside the 30-entry exclusion zone) — NOT an exact one:
            d = d + rng.normal(0, 1e-3, d.shape) * (d != 0)           # two identical ring keys tie exactly, and the reference's order of exact ties is whatever
                                                                      # nanoflann's traversal yields (SURVEY trap 15); the oracle canonicalises ties to (dist, idx)
        else:
            d = base[i].reshape(20, 60)
        ref.save_descriptor(d); orc.save_descriptor(d)
        l0, y0 = ref.detect(); l1, y1, _, _ = orc.detect()
        assert l0 == l1 and np.float32(y0).view(np.int32) == np.float32(y1).view(np.int32), (i, l0, l1, y0, y1)
        found += l0 >= 0
    assert found > 20 and ref.size() == orc.size() == 400
    for i in (0, 17, 399):
        d0, k0 = ref.get(i); d1, k1 = orc.get(i)
        assert np.array_equal(d0, d1) and np.array_equal(k0.view(np.int32), k1.view(np.int32))


def test_batch_through_the_reference_detect_loop_closure_id(refsc, synth):
    """bench.py --impl reference at N > 1: every query of a batch through the reference's own detectLoopClosureID, unchanged (the harness appends the query the way
    makeAndSaveScancontextAndKeys stores a keyframe and removes it again) — loop ids and shifts equal the oracle's batch search (what the GPU search is held to)."""
    import pyoracle as o
    K, Q = 4000, 96
    db = synth.sc_descriptors(K, seed=21)
    qd, src, shift = synth.sc_queries(db, Q, seed=22)
    R = o.RefSCManager()
    R.save_descriptors(db)
    loop, sh = R.query_batch(qd)
    assert R.size() == K                                                # the harness leaves the database as it found it
    keys = o.sc_keys_batch(db); qk = o.sc_keys_batch(qd)
    l2, s2, _, _ = o.sc_query_batch(keys, db, qk, qd)
    assert np.array_equal(loop, l2) and np.array_equal(sh, s2)
    planted = src >= 0
    assert planted.sum() >= Q // 3 and np.array_equal(loop[planted], src[planted]) and np.array_equal(sh[planted] % 60, shift[planted] % 60)
    loop_b, sh_b = R.query_batch(qd[:10])                               # a second batch: same answers (tree rebuilt in its first call)
    assert np.array_equal(loop_b, loop[:10]) and np.array_equal(sh_b, sh[:10])
