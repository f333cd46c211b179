"""ImageProjection::imuDeskewInfo (src/imageProjection.cpp:350-409) — liorf_host_imu_deskew_info (host-side scalar code of the C ABI, no GPU
needed) against the oracle's restatement over a std::deque (separately written) and against the numpy generator the synthetic sequences
use (tools/synth: imu_table).  Bit for bit: the tables feed findRotation's fp64 interpolation."""
import numpy as np
import pytest


def _stream(rng, t0, t1, rate, jitter=0.0):
    n = int((t1 - t0) * rate) + 1
    t = t0 + np.arange(n) / rate + (rng.uniform(-jitter, jitter, n) if jitter else 0.0)
    return np.sort(t), rng.normal(0, 0.3, (n, 3))


@pytest.mark.parametrize("rate", [100.0, 200.0, 500.0])
def test_matches_oracle_and_generator(oracle, synth, rate):
    import liorf_b200
    rng = np.random.default_rng(int(rate))
    for trial in range(40):
        cur = 1000.0 + 0.1 * trial + rng.uniform(0, 0.01); end = cur + rng.uniform(0.08, 0.1)
        stamp, gyro = _stream(rng, cur - 0.3, end + 0.3, rate, jitter=0.2 / rate if trial % 2 else 0.0)
        g = liorf_b200.imuDeskewInfo(stamp, gyro, cur, end)
        o = oracle.imu_deskew_info(stamp, gyro, cur, end)
        assert g["available"] == o["available"] is True
        assert g["imu_pointer_cur"] == o["imu_pointer_cur"] and g["n_pop"] == len(stamp) - o["queue_left"]
        assert np.array_equal(g["imu_time"].view(np.int64), o["imu_time"].view(np.int64))
        assert np.array_equal(g["imu_rot"].view(np.int64), o["imu_rot"].view(np.int64))
        # rows: first stamp >= cur - 0.01 ... last stamp <= end + 0.01; row 0 is all zeros
        assert g["imu_time"][0] >= cur - 0.01 and g["imu_time"][-1] <= end + 0.01 and not g["imu_rot"][0].any()
        assert abs(len(g["imu_time"]) - (end - cur + 0.02) * rate) <= 2               # ~12 rows at 100 Hz, ~22 at 200 Hz, ~52 at 500 Hz
        in_win = np.nonzero((stamp >= cur - 0.01) & (stamp <= cur))[0]            # :371-375 looks only at what survived the pop
        assert g["rpy_index"] == (in_win[-1] if len(in_win) else -1)
    # constant-rate stream on the generator's time lattice: the synthetic sequences' tables (numpy) come out of the same recurrence
    omega = np.array([0.02, -0.03, 1.0])
    cur, end = 20.0, 20.0987
    it, rot, ptr = synth.imu_table(cur, end, omega, rate_hz=rate)
    k = np.arange(int(np.floor((cur - 0.5) * rate)), int(np.ceil((end + 0.5) * rate)))
    stamp = k * (1.0 / rate)                                              # the generator's lattice (k * dt)
    g = liorf_b200.imuDeskewInfo(stamp, np.tile(omega, (len(stamp), 1)), cur, end)
    assert g["imu_pointer_cur"] == ptr and np.array_equal(g["imu_time"], it) and np.array_equal(g["imu_rot"], rot)


def test_gate_and_degenerate_queues(oracle):
    import liorf_b200
    cur, end = 50.0, 50.1
    stamp = 50.0 + np.arange(-5, 16) * 0.01
    gyro = np.ones((len(stamp), 3))
    # deskewInfo's gate (:337): the queue must bracket the scan
    assert not liorf_b200.imuDeskewInfo(stamp[stamp > cur], gyro[stamp > cur], cur, end)["available"]              # first sample after timeScanCur
    assert not liorf_b200.imuDeskewInfo(stamp[stamp < end - 0.02], gyro[stamp < end - 0.02], cur, end)["available"]  # last sample before timeScanEnd
    assert not liorf_b200.imuDeskewInfo(np.zeros(0), np.zeros((0, 3)), cur, end)["available"]
    # without the gate: everything popped / a single usable sample → imuPointerCur <= 0 → not available (:403-408), as the oracle
    for sel in (stamp < cur - 0.02, (stamp > end + 0.02), np.isclose(stamp, 50.0)):
        g = liorf_b200.imuDeskewInfo(stamp[sel], gyro[sel], cur, end, check_gate=False)
        o = oracle.imu_deskew_info(stamp[sel], gyro[sel], cur, end)
        assert g["available"] == o["available"] is False and g["imu_pointer_cur"] == o["imu_pointer_cur"]
    # a sample exactly at timeScanCur - 0.01 stays (strict <, :356); one exactly at timeScanEnd + 0.01 is still integrated (strict >, :378)
    s2 = np.array([cur - 0.01, cur, end, end + 0.01, end + 0.011])
    g = liorf_b200.imuDeskewInfo(s2, np.ones((5, 3)), cur, end)
    o = oracle.imu_deskew_info(s2, np.ones((5, 3)), cur, end)
    assert g["n_pop"] == 0 and g["imu_pointer_cur"] == o["imu_pointer_cur"] == 3 and np.array_equal(g["imu_rot"], o["imu_rot"])
    # queueLength = 2000 rows (:62): the reference would overrun its arrays; the library refuses
    dense = np.linspace(cur, end, 2500)
    with pytest.raises(liorf_b200.api.LiorfError):
        liorf_b200.imuDeskewInfo(dense, np.zeros((2500, 3)), cur, end)
