"""The numbers under profiles/ can be recomputed from the files next to them ("every number in `roofline` can be recomputed from
a file in profiles/"): roofline arithmetic of the committed bench lines, the captures traffic.json names, the files profiles/README.md lists.
CPU-only bookkeeping checks — nothing here measures anything."""
import json
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def _line(name):
    rows = [json.loads(l) for l in open(os.path.join(P, name)) if l.strip().startswith("{")]
    return rows[-1]


def test_headline_roofline_arithmetic():
    d = _line("r02_bench_n1.json")
    r, cfg = d["roofline"], d["config"]
    assert d["metric"] == "scan2map_ms_per_frame_64beam" and cfg["workload"].startswith("kitti64_single") and cfg["lm_iters"] == 30
    alg = 96.0 * cfg["n_ds"] * 30                                   # 16 B query point + 5 x 16 B neighbours per query and iteration
    assert r["algorithmic_bytes_per_launch"] == alg
    assert r["achieved"] == pytest.approx(alg / (r["avg_launch_ms"] * 1e-3) / 1e9, rel=1e-9)
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9)
    assert r["share_of_step"] == pytest.approx(r["avg_launch_ms"] / d["ms_per_step"], rel=1e-6)
    assert d["knn_queries_per_s"] == pytest.approx(30.0 * cfg["n_ds"] / (r["avg_launch_ms"] * 1e-3), rel=1e-9)
    t = json.load(open(os.path.join(P, "traffic.json")))["scan2map.kitti64_single"]
    assert r["traffic"] == pytest.approx(t["dram_bytes_per_launch"], rel=0.01)          # the same capture, re-taken once after the line was written
    assert r["l2_bytes_per_launch"] == pytest.approx(t["l2_bytes_per_launch"], rel=0.01)
    assert d["e2e"]["value"] > d["value"] and d["e2e"]["h2d_bytes_per_step"] == cfg["n_scan"] * 16 + 4
    assert d["value"] < d["target_ms"]


def test_traffic_json_names_committed_captures():
    t = json.load(open(os.path.join(P, "traffic.json")))
    for key, e in t.items():
        if key.startswith("_"):
            continue
        cap = os.path.join(ROOT, e["capture"])
        assert os.path.exists(cap), (key, e["capture"])
        s = json.load(open(cap))
        scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}
        total = sum(float(s[m]["value"]) * scale[s[m]["unit"].lower()] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        assert e["dram_bytes_per_launch"] == pytest.approx(total, rel=1e-3), key
        if "l2_bytes_per_launch" in e and "lts__t_sectors.sum" in s:
            assert e["l2_bytes_per_launch"] == pytest.approx(float(s["lts__t_sectors.sum"]["value"]) * 32, rel=1e-3), key


@pytest.mark.parametrize("n", [2, 4, 8])
def test_sharded_lines_are_bit_equal_and_same_lanes(n):
    d = _line("r02_bench_n%d.json" % n)
    s = d["sc"]
    assert d["metric"] == "sc_queries_per_s_100k" and d["n_gpus"] == n and s["shards"] == n and d["scaling"] == "strong"
    assert s["bit_equal_unsharded"] is True
    assert s["batches_in_flight"] == s["unsharded_same_run"]["batches_in_flight"]       # the same lanes at every N
    assert d["value"] == pytest.approx(s["Q"] / (d["ms_per_step"] * 1e-3), rel=1e-6)
    assert s["planted_loops_found"] == s["planted"]


def test_readme_lists_only_files_that_exist():
    txt = open(os.path.join(P, "README.md")).read()
    names = set(re.findall(r"`((?:r0[12]_|traffic)[A-Za-z0-9_.{},*]+)`", txt))
    have = set(os.listdir(P))
    for nm in names:
        if "*" in nm:
            assert any(h.startswith(nm.split("*")[0]) for h in have), nm
        elif "{" in nm:
            pre, rest = nm.split("{", 1)
            alts, post = rest.split("}", 1)
            for a in alts.split(","):
                assert pre + a + post in have, pre + a + post
        else:
            assert nm in have, nm
