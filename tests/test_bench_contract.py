"""CPU checks (-m "not gpu") of bench.py's contract: the reference arm runs here (it is the
unmodified reference on the host cores), prints ONE JSON line with the required keys on stdout, runs
on exactly the config our arm prints, and reproduces the golden checksum of the timed workload."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1",
                        "--steps", "1", "--warmup", "1"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[-2000:]            # one JSON line, nothing else on stdout
    d = json.loads(lines[0])
    import bench
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == "set-ops/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["config"] == bench.headline_config(19900, 179100)        # the arm runs on OUR config
    assert d["value"] > 0 and abs(d["value"] - 179100 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "set-ops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    with open(os.path.join(ROOT, "tests", "golden", "allpairs_golden.json")) as f:
        gold = json.load(f)
    assert d["checksum_sum_card"] == gold["bench_checksum_and_or_xor"] and d["parity"] is True


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "1"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_strong_scaling_pair_split_is_a_partition():
    import bench
    ia, ib = bench.all_pairs(200)
    assert len(ia) == 19900 and np.all(ia < ib)
    for world in (1, 2, 3, 4, 8):
        seen = np.zeros(len(ia), dtype=np.int64)
        for rank in range(world):
            seen[rank::world] += 1                      # bench.py: pairs[rank::world] on rank `rank`
        assert np.all(seen == 1)
