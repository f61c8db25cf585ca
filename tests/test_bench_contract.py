"""bench.py's reference arm (`--impl reference`: the reference's own CPU path from oracle/_ref on the host cores)
runs without a GPU, so its JSON line -- the keys the driver reads on both arms -- is checked here.  The b200 arm needs
a device; on a box without one it must fail loudly, not fall back."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libbq2ref.so")


def run_bench(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=300)


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/libbq2ref.so not built (needs /root/reference at build time)")
def test_reference_arm_line():
    r = run_bench("--impl", "reference", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1                                         # ONE JSON line
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "tris/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "f32" and d["scaling"] == "weak"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert d["gpu_launches"] == 0
    # throughput is consistent with the step time: 4,096 pairs per step
    assert abs(d["pairs_per_s"] * d["ms_per_step"] * 1e-3 - 4096) < 1


def test_b200_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the b200 arm is what tests/ -m gpu and the driver's bench run exercise")
    r = run_bench("--steps", "1", "--warmup", "0", "--no-cpu")
    assert r.returncode != 0                                       # no CPU fallback: the run fails, it prints no result line
    assert not [l for l in r.stdout.splitlines() if l.startswith("{") and '"value"' in l]


def test_byte_formulas_match_survey_8d():
    """SURVEY.md 8(d) at C1 / C2 (V=258, T=512): entity 2,180 R + 3,096 W; pair 9,232 R + 3,096 + 64 + 4 + 4*I W;
    `bytes = E*5,276 + E*L*(9,232 + 3,160 + 4*I + 4)`.  The layout's own minimum moves the 12-byte-per-triangle table
    once per entity and a 48-byte pair record."""
    sys.path.insert(0, ROOT)
    import bench
    E, Lt, V, T, I = 1024, 4, 258, 512, 2156
    assert bench.algorithmic_bytes(1, 0, V, T, 0) == 2180 + 3096
    assert bench.algorithmic_bytes(0, 1, V, T, I) == 9232 + 3096 + 64 + 4 * I + 4
    assert bench.algorithmic_bytes(E, E * Lt, V, T, E * Lt * I) == E * 5276 + E * Lt * (9232 + 3160 + 4 * I + 4)
    assert bench.min_traffic_bytes(1, 0, V, T, 0) == 2 * 4 * V + 128 + 12 * T + 12 * V
    assert bench.min_traffic_bytes(0, 1, V, T, I) == 48 + 12 * V + 64 + 4 + 4 * I
    assert bench.min_traffic_bytes(E, E * Lt, V, T, E * Lt * I) < bench.algorithmic_bytes(E, E * Lt, V, T, E * Lt * I)


def test_traffic_probe_parser_and_rank_pinning():
    """The in-run DRAM-traffic figure comes out of an ncu CSV log: the parser is checked on the committed round-2 capture
    (100 back-to-back C2 steps in one range: 53.2 MB per step against 52.5 MB of unique bytes), and on a per-launch log;
    every rank of an N-rank run gets its own CPUs."""
    sys.path.insert(0, ROOT)
    import bench
    rd, wr, rows = bench.parse_ncu_dram_csv(os.path.join(ROOT, "profiles", "r02_c2_traffic_range.csv"))
    assert rows == 1 and rd > 0 and wr > 0
    per_step = (rd + wr) / 100
    unique = bench.min_traffic_bytes(1024, 4096, 258, 512, 4096 * 1687.2)
    assert 0.9 < per_step / unique < 1.1
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False) as f:
        f.write('==PROF== Connected\n"ID","Kernel Name","Metric Name","Metric Unit","Metric Value"\n')
        for i in range(3):
            f.write(f'"{i}","k","dram__bytes_read.sum","byte","1,000"\n"{i}","k","dram__bytes_write.sum","byte","{2000 + i}"\n')
    assert bench.parse_ncu_dram_csv(f.name) == (3000.0, 6003.0, 3)
    os.unlink(f.name)
    sets = [set(bench.rank_cpus(r, 4)) for r in range(4)]
    assert all(sets) and (len(os.sched_getaffinity(0)) < 4 or all(not (sets[i] & sets[j]) for i in range(4) for j in range(i)))
