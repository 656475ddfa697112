"""Host-side tools (no GPU): the timeline report and bench.py's problem cache."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_trace_report_on_a_synthetic_timeline(tmp_path):
    """tools/trace_report.py turns the kernels' %globaltimer stamps ([iterations][8] ns) into per-phase medians."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import trace_report
    its = 40
    t0 = 1_000_000
    rows = []
    for i in range(its):
        s = t0 + i * 100_000                                      # 100 us per iteration
        rows.append([s, s + 50_000, s + 55_000, s + 57_000, s + 80_000, s + 84_000, s + 86_000, s + 40_000])
    path = tmp_path / "trace_c4_n2_r0.npy"
    np.save(path, np.array(rows, dtype=np.uint64))
    out = trace_report.report(str(path))
    for piece in ("spmv=50.0", "spmv_red=5.0", "gap1=2.0", "xr=23.0", "xr_red=4.0", "gap2=2.0", "d+gap3=14.0",
                  "iteration=100.0", "halo_ready_after_spmv_start=40.0"):
        assert piece in out, (piece, out)
    np.save(path, np.zeros((5, 8), dtype=np.uint64))               # nothing recorded
    assert "too few complete iterations" in trace_report.report(str(path))


def test_bench_problem_cache_round_trip(tmp_path, monkeypatch):
    """bench.make_problem under $CGB200_PROBLEM_CACHE returns the same CSR arrays and right-hand side."""
    sys.path.insert(0, ROOT)
    import bench
    monkeypatch.setenv("CGB200_PROBLEM_CACHE", str(tmp_path))
    A, B = bench.make_problem("c1", "f64", 1)
    A2, B2 = bench.make_problem("c1", "f64", 1)
    assert os.listdir(tmp_path) and (A != A2).nnz == 0 and np.array_equal(B, B2)
    assert A2.indices.dtype == np.int32 and A2.dtype == np.float64 and A2.has_sorted_indices


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the oracle port on the host cores) needs no GPU: one JSON line with the
    keys the driver reads."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "CG iters/sec" and line["unit"] == "iterations/s"
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["name"] == "c1" and line["higher_is_better"] is True
