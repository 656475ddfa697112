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


def test_clock_samples_are_taken_from_the_timed_region_or_the_post_roll():
    """bench.summarise_clock_rows: nvidia-smi rows carry a timestamp; only samples inside the timed region count, and when
    the region was shorter than the sampling period (sharded runs at 4 and 8 GPUs) the post-roll under the same load
    stands in -- and the `clocks` object says so."""
    import datetime
    sys.path.insert(0, ROOT)
    import bench
    base = datetime.datetime(2026, 10, 18, 22, 30, 15, 0)

    def row(ms, sm, power, cap="Not Active"):
        ts = (base + datetime.timedelta(milliseconds=ms)).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]
        return [ts, "0", str(sm), "1965", f"{power:.2f}", "0x0000000000000004", "Not Active", "Not Active", "Not Active", cap]
    t = base.timestamp()
    rows = [row(0, 345, 140.0), row(100, 1965, 700.0), row(200, 1800, 735.0, "Active"), row(300, 1800, 736.0, "Active"),
            row(400, 1965, 400.0), ["garbage"], row(500, 1950, 690.0)]
    c = bench.summarise_clock_rows(rows, window=(t + 0.15, t + 0.35))
    assert c["samples"] == 2 and c["sm_mhz"] == 1800 and c["sm_max_mhz"] == 1965 and c["reasons"] == ["sw_power_cap"]
    assert c["power_w_max"] == 736.0 and c["sampled_in"] == "timed region"
    # a 40 ms region between two samples: nothing inside -> the post-roll
    c = bench.summarise_clock_rows(rows, window=(t + 0.13, t + 0.17), post=(t + 0.45, t + 0.55))
    assert c["samples"] == 1 and c["sm_mhz"] == 1950 and c["sampled_in"].startswith("post-roll")
    # no window at all (older callers): every row; no rows: the empty object
    assert bench.summarise_clock_rows(rows)["samples"] == 6
    assert bench.summarise_clock_rows([], window=(t, t + 1)) == {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    # the tool's timestamps an hour off the host clock: the last samples before the stop stand in, and the object says so
    c = bench.summarise_clock_rows(rows, window=(t + 3600.0, t + 3600.2))
    assert c["samples"] == 2 and c["sm_mhz"] == 1957.5 and "did not match" in c["sampled_in"]
