"""the reference's own collation script, UNMODIFIED, run on a `whisper_b200_cli` output tree.

tests/golden/cli_sweep/gpu_1g/ is what `N_SYNTH=256 GPUS_LIST=1 scripts/run_gpu_benchmarks.sh` left on a B200
(`without_hf_pipeline_rust/` = the three result files of the Rust SUT, `logs/without_hf_pipeline_rust.time.txt` = the
two `/usr/bin/time -v` lines the script parses).  /root/reference/compare_container_benchmarks.py reads that tree in a
subprocess and must emit the table row of the Rust implementation with our numbers in it.  The schema half of the test
(what the script looks for: `latency_end_to_end_s.p95`, `config_used`) runs everywhere; the subprocess half only where
the reference checkout exists (this container, not the GPU box)."""
import csv
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TREE = os.path.join(ROOT, "tests", "golden", "cli_sweep", "gpu_1g")
SCRIPT = "/root/reference/compare_container_benchmarks.py"
RUST_ROW = "onnxruntime rust (no HF pipeline)"


def _summary():
    return json.load(open(os.path.join(TREE, "without_hf_pipeline_rust", "inference_summary.json")))


def test_tree_has_what_the_collation_script_reads():
    s = _summary()
    blk = s["latency_end_to_end_s"]
    assert list(blk.keys()) == ["max", "mean", "median", "min", "p90", "p95"] and blk["p95"] > 0
    assert isinstance(s["config_used"], dict) and s["n_files"] == 256 and s["max_new_tokens"] == 128
    rows = json.load(open(os.path.join(TREE, "without_hf_pipeline_rust", "inference_per_file.json")))
    assert len(rows) == 256 and [r["file"] for r in rows] == sorted(r["file"] for r in rows)
    with open(os.path.join(TREE, "without_hf_pipeline_rust", "inference_per_file.csv"), newline="") as f:
        table = list(csv.reader(f))
    assert table[0] == ["file", "duration_s", "end_to_end_s", "rtf", "text"] and len(table) == 257
    assert all(abs(float(t[2]) - r["end_to_end_s"]) < 5e-5 for t, r in zip(table[1:], rows))      # "{:.4}" of the same number
    log = open(os.path.join(TREE, "logs", "without_hf_pipeline_rust.time.txt")).read()
    assert "Elapsed (wall clock) time" in log and "Maximum resident set size (kbytes)" in log


@pytest.mark.skipif(not os.path.exists(SCRIPT), reason="reference checkout not present (GPU box)")
def test_reference_collation_script_runs_unmodified_on_our_tree(tmp_path):
    md, out_csv = tmp_path / "summary_table.md", tmp_path / "summary_table.csv"
    r = subprocess.run([sys.executable, SCRIPT, "--results-dir", TREE, "--log-dir", os.path.join(TREE, "logs"),
                        "--out-md", str(md), "--out-csv", str(out_csv)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "Wrote summary table:" in r.stdout
    with open(out_csv, newline="") as f:
        rows = {row["implementation"]: row for row in csv.DictReader(f)}
    assert len(rows) == 6                                           # the script's six fixed implementations
    ours = rows[RUST_ROW]
    p95 = _summary()["latency_end_to_end_s"]["p95"]
    assert float(ours["time_s"]) == round(p95, 3)                   # "Time" = latency_end_to_end_s.p95 (script :100-115, :195)
    assert ours["precision"] == "fp32" and ours["beam_size"] == "1"  # the script's fall-backs: config_used carries neither key
    rss_kb = int([l for l in open(os.path.join(TREE, "logs", "without_hf_pipeline_rust.time.txt")) if "Maximum resident" in l][0].split(":")[1])
    assert int(ours["ram_mb"]) == round(rss_kb / 1024.0)
    for name, row in rows.items():                                  # implementations we do not provide stay empty, not wrong
        if name != RUST_ROW:
            assert row["time_s"] == "" and row["ram_mb"] == ""
    line = [l for l in md.read_text().splitlines() if RUST_ROW in l][0]
    assert line.split("|")[4].strip() not in ("", "n/a") and line.split("|")[5].strip().endswith("MB")
    # and the same script on the reference's own committed tree gives the table the reference committed (the script
    # itself is what we think it is)
    ref_tree = "/root/reference/results.old/benchmarks/container_4c4g/epyc-9654"
    if os.path.isdir(ref_tree):
        md2, csv2 = tmp_path / "ref.md", tmp_path / "ref.csv"
        r2 = subprocess.run([sys.executable, SCRIPT, "--results-dir", ref_tree, "--log-dir", os.path.join(ref_tree, "logs"),
                             "--out-md", str(md2), "--out-csv", str(csv2)], capture_output=True, text=True, timeout=120)
        assert r2.returncode == 0, r2.stderr
        committed = os.path.join(ref_tree, "summary_table.csv")
        if os.path.exists(committed):
            assert csv2.read_text().replace("\r\n", "\n") == open(committed).read().replace("\r\n", "\n")
