"""CPU checks of the measurement helpers that feed bench.py (no GPU): the ncu launch-list summariser must find the full
hot-path step (flow + decoder + tail) in the committed CSV and reproduce the committed traffic figures."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_make_traffic_reproduces_committed_summary(tmp_path):
    src = os.path.join(ROOT, "profiles", "r02_launches.csv")
    out = tmp_path / "traffic.json"
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_traffic.py"), src, str(out)], check=True,
                   capture_output=True)
    got = json.load(open(out))
    ref = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    assert got["launches_in_step"] == ref["launches_in_step"] == 78
    assert got["conv_launches"] == ref["conv_launches"] == 76
    assert abs(got["conv_dram_bytes_per_launch"] - ref["conv_dram_bytes_per_launch"]) < 1.0
    assert abs(got["tail_dram_bytes_per_launch"] - ref["tail_dram_bytes_per_launch"]) < 1.0
