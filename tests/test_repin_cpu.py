"""scripts/repin.py -- the one-command re-pin of thal and the picker -- is exercised here with the
oracle standing in for primer3 (``--stub``) and with a parameter DIRECTORY in upstream's
primer3_config/ file format (``POLYOLIGO_P3_CONFIG``), so the day the real library or its tables
are available the command is known to work."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPIN = os.path.join(ROOT, "scripts", "repin.py")


def _run(tmp_path, extra_env, *flags):
    env = {k: v for k, v in os.environ.items() if k not in ("POLYOLIGO_P3_CONFIG", "POLYOLIGO_GOLDEN_P3")}
    env.update(extra_env)
    return subprocess.run([sys.executable, REPIN, "--n", "120", "--markers", "2", "--threads", "4", "--no-pytest",
                           "--out", str(tmp_path)] + list(flags), env=env, capture_output=True, text=True,
                          timeout=600)


def test_repin_without_a_source_reports_unpinned(tmp_path):
    r = _run(tmp_path, {})
    assert r.returncode == 2
    assert "PARITY UNPINNED" in r.stdout


def test_repin_stub_round_trip(tmp_path):
    r = _run(tmp_path, {}, "--stub")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "FAIL" not in r.stdout and r.stdout.count("PASS") >= 6
    for kind in ("any", "end1", "end2", "hairpin", "tailed"):
        z = np.load(os.path.join(tmp_path, "thal_%s.npz" % kind))
        assert z["res"].shape == (120, 5)
    with open(os.path.join(tmp_path, "design_kasp.json")) as f:
        d = json.load(f)
    assert len(d["cases"]) == 2 and all("pairs" in c for c in d["cases"])
    with open(os.path.join(tmp_path, "source.json")) as f:
        assert json.load(f)["pins"] == "nothing (stub)"


def test_repin_with_a_table_directory(tmp_path):
    params = os.path.join(ROOT, "polyoligo_b200", "params")   # upstream primer3_config/ file format
    r = _run(tmp_path, {"POLYOLIGO_P3_CONFIG": params})
    assert r.returncode == 0, r.stdout + r.stderr
    assert "table constants only" in r.stdout and "FAIL" not in r.stdout
