"""Keeps the evidence trail consistent: every profiles/ file that the documents cite exists, and
the traffic figure bench.py copies into `roofline.traffic` points at a committed capture."""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cited_profile_files_exist():
    missing = []
    for doc in ("DESIGN.md", "README.md", os.path.join("profiles", "README.md")):
        with open(os.path.join(ROOT, doc)) as f:
            text = f.read()
        names = set(re.findall(r"profiles/(r\d\d[a-z]_[A-Za-z0-9_]+\.(?:json|md|txt))", text))
        if doc.startswith("profiles"):
            names |= set(re.findall(r"`(r\d\d[a-z]_[A-Za-z0-9_]+\.(?:json|md|txt))`", text))
        missing += ["%s cites %s" % (doc, n) for n in sorted(names)
                    if not os.path.exists(os.path.join(ROOT, "profiles", n))]
    assert not missing, missing


def test_traffic_json_points_at_committed_captures():
    with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
        t = json.load(f)
    for kernel, rec in t.items():
        assert os.path.exists(os.path.join(ROOT, rec["source"])), (kernel, rec["source"])
        assert rec["dram_bytes_per_launch"] > 0
