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


def _lookup(d, dotted):
    for part in dotted.split("."):
        d = d[int(part)] if isinstance(d, list) else d[part]
    return d


def test_quoted_numbers_match_the_committed_bench_lines():
    """profiles/README.md carries a table `| file | key | quoted |`: every number DESIGN.md / README.md
    quote from a bench line is listed there and must equal the committed JSON within the rounding
    shown (a cited file once held a different experiment than the text claimed)."""
    with open(os.path.join(ROOT, "profiles", "README.md")) as f:
        text = f.read()
    rows = re.findall(r"^\| `(r\d\d[a-z]_[A-Za-z0-9_]+\.json)` \| `([A-Za-z0-9_.]+)` \| ([0-9.eE+-]+) \|", text, re.M)
    assert len(rows) >= 8, "the quoted-numbers table is missing"
    for name, key, quoted in rows:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            lines = [ln for ln in f if ln.startswith("{")]
        got = float(_lookup(json.loads(lines[-1]), key))
        want = float(quoted)
        digits = len(quoted.split(".")[1]) if "." in quoted else 0
        assert abs(got - want) <= 0.6 * 10 ** (-digits) * max(1.0, 0.0) or abs(got - want) <= 0.006 * abs(want), (name, key, got, quoted)
