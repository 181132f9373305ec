#!/usr/bin/env python
"""Compare a reference-run file produced by gen_golden.rs with the committed oracle-generated fixture
(tests/golden/scalar_golden.json, section "party_id_beaver_source", field bn254_fr).  Exit code 0 iff every share limb agrees."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def compare(ref_path: str):
    ref = json.load(open(ref_path))
    fix = [c for c in json.load(open(os.path.join(ROOT, "tests", "golden", "scalar_golden.json")))["party_id_beaver_source"]
           if c["field"] == ref["field"]][0]
    problems = []
    for key in ("x", "y", "batch_mul"):
        for p in (0, 1):
            if ref[key][p] != fix[key][p]:
                problems.append(f"{key}[party {p}] differs")
    if ref["opened"] != fix["opened"]:
        problems.append("opened values differ")
    for p in (0, 1):
        # the fixture stores n copies of the constant triple per party as [a_batch, b_batch, c_batch]
        a, b, c = ref["triple"][p]
        for got, want in zip(fix["triples"][p], (a, b, c)):
            if any(s != want for s in got):
                problems.append(f"triple[party {p}] differs")
    return problems


if __name__ == "__main__":
    bad = compare(sys.argv[1])
    for b in bad:
        print("MISMATCH:", b)
    print("reference run == committed fixture" if not bad else f"{len(bad)} mismatch(es)")
    sys.exit(1 if bad else 0)
