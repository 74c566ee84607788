"""Writes tests/golden/assoc_golden.json by EXECUTING the reference's `association` / `greedy_relational_association`
(lib/utils.py:461-526, :134-182; definitions extracted from the source file by oracle/assoc.py::reference_functions)
on the seeded clips of i2vsgg_b200/synth.py::clip_records.   python tests/golden/make_assoc_golden.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from i2vsgg_b200 import synth  # noqa: E402
from oracle import assoc  # noqa: E402

CASES = {"plain": dict(seed=1, frames=40), "gaps": dict(seed=2, frames=60, empty=(0, 7, 8, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 59)),
         "long": dict(seed=3, frames=300, tracks=20, clutter=30)}
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assoc_golden.json")


def main():
    ref = assoc.reference_functions()
    g = {}
    for name, kw in CASES.items():
        rec, cnt = synth.clip_records(**kw)
        rels = ref["association"]({"vid": synth.records_to_frame_relations(rec, cnt)})
        g[name] = rels.get("vid", [])
        print(name, len(g[name]), "relations; longest", max([r["duration"][1] - r["duration"][0] for r in g[name]] or [0]))
    json.dump(g, open(OUT, "w"))
    print("wrote", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
