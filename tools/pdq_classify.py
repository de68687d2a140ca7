#!/usr/bin/env python
"""Classify query lines against the reference: for each line of a
query file, run the compiled reference twice (plain, and with an LD_PRELOAD
rand() override that changes only the tie order) and a relational-truth
evaluator, and print one of
  PDQ-T  tie-invariant and relationally correct -- byte parity is asserted here
  PDQ-D  tie-invariant but not relational (deterministic quirk of the reference)
  TIE    output depends on the reference's random tie order: reference-undefined
  CRASH  the reference exits non-zero / dies
Usage: tools/pdq_classify.py <queries.txt> <relation files...>   (needs oracle/_ref: make -C oracle)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import workload as wl  # noqa: E402


def main():
    if len(sys.argv) < 3 or not wl.have_reference():
        print(__doc__)
        return 2
    paths = sys.argv[2:]
    db = [wl.read_relation(p) for p in paths]
    hist = {}
    for line in open(sys.argv[1]):
        line = line.strip()
        if not line or line.startswith("F"):
            continue
        cls, _ = wl.classify(paths, db, line + "\n")
        hist[cls] = hist.get(cls, 0) + 1
        print(f"{cls}\t{line}")
    print("#", hist, file=sys.stderr)
    return 0


if __name__ == "__main__":
    sys.exit(main())
