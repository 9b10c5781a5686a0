#!/usr/bin/env python3
"""Localises the gap between the oracle and the only published thal anchor.

primer3-py README: calcHairpin('CCCCCATCCGATCAGGGGG') ->
    structure_found=True, tm=34.15, dg=337.09, dh=-36300.00, ds=-118.13

The structure is a 5 bp C.G stem (4 CC/GG stacks) closed over a 9 nt loop whose
first mismatch is 5'-CA-3' / 3'-GA-5'.  thal's hairpin value is therefore

    dS = 4 * stack.ds[CC/GG] + loops.ds[hairpin, size 9] + tstack2.ds[C][A][G][A]
         + SEND5 seed (-1e-11) + (N/2 - 1) * saltCorrection          (drawHairpin)

and the same sum over the .dh files.  This script prints every term from a
parameter directory (default: the embedded literature set; pass upstream's
``primer3_config/`` or set POLYOLIGO_P3_CONFIG to check the real one) and states
which entries must move for the oracle to reproduce the anchor.

Test infrastructure: uses oracle/ (never imported by the product).
"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.oracle import Oracle  # noqa: E402

SEQ = "CCCCCATCCGATCAGGGGG"
UP = dict(tm=34.15, dg=337.09, dh=-36300.00, ds=-118.13)


def q(a, b, c, d):
    i = "ACGT".index
    return ((i(a) * 5 + i(b)) * 5 + i(c)) * 5 + i(d)


def main():
    param_dir = sys.argv[1] if len(sys.argv) > 1 else os.environ.get("POLYOLIGO_P3_CONFIG")
    orc = Oracle(param_dir)
    r = orc.calcHairpin(SEQ)
    stack_h, stack_s = orc.table(0)
    ts2_h, ts2_s = orc.table(3)
    hp_h, hp_s = orc.table(8)
    salt = 0.368 * math.log(50.0 / 1000.0)      # ThermoAnalysis() defaults: mv 50, dv 0
    n_pairs_counted = 9                          # drawHairpin counts paired bases before the last base
    terms = [
        ("4 x stack[CC/GG]", 4 * stack_h[q("C", "C", "G", "G")], 4 * stack_s[q("C", "C", "G", "G")]),
        ("loops hairpin size 9", hp_h[8], hp_s[8]),
        ("tstack2[C][A][G][A]", ts2_h[q("C", "A", "G", "A")], ts2_s[q("C", "A", "G", "A")]),
        ("SEND5 seed", 0.0, -0.00000000001),
        ("(N/2-1) x saltCorrection, N=%d" % n_pairs_counted, 0.0, (n_pairs_counted // 2 - 1) * salt),
    ]
    print("tables: %s" % (param_dir or "embedded literature set"))
    print("%-34s %12s %12s" % ("term", "dH", "dS"))
    for name, h, s in terms:
        print("%-34s %12.2f %12.5f" % (name, h, s))
    sum_h, sum_s = sum(t[1] for t in terms), sum(t[2] for t in terms)
    print("%-34s %12.2f %12.5f" % ("sum of terms", sum_h, sum_s))
    print("%-34s %12.2f %12.5f   tm %.2f  dg %.2f" % ("oracle calcHairpin", r.dh, r.ds, r.tm, r.dg))
    print("%-34s %12.2f %12.5f   tm %.2f  dg %.2f" % ("upstream README", UP["dh"], UP["ds"], UP["tm"], UP["dg"]))
    assert abs(sum_h - r.dh) < 1e-6 and abs(sum_s - r.ds) < 1e-6, "decomposition does not reproduce the oracle"
    gap = UP["ds"] - r.ds
    print("\ngap (upstream - oracle): dH %+.2f, dS %+.2f e.u." % (UP["dh"] - r.dh, gap))
    if abs(gap) < 0.006 and abs(UP["dh"] - r.dh) < 0.006:
        print("=> the anchor is REPRODUCED with these tables")
        return 0
    free = terms[1][2] + terms[2][2]
    print("dH matches, and the stack and salt terms are pinned by calcTm's anchor (same SantaLucia 1998 set),")
    print("so the whole gap sits in   loops.ds[hairpin 9] + tstack2.ds[C][A][G][A] = %.2f   (must be %.2f)."
          % (free, free + gap))
    print("With SantaLucia & Hicks' hairpin-9 value (4.5 kcal/mol -> dS %.2f) kept, upstream's" % terms[1][2])
    print("tstack2.ds[C][A][G][A] would be %.2f instead of %.2f; with the literature tstack2 kept, upstream's"
          % (terms[2][2] + gap, terms[2][2]))
    print("loops.ds hairpin-9 would be %.2f.  The anchor alone cannot split the two: ONE more published" %
          (terms[1][2] + gap))
    print("hairpin value with a different loop size or closing mismatch would.")
    return 1


if __name__ == "__main__":
    sys.exit(main())
