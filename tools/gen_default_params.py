#!/usr/bin/env python3
"""Generate the default thermodynamic parameter set in the upstream
``primer3_config/`` file format.

Why this exists: the reference (polyoligo) gets all of its thermodynamics from
primer3-py 0.6.0 (``/root/reference/setup.py:62``), whose bundled libprimer3
reads ``stack.ds/.dh``, ``stackmm.ds/.dh``, ``dangle.ds/.dh``, ``loops.ds/.dh``,
``tstack.dh``, ``tstack_tm_inf.ds``, ``tstack2.ds/.dh``, ``triloop.ds/.dh`` and
``tetraloop.ds/.dh``.  None of these files is vendored under ``/root/reference``
and there is no network, so this script rebuilds a default set from the
published literature the upstream tables are derived from.  Tables are DATA:
both the CPU oracle and the CUDA library parse these files (or the copy
embedded at build time), so dropping the real upstream directory in via
``po_init(param_dir=...)`` replaces every constant below without a rebuild.

Provenance of each table (cal/mol for dH, cal/(mol K) for dS):

* ``stack``    Watson-Crick nearest neighbours, SantaLucia (1998) PNAS 95:1460
               unified set (same numbers libprimer3's ``oligotm`` uses).
* ``stackmm``  single internal mismatches: Allawi & SantaLucia (1997) Biochem
               36:10581 (G.T), (1998) Biochem 37:2170 (G.A), (1998) NAR 26:2694
               (C.T), (1998) Biochem 37:9435 (A.C); Peyret et al. (1999) Biochem
               38:3468 (A.A, C.C, G.G, T.T).
* ``dangle``   Bommarito, Peyret & SantaLucia (2000) NAR 28:1929.
* ``tstack2``  terminal mismatches, SantaLucia & Peyret (2001) WO 01/94611.
* ``tstack``   interior-loop terminal mismatches: same set as ``tstack2``
               (upstream ships a separate ``tstack.dh`` / ``tstack_tm_inf.ds``
               pair whose values could not be recovered offline).
* ``loops``    SantaLucia & Hicks (2004) Annu Rev Biophys 33:415, table 4:
               dH = 0, dS = -dG37 / 310.15 truncated to 2 decimals; sizes the
               table does not list are filled with the Jacobson-Stockmayer
               extrapolation dS(n) = dS(x) - 2.44 R ln(n/x) from the largest
               tabulated x < n.
* ``triloop``  SantaLucia & Hicks (2004) supplementary triloop bonuses.
* ``tetraloop`` EMPTY by default: the upstream list (~100 hexamers) could not
               be recovered offline.  The lookup machinery is implemented and
               tested with synthetic tables.

Run:  python tools/gen_default_params.py
Writes ``polyoligo_b200/params/*`` and the embedded copies
``polyoligo_b200/csrc/default_params.inc`` and ``oracle/default_params.inc``.
"""
import math
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = "ACGT"
COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def wc(a, b):
    return COMP[a] == b


# -- Watson-Crick stacks, key = top-strand dinucleotide 5'->3' (kcal, e.u.) ----
NN = {
    "AA": (-7.9, -22.2), "AC": (-8.4, -22.4), "AG": (-7.8, -21.0), "AT": (-7.2, -20.4),
    "CA": (-8.5, -22.7), "CC": (-8.0, -19.9), "CG": (-10.6, -27.2), "CT": (-7.8, -21.0),
    "GA": (-8.2, -22.2), "GC": (-9.8, -24.4), "GG": (-8.0, -19.9), "GT": (-8.4, -22.4),
    "TA": (-7.2, -21.3), "TC": (-8.2, -22.2), "TG": (-8.5, -22.7), "TT": (-7.9, -22.2),
}

# -- single internal mismatches: key "ab/cd" = 5'-ab-3' over 3'-cd-5' ----------
IMM = {
    # G.T
    "AG/TT": (1.0, 0.9), "AT/TG": (-2.5, -8.3), "CG/GT": (-4.1, -11.7), "CT/GG": (-2.8, -8.0),
    "GG/CT": (3.3, 10.4), "GT/CG": (-4.4, -12.3), "TG/AT": (-0.1, -1.7), "TT/AG": (-1.3, -5.3),
    # G.A
    "AA/TG": (-0.6, -2.3), "AG/TA": (-0.7, -2.3), "CA/GG": (-0.7, -2.3), "CG/GA": (-4.0, -13.2),
    "GA/CG": (-0.6, -1.0), "GG/CA": (0.5, 3.2), "TA/AG": (0.7, 0.7), "TG/AA": (3.0, 7.4),
    # C.T
    "AC/TT": (0.7, 0.2), "AT/TC": (-1.2, -6.2), "CC/GT": (-0.8, -4.5), "CT/GC": (-1.5, -6.1),
    "GC/CT": (2.3, 5.4), "GT/CC": (5.2, 13.5), "TC/AT": (1.2, 0.7), "TT/AC": (1.0, 0.7),
    # A.C
    "AA/TC": (2.3, 4.6), "AC/TA": (5.3, 14.6), "CA/GC": (1.9, 3.7), "CC/GA": (0.6, -0.6),
    "GA/CC": (5.2, 14.2), "GC/CA": (-0.7, -3.8), "TA/AC": (3.4, 8.0), "TC/AA": (7.6, 20.2),
    # A.A, C.C, G.G, T.T
    "AA/TA": (1.2, 1.7), "CA/GA": (-0.9, -4.2), "GA/CA": (-2.9, -9.8), "TA/AA": (4.7, 12.9),
    "AC/TC": (0.0, -4.4), "CC/GC": (-1.5, -7.2), "GC/CC": (3.6, 8.9), "TC/AC": (6.1, 16.4),
    "AG/TG": (-3.1, -9.5), "CG/GG": (-4.9, -15.3), "GG/CG": (-6.0, -15.8), "TG/AG": (1.6, 3.6),
    "AT/TT": (-2.7, -10.8), "CT/GT": (-5.0, -15.8), "GT/CT": (-2.2, -8.4), "TT/AT": (0.2, -1.5),
}

# -- terminal mismatches: same key convention, first pair WC, second mismatched -
TMM = {
    "AA/TA": (-3.1, -7.8), "TA/AA": (-2.5, -6.3), "CA/GA": (-4.3, -10.7), "GA/CA": (-8.0, -22.5),
    "AC/TC": (-0.1, 0.5), "TC/AC": (-0.7, -1.3), "CC/GC": (-2.1, -5.1), "GC/CC": (-3.9, -10.6),
    "AG/TG": (-1.1, -2.1), "TG/AG": (-1.1, -2.7), "CG/GG": (-3.8, -9.5), "GG/CG": (-0.7, -19.2),
    "AT/TT": (-2.4, -6.5), "TT/AT": (-3.2, -8.9), "CT/GT": (-6.1, -16.9), "GT/CT": (-7.4, -21.2),
    "AA/TC": (-1.6, -4.0), "AC/TA": (-1.8, -3.8), "CA/GC": (-2.6, -5.9), "CC/GA": (-2.7, -6.0),
    "GA/CC": (-5.0, -13.8), "GC/CA": (-3.2, -7.1), "TA/AC": (-2.3, -5.9), "TC/AA": (-2.7, -7.0),
    "AC/TT": (-0.9, -1.7), "AT/TC": (-2.3, -6.3), "CC/GT": (-3.2, -8.0), "CT/GC": (-3.9, -10.6),
    "GC/CT": (-4.9, -13.5), "GT/CC": (-3.0, -7.8), "TC/AT": (-2.5, -6.3), "TT/AC": (-0.7, -1.2),
    "AA/TG": (-1.9, -4.4), "AG/TA": (-2.5, -5.9), "CA/GG": (-3.9, -9.6), "CG/GA": (-6.0, -15.5),
    "GA/CG": (-4.3, -11.1), "GG/CA": (-4.6, -11.4), "TA/AG": (-2.0, -4.7), "TG/AA": (-2.4, -5.8),
    "AG/TT": (-3.2, -8.7), "AT/TG": (-3.5, -9.4), "CG/GT": (-3.8, -9.0), "CT/GG": (-6.6, -18.7),
    "GG/CT": (-5.7, -15.9), "GT/CG": (-5.9, -16.1), "TG/AT": (-3.9, -10.5), "TT/AG": (-3.6, -9.8),
}

# -- dangling ends: "XY/.Z" = 5' dangling X on top; ".Y/XZ" = 3' dangling X on bottom
DE = {
    "AA/.T": (0.2, 2.3), "AC/.G": (-6.3, -17.1), "AG/.C": (-3.7, -10.0), "AT/.A": (-2.9, -7.6),
    "CA/.T": (0.6, 3.3), "CC/.G": (-4.4, -12.6), "CG/.C": (-4.0, -11.9), "CT/.A": (-4.1, -13.0),
    "GA/.T": (-1.1, -1.6), "GC/.G": (-5.1, -14.0), "GG/.C": (-3.9, -10.9), "GT/.A": (-4.2, -15.0),
    "TA/.T": (-6.9, -20.0), "TC/.G": (-4.0, -10.9), "TG/.C": (-4.9, -13.8), "TT/.A": (-0.2, -0.5),
    ".A/AT": (-0.7, -0.8), ".C/AG": (-2.1, -3.9), ".G/AC": (-5.9, -16.5), ".T/AA": (-0.5, -1.1),
    ".A/CT": (4.4, 14.9), ".C/CG": (-0.2, -0.1), ".G/CC": (-2.6, -7.4), ".T/CA": (4.7, 14.2),
    ".A/GT": (-1.6, -3.6), ".C/GG": (-3.9, -11.2), ".G/GC": (-3.2, -10.4), ".T/GA": (-4.1, -13.1),
    ".A/TT": (2.9, 10.4), ".C/TG": (-4.4, -13.1), ".G/TC": (-5.2, -15.0), ".T/TA": (-3.8, -12.6),
}

# -- loop free energies at 37 C (kcal/mol): size -> (interior, bulge, hairpin) --
LOOP_DG = {
    1: (None, 4.0, None), 2: (None, 2.9, None),
    3: (3.2, 3.1, 3.5), 4: (3.6, 3.2, 3.5), 5: (4.0, 3.3, 3.3), 6: (4.4, 3.5, 4.0),
    7: (4.6, 3.7, 4.2), 8: (4.8, 3.9, 4.3), 9: (4.9, 4.1, 4.5), 10: (4.9, 4.3, 4.6),
    12: (5.2, 4.5, 5.0), 14: (5.4, 4.8, 5.1), 16: (5.6, 5.0, 5.3), 18: (5.8, 5.2, 5.5),
    20: (5.9, 5.3, 5.7), 25: (6.3, 5.6, 6.1), 30: (6.6, 5.9, 6.3),
}

TRILOOP_DH = {
    "AGAAT": -1500, "AGCAT": -1500, "AGGAT": -1500, "AGTAT": -1500,
    "CGAAG": -2000, "CGCAG": -2000, "CGGAG": -2000, "CGTAG": -2000,
    "GGAAC": -2000, "GGCAC": -2000, "GGGAC": -2000, "GGTAC": -2000,
    "TGAAA": -1500, "TGCAA": -1500, "TGGAA": -1500, "TGTAA": -1500,
}


def rot(key):
    """Same duplex read from the other strand: 'ab/cd' -> 'dc/ba'."""
    top, bot = key.split("/")
    return bot[::-1] + "/" + top[::-1]


def fmt(v):
    if v is None or (isinstance(v, float) and math.isinf(v)):
        return "inf"
    s = ("%.2f" % v).rstrip("0").rstrip(".")
    return "0" if s in ("-0", "") else s


def quad_table(fn):
    """256 lines, nested a, b, c, d over ACGT: entry for 5'-ab-3' / 3'-cd-5'."""
    dh, ds = [], []
    for a in B:
        for b in B:
            for c in B:
                for d in B:
                    v = fn(a, b, c, d)
                    dh.append(fmt(None if v is None else v[0] * 1000.0))
                    ds.append(fmt(None if v is None else v[1]))
    return dh, ds


def stack_fn(a, b, c, d):
    return NN[a + b] if wc(a, c) and wc(b, d) else None


def mm_fn(table):
    def fn(a, b, c, d):
        key = a + b + "/" + c + d
        if wc(a, c) and not wc(b, d):
            assert key in table, key
            return table[key]
        return None
    return fn


def imm_fn(a, b, c, d):
    # stackint2[a][b][c][d]: exactly one of the two positions mismatched
    key = a + b + "/" + c + d
    if wc(a, c) != wc(b, d):
        v = IMM.get(key) or IMM.get(rot(key))
        assert v is not None, key
        return v
    return None


def dangle_tables():
    """dangle file = dangle3 block (a, c, X) then dangle5 block (a, c, Y), 64 lines each.

    dangle3[a][X][c]: pair a.c, X follows a on strand 1 (3' dangling end)
                      5'-aX-3' / 3'-c.-5'  ==  '.c/Xa' read from the other strand
    dangle5[a][c][Y]: pair a.c, Y follows c on the 3'->5' strand (5' dangling end)
                      5'-a.-3' / 3'-cY-5'  ==  'Yc/.a'
    """
    dh, ds = [], []
    for a in B:
        for c in B:
            for x in B:
                v = DE.get("." + c + "/" + x + a) if wc(a, c) else None
                dh.append(fmt(None if v is None else v[0] * 1000.0))
                ds.append(fmt(None if v is None else v[1]))
    for a in B:
        for c in B:
            for y in B:
                v = DE.get(y + c + "/." + a) if wc(a, c) else None
                dh.append(fmt(None if v is None else v[0] * 1000.0))
                ds.append(fmt(None if v is None else v[1]))
    return dh, ds


def loop_tables():
    R = 1.987
    cols = []
    for col in range(3):
        known = {n: v[col] for n, v in LOOP_DG.items() if v[col] is not None}
        ds = {}
        for n in range(1, 31):
            if n in known:
                ds[n] = math.trunc(-known[n] * 1000.0 / 310.15 * 100.0) / 100.0
            else:
                lower = [x for x in known if x < n]
                if not lower:
                    ds[n] = None
                else:
                    x = max(lower)
                    ds[n] = round(ds[x] - 2.44 * R * math.log(n / x), 2)
        cols.append(ds)
    ds_lines, dh_lines = [], []
    for n in range(1, 31):
        ds_lines.append("%d\t%s\t%s\t%s" % (n, fmt(cols[0][n]), fmt(cols[1][n]), fmt(cols[2][n])))
        dh_lines.append("%d\t%s\t%s\t%s" % (n, *["inf" if cols[c][n] is None else "0" for c in range(3)]))
    return dh_lines, ds_lines


def build():
    files = {}
    files["stack.dh"], files["stack.ds"] = quad_table(stack_fn)
    files["stackmm.dh"], files["stackmm.ds"] = quad_table(imm_fn)
    files["tstack2.dh"], files["tstack2.ds"] = quad_table(mm_fn(TMM))
    files["tstack.dh"], files["tstack_tm_inf.ds"] = quad_table(mm_fn(TMM))
    files["dangle.dh"], files["dangle.ds"] = dangle_tables()
    files["loops.dh"], files["loops.ds"] = loop_tables()
    files["triloop.dh"] = ["%s\t%d" % (k, v) for k, v in sorted(TRILOOP_DH.items())]
    files["triloop.ds"] = ["%s\t0" % k for k in sorted(TRILOOP_DH)]
    files["tetraloop.dh"] = []
    files["tetraloop.ds"] = []
    return files


def c_string(lines):
    body = "".join(l + "\n" for l in lines)
    out = []
    for i in range(0, len(body), 4000):   # stay below compilers' literal limits
        chunk = body[i:i + 4000].replace("\\", "\\\\").replace('"', '\\"').replace("\n", "\\n").replace("\t", "\\t")
        out.append('"%s"' % chunk)
    return "\n    ".join(out) if out else '""'


def main():
    files = build()
    pdir = os.path.join(ROOT, "polyoligo_b200", "params")
    os.makedirs(pdir, exist_ok=True)
    for name, lines in files.items():
        with open(os.path.join(pdir, name), "w") as f:
            f.write("".join(l + "\n" for l in lines))
    inc = ["/* GENERATED by tools/gen_default_params.py -- do not edit. */",
           "/* name, contents pairs in the upstream primer3_config file format. */"]
    for name in sorted(files):
        inc.append('{"%s",\n    %s},' % (name, c_string(files[name])))
    text = "\n".join(inc) + "\n"
    for dst in (os.path.join(ROOT, "polyoligo_b200", "csrc", "default_params.inc"),
                os.path.join(ROOT, "oracle", "default_params.inc")):
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        with open(dst, "w") as f:
            f.write(text)
    print("wrote %d parameter files to %s" % (len(files), pdir))


if __name__ == "__main__":
    main()
