#!/usr/bin/env python3
"""ONE command that (re)pins thal and the primer picker against the real primer3.

    python scripts/repin.py                      # needs `import primer3` (primer3-py) to work
    POLYOLIGO_P3_CONFIG=/path/to/primer3_config python scripts/repin.py
    python scripts/repin.py --stub --out /tmp/x  # CI exercise: the CPU oracle stands in for primer3

What it does
  1. picks the SOURCE of truth:
       * real primer3-py when it imports (ThermoAnalysis + bindings.designPrimers);
       * else upstream's parameter tables when POLYOLIGO_P3_CONFIG names a primer3_config/
         directory: the CPU oracle evaluated on THOSE tables (pins every constant, the algorithm
         is still the restatement);
       * else, only with --stub, the oracle on the embedded literature tables (exercises the
         machinery; pins nothing).
  2. dumps golden vectors under --out (default tests/golden/primer3/):
       thal_{any,end1,end2,hairpin,tailed}.npz  >= --n results per thal type (default 10^4):
            config-4 shaped random 18-30 nt pairs; `tailed` = 21-nt VIC/FAM reporter + 18-25 nt
            against 18-25 nt (reference cli_kasp.py:65,72, _lib_kasp.py:95-105)
       design_kasp.json   designPrimers answers for --markers (default 200) config-5 markers,
            one fixed allele-specific primer each, reference KASP settings at depth 100
       source.json        what produced them
  3. checks the CPU oracle and -- when a GPU is visible -- the CUDA library against the vectors
     at north_star's tolerance (Tm 0.01 C, dG 1e-3 kcal/mol = 1 cal/mol, structure_found and
     selected primer sequences identical) and prints one PASS / FAIL line per set;
  4. reruns the parity suite (pytest tests/, GPU tests included when a GPU is visible) with
     POLYOLIGO_GOLDEN_P3=<out>, so tests/test_primer3_golden.py gates on the new vectors.

Exit code 0 = everything agrees; 1 = some set differs (the report says which); 2 = no source.

Test infrastructure: imports oracle/; nothing in the product imports this file.
"""
import argparse
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.oracle import Oracle, ANY, END1, END2, HAIRPIN  # noqa: E402
from oracle import p3_oracle  # noqa: E402

VIC = "GAAGGTCGGAGTCAACGGATT"   # reference src/polyoligo/cli_kasp.py:65
FAM = "GAAGGTGACCAAGTTCATGCT"   # reference src/polyoligo/cli_kasp.py:72
TOL_TM, TOL_DG = 0.01, 1.0      # deg C; cal/mol (= 1e-3 kcal/mol)

# reference src/polyoligo/data/PRIMER3_KASP.yaml + _lib_kasp.py:654-656 (PRIMER_NUM_RETURN = depth)
KASP_GLOBALS = {
    "PRIMER_MIN_SIZE": 18, "PRIMER_OPT_SIZE": 20, "PRIMER_MAX_SIZE": 25,
    "PRIMER_MIN_TM": 55.0, "PRIMER_OPT_TM": 60.0, "PRIMER_MAX_TM": 65.0,
    "PRIMER_MIN_GC": 20.0, "PRIMER_MAX_GC": 80.0, "PRIMER_MAX_POLY_X": 100,
    "PRIMER_SALT_MONOVALENT": 50.0, "PRIMER_DNA_CONC": 50.0, "PRIMER_MAX_NS_ACCEPTED": 0,
    "PRIMER_PAIR_MAX_DIFF_TM": 6.0, "PRIMER_PRODUCT_SIZE_RANGE": [[50, 100], [100, 150]],
    "PRIMER_NUM_RETURN": 100,
}


# ---------------------------------------------------------------------------------------------
# sources
# ---------------------------------------------------------------------------------------------
class Primer3Source:
    """Real primer3-py (any version that has ThermoAnalysis and bindings.designPrimers)."""

    def __init__(self, primer3):
        from primer3.thermoanalysis import ThermoAnalysis
        self.p3 = primer3
        self.ta = ThermoAnalysis()      # the reference constructs it with defaults everywhere
        self.name = "primer3-py %s" % getattr(primer3, "__version__", "?")
        self.pins = "thal + designPrimers (real library)"

    @staticmethod
    def _row(r):
        return (r.tm, r.dg, r.dh, r.ds, int(bool(r.structure_found)))

    def thal(self, kind, a, b):
        out = []
        for x, y in zip(a, b if b is not None else a):
            if kind == "hairpin":
                r = self.ta.calcHairpin(x)
            elif kind == "end1":
                r = self.ta.calcEndStability(x, y)
            elif kind == "end2":        # upstream thal(): type END2 is END1 with the oligos swapped
                r = self.ta.calcEndStability(y, x)
            else:
                r = self.ta.calcHeterodimer(x, y)
            out.append(self._row(r))
        return np.array(out, dtype=np.float64)

    def design(self, template, fixed_left, excluded):
        seq_args = {"SEQUENCE_TEMPLATE": template, "SEQUENCE_PRIMER": fixed_left}
        if excluded:
            seq_args["SEQUENCE_EXCLUDED_REGION"] = excluded
        self.p3.setP3Globals(KASP_GLOBALS) if hasattr(self.p3, "setP3Globals") else None
        d = self.p3.bindings.designPrimers(seq_args, KASP_GLOBALS)
        n = int(d.get("PRIMER_PAIR_NUM_RETURNED", 0))
        return [{"right": d["PRIMER_RIGHT_%d_SEQUENCE" % i], "right_pos": list(d["PRIMER_RIGHT_%d" % i]),
                 "pair_penalty": d["PRIMER_PAIR_%d_PENALTY" % i],
                 "compl_any_th": d.get("PRIMER_PAIR_%d_COMPL_ANY_TH" % i),
                 "compl_end_th": d.get("PRIMER_PAIR_%d_COMPL_END_TH" % i),
                 "product_size": d["PRIMER_PAIR_%d_PRODUCT_SIZE" % i]} for i in range(n)]


class OracleSource:
    """The CPU oracle on a given table directory (or, for --stub, the embedded tables)."""

    def __init__(self, param_dir, threads):
        self.orc = Oracle(param_dir)
        self.threads = threads
        self.name = "oracle on %s" % (param_dir or "embedded literature tables (STUB)")
        self.pins = "table constants only" if param_dir else "nothing (stub)"

    def thal(self, kind, a, b):
        t = {"any": ANY, "tailed": ANY, "end1": END1, "end2": END2, "hairpin": HAIRPIN}[kind]
        r = self.orc.thal_batch(a, None if kind == "hairpin" else b, t, nthreads=self.threads)
        return np.stack([r["tm"], r["dg"], r["dh"], r["ds"], r["structure_found"].astype(np.float64)], axis=1)

    def design(self, template, fixed_left, excluded):
        pairs = p3_oracle.design(self.orc, settings_from(KASP_GLOBALS), template, excluded=excluded,
                                 left=fixed_left, nthreads=self.threads)
        return [{"right": e["right"]["seq"], "right_pos": [e["right"]["start"], e["right"]["len"]],
                 "pair_penalty": e["pair_penalty"], "compl_any_th": e["compl_any_th"],
                 "compl_end_th": e["compl_end_th"], "product_size": e["product_size"]} for e in pairs]


def settings_from(g):
    return dict(min_size=g["PRIMER_MIN_SIZE"], opt_size=g["PRIMER_OPT_SIZE"], max_size=g["PRIMER_MAX_SIZE"],
                num_return=g["PRIMER_NUM_RETURN"], min_tm=g["PRIMER_MIN_TM"], opt_tm=g["PRIMER_OPT_TM"],
                max_tm=g["PRIMER_MAX_TM"], min_gc=g["PRIMER_MIN_GC"], max_gc=g["PRIMER_MAX_GC"],
                max_poly_x=g["PRIMER_MAX_POLY_X"], pair_max_diff_tm=g["PRIMER_PAIR_MAX_DIFF_TM"],
                size_ranges=tuple(tuple(x) for x in g["PRIMER_PRODUCT_SIZE_RANGE"]))


def pick_source(args):
    try:
        import primer3  # noqa: F401
        if hasattr(primer3, "bindings") and not getattr(primer3, "__file__", None) is None:
            return Primer3Source(primer3)
    except ImportError:
        pass
    d = os.environ.get("POLYOLIGO_P3_CONFIG")
    if d:
        if not os.path.isdir(d):
            sys.exit("POLYOLIGO_P3_CONFIG=%s is not a directory" % d)
        return OracleSource(d, args.threads)
    if args.stub:
        return OracleSource(None, args.threads)
    return None


# ---------------------------------------------------------------------------------------------
# inputs
# ---------------------------------------------------------------------------------------------
def rand_oligos(rng, n, lo, hi):
    lens = rng.integers(lo, hi + 1, n)
    letters = np.array(list("ACGT"))
    return ["".join(letters[rng.integers(0, 4, l)]) for l in lens]


def thal_inputs(kind, n, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    if kind == "tailed":
        core = rand_oligos(rng, n, 18, 25)
        a = [(VIC if k % 2 == 0 else FAM) + s for k, s in enumerate(core)]
        return a, rand_oligos(rng, n, 18, 25)
    a = rand_oligos(rng, n, 18, 30)
    if kind == "hairpin":
        # a tenth of the set gets an engineered stem so that structures are common
        rc = str.maketrans("ACGT", "TGCA")
        for k in range(0, n, 10):
            stem = a[k][:rng.integers(4, 8)]
            a[k] = (a[k][:len(a[k]) - len(stem)] + stem[::-1].translate(rc))[:30]
        return a, None
    return a, rand_oligos(rng, n, 18, 30)


def design_inputs(n_markers, seed=20240608):
    """config-5 shaped markers: 500-nt templates, GC 0.42, SNP at index 249; the
    fixed primer is the 20-mer ending on the marker base (one allele-specific forward primer)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    letters = np.array(list("ACGT"))
    out = []
    for _ in range(n_markers):
        t = "".join(letters[rng.choice(4, 500, p=[0.29, 0.21, 0.21, 0.29])])
        ln = int(rng.integers(18, 26))
        out.append({"template": t, "fixed_left": t[250 - ln:250], "excluded": [[0, 1]]})
    return out


# ---------------------------------------------------------------------------------------------
# dump + check
# ---------------------------------------------------------------------------------------------
KINDS = ("any", "end1", "end2", "hairpin", "tailed")


def dump(src, args):
    os.makedirs(args.out, exist_ok=True)
    for i, kind in enumerate(KINDS):
        a, b = thal_inputs(kind, args.n, 1000 + i)
        res = src.thal(kind, a, b)
        np.savez_compressed(os.path.join(args.out, "thal_%s.npz" % kind), a=np.array(a),
                            b=np.array(b if b is not None else []), res=res)
        print("wrote thal_%s.npz (%d results, %d with a structure)" % (kind, len(a), int(res[:, 4].sum())))
    cases = design_inputs(args.markers)
    for c in cases:
        c["pairs"] = src.design(c["template"], c["fixed_left"], c["excluded"])
    with open(os.path.join(args.out, "design_kasp.json"), "w") as f:
        json.dump({"globals": KASP_GLOBALS, "cases": cases}, f)
    print("wrote design_kasp.json (%d markers, %d pairs)" % (len(cases), sum(len(c["pairs"]) for c in cases)))
    with open(os.path.join(args.out, "source.json"), "w") as f:
        json.dump({"source": src.name, "pins": src.pins, "n": args.n, "markers": args.markers}, f)


def compare_thal(name, got, want):
    """got / want: [n, 5] = tm, dg, dh, ds, structure_found.  Returns the number of bad rows."""
    sf = got[:, 4] != want[:, 4]
    tm = np.abs(got[:, 0] - want[:, 0]) > TOL_TM
    dg = np.abs(got[:, 1] - want[:, 1]) > TOL_DG
    bad = sf | tm | dg
    print("  %-34s %s  (%d rows: %d structure_found, %d Tm > %.2f C, %d dG > 1 cal/mol differ; max |dTm| %.4g)"
          % (name, "PASS" if not bad.any() else "FAIL", len(got), int(sf.sum()), int((tm & ~sf).sum()), TOL_TM,
             int((dg & ~sf).sum()), float(np.max(np.abs(got[:, 0] - want[:, 0])) if len(got) else 0.0)))
    return int(bad.sum())


def load_thal(out_dir, kind):
    z = np.load(os.path.join(out_dir, "thal_%s.npz" % kind))
    a = [str(x) for x in z["a"]]
    b = [str(x) for x in z["b"]] if len(z["b"]) else None
    return a, b, z["res"]


def gpu_context(param_dir):
    try:
        import polyoligo_b200 as po
        return po, po.Context(0, param_dir)
    except Exception as e:   # no GPU here: the product has no CPU fallback, so only the oracle is checked
        print("  (CUDA library not checked: %s)" % str(e).splitlines()[0][:100])
        return None, None


def check(args, param_dir):
    bad = 0
    orc = OracleSource(param_dir, args.threads)
    po, ctx = gpu_context(param_dir)
    print("parity against the vectors in %s:" % args.out)
    for kind in KINDS:
        a, b, want = load_thal(args.out, kind)
        bad += compare_thal("oracle  thal %s" % kind, orc.thal(kind, a, b), want)
        if ctx is not None:
            t = {"any": po.THAL_ANY, "tailed": po.THAL_ANY, "end1": po.THAL_END1, "end2": po.THAL_END2,
                 "hairpin": po.THAL_HAIRPIN}[kind]
            r = ctx.thal_batch(t, po.pack(a), po.pack(b) if b is not None else None)
            got = np.stack([r["tm"], r["dg"], r["dh"], r["ds"], (r["flags"] & 1).astype(np.float64)], axis=1)
            bad += compare_thal("CUDA    thal %s" % kind, got, want)
    with open(os.path.join(args.out, "design_kasp.json")) as f:
        cases = json.load(f)["cases"]
    miss = 0
    for c in cases:
        got = orc.design(c["template"], c["fixed_left"], [tuple(x) for x in c["excluded"]])
        if [p["right"] for p in got] != [p["right"] for p in c["pairs"]]:
            miss += 1
    print("  %-34s %s  (%d of %d markers pick different primers)"
          % ("oracle  designPrimers (KASP)", "PASS" if not miss else "FAIL", miss, len(cases)))
    bad += miss
    if ctx is not None:
        from polyoligo_b200 import primer3_bindings as p3b
        p3b.resetP3Globals()
        p3b.setP3Globals(KASP_GLOBALS)
        miss = 0
        res = p3b.design_primers_batch([{"SEQUENCE_TEMPLATE": c["template"], "SEQUENCE_PRIMER": c["fixed_left"],
                                         "SEQUENCE_EXCLUDED_REGION": c["excluded"]} for c in cases])
        for c, d in zip(cases, res):
            n = int(d.get("PRIMER_PAIR_NUM_RETURNED", 0))
            if [d["PRIMER_RIGHT_%d_SEQUENCE" % i] for i in range(n)] != [p["right"] for p in c["pairs"]]:
                miss += 1
        p3b.resetP3Globals()
        print("  %-34s %s  (%d of %d markers pick different primers)"
              % ("CUDA    designPrimers (KASP)", "PASS" if not miss else "FAIL", miss, len(cases)))
        bad += miss
        ctx.close()
    return bad


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--n", type=int, default=10000, help="thal results per type")
    ap.add_argument("--markers", type=int, default=200, help="config-5 markers for designPrimers")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "primer3"))
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--stub", action="store_true", help="let the oracle stand in for primer3 (pins nothing)")
    ap.add_argument("--no-pytest", action="store_true", help="skip step 4")
    args = ap.parse_args()
    src = pick_source(args)
    if src is None:
        print("repin: no source of truth here -- `import primer3` fails and POLYOLIGO_P3_CONFIG is unset.\n"
              "       thal and the picker stay PARITY UNPINNED (tests/test_oracle_cpu.py holds the strict xfail).")
        return 2
    print("source: %s   pins: %s" % (src.name, src.pins))
    dump(src, args)
    param_dir = os.environ.get("POLYOLIGO_P3_CONFIG")
    bad = check(args, param_dir)
    rc = 0
    if not args.no_pytest:
        env = dict(os.environ, POLYOLIGO_GOLDEN_P3=args.out)
        try:
            import torch
            have_gpu = torch.cuda.is_available()
        except Exception:
            have_gpu = False
        marks = [] if have_gpu else ["-m", "not gpu"]
        rc = subprocess.call([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests"), "-x", "-q"] + marks,
                             env=env, cwd=ROOT)
    print("repin: %s" % ("all sets agree" if bad == 0 and rc == 0 else
                         "%d differing rows/markers, pytest rc %d" % (bad, rc)))
    return 0 if bad == 0 and rc == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
