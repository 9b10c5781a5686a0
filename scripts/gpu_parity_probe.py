"""Quick GPU-vs-oracle parity probe (development aid; the judged tests live in tests/)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import polyoligo_b200 as po
from oracle.oracle import Oracle, ANY, HAIRPIN

def rand_seqs(rng, n, lo, hi, alphabet="ACGT"):
    lens = rng.integers(lo, hi + 1, n)
    return ["".join(rng.choice(list(alphabet), l)) for l in lens]

def compare(name, got, ref):
    sf = (got["flags"] & 1)
    bad = np.flatnonzero((sf != ref["structure_found"]) | (got["tm"] != ref["tm"]) | (got["dg"] != ref["dg"])
                         | (got["dh"] != ref["dh"]) | (got["ds"] != ref["ds"]))
    print("%-28s n=%d  bit-exact mismatches=%d  max|dTm|=%.3g" % (
        name, len(ref), len(bad), np.max(np.abs(got["tm"] - ref["tm"])) if len(ref) else 0))
    return bad

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    rng = np.random.default_rng(7)
    ctx = po.Context(0)
    orc = Oracle()
    s1 = rand_seqs(rng, n, 18, 30); s2 = rand_seqs(rng, n, 18, 30)
    # adversarial: low complexity, palindromes, long, tailed
    adv1 = rand_seqs(rng, 2000, 5, 60, "AT") + rand_seqs(rng, 2000, 2, 60, "ACGT") + rand_seqs(rng, 1000, 40, 60, "GC")
    adv2 = rand_seqs(rng, 2000, 5, 60, "AT") + rand_seqs(rng, 2000, 2, 60, "ACGT") + rand_seqs(rng, 1000, 40, 60, "GC")
    adv1 += ["A" * 60, "ACGT" * 15, "GCGCGCGCGCGCGCGCGCGC", "A", "AC"]
    adv2 += ["T" * 60, "ACGT" * 15, "GCGCGCGCGCGCGCGCGCGC", "T", "GT"]
    for label, a, b in (("random 18-30", s1, s2), ("adversarial", adv1, adv2)):
        pa, pb = po.pack(a), po.pack(b)
        t = time.time(); got = ctx.thal_batch(po.THAL_ANY, pa, pb); dt = time.time() - t
        ref = orc.thal_batch(a, b, ANY, nthreads=8)
        bad = compare("heterodimer " + label, got, ref)
        for k in bad[:5]:
            print("   ", a[k], b[k], got[k], ref[k])
        print("    gpu %.1f k pairs/s (incl. copies)" % (len(a) / dt / 1e3))
        got = ctx.thal_batch(po.THAL_ANY, pa); ref = orc.thal_batch(a, a, ANY, nthreads=8)
        bad = compare("homodimer " + label, got, ref)
        for k in bad[:5]:
            print("   ", a[k], got[k], ref[k])
        t = time.time(); got = ctx.thal_batch(po.THAL_HAIRPIN, pa); dt = time.time() - t
        ref = orc.thal_batch(a, None, HAIRPIN, nthreads=8)
        bad = compare("hairpin " + label, got, ref)
        for k in bad[:5]:
            print("   ", a[k], got[k], ref[k])
        print("    gpu %.1f k hairpins/s (incl. copies)" % (len(a) / dt / 1e3))
        tm, gc = ctx.tm_batch(pa, with_gc=True)
        rtm = orc.tm_batch(a)
        print("%-28s mismatches=%d" % ("tm " + label, int(np.sum(tm != rtm))))
        rel = ctx.count_relaxations(po.THAL_ANY, pa, pb)
        print("    relaxations gpu=%d oracle=%d" % (rel, int(orc.thal_batch(a, b, ANY, nthreads=8)["relaxations"].sum())))
        relh = ctx.count_relaxations(po.THAL_HAIRPIN, pa)
        print("    hairpin relaxations gpu=%d oracle=%d" % (relh, int(ref["relaxations"].sum())))
    print("fp64 peak TFLOP/s:", ctx.measure_fp64_peak())
    # throughput, device-resident-ish: big batch
    big = 400000
    a = rand_seqs(rng, big, 18, 30); b = rand_seqs(rng, big, 18, 30)
    pa, pb = po.pack(a), po.pack(b)
    for rep in range(2):
        t = time.time(); ctx.thal_batch(po.THAL_ANY, pa, pb); dt = time.time() - t
        print("het big: %.2f M pairs/s" % (big / dt / 1e6))
        t = time.time(); ctx.thal_batch(po.THAL_HAIRPIN, pa); dt = time.time() - t
        print("hairpin big: %.2f M/s" % (big / dt / 1e6))

main()
