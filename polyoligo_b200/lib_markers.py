"""Host-side mirror of the region objects the PCR / CAPS / HRM flows are written against
(reference src/polyoligo/lib_markers.py): ``ROI`` / ``Region`` (:13-77, :392-449), ``Marker``
(:297-389) and ``PCRproduct`` (:452-529, ``get_tm`` :586-587).

What stays outside (external tools): BLAST / MUSCLE homolog mapping
(``ROI.map_homologs``) -- a flank ROI arrives here with its ``p3_sequence_included_maps`` already
filled in -- and the VCF reader behind ``upload_mutations`` (any object with ``fetch_genotypes``).
Sequences come from ``blast_hook.fetch(query)`` exactly as in the reference (any object with that
method; ``lib_markers.SequenceStore`` below serves them from an in-memory genome), product melting
temperatures from the device (``lib_primer3.hrm_temps_batch`` -> ``po_tm_nn_batch``).
"""
import re

import numpy as np

from . import lib_primer3


class SequenceStore:
    """Minimal ``blast_hook``: serves 1-based closed ranges of in-memory chromosomes with the
    key format of the reference's BlastDB.fetch ("chr:start-stop")."""

    def __init__(self, chromosomes, temporary=None):
        self.chromosomes = dict(chromosomes)
        self.temporary = temporary

    def fetch(self, query):
        out = {}
        for q in query:
            seq = self.chromosomes[q["chr"]][max(int(q["start"]), 1) - 1:int(q["stop"])]
            out["{}:{}-{}".format(q["chr"], q["start"], q["stop"])] = seq
        return out


class ROI:
    def __init__(self, chrom=None, start=None, stop=None, blast_hook=None, malign_hook=None, vcf_hook=None,
                 name=None, marker=None, do_print_alt_subjects=False, is_indel=None):
        self.chrom, self.start, self.stop = chrom, int(start), int(stop)
        self.blast_hook, self.malign_hook, self.vcf_hook = blast_hook, malign_hook, vcf_hook
        self.marker, self.is_indel = marker, is_indel
        self.do_print_alt_subjects = do_print_alt_subjects
        self.name = name if name is not None else "{}:{}-{}".format(self.chrom, self.start, self.stop)
        self.seq = self.seq_alt = self.seq_x = self.n = None
        self.G = self.mutations = self.has_marker = None
        self.p3_sequence_target = None
        self.p3_sequence_included_maps = None
        self.p3_sequence_excluded_regions = None
        if blast_hook is not None:                                   # reference :53-62
            fetched = blast_hook.fetch([{"chr": self.chrom, "start": self.start, "stop": self.stop}])
            self.seq = next(iter(fetched.values())).upper()
            self.n = len(self.seq)
        if marker is not None and self.seq is not None:              # reference :64-77
            if self.start <= marker.pos1 <= self.stop:
                a = marker.pos1 - self.start
                b = a + marker.n
                self.seq_alt = self.seq[:a] + marker.alt + self.seq[b:]
                self.seq_x = self.seq[:a] + "x" * marker.n + self.seq[b:]
            else:
                self.seq_alt = self.seq_x = self.seq

    def upload_mutations(self):
        """reference :279-294: the mutations of the region, the marker itself removed."""
        self.mutations = []
        if self.vcf_hook:
            self.G, self.mutations = self.vcf_hook.fetch_genotypes(chrom=self.chrom, start=self.start, stop=self.stop)
            if self.marker is not None:
                i = np.where(np.array(self.G.columns) == self.marker.pos1)[0]
                if len(i) == 1:
                    self.G.pop(self.marker.pos1)
                    self.mutations.pop(i[0])


class Marker:
    """One input marker (reference :297-389): 0-based ``pos``, 1-based ``pos1``, '*' = empty allele."""

    def __init__(self, name=None, chrom=None, pos=None, ref_allele=None, alt_allele=None):
        self.chrom, self.pos, self.pos1 = chrom, int(pos), int(pos) + 1
        self.ref = "" if ref_allele == "*" else ref_allele
        self.alt = "" if alt_allele == "*" else alt_allele
        self.variant = "[{}/{}]".format(self.ref, self.alt)
        self.n = len(self.ref)
        self.is_indel = None
        if name is None or name == ".":
            name = "{}:{}-{}".format(self.chrom, self.pos1, self.pos1)
        self.name = name.replace("_", "-")

    def _fetch(self, blast_hook, start, stop):
        seq = "N"
        for seq in blast_hook.fetch([{"chr": self.chrom, "start": int(start), "stop": int(stop)}]).values():
            seq = seq.upper()
            break
        return seq

    def parse_indels(self, blast_hook):
        if self.ref == "":            # insertion: REF becomes the bases the insert replaces
            self.n = len(self.alt)
            self.ref = self._fetch(blast_hook, self.pos1, self.pos1 + self.n - 1)
            self.is_indel = "ins"
        if self.alt == "":            # deletion: ALT becomes the bases that follow
            self.n = len(self.ref)
            self.alt = self._fetch(blast_hook, self.pos1 + self.n, self.pos1 + 2 * self.n - 1)
            self.is_indel = "del"
        self.variant = "[{}/{}]".format(self.ref, self.alt)

    def assert_marker(self, blast_hook):
        if self.ref != "":
            seq = self._fetch(blast_hook, self.pos1, self.pos1 + self.n - 1)
            if self.ref != seq:
                return ("REF allele in the marker file does not match the genomic REF allele.\n"
                        "        SNP ID: {} | Marker {} vs Reference {}\n"
                        "        Please double check your marker file.".format(self.name, self.ref, seq))
        return None


class Region(ROI):
    def __init__(self, left_roi=None, right_roi=None, **kwargs):
        super().__init__(**kwargs)
        self.left_roi, self.right_roi = left_roi, right_roi

    def merge_with_primers(self):
        """The template the two-sided picker sees (reference :403-449): the span from the left
        flank's start to the right flank's end, ``SEQUENCE_TARGET`` = this region, and for every
        inclusion map of the flanks a merged map in which a base is usable iff a flank covers it,
        every covering flank includes it and it lies outside this region."""
        left, right = self.left_roi, self.right_roi
        region = Region(chrom=self.chrom, start=left.start, stop=right.stop, blast_hook=self.blast_hook,
                        malign_hook=self.malign_hook, vcf_hook=self.vcf_hook, name=self.name, marker=self.marker,
                        do_print_alt_subjects=self.do_print_alt_subjects, left_roi=left, right_roi=right)
        region.p3_sequence_target = [[self.start - left.start, self.n]]
        n = region.n

        def window(a, b):          # genome range [a, b] as a slice of the merged template
            return max(a - region.start, 0), max(min(b - region.start + 1, n), 0)

        la, lb = window(left.start, left.stop)
        ra, rb = window(right.start, right.stop)
        ca, cb = window(self.start, self.stop)
        region.p3_sequence_included_maps = {}
        for k, lmap in left.p3_sequence_included_maps.items():
            lmap = np.asarray(lmap, dtype=bool)
            rmap = np.asarray(right.p3_sequence_included_maps[k], dtype=bool)
            if lb - la != len(lmap) or rb - ra != len(rmap):
                raise ValueError("merge_with_primers: a flank map does not fit its ROI inside the merged region")
            covered = np.zeros(n, dtype=bool)
            blocked = np.zeros(n, dtype=bool)
            covered[la:lb] = True
            blocked[la:lb] |= ~lmap
            covered[ra:rb] = True
            blocked[ra:rb] |= ~rmap
            blocked[ca:cb] = True
            region.p3_sequence_included_maps[k] = covered & ~blocked
        if left.mutations is not None:
            region.mutations = left.mutations + right.mutations
        return region


def get_tm(seq, device=None):
    """Biopython ``MeltingTemp.Tm_NN(seq, strict=False)`` on the device (reference :586-587)."""
    return float(lib_primer3.get_context(device).tm_nn_batch([seq])[0])


class PCRproduct:
    """The amplicon of one primer pair (reference :452-529)."""

    def __init__(self, roi, pp):
        self.roi = Region(chrom=roi.chrom, start=pp.primers["F"].start, stop=pp.primers["R"].stop,
                          blast_hook=roi.blast_hook, malign_hook=roi.malign_hook, vcf_hook=roi.vcf_hook,
                          name=roi.marker.name, marker=roi.marker, do_print_alt_subjects=roi.do_print_alt_subjects)
        self.pp = pp
        self.enzymes = None
        self.hrm = None

    def will_it_cut(self, enzymes, min_fragment_size):
        """Enzymes that cut exactly one allele exactly once, both fragments >= min_fragment_size
        (reference :471-516).  Kept quirks: the accepted enzyme DICT is shared between products (the
        caller's list), so fragl / fragr / allele hold the values of the last product that accepted
        it; the cut position is the site start plus half the site length, rounded DOWN (:492,:498,
        :512) except for a forward-strand hit on the ALT allele, rounded UP (:506); a reverse-strand
        hit overrides a forward-strand one."""
        self.enzymes = []
        for enzyme in enzymes:
            fwd, rev = re.compile(enzyme["regex"]["F"]), re.compile(enzyme["regex"]["R"])
            n_ref = max(len(fwd.findall(self.roi.seq)), len(rev.findall(self.roi.seq)))
            n_alt = max(len(fwd.findall(self.roi.seq_alt)), len(rev.findall(self.roi.seq_alt)))
            fragl = fragr = 0
            allele = ""
            if n_ref == 1 and n_alt == 0:
                seq, allele, up_on_fwd = self.roi.seq, "REF", False
            elif n_ref == 0 and n_alt == 1:
                seq, allele, up_on_fwd = self.roi.seq_alt, "ALT", True
            else:
                seq = None
            if seq is not None:
                for rx, up in ((fwd, up_on_fwd), (rev, False)):
                    for m in rx.finditer(seq):
                        size = m.end() - m.start()
                        cut = m.start() + ((size + 1) // 2 if up else size // 2)
                        fragl, fragr = cut, self.roi.n - cut
            if fragl >= min_fragment_size and fragr >= min_fragment_size:
                enzyme["fragl"], enzyme["fragr"], enzyme["allele"] = fragl, fragr, allele
                self.enzymes.append(enzyme)

    def get_hrm_temps(self, device=None):
        self.hrm = lib_primer3.hrm_temps_batch([self.roi.seq], [self.roi.seq_alt], device)[0]
