// po_pipeline.cuh -- kernels around the thermodynamic core: KASP allele-specific
// candidate enumeration, reporter-tail concatenation and the
// heterodimer predicate / goodness score / per-marker top-n (rows a8, a9).
#pragma once
#include "po_common.h"
#include "po_thal.cuh"

// Python slice normalisation: seq[start:stop] on a sequence of length n
__device__ __forceinline__ void py_slice(int n, int start, int stop, int* lo, int* hi) {
    if (start < 0) start += n;
    if (stop < 0) stop += n;
    start = min(max(start, 0), n);
    stop = min(max(stop, 0), n);
    *lo = start;
    *hi = max(stop, start);
}

struct OligoBuilder {
    uint32_t w[4];
    int len;
    uint32_t flags;
    __device__ __forceinline__ void init() {
        w[0] = w[1] = w[2] = w[3] = 0;
        len = 0;
        flags = 0;
    }
    __device__ __forceinline__ void push(int code, bool is_n) {
        if (is_n) flags |= PO_OLIGO_HAS_N;
        if (len < PO_MAX_OLIGO_LEN) {
            const uint32_t v = (uint32_t)(code & 3) << (2 * (len & 15));
            if (len < 16) w[0] |= v;
            else if (len < 32) w[1] |= v;
            else if (len < 48) w[2] |= v;
            else w[3] |= v;
        } else {
            flags |= PO_OLIGO_BAD_CHAR;  // longer than a record: never evaluated
        }
        ++len;
    }
    __device__ __forceinline__ uint4 finish() const {
        const uint32_t meta = (uint32_t)min(len, PO_MAX_OLIGO_LEN) | flags;
        return make_uint4(w[0], w[1], w[2], (w[3] & 0x00ffffffu) | (meta << 24));
    }
};

// get_marker_primers (reference src/polyoligo/_lib_kasp.py:295-359): for every
// padding in [min_size, max_size] and direction F/R, the REF primer whose 3' end
// is the marker allele and its ALT twin.  One thread per (marker, padding, dir);
// slot = (padding - min_size) * 2 + dir, the reference's visiting order.
__global__ void k_kasp_enumerate(const po_kasp_marker* __restrict__ M, long long n, int min_size, int max_size,
                                 uint4* __restrict__ ref_out, uint4* __restrict__ alt_out) {
    const int slots = 2 * (max_size - min_size + 1);
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n * slots) return;
    const long long m = t / slots;
    const int s = (int)(t % slots);
    const int padding = min_size + s / 2, dir = s & 1;
    const po_kasp_marker& mk = M[m];
    const int L = mk.tmpl_len, ref_len = mk.ref_len, alt_len = mk.alt_len;
    const int d = ref_len - alt_len, ad = abs(d);
    int lo, hi;
    if (dir == 0) {
        const int x2 = mk.marker_pos + ref_len, x1 = x2 - padding - ad;  // cseq = seq[x1:x2]
        py_slice(L, x1, x2, &lo, &hi);
    } else {
        const int y2 = mk.marker_pos_rc, y1 = y2 - padding - ad - ref_len;  // cseq = seq_rc[y1:y2]
        py_slice(L, y1, y2, &lo, &hi);
    }
    const int clen = hi - lo;
    const int pos_a = clen - ref_len;  // position of the allele in the primer
    auto tmpl_at = [&](int k, bool* is_n) -> int {  // k-th base of cseq
        const int p = dir == 0 ? lo + k : L - 1 - (lo + k);
        const int code = (mk.tmpl[p >> 4] >> (2 * (p & 15))) & 3;
        *is_n = (mk.nmask[p >> 5] >> (p & 31)) & 1u;
        return dir == 0 ? code : 3 - code;
    };
    auto alt_at = [&](int k, bool* is_n) -> int {  // k-th base of calt_a (reverse complement for R)
        const int p = dir == 0 ? k : alt_len - 1 - k;
        const int code = (mk.alt[p >> 4] >> (2 * (p & 15))) & 3;
        *is_n = (mk.alt_nmask >> p) & 1u;
        return dir == 0 ? code : 3 - code;
    };
    OligoBuilder rb, ab;
    rb.init();
    ab.init();
    if (pos_a < 0 || clen <= 0) {  // degenerate slice (marker too close to the template edge): never valid
        rb.flags = ab.flags = PO_OLIGO_BAD_CHAR;
    } else {
        bool isn;
        for (int k = max(d, 0); k < clen; ++k) {  // pseq = cseq[d:] when the ALT allele is shorter
            const int c = tmpl_at(k, &isn);
            rb.push(c, isn);
        }
        const int alt_total = pos_a + alt_len;
        for (int k = max(-d, 0); k < alt_total; ++k) {  // pseq_alt = (cseq[:pos_a] + alt)[-d:]
            const int c = k < pos_a ? tmpl_at(k, &isn) : alt_at(k - pos_a, &isn);
            ab.push(c, isn);
        }
    }
    ref_out[t] = rb.finish();
    alt_out[t] = ab.finish();
}

// reporter + primer (reference src/polyoligo/_lib_kasp.py:75-97): out = tail[sel] + oligo,
// sel < 0 = no tail.
__global__ void k_concat_tail(const uint4* __restrict__ in, const int8_t* __restrict__ sel, const uint4 tail0,
                              const uint4 tail1, long long n, uint4* __restrict__ out) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint4 o = in[t];
    const int s = sel[t];
    if (s < 0) {
        out[t] = o;
        return;
    }
    const uint4 tl = s == 0 ? tail0 : tail1;
    OligoBuilder b;
    b.init();
    b.flags = ((o.w >> 24) | (tl.w >> 24)) & (PO_OLIGO_HAS_N | PO_OLIGO_BAD_CHAR);
    const int lt = po_len(tl), lo_ = po_len(o);
    for (int k = 0; k < lt; ++k) b.push(po_base(tl, k), false);
    for (int k = 0; k < lo_; ++k) b.push(po_base(o, k), false);
    out[t] = b.finish();
}

// Heterodimer predicates and goodness score of one primer pair / KASP triple.
//   PCR / CAPS / HRM: PrimerPair.check_heterodimerization (reference lib_primer3.py:147-164)
//   KASP:             its override, reference _lib_kasp.py:89-117 (tm_delta is applied on
//                     BOTH sides of the comparison there; restated literally)
//   score:            reference scoring.py:4-151
struct RankArgs {
    int assay;  // PO_ASSAY_*
    long long n;
    const double* tmF;
    const double* tmR;
    const double* tmA;
    const po_thal_out* het1;  // (F, R)
    const po_thal_out* het2;  // KASP: (A, R) when dir == F, (F, A) when dir == R
    const int32_t* n_off;
    const double* aaf;  // n x 3 (F, R, A)
    const int32_t* indel;
    const double* hrm_delta;
    double tm_delta;
    int32_t* goodness;
    int32_t* qbits;
};

__global__ void k_score_pairs(const RankArgs a) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= a.n) return;
    const bool kasp = a.assay == PO_ASSAY_KASP, hrm = a.assay == PO_ASSAY_HRM;
    const double tF = a.tmF[t], tR = a.tmR[t], tA = kasp ? a.tmA[t] : 0.0;
    const double dl = a.tm_delta;
    bool dimer;
    int bits = 0;
    if (kasp) {
        const double tm_ref = a.het1[t].tm - dl, tm_alt = a.het2[t].tm - dl;
        const bool ref_dimer = !((tm_ref <= (tF - dl)) && (tm_ref <= (tR - dl)));
        const bool alt_dimer = !((tm_alt <= (tF - dl)) && (tm_alt <= (tR - dl)));
        dimer = ref_dimer || alt_dimer;
        bits |= (ref_dimer ? PO_Q_REF_DIMER : 0) | (alt_dimer ? PO_Q_ALT_DIMER : 0);
    } else {
        const double tm_het = a.het1[t].tm;
        dimer = !((tm_het <= (tF - dl)) && (tm_het <= (tR - dl)));
    }
    int score = 0;
    if (hrm) {
        const double hd = a.hrm_delta[t];
        if (hd >= 0.5) {
            score += 2;
            if (hd >= 1) score += 1;
            else bits |= PO_Q_h;
        } else {
            bits |= PO_Q_H;
        }
    }
    // np.sum(np.abs(tms - np.mean(tms))) in numpy's evaluation order
    double l1;
    if (kasp) {
        const double mean = ((tF + tR) + tA) / 3.0;
        l1 = (fabs(tF - mean) + fabs(tR - mean)) + fabs(tA - mean);
    } else {
        const double mean = (tF + tR) / 2.0;
        l1 = fabs(tF - mean) + fabs(tR - mean);
    }
    if (l1 <= 5) score += 1;
    else bits |= PO_Q_t;
    if (a.n_off[t] == 0) score += hrm ? 1 : 3;
    else bits |= PO_Q_O;
    if (!dimer) score += 1;
    else bits |= PO_Q_d;
    const double a0 = a.aaf[3 * t], a1 = a.aaf[3 * t + 1], a2 = kasp ? a.aaf[3 * t + 2] : 0.0;
    if (a0 < 0.1 && a1 < 0.1 && a2 < 0.1) {
        score += hrm ? 1 : 2;
        if (a0 == 0 && a1 == 0 && a2 == 0) score += 1;
        else bits |= PO_Q_m;
    } else {
        bits |= PO_Q_M;
    }
    const int ind = a.indel[t];
    if (ind < 50) {
        score += 1;
        if (ind == 0) score += 1;
        else bits |= PO_Q_i;
    } else {
        bits |= PO_Q_I;
    }
    a.goodness[t] = score;
    a.qbits[t] = bits;
}

// PCR.classify + PCR.prune (reference lib_primer3.py:425-463, _lib_kasp.py:278-291):
// per marker, pairs in descending goodness, insertion order inside a class (HRM:
// descending delta first), first top_n.  One warp per marker; rank by counting.
__global__ void k_top_n(const int32_t* __restrict__ goodness, const double* __restrict__ hrm_delta, int use_delta,
                        const long long* __restrict__ seg, long long n_markers, int top_n,
                        long long* __restrict__ sel, int32_t* __restrict__ n_sel) {
    const long long m = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (m >= n_markers) return;
    const long long s = seg[m], e = seg[m + 1];
    for (long long i = s + lane; i < e; i += 32) {
        const int gi = goodness[i];
        const double di = use_delta ? hrm_delta[i] : 0.0;
        int rank = 0;
        for (long long j = s; j < e; ++j) {
            const int gj = goodness[j];
            const double dj = use_delta ? hrm_delta[j] : 0.0;
            const bool before = gj > gi || (gj == gi && (dj > di || (dj == di && j < i)));
            rank += before ? 1 : 0;
        }
        if (rank < top_n) sel[m * top_n + rank] = i;
    }
    const long long cnt = e - s;
    if (lane == 0) n_sel[m] = (int32_t)(cnt < top_n ? cnt : top_n);
    for (long long r = cnt + lane; r < top_n; r += 32) sel[m * top_n + r] = -1;
}

// second KASP heterodimer: (A', R') when the marker primer is F, (F', A') when it is R
// (reference _lib_kasp.py:102-105).  selF[k] == 0 marks dir F (F carries dye1).
__global__ void k_select_pair(const uint4* __restrict__ Ft, const uint4* __restrict__ Rt, const uint4* __restrict__ At,
                              const int8_t* __restrict__ selF, long long n, uint4* __restrict__ x1,
                              uint4* __restrict__ x2) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n) return;
    const bool dirF = selF[t] == 0;
    x1[t] = dirF ? At[t] : Ft[t];
    x2[t] = dirF ? Rt[t] : At[t];
}

// interleave per-pair Tm values for the host: tm[n][3] = F, R, A; het[n][2]
__global__ void k_gather_tm(const double* __restrict__ tF, const double* __restrict__ tR, const double* __restrict__ tA,
                            const po_thal_out* __restrict__ h1, const po_thal_out* __restrict__ h2, long long n,
                            double* __restrict__ tm3, double* __restrict__ het2) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n) return;
    tm3[3 * t] = tF[t];
    tm3[3 * t + 1] = tR[t];
    tm3[3 * t + 2] = tA ? tA[t] : 0.0;
    het2[2 * t] = h1[t].tm;
    het2[2 * t + 1] = h2 ? h2[t].tm : 0.0;
}

// Biopython 1.76 MeltingTemp.Tm_NN with default arguments and a perfect complement
// (reference lib_markers.py:586-587).  Sums run in Biopython's order, in kcal/mol and
// cal/(mol K) doubles, so the result is the same double Python computes.
struct TmNNConsts {
    double log_mon;   // math.log((Na + K + Tris / 2.0) * 1e-3)
    double r_log_k;   // 1.987 * math.log((dnac1 - dnac2 / 2.0) * 1e-9)
};
__constant__ double c_nn3_h[16] = {-7.9, -8.4, -7.8, -7.2, -8.5, -8.0, -10.6, -7.8, -8.2, -9.8, -8.0, -8.4, -7.2, -8.2, -8.5, -7.9};
__constant__ double c_nn3_s[16] = {-22.2, -22.4, -21.0, -20.4, -22.7, -19.9, -27.2, -21.0, -22.2, -24.4, -19.9, -22.4, -21.3, -22.2, -22.7, -22.2};

// MeltingTemp._check(seq, "Tm_NN") keeps A, C, G, T (U is back-transcribed to T) and I, and DROPS
// every other character (N and the other ambiguity codes included) before anything is summed --
// the reference calls Tm_NN(seq, strict=False), so nothing raises.  4 = inosine: stays in the
// sequence, but its neighbour pairs are in no table (strict=False: skipped) and it is not an
// A/T/G/C end.  -1 = dropped.
__device__ __forceinline__ int ascii_code(char ch) {
    switch (ch) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': case 'U': case 'u': return 3;
        case 'I': case 'i': return 4;
    }
    return -1;
}

__global__ void k_tm_nn(const char* __restrict__ ascii, const long long* __restrict__ off, long long n,
                        const TmNNConsts K, double* __restrict__ tm) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n) return;
    const char* s = ascii + off[t];
    const int len = (int)(off[t + 1] - off[t]);
    int kept = 0, first = -1, last = -1;
    for (int k = 0; k < len; ++k) {
        const int c = ascii_code(s[k]);
        if (c < 0) continue;
        if (first < 0) first = c;
        last = c;
        ++kept;
    }
    if (kept == 0) {   // Biopython: IndexError on seq[0]
        tm[t] = nan("");
        return;
    }
    // Biopython adds the initiation terms BEFORE zipping; addition order matters for the
    // last bit: init terms first, then the dinucleotides.
    const int at = (first == 0 || first == 3) + (last == 0 || last == 3);
    const int gc = (first == 1 || first == 2) + (last == 1 || last == 2);
    double h = 0.0, sS = 0.0;
    h += 2.3 * at;
    sS += 4.1 * at;
    h += 0.1 * gc;
    sS += -2.8 * gc;
    int prev = -1;
    for (int k = 0; k < len; ++k) {
        const int c = ascii_code(s[k]);
        if (c < 0) continue;
        if (prev >= 0 && prev < 4 && c < 4) {  // the 'zipping': nearest-neighbour sum along the kept bases
            h += c_nn3_h[4 * prev + c];
            sS += c_nn3_s[4 * prev + c];
        }
        prev = c;
    }
    const double corr = 0.368 * (kept - 1) * K.log_mon;
    sS += corr;
    tm[t] = (1000 * h) / (sS + K.r_log_k) - 273.15;
}

// PrimerPair.add_mutations, numeric part (reference lib_primer3.py:166-210)
__global__ void k_annotate_mutations(const long long* __restrict__ mut_pos, const double* __restrict__ mut_aaf,
                                     const int32_t* __restrict__ mut_indel, const long long* __restrict__ mut_seg,
                                     const long long* __restrict__ pair_seg, long long n_markers,
                                     const long long* __restrict__ pstart, const long long* __restrict__ pstop,
                                     int n_primers, long long n_pairs, double* __restrict__ max_aaf,
                                     int32_t* __restrict__ max_indel) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n_pairs) return;
    long long lo = 0, hi = n_markers - 1;  // marker of this pair: last m with pair_seg[m] <= t
    while (lo < hi) {
        const long long mid = (lo + hi + 1) >> 1;
        if (pair_seg[mid] <= t) lo = mid;
        else hi = mid - 1;
    }
    const long long ms = mut_seg[lo], me = mut_seg[lo + 1];
    double aaf[3] = {0.0, 0.0, 0.0};
    int indel = 0;
    const long long w0 = pstart[3 * t], w1 = pstop[3 * t + 1];  // product window: F.start .. R.stop
    for (long long m = ms; m < me; ++m) {
        const long long pos = mut_pos[m];
#pragma unroll
        for (int d = 0; d < 3; ++d)
            if (d < n_primers && pos >= pstart[3 * t + d] && pos <= pstop[3 * t + d]) aaf[d] = fmax(aaf[d], mut_aaf[m]);
        if (pos >= w0 && pos <= w1) indel = max(indel, mut_indel[m]);
    }
    max_aaf[3 * t] = aaf[0];
    max_aaf[3 * t + 1] = aaf[1];
    max_aaf[3 * t + 2] = n_primers > 2 ? aaf[2] : 0.0;
    max_indel[t] = indel;
}
