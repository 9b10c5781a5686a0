// po_p3.cuh -- primer3-core style picking of primers on the device.
//
// Replaces primer3.bindings.designPrimers as reached from the reference at
// src/polyoligo/lib_primer3.py:466-529 (get_primers), called from
// src/polyoligo/_lib_kasp.py:567-590 with one primer fixed.  libprimer3 is not
// part of the reference tree; the semantics restated here:
//   * candidates: every window of PRIMER_MIN_SIZE..PRIMER_MAX_SIZE nt that does
//     not touch an excluded / target base or a non-ACGT letter, with
//     GC %, GC clamp, 3'-end GC, poly-X and Tm (oligotm, SantaLucia) inside
//     their limits; penalty = wt_tm * |Tm - opt| + wt_size * |len - opt|
//   * lists sorted by (penalty, start descending, length ascending)
//   * the thermodynamic alignments (self any / self end / hairpin, pair any /
//     end) are evaluated LAZILY, only for oligos that reach a candidate pair
//   * pair constraints: product size inside the current PRIMER_PRODUCT_SIZE_RANGE
//     entry (entries are tried in order, a later entry never re-visits sizes an
//     earlier one covered), |Tm difference|, target flanked, thal limits
//   * pairs ordered by (pair penalty, compl_measure, left start descending,
//     right start ascending, left length, right length)
#pragma once

// ---------------------------------------------------------------------------
// candidate enumeration
// ---------------------------------------------------------------------------
__device__ __forceinline__ int p3_code(char c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 4;
    }
}

__device__ __forceinline__ long long p3_find_set(const long long* __restrict__ off, long long n_sets, long long g) {
    long long lo = 0, hi = n_sets;  // last s with off[s] <= g
    while (hi - lo > 1) {
        const long long mid = (lo + hi) >> 1;
        if (off[mid] <= g) lo = mid;
        else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ double p3_penalty(const po_p3_settings& S, double tm, int len) {
    double sum = 0.0;
    if (S.wt_tm_gt != 0.0 && tm > S.opt_tm) sum += S.wt_tm_gt * (tm - S.opt_tm);
    if (S.wt_tm_lt != 0.0 && tm < S.opt_tm) sum += S.wt_tm_lt * (S.opt_tm - tm);
    if (S.wt_size_lt != 0.0 && len < S.opt_size) sum += S.wt_size_lt * (S.opt_size - len);
    if (S.wt_size_gt != 0.0 && len > S.opt_size) sum += S.wt_size_gt * (len - S.opt_size);
    return sum;
}

// GC clamp, 3'-end GC count and poly-X of a packed primer (5'->3')
__device__ __forceinline__ bool p3_sequence_rules(const po_p3_settings& S, const uint4& o, int len) {
    if (S.gc_clamp > 0) {
        for (int k = 0; k < S.gc_clamp && k < len; ++k) {
            const int b = po_base(o, len - 1 - k);
            if (b != 1 && b != 2) return false;
        }
    }
    if (S.max_end_gc < 5) {
        int n = 0;
        for (int k = 0; k < 5 && k < len; ++k) {
            const int b = po_base(o, len - 1 - k);
            n += (b == 1 || b == 2);
        }
        if (n > S.max_end_gc) return false;
    }
    if (S.max_poly_x > 0 && S.max_poly_x < len) {
        int run = 1, prev = po_base(o, 0);
        for (int k = 1; k < len; ++k) {
            const int b = po_base(o, k);
            run = (b == prev) ? run + 1 : 1;
            prev = b;
            if (run > S.max_poly_x) return false;
        }
    }
    return true;
}

// One thread per (template base, strand): the windows whose 5' end is that base.
// FILL == false counts the accepted candidates per list, FILL == true writes them
// (unordered inside a list; k_p3_sort orders them).
template <bool FILL>
__global__ void k_p3_enum(const char* __restrict__ tmpl, const uint8_t* __restrict__ mask,
                          const long long* __restrict__ set_off, long long n_sets, const po_p3_settings S,
                          const PoConsts K, unsigned long long* __restrict__ counts,
                          const long long* __restrict__ list_off, po_p3_oligo* __restrict__ pool,
                          unsigned long long* __restrict__ vmask) {
    const long long total = set_off[n_sets];
    const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (g >= total) return;
    const int strand = blockIdx.y;
    // the count pass records which lengths were accepted (bit L - min_size); the fill pass only
    // re-evaluates those
    unsigned long long accepted = 0ull;
    const unsigned long long todo = FILL ? vmask[strand * total + g] : ~0ull;
    if (FILL && todo == 0ull) return;
    const long long set = p3_find_set(set_off, n_sets, g);
    const long long off = set_off[set];
    const int n = (int)(set_off[set + 1] - off), p = (int)(g - off);
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    int gc = 0;
    for (int L = 1; L <= S.max_size; ++L) {
        const int q = strand == 0 ? p + L - 1 : p - (L - 1);
        if (q < 0 || q >= n) break;
        const int code = p3_code(tmpl[off + q]);
        if (code > 3 || (mask && mask[off + q])) break;  // every longer window contains this base too
        const int b = strand ? 3 - code : code;
        w[(L - 1) >> 4] |= (uint32_t)b << (2 * ((L - 1) & 15));
        gc += (b == 1 || b == 2);
        if (L < S.min_size) continue;
        if (FILL && !((todo >> (L - S.min_size)) & 1ull)) continue;
        uint4 o = make_uint4(w[0], w[1], w[2], (w[3] & 0x00ffffffu) | ((uint32_t)L << 24));
        double gcp;
        const double tm = oligo_tm(o, K, &gcp);
        if (gcp < S.min_gc || gcp > S.max_gc) continue;
        if (tm < S.min_tm || tm > S.max_tm) continue;
        if (!p3_sequence_rules(S, o, L)) continue;
        accepted |= 1ull << (L - S.min_size);
        const unsigned long long slot = atomicAdd(&counts[2 * set + strand], 1ull);
        if (FILL) {
            po_p3_oligo r;
            memcpy(&r.seq, &o, 16);
            r.tm = tm;
            r.gc_percent = gcp;
            r.penalty = p3_penalty(S, tm, L);
            r.self_any_th = r.self_end_th = r.hairpin_th = CUDART_NAN;
            r.start = p;
            r.len = L;
            r.set = (int)set;
            r.flags = PO_P3_OK;
            pool[list_off[2 * set + strand] + (long long)slot] = r;
        }
    }
    if (!FILL) vmask[strand * total + g] = accepted;
}

// ---------------------------------------------------------------------------
// per-list sort: (penalty, start descending, length ascending), libprimer3's
// primer_rec_comp.  One CTA per list, bitonic network on two 64-bit key words;
// the keys live in shared memory when the padded list fits, else in `gscratch`.
// ---------------------------------------------------------------------------
#define P3_SORT_SMEM_ELEMS 4096
__global__ void __launch_bounds__(512)
k_p3_sort(const po_p3_oligo* __restrict__ pool, po_p3_oligo* __restrict__ sorted,
          const long long* __restrict__ list_off, const long long* __restrict__ scratch_off,
          unsigned long long* __restrict__ gscratch) {
    extern __shared__ unsigned long long p3_sort_smem[];
    const long long list = blockIdx.x;
    const long long off = list_off[list];
    const int n = (int)(list_off[list + 1] - off);
    if (n == 0) return;
    int P = 1;
    while (P < n) P <<= 1;
    unsigned long long *k1, *k2;
    if (P <= P3_SORT_SMEM_ELEMS) {
        k1 = p3_sort_smem;
        k2 = p3_sort_smem + P3_SORT_SMEM_ELEMS;
    } else {
        k1 = gscratch + scratch_off[list];
        k2 = k1 + P;
    }
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        if (i < n) {
            const po_p3_oligo& r = pool[off + i];
            k1[i] = ordered_bits(r.penalty);
            k2[i] = ((unsigned long long)(0xFFFFF - r.start) << 44) | ((unsigned long long)r.len << 36) |
                    (unsigned long long)i;
        } else {
            k1[i] = ~0ull;
            k2[i] = ~0ull;
        }
    }
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < P; i += blockDim.x) {
                const int x = i ^ j;
                if (x > i) {
                    const bool asc = (i & k) == 0;
                    const unsigned long long a1 = k1[i], a2 = k2[i], b1 = k1[x], b2 = k2[x];
                    const bool gt = (a1 > b1) | ((a1 == b1) & (a2 > b2));
                    if (gt == asc) {
                        k1[i] = b1;
                        k2[i] = b2;
                        k1[x] = a1;
                        k2[x] = a2;
                    }
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        sorted[off + i] = pool[off + (long long)(k2[i] & 0xFFFFFFFFFull)];
}

// ---------------------------------------------------------------------------
// fixed-primer tasks
// ---------------------------------------------------------------------------
struct P3TaskState {
    int interval;    // index of the product size range being searched
    int cursor;      // next entry of the other side's sorted list
    int n_found;
    int done;
    int exhausted;   // the list ended in this interval
    int tie_mode;    // num_return reached: only pairs tied with the last accepted one are still wanted
    int n_prop;      // proposals of the current round
    int stop_after;  // tie mode met a strictly worse pair: finish after this round
    double last_q;   // pair penalty of the last accepted pair
};

struct P3Result {
    int cand;        // index into the other side's list (relative to the list start)
    int interval;
    int product;
    int pad;
    double pq, compl_any, compl_end, compl_measure;
};

#define P3_K 64  // proposals per task and round

// the given primer's own record: cut from the template, cheap constraints applied
__global__ void k_p3_fixed_init(const char* __restrict__ tmpl, const uint8_t* __restrict__ mask,
                                const long long* __restrict__ set_off, const po_p3_task* __restrict__ tasks,
                                long long n_tasks, const po_p3_settings S, const PoConsts K,
                                po_p3_oligo* __restrict__ fixed, P3TaskState* __restrict__ state) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n_tasks) return;
    const po_p3_task tk = tasks[t];
    const long long off = set_off[tk.set];
    const int n = (int)(set_off[tk.set + 1] - off);
    po_p3_oligo r;
    memset(&r, 0, sizeof(r));
    // libprimer3 locates a given primer by a case-insensitive string search (the right primer's
    // reverse complement), i.e. at the FIRST occurrence of its sequence in the template
    // (reference lib_primer3.py:479 hands designPrimers the sequence, not a position): a task that
    // names a later copy of a repeated sequence is moved to the first one.
    int start = tk.start;
    {
        const int lo = tk.side == 1 ? tk.start : tk.start - tk.len + 1;
        if ((tk.side == 1 || tk.side == 2) && tk.len > 0 && lo > 0 && lo + tk.len <= n) {
            for (int p = 0; p < lo; ++p) {
                bool same = true;
                for (int k = 0; k < tk.len; ++k)
                    if ((tmpl[off + p + k] & 0xDF) != (tmpl[off + lo + k] & 0xDF)) {
                        same = false;
                        break;
                    }
                if (same) {
                    start = tk.side == 1 ? p : p + tk.len - 1;
                    break;
                }
            }
        }
    }
    r.start = start;
    r.len = tk.len;
    r.set = tk.set;
    r.self_any_th = r.self_end_th = r.hairpin_th = CUDART_NAN;
    bool ok = (tk.side == 1 || tk.side == 2) && tk.len >= S.min_size && tk.len <= S.max_size && tk.len <= PO_MAX_OLIGO_LEN;
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    for (int k = 0; k < tk.len && ok; ++k) {
        const int q = tk.side == 1 ? start + k : start - k;
        if (q < 0 || q >= n) {
            ok = false;
            break;
        }
        const int code = p3_code(tmpl[off + q]);
        // a given primer may lie inside an excluded region only if the caller says so; libprimer3
        // applies the same position constraints to it as to a picked one
        if (code > 3 || (mask && mask[off + q])) {
            ok = false;
            break;
        }
        const int b = tk.side == 2 ? 3 - code : code;
        w[k >> 4] |= (uint32_t)b << (2 * (k & 15));
    }
    if (ok) {
        uint4 o = make_uint4(w[0], w[1], w[2], (w[3] & 0x00ffffffu) | ((uint32_t)tk.len << 24));
        memcpy(&r.seq, &o, 16);
        double gcp;
        r.tm = oligo_tm(o, K, &gcp);
        r.gc_percent = gcp;
        r.penalty = p3_penalty(S, r.tm, tk.len);
        if (gcp < S.min_gc || gcp > S.max_gc) ok = false;
        if (r.tm < S.min_tm || r.tm > S.max_tm) ok = false;
        if (ok && !p3_sequence_rules(S, o, tk.len)) ok = false;
    } else {
        // an empty record still has to be a valid thal input
        const uint4 o = make_uint4(0u, 0u, 0u, 1u << 24);
        memcpy(&r.seq, &o, 16);
    }
    r.flags = ok ? PO_P3_OK : 0;
    fixed[t] = r;
    P3TaskState st;
    memset(&st, 0, sizeof(st));
    st.done = ok ? 0 : 1;
    state[t] = st;
}

// self alignments of the given primers (thal results of ANY(s,s), END1(s,s), HAIRPIN(s))
__global__ void k_p3_fixed_self(long long n_tasks, const po_thal_out* __restrict__ any,
                                const po_thal_out* __restrict__ end1, const po_thal_out* __restrict__ hp,
                                const po_p3_settings S, po_p3_oligo* __restrict__ fixed,
                                P3TaskState* __restrict__ state) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n_tasks) return;
    po_p3_oligo& r = fixed[t];
    if (!(r.flags & PO_P3_OK)) return;
    r.self_any_th = any[t].tm;
    r.self_end_th = end1[t].tm;
    r.hairpin_th = hp[t].tm;
    r.flags |= PO_P3_SELF_DONE;
    if (r.self_any_th > S.max_self_any_th || r.self_end_th > S.max_self_end_th || r.hairpin_th > S.max_hairpin_th) {
        r.flags = (r.flags & ~PO_P3_OK) | PO_P3_SELF_FAILED;
        state[t].done = 1;
    }
}

// product size of (given, candidate) and whether an EARLIER size range covers it
__device__ __forceinline__ bool p3_size_ok(const po_p3_settings& S, int interval, int product) {
    if (product < S.size_lo[interval] || product > S.size_hi[interval]) return false;
    for (int e = 0; e < interval; ++e)
        if (product >= S.size_lo[e] && product <= S.size_hi[e]) return false;
    return true;
}

// One warp per task: walk the other side's sorted list from the cursor and propose the next
// candidates that satisfy the cheap pair constraints.  Newly met candidates are queued for
// their self alignments; every proposal is queued for its pair alignments.
__global__ void k_p3_propose(const po_p3_task* __restrict__ tasks, long long n_tasks, const po_p3_settings S,
                             const int32_t* __restrict__ set_target, const long long* __restrict__ list_off,
                             po_p3_oligo* __restrict__ pool, const po_p3_oligo* __restrict__ fixed,
                             P3TaskState* __restrict__ state, int32_t* __restrict__ prop,
                             unsigned long long* __restrict__ ctr, long long* __restrict__ self_list,
                             long long* __restrict__ pair_list, int margin_div) {
    const long long t = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= n_tasks) return;
    P3TaskState st = state[t];
    if (st.done) return;
    const po_p3_task tk = tasks[t];
    const po_p3_oligo f = fixed[t];
    const long long list = 2ll * tk.set + (tk.side == 1 ? 1 : 0);
    const long long off = list_off[list];
    const int n = (int)(list_off[list + 1] - off);
    int tar_s = 0, tar_l = 0;
    if (set_target) {
        tar_s = set_target[2 * tk.set];
        tar_l = set_target[2 * tk.set + 1];
    }
    const int need = S.num_return - st.n_found;
    const int want = st.tie_mode ? P3_K : min(P3_K, need + need / margin_div + 2);  // few alignments fail the limits
    int r = 0, cur = st.cursor, stop_after = 0;
    while (cur < n && r < want && !stop_after) {
        const int c = cur + lane;
        const bool valid = c < n;
        const po_p3_oligo* rec = &pool[off + (valid ? c : 0)];
        const int cs = rec->start, cl = rec->len, cf = rec->flags;
        const double ctm = rec->tm, cq = rec->penalty;
        const int product = tk.side == 1 ? cs - f.start + 1 : f.start - cs + 1;
        const double pq = tk.side == 1 ? f.penalty + cq : cq + f.penalty;  // left + right
        bool ok = valid && p3_size_ok(S, st.interval, product) && !(cf & PO_P3_SELF_FAILED);
        ok = ok && !(fabs(f.tm - ctm) > S.pair_max_diff_tm);
        if (tar_l > 0) {
            const int last_of_left = tk.side == 1 ? f.start + f.len - 1 : cs + cl - 1;
            const int first_of_right = tk.side == 1 ? cs - cl + 1 : f.start - f.len + 1;
            ok = ok && last_of_left < tar_s && first_of_right >= tar_s + tar_l;
        }
        unsigned stopm = 0u;
        if (st.tie_mode) {
            ok = ok && (pq == st.last_q);
            stopm = __ballot_sync(FULL, valid && pq > st.last_q);
            if (stopm) {
                const int first_stop = __ffs(stopm) - 1;
                ok = ok && lane < first_stop;
                stop_after = 1;
            }
        }
        const unsigned okm = __ballot_sync(FULL, ok);
        const int rank = __popc(okm & ((1u << lane) - 1u));
        const int room = want - r;
        if (ok && rank < room) {
            prop[t * P3_K + r + rank] = c;
            const unsigned long long slot = atomicAdd(&ctr[1], 1ull);
            pair_list[slot] = t * P3_K + r + rank;
            if (!(cf & (PO_P3_SELF_DONE | PO_P3_QUEUED))) {
                const int old = atomicOr(&pool[off + c].flags, PO_P3_QUEUED);
                if (!(old & PO_P3_QUEUED)) self_list[atomicAdd(&ctr[0], 1ull)] = off + c;
            }
        }
        const int nok = __popc(okm);
        if (nok >= room) {
            // cursor: just past the candidate that filled the last slot
            unsigned m = okm;
            for (int s = 1; s < room; ++s) m &= m - 1;
            cur += __ffs(m);
            r = want;
        } else {
            r += nok;
            cur += 32;
        }
    }
    if (lane == 0) {
        st.cursor = min(cur, n);
        st.n_prop = r;
        st.exhausted = (cur >= n) ? 1 : 0;
        st.stop_after = stop_after;
        state[t] = st;
    }
}

__global__ void k_p3_gather_self(const long long* __restrict__ self_list, long long n,
                                 const po_p3_oligo* __restrict__ pool, uint4* __restrict__ seq) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    memcpy(&seq[i], &pool[self_list[i]].seq, 16);
}
__global__ void k_p3_gather_fixed(const po_p3_oligo* __restrict__ fixed, long long n, uint4* __restrict__ seq) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    memcpy(&seq[i], &fixed[i].seq, 16);
}
__global__ void k_p3_scatter_self(const long long* __restrict__ self_list, long long n,
                                  const po_thal_out* __restrict__ any, const po_thal_out* __restrict__ end1,
                                  const po_thal_out* __restrict__ hp, const po_p3_settings S,
                                  po_p3_oligo* __restrict__ pool) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    po_p3_oligo& r = pool[self_list[i]];
    r.self_any_th = any[i].tm;
    r.self_end_th = end1[i].tm;
    r.hairpin_th = hp[i].tm;
    int fl = (r.flags | PO_P3_SELF_DONE) & ~PO_P3_QUEUED;
    if (r.self_any_th > S.max_self_any_th || r.self_end_th > S.max_self_end_th || r.hairpin_th > S.max_hairpin_th)
        fl |= PO_P3_SELF_FAILED;
    r.flags = fl;
}
// (left, right) sequences of every queued proposal
__global__ void k_p3_gather_pairs(const long long* __restrict__ pair_list, long long n,
                                  const po_p3_task* __restrict__ tasks, const long long* __restrict__ list_off,
                                  const po_p3_oligo* __restrict__ pool, const po_p3_oligo* __restrict__ fixed,
                                  const int32_t* __restrict__ prop, uint4* __restrict__ left,
                                  uint4* __restrict__ right) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long slot = pair_list[i], t = slot / P3_K;
    const po_p3_task tk = tasks[t];
    const long long list = 2ll * tk.set + (tk.side == 1 ? 1 : 0);
    const po_p3_oligo& c = pool[list_off[list] + prop[slot]];
    if (tk.side == 1) {
        memcpy(&left[i], &fixed[t].seq, 16);
        memcpy(&right[i], &c.seq, 16);
    } else {
        memcpy(&left[i], &c.seq, 16);
        memcpy(&right[i], &fixed[t].seq, 16);
    }
}
__global__ void k_p3_scatter_pairs(const long long* __restrict__ pair_list, long long n,
                                   const po_thal_out* __restrict__ any, const po_thal_out* __restrict__ end1,
                                   const po_thal_out* __restrict__ end2, double* __restrict__ pairvals) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long slot = pair_list[i];
    pairvals[2 * slot] = any[i].tm;
    pairvals[2 * slot + 1] = fmax(end1[i].tm, end2[i].tm);
}

// One thread per task: accept the proposals whose alignments are inside the limits, in list
// order, and advance the task's search state.
__global__ void k_p3_accept(const po_p3_task* __restrict__ tasks, long long n_tasks, const po_p3_settings S,
                            const long long* __restrict__ list_off, const po_p3_oligo* __restrict__ pool,
                            const po_p3_oligo* __restrict__ fixed, P3TaskState* __restrict__ state,
                            const int32_t* __restrict__ prop, const double* __restrict__ pairvals,
                            P3Result* __restrict__ results, int rcap, unsigned long long* __restrict__ ctr) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n_tasks) return;
    P3TaskState st = state[t];
    if (st.done) return;
    const po_p3_task tk = tasks[t];
    const po_p3_oligo& f = fixed[t];
    const long long list = 2ll * tk.set + (tk.side == 1 ? 1 : 0);
    const long long off = list_off[list];
    for (int r = 0; r < st.n_prop; ++r) {
        const long long slot = t * P3_K + r;
        const int c = prop[slot];
        const po_p3_oligo& o = pool[off + c];
        if (o.flags & PO_P3_SELF_FAILED) continue;
        const double any = pairvals[2 * slot], end = pairvals[2 * slot + 1];
        if (any > S.pair_max_compl_any_th || end > S.pair_max_compl_end_th) continue;
        if (st.n_found >= rcap) break;
        P3Result res;
        res.cand = c;
        res.interval = st.interval;
        res.product = tk.side == 1 ? o.start - f.start + 1 : f.start - o.start + 1;
        res.pad = 0;
        res.pq = tk.side == 1 ? f.penalty + o.penalty : o.penalty + f.penalty;
        res.compl_any = any;
        res.compl_end = end;
        const po_p3_oligo& L = tk.side == 1 ? f : o;
        const po_p3_oligo& R = tk.side == 1 ? o : f;
        res.compl_measure = (R.self_end_th + L.self_end_th + end) * 1.1 + R.self_any_th + L.self_any_th + any;
        results[t * rcap + st.n_found] = res;
        st.n_found++;
        st.last_q = res.pq;
    }
    st.n_prop = 0;
    if (st.n_found >= rcap) st.done = 1;
    else if (st.tie_mode) {
        if (st.stop_after || st.exhausted) st.done = 1;
    } else if (st.n_found >= S.num_return) {
        if (st.exhausted) st.done = 1;
        else st.tie_mode = 1;
    } else if (st.exhausted) {
        st.interval++;
        if (st.interval >= S.n_size_ranges) st.done = 1;
        else {
            st.cursor = 0;
            st.exhausted = 0;
        }
    }
    state[t] = st;
    if (!st.done) atomicAdd(&ctr[2], 1ull);
}

// One thread per task: order the accepted pairs (full libprimer3 pair order); the chosen
// candidates are queued once for the optional Primer.is_valid pass.
__global__ void k_p3_order(const po_p3_task* __restrict__ tasks, long long n_tasks, const po_p3_settings S,
                           const long long* __restrict__ list_off, po_p3_oligo* __restrict__ pool,
                           const P3TaskState* __restrict__ state, P3Result* __restrict__ results, int rcap,
                           int want_valid, unsigned long long* __restrict__ ctr, long long* __restrict__ used_list) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n_tasks) return;
    const po_p3_task tk = tasks[t];
    const long long list = 2ll * tk.set + (tk.side == 1 ? 1 : 0);
    const long long off = list_off[list];
    const int n = state[t].n_found;
    P3Result* R = results + t * rcap;
    // insertion sort; the list is already ordered by (interval, pair penalty) except inside ties
    auto before = [&](const P3Result& a, const P3Result& b) {
        if (a.interval != b.interval) return a.interval < b.interval;
        if (a.pq != b.pq) return a.pq < b.pq;
        if (a.compl_measure != b.compl_measure) return a.compl_measure < b.compl_measure;
        const po_p3_oligo &oa = pool[off + a.cand], &ob = pool[off + b.cand];
        if (oa.start != ob.start) {
            // compare_primer_pair: left start descending, right start ascending
            return tk.side == 2 ? oa.start > ob.start : oa.start < ob.start;
        }
        return oa.len < ob.len;
    };
    for (int i = 1; i < n; ++i) {
        const P3Result x = R[i];
        int j = i - 1;
        while (j >= 0 && before(x, R[j])) {
            R[j + 1] = R[j];
            --j;
        }
        R[j + 1] = x;
    }
    if (want_valid) {
        const int m = min(n, S.num_return);
        for (int i = 0; i < m; ++i) {
            const long long c = off + R[i].cand;
            const int old = atomicOr(&pool[c].flags, PO_P3_QUEUED);
            if (!(old & PO_P3_QUEUED)) used_list[atomicAdd(&ctr[3], 1ull)] = c;
        }
    }
}

// Primer.is_valid of the queued candidates (props computed under the caller's conditions)
__global__ void k_p3_scatter_valid(const long long* __restrict__ used_list, long long n,
                                   const po_oligo_props* __restrict__ props, po_p3_oligo* __restrict__ pool) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    po_p3_oligo& r = pool[used_list[i]];
    int fl = r.flags & ~PO_P3_QUEUED;
    if (props[i].flags & PO_PRIMER_VALID) fl |= PO_P3_PRIMER_VALID;
    r.flags = fl;
}

__global__ void k_p3_emit(const po_p3_task* __restrict__ tasks, long long n_tasks, const po_p3_settings S,
                          const long long* __restrict__ list_off, const po_p3_oligo* __restrict__ pool,
                          const P3TaskState* __restrict__ state, const P3Result* __restrict__ results, int rcap,
                          po_p3_pair* __restrict__ out, int32_t* __restrict__ n_out) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n_tasks) return;
    const po_p3_task tk = tasks[t];
    const long long off = list_off[2ll * tk.set + (tk.side == 1 ? 1 : 0)];
    const P3Result* R = results + t * rcap;
    const int m = min(state[t].n_found, S.num_return);
    for (int i = 0; i < m; ++i) {
        po_p3_pair p;
        p.other = pool[off + R[i].cand];
        p.pair_penalty = R[i].pq;
        p.compl_any_th = R[i].compl_any;
        p.compl_end_th = R[i].compl_end;
        p.product_size = R[i].product;
        p.size_range = R[i].interval;
        out[t * S.num_return + i] = p;
    }
    n_out[t] = m;
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static int p3_check_settings(po_ctx* ctx, const po_p3_settings* s) {
    if (!s || s->min_size < 1 || s->max_size < s->min_size || s->max_size > PO_MAX_OLIGO_LEN || s->num_return < 1 ||
        s->n_size_ranges < 0 || s->n_size_ranges > 8) {
        ctx->err = "po_p3: invalid settings";
        return PO_E_INVALID;
    }
    return PO_OK;
}

struct P3Lists {
    std::vector<long long> list_off;     // 2*n_sets+1
    std::vector<long long> scratch_off;  // per list, for lists too long for the shared-memory sort
    long long total = 0, scratch_total = 0;
};

// enumerate + sort; leaves the sorted pool in ctx->p3[1] and the device copies of set_off /
// list_off in ctx->p3[2] / ctx->p3[3]
// `dev`: tmpl / mask / set_off are DEVICE pointers (copied device to device) and `dev_total_bases`
// is set_off[n_sets] (the fused KASP design keeps everything resident).
static int p3_build_lists(po_ctx* ctx, const po_p3_settings* s, const PoConsts& K, const char* tmpl,
                          const uint8_t* mask, const int64_t* set_off, int64_t n_sets, P3Lists* L, bool dev = false,
                          long long dev_total_bases = 0) {
    cudaStream_t st = ctx->stream;
    DevBuf* sb = ctx->p3;
    int rc;
    const long long total_bases = dev ? dev_total_bases : set_off[n_sets];
    const cudaMemcpyKind kind = dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    const size_t nl = (size_t)(2 * n_sets);
    if ((rc = ensure(ctx, sb[2], 8 * (size_t)(n_sets + 1)))) return rc;
    if ((rc = ensure(ctx, sb[3], 8 * (nl + 1)))) return rc;
    if ((rc = ensure(ctx, sb[4], (size_t)(total_bases > 0 ? total_bases : 1)))) return rc;
    if ((rc = ensure(ctx, sb[5], (size_t)(total_bases > 0 ? total_bases : 1)))) return rc;
    if ((rc = ensure(ctx, sb[6], 8 * (nl + 1)))) return rc;
    if ((rc = ensure(ctx, sb[7], 8 * (nl + 1)))) return rc;
    if ((rc = ensure(ctx, sb[26], 16 * (size_t)(total_bases > 0 ? total_bases : 1)))) return rc;  // accepted-length masks
    CK(cudaMemcpyAsync(sb[2].p, set_off, 8 * (size_t)(n_sets + 1), kind, st));
    if (total_bases > 0) {
        CK(cudaMemcpyAsync(sb[4].p, tmpl, (size_t)total_bases, kind, st));
        if (mask) CK(cudaMemcpyAsync(sb[5].p, mask, (size_t)total_bases, kind, st));
    }
    const uint8_t* d_mask = mask ? (const uint8_t*)sb[5].p : nullptr;
    unsigned long long* d_counts = (unsigned long long*)sb[6].p;
    L->list_off.assign(nl + 1, 0);
    L->scratch_off.assign(nl + 1, 0);
    if (total_bases == 0) return PO_OK;
    const int threads = 128;
    const dim3 grid((unsigned)((total_bases + threads - 1) / threads), 2);
    CK(cudaMemsetAsync(d_counts, 0, 8 * nl, st));
    k_p3_enum<false><<<grid, threads, 0, st>>>((const char*)sb[4].p, d_mask, (const long long*)sb[2].p, n_sets, *s, K,
                                               d_counts, nullptr, nullptr, (unsigned long long*)sb[26].p);
    ctx->launches++;
    CK(cudaGetLastError());
    std::vector<unsigned long long> counts(nl);
    CK(cudaMemcpyAsync(counts.data(), d_counts, 8 * nl, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (size_t l = 0; l < nl; ++l) {
        L->list_off[l + 1] = L->list_off[l] + (long long)counts[l];
        long long P = 1;
        while (P < (long long)counts[l]) P <<= 1;
        L->scratch_off[l + 1] = L->scratch_off[l] + (P > P3_SORT_SMEM_ELEMS ? 2 * P : 0);
    }
    L->total = L->list_off[nl];
    L->scratch_total = L->scratch_off[nl];
    if (L->total == 0) return PO_OK;
    if ((rc = ensure(ctx, sb[0], sizeof(po_p3_oligo) * (size_t)L->total))) return rc;
    if ((rc = ensure(ctx, sb[1], sizeof(po_p3_oligo) * (size_t)L->total))) return rc;
    if ((rc = ensure(ctx, sb[8], 8 * (size_t)(L->scratch_total > 0 ? L->scratch_total : 1)))) return rc;
    CK(cudaMemcpyAsync(sb[3].p, L->list_off.data(), 8 * (nl + 1), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(sb[7].p, L->scratch_off.data(), 8 * (nl + 1), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(d_counts, 0, 8 * nl, st));
    k_p3_enum<true><<<grid, threads, 0, st>>>((const char*)sb[4].p, d_mask, (const long long*)sb[2].p, n_sets, *s, K,
                                              d_counts, (const long long*)sb[3].p, (po_p3_oligo*)sb[0].p,
                                              (unsigned long long*)sb[26].p);
    ctx->launches++;
    CK(cudaGetLastError());
    const size_t smem = 2 * P3_SORT_SMEM_ELEMS * sizeof(unsigned long long);
    CK(cudaFuncSetAttribute(k_p3_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_p3_sort<<<(unsigned)nl, 512, smem, st>>>((const po_p3_oligo*)sb[0].p, (po_p3_oligo*)sb[1].p,
                                               (const long long*)sb[3].p, (const long long*)sb[7].p,
                                               (unsigned long long*)sb[8].p);
    ctx->launches++;
    CK(cudaGetLastError());
    return PO_OK;
}

static void p3_consts(const po_p3_settings* s, const PoConsts& base, PoConsts* K) {
    // libprimer3: thal at 37 C with maxLoop 30; oligotm with MAX_NN_TM_LENGTH 36
    resolve_conditions(K, s->mv_conc, s->dv_conc, s->dntp_conc, s->dna_conc, 37.0, 30, 36);
    (void)base;
}

extern "C" int po_p3_candidates_batch(po_ctx* ctx, const po_p3_settings* s, const char* tmpl, const uint8_t* mask,
                                      const int64_t* set_off, int64_t n_sets, int64_t* list_off, po_p3_oligo* out,
                                      int64_t out_cap) {
    if (!ctx || !set_off || !list_off || n_sets < 0 || (!tmpl && n_sets > 0 && set_off[n_sets] > 0)) return PO_E_INVALID;
    int rc = p3_check_settings(ctx, s);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    PoConsts K;
    p3_consts(s, ctx->K, &K);
    P3Lists L;
    if ((rc = p3_build_lists(ctx, s, K, tmpl, mask, set_off, n_sets, &L))) return rc;
    for (size_t l = 0; l < L.list_off.size(); ++l) list_off[l] = L.list_off[l];
    if (out && L.total > 0) {
        if (out_cap < L.total) {
            ctx->err = "po_p3_candidates_batch: output capacity too small";
            return PO_E_INVALID;
        }
        CK(cudaMemcpyAsync(out, ctx->p3[1].p, sizeof(po_p3_oligo) * (size_t)L.total, cudaMemcpyDeviceToHost,
                           ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return PO_OK;
}

// The search itself.  `dev`: tmpl / mask / set_off / tasks are device pointers (dev_total_bases =
// set_off[n_sets]) and the results stay on the device -- pairs in ctx->p3[24] (n_tasks x num_return
// po_p3_pair), pair counts in ctx->p3[25], the given primers' records in ctx->p3[10] -- unless the
// output pointers are given.
static int p3_design_fixed_impl(po_ctx* ctx, const po_p3_settings* s, const char* tmpl, const uint8_t* mask,
                                const int64_t* set_off, const int32_t* set_target, int64_t n_sets,
                                const po_p3_task* tasks, int64_t n_tasks, const po_valid_params* valid,
                                po_p3_oligo* fixed_out, po_p3_pair* pairs_out, int32_t* n_pairs_out, bool dev,
                                long long dev_total_bases) {
    int rc;
    cudaStream_t st = ctx->stream;
    DevBuf* sb = ctx->p3;
    PoConsts K;
    p3_consts(s, ctx->K, &K);
    P3Lists L;
    if ((rc = p3_build_lists(ctx, s, K, tmpl, mask, set_off, n_sets, &L, dev, dev_total_bases))) return rc;
    const size_t nt = (size_t)n_tasks;
    const int rcap = s->num_return + 2 * P3_K;
    const size_t nslot = nt * P3_K;
    if ((rc = ensure(ctx, sb[9], sizeof(po_p3_task) * nt))) return rc;
    if ((rc = ensure(ctx, sb[10], sizeof(po_p3_oligo) * nt))) return rc;
    if ((rc = ensure(ctx, sb[11], sizeof(P3TaskState) * nt))) return rc;
    if ((rc = ensure(ctx, sb[12], 4 * nslot))) return rc;                       // prop
    if ((rc = ensure(ctx, sb[13], 8 * 2 * nslot))) return rc;                   // pairvals
    if ((rc = ensure(ctx, sb[14], 8 * nslot))) return rc;                       // pair_list
    if ((rc = ensure(ctx, sb[15], 8 * (size_t)(L.total > 0 ? L.total : 1)))) return rc;  // self_list
    if ((rc = ensure(ctx, sb[16], sizeof(P3Result) * nt * rcap))) return rc;
    if ((rc = ensure(ctx, sb[17], 16 * nslot))) return rc;                      // left seqs / self seqs
    if ((rc = ensure(ctx, sb[18], 16 * nslot))) return rc;                      // right seqs
    if ((rc = ensure(ctx, sb[19], sizeof(po_thal_out) * nslot))) return rc;
    if ((rc = ensure(ctx, sb[20], sizeof(po_thal_out) * nslot))) return rc;
    if ((rc = ensure(ctx, sb[21], sizeof(po_thal_out) * nslot))) return rc;
    if ((rc = ensure(ctx, sb[22], 8 * 4))) return rc;                           // counters
    if ((rc = ensure(ctx, sb[23], 8 * (size_t)(n_sets > 0 ? n_sets : 1)))) return rc;  // targets
    if ((rc = ensure(ctx, sb[24], sizeof(po_p3_pair) * nt * s->num_return))) return rc;
    if ((rc = ensure(ctx, sb[25], 4 * nt))) return rc;
    // self alignments of candidates can outnumber the proposal slots only if a list is longer
    // than all slots; size the thal buffers for both uses
    const size_t nself_max = (size_t)(L.total > 0 ? L.total : 1);
    if (nself_max > nslot) {
        if ((rc = ensure(ctx, sb[17], 16 * nself_max))) return rc;
        if ((rc = ensure(ctx, sb[19], sizeof(po_thal_out) * nself_max))) return rc;
        if ((rc = ensure(ctx, sb[20], sizeof(po_thal_out) * nself_max))) return rc;
        if ((rc = ensure(ctx, sb[21], sizeof(po_thal_out) * nself_max))) return rc;
    }
    po_p3_task* d_tasks = (po_p3_task*)sb[9].p;
    po_p3_oligo* d_fixed = (po_p3_oligo*)sb[10].p;
    P3TaskState* d_state = (P3TaskState*)sb[11].p;
    po_p3_oligo* d_pool = (po_p3_oligo*)sb[1].p;
    const long long* d_list_off = (const long long*)sb[3].p;
    unsigned long long* d_ctr = (unsigned long long*)sb[22].p;
    const int32_t* d_target = nullptr;
    CK(cudaMemcpyAsync(d_tasks, tasks, sizeof(po_p3_task) * nt, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    if (L.total == 0) CK(cudaMemcpyAsync(sb[3].p, L.list_off.data(), 8 * L.list_off.size(), cudaMemcpyHostToDevice, st));
    if (set_target && n_sets > 0) {
        CK(cudaMemcpyAsync(sb[23].p, set_target, 8 * (size_t)n_sets, cudaMemcpyHostToDevice, st));
        d_target = (const int32_t*)sb[23].p;
    }
    const int threads = 128;
    const unsigned tb = (unsigned)((n_tasks + threads - 1) / threads);
    const uint8_t* d_mask = mask ? (const uint8_t*)sb[5].p : nullptr;
    k_p3_fixed_init<<<tb, threads, 0, st>>>((const char*)sb[4].p, d_mask, (const long long*)sb[2].p, d_tasks, n_tasks,
                                            *s, K, d_fixed, d_state);
    ctx->launches++;
    CK(cudaGetLastError());
    // thal under libprimer3's conditions
    const PoConsts saved = ctx->K;
    ctx->K = K;
    auto restore = [&](int code) {
        ctx->K = saved;
        return code;
    };
    po_thal_out* t0 = (po_thal_out*)sb[19].p;
    po_thal_out* t1 = (po_thal_out*)sb[20].p;
    po_thal_out* t2 = (po_thal_out*)sb[21].p;
    uint4* sa = (uint4*)sb[17].p;
    uint4* sbq = (uint4*)sb[18].p;
#define P3CK(call)                                                                    \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) {                                                      \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);           \
            return restore(PO_E_CUDA);                                                \
        }                                                                             \
    } while (0)
    {
        k_p3_gather_fixed<<<tb, threads, 0, st>>>(d_fixed, n_tasks, sa);
        ctx->launches++;
        P3CK(cudaGetLastError());
        if ((rc = thal_dev<false>(ctx, PO_THAL_ANY, sa, nullptr, n_tasks, t0, st, t1))) return restore(rc);
        if ((rc = thal_dev<false>(ctx, PO_THAL_HAIRPIN, sa, nullptr, n_tasks, t2, st))) return restore(rc);
        k_p3_fixed_self<<<tb, threads, 0, st>>>(n_tasks, t0, t1, t2, *s, d_fixed, d_state);
        ctx->launches++;
        P3CK(cudaGetLastError());
    }
    const unsigned wb = (unsigned)((n_tasks * 32 + threads - 1) / threads);
    // over-proposal per round: need + need / margin_div + 2 candidates (tunable for experiments)
    const char* md = getenv("PO_P3_MARGIN_DIV");
    const int margin_div = md && atoi(md) > 0 ? atoi(md) : 16;
    unsigned long long h_ctr[4];
    for (int round = 0; round < 100000; ++round) {
        P3CK(cudaMemsetAsync(d_ctr, 0, 8 * 4, st));
        k_p3_propose<<<wb, threads, 0, st>>>(d_tasks, n_tasks, *s, d_target, d_list_off, d_pool, d_fixed, d_state,
                                             (int32_t*)sb[12].p, d_ctr, (long long*)sb[15].p, (long long*)sb[14].p,
                                             margin_div);
        ctx->launches++;
        P3CK(cudaGetLastError());
        P3CK(cudaMemcpyAsync(h_ctr, d_ctr, 8 * 4, cudaMemcpyDeviceToHost, st));
        P3CK(cudaStreamSynchronize(st));
        const long long n_self = (long long)h_ctr[0], n_pair = (long long)h_ctr[1];
        if (n_self > 0) {
            const unsigned gb = (unsigned)((n_self + threads - 1) / threads);
            k_p3_gather_self<<<gb, threads, 0, st>>>((const long long*)sb[15].p, n_self, d_pool, sa);
            ctx->launches++;
            if ((rc = thal_dev<false>(ctx, PO_THAL_ANY, sa, nullptr, n_self, t0, st, t1))) return restore(rc);
            if ((rc = thal_dev<false>(ctx, PO_THAL_HAIRPIN, sa, nullptr, n_self, t2, st))) return restore(rc);
            k_p3_scatter_self<<<gb, threads, 0, st>>>((const long long*)sb[15].p, n_self, t0, t1, t2, *s, d_pool);
            ctx->launches++;
            P3CK(cudaGetLastError());
        }
        if (n_pair > 0) {
            const unsigned gb = (unsigned)((n_pair + threads - 1) / threads);
            k_p3_gather_pairs<<<gb, threads, 0, st>>>((const long long*)sb[14].p, n_pair, d_tasks, d_list_off, d_pool,
                                                      d_fixed, (const int32_t*)sb[12].p, sa, sbq);
            ctx->launches++;
            if ((rc = thal_dev<false>(ctx, PO_THAL_ANY, sa, sbq, n_pair, t0, st, t1))) return restore(rc);
            if ((rc = thal_dev<false>(ctx, PO_THAL_END2, sa, sbq, n_pair, t2, st))) return restore(rc);
            k_p3_scatter_pairs<<<gb, threads, 0, st>>>((const long long*)sb[14].p, n_pair, t0, t1, t2,
                                                       (double*)sb[13].p);
            ctx->launches++;
            P3CK(cudaGetLastError());
        }
        k_p3_accept<<<tb, threads, 0, st>>>(d_tasks, n_tasks, *s, d_list_off, d_pool, d_fixed, d_state,
                                            (const int32_t*)sb[12].p, (const double*)sb[13].p, (P3Result*)sb[16].p,
                                            rcap, d_ctr);
        ctx->launches++;
        P3CK(cudaGetLastError());
        P3CK(cudaMemcpyAsync(h_ctr, d_ctr, 8 * 4, cudaMemcpyDeviceToHost, st));
        P3CK(cudaStreamSynchronize(st));
        if (h_ctr[2] == 0) break;
    }
    ctx->K = saved;
    P3CK(cudaMemsetAsync(d_ctr, 0, 8 * 4, st));
    k_p3_order<<<tb, threads, 0, st>>>(d_tasks, n_tasks, *s, d_list_off, d_pool, d_state, (P3Result*)sb[16].p, rcap,
                                       valid ? 1 : 0, d_ctr, (long long*)sb[15].p);
    ctx->launches++;
    CK(cudaGetLastError());
    if (valid) {
        // Primer.is_valid (reference lib_primer3.py:66-97) of every picked primer, under the
        // context's own conditions (ThermoAnalysis()), once per distinct candidate
        CK(cudaMemcpyAsync(h_ctr, d_ctr, 8 * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const long long n_used = (long long)h_ctr[3];
        if (n_used > 0) {
            const unsigned gb = (unsigned)((n_used + threads - 1) / threads);
            if ((rc = ensure(ctx, ctx->props, sizeof(po_oligo_props) * (size_t)n_used))) return rc;
            k_p3_gather_self<<<gb, threads, 0, st>>>((const long long*)sb[15].p, n_used, d_pool, sa);
            ctx->launches++;
            if ((rc = thal_dev<false>(ctx, PO_THAL_HAIRPIN, sa, nullptr, n_used, t0, st))) return rc;
            if ((rc = thal_dev<false>(ctx, PO_THAL_ANY, sa, nullptr, n_used, t1, st))) return rc;
            k_props<<<gb, threads, 0, st>>>(sa, n_used, t0, t1, ctx->K, *valid, 1, (po_oligo_props*)ctx->props.p);
            ctx->launches++;
            k_p3_scatter_valid<<<gb, threads, 0, st>>>((const long long*)sb[15].p, n_used,
                                                       (const po_oligo_props*)ctx->props.p, d_pool);
            ctx->launches++;
            CK(cudaGetLastError());
        }
    }
    k_p3_emit<<<tb, threads, 0, st>>>(d_tasks, n_tasks, *s, d_list_off, d_pool, d_state, (const P3Result*)sb[16].p,
                                      rcap, (po_p3_pair*)sb[24].p, (int32_t*)sb[25].p);
    ctx->launches++;
    CK(cudaGetLastError());
    if (fixed_out) CK(cudaMemcpyAsync(fixed_out, d_fixed, sizeof(po_p3_oligo) * nt, cudaMemcpyDeviceToHost, st));
    if (pairs_out)
        CK(cudaMemcpyAsync(pairs_out, sb[24].p, sizeof(po_p3_pair) * nt * s->num_return, cudaMemcpyDeviceToHost, st));
    if (n_pairs_out) CK(cudaMemcpyAsync(n_pairs_out, sb[25].p, 4 * nt, cudaMemcpyDeviceToHost, st));
    if (!dev) CK(cudaStreamSynchronize(st));
#undef P3CK
    return PO_OK;
}

extern "C" int po_p3_design_fixed_batch(po_ctx* ctx, const po_p3_settings* s, const char* tmpl, const uint8_t* mask,
                                        const int64_t* set_off, const int32_t* set_target, int64_t n_sets,
                                        const po_p3_task* tasks, int64_t n_tasks, const po_valid_params* valid,
                                        po_p3_oligo* fixed_out, po_p3_pair* pairs_out, int32_t* n_pairs_out) {
    if (!ctx || !set_off || n_sets < 0 || n_tasks < 0 || (n_tasks > 0 && (!tasks || !pairs_out || !n_pairs_out)))
        return PO_E_INVALID;
    int rc = p3_check_settings(ctx, s);
    if (rc) return rc;
    if (s->n_size_ranges < 1) {
        ctx->err = "po_p3_design_fixed_batch: PRIMER_PRODUCT_SIZE_RANGE is empty";
        return PO_E_INVALID;
    }
    if (n_tasks == 0) return PO_OK;
    for (int64_t t = 0; t < n_tasks; ++t)
        if (tasks[t].set < 0 || tasks[t].set >= n_sets) {
            ctx->err = "po_p3_design_fixed_batch: task refers to an unknown template";
            return PO_E_INVALID;
        }
    CK(cudaSetDevice(ctx->device));
    return p3_design_fixed_impl(ctx, s, tmpl, mask, set_off, set_target, n_sets, tasks, n_tasks, valid, fixed_out,
                                pairs_out, n_pairs_out, false, 0);
}
