// po_kasp_design.cuh -- the design part of the reference's KASP worker (src/polyoligo/_lib_kasp.py:
// 671-743, BLAST left out) for a batch of markers behind ONE C entry, po_kasp_design_batch:
// everything between the marker records going in and the top-n primer triples coming out stays on
// the device.
//
//   a3  get_marker_primers           k_kasp_enumerate + props            _lib_kasp.py:295-359
//   a4  search masks                 k_kd_mutbits, k_kd_masks, k_kd_differs
//         include_mut_in_included_maps  lib_primer3.py:576-593 (mutation positions cleared in the
//                                       *_mut twins, numpy negative-index semantics included)
//         marker window                 _lib_kasp.py:681-684, literally: [min(1, start - size):] = True
//         get_sequence_excluded_regions lib_primer3.py:596-600: the picker takes the complement as a
//                                       per-base mask, which is what the [start, len] runs say (the
//                                       <= 200-run reduction of get_primers never triggers here: the
//                                       window leaves at most PRIMER_MAX_SIZE bases closed)
//   a5  designPrimers per marker primer and mask    p3_design_fixed_impl (po_p3.cuh), inputs resident
//   a6  round-robin merge + dedup + validity of the common primer    k_kd_merge   _lib_kasp.py:591-631
//   8f-3 add_mutations (numbers)     k_kd_pairs + k_annotate_mutations  lib_primer3.py:166-210
//   a8/a9 tailed heterodimers, score_kasp, classify, prune           rank core    _lib_kasp.py:89-117,278-291
//
// A search mask that equals an earlier one inside the window a marker's primers can reach gives the
// picker the same question and therefore nothing new to merge: those passes are skipped per marker
// (k_kd_differs), exactly like kasp_pipeline._design_chunk, whose arrays this entry reproduces.
#pragma once
#include "po_common.h"
#include "po_pipeline.cuh"

struct KdEntry {      // one merged (marker primer, common primer) pair
    uint4 other_seq;
    int32_t other_start, other_len;
    int32_t slot;
    int32_t pad;
};

// ---- a3 bookkeeping: which slots survive, where their primers sit, which template window matters ----
// One warp per marker, lane = enumeration slot (slots <= 32).
__global__ void k_kd_geometry(const po_kasp_marker* __restrict__ M, long long n, int slots, int min_size,
                              const po_oligo_props* __restrict__ pr, const po_oligo_props* __restrict__ pa, int maxprod,
                              uint32_t* __restrict__ keepbits, int32_t* __restrict__ p_start, int32_t* __restrict__ p_len,
                              int32_t* __restrict__ window) {
    const long long m = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (m >= n) return;
    const po_kasp_marker& mk = M[m];
    const int L = mk.tmpl_len, rl = mk.ref_len, d = mk.ref_len - mk.alt_len, trim = max(d, 0), ad = abs(d);
    bool keep = false;
    int start = 0, ln = 0;
    const bool is_r = lane & 1;
    if (lane < slots) {
        const long long t = m * slots + lane;
        keep = (pr[t].flags & PO_PRIMER_VALID) && (pa[t].flags & PO_PRIMER_VALID);
        // template position (primer3 convention) and length of the REF marker primer of the slot
        // (reference _lib_kasp.py:309-337; kasp_pipeline._slot_geometry)
        const int pad = min_size + lane / 2;
        const int x2 = mk.marker_pos + rl, x1 = x2 - pad - ad + trim;
        const int y2 = L - mk.marker_pos, y1 = y2 - pad - ad - rl + trim;
        start = is_r ? L - 1 - y1 : x1;
        ln = is_r ? y2 - y1 : x2 - x1;
        p_start[t] = start;
        p_len[t] = ln;
    }
    const unsigned kb = __ballot_sync(0xffffffffu, keep);
    // a primer that can pair with a marker primer lies inside the largest product that starts / ends at it
    const int big = 1 << 30;
    const int lo_pos = (keep && is_r) ? start - maxprod + 1 : big;
    const int hi_pos = (keep && !is_r) ? start + maxprod - 1 : -big;
    const int own_lo = keep ? (is_r ? start - ln + 1 : start) : big;
    const int own_hi = keep ? (is_r ? start : start + ln - 1) : -big;
    const int wlo = __reduce_min_sync(0xffffffffu, min(lo_pos, own_lo));
    const int whi = __reduce_max_sync(0xffffffffu, max(hi_pos, own_hi));
    if (lane == 0) {
        keepbits[m] = kb;
        window[2 * m] = wlo;
        window[2 * m + 1] = whi;
    }
}

// 2-bit template -> the ASCII the picker reads
__global__ void k_kd_templates(const po_kasp_marker* __restrict__ M, long long n, int L, char* __restrict__ ascii) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n * L) return;
    const long long m = t / L;
    const int p = (int)(t % L);
    const po_kasp_marker& mk = M[m];
    const int code = (mk.tmpl[p >> 4] >> (2 * (p & 15))) & 3;
    const bool isn = (mk.nmask[p >> 5] >> (p & 31)) & 1u;
    ascii[t] = isn ? 'N' : "ACGT"[code];
}

// ---- a4 -------------------------------------------------------------------------------------------
// mutation positions of every marker as a bit map: twin[pos] = False with pos = mutation.pos - hroi.start
// (numpy fancy indexing: a negative pos counts from the end)
__global__ void k_kd_mutbits(const long long* __restrict__ mut_pos, const long long* __restrict__ mut_seg,
                             const long long* __restrict__ starts, long long n_markers, long long n_mut, int L,
                             uint32_t* __restrict__ mutbits) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n_mut) return;
    long long lo = 0, hi = n_markers - 1;  // marker of this mutation: last m with mut_seg[m] <= t
    while (lo < hi) {
        const long long mid = (lo + hi + 1) >> 1;
        if (mut_seg[mid] <= t) lo = mid;
        else hi = mid - 1;
    }
    long long pos = mut_pos[t] - starts[lo];
    if (pos < 0) pos += L;
    if (pos >= 0 && pos < L) atomicOr(&mutbits[lo * 16 + (pos >> 5)], 1u << (pos & 31));
}

// excluded-base bit maps of every search type: thread per (marker, 32-base word).
// types (reference _lib_kasp.py:715-718): with mutations mism_mut, mism, partial_mism_mut, partial_mism,
// all_mut, all; without: mism, partial_mism, all.  maps: [n][3][16] words = mism, partial_mism, all.
__global__ void k_kd_masks(const uint32_t* __restrict__ maps, const uint32_t* __restrict__ mutbits,
                           const long long* __restrict__ starts, long long n, int L, int size, int with_mut,
                           int kasp_window, uint32_t* __restrict__ excl) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n * 16) return;
    const long long m = t >> 4;
    const int w = (int)(t & 15);
    // the marker window, literally `mmap[min(1, start - size):max(len, start + size)] = True`
    const long long s_idx = min(1ll, starts[m] - size);
    const int win_lo = s_idx >= 0 ? (int)s_idx : (int)max((long long)L + s_idx, 0ll);
    const int b0 = 32 * w;
    uint32_t win = 0u;
    if (!kasp_window) win = 0u;
    else if (win_lo <= b0) win = 0xffffffffu;
    else if (win_lo < b0 + 32) win = 0xffffffffu << (win_lo - b0);
    const uint32_t valid = L >= b0 + 32 ? 0xffffffffu : (L <= b0 ? 0u : ((1u << (L - b0)) - 1u));
    const uint32_t mut = with_mut ? mutbits[m * 16 + w] : 0u;
    const int n_types = with_mut ? 6 : 3;
    for (int ty = 0; ty < n_types; ++ty) {
        const int base = with_mut ? ty >> 1 : ty;
        const bool twin = with_mut && !(ty & 1);
        uint32_t incl = maps[(m * 3 + base) * 16 + w];
        if (twin) incl &= ~mut;
        incl |= win;
        excl[((long long)ty * n + m) * 16 + w] = ~incl & valid;
    }
}

// differs[ty][m]: the mask of type ty differs from EVERY earlier type's mask somewhere inside the
// marker's window (else the picker would be asked an earlier question again).  Warp per marker.
__global__ void k_kd_differs(const uint32_t* __restrict__ excl, const int32_t* __restrict__ window, long long n,
                             int n_types, uint8_t* __restrict__ differs) {
    const long long m = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (m >= n) return;
    const int wlo = window[2 * m], whi = window[2 * m + 1];
    uint32_t wb = 0u;
    if (lane < 16) {
        const int b0 = 32 * lane;
        const int lo = max(wlo - b0, 0), hi = min(whi - b0, 31);
        if (lo <= hi) wb = (hi >= 31 ? 0xffffffffu : ((1u << (hi + 1)) - 1u)) & (0xffffffffu << lo);
    }
    for (int ty = 0; ty < n_types; ++ty) {
        bool all = true;
        const uint32_t mine = lane < 16 ? excl[((long long)ty * n + m) * 16 + lane] : 0u;
        for (int p = 0; p < ty; ++p) {
            const uint32_t other = lane < 16 ? excl[((long long)p * n + m) * 16 + lane] : 0u;
            all = all && __any_sync(0xffffffffu, ((mine ^ other) & wb) != 0u);
        }
        if (lane == 0) differs[(long long)ty * n + m] = all ? 1 : 0;
    }
}

__global__ void k_kd_expand_mask(const uint32_t* __restrict__ excl_ty, long long n, int L, uint8_t* __restrict__ bytes) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n * L) return;
    const long long m = t / L;
    const int p = (int)(t % L);
    bytes[t] = (excl_ty[m * 16 + (p >> 5)] >> (p & 31)) & 1u;
}

// ---- a5 task lists ----------------------------------------------------------------------------------
__global__ void k_kd_task_count(const uint32_t* __restrict__ keepbits, const uint8_t* __restrict__ differs_ty,
                                const int32_t* __restrict__ count, int depth, long long n, int32_t* __restrict__ cnt) {
    const long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (m >= n) return;
    cnt[m] = (differs_ty[m] && count[m] < depth) ? __popc(keepbits[m]) : 0;
}

// exclusive prefix sum of n <= 1 << 20 int32 values by one block; out[n] = total (as int64 when WIDE)
template <typename OutT>
__global__ void k_kd_scan(const int32_t* __restrict__ in, long long n, OutT* __restrict__ out) {
    __shared__ long long part[1024];
    const int tid = threadIdx.x;
    const long long per = (n + blockDim.x - 1) / blockDim.x;
    const long long s = tid * per, e = min(n, s + per);
    long long sum = 0;
    for (long long k = s; k < e; ++k) sum += in[k];
    part[tid] = sum;
    __syncthreads();
    for (int off = 1; off < (int)blockDim.x; off <<= 1) {
        long long v = tid >= off ? part[tid - off] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    long long run = part[tid] - sum;
    for (long long k = s; k < e; ++k) {
        out[k] = (OutT)run;
        run += in[k];
    }
    if (tid == (int)blockDim.x - 1) out[n] = (OutT)part[tid];
}

__global__ void k_kd_task_fill(const uint32_t* __restrict__ keepbits, const int32_t* __restrict__ cnt,
                               const int32_t* __restrict__ task_off, const int32_t* __restrict__ p_start,
                               const int32_t* __restrict__ p_len, int slots, long long n, po_p3_task* __restrict__ tasks) {
    const long long m = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (m >= n || cnt[m] == 0) return;
    const uint32_t kb = keepbits[m];
    if ((kb >> lane) & 1u) {
        po_p3_task tk;
        tk.set = (int32_t)m;
        tk.side = 1 + (lane & 1);
        tk.start = p_start[m * slots + lane];
        tk.len = p_len[m * slots + lane];
        tasks[task_off[m] + __popc(kb & ((1u << lane) - 1u))] = tk;
    }
}

// ---- a6: round-robin merge (reference _lib_kasp.py:591-631) -------------------------------------------
// Warp per marker, lane = task of the marker (slot order).  Round r pops every task's r-th pair in
// marker-primer order; a pair is kept iff the common primer is valid (Primer.is_valid) and the
// (marker primer, common primer) combination is not in the repository yet; the repository stops at
// `depth` pairs.
__global__ void k_kd_merge(const uint32_t* __restrict__ keepbits, const int32_t* __restrict__ cnt,
                           const int32_t* __restrict__ task_off, const po_p3_pair* __restrict__ pairs,
                           const int32_t* __restrict__ n_pairs, int depth, long long n, int32_t* __restrict__ count,
                           KdEntry* __restrict__ merged) {
    const long long m = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (m >= n || cnt[m] == 0) return;
    const uint32_t kb = keepbits[m];
    const int nt = cnt[m];
    // lane -> its slot: the lane-th set bit of keepbits
    int slot = -1;
    {
        uint32_t b = kb;
        for (int k = 0; k < lane && b; ++k) b &= b - 1;
        if (lane < nt && b) slot = __ffs(b) - 1;
    }
    const long long t = task_off[m] + lane;
    const int np = lane < nt ? n_pairs[t] : 0;
    const int before = count[m];  // entries of earlier passes: the dedup set
    int total = before;
    KdEntry* out = merged + m * depth;
    const int rounds = __reduce_max_sync(0xffffffffu, np);
    for (int r = 0; r < rounds && total < depth; ++r) {
        bool ok = false;
        po_p3_oligo o;
        if (r < np) {
            o = pairs[t * depth + r].other;
            ok = (o.flags & PO_P3_PRIMER_VALID) != 0;
            for (int e = 0; ok && e < before; ++e)
                ok = !(out[e].slot == slot && out[e].other_start == o.start && out[e].other_len == o.len);
        }
        const unsigned okb = __ballot_sync(0xffffffffu, ok);
        const int pos = total + __popc(okb & ((1u << lane) - 1u));
        if (ok && pos < depth) {
            KdEntry en;
            memcpy(&en.other_seq, &o.seq, 16);
            en.other_start = o.start;
            en.other_len = o.len;
            en.slot = slot;
            en.pad = 0;
            out[pos] = en;
        }
        total = min(depth, total + __popc(okb));
    }
    if (lane == 0) count[m] = total;
}

// ---- assemble the merged pairs: F / R / A records, direction, genome coordinates ------------------------
// (get_primers :505-516 rebases the picked primer, merge_primers :559-562 keeps the marker primer's own
// coordinates)
__global__ void k_kd_pairs(const KdEntry* __restrict__ merged, const int32_t* __restrict__ count,
                           const long long* __restrict__ seg, const long long* __restrict__ starts,
                           const uint4* __restrict__ ref, const uint4* __restrict__ alt,
                           const int32_t* __restrict__ p_start, const int32_t* __restrict__ p_len, int slots, int depth,
                           long long n, uint4* __restrict__ F, uint4* __restrict__ R, uint4* __restrict__ A,
                           uint8_t* __restrict__ dir, long long* __restrict__ pstart, long long* __restrict__ pstop,
                           int32_t* __restrict__ n_off) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n * depth) return;
    const long long m = t / depth;
    const int e = (int)(t % depth);
    if (e >= count[m]) return;
    const KdEntry en = merged[t];
    const long long idx = seg[m] + e;
    const bool is_r = en.slot & 1;
    const uint4 mark = ref[m * slots + en.slot];
    F[idx] = is_r ? en.other_seq : mark;
    R[idx] = is_r ? mark : en.other_seq;
    A[idx] = alt[m * slots + en.slot];
    dir[idx] = is_r ? 1 : 0;
    n_off[idx] = 0;
    const long long g0 = starts[m];
    const long long ms = p_start[m * slots + en.slot], ml = p_len[m * slots + en.slot];
    const long long os = en.other_start, ol = en.other_len;
    const long long left_s = is_r ? os : ms, left_l = is_r ? ol : ml;
    const long long right_s = is_r ? ms : os, right_l = is_r ? ml : ol;
    const long long f_start = g0 + left_s, f_stop = g0 + left_s + left_l - 1;
    const long long r_stop = g0 + right_s, r_start = r_stop - (right_l - 1);
    pstart[3 * idx] = f_start;
    pstart[3 * idx + 1] = r_start;
    pstart[3 * idx + 2] = is_r ? r_start : f_start;
    pstop[3 * idx] = f_stop;
    pstop[3 * idx + 1] = r_stop;
    pstop[3 * idx + 2] = is_r ? r_stop : f_stop;
}

// tails of the three primers by direction (add_reporter_dyes, reference _lib_kasp.py:75-86):
// dir F -> F carries dye 1, A dye 2, R none; dir R -> R carries dye 1, A dye 2, F none
__global__ void k_kd_tail_sel(const uint8_t* __restrict__ dir, long long n, int8_t* __restrict__ sel) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n) return;
    const bool dirF = dir[t] == 0;
    sel[t] = dirF ? 0 : -1;
    sel[n + t] = dirF ? -1 : 0;
    sel[2 * n + t] = 1;
}

__global__ void k_kd_emit(const long long* __restrict__ sel, const int32_t* __restrict__ n_sel,
                          const uint4* __restrict__ F, const uint4* __restrict__ R, const uint4* __restrict__ A,
                          const uint8_t* __restrict__ dir, const int32_t* __restrict__ goodness,
                          const int32_t* __restrict__ qbits, int top_n, long long n, po_kasp_pair* __restrict__ out) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n * top_n) return;
    const long long m = t / top_n;
    const int j = (int)(t % top_n);
    po_kasp_pair o;
    memset(&o, 0, sizeof(o));
    if (j < n_sel[m]) {
        const long long idx = sel[t];
        memcpy(&o.f, &F[idx], 16);
        memcpy(&o.r, &R[idx], 16);
        memcpy(&o.a, &A[idx], 16);
        o.goodness = goodness[idx];
        o.qbits = qbits[idx];
        o.dir = dir[idx];
    }
    out[t] = o;
}

// ======================================================================================================
// host side
// ======================================================================================================
static int kd_chunk(po_ctx* ctx, const po_p3_settings* S, const po_kasp_design_params* P,
                    const po_kasp_marker* markers, const int64_t* starts, int64_t n, const uint32_t* maps,
                    const int64_t* mut_pos, const double* mut_aaf, const int32_t* mut_indel, const int64_t* mut_seg,
                    int64_t mut_base, po_kasp_pair* out, int32_t* n_out, int32_t* n_merged) {
    cudaStream_t st = ctx->stream;
    DevBuf* kd = ctx->kd;
    int rc;
    const int slots = 2 * (P->max_size - P->min_size + 1);
    const int L = markers[0].tmpl_len;
    const int depth = S->num_return, top_n = P->top_n;
    const int n_types = P->with_mutations ? 6 : 3;
    const size_t ns = (size_t)n * slots, nL = (size_t)n * L;
    const int64_t n_mut = P->with_mutations ? mut_seg[n] - mut_base : 0;
    enum { B_MARK = 0, B_START, B_MAPS, B_MPOS, B_MAAF, B_MIND, B_MSEG, B_REF, B_ALT, B_PR, B_PA, B_KEEP, B_PSTART,
           B_PLEN, B_WIN, B_ASCII, B_SETOFF, B_MUTB, B_EXCL, B_DIFF, B_BYTES, B_COUNT, B_CNT, B_TOFF, B_TASKS,
           B_MERGED, B_SEG, B_F, B_R, B_A, B_DIR, B_PS, B_PE, B_AAF, B_IND, B_NOFF, B_TF, B_TR, B_TA, B_SEL8, B_FT,
           B_RT, B_AT, B_X, B_H1, B_H2, B_GOOD, B_SEL, B_NSEL, B_OUT, B_TOTAL };
    static_assert(B_TOTAL <= 56, "po_ctx::kd too small");
#define KD_ENSURE(b, bytes) if ((rc = ensure(ctx, kd[b], (bytes) > 0 ? (size_t)(bytes) : 16))) return rc
    KD_ENSURE(B_MARK, sizeof(po_kasp_marker) * n);
    KD_ENSURE(B_START, 8 * n);
    KD_ENSURE(B_MAPS, 4 * 48 * n);
    KD_ENSURE(B_MPOS, 8 * n_mut);
    KD_ENSURE(B_MAAF, 8 * n_mut);
    KD_ENSURE(B_MIND, 4 * n_mut);
    KD_ENSURE(B_MSEG, 8 * (n + 1));
    KD_ENSURE(B_REF, 16 * ns);
    KD_ENSURE(B_ALT, 16 * ns);
    KD_ENSURE(B_PR, sizeof(po_oligo_props) * ns);
    KD_ENSURE(B_PA, sizeof(po_oligo_props) * ns);
    KD_ENSURE(B_KEEP, 4 * n);
    KD_ENSURE(B_PSTART, 4 * ns);
    KD_ENSURE(B_PLEN, 4 * ns);
    KD_ENSURE(B_WIN, 8 * n);
    KD_ENSURE(B_ASCII, nL);
    KD_ENSURE(B_SETOFF, 8 * (n + 1));
    KD_ENSURE(B_MUTB, 64 * n);
    KD_ENSURE(B_EXCL, 64 * n * n_types);
    KD_ENSURE(B_DIFF, n * n_types);
    KD_ENSURE(B_BYTES, nL);
    KD_ENSURE(B_COUNT, 4 * n);
    KD_ENSURE(B_CNT, 4 * n);
    KD_ENSURE(B_TOFF, 4 * (n + 1));
    KD_ENSURE(B_TASKS, sizeof(po_p3_task) * ns);
    KD_ENSURE(B_MERGED, sizeof(KdEntry) * (size_t)n * depth);
    KD_ENSURE(B_SEG, 8 * (n + 1));
    KD_ENSURE(B_NSEL, 4 * n);
    KD_ENSURE(B_SEL, 8 * (size_t)n * top_n);
    KD_ENSURE(B_OUT, sizeof(po_kasp_pair) * (size_t)n * top_n);
    const int threads = 256;
    auto blocks = [&](long long work) { return (unsigned)((work + threads - 1) / threads); };
    // ---- inputs ---------------------------------------------------------------------------------------
    CK(cudaMemcpyAsync(kd[B_MARK].p, markers, sizeof(po_kasp_marker) * n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(kd[B_START].p, starts, 8 * n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(kd[B_MAPS].p, maps, 4 * 48 * n, cudaMemcpyHostToDevice, st));
    std::vector<long long> h_off((size_t)n + 1), h_mseg((size_t)n + 1, 0);
    for (int64_t k = 0; k <= n; ++k) h_off[k] = k * (long long)L;
    if (P->with_mutations)
        for (int64_t k = 0; k <= n; ++k) h_mseg[k] = mut_seg[k] - mut_base;
    CK(cudaMemcpyAsync(kd[B_SETOFF].p, h_off.data(), 8 * (n + 1), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(kd[B_MSEG].p, h_mseg.data(), 8 * (n + 1), cudaMemcpyHostToDevice, st));
    if (n_mut > 0) {
        CK(cudaMemcpyAsync(kd[B_MPOS].p, mut_pos + mut_base, 8 * n_mut, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(kd[B_MAAF].p, mut_aaf + mut_base, 8 * n_mut, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(kd[B_MIND].p, mut_indel + mut_base, 4 * n_mut, cudaMemcpyHostToDevice, st));
    }
    const po_kasp_marker* dM = (const po_kasp_marker*)kd[B_MARK].p;
    const long long* dStart = (const long long*)kd[B_START].p;
    const long long* dMseg = (const long long*)kd[B_MSEG].p;
    uint4* dRef = (uint4*)kd[B_REF].p;
    uint4* dAlt = (uint4*)kd[B_ALT].p;
    // ---- a3 + a1/a2 -----------------------------------------------------------------------------------
    k_kasp_enumerate<<<blocks((long long)ns), threads, 0, st>>>(dM, n, P->min_size, P->max_size, dRef, dAlt);
    ctx->launches++;
    if ((rc = ensure(ctx, ctx->out, sizeof(po_thal_out) * ns))) return rc;
    if ((rc = ensure(ctx, ctx->out2, sizeof(po_thal_out) * ns))) return rc;
    for (int which = 0; which < 2; ++which) {
        const uint4* src = which ? dAlt : dRef;
        if ((rc = thal_dev<false>(ctx, PO_THAL_HAIRPIN, src, nullptr, (int64_t)ns, ctx->out.p, st))) return rc;
        if ((rc = thal_dev<false>(ctx, PO_THAL_ANY, src, nullptr, (int64_t)ns, ctx->out2.p, st))) return rc;
        k_props<<<blocks((long long)ns), threads, 0, st>>>(src, (long long)ns, (const po_thal_out*)ctx->out.p,
                                                          (const po_thal_out*)ctx->out2.p, ctx->K, P->valid, 1,
                                                          (po_oligo_props*)kd[which ? B_PA : B_PR].p);
        ctx->launches++;
    }
    int maxprod = 0;
    for (int k = 0; k < S->n_size_ranges; ++k) maxprod = S->size_hi[k] > maxprod ? S->size_hi[k] : maxprod;
    k_kd_geometry<<<blocks(32 * n), threads, 0, st>>>(dM, n, slots, P->min_size, (const po_oligo_props*)kd[B_PR].p,
                                                      (const po_oligo_props*)kd[B_PA].p, maxprod, (uint32_t*)kd[B_KEEP].p,
                                                      (int32_t*)kd[B_PSTART].p, (int32_t*)kd[B_PLEN].p,
                                                      (int32_t*)kd[B_WIN].p);
    k_kd_templates<<<blocks((long long)nL), threads, 0, st>>>(dM, n, L, (char*)kd[B_ASCII].p);
    ctx->launches += 2;
    // ---- a4 -------------------------------------------------------------------------------------------
    CK(cudaMemsetAsync(kd[B_MUTB].p, 0, 64 * n, st));
    if (n_mut > 0) {
        k_kd_mutbits<<<blocks(n_mut), threads, 0, st>>>((const long long*)kd[B_MPOS].p, dMseg, dStart, n, n_mut, L,
                                                        (uint32_t*)kd[B_MUTB].p);
        ctx->launches++;
    }
    k_kd_masks<<<blocks(16 * n), threads, 0, st>>>((const uint32_t*)kd[B_MAPS].p, (const uint32_t*)kd[B_MUTB].p, dStart, n,
                                                   L, P->max_size, P->with_mutations, 1, (uint32_t*)kd[B_EXCL].p);
    k_kd_differs<<<blocks(32 * n), threads, 0, st>>>((const uint32_t*)kd[B_EXCL].p, (const int32_t*)kd[B_WIN].p, n,
                                                     n_types, (uint8_t*)kd[B_DIFF].p);
    ctx->launches += 2;
    CK(cudaGetLastError());
    // ---- a5 + a6, one pass per search type ---------------------------------------------------------------
    CK(cudaMemsetAsync(kd[B_COUNT].p, 0, 4 * n, st));
    for (int ty = 0; ty < n_types; ++ty) {
        k_kd_task_count<<<blocks(n), threads, 0, st>>>((const uint32_t*)kd[B_KEEP].p, (const uint8_t*)kd[B_DIFF].p + (size_t)ty * n,
                                                       (const int32_t*)kd[B_COUNT].p, depth, n, (int32_t*)kd[B_CNT].p);
        k_kd_scan<int32_t><<<1, 1024, 0, st>>>((const int32_t*)kd[B_CNT].p, n, (int32_t*)kd[B_TOFF].p);
        ctx->launches += 2;
        int32_t n_tasks = 0;
        CK(cudaMemcpyAsync(&n_tasks, (int32_t*)kd[B_TOFF].p + n, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (n_tasks == 0) continue;
        k_kd_task_fill<<<blocks(32 * n), threads, 0, st>>>((const uint32_t*)kd[B_KEEP].p, (const int32_t*)kd[B_CNT].p,
                                                           (const int32_t*)kd[B_TOFF].p, (const int32_t*)kd[B_PSTART].p,
                                                           (const int32_t*)kd[B_PLEN].p, slots, n, (po_p3_task*)kd[B_TASKS].p);
        k_kd_expand_mask<<<blocks((long long)nL), threads, 0, st>>>((const uint32_t*)kd[B_EXCL].p + (size_t)ty * n * 16, n, L,
                                                                    (uint8_t*)kd[B_BYTES].p);
        ctx->launches += 2;
        CK(cudaGetLastError());
        if ((rc = p3_design_fixed_impl(ctx, S, (const char*)kd[B_ASCII].p, (const uint8_t*)kd[B_BYTES].p,
                                       (const int64_t*)kd[B_SETOFF].p, nullptr, n, (const po_p3_task*)kd[B_TASKS].p, n_tasks,
                                       &P->valid, nullptr, nullptr, nullptr, true, (long long)nL)))
            return rc;
        k_kd_merge<<<blocks(32 * n), threads, 0, st>>>((const uint32_t*)kd[B_KEEP].p, (const int32_t*)kd[B_CNT].p,
                                                       (const int32_t*)kd[B_TOFF].p, (const po_p3_pair*)ctx->p3[24].p,
                                                       (const int32_t*)ctx->p3[25].p, depth, n, (int32_t*)kd[B_COUNT].p,
                                                       (KdEntry*)kd[B_MERGED].p);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    // ---- merged pairs -> arrays --------------------------------------------------------------------------
    k_kd_scan<long long><<<1, 1024, 0, st>>>((const int32_t*)kd[B_COUNT].p, n, (long long*)kd[B_SEG].p);
    ctx->launches++;
    long long np = 0;
    CK(cudaMemcpyAsync(&np, (long long*)kd[B_SEG].p + n, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const size_t nn = (size_t)(np > 0 ? np : 1);
    KD_ENSURE(B_F, 16 * nn);
    KD_ENSURE(B_R, 16 * nn);
    KD_ENSURE(B_A, 16 * nn);
    KD_ENSURE(B_DIR, nn);
    KD_ENSURE(B_PS, 24 * nn);
    KD_ENSURE(B_PE, 24 * nn);
    KD_ENSURE(B_AAF, 24 * nn);
    KD_ENSURE(B_IND, 4 * nn);
    KD_ENSURE(B_NOFF, 4 * nn);
    KD_ENSURE(B_TF, 8 * nn);
    KD_ENSURE(B_TR, 8 * nn);
    KD_ENSURE(B_TA, 8 * nn);
    KD_ENSURE(B_SEL8, 3 * nn);
    KD_ENSURE(B_FT, 16 * nn);
    KD_ENSURE(B_RT, 16 * nn);
    KD_ENSURE(B_AT, 16 * nn);
    KD_ENSURE(B_X, 32 * nn);
    KD_ENSURE(B_H1, sizeof(po_thal_out) * nn);
    KD_ENSURE(B_H2, sizeof(po_thal_out) * nn);
    KD_ENSURE(B_GOOD, 8 * nn);
    const uint4* dF = (const uint4*)kd[B_F].p;
    const uint4* dR = (const uint4*)kd[B_R].p;
    const uint4* dA = (const uint4*)kd[B_A].p;
    if (np > 0) {
        k_kd_pairs<<<blocks((long long)n * depth), threads, 0, st>>>(
            (const KdEntry*)kd[B_MERGED].p, (const int32_t*)kd[B_COUNT].p, (const long long*)kd[B_SEG].p, dStart, dRef, dAlt,
            (const int32_t*)kd[B_PSTART].p, (const int32_t*)kd[B_PLEN].p, slots, depth, n, (uint4*)kd[B_F].p, (uint4*)kd[B_R].p,
            (uint4*)kd[B_A].p, (uint8_t*)kd[B_DIR].p, (long long*)kd[B_PS].p, (long long*)kd[B_PE].p, (int32_t*)kd[B_NOFF].p);
        ctx->launches++;
        // ---- add_mutations (numbers) ----
        if (P->with_mutations) {
            k_annotate_mutations<<<(unsigned)((np + 127) / 128), 128, 0, st>>>(
                (const long long*)kd[B_MPOS].p, (const double*)kd[B_MAAF].p, (const int32_t*)kd[B_MIND].p, dMseg,
                (const long long*)kd[B_SEG].p, n, (const long long*)kd[B_PS].p, (const long long*)kd[B_PE].p, 3, np,
                (double*)kd[B_AAF].p, (int32_t*)kd[B_IND].p);
            ctx->launches++;
        } else {
            CK(cudaMemsetAsync(kd[B_AAF].p, 0, 24 * nn, st));
            CK(cudaMemsetAsync(kd[B_IND].p, 0, 4 * nn, st));
        }
        // ---- reporter-tailed heterodimers, score_kasp ----
        if ((rc = po_tm_batch_dev(ctx, dF, np, kd[B_TF].p, nullptr, st))) return rc;
        if ((rc = po_tm_batch_dev(ctx, dR, np, kd[B_TR].p, nullptr, st))) return rc;
        if ((rc = po_tm_batch_dev(ctx, dA, np, kd[B_TA].p, nullptr, st))) return rc;
        int8_t* s8 = (int8_t*)kd[B_SEL8].p;
        k_kd_tail_sel<<<blocks(np), threads, 0, st>>>((const uint8_t*)kd[B_DIR].p, np, s8);
        uint4 t0, t1;
        memcpy(&t0, &P->reporters[0], 16);
        memcpy(&t1, &P->reporters[1], 16);
        uint4* Ft = (uint4*)kd[B_FT].p;
        uint4* Rt = (uint4*)kd[B_RT].p;
        uint4* At = (uint4*)kd[B_AT].p;
        k_concat_tail<<<blocks(np), threads, 0, st>>>(dF, s8, t0, t1, np, Ft);
        k_concat_tail<<<blocks(np), threads, 0, st>>>(dR, s8 + np, t0, t1, np, Rt);
        k_concat_tail<<<blocks(np), threads, 0, st>>>(dA, s8 + 2 * np, t0, t1, np, At);
        ctx->launches += 4;
        if ((rc = thal_dev<false>(ctx, PO_THAL_ANY, Ft, Rt, np, kd[B_H1].p, st))) return rc;
        uint4* X1 = (uint4*)kd[B_X].p;
        uint4* X2 = X1 + np;
        k_select_pair<<<blocks(np), threads, 0, st>>>(Ft, Rt, At, s8, np, X1, X2);
        ctx->launches++;
        if ((rc = thal_dev<false>(ctx, PO_THAL_ANY, X1, X2, np, kd[B_H2].p, st))) return rc;
        RankArgs ra;
        ra.assay = PO_ASSAY_KASP;
        ra.n = np;
        ra.tmF = (const double*)kd[B_TF].p;
        ra.tmR = (const double*)kd[B_TR].p;
        ra.tmA = (const double*)kd[B_TA].p;
        ra.het1 = (const po_thal_out*)kd[B_H1].p;
        ra.het2 = (const po_thal_out*)kd[B_H2].p;
        ra.n_off = (const int32_t*)kd[B_NOFF].p;
        ra.aaf = (const double*)kd[B_AAF].p;
        ra.indel = (const int32_t*)kd[B_IND].p;
        ra.hrm_delta = nullptr;
        ra.tm_delta = P->tm_delta;
        ra.goodness = (int32_t*)kd[B_GOOD].p;
        ra.qbits = ra.goodness + np;
        k_score_pairs<<<blocks(np), threads, 0, st>>>(ra);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    k_top_n<<<blocks(32 * n), threads, 0, st>>>((const int32_t*)kd[B_GOOD].p, nullptr, 0, (const long long*)kd[B_SEG].p, n,
                                                top_n, (long long*)kd[B_SEL].p, (int32_t*)kd[B_NSEL].p);
    k_kd_emit<<<blocks((long long)n * top_n), threads, 0, st>>>(
        (const long long*)kd[B_SEL].p, (const int32_t*)kd[B_NSEL].p, dF, dR, dA, (const uint8_t*)kd[B_DIR].p,
        (const int32_t*)kd[B_GOOD].p, (const int32_t*)kd[B_GOOD].p + np, top_n, n, (po_kasp_pair*)kd[B_OUT].p);
    ctx->launches += 2;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, kd[B_OUT].p, sizeof(po_kasp_pair) * (size_t)n * top_n, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(n_out, kd[B_NSEL].p, 4 * n, cudaMemcpyDeviceToHost, st));
    if (n_merged) CK(cudaMemcpyAsync(n_merged, kd[B_COUNT].p, 4 * n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
#undef KD_ENSURE
    return PO_OK;
}

extern "C" int po_kasp_design_batch(po_ctx* ctx, const po_p3_settings* picker, const po_kasp_design_params* P,
                                    const po_kasp_marker* markers, const int64_t* starts, int64_t n_markers,
                                    const uint32_t* maps, const int64_t* mut_pos, const double* mut_aaf,
                                    const int32_t* mut_indel, const int64_t* mut_seg, po_kasp_pair* out,
                                    int32_t* n_out, int32_t* n_merged) {
    if (!ctx || !picker || !P || n_markers < 0 || (n_markers > 0 && (!markers || !starts || !maps || !out || !n_out)))
        return PO_E_INVALID;
    int rc = p3_check_settings(ctx, picker);
    if (rc) return rc;
    const int slots = 2 * (P->max_size - P->min_size + 1);
    if (P->min_size < 1 || P->max_size < P->min_size || slots > 32 || P->top_n < 1 || picker->n_size_ranges < 1 ||
        (P->with_mutations && !mut_seg)) {
        ctx->err = "po_kasp_design_batch: invalid parameters (at most 16 primer sizes, top_n >= 1, a product size range)";
        return PO_E_INVALID;
    }
    if (n_markers == 0) return PO_OK;
    const int L = markers[0].tmpl_len;
    for (int64_t k = 0; k < n_markers; ++k)
        if (markers[k].tmpl_len != L || L < 1 || L > 512) {
            ctx->err = "po_kasp_design_batch: templates must share one length of 1..512 nt";
            return PO_E_INVALID;
        }
    CK(cudaSetDevice(ctx->device));
    const char* cs = getenv("PO_KASP_CHUNK");
    const int64_t chunk = cs && atoll(cs) > 0 ? atoll(cs) : 4096;
    for (int64_t s = 0; s < n_markers; s += chunk) {
        const int64_t m = (n_markers - s) < chunk ? (n_markers - s) : chunk;
        const int64_t mut_base = P->with_mutations ? mut_seg[s] : 0;
        if ((rc = kd_chunk(ctx, picker, P, markers + s, starts + s, m, maps + 48 * s, mut_pos, mut_aaf, mut_indel,
                           P->with_mutations ? mut_seg + s : nullptr, mut_base, out + s * P->top_n, n_out + s,
                           n_merged ? n_merged + s : nullptr)))
            return rc;
    }
    return PO_OK;
}

extern "C" int po_kasp_masks_batch(po_ctx* ctx, const int64_t* starts, int64_t n, int32_t tmpl_len, const uint32_t* maps,
                                   const int64_t* mut_pos, const int64_t* mut_seg, int32_t with_mutations,
                                   int32_t kasp_window, int32_t max_size, uint32_t* excl) {
    if (!ctx || n < 0 || tmpl_len < 1 || tmpl_len > 512 || (n > 0 && (!starts || !maps || !excl)) ||
        (with_mutations && !mut_seg))
        return PO_E_INVALID;
    if (n == 0) return PO_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    DevBuf* kd = ctx->kd;
    int rc;
    const int n_types = with_mutations ? 6 : 3;
    const int64_t n_mut = with_mutations ? mut_seg[n] : 0;
    if ((rc = ensure(ctx, kd[1], 8 * n))) return rc;
    if ((rc = ensure(ctx, kd[2], 4 * 48 * n))) return rc;
    if ((rc = ensure(ctx, kd[3], n_mut > 0 ? 8 * n_mut : 16))) return rc;
    if ((rc = ensure(ctx, kd[6], 8 * (n + 1)))) return rc;
    if ((rc = ensure(ctx, kd[17], 64 * n))) return rc;
    if ((rc = ensure(ctx, kd[18], 64 * n * n_types))) return rc;
    CK(cudaMemcpyAsync(kd[1].p, starts, 8 * n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(kd[2].p, maps, 4 * 48 * n, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(kd[17].p, 0, 64 * n, st));
    const int threads = 256;
    if (n_mut > 0) {
        CK(cudaMemcpyAsync(kd[3].p, mut_pos, 8 * n_mut, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(kd[6].p, mut_seg, 8 * (n + 1), cudaMemcpyHostToDevice, st));
        k_kd_mutbits<<<(unsigned)((n_mut + threads - 1) / threads), threads, 0, st>>>(
            (const long long*)kd[3].p, (const long long*)kd[6].p, (const long long*)kd[1].p, n, n_mut, tmpl_len,
            (uint32_t*)kd[17].p);
        ctx->launches++;
    }
    k_kd_masks<<<(unsigned)((16 * n + threads - 1) / threads), threads, 0, st>>>(
        (const uint32_t*)kd[2].p, (const uint32_t*)kd[17].p, (const long long*)kd[1].p, n, tmpl_len, max_size,
        with_mutations, kasp_window, (uint32_t*)kd[18].p);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(excl, kd[18].p, 64 * n * n_types, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return PO_OK;
}
