// po_thal_dimer.cuh -- thal type ANY (calcHomodimer / calcHeterodimer) for one
// warp.  Restates upstream thal.c: initMatrix, fillMatrix (LSH, maxTM,
// calc_bulge_internal), the RSH end scan, traceback and drawDimer.
//
// The cells of one row do not depend on each other, so the warp relaxes 32/G of
// them side by side, one group of G lanes each (Grp<G>).  Control flow is kept
// warp-uniform (loops run a uniform number of trips with clamped indices and
// predicated stores): a lane that leaves a loop early would spin on the next
// warp barrier and burn issue slots.
#pragma once
#include "po_thal.cuh"

struct DimerCellCtx {
    int ci, cj;
    int b1q;  // cb * 16 + c2: cell part of the bulge-of-one stack index
};

// upstream calc_bulge_internal(pi, pj, ci, cj): value {dH,dS} of reaching the
// cell through the loop that starts at list entry kp; {inf,-1} when rejected.
template <class WS>
__device__ __forceinline__ double2 dimer_loop_value(const TabAddr& A, const WS& ws, const double2* ycell,
                                                    const DimerCellCtx& c, int kp, int* key_out) {
    const uint32_t code = ws.code[kp];
    const double2 pv = ws.cell[kp];
    const int pb = CODE_B(code);
    const int l1 = c.ci - CODE_I(code) - 1, l2 = c.cj - CODE_J(code) - 1;
    // stack[s1[pi]][s1[ci]][s2[pj]][s2[cj]] = pb*64 + cb*16 + (3-pb)*4 + c2
    const LoopOps o = loop_ops(A, ycell, l1, l2, CODE_TQ(code), pb, pb * 60 + 12 + c.b1q);
    *key_out = o.key;
    // Upstream rejects in three places: a bulge of one whose own term is positive (before the
    // earlier cell is added), a non-finite sum, and any other loop whose sum is positive in both
    // components.  A rejected value is {inf, -1} and stays so through the later steps, so the three
    // tests fold into one select.
    const double H0 = (o.L.x + o.P.x) + o.C.x;
    const double S0 = ((o.L.y + o.P.y) + o.C.y) + o.asymS;
    const double H = H0 + pv.x, S = S0 + pv.y;
    // x > 0 for a non-NaN double is "sign clear and not zero": one 64-bit integer compare
    const bool own_positive = (__double_as_longlong(H0) > 0ll) | (__double_as_longlong(S0) > 0ll);
    const bool sum_positive = (H > 0.0) && (S > 0.0);
    const bool reject = !fin(H) || (o.isB1 ? own_positive : sum_positive);
    return reject ? make_double2(PO_INF, -1.0) : make_double2(H, S);
}

// Candidate predecessors of cell (ci,cj): every finite (pi,pj) with pi < ci,
// pj < cj and 1 <= (ci-pi-1)+(cj-pj-1) <= max_loop.  Virtual run l1 owns row
// ci-1-l1; the row's candidates are one contiguous range of the row-major list.
// Fills the group's run table and cell-side loop operands (ycell, stored
// identically by every lane of the group) and returns the number of candidates.
template <int G, class WS>
__device__ __forceinline__ int dimer_build_runs(const Grp<G>& g, const SmemTables& tb, WS& ws, int ci, int cj,
                                                int max_loop, bool prune, DimerCellCtx* ctx) {
    const int cb = ws.s1[ci], c2 = ws.s2[cj];
    const int q = quad(c2, ws.s2[cj - 1], cb, ws.s1[ci - 1]);
    const double2 y0 = tb.tstack[q], y3 = tb.stackint2[q];
    // When both cell-side terms are infinite every interior / 1x1 candidate is rejected
    // (dH = +inf never wins and never matches in the traceback): only bulges remain, i.e.
    // row ci-1 (l1 == 0) and column cj-1 (l2 == 0).
    const bool interior_ok = fin(y0.x) | fin(y3.x);
    RunTable& rt = ws.runs[g.idx];
    int carry = 0;
#pragma unroll
    for (int sl = 0; sl < 32 / G; ++sl) {
        if (sl * G >= ci - 1) break;  // no predecessor row that far up (warp-uniform: one row per batch)
        const int l1 = sl * G + g.lane, r = ci - 1 - l1;
        const bool have = (r >= 1) & (l1 <= max_loop);
        const int rr = have ? r : 1;
        const int hi_col = (l1 == 0) ? cj - 2 : cj - 1;       // l1 + l2 >= 1
        int lo_col = max(1, cj - 1 - (max_loop - l1));        // l1 + l2 <= max_loop
        if (prune && !interior_ok && l1 > 0) lo_col = max(lo_col, cj - 1);
        const unsigned long long rm = ws.mask[3 - ws.s1[rr]];
        // columns lo_col .. hi_col (1-based) are bits lo_col-1 .. hi_col-1; lo_col >= 1, hi_col <= 60
        const int below = popc_below(rm, lo_col - 1);
        const int cnt = have ? max(popc_below(rm, max(hi_col, 0)) - below, 0) : 0;
        const int first = ws.start[rr] + below;
        const int incl = g.scan_incl(cnt) + carry;
        rt.end[l1] = (uint16_t)incl;
        rt.adj[l1] = (int16_t)(first - (incl - cnt));
        carry = g.shfl(incl, G - 1);
    }
    const bool cAT = (cb == 0) | (cb == 3);
    double2* yc = ws.ycell[g.idx];
    yc[0] = y0;
    yc[1] = make_double2(cAT ? PO_AT_H : 0.0, cAT ? PO_AT_S : 0.0);
    yc[2] = make_double2(0.0, 0.0);
    yc[3] = y3;
    ctx->ci = ci;
    ctx->cj = cj;
    ctx->b1q = cb * 16 + c2;
    g.sync();
    return carry;
}

// Forward scan of the group's candidates: every lane keeps its best (largest T,
// earliest key).  `trip_total` is the warp-wide maximum candidate count.
template <int G, class WS>
__device__ __forceinline__ void dimer_scan(const Grp<G>& g, const TabAddr& A, const WS& ws,
                                           const DimerCellCtx& c, int total, int trip_total, double iH, double iS,
                                           double RC, int mode, double* bT, int* bKey, double2* bV) {
    const RunTable& rt = ws.runs[g.idx];
    const double2* yc = ws.ycell[g.idx];
    double bestT = -PO_INF;
    int bestKey = KEY_NONE;
    double2 bestV = make_double2(PO_INF, -1.0);
    const bool half = c.ci - 1 <= 16;
    for (int t0 = 0; t0 < trip_total; t0 += G) {
        const int t = t0 + g.lane;
        bool act = t < total;
        int key;
        const double2 v = dimer_loop_value(A, ws, yc, c, run_lookup(rt, act ? t : 0, half), &key);
        if (mode == SCAN_NO_STAR) act &= key != KEY_STAR;
        if (mode == SCAN_BEFORE_STAR) act &= key < KEY_STAR;
        const double T1 = po_div(v.x + iH, (v.y + iS) + RC);
        if (act & ((T1 > bestT) | ((T1 == bestT) & (key < bestKey)))) {
            bestT = T1;
            bestKey = key;
            bestV = v;
        }
    }
    *bT = bestT;
    *bKey = bestKey;
    *bV = bestV;
}

template <int G, int CAP, bool COUNT>
__device__ void thal_dimer_warp(const SmemTables& tb, const SmemExtra& X, const TabAddr& A, DimerWs<CAP>& ws,
                                const uint4 oa, const uint4 ob, const PoConsts& K, const bool end_only,
                                po_thal_out* out, po_thal_out* out2, unsigned long long* relax) {
    const int lane = threadIdx.x & 31;
    const Grp<G> g = Grp<G>::make();
    const Grp<32> gw = Grp<32>::make();
    constexpr int NG = 32 / G;
    const int len1 = po_len(oa), len2 = po_len(ob);
    po_thal_out r;
    r.tm = 0.0;
    r.dg = r.dh = r.ds = 0.0;
    r.flags = 0;
    r.align_end_1 = r.align_end_2 = -1;
    if (po_bad(oa) || po_bad(ob) || len1 == 0 || len2 == 0 || len1 > PO_MAX_OLIGO_LEN || len2 > PO_MAX_OLIGO_LEN) {
        r.flags = PO_ITEM_SKIPPED;
        if (lane == 0) {
            *out = r;
            if (out2) *out2 = r;
        }
        return;
    }
    __syncwarp();
    // ---- decode; the second oligo is reversed so it runs 3'->5' ------------
    for (int p = lane; p < 64; p += 32) {
        if (p < len1) ws.s1[p + 1] = (uint8_t)po_base(oa, p);
        if (p < len2) ws.s2[len2 - p] = (uint8_t)po_base(ob, p);
    }
    if (lane == 0) {
        ws.s1[0] = ws.s1[len1 + 1] = 4;
        ws.s2[0] = ws.s2[len2 + 1] = 4;
    }
    // upstream symmetry_thermo() on both oligos
    bool asym = false;
    for (int p = lane; p < 32; p += 32) {
        if (p < len1 / 2) asym |= (po_base(oa, p) + po_base(oa, len1 - 1 - p) != 3);
        if (p < len2 / 2) asym |= (po_base(ob, p) + po_base(ob, len2 - 1 - p) != 3);
    }
    const bool sym = !__any_sync(FULL, asym) && (len1 % 2 == 0) && (len2 % 2 == 0);
    const double RC = sym ? K.rc_sym : K.rc_nonsym;
    const double2* lshT = X.lsh[sym ? 1 : 0];
    const double2* rshT = X.rsh[sym ? 1 : 0];
    const double iH = 200.0, iS = -5.7;
    const int max_loop = K.max_loop;
    __syncwarp();

    // ---- finite cells: bit (j-1) of mask[b] set iff s2[j] == b ----------------
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const unsigned lo = __ballot_sync(FULL, lane + 1 <= len2 && ws.s2[lane + 1] == b);
        const unsigned hi = __ballot_sync(FULL, lane + 33 <= len2 && ws.s2[lane + 33] == b);
        if (lane == 0) ws.mask[b] = (unsigned long long)lo | ((unsigned long long)hi << 32);
    }
    __syncwarp();
#define ROWMASK(i) (ws.mask[3 - ws.s1[i]])
#define LSH_AT(i, j) (lshT[END_IDX(ws.s1[i], ws.s1[(i)-1], ws.s2[(j)-1])])
#define RSH_AT(i, j) (rshT[END_IDX(ws.s1[i], ws.s1[(i) + 1], ws.s2[(j) + 1])])
    // row starts (exclusive prefix of per-row popcounts), two rows per lane
    {
        const int i0 = lane + 1, i1 = lane + 33;
        const int c0 = i0 <= len1 ? __popcll(ROWMASK(i0)) : 0;
        const int c1 = i1 <= len1 ? __popcll(ROWMASK(i1)) : 0;
        const int x0 = gw.scan_incl(c0), x1 = gw.scan_incl(c1);
        const int tot0 = __shfl_sync(FULL, x0, 31);
        ws.start[i0] = (uint16_t)(x0 - c0);
        if (i1 < 64) ws.start[i1] = (uint16_t)(tot0 + x1 - c1);
    }
    __syncwarp();
    const int ncell = ws.start[len1 + 1];
    for (int i = 1; i <= len1; ++i) {
        const unsigned long long m = ROWMASK(i);
        const int base = ws.start[i];
        const int a = ws.s1[i], an = ws.s1[i + 1] & 3;
        const uint32_t rowcode = (uint32_t)i | ((uint32_t)a << 20);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int bit = lane + 32 * h;  // column j = bit + 1
            if ((m >> bit) & 1ull) {
                const int tq = quad(a, an, 3 - a, ws.s2[bit + 2] & 3);
                ws.code[base + __popcll(m & ((1ull << bit) - 1ull))] =
                    rowcode | ((uint32_t)(bit + 1) << 6) | ((uint32_t)tq << 12);
            }
        }
    }
    __syncwarp();

    int cnt = 0;
    // ---- fillMatrix ----------------------------------------------------------
    for (int i = 1; i <= len1; ++i) {
        const int rs = ws.start[i], re = ws.start[i + 1];
        if (rs == re) continue;
        // phase A: LSH seed and maxTM (stack) for every finite cell of the row
        for (int k0 = rs; k0 < re; k0 += 32) {
            const bool act = k0 + lane < re;
            const int k = act ? k0 + lane : re - 1;
            const int j = CODE_J(ws.code[k]);
            double2 v = LSH_AT(i, j);
            if (COUNT) cnt += act ? 1 : 0;
            if (i > 1 && j > 1) {
                if (COUNT) cnt += act ? 1 : 0;
                const double S0 = v.y, H0 = v.x;
                const double2 sh = RSH_AT(i, j);
                const double T0 = po_div(H0 + iH + sh.x, S0 + iS + RC + sh.y);
                double S1 = -1.0, H1 = PO_INF;
                const unsigned long long pm = ROWMASK(i - 1);
                const bool pfin = (pm >> (j - 2)) & 1ull;
                const int kp = pfin ? ws.start[i - 1] + __popcll(pm & ((1ull << (j - 2)) - 1ull)) : 0;
                const double2 pv = ws.cell[kp];
                const double2 st = tb.stack[quad(ws.s1[i - 1], ws.s1[i], ws.s2[j - 1], ws.s2[j])];
                const bool have = pfin && fin(pv.x) && fin(st.x);
                if (have) {
                    S1 = pv.y + st.y;
                    H1 = pv.x + st.x;
                }
                // without a stackable predecessor upstream divides {inf + iH} by {-1 + iS + RC} < 0: -inf
                const double T1 = have ? po_div(H1 + iH + sh.x, S1 + iS + RC + sh.y) : -PO_INF;
                if (S1 < PO_MIN_ENTROPY_CUTOFF) {
                    S1 = PO_MIN_ENTROPY;
                    H1 = 0.0;
                }
                double S0c = S0, H0c = H0;
                if (S0c < PO_MIN_ENTROPY_CUTOFF) {
                    S0c = PO_MIN_ENTROPY;
                    H0c = 0.0;
                }
                if (T1 > T0) v = make_double2(H1, S1);
                else if (T0 >= T1) v = make_double2(H0c, S0c);
            }
            if (act) ws.cell[k] = v;
        }
        __syncwarp();
        if (i == 1) continue;
        // phase B: bulges and interior loops, 32/G cells of the row at a time (column 1 has
        // no predecessor)
        const int kb = rs + (int)(ROWMASK(i) & 1ull);
        for (int k0 = kb; k0 < re; k0 += NG) {
            const bool valid = k0 + g.idx < re;
            const int k = valid ? k0 + g.idx : re - 1;
            const int j = CODE_J(ws.code[k]);
            DimerCellCtx c;
            int total = dimer_build_runs<G>(g, tb, ws, i, j, max_loop, !COUNT, &c);
            total = valid ? total : 0;
            if (COUNT) cnt += (g.lane == 0) ? total : 0;
            const int trip_total = NG == 1 ? total : warp_max_int(total);
            if (trip_total > 0) {
                const double2 cur = ws.cell[k];
                const double Tcur = po_div(cur.x + iH, (cur.y + iS) + RC);
                double bT, wT = -PO_INF;
                int bKey, wKey = KEY_NONE;
                double2 bV;
                dimer_scan<G>(g, A, ws, c, total, trip_total, iH, iS, RC, SCAN_ALL, &bT, &bKey, &bV);
                int src = g.argmax(bT, bKey, &wT, &wKey);
                const bool star = src >= 0 && wKey == KEY_STAR;
                if (__any_sync(FULL, star)) {
                    // The 1x1 loop is the arg-max.  Upstream accepts it only if it beats the
                    // running best by SMALL_NON_ZERO at the moment it is visited; re-derive that
                    // running best (current cell + the three candidates visited earlier).
                    double pT, qT = -PO_INF;
                    int pKey, qKey = KEY_NONE;
                    double2 pV;
                    dimer_scan<G>(g, A, ws, c, total, trip_total, iH, iS, RC, SCAN_BEFORE_STAR, &pT, &pKey, &pV);
                    const int psrc = g.argmax(pT, pKey, &qT, &qKey);
                    const double Tm = (psrc >= 0 && qT > Tcur) ? qT : Tcur;
                    const bool redo = star && !((wT - Tm) >= PO_SMALL_NON_ZERO && wT > Tm);
                    if (__any_sync(FULL, redo)) {
                        double nT, xT = -PO_INF;
                        int nKey, xKey = KEY_NONE;
                        double2 nV;
                        dimer_scan<G>(g, A, ws, c, total, trip_total, iH, iS, RC, SCAN_NO_STAR, &nT, &nKey, &nV);
                        const int nsrc = g.argmax(nT, nKey, &xT, &xKey);
                        if (redo) {
                            bV = nV;
                            src = nsrc;
                            wT = xT;
                            wKey = xKey;
                        }
                    }
                }
                double H = g.shfl(bV.x, max(src, 0)), S = g.shfl(bV.y, max(src, 0));
                if (src >= 0 && wT > Tcur) {
                    if (S < PO_MIN_ENTROPY_CUTOFF) {
                        S = PO_MIN_ENTROPY;
                        H = 0.0;
                    }
                    if (fin(H)) ws.cell[k] = make_double2(H, S);  // the group's lanes store the same value
                }
            }
            __syncwarp();  // run tables are rewritten by the next batch
        }
    }

    // The filled table serves every alignment type: with out2 the ANY result goes to `out` and
    // the END1 result (same fill, end scan over the last row only) to `out2`.
    const int npass = out2 ? 2 : 1;
    for (int pass = 0; pass < npass; ++pass) {
    const bool eo = pass == 1 ? true : end_only;
    po_thal_out* dst = pass == 1 ? out2 : out;
    // ---- end scan: minimum dG over all cells, first in row-major order wins ---
    // (thal types END1 / END2: only the cells of the last row, i == len1)
    double bestG = PO_INF;
    int bestK = KEY_NONE;
    for (int k0 = eo ? ws.start[len1] : 0; k0 < ncell; k0 += 32) {
        const bool act = k0 + lane < ncell;
        const int k = act ? k0 + lane : 0;
        const uint32_t code = ws.code[k];
        const double2 v = ws.cell[k];
        double2 sh = RSH_AT(CODE_I(code), CODE_J(code));
        sh.y = sh.y + PO_SMALL_NON_ZERO;
        sh.x = sh.x + PO_SMALL_NON_ZERO;
        const double G1 = (v.x + sh.x + iH) - PO_TEMP_KELVIN * (v.y + sh.y + iS);
        if (act & (G1 < bestG)) {
            bestG = G1;
            bestK = k;
        }
    }
    if (COUNT) cnt += (lane == 0) ? (eo ? len2 : len1 * len2) : 0;
    double wG = PO_INF;
    int wK = KEY_NONE;
    const int gsrc = gw.argmin(bestG, bestK, &wG, &wK);
    if (COUNT && relax && pass == npass - 1) {
        __syncwarp();
        const unsigned total_cnt = __reduce_add_sync(FULL, (unsigned)cnt);
        if (lane == 0) atomicAdd(relax, (unsigned long long)total_cnt);
    }
    if (gsrc < 0 || !fin(wG)) {
        // upstream falls back to cell (1,1); without a finite value there: no structure, tm = 0
        if (!(ROWMASK(1) & 1ull)) {
            if (lane == 0) *dst = r;
            continue;
        }
        wK = 0;
    }
    int i = CODE_I(ws.code[wK]), j = CODE_J(ws.code[wK]);
    const int bestI = i, bestJ = j;
    double dH, dS;
    {
        const double2 v = ws.cell[wK];
        const double2 sh = RSH_AT(i, j);
        dH = v.x + sh.x + iH;
        dS = v.y + sh.y + iS;
    }

    // ---- traceback: count the base pairs of the chosen structure ---------------
    int npairs = 1;
    for (;;) {
        const unsigned long long rm = ROWMASK(i);
        const int k = ws.start[i] + __popcll(rm & ((1ull << (j - 1)) - 1ull));
        const double2 cur = ws.cell[k];
        const double2 l = LSH_AT(i, j);
        if (eq5(cur.y, l.y) && eq5(cur.x, l.x)) break;
        if (i == 1 || j == 1) break;  // guard: nothing can precede the first row / column
        {
            const unsigned long long pm = ROWMASK(i - 1);
            if ((pm >> (j - 2)) & 1ull) {
                const double2 pv = ws.cell[ws.start[i - 1] + __popcll(pm & ((1ull << (j - 2)) - 1ull))];
                const double2 st = tb.stack[quad(ws.s1[i - 1], ws.s1[i], ws.s2[j - 1], ws.s2[j])];
                if (eq5(cur.y, st.y + pv.y) && eq5(cur.x, st.x + pv.x)) {
                    --i;
                    --j;
                    ++npairs;
                    continue;
                }
            }
        }
        // loops: first candidate in upstream order whose value reproduces the cell
        const double Tcur = po_div(cur.x + iH, (cur.y + iS) + RC);
        DimerCellCtx c;
        const int total = dimer_build_runs<32>(gw, tb, ws, i, j, max_loop, true, &c);
        int myKey = KEY_NONE;
        for (int t0 = 0; t0 < total; t0 += 32) {
            const int t = t0 + lane;
            const bool act = t < total;
            int key;
            const double2 v = dimer_loop_value(A, ws, ws.ycell[0], c, run_lookup(ws.runs[0], act ? t : 0, i - 1 <= 16), &key);
            const double T1 = po_div(v.x + iH, (v.y + iS) + RC);
            // traceback mode of the 1x1 branch reports its value only when T1 >= T2
            const bool ok = act & (key < myKey) & ((key != KEY_STAR) | (T1 >= Tcur));
            if (ok && eq5(cur.y, v.y) && eq5(cur.x, v.x)) myKey = key;
        }
        const int wKey = gw.min_s32(myKey);
        if (wKey == KEY_NONE) break;  // guard: upstream would spin here (unreachable for <= 60 nt)
        const int sum = wKey >> 6, l1 = wKey & 63;
        i = i - l1 - 1;
        j = j - (sum - l1) - 1;
        ++npairs;
    }
    // ---- drawDimer ---------------------------------------------------------------
    const int N = npairs - 1;  // (2 * npairs) / 2 - 1
    const double t = (dH / (dS + (N * K.salt_corr) + RC)) - PO_ABSOLUTE_ZERO;
    const double dG = dH - (K.temp_k * (dS + (N * K.salt_corr)));
    const double S = dS + (N * K.salt_corr);
    po_thal_out res = r;
    res.tm = t;
    res.dg = dG;
    res.dh = dH;
    res.ds = S;
    res.flags = PO_ITEM_STRUCTURE_FOUND | (sym ? PO_ITEM_SYMMETRIC : 0);
    res.align_end_1 = (int16_t)bestI;
    res.align_end_2 = (int16_t)bestJ;
    if (lane == 0) *dst = res;
    }
#undef ROWMASK
#undef LSH_AT
#undef RSH_AT
}
