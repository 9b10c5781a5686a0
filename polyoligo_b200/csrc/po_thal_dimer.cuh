// po_thal_dimer.cuh -- thal type ANY (calcHomodimer / calcHeterodimer) for one
// warp.  Restates upstream thal.c: initMatrix, fillMatrix (LSH, maxTM,
// calc_bulge_internal), the RSH end scan, traceback and drawDimer.
#pragma once
#include "po_thal.cuh"

// One loop candidate of calc_bulge_internal(): predecessor (pi,pj) with value
// {Hp,Sp}, current cell (ci,cj).  Returns {dH,dS} after upstream's validity
// rewrites ({inf,-1} when rejected).
__device__ __forceinline__ double2 dimer_loop_term(const SmemTables& tb, const uint8_t* s1, const uint8_t* s2, int pi,
                                                   int pj, int ci, int cj, int l1, int l2, double2 pv) {
    const int ls = l1 + l2 - 1;
    double H, S;
    if (l1 == 0 || l2 == 0) {
        if (l1 + l2 == 1) {  // bulge of one: the intervening nearest-neighbour pair is added
            const double2 st = tb.stack[quad(s1[pi], s1[ci], s2[pj], s2[cj])];
            H = tb.bulge[ls].x + st.x;
            S = tb.bulge[ls].y + st.y;
            if (H > 0 || S > 0) {
                H = PO_INF;
                S = -1.0;
            }
            H += pv.x;
            S += pv.y;
            if (!fin(H)) {
                H = PO_INF;
                S = -1.0;
            }
            return make_double2(H, S);
        }
        H = tb.bulge[ls].x + at_h(s1[pi], s2[pj]) + at_h(s1[ci], s2[cj]);
        H += pv.x;
        S = tb.bulge[ls].y + at_s(s1[pi], s2[pj]) + at_s(s1[ci], s2[cj]);
        S += pv.y;
    } else if (l1 == 1 && l2 == 1) {
        const double2 a = tb.stackint2[quad(s1[pi], s1[pi + 1], s2[pj], s2[pj + 1])];
        const double2 b = tb.stackint2[quad(s2[cj], s2[cj - 1], s1[ci], s1[ci - 1])];
        S = a.y + b.y;
        S += pv.y;
        H = a.x + b.x;
        H += pv.x;
    } else {
        const double2 a = tb.tstack[quad(s1[pi], s1[pi + 1], s2[pj], s2[pj + 1])];
        const double2 b = tb.tstack[quad(s2[cj], s2[cj - 1], s1[ci], s1[ci - 1])];
        const int asym = abs(l1 - l2);
        H = tb.interior[ls].x + a.x + b.x + (ILAH_D * asym);
        H += pv.x;
        S = tb.interior[ls].y + a.y + b.y + (ILAS_D * asym);
        S += pv.y;
    }
    if (!fin(H)) {
        H = PO_INF;
        S = -1.0;
    }
    if (H > 0 && S > 0) {
        H = PO_INF;
        S = -1.0;
    }
    return make_double2(H, S);
}

// Forward scan over the candidate predecessors of cell (ci,cj): every lane
// keeps its best (largest T, earliest key).  `cnt` accumulates the number of
// candidates evaluated (the DP relaxation count).
template <int CAP>
__device__ __forceinline__ void dimer_scan(const SmemTables& tb, const WarpWs<CAP>& ws, int lane, int ci, int cj,
                                           int k_lo, int k_hi, int max_loop, double iH, double iS, double RC,
                                           int mode, double* bT, int* bKey, double2* bV, int* cnt) {
    double bestT = -PO_INF;
    int bestKey = KEY_NONE;
    double2 bestV = make_double2(PO_INF, -1.0);
    for (int kp = k_lo + lane; kp < k_hi; kp += 32) {
        const int c = ws.coord[kp];
        const int pi = c >> 8, pj = c & 255;
        if (pj >= cj) continue;
        const int l1 = ci - pi - 1, l2 = cj - pj - 1, sum = l1 + l2;
        if (sum < 1 || sum > max_loop) continue;
        const int key = KEY_OF(sum, l1);
        if (mode == SCAN_NO_STAR && key == KEY_STAR) continue;
        if (mode == SCAN_BEFORE_STAR && key >= KEY_STAR) continue;
        if (cnt) ++*cnt;
        const double2 v = dimer_loop_term(tb, ws.s1, ws.s2, pi, pj, ci, cj, l1, l2, ws.cell[kp]);
        const double T1 = (v.x + iH) / ((v.y + iS) + RC);
        if (T1 > bestT || (T1 == bestT && key < bestKey)) {
            bestT = T1;
            bestKey = key;
            bestV = v;
        }
    }
    *bT = bestT;
    *bKey = bestKey;
    *bV = bestV;
}

template <int CAP, bool COUNT>
__device__ void thal_dimer_warp(const SmemTables& tb, WarpWs<CAP>& ws, const uint4 oa, const uint4 ob,
                                const PoConsts& K, po_thal_out* out, unsigned long long* relax) {
    const int lane = threadIdx.x & 31;
    const int len1 = po_len(oa), len2 = po_len(ob);
    po_thal_out r;
    r.tm = 0.0;
    r.dg = r.dh = r.ds = 0.0;
    r.flags = 0;
    r.align_end_1 = r.align_end_2 = -1;
    if (po_bad(oa) || po_bad(ob) || len1 == 0 || len2 == 0 || len1 > PO_MAX_OLIGO_LEN || len2 > PO_MAX_OLIGO_LEN) {
        r.flags = PO_ITEM_SKIPPED;
        if (lane == 0) *out = r;
        return;
    }
    __syncwarp();
    // ---- decode; the second oligo is reversed so it runs 3'->5' ------------
    for (int p = lane; p < len1; p += 32) ws.s1[p + 1] = (uint8_t)po_base(oa, p);
    for (int p = lane; p < len2; p += 32) ws.s2[len2 - p] = (uint8_t)po_base(ob, p);
    if (lane == 0) {
        ws.s1[0] = ws.s1[len1 + 1] = 4;
        ws.s2[0] = ws.s2[len2 + 1] = 4;
    }
    // upstream symmetry_thermo() on both oligos
    bool asym = false;
    for (int p = lane; p < len1 / 2; p += 32) asym |= (po_base(oa, p) + po_base(oa, len1 - 1 - p) != 3);
    for (int p = lane; p < len2 / 2; p += 32) asym |= (po_base(ob, p) + po_base(ob, len2 - 1 - p) != 3);
    const bool sym = !__any_sync(FULL, asym) && (len1 % 2 == 0) && (len2 % 2 == 0);
    const double RC = sym ? K.rc_sym : K.rc_nonsym;
    const double iH = 200.0, iS = -5.7;
    const int max_loop = K.max_loop;
    __syncwarp();

    // ---- finite cells: bit (j-1) of cm[b] set iff s2[j] == b ----------------
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const unsigned lo = __ballot_sync(FULL, lane + 1 <= len2 && ws.s2[lane + 1] == b);
        const unsigned hi = __ballot_sync(FULL, lane + 33 <= len2 && ws.s2[lane + 33] == b);
        if (lane == 0) ws.mask[b] = (unsigned long long)lo | ((unsigned long long)hi << 32);
    }
    __syncwarp();
#define ROWMASK(i) (ws.mask[3 - ws.s1[i]])
    // row starts (exclusive prefix of per-row popcounts), two rows per lane
    {
        const int i0 = lane + 1, i1 = lane + 33;
        const int c0 = i0 <= len1 ? __popcll(ROWMASK(i0)) : 0;
        const int c1 = i1 <= len1 ? __popcll(ROWMASK(i1)) : 0;
        int x0 = c0, x1 = c1;
#pragma unroll
        for (int m = 1; m < 32; m <<= 1) {
            const int y0 = __shfl_up_sync(FULL, x0, m), y1 = __shfl_up_sync(FULL, x1, m);
            if (lane >= m) {
                x0 += y0;
                x1 += y1;
            }
        }
        const int tot0 = __shfl_sync(FULL, x0, 31);
        ws.start[i0] = (uint16_t)(x0 - c0);
        if (i1 < 64) ws.start[i1] = (uint16_t)(tot0 + x1 - c1);
    }
    __syncwarp();
    const int ncell = ws.start[len1 + 1];
    for (int i = 1; i <= len1; ++i) {
        const unsigned long long m = ROWMASK(i);
        const int base = ws.start[i];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int bit = lane + 32 * h;
            if ((m >> bit) & 1ull) ws.coord[base + __popcll(m & ((1ull << bit) - 1ull))] = (uint16_t)((i << 8) | (bit + 1));
        }
    }
    __syncwarp();

    int cnt = 0;
    // ---- fillMatrix ----------------------------------------------------------
    for (int i = 1; i <= len1; ++i) {
        const int rs = ws.start[i], re = ws.start[i + 1];
        if (rs == re) continue;
        // phase A: LSH seed and maxTM (stack) for every finite cell of the row
        for (int k = rs + lane; k < re; k += 32) {
            const int j = ws.coord[k] & 255;
            double2 v = lsh(tb, ws.s1, ws.s2, i, j, iH, iS, RC);
            if (COUNT) ++cnt;
            if (i > 1 && j > 1) {
                if (COUNT) ++cnt;
                double S0 = v.y, H0 = v.x;
                const double2 sh = rsh(tb, ws.s1, ws.s2, i, j, iH, iS, RC);
                const double T0 = (H0 + iH + sh.x) / (S0 + iS + RC + sh.y);
                double S1 = -1.0, H1 = PO_INF, T1;
                const unsigned long long pm = ROWMASK(i - 1);
                bool have = false;
                if ((pm >> (j - 2)) & 1ull) {
                    const double2 pv = ws.cell[ws.start[i - 1] + __popcll(pm & ((1ull << (j - 2)) - 1ull))];
                    const double2 st = tb.stack[quad(ws.s1[i - 1], ws.s1[i], ws.s2[j - 1], ws.s2[j])];
                    if (fin(pv.x) && fin(st.x)) {
                        S1 = pv.y + st.y;
                        H1 = pv.x + st.x;
                        have = true;
                    }
                }
                if (have) T1 = (H1 + iH + sh.x) / (S1 + iS + RC + sh.y);
                else T1 = (H1 + iH) / (S1 + iS + RC);
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
            ws.cell[k] = v;
        }
        __syncwarp();
        if (i == 1) continue;
        // phase B: bulges and interior loops, one cell at a time, lanes over predecessors
        const int k_lo = ws.start[max(1, i - 1 - max_loop)];
        for (int k = rs; k < re; ++k) {
            const int j = ws.coord[k] & 255;
            if (j == 1) continue;
            const double2 cur = ws.cell[k];
            const double Tcur = (cur.x + iH) / ((cur.y + iS) + RC);
            double bT, wT;
            int bKey, wKey;
            double2 bV;
            dimer_scan<CAP>(tb, ws, lane, i, j, k_lo, rs, max_loop, iH, iS, RC, SCAN_ALL, &bT, &bKey, &bV,
                            COUNT ? &cnt : nullptr);
            int src = warp_argmax(bT, bKey, &wT, &wKey);
            if (src >= 0 && wKey == KEY_STAR) {
                // The 1x1 loop is the arg-max.  Upstream accepts it only if it beats the
                // running best by SMALL_NON_ZERO at the moment it is visited; re-derive that
                // running best (current cell + the three candidates visited earlier).
                double pT, qT;
                int pKey, qKey;
                double2 pV;
                dimer_scan<CAP>(tb, ws, lane, i, j, k_lo, rs, max_loop, iH, iS, RC, SCAN_BEFORE_STAR, &pT, &pKey, &pV,
                                nullptr);
                warp_argmax(pT, pKey, &qT, &qKey);
                const double Tm = qT > Tcur ? qT : Tcur;
                if (!((wT - Tm) >= PO_SMALL_NON_ZERO && wT > Tm)) {
                    dimer_scan<CAP>(tb, ws, lane, i, j, k_lo, rs, max_loop, iH, iS, RC, SCAN_NO_STAR, &bT, &bKey, &bV,
                                    nullptr);
                    src = warp_argmax(bT, bKey, &wT, &wKey);
                }
            }
            if (src >= 0 && wT > Tcur) {
                double H = shfl_d(bV.x, src), S = shfl_d(bV.y, src);
                if (S < PO_MIN_ENTROPY_CUTOFF) {
                    S = PO_MIN_ENTROPY;
                    H = 0.0;
                }
                if (fin(H) && lane == 0) ws.cell[k] = make_double2(H, S);
            }
            __syncwarp();
        }
    }

    // ---- end scan: minimum dG over all cells, first in row-major order wins ---
    double bestG = PO_INF;
    int bestK = KEY_NONE;
    for (int k = lane; k < ncell; k += 32) {
        const int c = ws.coord[k];
        const int i = c >> 8, j = c & 255;
        const double2 v = ws.cell[k];
        double2 sh = rsh(tb, ws.s1, ws.s2, i, j, iH, iS, RC);
        sh.y = sh.y + PO_SMALL_NON_ZERO;
        sh.x = sh.x + PO_SMALL_NON_ZERO;
        const double G1 = (v.x + sh.x + iH) - PO_TEMP_KELVIN * (v.y + sh.y + iS);
        if (G1 < bestG) {
            bestG = G1;
            bestK = k;
        }
    }
    if (COUNT) cnt += (lane == 0) ? len1 * len2 : 0;
    double wG;
    int wK;
    warp_argmin(bestG, bestK, &wG, &wK);
    if (COUNT && relax) {
        unsigned long long c64 = (unsigned long long)cnt;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) c64 += __shfl_xor_sync(FULL, c64, m);
        if (lane == 0) atomicAdd(relax, c64);
    }
    if (wK == KEY_NONE || !fin(wG)) {  // no finite cell: no structure, tm = 0
        if (lane == 0) *out = r;
        return;
    }
    int i = ws.coord[wK] >> 8, j = ws.coord[wK] & 255;
    const int bestI = i, bestJ = j;
    double dH, dS;
    {
        const double2 v = ws.cell[wK];
        const double2 sh = rsh(tb, ws.s1, ws.s2, i, j, iH, iS, RC);
        dH = v.x + sh.x + iH;
        dS = v.y + sh.y + iS;
    }

    // ---- traceback: count the base pairs of the chosen structure ---------------
    int npairs = 1;
    for (;;) {
        const unsigned long long rm = ROWMASK(i);
        const int k = ws.start[i] + __popcll(rm & ((1ull << (j - 1)) - 1ull));
        const double2 cur = ws.cell[k];
        const double2 l = lsh(tb, ws.s1, ws.s2, i, j, iH, iS, RC);
        if (eq5(cur.y, l.y) && eq5(cur.x, l.x)) break;
        if (i > 1 && j > 1) {
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
        const double Tcur = (cur.x + iH) / ((cur.y + iS) + RC);
        int myKey = KEY_NONE;
        const int k_lo = ws.start[max(1, i - 1 - max_loop)], k_hi = ws.start[i];
        for (int kp = k_lo + lane; kp < k_hi; kp += 32) {
            const int c = ws.coord[kp];
            const int pi = c >> 8, pj = c & 255;
            if (pj >= j) continue;
            const int l1 = i - pi - 1, l2 = j - pj - 1, sum = l1 + l2;
            if (sum < 1 || sum > max_loop) continue;
            const int key = KEY_OF(sum, l1);
            if (key >= myKey) continue;
            const double2 v = dimer_loop_term(tb, ws.s1, ws.s2, pi, pj, i, j, l1, l2, ws.cell[kp]);
            if (key == KEY_STAR) {  // traceback mode of the 1x1 branch: only when T1 >= T2
                const double T1 = (v.x + iH) / ((v.y + iS) + RC);
                if (!(T1 >= Tcur)) continue;
            }
            if (eq5(cur.y, v.y) && eq5(cur.x, v.x)) myKey = key;
        }
        int wKey = myKey;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) wKey = min(wKey, __shfl_xor_sync(FULL, wKey, m));
        if (wKey == KEY_NONE) break;  // guard: upstream would spin here (unreachable for <= 60 nt)
        const int sum = wKey >> 6, l1 = wKey & 63;
        i = i - l1 - 1;
        j = j - (sum - l1) - 1;
        ++npairs;
    }
#undef ROWMASK

    // ---- drawDimer ---------------------------------------------------------------
    const int N = npairs - 1;  // (2 * npairs) / 2 - 1
    const double t = (dH / (dS + (N * K.salt_corr) + RC)) - PO_ABSOLUTE_ZERO;
    const double G = dH - (K.temp_k * (dS + (N * K.salt_corr)));
    const double S = dS + (N * K.salt_corr);
    r.tm = t;
    r.dg = G;
    r.dh = dH;
    r.ds = S;
    r.flags = PO_ITEM_STRUCTURE_FOUND | (sym ? PO_ITEM_SYMMETRIC : 0);
    r.align_end_1 = (int16_t)bestI;
    r.align_end_2 = (int16_t)bestJ;
    if (lane == 0) *out = r;
}
