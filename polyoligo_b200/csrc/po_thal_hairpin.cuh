// po_thal_hairpin.cuh -- thal type HAIRPIN (calcHairpin) for one warp.
// Restates upstream thal.c: initMatrix2, fillMatrix2 (maxTM2, CBI /
// calc_bulge_internal2, calc_hairpin), calc_terminal_bp (END5_1..4),
// tracebacku and drawHairpin.
//
// Finite cells (i,j): j - i >= 4 and bases complementary.  They are kept in
// upstream's fill order (j ascending, i descending), so the cells enclosed by
// (i,j) are a prefix of the list.  The cells of one column do not depend on
// each other, so the warp relaxes 32/G of them side by side, one group of G
// lanes each (Grp<G>); control flow stays warp-uniform (uniform trip counts,
// clamped indices, predicated stores).
#pragma once
#include "po_thal.cuh"

// rows i <= j-4 of column j that pair with base j
#define HP_COLMASK(ws, j) ((j) >= 5 ? ((ws).mask[3 - (ws).s1[j]] & ((1ull << ((j)-4)) - 1ull)) : 0ull)

struct HairpinCellCtx {
    int i, j;
    int b1q;  // bi * 64 + bj * 4: closing-pair part of the bulge-of-one stack index
};

// upstream calc_bulge_internal2(i, j, ii, jj) WITHOUT the enclosed cell's own
// value (its traceback == 1 form); the forward pass adds that value afterwards
// in the same association order.  In upstream's sums the closing pair's term
// comes first, the enclosed pair's second.
template <class WS>
__device__ __forceinline__ double2 hairpin_loop_term(const TabAddr& A, const WS& ws, const double2* ycell,
                                                     const HairpinCellCtx& c, int kp, int* key_out) {
    const uint32_t code = ws.code[kp];
    const int b = CODE_B(code);
    const int l1 = CODE_I(code) - c.i - 1, l2 = c.j - CODE_J(code) - 1;
    // stack[s[i]][s[ii]][s[j]][s[jj]] = bi*64 + b*16 + bj*4 + (3-b)
    const LoopOps o = loop_ops(A, ycell, l1, l2, CODE_TQ(code), b, c.b1q + b * 15 + 3);
    *key_out = o.key;
    // closing side first, then the enclosed side.  For a bulge of one upstream sums bulge + stack
    // only; there C is the {+0, +0} placeholder and L + (+0) is L exactly, so the same expression
    // gives the same bits.
    const double H = (o.L.x + o.C.x) + o.P.x;
    const double S = ((o.L.y + o.C.y) + o.P.y) + o.asymS;
    return make_double2(H, S);
}

__device__ __forceinline__ double2 hp_fix(double H, double S) {
    if (!fin(H)) return make_double2(PO_INF, -1.0);
    return make_double2(H, S);
}

// binary search of a loop-bonus list (global memory, sorted keys)
__device__ __forceinline__ double2 bonus_lookup(const int32_t* keys, const double2* vals, int n, int key) {
    int lo = 0, hi = n - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const int k = keys[mid];
        if (k == key) return vals[mid];
        if (k < key) lo = mid + 1;
        else hi = mid - 1;
    }
    return make_double2(0.0, 0.0);
}

// The sequence-only part of upstream calc_hairpin(): loop entropy / enthalpy, terminal mismatch or AT
// penalty, tri- / tetraloop bonus of the hairpin loop closed by (i,j).
__device__ __noinline__ double2 hairpin_closure_static(const SmemTables& tb, const SmemExtra& X, const PoTables* gt,
                                                          const uint8_t* s, int i, int j) {
    const int loopSize = j - i - 1;
    double H, S;
    const int li = loopSize <= 30 ? loopSize - 1 : 29;
    H = tb.hairpin[li].x;
    S = tb.hairpin[li].y;
    const double2 t = tb.tstack2[quad(s[i], s[i + 1] & 3, s[j], s[j - 1] & 3)];
    if (loopSize > 3) {
        H += t.x;
        S += t.y;
    } else if (loopSize == 3) {
        H += at_h(s[i], s[j]);
        S += at_s(s[i], s[j]);
    }
    if (loopSize == 3) {
        if (X.n_tri) {
            int key = 0;
#pragma unroll
            for (int c = 0; c < 5; ++c) key |= (s[i + c] & 3) << (2 * c);
            double2 b = make_double2(0.0, 0.0);
            if (X.n_tri <= 16) {  // shared-memory copy, sorted: 4-step branch-free search
                int pos = 0;
#pragma unroll
                for (int st = 8; st >= 1; st >>= 1) pos += (X.tri_key[pos + st - 1] < key) ? st : 0;
                pos = min(pos, 15);
                if (X.tri_key[pos] == key) b = X.tri_val[pos];
            } else {
                b = bonus_lookup(gt->tri_key, gt->tri_val, gt->n_tri, key);
            }
            H += b.x;
            S += b.y;
        }
    } else if (loopSize == 4) {
        if (X.n_tetra) {
            int key = 0;
#pragma unroll
            for (int c = 0; c < 6; ++c) key |= (s[i + c] & 3) << (2 * c);
            const double2 b = bonus_lookup(gt->tetra_key, gt->tetra_val, gt->n_tetra, key);
            H += b.x;
            S += b.y;
        }
    }
    if (!fin(H)) {
        H = PO_INF;
        S = -1.0;
    }
    return make_double2(H, S);
}

// the part of calc_hairpin() that looks at the cell's current value
__device__ __forceinline__ double2 hairpin_closure_vs(double2 c, double2 cur) {
    if (c.x > 0 && c.y > 0 && (!(cur.x > 0) || !(cur.y > 0))) c = make_double2(PO_INF, -1.0);
    return c;
}

// upstream calc_hairpin() without its final dG comparison: the hairpin-loop
// closure value of (i,j); `cur` is the current DP value of the cell.
__device__ __forceinline__ double2 hairpin_closure(const SmemTables& tb, const SmemExtra& X, const PoTables* gt,
                                                   const uint8_t* s, int i, int j, double2 cur) {
    return hairpin_closure_vs(hairpin_closure_static(tb, X, gt, s, i, j), cur);
}

// index of finite cell (i,j) in the list; caller guarantees it is finite
template <class WS>
__device__ __forceinline__ int hp_index(const WS& ws, int i, int j) {
    return ws.start[j] + __popcll(HP_COLMASK(ws, j) >> i);
}
template <class WS>
__device__ __forceinline__ bool hp_finite(const WS& ws, int i, int j) {
    if (j - i < PO_MIN_HRPN_LOOP + 1 || i < 1) return false;
    return ((ws.mask[3 - ws.s1[j]] >> (i - 1)) & 1ull) != 0;
}

// Candidates enclosed by (i,j): every finite (ii,jj) with i < ii, jj < j and
// 1 <= (ii-i-1)+(j-jj-1) <= max_loop.  Virtual run l2 owns column j-1-l2; the
// column's candidates are one contiguous range of the list (rows descending).
// Fills the group's run table and closing-pair loop operands (ycell, stored
// identically by every lane of the group) and returns the number of candidates.
template <int G, class WS>
__device__ __forceinline__ int hairpin_build_runs(const Grp<G>& g, const SmemTables& tb, WS& ws, int i,
                                                  int j, int max_loop, bool prune, HairpinCellCtx* ctx) {
    const int bi = ws.s1[i], bj = ws.s1[j];
    const int q = quad(bi, ws.s1[i + 1] & 3, bj, ws.s1[j - 1] & 3);
    const double2 y0 = tb.tstack[q], y3 = tb.stackint2[q];
    // both closing-pair terms infinite: only bulges (l2 == 0 or l1 == 0) can be finite
    const bool interior_ok = fin(y0.x) | fin(y3.x);
    RunTable& rt = ws.runs[g.idx];
    int carry = 0;
#pragma unroll
    for (int sl = 0; sl < 32 / G; ++sl) {
        if (sl * G > j - 6) break;  // columns j-1 .. 5 only: no run that far (warp-uniform: one column per batch)
        const int l2 = sl * G + g.lane, jj = j - 1 - l2;
        const bool have = (jj >= 5) & (l2 <= max_loop);
        const int jc = have ? jj : 5;
        const int lo_row = (l2 == 0) ? i + 2 : i + 1;            // l1 + l2 >= 1
        int hi_row = min(jc - 4, i + 1 + (max_loop - l2));        // l1 + l2 <= max_loop
        if (prune && !interior_ok && l2 > 0) hi_row = min(hi_row, i + 1);
        const unsigned long long cm = HP_COLMASK(ws, jc);
        // rows lo_row .. hi_row (1-based) are bits lo_row-1 .. hi_row-1; lo_row >= 2
        const int cnt = have ? max(popc_below(cm, min(max(hi_row, 0), 63)) - popc_below(cm, lo_row - 1), 0) : 0;
        const int first = ws.start[jc] + __popcll(cm >> min(max(hi_row, 0), 63));
        const int incl = g.scan_incl(cnt) + carry;
        rt.end[l2] = (uint16_t)incl;
        rt.adj[l2] = (int16_t)(first - (incl - cnt));
        carry = g.shfl(incl, G - 1);
    }
    const bool cAT = (bi == 0) | (bi == 3);
    double2* yc = ws.ycell[g.idx];
    yc[0] = y0;
    yc[1] = make_double2(cAT ? PO_AT_H : 0.0, cAT ? PO_AT_S : 0.0);
    yc[2] = make_double2(0.0, 0.0);
    yc[3] = y3;
    ctx->i = i;
    ctx->j = j;
    ctx->b1q = bi * 64 + bj * 4;
    g.sync();
    return carry;
}

// Scan of the group's candidates; `trip_total` is the warp-wide maximum of the
// groups' candidate counts (uniform trip count).  SCAN_LAST_GE selects the LAST
// candidate (largest key) with T1 >= Tcur, which is what upstream's
// CBI(traceback = 2) leaves in its output argument.
template <int G, class WS>
__device__ __forceinline__ void hairpin_scan(const Grp<G>& g, const TabAddr& A, const WS& ws,
                                             const HairpinCellCtx& c, int total, int trip_total, double iH, double iS,
                                             double RC, int mode, double Tcur, double* bT, int* bKey, double2* bV) {
    const RunTable& rt = ws.runs[g.idx];
    const double2* yc = ws.ycell[g.idx];
    double bestT = -PO_INF;
    int bestKey = KEY_NONE;
    double2 bestV = make_double2(PO_INF, -1.0);
    for (int t0 = 0; t0 < trip_total; t0 += G) {
        const int t = t0 + g.lane;
        bool act = t < total;
        const int kp = run_lookup(rt, act ? t : 0, G < 32 && c.j - 6 < 16);
        int key;
        double2 v = hairpin_loop_term(A, ws, yc, c, kp, &key);
        if (mode == SCAN_NO_STAR) act &= key != KEY_STAR;
        if (mode == SCAN_BEFORE_STAR) act &= key < KEY_STAR;
        const double2 pv = ws.cell[kp];
        v.x += pv.x;
        v.y += pv.y;
        v = hp_fix(v.x, v.y);
        const double T1 = po_div(v.x + iH, (v.y + iS) + RC);
        bool take;
        if (mode == SCAN_LAST_GE) take = act & (T1 >= Tcur) & ((bestKey == KEY_NONE) | (key > bestKey));
        else take = act & ((T1 > bestT) | ((T1 == bestT) & (key < bestKey)));
        if (take) {
            bestT = T1;
            bestKey = key;
            bestV = v;
        }
    }
    *bT = bestT;
    *bKey = bestKey;
    *bV = bestV;
}

// upstream END5_1..END5_4 for FOUR consecutive positions i0 .. i0+3 at once: lanes 8b .. 8b+7 serve position
// i0 + b, inside them lanes 2(v-1), 2(v-1)+1 serve variant v.  A candidate of position i reads the
// exterior prefix values HEND5 / SEND5 of k <= i - 5 only (the paired base p is at least four below its
// partner, k = p - 1 or p - 2), so the four positions of a block depend on nothing inside the block
// and their candidates are evaluated side by side; the caller then runs upstream's max5() selection
// position by position.  Returns the variant's {H_max, S_max} in both lanes of its pair.
template <class WS>
__device__ __noinline__ double2 end5_block(const SmemTables& tb, const WS& ws, int lane, int i0, int len,
                                              double iH, double iS, double RC, double T2) {
    const int i = i0 + (lane >> 3), v = ((lane >> 1) & 3) + 1, sub = lane & 1;
    const uint8_t* s = ws.s1;
    const bool in = i <= len;
    const int q = (v <= 2) ? i : i - 1;  // paired 3' base
    double bestT = -PO_INF;
    int bestKey = KEY_NONE;
    double2 best = make_double2(PO_INF, -1.0);
    const int ks = (in && q >= 5) ? ws.start[q] : 0, ke = (in && q >= 5) ? ws.start[q + 1] : 0;
    const int trips = warp_max_int((ke - ks + 1) >> 1);
    const int ic = in ? i : 2;   // positions past the oligo only keep the lanes company
    const int bi = s[ic] & 3, bim = s[ic - 1] & 3;
    for (int it = 0; it < trips; ++it) {
        const int kc0 = ks + sub + 2 * it;
        bool act = kc0 < ke;
        const int kc = act ? kc0 : 0;
        const int p = CODE_I(ws.code[kc]);
        const int k = (v == 1 || v == 3) ? p - 1 : p - 2;  // the exterior prefix ends at k
        act &= k >= 0;
        const int kk = act ? k : 0;
        const int bp = s[kk + 1] & 3, bp2 = s[kk + 2] & 3;
        double2 x = make_double2(0.0, 0.0);
        if (v == 2) x = tb.dangle5[(bi * 4 + bp2) * 4 + bp];            // Hd5(i, k+2)
        else if (v == 3) x = tb.dangle3[(bim * 4 + bi) * 4 + bp];       // Hd3(i-1, k+1)
        else if (v == 4) x = tb.tstack2[quad(bim, bi, bp2, bp)];        // Htstack(i-1, k+2)
        const double2 dv = ws.cell[kc];
        const double hk = ws.end5[0][kk], sk = ws.end5[1][kk];
        const double T1p = ws.end5t[kk];   // == po_div(hk + iH, sk + iS + RC), kept when end5[kk] was set
        double H, S;
        const int pp = act ? p : 1, qq = act ? q : 1;
        const double aH = at_h(s[pp], s[qq]), aS = at_s(s[pp], s[qq]);
        if (T1p >= T2) {
            if (v == 1) {
                H = hk + aH + dv.x;
                S = sk + aS + dv.y;
            } else {
                H = hk + aH + x.x + dv.x;
                S = sk + aS + x.y + dv.y;
            }
        } else {
            if (v == 1) {
                H = 0 + aH + dv.x;
                S = 0 + aS + dv.y;
            } else {
                H = 0 + aH + x.x + dv.x;
                S = 0 + aS + x.y + dv.y;
            }
        }
        if (!fin(H) || H > 0 || S > 0) {
            H = PO_INF;
            S = -1.0;
        }
        const double T1 = po_div_rej(H + iH, S + iS + RC);  // H <= 0 and S <= 0 here, or {inf, -1}
        // upstream keeps the first k with the strictly largest T1; -inf never wins
        if (act & (S > PO_MIN_ENTROPY_CUTOFF) & (T1 > -PO_INF) & ((T1 > bestT) | ((T1 == bestT) & (k < bestKey)))) {
            bestT = T1;
            bestKey = k;
            best = make_double2(H, S);
        }
    }
    // arg-max inside each pair of lanes
    {
        const double ot = __shfl_xor_sync(FULL, bestT, 1);
        const int ok = __shfl_xor_sync(FULL, bestKey, 1);
        const double oh = __shfl_xor_sync(FULL, best.x, 1), os = __shfl_xor_sync(FULL, best.y, 1);
        if ((ot > bestT) | ((ot == bestT) & (ok < bestKey))) best = make_double2(oh, os);
    }
    return best;
}

// calc_terminal_bp (END5_1..4), tracebacku and drawHairpin on a filled table; shared by every
// hairpin tier (WS: HairpinWs<CAP> or Hairpin32Ws<CAP>).  `cnt` carries the relaxation count of the
// fill (COUNT variant).
template <class WS, bool COUNT>
__device__ __noinline__ void hairpin_finish(const SmemTables& tb, const SmemExtra& X, const TabAddr& A,
                                               const PoTables* gt, WS& ws, const int len, const PoConsts& K,
                                               po_thal_out r, po_thal_out* out, unsigned long long* relax, int cnt) {
    const int lane = threadIdx.x & 31;
    const Grp<32> gw = Grp<32>::make();
    const double iH = 0.0, iS = -0.00000000001, RC = 0.0;
    const int max_loop = K.max_loop;
    // ---- calc_terminal_bp --------------------------------------------------------
    const double T2e = po_div(0 + iH, 0 + iS + RC);  // quotient of the empty 5' prefix, END5_1..4's T2
    if (lane == 0) {
        ws.end5[1][0] = ws.end5[1][1] = -1.0;
        ws.end5[0][0] = ws.end5[0][1] = PO_INF;
        ws.end5t[0] = ws.end5t[1] = po_div(PO_INF + iH, -1.0 + iS + RC);
    }
    __syncwarp();
    {
        double hn = ws.end5[0][1], sn = ws.end5[1][1];
        double tn = ws.end5t[1];  // T of (hn, sn): the previous quotient, or the chosen variant's
        for (int i0 = 2; i0 <= len; i0 += 4) {
            // no finite cell in columns i0-1 .. i0+3: every END5 variant of the block is {inf,-1}
            const int cfirst = max(i0 - 1, 5), clast = min(i0 + 3, len);
            const bool any_blk = clast >= cfirst && ws.start[clast + 1] > ws.start[cfirst];
            // Per position (lanes 8b .. 8b+7 hold position i0 + b): the variant upstream's max5() would take
            // if the carried value does not beat all four, its values, and whether it is accepted (dG < 0).
            // None of this depends on the carried value, so the four positions are prepared side by side;
            // the carry (T1 = the previous position's quotient) is then resolved position by position.
            double Tmax4 = -PO_INF, et = -PO_INF, eh = PO_INF, es = -1.0;
            bool take = false;
            if (any_blk) {
                const double2 e = end5_block(tb, ws, lane, i0, len, iH, iS, RC, T2e);
                const double Tv = po_div_rej(e.x + iH, e.y + iS + RC);  // {inf, -1}, or dH <= 0 and dS <= 0
                const double T2 = __shfl_sync(FULL, Tv, 0, 8), T3 = __shfl_sync(FULL, Tv, 2, 8),
                             T4 = __shfl_sync(FULL, Tv, 4, 8), T5 = __shfl_sync(FULL, Tv, 6, 8);
                int mx;
                if (T2 > T3 && T2 > T4 && T2 > T5) mx = 2;
                else if (T3 > T4 && T3 > T5) mx = 3;
                else if (T4 > T5) mx = 4;
                else mx = 5;
                eh = __shfl_sync(FULL, e.x, (mx - 2) * 2, 8);
                es = __shfl_sync(FULL, e.y, (mx - 2) * 2, 8);
                et = mx == 2 ? T2 : mx == 3 ? T3 : mx == 4 ? T4 : T5;
                Tmax4 = fmax(fmax(T2, T3), fmax(T4, T5));   // quotients or -inf, never NaN
                const int i = i0 + (lane >> 3);
                // no finite cell in columns i and i-1: every END5 variant is {inf,-1}, T = -inf, and
                // upstream's max5() falls through to its last case, which keeps the previous value
                const bool any = i <= len && ((i >= 5 && ws.start[i + 1] > ws.start[i]) ||
                                              (i >= 6 && ws.start[i] > ws.start[i - 1]));
                take = any && (eh - (K.temp_k * es)) < 0.0;
            }
#pragma unroll 1
            for (int b = 0; b < 4; ++b) {
                const int i = i0 + b;
                if (i > len) break;
                if (any_blk) {
                    // lanes of position b decide with the carried quotient; everybody receives the result
                    const bool d = take && !(tn > Tmax4);
                    const double nh = d ? eh : hn, ns = d ? es : sn, nt = d ? et : tn;
                    hn = gw.shfl(nh, 8 * b);
                    sn = gw.shfl(ns, 8 * b);
                    tn = gw.shfl(nt, 8 * b);
                }
                if (COUNT) cnt += (lane == 0) ? 4 : 0;
                if (lane == 0) {
                    ws.end5[0][i] = hn;
                    ws.end5[1][i] = sn;
                    ws.end5t[i] = tn;
                }
            }
            __syncwarp();
        }
    }
    if (COUNT && relax) {
        __syncwarp();
        const unsigned total_cnt = __reduce_add_sync(FULL, (unsigned)cnt);
        if (lane == 0) atomicAdd(relax, (unsigned long long)total_cnt);
    }
    double mh = ws.end5[0][len], ms = ws.end5[1][len];
    if (!fin(mh) || !fin(ms)) {
        if (lane == 0) *out = r;
        return;
    }

    // ---- tracebacku: collect the paired positions ---------------------------------
    unsigned long long paired = 0ull;
    int sp_ = 0;
    if (lane == 0) ws.stk[0] = (1 << 16) | (len << 8);
    sp_ = 1;
    __syncwarp();
    int guard = 0;
    while (sp_ > 0 && guard++ < 256) {
        --sp_;
        const int item = ws.stk[sp_];
        __syncwarp();
        int i = (item >> 8) & 255, j = item & 255;
        if (item >> 16) {
            // exterior loop: find the structure that ends at or before i
            while (i > 0 && eq5(ws.end5[1][i], ws.end5[1][i - 1]) && eq5(ws.end5[0][i], ws.end5[0][i - 1])) --i;
            if (i == 0) continue;
            // the block form evaluated for position i alone: its four variants sit in lanes 0, 2, 4, 6
            const double2 e = end5_block(tb, ws, lane, i, len, iH, iS, RC, T2e);
            const double hi_ = ws.end5[0][i], si_ = ws.end5[1][i];
            const unsigned match = __ballot_sync(FULL, lane < 8 && (lane & 1) == 0 && eq5(si_, e.y) && eq5(hi_, e.x));
            if (!match) continue;
            const int v = (__ffs(match) - 1) / 2 + 1;  // upstream's else-if chain: first matching variant only
            const uint8_t* s = ws.s1;
            const int q = (v <= 2) ? i : i - 1;
            int myKey = KEY_NONE;
            if (q >= 5) {
                const int ks = ws.start[q], ke = ws.start[q + 1];
                for (int kc0 = ks; kc0 < ke; kc0 += 32) {
                    bool act = kc0 + lane < ke;
                    const int kc = act ? kc0 + lane : ks;
                    const int p = CODE_I(ws.code[kc]);
                    const int k = (v == 1 || v == 3) ? p - 1 : p - 2;
                    act &= k >= 0;
                    const int kk = act ? k : 0;
                    const int bp = s[kk + 1] & 3, bp2 = s[kk + 2] & 3, bi = s[i] & 3, bim = s[i - 1] & 3;
                    double2 x = make_double2(0.0, 0.0);
                    if (v == 2) x = tb.dangle5[(bi * 4 + bp2) * 4 + bp];
                    else if (v == 3) x = tb.dangle3[(bim * 4 + bi) * 4 + bp];
                    else if (v == 4) x = tb.tstack2[quad(bim, bi, bp2, bp)];
                    const double2 dv = ws.cell[kc];
                    const double aH = at_h(s[p], s[q]), aS = at_s(s[p], s[q]);
                    const double hk = ws.end5[0][kk], sk = ws.end5[1][kk];
                    double s0, h0, s1v, h1v;
                    if (v == 1) {
                        s0 = aS + dv.y;
                        h0 = aH + dv.x;
                        s1v = sk + aS + dv.y;
                        h1v = hk + aH + dv.x;
                    } else {
                        s0 = aS + x.y + dv.y;
                        h0 = aH + x.x + dv.x;
                        s1v = sk + aS + x.y + dv.y;
                        h1v = hk + aH + x.x + dv.x;
                    }
                    // key: smaller k first; bit 0 = continues into the prefix structure
                    if (act && eq5(si_, s0) && eq5(hi_, h0)) myKey = min(myKey, k * 2);
                    else if (act && eq5(si_, s1v) && eq5(hi_, h1v)) myKey = min(myKey, k * 2 + 1);
                }
            }
            const int wKey = gw.min_s32(myKey);
            if (wKey == KEY_NONE) continue;
            const int k = wKey >> 1;
            const int p = (v == 1 || v == 3) ? k + 1 : k + 2;
            if (lane == 0) {
                // upstream pushes the pair first, then the prefix: the prefix is popped first
                ws.stk[sp_] = (p << 8) | q;
                if (wKey & 1) ws.stk[sp_ + 1] = (1 << 16) | (k << 8);
            }
            sp_ += 1 + (wKey & 1);
            __syncwarp();
        } else {
            paired |= (1ull << (i - 1)) | (1ull << (j - 1));
            const int k = hp_index(ws, i, j);
            const double2 cur = ws.cell[k];
            // stack
            if (hp_finite(ws, i + 1, j - 1)) {
                const double2 pv = ws.cell[hp_index(ws, i + 1, j - 1)];
                const double2 st = tb.stack[quad(ws.s1[i], ws.s1[i + 1], ws.s1[j], ws.s1[j - 1])];
                if (eq5(cur.y, st.y + pv.y) && eq5(cur.x, st.x + pv.x)) {
                    if (lane == 0) ws.stk[sp_] = ((i + 1) << 8) | (j - 1);
                    ++sp_;
                    __syncwarp();
                    continue;
                }
            }
            // hairpin loop closes here
            const double2 hv = hairpin_closure(tb, X, gt, ws.s1, i, j, cur);
            if (eq5(cur.y, hv.y) && eq5(cur.x, hv.x)) continue;
            // bulge / interior loop
            if (j - i < 7) continue;
            const double Tcur = po_div(cur.x + iH, (cur.y + iS) + RC);
            HairpinCellCtx c;
            const int total = hairpin_build_runs<32>(gw, tb, ws, i, j, max_loop, true, &c);
            if (total == 0) continue;
            double bT;
            int bKey;
            double2 bV;
            hairpin_scan<32>(gw, A, ws, c, total, total, iH, iS, RC, SCAN_LAST_GE, Tcur, &bT, &bKey, &bV);
            // LAST candidate with T1 >= Tcur: largest key
            const int mxKey = gw.max_s32(bKey == KEY_NONE ? -1 : bKey);
            if (mxKey < 0) continue;
            const unsigned who = __ballot_sync(FULL, bKey == mxKey);
            const int src = __ffs(who) - 1;
            double H2 = gw.shfl(bV.x, src), S2 = gw.shfl(bV.y, src);
            if (fin(H2) && S2 < PO_MIN_ENTROPY_CUTOFF) {
                S2 = PO_MIN_ENTROPY;
                H2 = 0.0;
            }
            if (!(eq5(cur.y, S2) && eq5(cur.x, H2))) continue;
            // first candidate (upstream order) whose value reproduces the cell
            int myKey = KEY_NONE;
            for (int t0 = 0; t0 < total; t0 += 32) {
                const int t = t0 + lane;
                const bool act = t < total;
                const int kp = run_lookup(ws.runs[0], act ? t : 0, false);
                int key;
                double2 v = hairpin_loop_term(A, ws, ws.ycell[0], c, kp, &key);
                v = hp_fix(v.x, v.y);  // traceback == 1 form: every branch reports its value
                const double2 pv = ws.cell[kp];
                if (act && key < myKey && eq5(cur.y, v.y + pv.y) && eq5(cur.x, v.x + pv.x)) myKey = key;
            }
            const int wKey = gw.min_s32(myKey);
            if (wKey == KEY_NONE) continue;
            const int sum = wKey >> 6, l1 = wKey & 63;
            const int ii = i + l1 + 1, jj = j - (sum - l1) - 1;
            if (lane == 0) ws.stk[sp_] = (ii << 8) | jj;
            ++sp_;
            __syncwarp();
        }
    }

    // ---- drawHairpin: the count stops before the last base ---------------------------
    const int N = __popcll(paired & ((1ull << (len - 1)) - 1ull));
    const double t = (mh / (ms + (((N / 2) - 1) * K.salt_corr))) - PO_ABSOLUTE_ZERO;
    const double mg = mh - (K.temp_k * (ms + (((N / 2) - 1) * K.salt_corr)));
    ms = ms + (((N / 2) - 1) * K.salt_corr);
    r.tm = t;
    r.dg = mg;
    r.dh = mh;
    r.ds = ms;
    r.flags = PO_ITEM_STRUCTURE_FOUND;
    if (lane == 0) *out = r;
}

template <int G, int CAP, bool COUNT>
__device__ void thal_hairpin_warp(const SmemTables& tb, const SmemExtra& X, const TabAddr& A, const PoTables* gt,
                                  HairpinWs<CAP>& ws, const uint4 oa, const PoConsts& K, po_thal_out* out,
                                  unsigned long long* relax) {
    const int lane = threadIdx.x & 31;
    const Grp<G> g = Grp<G>::make();
    const Grp<32> gw = Grp<32>::make();
    constexpr int NG = 32 / G;
    const int len = po_len(oa);
    po_thal_out r;
    r.tm = 0.0;
    r.dg = r.dh = r.ds = 0.0;
    r.flags = 0;
    r.align_end_1 = r.align_end_2 = 0;
    if (po_bad(oa) || len == 0 || len > PO_MAX_OLIGO_LEN) {
        r.flags = PO_ITEM_SKIPPED;
        if (lane == 0) *out = r;
        return;
    }
    __syncwarp();
    for (int p = lane; p < 64; p += 32)
        if (p < len) ws.s1[p + 1] = (uint8_t)po_base(oa, p);
    if (lane == 0) ws.s1[0] = ws.s1[len + 1] = 4;
    const double iH = 0.0, iS = -0.00000000001, RC = 0.0;
    const double2* rshT = X.rsh[2];
    const int max_loop = K.max_loop;
    __syncwarp();
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const unsigned lo = __ballot_sync(FULL, lane + 1 <= len && ws.s1[lane + 1] == b);
        const unsigned hi = __ballot_sync(FULL, lane + 33 <= len && ws.s1[lane + 33] == b);
        if (lane == 0) ws.mask[b] = (unsigned long long)lo | ((unsigned long long)hi << 32);
    }
    __syncwarp();
    {
        const int j0 = lane + 1, j1 = lane + 33;
        const int c0 = j0 <= len ? __popcll(HP_COLMASK(ws, j0)) : 0;
        const int c1 = j1 <= len ? __popcll(HP_COLMASK(ws, j1)) : 0;
        const int x0 = gw.scan_incl(c0), x1 = gw.scan_incl(c1);
        const int tot0 = __shfl_sync(FULL, x0, 31);
        ws.start[j0] = (uint16_t)(x0 - c0);
        if (j1 < 64) ws.start[j1] = (uint16_t)(tot0 + x1 - c1);
    }
    if (lane == 0) ws.start[0] = 0;
    __syncwarp();
    for (int j = 5; j <= len; ++j) {
        const unsigned long long m = HP_COLMASK(ws, j);
        const int base = ws.start[j];
        const int bj = ws.s1[j], bjn = ws.s1[j + 1] & 3;
        const uint32_t colcode = ((uint32_t)j << 6);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int bit = lane + 32 * h;  // row i = bit + 1; list order is i descending
            if ((m >> bit) & 1ull) {
                const int bi = ws.s1[bit + 1];
                const int tq = quad(bj, bjn, bi, ws.s1[bit] & 3);  // [s[j]][s[j+1]][s[i]][s[i-1]]
                ws.code[base + __popcll(m >> (bit + 1))] =
                    colcode | (uint32_t)(bit + 1) | ((uint32_t)tq << 12) | ((uint32_t)bi << 20);
            }
        }
    }
    __syncwarp();

    int cnt = 0;
    // maxTM2's empty-structure quotient (0 + iH) / (MIN_ENTROPY + iS + RC): the same for every cell
    const double T0_const = po_div(0.0 + iH, PO_MIN_ENTROPY + iS + RC);
    // ---- fillMatrix2 -----------------------------------------------------------
    for (int j = 5; j <= len; ++j) {
        const int cs = ws.start[j], ce = ws.start[j + 1];
        if (cs == ce) continue;
        // phase A: maxTM2 (stack on (i+1, j-1)) for every finite cell of the column
        for (int k0 = cs; k0 < ce; k0 += 32) {
            const bool act = k0 + lane < ce;
            const int k = act ? k0 + lane : ce - 1;
            const int i = CODE_I(ws.code[k]);
            double S0 = PO_MIN_ENTROPY, H0 = 0.0;
            const double T0 = T0_const;
            const bool pfin = hp_finite(ws, i + 1, j - 1);
            double2 pv = ws.cell[pfin ? hp_index(ws, i + 1, j - 1) : 0];
            if (!pfin) pv = make_double2(PO_INF, -1.0);
            const double2 st = tb.stack[quad(ws.s1[i], ws.s1[i + 1], ws.s1[j], ws.s1[j - 1])];
            double S1 = pv.y + st.y;
            double H1 = pv.x + st.x;
            const double T1 = po_div(H1 + iH, S1 + iS + RC);
            if (S1 < PO_MIN_ENTROPY_CUTOFF) {
                S1 = PO_MIN_ENTROPY;
                H1 = 0.0;
            }
            if (S0 < PO_MIN_ENTROPY_CUTOFF) {
                S0 = PO_MIN_ENTROPY;
                H0 = 0.0;
            }
            if (act) ws.cell[k] = (T1 > T0) ? make_double2(H1, S1) : make_double2(H0, S0);
        }
        __syncwarp();
        // phase B: CBI, 32/G cells of the column at a time.  Cells with j - i < 7 have no room
        // for an enclosed pair plus a loop; rows descend along the list, so they come first.
        const int kb = j >= 8 ? cs + __popcll(HP_COLMASK(ws, j) >> (j - 7)) : ce;
        for (int k0 = kb; k0 < ce; k0 += NG) {
            const bool valid = k0 + g.idx < ce;
            const int k = valid ? k0 + g.idx : ce - 1;
            const int i = CODE_I(ws.code[k]);
            HairpinCellCtx c;
            int total = hairpin_build_runs<G>(g, tb, ws, i, j, max_loop, !COUNT, &c);
            total = valid ? total : 0;
            if (COUNT) cnt += (g.lane == 0) ? total : 0;
            const int trip_total = NG == 1 ? total : warp_max_int(total);
            if (trip_total > 0) {
                const double2 cur = ws.cell[k];
                const double Tcur = po_div(cur.x + iH, (cur.y + iS) + RC);
                double bT, wT = -PO_INF;
                int bKey, wKey = KEY_NONE;
                double2 bV;
                hairpin_scan<G>(g, A, ws, c, total, trip_total, iH, iS, RC, SCAN_ALL, Tcur, &bT, &bKey, &bV);
                int src = g.argmax(bT, bKey, &wT, &wKey);
                const bool star = src >= 0 && wKey == KEY_STAR;
                if (__any_sync(FULL, star)) {
                    double pT, qT = -PO_INF;
                    int pKey, qKey = KEY_NONE;
                    double2 pV;
                    hairpin_scan<G>(g, A, ws, c, total, trip_total, iH, iS, RC, SCAN_BEFORE_STAR, Tcur, &pT, &pKey,
                                         &pV);
                    const int psrc = g.argmax(pT, pKey, &qT, &qKey);
                    const double Tm = (psrc >= 0 && qT > Tcur) ? qT : Tcur;
                    const bool redo = star && !((wT - Tm) >= PO_SMALL_NON_ZERO && wT > Tm);
                    if (__any_sync(FULL, redo)) {
                        double nT, xT = -PO_INF;
                        int nKey, xKey = KEY_NONE;
                        double2 nV;
                        hairpin_scan<G>(g, A, ws, c, total, trip_total, iH, iS, RC, SCAN_NO_STAR, Tcur, &nT, &nKey,
                                             &nV);
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
                if (src >= 0 && wT > Tcur && fin(H)) {
                    if (S < PO_MIN_ENTROPY_CUTOFF) {
                        S = PO_MIN_ENTROPY;
                        H = 0.0;
                    }
                    ws.cell[k] = make_double2(H, S);  // the group's lanes store the same value
                }
            }
            __syncwarp();  // run tables are rewritten by the next batch
        }
        // phase C: calc_hairpin, keep the lower dG of {loop closure, current value}
        for (int k0 = cs; k0 < ce; k0 += 32) {
            const bool act = k0 + lane < ce;
            const int k = act ? k0 + lane : ce - 1;
            const int i = CODE_I(ws.code[k]);
            const double2 cur = ws.cell[k];
            double2 v = hairpin_closure(tb, X, gt, ws.s1, i, j, cur);
            const double2 sh = rshT[END_IDX(ws.s1[i], ws.s1[i + 1], ws.s1[j + 1])];
            const double G1 = v.x + sh.x - PO_TEMP_KELVIN * (v.y + sh.y);
            const double G2 = cur.x + sh.x - PO_TEMP_KELVIN * (cur.y + sh.y);
            if (G2 < G1) v = cur;
            if (COUNT) cnt += act ? 2 : 0;
            if (fin(v.x)) {
                if (v.y < PO_MIN_ENTROPY_CUTOFF) v = make_double2(0.0, PO_MIN_ENTROPY);
                if (act) ws.cell[k] = v;
            }
        }
        __syncwarp();
    }

    hairpin_finish<HairpinWs<CAP>, COUNT>(tb, X, A, gt, ws, len, K, r, out, relax, cnt);
}
