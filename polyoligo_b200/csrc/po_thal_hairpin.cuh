// po_thal_hairpin.cuh -- thal type HAIRPIN (calcHairpin) for one warp.
// Restates upstream thal.c: initMatrix2, fillMatrix2 (maxTM2, CBI /
// calc_bulge_internal2, calc_hairpin), calc_terminal_bp (END5_1..4),
// tracebacku and drawHairpin.
//
// Finite cells (i,j): j - i >= 4 and bases complementary.  They are kept in
// upstream's fill order (j ascending, i descending), so the cells enclosed by
// (i,j) are a prefix of the list.
#pragma once
#include "po_thal.cuh"

// loop closed by (i,j) outside and (ii,jj) inside, WITHOUT the enclosed cell's
// own value (upstream calc_bulge_internal2 in traceback==1 form); the forward
// pass adds the enclosed value afterwards in the same association order.
__device__ __forceinline__ double2 hairpin_loop_term(const SmemTables& tb, const uint8_t* s, int i, int j, int ii,
                                                     int jj, int l1, int l2) {
    const int ls = l1 + l2 - 1;
    double H, S;
    if (l1 == 0 || l2 == 0) {
        if (l1 + l2 == 1) {
            const double2 st = tb.stack[quad(s[i], s[ii], s[j], s[jj])];
            H = tb.bulge[ls].x + st.x;
            S = tb.bulge[ls].y + st.y;
        } else {
            H = tb.bulge[ls].x + at_h(s[i], s[j]) + at_h(s[ii], s[jj]);
            S = tb.bulge[ls].y + at_s(s[i], s[j]) + at_s(s[ii], s[jj]);
        }
    } else if (l1 == 1 && l2 == 1) {
        const double2 a = tb.stackint2[quad(s[i], s[i + 1], s[j], s[j - 1])];
        const double2 b = tb.stackint2[quad(s[jj], s[jj + 1], s[ii], s[ii - 1])];
        S = a.y + b.y;
        H = a.x + b.x;
    } else {
        const double2 a = tb.tstack[quad(s[i], s[i + 1], s[j], s[j - 1])];
        const double2 b = tb.tstack[quad(s[jj], s[jj + 1], s[ii], s[ii - 1])];
        const int asym = abs(l1 - l2);
        H = tb.interior[ls].x + a.x + b.x + (ILAH_D * asym);
        S = tb.interior[ls].y + a.y + b.y + (ILAS_D * asym);
    }
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

// upstream calc_hairpin() without its final dG comparison: the hairpin-loop
// closure value of (i,j); `cur` is the current DP value of the cell.
__device__ __forceinline__ double2 hairpin_closure(const SmemTables& tb, const PoTables* gt, const uint8_t* s, int i,
                                                   int j, double2 cur) {
    const int loopSize = j - i - 1;
    double H, S;
    const int li = loopSize <= 30 ? loopSize - 1 : 29;
    H = tb.hairpin[li].x;
    S = tb.hairpin[li].y;
    if (loopSize > 3) {
        const int a = s[i], b = s[i + 1], c = s[j], d = s[j - 1];
        const double2 t = tb.tstack2[quad(a, b, c, d)];
        H += t.x;
        S += t.y;
    } else if (loopSize == 3) {
        H += at_h(s[i], s[j]);
        S += at_s(s[i], s[j]);
    }
    if (loopSize == 3) {
        if (gt->n_tri) {
            int key = 0;
#pragma unroll
            for (int c = 0; c < 5; ++c) key |= s[i + c] << (2 * c);
            const double2 b = bonus_lookup(gt->tri_key, gt->tri_val, gt->n_tri, key);
            H += b.x;
            S += b.y;
        }
    } else if (loopSize == 4) {
        if (gt->n_tetra) {
            int key = 0;
#pragma unroll
            for (int c = 0; c < 6; ++c) key |= s[i + c] << (2 * c);
            const double2 b = bonus_lookup(gt->tetra_key, gt->tetra_val, gt->n_tetra, key);
            H += b.x;
            S += b.y;
        }
    }
    if (!fin(H)) {
        H = PO_INF;
        S = -1.0;
    }
    if (H > 0 && S > 0 && (!(cur.x > 0) || !(cur.y > 0))) {
        H = PO_INF;
        S = -1.0;
    }
    return make_double2(H, S);
}

// index of finite cell (i,j) in the list; caller guarantees it is finite
template <int CAP>
__device__ __forceinline__ int hp_index(const WarpWs<CAP>& ws, int i, int j) {
    const unsigned long long cmask = ws.mask[3 - ws.s1[j]] & ((1ull << (j - 4)) - 1ull);
    return ws.start[j] + __popcll(cmask >> i);
}
template <int CAP>
__device__ __forceinline__ bool hp_finite(const WarpWs<CAP>& ws, int i, int j) {
    if (j - i < PO_MIN_HRPN_LOOP + 1 || i < 1) return false;
    return ((ws.mask[3 - ws.s1[j]] >> (i - 1)) & 1ull) != 0;
}

// Scan of the candidates enclosed by (i,j).  mode as in the dimer; additionally
// LAST_GE selects the LAST candidate (largest key) with T1 >= Tcur, which is
// what upstream's CBI(traceback = 2) leaves in its output argument.
enum { HP_SCAN_LAST_GE = 3 };
template <int CAP>
__device__ __forceinline__ void hairpin_scan(const SmemTables& tb, const WarpWs<CAP>& ws, int lane, int i, int j,
                                             int k_lo, int k_hi, int max_loop, double iH, double iS, double RC,
                                             int mode, double Tcur, double* bT, int* bKey, double2* bV, int* cnt) {
    double bestT = -PO_INF;
    int bestKey = KEY_NONE;
    double2 bestV = make_double2(PO_INF, -1.0);
    for (int kp = k_lo + lane; kp < k_hi; kp += 32) {
        const int c = ws.coord[kp];
        const int ii = c >> 8, jj = c & 255;
        if (ii <= i) continue;
        const int l1 = ii - i - 1, l2 = j - jj - 1, sum = l1 + l2;
        if (sum < 1 || sum > max_loop) continue;
        const int key = KEY_OF(sum, l1);
        if (mode == SCAN_NO_STAR && key == KEY_STAR) continue;
        if (mode == SCAN_BEFORE_STAR && key >= KEY_STAR) continue;
        if (cnt) ++*cnt;
        double2 v = hairpin_loop_term(tb, ws.s1, i, j, ii, jj, l1, l2);
        const double2 pv = ws.cell[kp];
        v.x += pv.x;
        v.y += pv.y;
        v = hp_fix(v.x, v.y);
        const double T1 = (v.x + iH) / ((v.y + iS) + RC);
        if (mode == HP_SCAN_LAST_GE) {
            if (T1 >= Tcur && (bestKey == KEY_NONE || key > bestKey)) {
                bestT = T1;
                bestKey = key;
                bestV = v;
            }
        } else if (T1 > bestT || (T1 == bestT && key < bestKey)) {
            bestT = T1;
            bestKey = key;
            bestV = v;
        }
    }
    *bT = bestT;
    *bKey = bestKey;
    *bV = bestV;
}

// upstream END5_1..END5_4 for position i, all four variants at once: lanes
// 8g..8g+7 serve variant g+1.  Returns the variant's {H_max, S_max} in every
// lane of its group.
template <int CAP>
__device__ __forceinline__ double2 end5_all(const SmemTables& tb, const WarpWs<CAP>& ws, int lane, int i, double iH,
                                            double iS, double RC) {
    const int v = (lane >> 3) + 1, sub = lane & 7;
    const uint8_t* s = ws.s1;
    const int q = (v <= 2) ? i : i - 1;  // paired 3' base
    double bestT = -PO_INF;
    int bestKey = KEY_NONE;
    double2 best = make_double2(PO_INF, -1.0);
    if (q >= 5) {
        const int ks = ws.start[q], ke = ws.start[q + 1];
        const double T2 = (0 + iH) / (0 + iS + RC);
        for (int kc = ks + sub; kc < ke; kc += 8) {
            const int p = ws.coord[kc] >> 8;
            const int k = (v == 1 || v == 3) ? p - 1 : p - 2;  // unpaired/previous-structure prefix ends at k
            if (k < 0) continue;
            double2 x = make_double2(0.0, 0.0);
            if (v == 2) x = tb.dangle5[(s[i] * 4 + s[k + 2]) * 4 + s[k + 1]];                    // Hd5(i, k+2)
            else if (v == 3) x = tb.dangle3[(s[i - 1] * 4 + s[i]) * 4 + s[k + 1]];               // Hd3(i-1, k+1)
            else if (v == 4) x = tb.tstack2[quad(s[i - 1], s[i], s[k + 2], s[k + 1])];           // Htstack(i-1, k+2)
            const double2 dv = ws.cell[kc];
            const double hk = ws.end5[0][k], sk = ws.end5[1][k];
            const double T1p = (hk + iH) / (sk + iS + RC);
            double H, S;
            const double aH = at_h(s[p], s[q]), aS = at_s(s[p], s[q]);
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
            const double T1 = (H + iH) / (S + iS + RC);
            if (!(S > PO_MIN_ENTROPY_CUTOFF)) continue;
            // upstream keeps the first k with the strictly largest T1; -inf never wins
            if (T1 > bestT || (T1 == bestT && k < bestKey)) {
                if (T1 > -PO_INF) {
                    bestT = T1;
                    bestKey = k;
                    best = make_double2(H, S);
                }
            }
        }
    }
    // arg-max inside each 8-lane group
#pragma unroll
    for (int m = 4; m >= 1; m >>= 1) {
        const double ot = shfl_xor_d(bestT, m);
        const int ok = __shfl_xor_sync(FULL, bestKey, m);
        const double oh = shfl_xor_d(best.x, m), os = shfl_xor_d(best.y, m);
        if (ot > bestT || (ot == bestT && ok < bestKey)) {
            bestT = ot;
            bestKey = ok;
            best = make_double2(oh, os);
        }
    }
    return best;
}

template <int CAP, bool COUNT>
__device__ void thal_hairpin_warp(const SmemTables& tb, const PoTables* gt, WarpWs<CAP>& ws, const uint4 oa,
                                  const PoConsts& K, po_thal_out* out, unsigned long long* relax) {
    const int lane = threadIdx.x & 31;
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
    for (int p = lane; p < len; p += 32) ws.s1[p + 1] = ws.s2[p + 1] = (uint8_t)po_base(oa, p);
    if (lane == 0) {
        ws.s1[0] = ws.s1[len + 1] = 4;
        ws.s2[0] = ws.s2[len + 1] = 4;
    }
    const double iH = 0.0, iS = -0.00000000001, RC = 0.0;
    const int max_loop = K.max_loop;
    __syncwarp();
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const unsigned lo = __ballot_sync(FULL, lane + 1 <= len && ws.s1[lane + 1] == b);
        const unsigned hi = __ballot_sync(FULL, lane + 33 <= len && ws.s1[lane + 33] == b);
        if (lane == 0) ws.mask[b] = (unsigned long long)lo | ((unsigned long long)hi << 32);
    }
    __syncwarp();
    // rows i <= j-4 of column j that pair with base j
#define COLMASK(j) ((j) >= 5 ? (ws.mask[3 - ws.s1[j]] & ((1ull << ((j)-4)) - 1ull)) : 0ull)
    {
        const int j0 = lane + 1, j1 = lane + 33;
        const int c0 = j0 <= len ? __popcll(COLMASK(j0)) : 0;
        const int c1 = j1 <= len ? __popcll(COLMASK(j1)) : 0;
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
        ws.start[j0] = (uint16_t)(x0 - c0);
        if (j1 < 64) ws.start[j1] = (uint16_t)(tot0 + x1 - c1);
    }
    if (lane == 0) ws.start[0] = 0;
    __syncwarp();
    for (int j = 5; j <= len; ++j) {
        const unsigned long long m = COLMASK(j);
        const int base = ws.start[j];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int bit = lane + 32 * h;  // row i = bit + 1; list order is i descending
            if ((m >> bit) & 1ull) ws.coord[base + __popcll(m >> (bit + 1))] = (uint16_t)(((bit + 1) << 8) | j);
        }
    }
    __syncwarp();

    int cnt = 0;
    // ---- fillMatrix2 -----------------------------------------------------------
    for (int j = 5; j <= len; ++j) {
        const int cs = ws.start[j], ce = ws.start[j + 1];
        if (cs == ce) continue;
        // phase A: maxTM2 (stack on (i+1, j-1)) for every finite cell of the column
        for (int k = cs + lane; k < ce; k += 32) {
            const int i = ws.coord[k] >> 8;
            double S0 = PO_MIN_ENTROPY, H0 = 0.0;
            const double T0 = (H0 + iH) / (S0 + iS + RC);
            double2 pv = make_double2(PO_INF, -1.0);
            if (hp_finite<CAP>(ws, i + 1, j - 1)) pv = ws.cell[hp_index<CAP>(ws, i + 1, j - 1)];
            const double2 st = tb.stack[quad(ws.s1[i], ws.s1[i + 1], ws.s1[j], ws.s1[j - 1])];
            double S1 = pv.y + st.y;
            double H1 = pv.x + st.x;
            const double T1 = (H1 + iH) / (S1 + iS + RC);
            if (S1 < PO_MIN_ENTROPY_CUTOFF) {
                S1 = PO_MIN_ENTROPY;
                H1 = 0.0;
            }
            if (S0 < PO_MIN_ENTROPY_CUTOFF) {
                S0 = PO_MIN_ENTROPY;
                H0 = 0.0;
            }
            ws.cell[k] = (T1 > T0) ? make_double2(H1, S1) : make_double2(H0, S0);
        }
        __syncwarp();
        // phase B: CBI, one cell at a time, lanes over the enclosed cells
        const int k_lo = ws.start[max(0, j - 1 - max_loop)];
        for (int k = cs; k < ce; ++k) {
            const int i = ws.coord[k] >> 8;
            if (j - i < 6) continue;  // no room for an enclosed pair plus a loop
            const double2 cur = ws.cell[k];
            const double Tcur = (cur.x + iH) / ((cur.y + iS) + RC);
            double bT, wT;
            int bKey, wKey;
            double2 bV;
            hairpin_scan<CAP>(tb, ws, lane, i, j, k_lo, cs, max_loop, iH, iS, RC, SCAN_ALL, Tcur, &bT, &bKey, &bV,
                              COUNT ? &cnt : nullptr);
            int src = warp_argmax(bT, bKey, &wT, &wKey);
            if (src >= 0 && wKey == KEY_STAR) {
                double pT, qT;
                int pKey, qKey;
                double2 pV;
                hairpin_scan<CAP>(tb, ws, lane, i, j, k_lo, cs, max_loop, iH, iS, RC, SCAN_BEFORE_STAR, Tcur, &pT,
                                  &pKey, &pV, nullptr);
                warp_argmax(pT, pKey, &qT, &qKey);
                const double Tm = qT > Tcur ? qT : Tcur;
                if (!((wT - Tm) >= PO_SMALL_NON_ZERO && wT > Tm)) {
                    hairpin_scan<CAP>(tb, ws, lane, i, j, k_lo, cs, max_loop, iH, iS, RC, SCAN_NO_STAR, Tcur, &bT,
                                      &bKey, &bV, nullptr);
                    src = warp_argmax(bT, bKey, &wT, &wKey);
                }
            }
            if (src >= 0 && wT > Tcur) {
                double H = shfl_d(bV.x, src), S = shfl_d(bV.y, src);
                if (fin(H)) {
                    if (S < PO_MIN_ENTROPY_CUTOFF) {
                        S = PO_MIN_ENTROPY;
                        H = 0.0;
                    }
                    if (lane == 0) ws.cell[k] = make_double2(H, S);
                }
            }
            __syncwarp();
        }
        // phase C: calc_hairpin, keep the lower dG of {loop closure, current value}
        for (int k = cs + lane; k < ce; k += 32) {
            const int i = ws.coord[k] >> 8;
            const double2 cur = ws.cell[k];
            double2 v = hairpin_closure(tb, gt, ws.s1, i, j, cur);
            const double2 sh = rsh(tb, ws.s1, ws.s2, i, j, iH, iS, RC);
            const double G1 = v.x + sh.x - PO_TEMP_KELVIN * (v.y + sh.y);
            const double G2 = cur.x + sh.x - PO_TEMP_KELVIN * (cur.y + sh.y);
            if (G2 < G1) v = cur;
            if (COUNT) cnt += 2;
            if (fin(v.x)) {
                if (v.y < PO_MIN_ENTROPY_CUTOFF) v = make_double2(0.0, PO_MIN_ENTROPY);
                ws.cell[k] = v;
            }
        }
        __syncwarp();
    }

    // ---- calc_terminal_bp --------------------------------------------------------
    if (lane == 0) {
        ws.end5[1][0] = ws.end5[1][1] = -1.0;
        ws.end5[0][0] = ws.end5[0][1] = PO_INF;
    }
    __syncwarp();
    for (int i = 2; i <= len; ++i) {
        const double2 e = end5_all<CAP>(tb, ws, lane, i, iH, iS, RC);
        const double hp = ws.end5[0][i - 1], sp = ws.end5[1][i - 1];
        const double T1 = (hp + iH) / (sp + iS + RC);
        const double Tv = (e.x + iH) / (e.y + iS + RC);
        const double T2 = shfl_d(Tv, 0), T3 = shfl_d(Tv, 8), T4 = shfl_d(Tv, 16), T5 = shfl_d(Tv, 24);
        int mx;
        if (T1 > T2 && T1 > T3 && T1 > T4 && T1 > T5) mx = 1;
        else if (T2 > T3 && T2 > T4 && T2 > T5) mx = 2;
        else if (T3 > T4 && T3 > T5) mx = 3;
        else if (T4 > T5) mx = 4;
        else mx = 5;
        double hn = hp, sn = sp;
        if (mx > 1) {
            const double eh = shfl_d(e.x, (mx - 2) * 8), es = shfl_d(e.y, (mx - 2) * 8);
            const double G = eh - (K.temp_k * es);
            if (G < 0.0) {
                hn = eh;
                sn = es;
            }
        }
        if (COUNT) cnt += (lane == 0) ? 4 : 0;
        if (lane == 0) {
            ws.end5[0][i] = hn;
            ws.end5[1][i] = sn;
        }
        __syncwarp();
    }
    if (COUNT && relax) {
        unsigned long long c64 = (unsigned long long)cnt;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) c64 += __shfl_xor_sync(FULL, c64, m);
        if (lane == 0) atomicAdd(relax, c64);
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
            const double2 e = end5_all<CAP>(tb, ws, lane, i, iH, iS, RC);
            const double hi_ = ws.end5[0][i], si_ = ws.end5[1][i];
            const unsigned match = __ballot_sync(FULL, (lane & 7) == 0 && eq5(si_, e.y) && eq5(hi_, e.x));
            if (!match) continue;
            const int v = (__ffs(match) - 1) / 8 + 1;  // upstream's else-if chain: first matching variant only
            const uint8_t* s = ws.s1;
            const int q = (v <= 2) ? i : i - 1;
            int myKey = KEY_NONE;
            if (q >= 5) {
                const int ks = ws.start[q], ke = ws.start[q + 1];
                for (int kc = ks + lane; kc < ke; kc += 32) {
                    const int p = ws.coord[kc] >> 8;
                    const int k = (v == 1 || v == 3) ? p - 1 : p - 2;
                    if (k < 0) continue;
                    double2 x = make_double2(0.0, 0.0);
                    if (v == 2) x = tb.dangle5[(s[i] * 4 + s[k + 2]) * 4 + s[k + 1]];
                    else if (v == 3) x = tb.dangle3[(s[i - 1] * 4 + s[i]) * 4 + s[k + 1]];
                    else if (v == 4) x = tb.tstack2[quad(s[i - 1], s[i], s[k + 2], s[k + 1])];
                    const double2 dv = ws.cell[kc];
                    const double aH = at_h(s[p], s[q]), aS = at_s(s[p], s[q]);
                    const double hk = ws.end5[0][k], sk = ws.end5[1][k];
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
                    if (eq5(si_, s0) && eq5(hi_, h0)) myKey = min(myKey, k * 2);
                    else if (eq5(si_, s1v) && eq5(hi_, h1v)) myKey = min(myKey, k * 2 + 1);
                }
            }
            int wKey = myKey;
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) wKey = min(wKey, __shfl_xor_sync(FULL, wKey, m));
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
            const int k = hp_index<CAP>(ws, i, j);
            const double2 cur = ws.cell[k];
            // stack
            if (hp_finite<CAP>(ws, i + 1, j - 1)) {
                const double2 pv = ws.cell[hp_index<CAP>(ws, i + 1, j - 1)];
                const double2 st = tb.stack[quad(ws.s1[i], ws.s1[i + 1], ws.s1[j], ws.s1[j - 1])];
                if (eq5(cur.y, st.y + pv.y) && eq5(cur.x, st.x + pv.x)) {
                    if (lane == 0) ws.stk[sp_] = ((i + 1) << 8) | (j - 1);
                    ++sp_;
                    __syncwarp();
                    continue;
                }
            }
            // hairpin loop closes here
            const double2 hv = hairpin_closure(tb, gt, ws.s1, i, j, cur);
            if (eq5(cur.y, hv.y) && eq5(cur.x, hv.x)) continue;
            // bulge / interior loop
            if (j - i < 6) continue;
            const double Tcur = (cur.x + iH) / ((cur.y + iS) + RC);
            const int k_lo = ws.start[max(0, j - 1 - max_loop)], k_hi = ws.start[j];
            double bT, wT;
            int bKey, wKey;
            double2 bV;
            hairpin_scan<CAP>(tb, ws, lane, i, j, k_lo, k_hi, max_loop, iH, iS, RC, HP_SCAN_LAST_GE, Tcur, &bT, &bKey,
                              &bV, nullptr);
            // LAST candidate with T1 >= Tcur: largest key
            int mxKey = (bKey == KEY_NONE) ? -1 : bKey;
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) mxKey = max(mxKey, __shfl_xor_sync(FULL, mxKey, m));
            if (mxKey < 0) continue;
            const unsigned who = __ballot_sync(FULL, bKey == mxKey);
            const int src = __ffs(who) - 1;
            double H2 = shfl_d(bV.x, src), S2 = shfl_d(bV.y, src);
            if (fin(H2) && S2 < PO_MIN_ENTROPY_CUTOFF) {
                S2 = PO_MIN_ENTROPY;
                H2 = 0.0;
            }
            if (!(eq5(cur.y, S2) && eq5(cur.x, H2))) continue;
            // first candidate (upstream order) whose value reproduces the cell
            int myKey = KEY_NONE;
            for (int kp = k_lo + lane; kp < k_hi; kp += 32) {
                const int c = ws.coord[kp];
                const int ii = c >> 8, jj = c & 255;
                if (ii <= i) continue;
                const int l1 = ii - i - 1, l2 = j - jj - 1, sum = l1 + l2;
                if (sum < 1 || sum > max_loop) continue;
                const int key = KEY_OF(sum, l1);
                if (key >= myKey) continue;
                double2 v = hairpin_loop_term(tb, ws.s1, i, j, ii, jj, l1, l2);
                v = hp_fix(v.x, v.y);  // traceback == 1 form: every branch reports its value
                const double2 pv = ws.cell[kp];
                if (eq5(cur.y, v.y + pv.y) && eq5(cur.x, v.x + pv.x)) myKey = key;
            }
            wKey = myKey;
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) wKey = min(wKey, __shfl_xor_sync(FULL, wKey, m));
            if (wKey == KEY_NONE) continue;
            const int sum = wKey >> 6, l1 = wKey & 63;
            const int ii = i + l1 + 1, jj = j - (sum - l1) - 1;
            if (lane == 0) ws.stk[sp_] = (ii << 8) | jj;
            ++sp_;
            __syncwarp();
        }
    }
#undef COLMASK

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
