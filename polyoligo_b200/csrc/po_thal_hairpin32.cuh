// po_thal_hairpin32.cuh -- thal type HAIRPIN for one warp, FIRST TIER: oligo <= 32 nt and at most
// CAP finite cells.  Same arithmetic as thal_hairpin_warp (po_thal_hairpin.cuh: upstream initMatrix2,
// fillMatrix2 with maxTM2 / CBI / calc_bulge_internal2 / calc_hairpin, then calc_terminal_bp,
// tracebacku, drawHairpin through the shared hairpin_finish) -- bit-identical results -- with the fill
// organised like the dimer's first tier (po_thal_dimer32.cuh), mirrored for the unimolecular table:
//
//   * finite cells (i,j), j - i >= 4, are filled column by column (j ascending, i descending inside a
//     column).  The candidates ENCLOSED by (i,j) with an interior loop (l1 = ii-i-1 >= 1, l2 = j-jj-1
//     >= 1, not 1x1) lie in rows >= i+2 of columns <= j-2: `byrow` holds the interior-capable cells of
//     columns <= j-2 sorted by (row descending, column ascending), so the set is a PREFIX of that list
//     for every cell of column j (its last entry is the 1x1 cell (i+2,j-2), taken out).  Column j-2 is
//     merged in before column j is relaxed.
//   * bulges (l2 == 0: column j-1, contiguous in the fill-order list; l1 == 0: row i+1, through a
//     static row-major index) and the 1x1 loop form the short list evaluated with the general term.
//   * maxTM2, the candidate-set descriptor and T of the current value are computed one lane per cell of
//     the column; the cells are then relaxed 32/G at a time.
//   * call-free Newton-Raphson quotients (po_div_nr) where the operands are known finite; 32-bit masks.
//
// Replaces libprimer3 thal() type 4 as reached from the reference at src/polyoligo/lib_primer3.py:60.
#pragma once
#include "po_thal_dimer32.cuh"
#include "po_thal_hairpin.cuh"

template <int CAP>
struct Hairpin32Ws {
    static_assert(CAP <= 512, "a list index must fit the 9 spare bits of a cell code");
    double2 cell[CAP];    // fill order: j ascending, i descending
    uint32_t code[CAP];   // i | j << 6 | tq << 12 | b << 20 | capable << 22 | list index << 23
    uint32_t byrow[CAP];  // codes of the interior-capable cells of columns <= j-2, (row desc, column asc)
    double2 ycell[4][4];
    struct Desc {
        double Tcur;
        uint16_t n_fast, total, nA, nB;
        uint16_t kA0, kB0, kstar;
        uint8_t q, i;
    } desc[32];
    double end5[2][36];   // HEND5 / SEND5
    double end5t[36];
    unsigned long long mask[4];  // 64-bit copy of m32 for hairpin_finish and its helpers
    RunTable runs[1];            // traceback
    int stk[64];                 // traceback stack
    uint32_t m32[4];      // bit i-1 set iff s[i] == b
    uint16_t start[36];   // first list index of column j
    uint16_t rstart[36];  // first rowmajor index of row r
    uint16_t rowge[36];   // number of byrow entries with row >= r
    typename std::conditional<(CAP <= 256), uint8_t, uint16_t>::type rowmajor[CAP];  // every finite cell, row-major
    uint8_t s1[40];
    uint2 dpre[CAP];      // per cell: the sequence-only part of its candidate-set descriptor (fill kernel)
    uint32_t capcol[36];  // interior-capable cells of column c (bit i-1)
};

// rows i <= j-4 of column j that pair with base j (bit i-1)
#define HP32_COLM(ws, j) ((j) >= 5 ? ((ws).m32[3 - (ws).s1[j]] & ((1u << ((j)-4)) - 1u)) : 0u)
// columns jj >= r+4 of row r that pair with base r (bit jj-1)
#define HP32_ROWM(ws, r) ((r) + 3 < 32 ? ((ws).m32[3 - (ws).s1[r]] & ~((1u << ((r) + 3)) - 1u)) : 0u)

__device__ __forceinline__ int popc_from32(uint32_t m, int x) {  // set bits at or above bit x, 0 <= x <= 32
    return x >= 32 ? 0 : __popc(m >> x);
}

// merges column c (final since column c+1 was relaxed) into ws.byrow; returns the new length
template <int CAP>
__device__ __forceinline__ int byrow32_merge_col(const SmemTables& tb, Hairpin32Ws<CAP>& ws, int c, int n_by, int lane) {
    const uint32_t cm = HP32_COLM(ws, c);
    // interior-capable cells of the column (tabulated per column by the prologue)
    const uint32_t newm = ws.capcol[c];
    if (newm == 0u) return n_by;  // warp-uniform
    // existing entries move up by the number of new cells in higher rows; chunks from the top down
    for (int base = ((n_by - 1) >> 5) << 5; base >= 0; base -= 32) {
        const int t = base + lane;
        const bool act = t < n_by;
        const uint32_t e = ws.byrow[act ? t : 0];
        const int np = t + popc_from32(newm, CODE_I(e));  // rows > pi are bits >= pi
        __syncwarp();
        if (act) ws.byrow[np] = e;
        __syncwarp();
    }
    if ((newm >> lane) & 1u) {  // row lane + 1
        const int pos = ws.rowge[lane + 1] + popc_from32(newm, lane + 1);
        ws.byrow[pos] = ws.code[ws.start[c] + popc_from32(cm, lane + 1)];
    }
    __syncwarp();
    ws.rowge[lane + 1] = (uint16_t)(ws.rowge[lane + 1] + popc_from32(newm, lane));
    __syncwarp();
    return n_by + __popc(newm);
}

struct Hairpin32Cell {
    int i, j, b1q, kself;
    int n_fast, nA, kA0, nB, kB0, kstar, total;
    double2 y0;
};

// One pass over the candidates of the group's cell (see dimer32_scan).  The closing pair's term
// comes first in upstream's sums, the enclosed pair's second; only a non-finite sum is rejected.
template <int G, int CAP>
__device__ __forceinline__ void hairpin32_scan(const Grp<G>& g, const TabAddr& A, const Hairpin32Ws<CAP>& ws,
                                               const Hairpin32Cell& c, int trip_f, int trip_s, int max_loop,
                                               double iH, double iS, double RC, int mode, double& bestT,
                                               int& bestKey, double2& bestV) {
    if (mode != SCAN_BEFORE_STAR) {
        const unsigned Lbase = A.ltab - 16u;          // interior[sum - 1]
        const unsigned Pbase = A.stack + 512u * 16u;  // tstack[tq]
#pragma unroll 1
        for (int t0 = 0; t0 < trip_f; t0 += G) {
            const int t = t0 + g.lane;
            bool act = t < c.n_fast;
            const uint32_t e = ws.byrow[act ? t : 0];
            const double2 pv = ws.cell[CODE_K(e)];
            const int l1 = CODE_I(e) - c.i - 1, l2 = c.j - CODE_J(e) - 1;
            const int sum = l1 + l2;
            act &= sum <= max_loop;
            const double2 L = lds_d2(Lbase + (unsigned)sum * 16u);
            const double2 P = lds_d2(Pbase + ((e >> 8) & 0xff0u));
            const double asymS = lds_d(A.asym + (unsigned)abs(l1 - l2) * 8u);
            const double H0 = (L.x + c.y0.x) + P.x;
            const double S0 = ((L.y + c.y0.y) + P.y) + asymS;
            const double H = H0 + pv.x, S = S0 + pv.y;
            const bool ok = act & fin(H);
            const double T1 = po_div_nr(ok ? H + iH : 1.0, ok ? (S + iS) + RC : 1.0);
            const int key = KEY_OF(sum, l1);
            if (ok & ((T1 > bestT) | ((T1 == bestT) & (key < bestKey)))) {
                bestT = T1;
                bestKey = key;
                bestV = make_double2(H, S);
            }
        }
    }
    const double2* yc = ws.ycell[g.idx];
    HairpinCellCtx cc;
    cc.i = c.i;
    cc.j = c.j;
    cc.b1q = c.b1q;
    for (int t0 = 0; t0 < trip_s; t0 += G) {
        const int t = t0 + g.lane;
        bool act = t < c.total;
        const int tb_ = min(max(t - c.nA, 0), max(c.nB - 1, 0));
        const int krow = ws.rowmajor[min(c.kB0 + tb_, CAP - 1)];
        // idle lanes evaluate the cell against itself (l1 = l2 = -1: every table offset stays in range)
        const int kp = !act ? c.kself : (t < c.nA ? c.kA0 + t : (t < c.nA + c.nB ? krow : c.kstar));
        int key;
        double2 v = hairpin_loop_term(A, ws, yc, cc, kp, &key);
        if (mode == SCAN_NO_STAR) act &= key != KEY_STAR;
        if (mode == SCAN_BEFORE_STAR) act &= key < KEY_STAR;
        const double2 pv = ws.cell[kp];
        v.x += pv.x;
        v.y += pv.y;
        const bool ok = act & fin(v.x);
        const double T1 = po_div_nr(ok ? v.x + iH : 1.0, ok ? (v.y + iS) + RC : 1.0);
        if (ok & ((T1 > bestT) | ((T1 == bestT) & (key < bestKey)))) {
            bestT = T1;
            bestKey = key;
            bestV = v;
        }
    }
}

// PHASE 1: build the table and write its cell values to `cells` (CAP double2 per item, global memory);
// PHASE 2: rebuild the index structures, read the values back and run hairpin_finish.  Two kernels
// instead of one because the fill and the finish are each ~20 KB of hot code: resident together in
// one kernel, with every warp at its own place, they overflow the SM's 32 KB instruction cache (ncu:
// 4.2 no_instruction stall cycles per issued instruction, issue slots 57 % busy); the 2 KB per hairpin
// that cross HBM instead cost ~1 % of the kernels' time.
template <int G, int CAP, int PHASE>
__device__ void thal_hairpin32_warp(const SmemTables& tb, const SmemExtra& X, const TabAddr& A, const PoTables* gt,
                                    Hairpin32Ws<CAP>& ws, const uint4 oa, const PoConsts& K, po_thal_out* out,
                                    double2* __restrict__ cells) {
    const int lane = threadIdx.x & 31;
    const Grp<G> g = Grp<G>::make();
    const Grp<32> gw = Grp<32>::make();
    constexpr int NG = 32 / G;
    static_assert(NG <= 4, "ycell holds four concurrent cells");
    const int len = po_len(oa);  // 1 .. 32 (checked by the caller)
    po_thal_out r;
    r.tm = 0.0;
    r.dg = r.dh = r.ds = 0.0;
    r.flags = 0;
    r.align_end_1 = r.align_end_2 = 0;
    __syncwarp();
    if (lane < len) ws.s1[lane + 1] = (uint8_t)po_base(oa, lane);
    if (lane == 0) ws.s1[0] = ws.s1[len + 1] = 4;
    const double iH = 0.0, iS = -0.00000000001, RC = 0.0;
    const double2* rshT = X.rsh[2];
    const int max_loop = K.max_loop;
    __syncwarp();
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const unsigned mb = __ballot_sync(FULL, lane < len && ws.s1[lane + 1] == b);
        if (lane == 0) {
            ws.m32[b] = mb;
            ws.mask[b] = (unsigned long long)mb;
        }
    }
    __syncwarp();
    {
        // column starts (fill-order list) and row starts (row-major index)
        const int jr = lane + 1;
        const int c0 = jr <= len ? __popc(HP32_COLM(ws, jr)) : 0;
        const int x0 = gw.scan_incl(c0);
        ws.start[jr] = (uint16_t)(x0 - c0);
        if (lane == 31) ws.start[33] = (uint16_t)x0;
        const int d0 = jr <= len ? __popc(HP32_ROWM(ws, jr)) : 0;
        const int y0 = gw.scan_incl(d0);
        ws.rstart[jr] = (uint16_t)(y0 - d0);
        ws.rowge[lane] = 0;
        if (lane < 4) ws.rowge[32 + lane] = 0;
        if (lane == 0) {
            ws.start[0] = 0;
            ws.byrow[0] = 0;  // read (and ignored) by idle lanes while the list is empty
        }
    }
    __syncwarp();
    const int ncell = ws.start[len + 1];
    // cell codes: built by the fill kernel, which hands them to the finish kernel behind the cell values
    if constexpr (PHASE == 2) {
        const uint32_t* codes = reinterpret_cast<const uint32_t*>(cells + CAP);
        for (int k = lane; k < ncell; k += 32) {
            ws.code[k] = codes[k];
            ws.cell[k] = cells[k];
        }
        __syncwarp();
        hairpin_finish<Hairpin32Ws<CAP>, false>(tb, X, A, gt, ws, len, K, r, out, nullptr, 0);
        return;
    }
    for (int j = 5; j <= len; ++j) {
        const uint32_t m = HP32_COLM(ws, j);
        if ((m >> lane) & 1u) {  // row i = lane + 1; list order is i descending
            const int bj = ws.s1[j], bjn = ws.s1[j + 1] & 3;
            const int bi = ws.s1[lane + 1];
            const int tq = quad(bj, bjn, bi, ws.s1[lane] & 3);  // [s[j]][s[j+1]][s[i]][s[i-1]]
            const int k = ws.start[j] + popc_from32(m, lane + 1);
            ws.code[k] = (uint32_t)(lane + 1) | ((uint32_t)j << 6) | ((uint32_t)tq << 12) | ((uint32_t)bi << 20) |
                         (fin(tb.tstack[tq].x) ? CODE_CAP_BIT : 0u) | ((uint32_t)k << 23);
        }
    }
    __syncwarp();
    for (int k0 = 0; k0 < ncell; k0 += 32) {
        const int k = k0 + lane;
        if (k < ncell) {
            const uint32_t code = ws.code[k];
            const int ci = CODE_I(code);
            ws.rowmajor[ws.rstart[ci] + popc_below32(HP32_ROWM(ws, ci), CODE_J(code) - 1)] =
                (typename std::remove_reference<decltype(ws.rowmajor[0])>::type)k;
        }
    }
    __syncwarp();

    // The hairpin-loop closure of every cell (sequence only), 32 cells at a time; it waits in the cell's
    // own slot, which nothing reads before the cell's column is reached.
    for (int k0 = 0; k0 < ncell; k0 += 32) {
        const int k = min(k0 + lane, ncell - 1);
        const uint32_t code = ws.code[k];
        const double2 c = hairpin_closure_static(tb, X, gt, ws.s1, CODE_I(code), CODE_J(code));
        if (k0 + lane < ncell) ws.cell[k] = c;
    }
    __syncwarp();

    {
        // interior-capable cells of every column: tstack[s[c]][s[c+1]][s[i]][s[i-1]] finite (the sentinel
        // before the first base reads as base 0, exactly like the tq field of the cell codes)
        const int c = lane + 1;
        uint32_t cap = 0u;
        if (c >= 5 && c <= len) {
            const int bc = ws.s1[c], bcn = ws.s1[c + 1] & 3;
            uint32_t capm = 0u;
#pragma unroll
            for (int nb = 0; nb < 4; ++nb)
                if (fin(tb.tstack[quad(bc, bcn, 3 - bc, nb)].x)) capm |= (ws.m32[nb] << 1) | (nb == 0 ? 1u : 0u);
            cap = HP32_COLM(ws, c) & capm;
        }
        ws.capcol[c] = cap;
        if (lane < 4) ws.capcol[(lane == 0) ? 0 : 32 + lane] = 0u;
    }
    // The sequence-only part of every cell's candidate-set descriptor, 32 cells at a time (the per-column
    // pass runs with 3-4 lanes busy): bulge ranges, the 1x1 cell, table indices.  The interior-loop count
    // depends on the merge state of the candidate list and stays in the column pass.
    for (int k0 = 0; k0 < ncell; k0 += 32) {
        const int k = min(k0 + lane, ncell - 1);
        const uint32_t code = ws.code[k];
        const int i = CODE_I(code), j = CODE_J(code);
        uint2 pk = make_uint2(0u, 0u);
        if (j - i >= 7) {
            const int bi = ws.s1[i], bin = ws.s1[i + 1], bj = ws.s1[j], bjm = ws.s1[j - 1];
            const uint32_t cmj1 = HP32_COLM(ws, j - 1);
            const uint32_t cmj2 = j >= 7 ? HP32_COLM(ws, j - 2) : 0u;
            const int q = quad(bi, bin & 3, bj, bjm & 3);
            const double2 y0 = tb.tstack[q], y3 = tb.stackint2[q];
            // 1x1 loop: cell (i+2, j-2)
            const bool star_fin = (j - i >= 8) && ((cmj2 >> (i + 1)) & 1u);
            const int kstar = star_fin ? ws.start[j - 2] + popc_from32(cmj2, i + 2) : 0;
            const bool star_cap = star_fin && (ws.code[kstar] & CODE_CAP_BIT);
            // bulges: column j-1, rows i+2 .. min(j-5, i+1+max_loop) (descending in the list);
            //         row i+1, columns max(i+5, j-1-max_loop) .. j-2
            const int hiA = min(j - 5, i + 1 + max_loop);
            const int aboveA = popc_from32(cmj1, max(hiA, 0));
            const int nA = max(popc_from32(cmj1, i + 1) - aboveA, 0);
            const uint32_t rm = HP32_ROWM(ws, i + 1);
            const int loB = max(i + 5, j - 1 - max_loop);
            const int belowB = popc_below32(rm, loB - 1);
            const int nB = max(popc_below32(rm, j - 2) - belowB, 0);
            const int nS = (star_fin && fin(y3.x) && max_loop >= 2) ? 1 : 0;
            const unsigned flags = (fin(y0.x) ? 1u : 0u) | (star_cap ? 2u : 0u);
            pk.x = (unsigned)(nA + nB + nS) | ((unsigned)nA << 8) | ((unsigned)nB << 16) | (flags << 24);
            pk.y = (unsigned)(ws.start[j - 1] + aboveA) | ((unsigned)(ws.rstart[i + 1] + belowB) << 8) |
                   ((unsigned)kstar << 16) | ((unsigned)q << 24);
        }
        if (k0 + lane < ncell) ws.dpre[k] = pk;
    }
    __syncwarp();

    // maxTM2's empty-structure quotient (0 + iH) / (MIN_ENTROPY + iS + RC): the same for every cell
    const double T0_const = po_div(0.0 + iH, PO_MIN_ENTROPY + iS + RC);
    // ---- fillMatrix2 -----------------------------------------------------------
    int n_by = 0;  // length of ws.byrow
    for (int j = 5; j <= len; ++j) {
        if (j >= 7) n_by = byrow32_merge_col<CAP>(tb, ws, j - 2, n_by, lane);
        const int cs = ws.start[j], ce = ws.start[j + 1];
        if (cs == ce) continue;
        const int bj = ws.s1[j];
        const uint32_t cmj1 = HP32_COLM(ws, j - 1);
        const uint32_t cmj2 = j >= 7 ? HP32_COLM(ws, j - 2) : 0u;
        // One lane per cell of the column (<= 28): maxTM2 (stack on (i+1, j-1)), candidate-set descriptor
        double2 closure;   // the cell's loop closure, parked in its slot by the prologue; used in phase C
        {
            const bool act = cs + lane < ce;
            const int k = act ? cs + lane : ce - 1;
            closure = ws.cell[k];
            const int i = CODE_I(ws.code[k]);
            const int bi = ws.s1[i], bin = ws.s1[i + 1], bjm = ws.s1[j - 1];
            const double S0 = PO_MIN_ENTROPY, H0 = 0.0;
            const bool pfin = hp_finite(ws, i + 1, j - 1);
            double2 pv = ws.cell[pfin ? hp_index(ws, i + 1, j - 1) : 0];
            if (!pfin) pv = make_double2(PO_INF, -1.0);
            const double2 st = tb.stack[quad(bi, bin, bj, bjm)];
            double S1 = pv.y + st.y;
            double H1 = pv.x + st.x;
            // H1 is finite, or +inf without a stackable predecessor: +inf over the (finite) denominator
            const double den1 = S1 + iS + RC;
            const bool f1 = fin(H1);
            const double q1 = po_div_nr(f1 ? H1 + iH : 1.0, den1);
            const double T1 = f1 ? q1 : (den1 < 0.0 ? -PO_INF : PO_INF);
            if (S1 < PO_MIN_ENTROPY_CUTOFF) {
                S1 = PO_MIN_ENTROPY;
                H1 = 0.0;
            }
            const double2 v = (T1 > T0_const) ? make_double2(H1, S1) : make_double2(H0, S0);
            typename Hairpin32Ws<CAP>::Desc d;
            d.n_fast = d.total = d.nA = d.nB = d.kA0 = d.kB0 = d.kstar = 0;
            d.q = 0;
            d.i = (uint8_t)i;
            d.Tcur = 0.0;
            if (j - i >= 7) {
                const uint2 pk = ws.dpre[k];
                const unsigned flags = pk.x >> 24;
                // interior loops: byrow prefix of rows >= i+2, minus the 1x1 cell (its last entry)
                d.n_fast = (uint16_t)((flags & 1u) ? (int)ws.rowge[i + 2] - (int)((flags >> 1) & 1u) : 0);
                d.total = (uint16_t)(pk.x & 255u);
                d.nA = (uint16_t)((pk.x >> 8) & 255u);
                d.nB = (uint16_t)((pk.x >> 16) & 255u);
                d.kA0 = (uint16_t)(pk.y & 255u);
                d.kB0 = (uint16_t)((pk.y >> 8) & 255u);
                d.kstar = (uint16_t)((pk.y >> 16) & 255u);
                d.q = (uint8_t)(pk.y >> 24);
                d.Tcur = po_div_nr(v.x + iH, (v.y + iS) + RC);
            }
            if (act) {
                ws.cell[k] = v;
                ws.desc[lane] = d;
            }
        }
        __syncwarp();
        // phase B: CBI, 32/G cells of the column at a time.  Cells with j - i < 7 have no room
        // for an enclosed pair plus a loop; rows descend along the list, so they come first.
        const int kb = j >= 8 ? cs + popc_from32(HP32_COLM(ws, j), j - 7) : ce;
        const bool cAT = (bj == 0) | (bj == 3);   // the closing pair's AT penalty depends on either base
        for (int k0 = kb; k0 < ce; k0 += NG) {
            const bool valid = k0 + g.idx < ce;
            const int k = valid ? k0 + g.idx : ce - 1;
            const typename Hairpin32Ws<CAP>::Desc d = ws.desc[k - cs];
            Hairpin32Cell c;
            c.i = d.i;
            c.j = j;
            c.b1q = ws.s1[c.i] * 64 + bj * 4;
            c.kself = k;
            c.n_fast = valid ? (int)d.n_fast : 0;
            c.nA = d.nA;
            c.kA0 = d.kA0;
            c.nB = d.nB;
            c.kB0 = d.kB0;
            c.kstar = d.kstar;
            c.total = valid ? (int)d.total : 0;
            c.y0 = tb.tstack[d.q];
            double2* yc = ws.ycell[g.idx];
            yc[0] = c.y0;
            yc[1] = make_double2(cAT ? PO_AT_H : 0.0, cAT ? PO_AT_S : 0.0);
            yc[2] = make_double2(0.0, 0.0);
            yc[3] = tb.stackint2[d.q];
            const int trip_f = NG == 1 ? c.n_fast : warp_max_int(c.n_fast);
            const int trip_s = NG == 1 ? c.total : warp_max_int(c.total);
            __syncwarp();
            if (trip_f + trip_s > 0) {
                const double Tcur = d.Tcur;
                // passes: arg-max over all candidates; if it is the 1x1 loop, the running best at the
                // moment upstream visits it (DBL_EQ rule); if rejected, the arg-max without it
                int mode = SCAN_ALL, src = -1, wKey = KEY_NONE;
                double wT = -PO_INF;
                double2 wV = make_double2(PO_INF, -1.0);
                bool star = false, redo = false;
#pragma unroll 1
                for (int pass = 0; pass < 3; ++pass) {
                    double bT = -PO_INF, xT = -PO_INF;
                    int bKey = KEY_NONE, xKey = KEY_NONE;
                    double2 bV = make_double2(PO_INF, -1.0);
                    hairpin32_scan<G, CAP>(g, A, ws, c, trip_f, trip_s, max_loop, iH, iS, RC, mode, bT, bKey, bV);
                    const int s = g.argmax(bT, bKey, &xT, &xKey);
                    if (pass == 0) {
                        src = s;
                        wT = xT;
                        wKey = xKey;
                        wV = bV;
                        star = s >= 0 && xKey == KEY_STAR;
                        if (!__any_sync(FULL, star)) break;
                        mode = SCAN_BEFORE_STAR;
                    } else if (pass == 1) {
                        const double Tm = (s >= 0 && xT > Tcur) ? xT : Tcur;
                        redo = star && !((wT - Tm) >= PO_SMALL_NON_ZERO && wT > Tm);
                        if (!__any_sync(FULL, redo)) break;
                        mode = SCAN_NO_STAR;
                    } else if (redo) {
                        src = s;
                        wT = xT;
                        wKey = xKey;
                        wV = bV;
                    }
                }
                double H = g.shfl(wV.x, max(src, 0)), S = g.shfl(wV.y, max(src, 0));
                if (valid && src >= 0 && wT > Tcur && fin(H)) {
                    if (S < PO_MIN_ENTROPY_CUTOFF) {
                        S = PO_MIN_ENTROPY;
                        H = 0.0;
                    }
                    ws.cell[k] = make_double2(H, S);  // the group's lanes store the same value
                }
            }
            __syncwarp();  // ycell is rewritten by the next batch
        }
        // phase C: calc_hairpin, keep the lower dG of {loop closure, current value}
        {
            const bool act = cs + lane < ce;
            const int k = act ? cs + lane : ce - 1;
            const int i = CODE_I(ws.code[k]);
            const double2 cur = ws.cell[k];
            double2 v = hairpin_closure_vs(closure, cur);
            const double2 sh = rshT[END_IDX(ws.s1[i], ws.s1[i + 1], ws.s1[j + 1])];
            const double G1 = v.x + sh.x - PO_TEMP_KELVIN * (v.y + sh.y);
            const double G2 = cur.x + sh.x - PO_TEMP_KELVIN * (cur.y + sh.y);
            if (G2 < G1) v = cur;
            if (fin(v.x)) {
                if (v.y < PO_MIN_ENTROPY_CUTOFF) v = make_double2(0.0, PO_MIN_ENTROPY);
                if (act) ws.cell[k] = v;
            }
        }
        __syncwarp();
    }

    if constexpr (PHASE == 1) {
        uint32_t* codes = reinterpret_cast<uint32_t*>(cells + CAP);
        for (int k = lane; k < ncell; k += 32) {
            cells[k] = ws.cell[k];
            codes[k] = ws.code[k];
        }
    } else {
        hairpin_finish<Hairpin32Ws<CAP>, false>(tb, X, A, gt, ws, len, K, r, out, nullptr, 0);
    }
}
