// po_common.h -- shared host/device declarations of the polyoligo_b200 CUDA library.
//
// Thermodynamic model: libprimer3 `oligotm` / `thal` as called by polyoligo
// through primer3-py (reference src/polyoligo/lib_primer3.py:56-64,147-164).
// Parameter tables are runtime DATA in the upstream primer3_config format;
// the device copy keeps {dH, dS} interleaved as double2 so that one 16-byte
// shared-memory load fetches both numbers of a nearest-neighbour term.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/polyoligo_b200.h"

#define PO_WARPS_PER_CTA 8
#define PO_CTA_THREADS (PO_WARPS_PER_CTA * 32)
#define PO_MAX_BONUS 256  // capacity of the triloop / tetraloop lists

// upstream thal.c constants
#define PO_MIN_ENTROPY_CUTOFF (-2500.0)
#define PO_MIN_ENTROPY (-3224.0)
#define PO_AT_H 2200.0
#define PO_AT_S 6.9
#define PO_TEMP_KELVIN 310.15
#define PO_ABSOLUTE_ZERO 273.15
#define PO_SMALL_NON_ZERO 0.000001
#define PO_MIN_HRPN_LOOP 3

// 4-letter tables, {x = dH, y = dS}.  The sentinel N that thal puts before and
// after each sequence is handled by rule in the lookup helpers (po_thal.cuh),
// exactly as upstream's table loaders hard-code it.
struct PoTables {
    double2 stack[256];      // [a][b][c][d]: 5'-ab-3' / 3'-cd-5'
    double2 stackint2[256];  // single internal mismatch
    double2 tstack[256];     // interior-loop terminal mismatch
    double2 tstack2[256];    // duplex-end / hairpin terminal mismatch
    double2 dangle3[64];     // [a][x][c]
    double2 dangle5[64];     // [a][c][y]
    double2 interior[30];
    double2 bulge[30];
    double2 hairpin[30];
    int32_t n_tri, n_tetra;
    int32_t tri_key[PO_MAX_BONUS];    // 2 bits per base, first base in the low bits
    int32_t tetra_key[PO_MAX_BONUS];  // both lists sorted ascending by key
    double2 tri_val[PO_MAX_BONUS];
    double2 tetra_val[PO_MAX_BONUS];
};
// portion staged into shared memory by every CTA (everything before n_tri)
#define PO_TABLES_SMEM_BYTES (sizeof(double2) * (4 * 256 + 2 * 64 + 3 * 30))

// Conditions resolved on the host so that the device never calls log():
// the oracle and libprimer3 evaluate these with the C library's log().
struct PoConsts {
    double salt_corr;     // thal saltCorrectS()
    double rc_sym;        // R ln(C / 1e9)
    double rc_nonsym;     // R ln(C / 4e9)
    double temp_k;        // thal_args.temp
    double tm_log_mv;     // oligotm: ln(K / 1000)
    double tm_rlog_sym;   // oligotm: 1.987 ln(C / 1e9)
    double tm_rlog_non;   // oligotm: 1.987 ln(C / 4e9)
    double tm_log10_mv;   // long_seq_tm: log10(K / 1000)
    int32_t max_loop;
    int32_t max_nn_length;
};

// Host-side parameter loading (po_tables.cpp)
bool po_load_tables(const char* dir_or_null, PoTables* out, std::string* err);
