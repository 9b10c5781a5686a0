/*
 * po_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Scalar double-precision restatement of the thermodynamic routines the
 * reference (MirkoLedda/polyoligo) reaches through primer3-py 0.6.0:
 *   calcTm          <- reference src/polyoligo/lib_primer3.py:59, _lib_crispr.py:42
 *   calcHairpin     <- reference src/polyoligo/lib_primer3.py:60
 *   calcHomodimer   <- reference src/polyoligo/lib_primer3.py:61
 *   calcHeterodimer <- reference src/polyoligo/lib_primer3.py:158, _lib_kasp.py:100-105
 * The arithmetic itself lives in primer3-py's bundled libprimer3 (oligotm.c,
 * thal.c; pinned primer3-py==0.6.0 in reference setup.py:62), which is NOT
 * vendored under /root/reference and is not installable offline.  This file
 * restates the published upstream algorithm function by function (upstream
 * function names are kept in comments).
 *
 * PARITY STATUS: calcTm is pinned bit-exactly by the primer3-py README anchor
 * calcTm('GTAAAACGACGGCCAGT') = 49.16808228911765.  thal is PARITY UNPINNED
 * against real primer3: the reference's tests hold no numeric vectors and the
 * upstream parameter files are absent (see tools/gen_default_params.py for the
 * literature set used instead).  The README hairpin anchor is reproduced for
 * structure and dH; its dS/Tm depend on the missing upstream tables.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * arm may call this library.
 */
#ifndef PO_ORACLE_H
#define PO_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct po_oracle_params po_oracle_params;

/* thal alignment types (upstream thal.h: thal_alignment_type) */
enum { PO_ORACLE_ANY = 1, PO_ORACLE_END1 = 2, PO_ORACLE_END2 = 3, PO_ORACLE_HAIRPIN = 4 };

typedef struct {
    double mv;        /* monovalent cations, mM   (ThermoAnalysis default 50)  */
    double dv;        /* divalent cations, mM     (default 0)                  */
    double dntp;      /* dNTP, mM                 (default 0.8)                */
    double dna_conc;  /* oligo concentration, nM  (default 50)                 */
    double temp_k;    /* temperature for dG, K    (default 37 C = 310.15)      */
    int32_t max_loop; /* largest loop size        (default 30)                 */
    int32_t type;     /* 1 any, 4 hairpin (2/3 = END1/END2)                    */
} po_oracle_args;

typedef struct {
    double tm;          /* deg C; 0.0 when no structure                         */
    double dg;          /* cal/mol                                              */
    double dh;          /* cal/mol                                              */
    double ds;          /* cal/(mol K)                                          */
    int32_t structure_found;
    int32_t align_end_1;
    int32_t align_end_2;
    int32_t status;     /* 0 ok, 1 empty input, 2 too long, 3 bad type          */
    int64_t relaxations;/* DP relaxations performed (SURVEY 8d work unit)       */
} po_oracle_result;

/* dir == NULL -> the compiled-in default set (oracle/default_params.inc). */
po_oracle_params *po_oracle_params_load(const char *dir, char *err, int errlen);
void po_oracle_params_free(po_oracle_params *p);
/* copy table t (0 stack,1 stackint2,2 tstack,3 tstack2: 625 each; 4 dangle3,
 * 5 dangle5: 125 each; 6 interior,7 bulge,8 hairpin: 30 each) */
int po_oracle_params_get(const po_oracle_params *p, int table, double *dh, double *ds);

void po_oracle_default_args(po_oracle_args *a, int type);

/* upstream oligotm.c: seqtm() with santalucia tm + santalucia salt correction */
double po_oracle_seqtm(const char *seq, double dna_conc, double mv, double dv,
                       double dntp, int nn_max_len);

/* upstream thal.c: thal() */
void po_oracle_thal(const po_oracle_params *p, const char *s1, const char *s2,
                    const po_oracle_args *a, po_oracle_result *o);

/* Batch drivers: sequences are one ASCII blob + n+1 offsets; OpenMP over items. */
void po_oracle_tm_batch(const char *blob, const int64_t *off, int64_t n,
                        double dna_conc, double mv, double dv, double dntp,
                        int nn_max_len, double *tm, int nthreads);
void po_oracle_thal_batch(const po_oracle_params *p, const po_oracle_args *a,
                          const char *blob1, const int64_t *off1,
                          const char *blob2, const int64_t *off2, /* NULL,NULL: s2 = s1 */
                          int64_t n, po_oracle_result *out, int nthreads);
int po_oracle_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
