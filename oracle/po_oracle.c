/*
 * po_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See po_oracle.h for scope, provenance and parity status.
 *
 * Structure follows upstream primer3 (libprimer3 as bundled by primer3-py
 * 0.6.0; not present under /root/reference): oligotm.c:seqtm/oligotm and
 * thal.c:thal with its helpers.  Upstream keeps the DP state in file-static
 * globals; here it lives in a per-call context so batches can run on all host
 * cores (the reference gets its parallelism from one process per marker,
 * reference src/polyoligo/cli_kasp.py:354).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).
 */
#include "po_oracle.h"

#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* constants (upstream thal.c)                                               */
/* ------------------------------------------------------------------------- */
#define SMALL_NON_ZERO 0.000001
#define DBL_EQ(X, Y) (((X) - (Y)) < (SMALL_NON_ZERO) ? (1) : (2)) /* 1 when "equal" */
#define MIN_HRPN_LOOP 3
#define THAL_MAX_ALIGN 60
#define THAL_MAX_SEQ 10000

static const double PINF = INFINITY;
static const double R_GAS = 1.9872;            /* cal/(K mol) */
static const double ILAS = (-300 / 310.15);    /* internal loop entropy asymmetry */
static const double ILAH = 0.0;                /* internal loop enthalpy asymmetry */
static const double AT_H = 2200.0;             /* terminal AT penalty */
static const double AT_S = 6.9;
static const double MinEntropyCutoff = -2500.0;
static const double MinEntropy = -3224.0;
static const double G2_LIMIT = 0.0;            /* upstream G2: structures with higher G are unstable */
static const double ABSOLUTE_ZERO = 273.15;
static const double TEMP_KELVIN = 310.15;

#define isFinite(x) isfinite(x)
#define isPositive(x) ((x) > 0 ? (1) : (0))

struct loop_bonus {
    unsigned char loop[6];
    double value;
};

struct po_oracle_params {
    double stackS[5][5][5][5], stackH[5][5][5][5];
    double stackint2S[5][5][5][5], stackint2H[5][5][5][5];
    double tstackS[5][5][5][5], tstackH[5][5][5][5];
    double tstack2S[5][5][5][5], tstack2H[5][5][5][5];
    double dangle3S[5][5][5], dangle3H[5][5][5];
    double dangle5S[5][5][5], dangle5H[5][5][5];
    double interiorS[30], interiorH[30];
    double bulgeS[30], bulgeH[30];
    double hairpinS[30], hairpinH[30];
    struct loop_bonus *triS, *triH, *tetraS, *tetraH;
    int numTri, numTetra;
};

/* ------------------------------------------------------------------------- */
/* parameter files (upstream thal.c: get_thermodynamic_values and friends)   */
/* ------------------------------------------------------------------------- */
struct embedded_file {
    const char *name;
    const char *text;
};
static const struct embedded_file EMBEDDED[] = {
#include "default_params.inc"
    {NULL, NULL}};

static char *slurp(const char *dir, const char *name, char *err, int errlen)
{
    if (dir == NULL) {
        for (const struct embedded_file *e = EMBEDDED; e->name; ++e)
            if (strcmp(e->name, name) == 0) {
                char *c = (char *)malloc(strlen(e->text) + 1);
                strcpy(c, e->text);
                return c;
            }
        snprintf(err, errlen, "no embedded parameter file %s", name);
        return NULL;
    }
    char path[4096];
    snprintf(path, sizeof path, "%s/%s", dir, name);
    FILE *f = fopen(path, "rb");
    if (!f) {
        snprintf(err, errlen, "cannot open %s", path);
        return NULL;
    }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    char *c = (char *)malloc((size_t)n + 1);
    if (fread(c, 1, (size_t)n, f) != (size_t)n) n = 0;
    c[n] = 0;
    fclose(f);
    return c;
}

/* upstream readDouble: next whitespace-delimited token, "inf" -> +infinity */
static double read_double(char **pt, int *ok)
{
    char *p = *pt;
    while (*p && isspace((unsigned char)*p)) ++p;
    if (!*p) {
        *ok = 0;
        return PINF;
    }
    char *q = p;
    while (*q && !isspace((unsigned char)*q)) ++q;
    char save = *q;
    *q = 0;
    double v;
    if (strcmp(p, "inf") == 0) v = PINF;
    else v = strtod(p, NULL);
    *q = save;
    *pt = q;
    return v;
}

static int base_code(int c) /* upstream str2int */
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    }
    return 4;
}

/* upstream getStack / getStackint2: 256 values in i, ii, j, jj order; any N -> (inf,-1) */
static int load_quad_strict(const char *dir, const char *fs, const char *fh,
                            double S[5][5][5][5], double H[5][5][5][5], char *err, int errlen)
{
    char *ts = slurp(dir, fs, err, errlen), *th = slurp(dir, fh, err, errlen);
    if (!ts || !th) { free(ts); free(th); return -1; }
    char *ps = ts, *ph = th;
    int ok = 1;
    for (int i = 0; i < 5; ++i)
        for (int ii = 0; ii < 5; ++ii)
            for (int j = 0; j < 5; ++j)
                for (int jj = 0; jj < 5; ++jj) {
                    if (i == 4 || j == 4 || ii == 4 || jj == 4) {
                        S[i][ii][j][jj] = -1.0;
                        H[i][ii][j][jj] = PINF;
                    } else {
                        S[i][ii][j][jj] = read_double(&ps, &ok);
                        H[i][ii][j][jj] = read_double(&ph, &ok);
                        if (!isFinite(S[i][ii][j][jj]) || !isFinite(H[i][ii][j][jj])) {
                            S[i][ii][j][jj] = -1.0;
                            H[i][ii][j][jj] = PINF;
                        }
                    }
                }
    free(ts); free(th);
    if (!ok) { snprintf(err, errlen, "short parameter file %s/%s", fs, fh); return -1; }
    return 0;
}

/* upstream getTstack / getTstack2: N in the first pair -> (inf,-1); N in the
 * second position -> (H 0, S 1e-11) */
static int load_quad_terminal(const char *dir, const char *fs, const char *fh,
                              double S[5][5][5][5], double H[5][5][5][5], char *err, int errlen)
{
    char *ts = slurp(dir, fs, err, errlen), *th = slurp(dir, fh, err, errlen);
    if (!ts || !th) { free(ts); free(th); return -1; }
    char *ps = ts, *ph = th;
    int ok = 1;
    for (int i1 = 0; i1 < 5; ++i1)
        for (int i2 = 0; i2 < 5; ++i2)
            for (int j1 = 0; j1 < 5; ++j1)
                for (int j2 = 0; j2 < 5; ++j2) {
                    if (i1 == 4 || j1 == 4) {
                        H[i1][i2][j1][j2] = PINF;
                        S[i1][i2][j1][j2] = -1.0;
                    } else if (i2 == 4 || j2 == 4) {
                        S[i1][i2][j1][j2] = 0.00000000001;
                        H[i1][i2][j1][j2] = 0.0;
                    } else {
                        S[i1][i2][j1][j2] = read_double(&ps, &ok);
                        H[i1][i2][j1][j2] = read_double(&ph, &ok);
                        if (!isFinite(S[i1][i2][j1][j2]) || !isFinite(H[i1][i2][j1][j2])) {
                            S[i1][i2][j1][j2] = -1.0;
                            H[i1][i2][j1][j2] = PINF;
                        }
                    }
                }
    free(ts); free(th);
    if (!ok) { snprintf(err, errlen, "short parameter file %s/%s", fs, fh); return -1; }
    return 0;
}

/* upstream getDangle: dangle3 stored [i][k][j] while read in i, j, k order */
static int load_dangle(const char *dir, po_oracle_params *P, char *err, int errlen)
{
    char *ts = slurp(dir, "dangle.ds", err, errlen), *th = slurp(dir, "dangle.dh", err, errlen);
    if (!ts || !th) { free(ts); free(th); return -1; }
    char *ps = ts, *ph = th;
    int ok = 1;
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 5; ++j)
            for (int k = 0; k < 5; ++k) {
                if (i == 4 || j == 4 || k == 4) {
                    P->dangle3S[i][k][j] = -1.0;
                    P->dangle3H[i][k][j] = PINF;
                } else {
                    P->dangle3S[i][k][j] = read_double(&ps, &ok);
                    P->dangle3H[i][k][j] = read_double(&ph, &ok);
                    if (!isFinite(P->dangle3S[i][k][j]) || !isFinite(P->dangle3H[i][k][j])) {
                        P->dangle3S[i][k][j] = -1.0;
                        P->dangle3H[i][k][j] = PINF;
                    }
                }
            }
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 5; ++j)
            for (int k = 0; k < 5; ++k) {
                if (i == 4 || j == 4 || k == 4) {
                    P->dangle5S[i][j][k] = -1.0;
                    P->dangle5H[i][j][k] = PINF;
                } else {
                    P->dangle5S[i][j][k] = read_double(&ps, &ok);
                    P->dangle5H[i][j][k] = read_double(&ph, &ok);
                    if (!isFinite(P->dangle5S[i][j][k]) || !isFinite(P->dangle5H[i][j][k])) {
                        P->dangle5S[i][j][k] = -1.0;
                        P->dangle5H[i][j][k] = PINF;
                    }
                }
            }
    free(ts); free(th);
    if (!ok) { snprintf(err, errlen, "short parameter file dangle"); return -1; }
    return 0;
}

/* upstream getLoop/readLoop: "size interior bulge hairpin" per line, 30 lines */
static int load_loops(const char *dir, po_oracle_params *P, char *err, int errlen)
{
    char *ts = slurp(dir, "loops.ds", err, errlen), *th = slurp(dir, "loops.dh", err, errlen);
    if (!ts || !th) { free(ts); free(th); return -1; }
    char *ps = ts, *ph = th;
    int ok = 1;
    for (int k = 0; k < 30; ++k) {
        (void)read_double(&ps, &ok);
        P->interiorS[k] = read_double(&ps, &ok);
        P->bulgeS[k] = read_double(&ps, &ok);
        P->hairpinS[k] = read_double(&ps, &ok);
        (void)read_double(&ph, &ok);
        P->interiorH[k] = read_double(&ph, &ok);
        P->bulgeH[k] = read_double(&ph, &ok);
        P->hairpinH[k] = read_double(&ph, &ok);
    }
    free(ts); free(th);
    if (!ok) { snprintf(err, errlen, "short parameter file loops"); return -1; }
    return 0;
}

static int cmp_bonus5(const void *a, const void *b) { return memcmp(a, b, 5); }
static int cmp_bonus6(const void *a, const void *b) { return memcmp(a, b, 6); }

/* upstream getTriloop/getTetraloop: "SEQ value" per line; kept sorted for bsearch */
static int load_bonus(const char *dir, const char *name, int k, struct loop_bonus **out,
                      int *n_out, char *err, int errlen)
{
    char *t = slurp(dir, name, err, errlen);
    if (!t) return -1;
    int cap = 16, n = 0;
    struct loop_bonus *v = (struct loop_bonus *)malloc(sizeof(*v) * cap);
    char *p = t;
    for (;;) {
        while (*p && isspace((unsigned char)*p)) ++p;
        if (!*p) break;
        char *q = p;
        while (*q && !isspace((unsigned char)*q)) ++q;
        if ((int)(q - p) != k) { free(v); free(t); snprintf(err, errlen, "bad key in %s", name); return -1; }
        if (n == cap) { cap *= 2; v = (struct loop_bonus *)realloc(v, sizeof(*v) * cap); }
        memset(v[n].loop, 0, 6);
        for (int c = 0; c < k; ++c) v[n].loop[c] = (unsigned char)base_code(p[c]);
        int ok = 1;
        p = q;
        v[n].value = read_double(&p, &ok);
        if (!ok) { free(v); free(t); snprintf(err, errlen, "missing value in %s", name); return -1; }
        ++n;
    }
    qsort(v, n, sizeof(*v), k == 5 ? cmp_bonus5 : cmp_bonus6);
    free(t);
    *out = v;
    *n_out = n;
    return 0;
}

po_oracle_params *po_oracle_params_load(const char *dir, char *err, int errlen)
{
    char local[256];
    if (!err) { err = local; errlen = sizeof local; }
    err[0] = 0;
    po_oracle_params *P = (po_oracle_params *)calloc(1, sizeof(*P));
    int n1 = 0, n2 = 0;
    if (load_quad_strict(dir, "stack.ds", "stack.dh", P->stackS, P->stackH, err, errlen) ||
        load_quad_strict(dir, "stackmm.ds", "stackmm.dh", P->stackint2S, P->stackint2H, err, errlen) ||
        load_quad_terminal(dir, "tstack_tm_inf.ds", "tstack.dh", P->tstackS, P->tstackH, err, errlen) ||
        load_quad_terminal(dir, "tstack2.ds", "tstack2.dh", P->tstack2S, P->tstack2H, err, errlen) ||
        load_dangle(dir, P, err, errlen) || load_loops(dir, P, err, errlen) ||
        load_bonus(dir, "triloop.ds", 5, &P->triS, &n1, err, errlen) ||
        load_bonus(dir, "triloop.dh", 5, &P->triH, &n2, err, errlen)) {
        po_oracle_params_free(P);
        return NULL;
    }
    if (n1 != n2) { snprintf(err, errlen, "triloop.ds/.dh differ in length"); po_oracle_params_free(P); return NULL; }
    P->numTri = n1;
    if (load_bonus(dir, "tetraloop.ds", 6, &P->tetraS, &n1, err, errlen) ||
        load_bonus(dir, "tetraloop.dh", 6, &P->tetraH, &n2, err, errlen)) {
        po_oracle_params_free(P);
        return NULL;
    }
    if (n1 != n2) { snprintf(err, errlen, "tetraloop.ds/.dh differ in length"); po_oracle_params_free(P); return NULL; }
    P->numTetra = n1;
    return P;
}

void po_oracle_params_free(po_oracle_params *P)
{
    if (!P) return;
    free(P->triS); free(P->triH); free(P->tetraS); free(P->tetraH);
    free(P);
}

int po_oracle_params_get(const po_oracle_params *P, int table, double *dh, double *ds)
{
    const double *h = NULL, *s = NULL;
    int n = 0;
    switch (table) {
    case 0: h = &P->stackH[0][0][0][0]; s = &P->stackS[0][0][0][0]; n = 625; break;
    case 1: h = &P->stackint2H[0][0][0][0]; s = &P->stackint2S[0][0][0][0]; n = 625; break;
    case 2: h = &P->tstackH[0][0][0][0]; s = &P->tstackS[0][0][0][0]; n = 625; break;
    case 3: h = &P->tstack2H[0][0][0][0]; s = &P->tstack2S[0][0][0][0]; n = 625; break;
    case 4: h = &P->dangle3H[0][0][0]; s = &P->dangle3S[0][0][0]; n = 125; break;
    case 5: h = &P->dangle5H[0][0][0]; s = &P->dangle5S[0][0][0]; n = 125; break;
    case 6: h = P->interiorH; s = P->interiorS; n = 30; break;
    case 7: h = P->bulgeH; s = P->bulgeS; n = 30; break;
    case 8: h = P->hairpinH; s = P->hairpinS; n = 30; break;
    default: return -1;
    }
    memcpy(dh, h, sizeof(double) * n);
    memcpy(ds, s, sizeof(double) * n);
    return n;
}

void po_oracle_default_args(po_oracle_args *a, int type)
{
    /* primer3-py 0.6.0 ThermoAnalysis() defaults; the reference never overrides
     * them (reference src/polyoligo/lib_primer3.py:58,157) */
    a->mv = 50.0;
    a->dv = 0.0;
    a->dntp = 0.8;
    a->dna_conc = 50.0;
    a->temp_k = 37.0 + 273.15;
    a->max_loop = 30;
    a->type = type;
}

/* ------------------------------------------------------------------------- */
/* oligotm (upstream oligotm.c)                                              */
/* ------------------------------------------------------------------------- */
/* SantaLucia (1998) nearest-neighbour sums, sign-flipped integers:
 * dH in 100 cal/mol, dS in 0.1 cal/(K mol); index = 4*first + second, ACGT. */
static const int TM_DH[16] = {79, 84, 78, 72, 85, 80, 106, 78, 82, 98, 80, 84, 72, 82, 85, 79};
static const int TM_DS[16] = {222, 224, 210, 204, 227, 199, 272, 210, 222, 244, 199, 224, 213, 222, 227, 222};

/* upstream oligotm.c: symmetry() -- even length and self reverse-complementary */
static int tm_symmetry(const char *seq, int len)
{
    if (len % 2 == 1) return 0;
    for (int i = 0; i < len / 2; ++i) {
        int s = toupper((unsigned char)seq[i]), e = toupper((unsigned char)seq[len - 1 - i]);
        if ((s == 'A' && e != 'T') || (s == 'T' && e != 'A') || (e == 'A' && s != 'T') || (e == 'T' && s != 'A')) return 0;
        if ((s == 'C' && e != 'G') || (s == 'G' && e != 'C') || (e == 'C' && s != 'G') || (e == 'G' && s != 'C')) return 0;
    }
    return 1;
}

#define OLIGOTM_ERROR -999999.9999

/* upstream oligotm.c: divalent_to_monovalent() */
static double divalent_to_monovalent(double dv, double dntp)
{
    if (dv == 0) dntp = 0;
    if (dv < 0 || dntp < 0) return OLIGOTM_ERROR;
    if (dv < dntp) dv = dntp;
    return 120 * (sqrt(dv - dntp));
}

double po_oracle_seqtm(const char *seq, double dna_conc, double mv, double dv,
                       double dntp, int nn_max_len)
{
    int len = (int)strlen(seq);
    double d2m = divalent_to_monovalent(dv, dntp);
    if (d2m == OLIGOTM_ERROR) return OLIGOTM_ERROR;
    mv = mv + d2m;
    if (len < 2) return OLIGOTM_ERROR;
    if (len > nn_max_len) {
        /* upstream long_seq_tm(): GC-content formula */
        int gc = 0;
        for (int i = 0; i < len; ++i) {
            int c = toupper((unsigned char)seq[i]);
            if (c == 'G' || c == 'C') ++gc;
            else if (c != 'A' && c != 'T' && c != 'N') return OLIGOTM_ERROR;
        }
        return 81.5 + (16.6 * log10(mv / 1000.0)) + (41.0 * (((double)gc) / len)) - (600.0 / len);
    }
    /* upstream oligotm() with tm_method = santalucia_auto, salt = santalucia */
    int dh = 0, ds = 0;
    int sym = tm_symmetry(seq, len);
    if (sym) ds += 14;
    int first = toupper((unsigned char)seq[0]), last = toupper((unsigned char)seq[len - 1]);
    if (first == 'A' || first == 'T') { ds += -41; dh += -23; }
    else if (first == 'C' || first == 'G') { ds += 28; dh += -1; }
    if (last == 'A' || last == 'T') { ds += -41; dh += -23; }
    else if (last == 'C' || last == 'G') { ds += 28; dh += -1; }
    for (int i = 0; i + 1 < len; ++i) {
        int a = base_code(seq[i]), b = base_code(seq[i + 1]);
        /* ORACLE LIMIT: upstream carries extra table rows for 'N'; the reference
         * rejects every oligo containing N regardless of its Tm (reference
         * src/polyoligo/lib_primer3.py:72-74) so they are not restated. */
        if (a > 3 || b > 3) return OLIGOTM_ERROR;
        dh += TM_DH[4 * a + b];
        ds += TM_DS[4 * a + b];
    }
    double delta_H = dh * -100.0;
    double delta_S = ds * -0.1;
    delta_S = delta_S + 0.368 * (len - 1) * log(mv / 1000.0);
    double Tm;
    if (sym) Tm = delta_H / (delta_S + 1.987 * log(dna_conc / 1000000000.0)) - 273.15;
    else Tm = delta_H / (delta_S + 1.987 * log(dna_conc / 4000000000.0)) - 273.15;
    return Tm;
}

/* ------------------------------------------------------------------------- */
/* thal (upstream thal.c)                                                    */
/* ------------------------------------------------------------------------- */
typedef struct {
    const po_oracle_params *P;
    int len1, len2, stride;
    unsigned char *numSeq1, *numSeq2;
    double *enthalpyDPT, *entropyDPT;
    double *send5, *hend5;
    double dplx_init_H, dplx_init_S, RC, saltCorrection;
    int64_t relax;
} thal_ctx;

#define EnthalpyDPT(i, j) (c->enthalpyDPT[(i) * c->stride + (j)])
#define EntropyDPT(i, j) (c->entropyDPT[(i) * c->stride + (j)])
#define SEND5(i) (c->send5[i])
#define HEND5(i) (c->hend5[i])
#define numSeq1 (c->numSeq1)
#define numSeq2 (c->numSeq2)
#define PAR (c->P)

static const int BPI[5][5] = {
    {0, 0, 0, 1, 0}, {0, 0, 1, 0, 0}, {0, 1, 0, 0, 0}, {1, 0, 0, 0, 0}, {0, 0, 0, 0, 0}};
#define bpIndx(a, b) BPI[a][b]

static double atPenaltyS(int a, int b) { return ((a == 0 && b == 3) || (a == 3 && b == 0)) ? AT_S : 0.0; }
static double atPenaltyH(int a, int b) { return ((a == 0 && b == 3) || (a == 3 && b == 0)) ? AT_H : 0.0; }

/* upstream equal() */
static int equal(double a, double b)
{
    if (!isfinite(a) || !isfinite(b)) return 0;
    return fabs(a - b) < 1e-5;
}

/* upstream symmetry_thermo() */
static int symmetry_thermo(const char *seq)
{
    return tm_symmetry(seq, (int)strlen(seq));
}

/* upstream saltCorrectS() */
static double saltCorrectS(double mv, double dv, double dntp)
{
    if (dv <= 0) dntp = dv;
    return 0.368 * ((log((mv + 120 * (sqrt(fmax(0.0, dv - dntp)))) / 1000)));
}

/* upstream Ss()/Hs(): k==1 dimer orientation, k==2 hairpin orientation */
static double Ss(thal_ctx *c, int i, int j, int k)
{
    if (k == 2) {
        if (i >= j) return -1.0;
        if (i == c->len1 || j == c->len2 + 1) return -1.0;
        return PAR->stackS[numSeq1[i]][numSeq1[i + 1]][numSeq2[j]][numSeq2[j - 1]];
    }
    return PAR->stackS[numSeq1[i]][numSeq1[i + 1]][numSeq2[j]][numSeq2[j + 1]];
}
static double Hs(thal_ctx *c, int i, int j, int k)
{
    if (k == 2) {
        if (i >= j) return PINF;
        if (i == c->len1 || j == c->len2 + 1) return PINF;
        return PAR->stackH[numSeq1[i]][numSeq1[i + 1]][numSeq2[j]][numSeq2[j - 1]];
    }
    return PAR->stackH[numSeq1[i]][numSeq1[i + 1]][numSeq2[j]][numSeq2[j + 1]];
}

/* upstream initMatrix(): dimer DP table, finite iff bases are complementary */
static void initMatrix(thal_ctx *c)
{
    for (int i = 1; i <= c->len1; ++i)
        for (int j = 1; j <= c->len2; ++j) {
            if (bpIndx(numSeq1[i], numSeq2[j]) == 0) {
                EnthalpyDPT(i, j) = PINF;
                EntropyDPT(i, j) = -1.0;
            } else {
                EnthalpyDPT(i, j) = 0.0;
                EntropyDPT(i, j) = MinEntropy;
            }
        }
}

/* upstream initMatrix2(): hairpin DP table */
static void initMatrix2(thal_ctx *c)
{
    for (int i = 1; i <= c->len1; ++i)
        for (int j = i; j <= c->len2; ++j) {
            if (j - i < MIN_HRPN_LOOP + 1 || (bpIndx(numSeq1[i], numSeq1[j]) == 0)) {
                EnthalpyDPT(i, j) = PINF;
                EntropyDPT(i, j) = -1.0;
            } else {
                EnthalpyDPT(i, j) = 0.0;
                EntropyDPT(i, j) = MinEntropy;
            }
        }
}

/* upstream LSH(): best left-end term of a duplex starting at (i,j) */
static void LSH(thal_ctx *c, int i, int j, double *EntropyEnthalpy)
{
    double S1, H1, T1, G1;
    double S2, H2, T2, G2;
    S1 = S2 = -1.0;
    H1 = H2 = -PINF;
    T1 = T2 = -PINF;
    if (bpIndx(numSeq1[i], numSeq2[j]) == 0) {
        EntropyDPT(i, j) = -1.0;
        EnthalpyDPT(i, j) = PINF;
        return;
    }
    const double iH = c->dplx_init_H, iS = c->dplx_init_S, RC = c->RC;
    S1 = atPenaltyS(numSeq1[i], numSeq2[j]) + PAR->tstack2S[numSeq2[j]][numSeq2[j - 1]][numSeq1[i]][numSeq1[i - 1]];
    H1 = atPenaltyH(numSeq1[i], numSeq2[j]) + PAR->tstack2H[numSeq2[j]][numSeq2[j - 1]][numSeq1[i]][numSeq1[i - 1]];
    G1 = H1 - TEMP_KELVIN * S1;
    if (!isFinite(H1) || G1 > 0) {
        H1 = PINF;
        S1 = -1.0;
        G1 = 1.0;
    }
    const double d3H = PAR->dangle3H[numSeq2[j]][numSeq2[j - 1]][numSeq1[i]];
    const double d3S = PAR->dangle3S[numSeq2[j]][numSeq2[j - 1]][numSeq1[i]];
    const double d5H = PAR->dangle5H[numSeq2[j]][numSeq1[i]][numSeq1[i - 1]];
    const double d5S = PAR->dangle5S[numSeq2[j]][numSeq1[i]][numSeq1[i - 1]];
    int have2 = 0;
    /* two dangling ends at the same end of the duplex / 3' only / 5' only */
    if ((bpIndx(numSeq1[i - 1], numSeq2[j - 1]) != 1) && isFinite(d3H) && isFinite(d5H)) {
        S2 = atPenaltyS(numSeq1[i], numSeq2[j]) + d3S + d5S;
        H2 = atPenaltyH(numSeq1[i], numSeq2[j]) + d3H + d5H;
        have2 = 1;
    } else if ((bpIndx(numSeq1[i - 1], numSeq2[j - 1]) != 1) && isFinite(d3H)) {
        S2 = atPenaltyS(numSeq1[i], numSeq2[j]) + d3S;
        H2 = atPenaltyH(numSeq1[i], numSeq2[j]) + d3H;
        have2 = 1;
    } else if ((bpIndx(numSeq1[i - 1], numSeq2[j - 1]) != 1) && isFinite(d5H)) {
        S2 = atPenaltyS(numSeq1[i], numSeq2[j]) + d5S;
        H2 = atPenaltyH(numSeq1[i], numSeq2[j]) + d5H;
        have2 = 1;
    }
    if (have2) { /* identical tail in all three upstream branches */
        G2 = H2 - TEMP_KELVIN * S2;
        if (!isFinite(H2) || G2 > 0) {
            H2 = PINF;
            S2 = -1.0;
            G2 = 1.0;
        }
        T2 = (H2 + iH) / (S2 + iS + RC);
        if (isFinite(H1) && G1 < 0) {
            T1 = (H1 + iH) / (S1 + iS + RC);
            if (T1 < T2 && G2 < 0) {
                S1 = S2;
                H1 = H2;
                T1 = T2;
            }
        } else if (G2 < 0) {
            S1 = S2;
            H1 = H2;
            T1 = T2;
        }
    }
    S2 = atPenaltyS(numSeq1[i], numSeq2[j]);
    H2 = atPenaltyH(numSeq1[i], numSeq2[j]);
    T2 = (H2 + iH) / (S2 + iS + RC);
    if (isFinite(H1)) {
        if (T1 < T2) {
            EntropyEnthalpy[0] = S2;
            EntropyEnthalpy[1] = H2;
        } else {
            EntropyEnthalpy[0] = S1;
            EntropyEnthalpy[1] = H1;
        }
    } else {
        EntropyEnthalpy[0] = S2;
        EntropyEnthalpy[1] = H2;
    }
}

/* upstream RSH(): mirror of LSH on the (i+1, j+1) side */
static void RSH(thal_ctx *c, int i, int j, double *EntropyEnthalpy)
{
    double G1, G2;
    double S1, S2;
    double H1, H2;
    double T1, T2;
    S1 = S2 = -1.0;
    H1 = H2 = PINF;
    T1 = T2 = -PINF;
    if (bpIndx(numSeq1[i], numSeq2[j]) == 0) {
        EntropyEnthalpy[0] = -1.0;
        EntropyEnthalpy[1] = PINF;
        return;
    }
    const double iH = c->dplx_init_H, iS = c->dplx_init_S, RC = c->RC;
    S1 = atPenaltyS(numSeq1[i], numSeq2[j]) + PAR->tstack2S[numSeq1[i]][numSeq1[i + 1]][numSeq2[j]][numSeq2[j + 1]];
    H1 = atPenaltyH(numSeq1[i], numSeq2[j]) + PAR->tstack2H[numSeq1[i]][numSeq1[i + 1]][numSeq2[j]][numSeq2[j + 1]];
    G1 = H1 - TEMP_KELVIN * S1;
    if (!isFinite(H1) || G1 > 0) {
        H1 = PINF;
        S1 = -1.0;
        G1 = 1.0;
    }
    const double d3H = PAR->dangle3H[numSeq1[i]][numSeq1[i + 1]][numSeq2[j]];
    const double d3S = PAR->dangle3S[numSeq1[i]][numSeq1[i + 1]][numSeq2[j]];
    const double d5H = PAR->dangle5H[numSeq1[i]][numSeq2[j]][numSeq2[j + 1]];
    const double d5S = PAR->dangle5S[numSeq1[i]][numSeq2[j]][numSeq2[j + 1]];
    int have2 = 0;
    if (bpIndx(numSeq1[i + 1], numSeq2[j + 1]) == 0 && isFinite(d3H) && isFinite(d5H)) {
        S2 = atPenaltyS(numSeq1[i], numSeq2[j]) + d3S + d5S;
        H2 = atPenaltyH(numSeq1[i], numSeq2[j]) + d3H + d5H;
        have2 = 1;
    } else if (bpIndx(numSeq1[i + 1], numSeq2[j + 1]) == 0 && isFinite(d3H)) {
        S2 = atPenaltyS(numSeq1[i], numSeq2[j]) + d3S;
        H2 = atPenaltyH(numSeq1[i], numSeq2[j]) + d3H;
        have2 = 1;
    } else if (bpIndx(numSeq1[i + 1], numSeq2[j + 1]) == 0 && isFinite(d5H)) {
        S2 = atPenaltyS(numSeq1[i], numSeq2[j]) + d5S;
        H2 = atPenaltyH(numSeq1[i], numSeq2[j]) + d5H;
        have2 = 1;
    }
    if (have2) {
        G2 = H2 - TEMP_KELVIN * S2;
        if (!isFinite(H2) || G2 > 0) {
            H2 = PINF;
            S2 = -1.0;
            G2 = 1.0;
        }
        T2 = (H2 + iH) / (S2 + iS + RC);
        if (isFinite(H1) && G1 < 0) {
            T1 = (H1 + iH) / (S1 + iS + RC);
            if (T1 < T2 && G2 < 0) {
                S1 = S2;
                H1 = H2;
                T1 = T2;
            }
        } else if (G2 < 0) {
            S1 = S2;
            H1 = H2;
            T1 = T2;
        }
    }
    S2 = atPenaltyS(numSeq1[i], numSeq2[j]);
    H2 = atPenaltyH(numSeq1[i], numSeq2[j]);
    T2 = (H2 + iH) / (S2 + iS + RC);
    if (isFinite(H1)) {
        if (T1 < T2) {
            EntropyEnthalpy[0] = S2;
            EntropyEnthalpy[1] = H2;
        } else {
            EntropyEnthalpy[0] = S1;
            EntropyEnthalpy[1] = H1;
        }
    } else {
        EntropyEnthalpy[0] = S2;
        EntropyEnthalpy[1] = H2;
    }
}

/* upstream maxTM(): stack (i-1,j-1)->(i,j) versus the current cell, judged
 * with the right-end term of (i,j) included */
static void maxTM(thal_ctx *c, int i, int j)
{
    double T0, T1;
    double S0, S1;
    double H0, H1;
    double SH[2];
    const double iH = c->dplx_init_H, iS = c->dplx_init_S, RC = c->RC;
    S0 = EntropyDPT(i, j);
    H0 = EnthalpyDPT(i, j);
    RSH(c, i, j, SH);
    T0 = (H0 + iH + SH[1]) / (S0 + iS + RC + SH[0]);
    if (isFinite(EnthalpyDPT(i - 1, j - 1)) && isFinite(Hs(c, i - 1, j - 1, 1))) {
        S1 = (EntropyDPT(i - 1, j - 1) + Ss(c, i - 1, j - 1, 1));
        H1 = (EnthalpyDPT(i - 1, j - 1) + Hs(c, i - 1, j - 1, 1));
        T1 = (H1 + iH + SH[1]) / (S1 + iS + RC + SH[0]);
    } else {
        S1 = -1.0;
        H1 = PINF;
        T1 = (H1 + iH) / (S1 + iS + RC);
    }
    if (S1 < MinEntropyCutoff) {
        S1 = MinEntropy;
        H1 = 0.0;
    }
    if (S0 < MinEntropyCutoff) {
        S0 = MinEntropy;
        H0 = 0.0;
    }
    if (T1 > T0) {
        EntropyDPT(i, j) = S1;
        EnthalpyDPT(i, j) = H1;
    } else if (T0 >= T1) {
        EntropyDPT(i, j) = S0;
        EnthalpyDPT(i, j) = H0;
    }
}

/* upstream calc_bulge_internal(): loop closed by (i,j) [earlier] and (ii,jj) [current cell] */
static void calc_bulge_internal(thal_ctx *c, int i, int j, int ii, int jj,
                                double *EntropyEnthalpy, int traceback)
{
    int loopSize1, loopSize2, loopSize;
    double T1, T2;
    double S, H;
    const double iH = c->dplx_init_H, iS = c->dplx_init_S, RC = c->RC;
    S = MinEntropy;
    H = 0;
    loopSize1 = ii - i - 1;
    loopSize2 = jj - j - 1;
    loopSize = loopSize1 + loopSize2 - 1;
    if ((loopSize1 == 0 && loopSize2 > 0) || (loopSize2 == 0 && loopSize1 > 0)) { /* bulges */
        if (loopSize2 == 1 || loopSize1 == 1) { /* size 1: the intervening nn-pair is added */
            H = PAR->bulgeH[loopSize] + PAR->stackH[numSeq1[i]][numSeq1[ii]][numSeq2[j]][numSeq2[jj]];
            S = PAR->bulgeS[loopSize] + PAR->stackS[numSeq1[i]][numSeq1[ii]][numSeq2[j]][numSeq2[jj]];
            if (isPositive(H) || isPositive(S)) {
                H = PINF;
                S = -1.0;
            }
            H += EnthalpyDPT(i, j);
            S += EntropyDPT(i, j);
            if (!isFinite(H)) {
                H = PINF;
                S = -1.0;
            }
            T1 = (H + iH) / ((S + iS) + RC);
            T2 = (EnthalpyDPT(ii, jj) + iH) / ((EntropyDPT(ii, jj)) + iS + RC);
            if ((T1 > T2) || ((traceback && T1 >= T2) || traceback == 1)) {
                EntropyEnthalpy[0] = S;
                EntropyEnthalpy[1] = H;
            }
        } else {
            H = PAR->bulgeH[loopSize] + atPenaltyH(numSeq1[i], numSeq2[j]) + atPenaltyH(numSeq1[ii], numSeq2[jj]);
            H += EnthalpyDPT(i, j);
            S = PAR->bulgeS[loopSize] + atPenaltyS(numSeq1[i], numSeq2[j]) + atPenaltyS(numSeq1[ii], numSeq2[jj]);
            S += EntropyDPT(i, j);
            if (!isFinite(H)) {
                H = PINF;
                S = -1.0;
            }
            if (isPositive(H) && isPositive(S)) {
                H = PINF;
                S = -1.0;
            }
            T1 = (H + iH) / ((S + iS) + RC);
            T2 = (EnthalpyDPT(ii, jj) + iH) / (EntropyDPT(ii, jj) + iS + RC);
            if ((T1 > T2) || ((traceback && T1 >= T2) || (traceback == 1))) {
                EntropyEnthalpy[0] = S;
                EntropyEnthalpy[1] = H;
            }
        }
    } else if (loopSize1 == 1 && loopSize2 == 1) {
        S = PAR->stackint2S[numSeq1[i]][numSeq1[i + 1]][numSeq2[j]][numSeq2[j + 1]] +
            PAR->stackint2S[numSeq2[jj]][numSeq2[jj - 1]][numSeq1[ii]][numSeq1[ii - 1]];
        S += EntropyDPT(i, j);
        H = PAR->stackint2H[numSeq1[i]][numSeq1[i + 1]][numSeq2[j]][numSeq2[j + 1]] +
            PAR->stackint2H[numSeq2[jj]][numSeq2[jj - 1]][numSeq1[ii]][numSeq1[ii - 1]];
        H += EnthalpyDPT(i, j);
        if (!isFinite(H)) {
            H = PINF;
            S = -1.0;
        }
        if (isPositive(H) && isPositive(S)) {
            H = PINF;
            S = -1.0;
        }
        T1 = (H + iH) / ((S + iS) + RC);
        T2 = (EnthalpyDPT(ii, jj) + iH) / (EntropyDPT(ii, jj) + iS + RC);
        if ((DBL_EQ(T1, T2) == 2) || traceback) {
            if ((T1 > T2) || (traceback && T1 >= T2)) {
                EntropyEnthalpy[0] = S;
                EntropyEnthalpy[1] = H;
            }
        }
        return;
    } else { /* interior loops */
        H = PAR->interiorH[loopSize] + PAR->tstackH[numSeq1[i]][numSeq1[i + 1]][numSeq2[j]][numSeq2[j + 1]] +
            PAR->tstackH[numSeq2[jj]][numSeq2[jj - 1]][numSeq1[ii]][numSeq1[ii - 1]] +
            (ILAH * abs(loopSize1 - loopSize2));
        H += EnthalpyDPT(i, j);
        S = PAR->interiorS[loopSize] + PAR->tstackS[numSeq1[i]][numSeq1[i + 1]][numSeq2[j]][numSeq2[j + 1]] +
            PAR->tstackS[numSeq2[jj]][numSeq2[jj - 1]][numSeq1[ii]][numSeq1[ii - 1]] +
            (ILAS * abs(loopSize1 - loopSize2));
        S += EntropyDPT(i, j);
        if (!isFinite(H)) {
            H = PINF;
            S = -1.0;
        }
        if (isPositive(H) && isPositive(S)) {
            H = PINF;
            S = -1.0;
        }
        T1 = (H + iH) / ((S + iS) + RC);
        T2 = (EnthalpyDPT(ii, jj) + iH) / ((EntropyDPT(ii, jj)) + iS + RC);
        if ((T1 > T2) || ((traceback && T1 >= T2) || (traceback == 1))) {
            EntropyEnthalpy[0] = S;
            EntropyEnthalpy[1] = H;
        }
    }
}

/* upstream fillMatrix(): dimer DP */
static void fillMatrix(thal_ctx *c, int maxLoop)
{
    int d, i, j, ii, jj;
    double SH[2];
    for (i = 1; i <= c->len1; ++i) {
        for (j = 1; j <= c->len2; ++j) {
            if (isFinite(EnthalpyDPT(i, j))) {
                SH[0] = -1.0;
                SH[1] = PINF;
                LSH(c, i, j, SH);
                c->relax += 1;
                if (isFinite(SH[1])) {
                    EntropyDPT(i, j) = SH[0];
                    EnthalpyDPT(i, j) = SH[1];
                }
                if (i > 1 && j > 1) {
                    maxTM(c, i, j);
                    c->relax += 1;
                    for (d = 3; d <= maxLoop + 2; d++) {
                        ii = i - 1;
                        jj = -ii - d + (j + i);
                        if (jj < 1) {
                            ii -= abs(jj - 1);
                            jj = 1;
                        }
                        for (; ii > 0 && jj < j; --ii, ++jj) {
                            if (isFinite(EnthalpyDPT(ii, jj))) {
                                SH[0] = -1.0;
                                SH[1] = PINF;
                                calc_bulge_internal(c, ii, jj, i, j, SH, 0);
                                c->relax += 1;
                                if (SH[0] < MinEntropyCutoff) {
                                    SH[0] = MinEntropy;
                                    SH[1] = 0.0;
                                }
                                if (isFinite(SH[1])) {
                                    EnthalpyDPT(i, j) = SH[1];
                                    EntropyDPT(i, j) = SH[0];
                                }
                            }
                        }
                    }
                }
            }
        }
    }
}

/* upstream traceback(): dimer */
static void traceback(thal_ctx *c, int i, int j, int *ps1, int *ps2, int maxLoop)
{
    int d, ii, jj, done;
    double SH[2];
    ps1[i - 1] = j;
    ps2[j - 1] = i;
    while (1) {
        SH[0] = -1.0;
        SH[1] = PINF;
        LSH(c, i, j, SH);
        if (equal(EntropyDPT(i, j), SH[0]) && equal(EnthalpyDPT(i, j), SH[1])) break;
        done = 0;
        if (i > 1 && j > 1 && equal(EntropyDPT(i, j), Ss(c, i - 1, j - 1, 1) + EntropyDPT(i - 1, j - 1)) &&
            equal(EnthalpyDPT(i, j), Hs(c, i - 1, j - 1, 1) + EnthalpyDPT(i - 1, j - 1))) {
            i = i - 1;
            j = j - 1;
            ps1[i - 1] = j;
            ps2[j - 1] = i;
            done = 1;
        }
        for (d = 3; !done && d <= maxLoop + 2; ++d) {
            ii = i - 1;
            jj = -ii - d + (j + i);
            if (jj < 1) {
                ii -= abs(jj - 1);
                jj = 1;
            }
            for (; !done && ii > 0 && jj < j; --ii, ++jj) {
                SH[0] = -1.0;
                SH[1] = PINF;
                calc_bulge_internal(c, ii, jj, i, j, SH, 1);
                if (equal(EntropyDPT(i, j), SH[0]) && equal(EnthalpyDPT(i, j), SH[1])) {
                    i = ii;
                    j = jj;
                    ps1[i - 1] = j;
                    ps2[j - 1] = i;
                    done = 1;
                    break;
                }
            }
        }
        /* ORACLE GUARD: upstream loops forever when no predecessor matches
         * (only reachable through the MinEntropyCutoff clamp, i.e. never for
         * oligos <= 60 nt); stop instead. */
        if (!done) break;
    }
}

/* ---- hairpin side -------------------------------------------------------- */

/* upstream maxTM2() */
static void maxTM2(thal_ctx *c, int i, int j)
{
    double T0, T1;
    double S0, S1;
    double H0, H1;
    const double iH = c->dplx_init_H, iS = c->dplx_init_S, RC = c->RC;
    S0 = EntropyDPT(i, j);
    H0 = EnthalpyDPT(i, j);
    T0 = (H0 + iH) / (S0 + iS + RC);
    if (isFinite(EnthalpyDPT(i, j))) {
        S1 = (EntropyDPT(i + 1, j - 1) + Ss(c, i, j, 2));
        H1 = (EnthalpyDPT(i + 1, j - 1) + Hs(c, i, j, 2));
    } else {
        S1 = -1.0;
        H1 = PINF;
    }
    T1 = (H1 + iH) / (S1 + iS + RC);
    if (S1 < MinEntropyCutoff) {
        S1 = MinEntropy;
        H1 = 0.0;
    }
    if (S0 < MinEntropyCutoff) {
        S0 = MinEntropy;
        H0 = 0.0;
    }
    if (T1 > T0) {
        EntropyDPT(i, j) = S1;
        EnthalpyDPT(i, j) = H1;
    } else {
        EntropyDPT(i, j) = S0;
        EnthalpyDPT(i, j) = H0;
    }
}

/* upstream calc_bulge_internal2(): (i,j) encloses (ii,jj) */
static void calc_bulge_internal2(thal_ctx *c, int i, int j, int ii, int jj,
                                 double *EntropyEnthalpy, int traceback, int maxLoop)
{
    int loopSize1, loopSize2, loopSize;
    double T1, T2;
    double S, H;
    const double iH = c->dplx_init_H, iS = c->dplx_init_S, RC = c->RC;
    S = MinEntropy;
    H = 0.0;
    loopSize1 = ii - i - 1;
    loopSize2 = j - jj - 1;
    if (loopSize1 + loopSize2 > maxLoop) {
        EntropyEnthalpy[0] = -1.0;
        EntropyEnthalpy[1] = PINF;
        return;
    }
    loopSize = loopSize1 + loopSize2 - 1;
    if ((loopSize1 == 0 && loopSize2 > 0) || (loopSize2 == 0 && loopSize1 > 0)) {
        if (loopSize2 == 1 || loopSize1 == 1) {
            H = PAR->bulgeH[loopSize] + PAR->stackH[numSeq1[i]][numSeq1[ii]][numSeq2[j]][numSeq2[jj]];
            S = PAR->bulgeS[loopSize] + PAR->stackS[numSeq1[i]][numSeq1[ii]][numSeq2[j]][numSeq2[jj]];
            if (traceback != 1) {
                H += EnthalpyDPT(ii, jj);
                S += EntropyDPT(ii, jj);
            }
            if (!isFinite(H)) {
                H = PINF;
                S = -1.0;
            }
            T1 = (H + iH) / ((S + iS) + RC);
            T2 = (EnthalpyDPT(i, j) + iH) / ((EntropyDPT(i, j)) + iS + RC);
            if ((T1 > T2) || ((traceback && T1 >= T2) || traceback == 1)) {
                EntropyEnthalpy[0] = S;
                EntropyEnthalpy[1] = H;
            }
        } else {
            H = PAR->bulgeH[loopSize] + atPenaltyH(numSeq1[i], numSeq2[j]) + atPenaltyH(numSeq1[ii], numSeq2[jj]);
            if (traceback != 1) H += EnthalpyDPT(ii, jj);
            S = PAR->bulgeS[loopSize] + atPenaltyS(numSeq1[i], numSeq2[j]) + atPenaltyS(numSeq1[ii], numSeq2[jj]);
            if (traceback != 1) S += EntropyDPT(ii, jj);
            if (!isFinite(H)) {
                H = PINF;
                S = -1.0;
            }
            T1 = (H + iH) / ((S + iS) + RC);
            T2 = (EnthalpyDPT(i, j) + iH) / (EntropyDPT(i, j) + iS + RC);
            if ((T1 > T2) || ((traceback && T1 >= T2) || (traceback == 1))) {
                EntropyEnthalpy[0] = S;
                EntropyEnthalpy[1] = H;
            }
        }
    } else if (loopSize1 == 1 && loopSize2 == 1) {
        S = PAR->stackint2S[numSeq1[i]][numSeq1[i + 1]][numSeq2[j]][numSeq2[j - 1]] +
            PAR->stackint2S[numSeq2[jj]][numSeq2[jj + 1]][numSeq1[ii]][numSeq1[ii - 1]];
        if (traceback != 1) S += EntropyDPT(ii, jj);
        H = PAR->stackint2H[numSeq1[i]][numSeq1[i + 1]][numSeq2[j]][numSeq2[j - 1]] +
            PAR->stackint2H[numSeq2[jj]][numSeq2[jj + 1]][numSeq1[ii]][numSeq1[ii - 1]];
        if (traceback != 1) H += EnthalpyDPT(ii, jj);
        if (!isFinite(H)) {
            H = PINF;
            S = -1.0;
        }
        T1 = (H + iH) / ((S + iS) + RC);
        T2 = (EnthalpyDPT(i, j) + iH) / (EntropyDPT(i, j) + iS + RC);
        if ((DBL_EQ(T1, T2) == 2) || traceback) {
            if ((T1 > T2) || ((traceback && T1 >= T2) || traceback == 1)) {
                EntropyEnthalpy[0] = S;
                EntropyEnthalpy[1] = H;
            }
        }
        return;
    } else {
        H = PAR->interiorH[loopSize] + PAR->tstackH[numSeq1[i]][numSeq1[i + 1]][numSeq2[j]][numSeq2[j - 1]] +
            PAR->tstackH[numSeq2[jj]][numSeq2[jj + 1]][numSeq1[ii]][numSeq1[ii - 1]] +
            (ILAH * abs(loopSize1 - loopSize2));
        if (traceback != 1) H += EnthalpyDPT(ii, jj);
        S = PAR->interiorS[loopSize] + PAR->tstackS[numSeq1[i]][numSeq1[i + 1]][numSeq2[j]][numSeq2[j - 1]] +
            PAR->tstackS[numSeq2[jj]][numSeq2[jj + 1]][numSeq1[ii]][numSeq1[ii - 1]] +
            (ILAS * abs(loopSize1 - loopSize2));
        if (traceback != 1) S += EntropyDPT(ii, jj);
        if (!isFinite(H)) {
            H = PINF;
            S = -1.0;
        }
        T1 = (H + iH) / ((S + iS) + RC);
        T2 = (EnthalpyDPT(i, j) + iH) / ((EntropyDPT(i, j)) + iS + RC);
        if ((T1 > T2) || ((traceback && T1 >= T2) || (traceback == 1))) {
            EntropyEnthalpy[0] = S;
            EntropyEnthalpy[1] = H;
        }
    }
}

/* upstream CBI(): bulges and interior loops enclosed by (i,j) */
static void CBI(thal_ctx *c, int i, int j, double *EntropyEnthalpy, int traceback, int maxLoop)
{
    int d, ii, jj;
    for (d = j - i - 3; d >= MIN_HRPN_LOOP + 1 && d >= j - i - 2 - maxLoop; --d)
        for (ii = i + 1; ii < j - d && ii <= c->len1; ++ii) {
            jj = d + ii;
            if (traceback == 0) {
                EntropyEnthalpy[0] = -1.0;
                EntropyEnthalpy[1] = PINF;
            }
            if (isFinite(EnthalpyDPT(ii, jj)) && isFinite(EnthalpyDPT(i, j))) {
                calc_bulge_internal2(c, i, j, ii, jj, EntropyEnthalpy, traceback, maxLoop);
                if (traceback == 0) c->relax += 1;
                if (isFinite(EntropyEnthalpy[1])) {
                    if (EntropyEnthalpy[0] < MinEntropyCutoff) {
                        EntropyEnthalpy[0] = MinEntropy;
                        EntropyEnthalpy[1] = 0.0;
                    }
                    if (traceback == 0) {
                        EnthalpyDPT(i, j) = EntropyEnthalpy[1];
                        EntropyDPT(i, j) = EntropyEnthalpy[0];
                    }
                }
            }
        }
}

/* upstream calc_hairpin(): hairpin loop closed by (i,j) */
static void calc_hairpin(thal_ctx *c, int i, int j, double *EntropyEnthalpy, int traceback)
{
    int loopSize = j - i - 1;
    double G1, G2;
    double SH[2];
    SH[0] = -1.0;
    SH[1] = PINF;
    if (loopSize < MIN_HRPN_LOOP) {
        EntropyEnthalpy[0] = -1.0;
        EntropyEnthalpy[1] = PINF;
        return;
    }
    if (i <= c->len1 && c->len2 < j) {
        EntropyEnthalpy[0] = -1.0;
        EntropyEnthalpy[1] = PINF;
        return;
    }
    if (loopSize <= 30) {
        EntropyEnthalpy[1] = PAR->hairpinH[loopSize - 1];
        EntropyEnthalpy[0] = PAR->hairpinS[loopSize - 1];
    } else {
        EntropyEnthalpy[1] = PAR->hairpinH[29];
        EntropyEnthalpy[0] = PAR->hairpinS[29];
    }
    if (loopSize > 3) { /* terminal mismatch */
        EntropyEnthalpy[1] += PAR->tstack2H[numSeq1[i]][numSeq1[i + 1]][numSeq1[j]][numSeq1[j - 1]];
        EntropyEnthalpy[0] += PAR->tstack2S[numSeq1[i]][numSeq1[i + 1]][numSeq1[j]][numSeq1[j - 1]];
    } else if (loopSize == 3) { /* closing AT penalty */
        EntropyEnthalpy[1] += atPenaltyH(numSeq1[i], numSeq1[j]);
        EntropyEnthalpy[0] += atPenaltyS(numSeq1[i], numSeq1[j]);
    }
    if (loopSize == 3) { /* triloop bonus */
        if (PAR->numTri) {
            const struct loop_bonus *b;
            if ((b = (const struct loop_bonus *)bsearch(numSeq1 + i, PAR->triH, PAR->numTri, sizeof(struct loop_bonus), cmp_bonus5)))
                EntropyEnthalpy[1] += b->value;
            if ((b = (const struct loop_bonus *)bsearch(numSeq1 + i, PAR->triS, PAR->numTri, sizeof(struct loop_bonus), cmp_bonus5)))
                EntropyEnthalpy[0] += b->value;
        }
    } else if (loopSize == 4) { /* tetraloop bonus */
        if (PAR->numTetra) {
            const struct loop_bonus *b;
            if ((b = (const struct loop_bonus *)bsearch(numSeq1 + i, PAR->tetraH, PAR->numTetra, sizeof(struct loop_bonus), cmp_bonus6)))
                EntropyEnthalpy[1] += b->value;
            if ((b = (const struct loop_bonus *)bsearch(numSeq1 + i, PAR->tetraS, PAR->numTetra, sizeof(struct loop_bonus), cmp_bonus6)))
                EntropyEnthalpy[0] += b->value;
        }
    }
    if (!isFinite(EntropyEnthalpy[1])) {
        EntropyEnthalpy[1] = PINF;
        EntropyEnthalpy[0] = -1.0;
    }
    if (isPositive(EntropyEnthalpy[1]) && isPositive(EntropyEnthalpy[0]) &&
        (!isPositive(EnthalpyDPT(i, j)) || !isPositive(EntropyDPT(i, j)))) {
        EntropyEnthalpy[1] = PINF;
        EntropyEnthalpy[0] = -1.0;
    }
    RSH(c, i, j, SH);
    G1 = EntropyEnthalpy[1] + SH[1] - TEMP_KELVIN * (EntropyEnthalpy[0] + SH[0]);
    G2 = EnthalpyDPT(i, j) + SH[1] - TEMP_KELVIN * (EntropyDPT(i, j) + SH[0]);
    if (G2 < G1 && traceback == 0) {
        EntropyEnthalpy[0] = EntropyDPT(i, j);
        EntropyEnthalpy[1] = EnthalpyDPT(i, j);
    }
}

/* upstream fillMatrix2(): hairpin DP */
static void fillMatrix2(thal_ctx *c, int maxLoop)
{
    int i, j;
    double SH[2];
    for (j = 2; j <= c->len2; ++j)
        for (i = j - MIN_HRPN_LOOP - 1; i >= 1; --i) {
            if (isFinite(EnthalpyDPT(i, j))) {
                SH[0] = -1.0;
                SH[1] = PINF;
                maxTM2(c, i, j);
                CBI(c, i, j, SH, 0, maxLoop);
                SH[0] = -1.0;
                SH[1] = PINF;
                calc_hairpin(c, i, j, SH, 0);
                c->relax += 2;
                if (isFinite(SH[1])) {
                    if (SH[0] < MinEntropyCutoff) {
                        SH[0] = MinEntropy;
                        SH[1] = 0.0;
                    }
                    EntropyDPT(i, j) = SH[0];
                    EnthalpyDPT(i, j) = SH[1];
                }
            }
        }
}

/* upstream Sd5/Hd5/Sd3/Hd3/Ststack/Htstack (hairpin exterior terms) */
static double Sd5(thal_ctx *c, int i, int j) { return PAR->dangle5S[numSeq1[i]][numSeq1[j]][numSeq1[j - 1]]; }
static double Hd5(thal_ctx *c, int i, int j) { return PAR->dangle5H[numSeq1[i]][numSeq1[j]][numSeq1[j - 1]]; }
static double Sd3(thal_ctx *c, int i, int j) { return PAR->dangle3S[numSeq1[i]][numSeq1[i + 1]][numSeq1[j]]; }
static double Hd3(thal_ctx *c, int i, int j) { return PAR->dangle3H[numSeq1[i]][numSeq1[i + 1]][numSeq1[j]]; }
static double Ststack(thal_ctx *c, int i, int j) { return PAR->tstack2S[numSeq1[i]][numSeq1[i + 1]][numSeq1[j]][numSeq1[j - 1]]; }
static double Htstack(thal_ctx *c, int i, int j) { return PAR->tstack2H[numSeq1[i]][numSeq1[i + 1]][numSeq1[j]][numSeq1[j - 1]]; }

/* upstream END5_1..END5_4 share one shape: scan k, pair (k+a, i-b) plus an
 * exterior term; variant 1 none, 2 5' dangle, 3 3' dangle, 4 terminal mismatch.
 * hs == 1 returns H, otherwise S. */
static double END5_x(thal_ctx *c, int variant, int i, int hs)
{
    int k;
    double max_tm;
    double T1, T2;
    double H, S;
    double H_max, S_max;
    const double iH = c->dplx_init_H, iS = c->dplx_init_S, RC = c->RC;
    H_max = H = PINF;
    S_max = S = -1.0;
    T1 = T2 = -PINF;
    max_tm = -PINF;
    int kmax, p, q; /* paired bases are (p, q) */
    switch (variant) {
    case 1: kmax = i - MIN_HRPN_LOOP - 2; break;
    case 2: kmax = i - MIN_HRPN_LOOP - 3; break;
    case 3: kmax = i - MIN_HRPN_LOOP - 3; break;
    default: kmax = i - MIN_HRPN_LOOP - 4; break;
    }
    for (k = 0; k <= kmax; ++k) {
        double xH, xS;
        switch (variant) {
        case 1: p = k + 1; q = i; xH = 0; xS = 0; break;
        case 2: p = k + 2; q = i; xH = Hd5(c, i, k + 2); xS = Sd5(c, i, k + 2); break;
        case 3: p = k + 1; q = i - 1; xH = Hd3(c, i - 1, k + 1); xS = Sd3(c, i - 1, k + 1); break;
        default: p = k + 2; q = i - 1; xH = Htstack(c, i - 1, k + 2); xS = Ststack(c, i - 1, k + 2); break;
        }
        T1 = (HEND5(k) + iH) / (SEND5(k) + iS + RC);
        T2 = (0 + iH) / (0 + iS + RC);
        if (T1 >= T2) {
            if (variant == 1) {
                H = HEND5(k) + atPenaltyH(numSeq1[p], numSeq1[q]) + EnthalpyDPT(p, q);
                S = SEND5(k) + atPenaltyS(numSeq1[p], numSeq1[q]) + EntropyDPT(p, q);
            } else {
                H = HEND5(k) + atPenaltyH(numSeq1[p], numSeq1[q]) + xH + EnthalpyDPT(p, q);
                S = SEND5(k) + atPenaltyS(numSeq1[p], numSeq1[q]) + xS + EntropyDPT(p, q);
            }
        } else {
            if (variant == 1) {
                H = 0 + atPenaltyH(numSeq1[p], numSeq1[q]) + EnthalpyDPT(p, q);
                S = 0 + atPenaltyS(numSeq1[p], numSeq1[q]) + EntropyDPT(p, q);
            } else {
                H = 0 + atPenaltyH(numSeq1[p], numSeq1[q]) + xH + EnthalpyDPT(p, q);
                S = 0 + atPenaltyS(numSeq1[p], numSeq1[q]) + xS + EntropyDPT(p, q);
            }
        }
        if (!isFinite(H) || H > 0 || S > 0) {
            H = PINF;
            S = -1.0;
        }
        T1 = (H + iH) / (S + iS + RC);
        if (max_tm < T1) {
            if (S > MinEntropyCutoff) {
                H_max = H;
                S_max = S;
                max_tm = T1;
            }
        }
    }
    if (hs == 1) return H_max;
    return S_max;
}

static int max5(double a, double b, double c_, double d, double e)
{
    if (a > b && a > c_ && a > d && a > e) return 1;
    else if (b > c_ && b > d && b > e) return 2;
    else if (c_ > d && c_ > e) return 3;
    else if (d > e) return 4;
    else return 5;
}

/* upstream calc_terminal_bp(): exterior loop of the hairpin, 5'->3' */
static void calc_terminal_bp(thal_ctx *c, double temp)
{
    int i, max;
    const double iH = c->dplx_init_H, iS = c->dplx_init_S, RC = c->RC;
    SEND5(0) = SEND5(1) = -1.0;
    HEND5(0) = HEND5(1) = PINF;
    for (i = 2; i <= c->len1; i++) {
        SEND5(i) = MinEntropy;
        HEND5(i) = 0;
    }
    double T1, T2, T3, T4, T5;
    double G;
    for (i = 2; i <= c->len1; ++i) {
        double eH[5], eS[5];
        for (int v = 1; v <= 4; ++v) {
            eH[v] = END5_x(c, v, i, 1);
            eS[v] = END5_x(c, v, i, 2);
        }
        c->relax += 4;
        T1 = (HEND5(i - 1) + iH) / (SEND5(i - 1) + iS + RC);
        T2 = (eH[1] + iH) / (eS[1] + iS + RC);
        T3 = (eH[2] + iH) / (eS[2] + iS + RC);
        T4 = (eH[3] + iH) / (eS[3] + iS + RC);
        T5 = (eH[4] + iH) / (eS[4] + iS + RC);
        max = max5(T1, T2, T3, T4, T5);
        if (max == 1) {
            SEND5(i) = SEND5(i - 1);
            HEND5(i) = HEND5(i - 1);
        } else {
            int v = max - 1;
            G = eH[v] - (temp * (eS[v]));
            if (G < G2_LIMIT) {
                SEND5(i) = eS[v];
                HEND5(i) = eH[v];
            } else {
                SEND5(i) = SEND5(i - 1);
                HEND5(i) = HEND5(i - 1);
            }
        }
    }
}

struct tracer {
    int i, j, mtrx;
};

/* upstream tracebacku(): hairpin traceback with an explicit stack */
static void tracebacku(thal_ctx *c, int *bp, int maxLoop)
{
    int i, j, ii, jj, k;
    int cap = 4 * (c->len1 + 4), sp = 0;
    struct tracer *stack = (struct tracer *)malloc(sizeof(struct tracer) * cap);
    double SH1[2], SH2[2], EE[2];
#define PUSH(I, J, M)                                                      \
    do {                                                                   \
        if (sp == cap) {                                                   \
            cap *= 2;                                                      \
            stack = (struct tracer *)realloc(stack, sizeof(*stack) * cap); \
        }                                                                  \
        stack[sp].i = (I);                                                 \
        stack[sp].j = (J);                                                 \
        stack[sp].mtrx = (M);                                              \
        ++sp;                                                              \
    } while (0)
    PUSH(c->len1, 0, 1);
    while (sp > 0) {
        --sp;
        i = stack[sp].i;
        j = stack[sp].j;
        int mtrx = stack[sp].mtrx;
        if (mtrx == 1) {
            while (i > 0 && equal(SEND5(i), SEND5(i - 1)) && equal(HEND5(i), HEND5(i - 1))) --i;
            if (i == 0) continue;
            int matched = 0;
            for (int v = 1; v <= 4 && !matched; ++v) {
                if (!(equal(SEND5(i), END5_x(c, v, i, 2)) && equal(HEND5(i), END5_x(c, v, i, 1)))) continue;
                matched = 1; /* upstream else-if chain: only the first matching variant is searched */
                int kmax = (v == 1) ? i - MIN_HRPN_LOOP - 2 : (v == 4) ? i - MIN_HRPN_LOOP - 4 : i - MIN_HRPN_LOOP - 3;
                for (k = 0; k <= kmax; ++k) {
                    int p, q;
                    double xH, xS;
                    switch (v) {
                    case 1: p = k + 1; q = i; xH = 0; xS = 0; break;
                    case 2: p = k + 2; q = i; xH = Hd5(c, i, k + 2); xS = Sd5(c, i, k + 2); break;
                    case 3: p = k + 1; q = i - 1; xH = Hd3(c, i - 1, k + 1); xS = Sd3(c, i - 1, k + 1); break;
                    default: p = k + 2; q = i - 1; xH = Htstack(c, i - 1, k + 2); xS = Ststack(c, i - 1, k + 2); break;
                    }
                    double s0, h0, s1, h1;
                    if (v == 1) {
                        s0 = atPenaltyS(numSeq1[p], numSeq1[q]) + EntropyDPT(p, q);
                        h0 = atPenaltyH(numSeq1[p], numSeq1[q]) + EnthalpyDPT(p, q);
                        s1 = SEND5(k) + atPenaltyS(numSeq1[p], numSeq1[q]) + EntropyDPT(p, q);
                        h1 = HEND5(k) + atPenaltyH(numSeq1[p], numSeq1[q]) + EnthalpyDPT(p, q);
                    } else {
                        s0 = atPenaltyS(numSeq1[p], numSeq1[q]) + xS + EntropyDPT(p, q);
                        h0 = atPenaltyH(numSeq1[p], numSeq1[q]) + xH + EnthalpyDPT(p, q);
                        s1 = SEND5(k) + atPenaltyS(numSeq1[p], numSeq1[q]) + xS + EntropyDPT(p, q);
                        h1 = HEND5(k) + atPenaltyH(numSeq1[p], numSeq1[q]) + xH + EnthalpyDPT(p, q);
                    }
                    if (equal(SEND5(i), s0) && equal(HEND5(i), h0)) {
                        PUSH(p, q, 0);
                        break;
                    } else if (equal(SEND5(i), s1) && equal(HEND5(i), h1)) {
                        PUSH(p, q, 0);
                        PUSH(k, 0, 1);
                        break;
                    }
                }
            }
        } else if (mtrx == 0) {
            bp[i - 1] = j;
            bp[j - 1] = i;
            SH1[0] = -1.0;
            SH1[1] = PINF;
            calc_hairpin(c, i, j, SH1, 1);
            SH2[0] = -1.0;
            SH2[1] = PINF;
            CBI(c, i, j, SH2, 2, maxLoop);
            if (equal(EntropyDPT(i, j), Ss(c, i, j, 2) + EntropyDPT(i + 1, j - 1)) &&
                equal(EnthalpyDPT(i, j), Hs(c, i, j, 2) + EnthalpyDPT(i + 1, j - 1))) {
                PUSH(i + 1, j - 1, 0);
            } else if (equal(EntropyDPT(i, j), SH1[0]) && equal(EnthalpyDPT(i, j), SH1[1])) {
                /* hairpin loop closes here */
            } else if (equal(EntropyDPT(i, j), SH2[0]) && equal(EnthalpyDPT(i, j), SH2[1])) {
                int d, done;
                for (done = 0, d = j - i - 3; d >= MIN_HRPN_LOOP + 1 && d >= j - i - 2 - maxLoop && !done; --d)
                    for (ii = i + 1; ii < j - d; ++ii) {
                        jj = d + ii;
                        EE[0] = -1.0;
                        EE[1] = PINF;
                        calc_bulge_internal2(c, i, j, ii, jj, EE, 1, maxLoop);
                        if (equal(EntropyDPT(i, j), EE[0] + EntropyDPT(ii, jj)) &&
                            equal(EnthalpyDPT(i, j), EE[1] + EnthalpyDPT(ii, jj))) {
                            PUSH(ii, jj, 0);
                            ++done;
                            break;
                        }
                    }
            }
        }
    }
#undef PUSH
    free(stack);
}

static void set_no_structure(po_oracle_result *o)
{
    o->tm = 0.0;
    o->dg = o->dh = o->ds = 0.0;
    o->structure_found = 0;
}

void po_oracle_thal(const po_oracle_params *P, const char *oligo_f, const char *oligo_r,
                    const po_oracle_args *a, po_oracle_result *o)
{
    memset(o, 0, sizeof(*o));
    o->align_end_1 = o->align_end_2 = -1;
    int len_f = (int)strlen(oligo_f), len_r = (int)strlen(oligo_r);
    if (a->type < 1 || a->type > 4) { o->status = 3; return; }
    if ((len_f > THAL_MAX_ALIGN) && (len_r > THAL_MAX_ALIGN)) { o->status = 2; return; }
    if (len_f > THAL_MAX_SEQ || len_r > THAL_MAX_SEQ) { o->status = 2; return; }
    if (a->type == PO_ORACLE_HAIRPIN && len_f > THAL_MAX_ALIGN) { o->status = 2; return; }
    if (len_f == 0 || len_r == 0) { o->status = 1; return; }

    thal_ctx ctx, *c = &ctx;
    memset(c, 0, sizeof(*c));
    c->P = P;
    const char *o1 = (a->type != 3) ? oligo_f : oligo_r;
    const char *o2 = (a->type != 3) ? oligo_r : oligo_f;
    c->len1 = (int)strlen(o1);
    c->len2 = (int)strlen(o2);
    c->stride = c->len2 + 2;
    numSeq1 = (unsigned char *)malloc(c->len1 + 2);
    numSeq2 = (unsigned char *)malloc(c->len2 + 2);
    for (int i = 1; i <= c->len1; ++i) numSeq1[i] = (unsigned char)base_code(toupper((unsigned char)o1[i - 1]));
    if (a->type == PO_ORACLE_HAIRPIN) {
        for (int i = 1; i <= c->len2; ++i) numSeq2[i] = (unsigned char)base_code(toupper((unsigned char)o2[i - 1]));
        c->dplx_init_H = 0.0;
        c->dplx_init_S = -0.00000000001;
        c->RC = 0;
    } else {
        /* oligo2 is reversed so it runs 3'->5' along the DP table */
        for (int i = 1; i <= c->len2; ++i) numSeq2[i] = (unsigned char)base_code(toupper((unsigned char)o2[c->len2 - i]));
        c->dplx_init_H = 200;
        c->dplx_init_S = -5.7;
        if (symmetry_thermo(o1) && symmetry_thermo(o2)) c->RC = R_GAS * log(a->dna_conc / 1000000000.0);
        else c->RC = R_GAS * log(a->dna_conc / 4000000000.0);
    }
    numSeq1[0] = numSeq1[c->len1 + 1] = numSeq2[0] = numSeq2[c->len2 + 1] = 4;
    c->saltCorrection = saltCorrectS(a->mv, a->dv, a->dntp);
    size_t cells = (size_t)(c->len1 + 2) * (size_t)c->stride;
    c->enthalpyDPT = (double *)calloc(cells, sizeof(double));
    c->entropyDPT = (double *)calloc(cells, sizeof(double));

    if (a->type == PO_ORACLE_HAIRPIN) {
        c->send5 = (double *)calloc(c->len1 + 1, sizeof(double));
        c->hend5 = (double *)calloc(c->len1 + 1, sizeof(double));
        initMatrix2(c);
        fillMatrix2(c, a->max_loop);
        calc_terminal_bp(c, a->temp_k);
        double mh = HEND5(c->len1), ms = SEND5(c->len1);
        o->align_end_1 = isfinite(mh) ? (int)mh : 0;
        o->align_end_2 = isfinite(ms) ? (int)ms : 0;
        if (isFinite(mh) && isFinite(ms)) {
            int *bp = (int *)calloc(c->len1, sizeof(int));
            tracebacku(c, bp, a->max_loop);
            /* upstream drawHairpin(): note the count stops before the last base */
            int N = 0;
            for (int i = 1; i < c->len1; ++i)
                if (bp[i - 1] > 0) N++;
            double t = (mh / (ms + (((N / 2) - 1) * c->saltCorrection))) - ABSOLUTE_ZERO;
            double mg = mh - (a->temp_k * (ms + (((N / 2) - 1) * c->saltCorrection)));
            ms = ms + (((N / 2) - 1) * c->saltCorrection);
            o->tm = t;
            o->dg = mg;
            o->dh = mh;
            o->ds = ms;
            o->structure_found = 1;
            free(bp);
        } else {
            set_no_structure(o);
        }
        free(c->send5);
        free(c->hend5);
    } else {
        double SH[2];
        int i, j, bestI = 0, bestJ = 0;
        double G1, bestG;
        initMatrix(c);
        fillMatrix(c, a->max_loop);
        G1 = bestG = PINF;
        if (a->type == 1) {
            for (i = 1; i <= c->len1; i++)
                for (j = 1; j <= c->len2; j++) {
                    RSH(c, i, j, SH);
                    c->relax += 1;
                    SH[0] = SH[0] + SMALL_NON_ZERO;
                    SH[1] = SH[1] + SMALL_NON_ZERO;
                    G1 = (EnthalpyDPT(i, j) + SH[1] + c->dplx_init_H) -
                         TEMP_KELVIN * (EntropyDPT(i, j) + SH[0] + c->dplx_init_S);
                    if (G1 < bestG) {
                        bestG = G1;
                        bestI = i;
                        bestJ = j;
                    }
                }
        } else {
            bestI = i = c->len1;
            for (j = 1; j <= c->len2; ++j) {
                RSH(c, i, j, SH);
                c->relax += 1;
                SH[0] = SH[0] + SMALL_NON_ZERO;
                SH[1] = SH[1] + SMALL_NON_ZERO;
                G1 = (EnthalpyDPT(i, j) + SH[1] + c->dplx_init_H) -
                     TEMP_KELVIN * (EntropyDPT(i, j) + SH[0] + c->dplx_init_S);
                if (G1 < bestG) {
                    bestG = G1;
                    bestJ = j;
                }
            }
        }
        if (!isFinite(bestG)) bestI = bestJ = 1;
        double dH, dS;
        RSH(c, bestI, bestJ, SH);
        dH = EnthalpyDPT(bestI, bestJ) + SH[1] + c->dplx_init_H;
        dS = (EntropyDPT(bestI, bestJ) + SH[0] + c->dplx_init_S);
        if (isFinite(EnthalpyDPT(bestI, bestJ))) {
            int *ps1 = (int *)calloc(c->len1, sizeof(int));
            int *ps2 = (int *)calloc(c->len2, sizeof(int));
            traceback(c, bestI, bestJ, ps1, ps2, a->max_loop);
            /* upstream drawDimer() */
            int N = 0;
            for (i = 0; i < c->len1; i++)
                if (ps1[i] > 0) ++N;
            for (i = 0; i < c->len2; i++)
                if (ps2[i] > 0) ++N;
            N = (N / 2) - 1;
            double t = ((dH) / (dS + (N * c->saltCorrection) + c->RC)) - ABSOLUTE_ZERO;
            double G = (dH) - (a->temp_k * (dS + (N * c->saltCorrection)));
            double S = dS + (N * c->saltCorrection);
            o->tm = t;
            o->dg = G;
            o->dh = dH;
            o->ds = S;
            o->structure_found = 1;
            o->align_end_1 = bestI;
            o->align_end_2 = bestJ;
            free(ps1);
            free(ps2);
        } else {
            set_no_structure(o);
        }
    }
    o->relaxations = c->relax;
    free(c->enthalpyDPT);
    free(c->entropyDPT);
    free(numSeq1);
    free(numSeq2);
}

/* ------------------------------------------------------------------------- */
/* batch drivers                                                             */
/* ------------------------------------------------------------------------- */
int po_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static char *item(const char *blob, const int64_t *off, int64_t k, char *buf, size_t cap)
{
    int64_t n = off[k + 1] - off[k];
    char *s = buf;
    if ((size_t)n + 1 > cap) s = (char *)malloc((size_t)n + 1);
    memcpy(s, blob + off[k], (size_t)n);
    s[n] = 0;
    return s;
}

void po_oracle_tm_batch(const char *blob, const int64_t *off, int64_t n,
                        double dna_conc, double mv, double dv, double dntp,
                        int nn_max_len, double *tm, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int64_t k = 0; k < n; ++k) {
        char buf[256];
        char *s = item(blob, off, k, buf, sizeof buf);
        tm[k] = po_oracle_seqtm(s, dna_conc, mv, dv, dntp, nn_max_len);
        if (s != buf) free(s);
    }
}

void po_oracle_thal_batch(const po_oracle_params *p, const po_oracle_args *a,
                          const char *blob1, const int64_t *off1,
                          const char *blob2, const int64_t *off2,
                          int64_t n, po_oracle_result *out, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    /* static partition: mirrors the reference's one-worker-per-core fan-out
     * (reference src/polyoligo/cli_kasp.py:354) */
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int64_t k = 0; k < n; ++k) {
        char b1[256], b2[256];
        char *s1 = item(blob1, off1, k, b1, sizeof b1);
        char *s2 = blob2 ? item(blob2, off2, k, b2, sizeof b2) : s1;
        po_oracle_thal(p, s1, s2, a, &out[k]);
        if (s1 != b1) free(s1);
        if (blob2 && s2 != b2) free(s2);
    }
}
