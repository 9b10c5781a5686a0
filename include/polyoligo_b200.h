/*
 * polyoligo_b200.h -- C ABI of the B200 (sm_100a) thermodynamic scoring library.
 *
 * This is the drop-in boundary for the one hot path of MirkoLedda/polyoligo:
 * batch nearest-neighbour Tm and primer3-style `thal` (hairpin / homodimer /
 * heterodimer) evaluation, the KASP allele-specific candidate enumeration that
 * feeds it and the validity filter / heterodimer check / goodness score /
 * per-marker top-n that consume it.  Each entry point names the reference call
 * (file:line under the reference tree) it replaces.  The reference reaches the
 * arithmetic through primer3-py 0.6.0 (Cython -> libprimer3); a maintainer
 * binds this library with ctypes instead.
 *
 * Conventions: every function returns 0 on success or a PO_E_* code; the
 * message of the last failure is available from po_last_error().  The caller
 * owns every buffer passed in; the library owns device memory inside po_ctx.
 * No torch types, no C++ types.  One po_ctx per process and GPU.
 */
#ifndef POLYOLIGO_B200_H
#define POLYOLIGO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PO_MAX_OLIGO_LEN 60 /* thal.c THAL_MAX_ALIGN; also the 2-bit record capacity */

enum {
    PO_OK = 0,
    PO_E_INVALID = 1, /* bad argument                                         */
    PO_E_CUDA = 2,    /* CUDA runtime failure (message has the CUDA string)   */
    PO_E_PARAMS = 3,  /* thermodynamic parameter files missing / malformed    */
    PO_E_NOMEM = 4
};

/* thal alignment types (libprimer3 thal.h thal_alignment_type) */
enum { PO_THAL_ANY = 1, PO_THAL_END1 = 2, PO_THAL_END2 = 3, PO_THAL_HAIRPIN = 4 };

/* One oligo, 16 bytes, self-contained and vector-loadable:
 *   bits 0..119  2-bit bases, base k at bits 2k..2k+1 (A=0 C=1 G=2 T=3)
 *   byte 15      bits 0..5 length (0..60), bit 6 = invalid character seen,
 *                bit 7 = contains N (or any other non-ACGT IUPAC letter)
 * Oligos with bit 6/7 set or longer than 60 nt are not evaluated: their results
 * carry PO_ITEM_SKIPPED.  (The reference rejects every N-containing primer
 * regardless of its thermodynamics, reference src/polyoligo/lib_primer3.py:72-74.) */
typedef struct {
    uint32_t w[4];
} po_oligo;

#define PO_OLIGO_HAS_N 0x80u
#define PO_OLIGO_BAD_CHAR 0x40u
#define PO_OLIGO_LEN_MASK 0x3Fu

/* Result of one thal evaluation (40 bytes).  Units follow primer3-py's
 * ThermoResult: tm in deg C, dg and dh in cal/mol, ds in cal/(mol K). */
typedef struct {
    double tm; /* 0.0 when no structure was found (libprimer3 convention)   */
    double dg;
    double dh;
    double ds;
    int32_t flags;     /* PO_ITEM_* bits                                      */
    int16_t align_end_1;
    int16_t align_end_2;
} po_thal_out;

#define PO_ITEM_STRUCTURE_FOUND 0x1
#define PO_ITEM_SKIPPED 0x2   /* empty / too long / non-ACGT input             */
#define PO_ITEM_SYMMETRIC 0x4 /* both strands self-complementary (RC used 1e9) */

/* Per-oligo properties computed by po_oligo_props_batch: the quantities
 * Primer.calc_thermals gathers (reference src/polyoligo/lib_primer3.py:56-64). */
typedef struct {
    double tm;           /* calcTm                                            */
    double hairpin_tm;   /* calcHairpin(seq).tm                               */
    double homodimer_tm; /* calcHomodimer(seq).tm                             */
    double gc_content;   /* Bio.SeqUtils.GC: 100 * (G + C) / len              */
    int32_t length;
    int32_t flags;       /* PO_ITEM_SKIPPED, PO_PRIMER_VALID                  */
} po_oligo_props;

#define PO_PRIMER_VALID 0x100 /* Primer.is_valid() (lib_primer3.py:66-97) passed */

/* Thresholds of Primer.is_valid: PRIMER3_GLOBALS values from the assay YAML
 * (reference src/polyoligo/data/PRIMER3_KASP.yaml) plus the CLI's tm_delta. */
typedef struct {
    int32_t min_size, max_size; /* PRIMER_MIN_SIZE / PRIMER_MAX_SIZE          */
    double min_gc, max_gc;      /* PRIMER_MIN_GC / PRIMER_MAX_GC               */
    double min_tm, max_tm;      /* PRIMER_MIN_TM / PRIMER_MAX_TM               */
    double tm_delta;            /* lib_primer3.TM_DELTA (cli_kasp.py --tm_delta) */
} po_valid_params;

typedef struct po_ctx po_ctx;

/* ---- lifetime ------------------------------------------------------------ */
/* Creates the context on CUDA device `device`.  param_dir: directory in the
 * libprimer3 primer3_config format (stack.ds ... tetraloop.dh) or NULL for the
 * compiled-in default set.  Replaces: ThermoAnalysis() construction,
 * reference src/polyoligo/lib_primer3.py:58,157; _lib_kasp.py:99. */
int po_init(int device, const char *param_dir, po_ctx **out);
void po_destroy(po_ctx *ctx);
/* ctx may be NULL: returns the message of the last failed po_init. */
const char *po_last_error(const po_ctx *ctx);

/* ThermoAnalysis attributes mv_conc, dv_conc, dntp_conc, dna_conc, temp_c,
 * max_loop, max_nn_length (defaults 50, 0, 0.8, 50, 37, 30, 60: the reference
 * never changes them). */
int po_set_conditions(po_ctx *ctx, double mv, double dv, double dntp, double dna_nM,
                      double temp_c, int max_loop, int max_nn_length);

/* ---- packing (host, no GPU needed) --------------------------------------- */
/* n ASCII sequences stored back to back in `ascii`, item k = bytes
 * [offsets[k], offsets[k+1]).  Case-insensitive. */
int po_pack(const char *ascii, const int64_t *offsets, int64_t n, po_oligo *packed);
/* Inverse; out must hold 61 bytes per item (NUL terminated, stride 61). */
int po_unpack(const po_oligo *packed, int64_t n, char *out);

/* ---- scoring, HOST buffers (copies are inside the call) ------------------- */
/* calcTm for n oligos (lib_primer3.py:59, _lib_crispr.py:42).  gc may be NULL. */
int po_tm_batch(po_ctx *ctx, const po_oligo *oligos, int64_t n, double *tm, double *gc);

/* thal for n (a[k], b[k]) pairs.  type ANY with b == NULL is calcHomodimer
 * (lib_primer3.py:61); with b it is calcHeterodimer (lib_primer3.py:158,
 * _lib_kasp.py:100-105); type HAIRPIN ignores b (calcHairpin, lib_primer3.py:60). */
int po_thal_batch(po_ctx *ctx, int type, const po_oligo *a, const po_oligo *b, int64_t n,
                  po_thal_out *out);

/* Primer.calc_thermals + Primer.is_valid for n oligos (lib_primer3.py:56-97).
 * valid may be NULL (no validity bit is computed). */
int po_oligo_props_batch(po_ctx *ctx, const po_oligo *oligos, int64_t n,
                         const po_valid_params *valid, po_oligo_props *out);

/* Score-only path for n given primer pairs (the shape of reverse_engineer.py:
 * process_assay, reference reverse_engineer.py:61-112, and of BASELINE config 4):
 * calcHeterodimer(a,b) + calcHairpin(a) + calcHairpin(b) + calcTm(a) + calcTm(b).
 * Any output pointer may be NULL to skip that evaluation.  Large batches are
 * processed in bounded device-memory chunks. */
int po_score_pairs_batch(po_ctx *ctx, const po_oligo *a, const po_oligo *b, int64_t n,
                         po_thal_out *het, po_thal_out *hairpin_a, po_thal_out *hairpin_b,
                         double *tm_a, double *tm_b);

/* ---- scoring, DEVICE buffers (no copies; asynchronous on `stream`) -------- */
/* Pointers are CUDA device pointers; stream is a cudaStream_t (NULL = the
 * context's own stream).  The call returns after enqueueing. */
int po_tm_batch_dev(po_ctx *ctx, const void *d_oligos, int64_t n, void *d_tm, void *d_gc,
                    void *stream);
int po_thal_batch_dev(po_ctx *ctx, int type, const void *d_a, const void *d_b, int64_t n,
                      void *d_out, void *stream);
int po_score_pairs_batch_dev(po_ctx *ctx, const void *d_a, const void *d_b, int64_t n,
                             void *d_het, void *d_hairpin_a, void *d_hairpin_b, void *d_tm_a,
                             void *d_tm_b, void *stream);
/* Blocks until everything enqueued on `stream` (NULL = context stream) is done. */
int po_sync(po_ctx *ctx, void *stream);

/* ---- KASP candidate enumeration ------------------------- */
/* One marker: the homolog-mapped template around it (hroi.seq, reference
 * src/polyoligo/_lib_kasp.py:296, OUTER_LEN = 500 nt) and the marker alleles. */
typedef struct {
    uint32_t tmpl[32];     /* 2-bit template, up to 512 nt, base k at bits 2k..2k+1 of word k/16 */
    uint32_t nmask[16];    /* bit k set: template position k is not A/C/G/T               */
    uint32_t alt[2];       /* ALT allele, 2-bit, up to 32 nt                               */
    uint32_t alt_nmask;
    int32_t tmpl_len;
    int32_t marker_pos;    /* hroi.marker.pos1 - hroi.start        (_lib_kasp.py:297)      */
    int32_t marker_pos_rc; /* hroi.stop - hroi.marker.pos1 + 1     (_lib_kasp.py:298)      */
    int32_t ref_len;       /* len(hroi.marker.ref)                                         */
    int32_t alt_len;       /* len(hroi.marker.alt)                                         */
    int32_t reserved[8];
} po_kasp_marker; /* 256 bytes */

/* get_marker_primers (reference src/polyoligo/_lib_kasp.py:295-359), enumeration
 * only: for marker m, padding p in [min_size, max_size] and direction d (0 = F,
 * 1 = R), slot = (p - min_size) * 2 + d holds the REF primer (ref_out) and its
 * ALT twin (alt_out) at index m * slots + slot, slots = 2 * (max_size - min_size
 * + 1), in the reference's visiting order.  Validity (the reference keeps a
 * slot iff both primers pass Primer.is_valid) comes from po_oligo_props_batch. */
int po_kasp_enumerate_batch(po_ctx *ctx, const po_kasp_marker *markers, int64_t n, int min_size,
                            int max_size, po_oligo *ref_out, po_oligo *alt_out);

/* ---- heterodimer predicate, goodness score, per-marker top-n (rows a8, a9) --- */
enum { PO_ASSAY_PCR = 0 /* also CAPS */, PO_ASSAY_KASP = 1, PO_ASSAY_HRM = 2 };

/* qcode letters of scoring.py as bits, plus the KASP dimer flags */
#define PO_Q_t 0x001
#define PO_Q_O 0x002
#define PO_Q_d 0x004
#define PO_Q_m 0x008
#define PO_Q_M 0x010
#define PO_Q_i 0x020
#define PO_Q_I 0x040
#define PO_Q_h 0x080
#define PO_Q_H 0x100
#define PO_Q_REF_DIMER 0x1000
#define PO_Q_ALT_DIMER 0x2000

typedef struct {
    int32_t assay;            /* PO_ASSAY_*                                              */
    int32_t top_n;            /* n of PCR.prune(n)                                       */
    int64_t n_pairs;
    int64_t n_markers;
    const po_oligo *f, *r;    /* forward / reverse primer of every pair, 5'->3'          */
    const po_oligo *a;        /* KASP: ALT allele primer; NULL otherwise                 */
    const uint8_t *dir;       /* KASP: 0 = marker primer is F, 1 = it is R (pp.dir)      */
    const int64_t *seg;       /* n_markers + 1 offsets: pairs of marker m, in pps order  */
    const int32_t *n_offtargets; /* len(pp.offtargets) (BLAST stays external glue)       */
    const double *max_aaf;    /* n_pairs x 3: primers F, R, A (A ignored unless KASP)    */
    const int32_t *max_indel; /* pp.max_indel_size                                        */
    const double *hrm_delta;  /* pp.hrm["delta"]; HRM only, else NULL                     */
    po_oligo reporters[2];    /* KASP reporter dyes (cli_kasp.py --dye1 / --dye2)        */
    double tm_delta;
} po_rank_in;

typedef struct {
    int32_t *goodness;        /* n_pairs                                                 */
    int32_t *qbits;           /* n_pairs: PO_Q_* bits                                    */
    double *tm;               /* n_pairs x 3: calcTm of F, R, A                          */
    double *het_tm;           /* n_pairs x 2: heterodimer Tm (KASP: ref pair, alt pair)  */
    int64_t *sel;             /* n_markers x top_n: pair indices in report order, -1 pad */
    int32_t *n_sel;           /* n_markers                                               */
} po_rank_out;

/* PrimerPair.check_heterodimerization (reference lib_primer3.py:147-164; KASP
 * override _lib_kasp.py:89-117 after add_reporter_dyes :75-86) + PCR.classify
 * (scoring.py:4-151 via lib_primer3.py:425-431) + PCR.prune (lib_primer3.py:
 * 433-463, _lib_kasp.py:278-291) for all markers of a batch. */
int po_rank_pairs_batch(po_ctx *ctx, const po_rank_in *in, const po_rank_out *out);

/* ---- "next" rows of SURVEY 8f ------------------------------------------------ */
/* HRM product melting temperature: Biopython 1.76 MeltingTemp.Tm_NN(seq,
 * strict=False) with its defaults (Allawi & SantaLucia 1997 table DNA_NN3,
 * dnac1 = dnac2 = 25 nM, Na = 50 mM, saltcorr = 5) as called by get_tm
 * (reference src/polyoligo/lib_markers.py:586-587) from
 * PCRproduct.get_hrm_temps (:518-529).  Products are 40-150 nt, longer than an
 * oligo record, so the input is ASCII: item k = bytes [offsets[k], offsets[k+1]).
 * Like Biopython's _check(), letters other than A/C/G/T/U/I are DROPPED before anything is summed
 * (the reference calls Tm_NN(seq, strict=False)); NaN only when nothing is left. */
int po_tm_nn_batch(po_ctx *ctx, const char *ascii, const int64_t *offsets, int64_t n, double *tm);

/* Numeric part of PrimerPair.add_mutations (reference src/polyoligo/
 * lib_primer3.py:166-210): max alternative-allele frequency of the mutations
 * inside every primer window [start, stop] and the largest indel inside the
 * product window [F.start, R.stop].  Mutations and pairs are grouped per marker
 * (mut_seg / pair_seg: n_markers + 1 offsets).  primer_start / primer_stop are
 * n_pairs x 3 (F, R, A; A ignored when n_primers == 2). */
int po_annotate_mutations_batch(po_ctx *ctx, const int64_t *mut_pos, const double *mut_aaf,
                                const int32_t *mut_indel, const int64_t *mut_seg,
                                const int64_t *pair_seg, int64_t n_markers,
                                const int64_t *primer_start, const int64_t *primer_stop,
                                int n_primers, double *max_aaf, int32_t *max_indel);

/* ---- primer3-core style picking of the complementary primer ------------------
 * Replaces primer3.bindings.designPrimers as reached from lib_primer3.get_primers
 * (reference src/polyoligo/lib_primer3.py:466-529, globals set at :539-544; callers
 * _lib_kasp.py:567-590 with SEQUENCE_PRIMER / SEQUENCE_PRIMER_REVCOMP fixed, _lib_pcr.py:
 * 167-180 with both primers free).  libprimer3 is third-party and absent from the reference
 * tree: the restated semantics (candidate filters, penalty, pair constraints, ordering) are
 * listed in DESIGN.md section 7; parity with the real library is UNPINNED. */
typedef struct {
    int32_t min_size, opt_size, max_size;            /* PRIMER_{MIN,OPT,MAX}_SIZE                 */
    int32_t num_return;                               /* PRIMER_NUM_RETURN                          */
    double min_tm, opt_tm, max_tm;                    /* PRIMER_{MIN,OPT,MAX}_TM                    */
    double min_gc, max_gc;                            /* PRIMER_{MIN,MAX}_GC                        */
    double max_self_any_th, max_self_end_th, max_hairpin_th;      /* PRIMER_MAX_*_TH              */
    double pair_max_compl_any_th, pair_max_compl_end_th;          /* PRIMER_PAIR_MAX_COMPL_*_TH   */
    double pair_max_diff_tm;                          /* PRIMER_PAIR_MAX_DIFF_TM                    */
    double wt_tm_gt, wt_tm_lt, wt_size_lt, wt_size_gt;            /* PRIMER_WT_*                  */
    double mv_conc, dv_conc, dntp_conc, dna_conc;     /* PRIMER_SALT_*, PRIMER_DNTP_CONC, PRIMER_DNA_CONC */
    int32_t max_poly_x, gc_clamp, max_end_gc;         /* PRIMER_MAX_POLY_X, PRIMER_GC_CLAMP, PRIMER_MAX_END_GC */
    int32_t n_size_ranges;                            /* PRIMER_PRODUCT_SIZE_RANGE, tried in order  */
    int32_t size_lo[8], size_hi[8];
} po_p3_settings;

typedef struct {
    po_oligo seq;                 /* the primer, 5'->3'                                            */
    double tm, gc_percent, penalty;
    double self_any_th, self_end_th, hairpin_th;      /* NaN until evaluated (lazily, like libprimer3) */
    int32_t start;                /* primer3 convention: index of the primer's 5' base on the template
                                     (for a right primer that is its rightmost template base)        */
    int32_t len;
    int32_t set;                  /* template the oligo was cut from                                 */
    int32_t flags;                /* PO_P3_*                                                          */
} po_p3_oligo;                    /* 80 bytes                                                         */
enum { PO_P3_OK = 1, PO_P3_SELF_DONE = 2, PO_P3_SELF_FAILED = 4, PO_P3_QUEUED = 8,
       PO_P3_PRIMER_VALID = 16 /* passes Primer.is_valid (only with `valid` below) */ };

typedef struct {
    int32_t set;                  /* template index                                                   */
    int32_t side;                 /* 1: left primer given (SEQUENCE_PRIMER), 2: right primer given
                                     (SEQUENCE_PRIMER_REVCOMP)                                        */
    int32_t start, len;           /* the given primer on the template, primer3 convention            */
} po_p3_task;

typedef struct {
    po_p3_oligo other;            /* the picked complementary primer                                  */
    double pair_penalty, compl_any_th, compl_end_th;
    int32_t product_size;
    int32_t size_range;           /* index of the PRIMER_PRODUCT_SIZE_RANGE entry it was found in     */
} po_p3_pair;                     /* 112 bytes                                                        */

/* Candidate oligos of n_sets templates: every window of min..max_size nt that avoids masked
 * bases (mask != 0: SEQUENCE_EXCLUDED_REGION / SEQUENCE_TARGET / outside SEQUENCE_INCLUDED_REGION)
 * and non-ACGT letters and passes the GC, clamp, poly-X and Tm limits, with its penalty.
 * Lists are sorted like libprimer3 sorts them (penalty, then start descending, then length):
 * list 2*s = left primers of set s, list 2*s+1 = right primers.  list_off has 2*n_sets+1
 * entries.  With out == NULL only list_off is produced (sizing call). */
int po_p3_candidates_batch(po_ctx *ctx, const po_p3_settings *s, const char *tmpl, const uint8_t *mask,
                           const int64_t *set_off, int64_t n_sets, int64_t *list_off,
                           po_p3_oligo *out, int64_t out_cap);

/* designPrimers with one primer fixed, for n_tasks (template, given primer) tasks: the best
 * num_return complementary primers per task in libprimer3's order.  set_target is NULL or
 * 2 x n_sets (start, len; len <= 0: none).  fixed_out (n_tasks) receives the given primer's own
 * record (flags without PO_P3_OK: it violates the constraints and the task returns nothing);
 * pairs_out is n_tasks x num_return, n_pairs_out the valid count per task.
 * valid != NULL additionally runs Primer.is_valid (reference lib_primer3.py:66-97, what the KASP
 * merge asks of every picked common primer, _lib_kasp.py:621-625) on each distinct picked
 * primer under the context's own conditions and reports it as PO_P3_PRIMER_VALID in
 * pairs_out[].other.flags. */
int po_p3_design_fixed_batch(po_ctx *ctx, const po_p3_settings *s, const char *tmpl, const uint8_t *mask,
                             const int64_t *set_off, const int32_t *set_target, int64_t n_sets,
                             const po_p3_task *tasks, int64_t n_tasks, const po_valid_params *valid,
                             po_p3_oligo *fixed_out, po_p3_pair *pairs_out, int32_t *n_pairs_out);

/* ---- the KASP design of a marker batch behind one call (rows a3-a9) -------------------------- */
/* The design part of the reference's per-marker worker, src/polyoligo/_lib_kasp.py:671-743 (main)
 * with the BLAST off-target search left out (no off-targets): allele-specific candidate
 * enumeration, validity, the search masks (include_mut_in_included_maps lib_primer3.py:576-593, the
 * marker window _lib_kasp.py:681-684, get_sequence_excluded_regions lib_primer3.py:596-600), one
 * designPrimers per marker primer and mask, the round-robin merge (_lib_kasp.py:591-631), mutation
 * annotation, reporter-tailed heterodimers, score_kasp, classify and prune.  Everything between the
 * inputs and the top-n triples stays on the device. */
typedef struct {
    int32_t min_size, max_size;   /* PRIMER_MIN_SIZE / PRIMER_MAX_SIZE: marker-primer paddings (at most 16 sizes) */
    int32_t top_n;                /* pairs kept per marker (cli_kasp -n, default 10)                  */
    int32_t with_mutations;       /* 1: VCF hook on, six search masks (x_mut twins first); 0: three   */
    double tm_delta;              /* cli_kasp --tm_delta, default 5.0                                 */
    po_valid_params valid;        /* Primer.is_valid thresholds                                       */
    po_oligo reporters[2];        /* reporter dyes (cli_kasp.py:65,72); length 0 for none             */
} po_kasp_design_params;

typedef struct {
    po_oligo f, r, a;             /* forward, reverse and ALT-allele primer, 5'->3'                   */
    int32_t goodness;             /* score_kasp                                                       */
    int32_t qbits;                /* PO_Q_* | PO_Q_REF_DIMER | PO_Q_ALT_DIMER                         */
    uint8_t dir;                  /* 0: the marker primers are forward, 1: reverse                    */
    uint8_t pad[7];
} po_kasp_pair;                   /* 64 bytes */

/* markers[n]: templates of ONE common length (<= 512 nt); starts[n]: hroi.start (genome coordinate of
 * template base 0); maps[n][3][16]: inclusion maps mism, partial_mism, all as bit maps (bit p % 32 of
 * word p / 32: base p may be covered by a primer); mutations of marker k: entries mut_seg[k] ..
 * mut_seg[k+1]-1 of mut_pos (genome coordinate) / mut_aaf / mut_indel (ignored without
 * with_mutations).  `picker`: designPrimers settings, num_return = depth (cli_kasp --depth, 100).
 * out[n][top_n] / n_out[n]: the selected triples of every marker in report order (descending
 * goodness, merge order inside a class); n_merged[n] (optional): pairs in the marker's repository. */
int po_kasp_design_batch(po_ctx *ctx, const po_p3_settings *picker, const po_kasp_design_params *params,
                         const po_kasp_marker *markers, const int64_t *starts, int64_t n_markers,
                         const uint32_t *maps, const int64_t *mut_pos, const double *mut_aaf,
                         const int32_t *mut_indel, const int64_t *mut_seg, po_kasp_pair *out,
                         int32_t *n_out, int32_t *n_merged);

/* Row a4 alone: the excluded-base bit maps of every search type, excl[n_types][n][16] (bit set: no
 * primer may cover the base), types in the reference's order (_lib_kasp.py:715-718: mism_mut, mism,
 * partial_mism_mut, partial_mism, all_mut, all with mutations; mism, partial_mism, all without).
 * kasp_window != 0 applies the KASP marker-window step with PRIMER_MAX_SIZE = max_size; 0 leaves the
 * maps as PCR / CAPS / HRM use them. */
int po_kasp_masks_batch(po_ctx *ctx, const int64_t *starts, int64_t n_markers, int32_t tmpl_len,
                        const uint32_t *maps, const int64_t *mut_pos, const int64_t *mut_seg,
                        int32_t with_mutations, int32_t kasp_window, int32_t max_size, uint32_t *excl);

/* ---- instrumentation ------------------------------------------------------ */
/* Number of kernels this context has launched since po_init. */
int64_t po_launch_count(const po_ctx *ctx);
/* Number of thal evaluations (hairpin / dimer items of any entry point, including the ones the
 * primer picker and the ranking issue internally) this context has run since po_init. */
int64_t po_thal_item_count(const po_ctx *ctx);
/* DP relaxations (SURVEY 8d work unit, counted exactly as the upstream
 * algorithm performs them) of one po_thal_batch over host buffers; runs the
 * instrumented kernel variant, not for timing. */
int po_thal_count_relaxations(po_ctx *ctx, int type, const po_oligo *a, const po_oligo *b,
                              int64_t n, int64_t *relaxations);
/* Register-resident FP64 FMA microbenchmark: measured non-tensor FP64 peak of
 * the device in TFLOP/s (the compute-roofline denominator). */
int po_measure_fp64_peak(po_ctx *ctx, double *tflops);

#ifdef __cplusplus
}
#endif
#endif
