// po_lib.cu -- kernels, launchers and the C ABI of libpolyoligo_b200.so.
// The ABI is declared in include/polyoligo_b200.h; each entry point names the
// reference call it replaces there.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

#include "po_common.h"
#include "po_thal.cuh"
#include "po_thal_dimer.cuh"
#include "po_thal_dimer32.cuh"
#include "po_thal_hairpin.cuh"
#include "po_thal_hairpin32.cuh"
#include "po_pipeline.cuh"

// ============================================================================
// kernels
// ============================================================================

__device__ __forceinline__ void stage_tables(SmemTables* tb, const PoTables* gt) {
    const uint4* src = reinterpret_cast<const uint4*>(gt);
    uint4* dst = reinterpret_cast<uint4*>(tb);
    for (int idx = threadIdx.x; idx < (int)(sizeof(SmemTables) / 16); idx += blockDim.x) dst[idx] = src[idx];
    __syncthreads();
}

// lanes per concurrently relaxed cell (see Grp<G> in po_thal.cuh).  Measured on B200 with
// config 4 (3M pairs): dimer 16 lanes x 2 cells 226.8 ms vs 32 x 1 241.8 ms vs 8 x 4 252.4 ms;
// hairpin 32 x 1 210.3 ms vs 16 x 2 224.4 ms vs 8 x 4 254.3 ms (fewer candidates per cell).
#ifndef PO_HAIRPIN_GROUP
#define PO_HAIRPIN_GROUP 32
#endif
#ifndef PO_DIMER_GROUP
#define PO_DIMER_GROUP 16
#endif

// Work distribution: every warp pulls the next item from a global counter, so
// long pairs do not stall a fixed partition.  Items whose finite-cell count
// exceeds CAP are appended to `overflow` and picked up by the large-CAP pass.
//   list == nullptr : items are 0..n-1
//   list != nullptr : items are list[0 .. *list_n)
// warps per CTA of the first-tier kernels (2 CTAs per SM; 72 registers per thread).  Measured
// on B200 (3M config-4 pairs): 13 / 14 warps 227.3 / 206.8 ms vs 12 / 12 warps 229.4 / 210.2 ms.
#ifndef PO_DIMER_WARPS
#define PO_DIMER_WARPS 24
#endif
#ifndef PO_HAIRPIN_WARPS
#define PO_HAIRPIN_WARPS 28
#endif
#define PO_DIMER32_CAP 256
#ifndef PO_DIMER32_GROUP
#define PO_DIMER32_GROUP 8
#endif
#define PO_HAIRPIN32_CAP 128
#ifndef PO_HAIRPIN_CHUNK
#define PO_HAIRPIN_CHUNK (1ll << 20)  // hairpins per fill / finish kernel pair: 2 KB of cell values each
#endif
// double2 units a hairpin hands from the fill to the finish kernel: CAP cell values + CAP 32-bit cell codes
#define PO_HAIRPIN_XFER(cap) ((cap) + (cap) / 4)
#ifndef PO_HAIRPIN32_GROUP
#define PO_HAIRPIN32_GROUP 8
#endif
template <int CAP, bool COUNT, bool HAIRPIN>
struct ThalWs {
    using type = typename std::conditional<
        HAIRPIN,
        typename std::conditional<(!COUNT && CAP == PO_HAIRPIN32_CAP), Hairpin32Ws<CAP>, HairpinWs<CAP>>::type,
        typename std::conditional<(!COUNT && CAP == PO_DIMER32_CAP), Dimer32Ws<CAP>, DimerWs<CAP>>::type>::type;
};
// PHASE (first hairpin tier only): 0 = whole evaluation, 1 = table fill writing the cell values to
// `cells`, 2 = finish from `cells` (see thal_hairpin32_warp).  `first`: items are first .. first+n-1
// when no list is given (sub-batches of one call share the overflow list).
template <int CAP, int WARPS, bool COUNT, bool HAIRPIN, int PHASE = 0>
__global__ void __launch_bounds__(WARPS * 32, 1)
k_thal(const uint4* __restrict__ A, const uint4* __restrict__ B, long long n, po_thal_out* __restrict__ out,
       const PoTables* __restrict__ gt, const PoConsts K, unsigned long long* work,
       const long long* __restrict__ list, const unsigned long long* __restrict__ list_n,
       long long* overflow, unsigned long long* overflow_n, unsigned long long* relax, int end_mode,
       po_thal_out* __restrict__ out2, long long first = 0, double2* __restrict__ cells = nullptr) {
    // first dimer tier: oligos <= 32 nt, column-sorted fill (po_thal_dimer32.cuh)
    constexpr bool FAST32 = !HAIRPIN && !COUNT && CAP == PO_DIMER32_CAP;
    using Ws = typename ThalWs<CAP, COUNT, HAIRPIN>::type;
    extern __shared__ __align__(16) unsigned char smem[];
    if (list && *list_n == 0ull) return;  // an overflow tier nobody reached: skip the table staging (block-uniform)
    SmemTables* tb = reinterpret_cast<SmemTables*>(smem);
    SmemExtra* xt = reinterpret_cast<SmemExtra*>(smem + sizeof(SmemTables));
    Ws* wsAll = reinterpret_cast<Ws*>(smem + sizeof(SmemTables) + sizeof(SmemExtra));
    stage_tables(tb, gt);
    fill_extra(xt, *tb, K, gt);
    const TabAddr TA = tab_addr(*tb, *xt);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Ws& ws = wsAll[warp];
    const unsigned long long total = list ? *list_n : (unsigned long long)n;
    for (;;) {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(work, 1ull);
        t = __shfl_sync(FULL, t, 0);
        if (t >= total) break;
        const long long item = list ? list[t] : first + (long long)t;
        uint4 oa = A[item];
        // upper bound of the finite-cell count: exact count is cheap from base composition
        if constexpr (HAIRPIN) {
            int a, c, g, t_;
            po_base_counts(oa, po_len(oa), a, c, g, t_);
            const int ub = a * t_ + c * g;
            if constexpr (!COUNT && CAP == PO_HAIRPIN32_CAP) {
                const int len = po_len(oa);
                if (po_bad(oa) || len == 0) {
                    if (lane == 0 && PHASE != 2) {
                        po_thal_out r;
                        r.tm = r.dg = r.dh = r.ds = 0.0;
                        r.flags = PO_ITEM_SKIPPED;
                        r.align_end_1 = r.align_end_2 = 0;
                        out[item] = r;
                    }
                    continue;
                }
                if (ub > CAP || len > 32) {
                    if (lane == 0 && PHASE != 2) overflow[atomicAdd(overflow_n, 1ull)] = item;
                    continue;
                }
                thal_hairpin32_warp<PO_HAIRPIN32_GROUP, CAP, PHASE>(*tb, *xt, TA, gt, ws, oa, K, out + item,
                                                                    cells ? cells + (size_t)(item - first) * PO_HAIRPIN_XFER(CAP) : nullptr);
            } else {
                if (ub > CAP) {
                    if (lane == 0) overflow[atomicAdd(overflow_n, 1ull)] = item;
                    continue;
                }
                thal_hairpin_warp<PO_HAIRPIN_GROUP, CAP, COUNT>(*tb, *xt, TA, gt, ws, oa, K, out + item, relax);
            }
        } else {
            uint4 ob = B ? B[item] : oa;
            if (end_mode == 2) {  // thal type END2 is END1 with the oligos swapped (upstream thal())
                const uint4 tmp = oa;
                oa = ob;
                ob = tmp;
            }
            int a1, c1, g1, t1, a2, c2, g2, t2;
            po_base_counts(oa, po_len(oa), a1, c1, g1, t1);
            po_base_counts(ob, po_len(ob), a2, c2, g2, t2);
            const int ncell = a1 * t2 + c1 * g2 + g1 * c2 + t1 * a2;
            if constexpr (FAST32) {
                const int len1 = po_len(oa), len2 = po_len(ob);
                if (po_bad(oa) || po_bad(ob) || len1 == 0 || len2 == 0) {
                    if (lane == 0) {
                        po_thal_out r;
                        r.tm = r.dg = r.dh = r.ds = 0.0;
                        r.flags = PO_ITEM_SKIPPED;
                        r.align_end_1 = r.align_end_2 = -1;
                        out[item] = r;
                        if (out2) out2[item] = r;
                    }
                    continue;
                }
                if (ncell > CAP || len1 > 32 || len2 > 32) {
                    if (lane == 0) overflow[atomicAdd(overflow_n, 1ull)] = item;
                    continue;
                }
                thal_dimer32_warp<PO_DIMER32_GROUP, CAP>(*tb, *xt, TA, ws, oa, ob, K, end_mode != 0, out + item,
                                                       out2 ? out2 + item : nullptr);
            } else {
                if (ncell > CAP) {
                    if (lane == 0) overflow[atomicAdd(overflow_n, 1ull)] = item;
                    continue;
                }
                thal_dimer_warp<PO_DIMER_GROUP, CAP, COUNT>(*tb, *xt, TA, ws, oa, ob, K, end_mode != 0, out + item,
                                                            out2 ? out2 + item : nullptr, relax);
            }
        }
    }
}

// calcTm + GC content, one thread per oligo (upstream oligotm.c: seqtm/oligotm
// with SantaLucia tables and SantaLucia salt correction).
__constant__ int c_tm_dh[16] = {79, 84, 78, 72, 85, 80, 106, 78, 82, 98, 80, 84, 72, 82, 85, 79};
__constant__ int c_tm_ds[16] = {222, 224, 210, 204, 227, 199, 272, 210, 222, 244, 199, 224, 213, 222, 227, 222};

__device__ __forceinline__ double oligo_tm(const uint4& o, const PoConsts& K, double* gc_out) {
    const int len = po_len(o);
    int gc = 0;
    for (int p = 0; p < len; ++p) {
        const int b = po_base(o, p);
        gc += (b == 1 || b == 2);
    }
    if (gc_out) *gc_out = len ? gc * 100.0 / len : 0.0;
    if (po_bad(o) || len < 2) return -999999.9999;
    if (len > K.max_nn_length)
        return 81.5 + (16.6 * K.tm_log10_mv) + (41.0 * (((double)gc) / len)) - (600.0 / len);
    int dh = 0, ds = 0;
    bool sym = (len % 2 == 0);
    for (int p = 0; p < len / 2 && sym; ++p) sym = (po_base(o, p) + po_base(o, len - 1 - p) == 3);
    if (sym) ds += 14;
    const int first = po_base(o, 0), last = po_base(o, len - 1);
    if (first == 0 || first == 3) { ds += -41; dh += -23; } else { ds += 28; dh += -1; }
    if (last == 0 || last == 3) { ds += -41; dh += -23; } else { ds += 28; dh += -1; }
    int prev = first;
    for (int p = 1; p < len; ++p) {
        const int b = po_base(o, p);
        dh += c_tm_dh[4 * prev + b];
        ds += c_tm_ds[4 * prev + b];
        prev = b;
    }
    const double delta_H = dh * -100.0;
    double delta_S = ds * -0.1;
    delta_S = delta_S + 0.368 * (len - 1) * K.tm_log_mv;
    return delta_H / (delta_S + (sym ? K.tm_rlog_sym : K.tm_rlog_non)) - 273.15;
}

__global__ void k_tm(const uint4* __restrict__ A, long long n, double* __restrict__ tm, double* __restrict__ gc,
                     const PoConsts K) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n) return;
    double g;
    const double v = oligo_tm(A[t], K, &g);
    tm[t] = v;
    if (gc) gc[t] = g;
}

// Primer.calc_thermals + Primer.is_valid (reference lib_primer3.py:56-97) from
// the three per-oligo evaluations.
__global__ void k_props(const uint4* __restrict__ A, long long n, const po_thal_out* __restrict__ hp,
                        const po_thal_out* __restrict__ hd, const PoConsts K, const po_valid_params V,
                        int have_valid, po_oligo_props* __restrict__ out) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint4 o = A[t];
    po_oligo_props r;
    double g;
    r.tm = oligo_tm(o, K, &g);
    r.gc_content = g;
    r.hairpin_tm = hp[t].tm;
    r.homodimer_tm = hd[t].tm;
    r.length = po_len(o);
    r.flags = (hp[t].flags | hd[t].flags) & PO_ITEM_SKIPPED;
    if (have_valid && !(r.flags & PO_ITEM_SKIPPED)) {
        bool ok = true;
        if (r.length < V.min_size || r.length > V.max_size) ok = false;
        if (r.gc_content < V.min_gc || r.gc_content > V.max_gc) ok = false;
        if (r.tm < V.min_tm || r.tm > V.max_tm) ok = false;
        if (r.hairpin_tm >= (r.tm - V.tm_delta)) ok = false;
        if (r.homodimer_tm >= (r.tm - V.tm_delta)) ok = false;
        if (ok) r.flags |= PO_PRIMER_VALID;
    }
    out[t] = r;
}

// register-resident FP64 FMA chains: the compute-roofline denominator
__global__ void k_fp64_peak(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ============================================================================
// context
// ============================================================================

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct po_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    PoTables* d_tables = nullptr;
    PoConsts K;
    int num_sms = 0;
    std::string err;
    int64_t launches = 0;
    int64_t thal_items = 0;  // thal evaluations requested since po_init
    // grow-only device scratch
    DevBuf a, b, out, out2, out3, tmv, gcv, props, overflow, hp_cells;
    DevBuf scratch[12];
    DevBuf p3[28];  // primer picking (po_p3.cuh)
    DevBuf kd[56];  // fused KASP design (po_kasp_design.cuh)
    unsigned long long* d_ctr = nullptr;  // [0] work, [1] overflow_n, [2] work2, [4] overflow2_n, [7] relax
    // pinned staging for the host entry points
    DevBuf h_in, h_out;
    // po_score_pairs_batch: two buffer sets so that the copies of one chunk overlap the kernels of another
    struct PipeSlot {
        DevBuf buf[7];  // a, b, het, hairpin_a, hairpin_b, tm_a, tm_b
        cudaEvent_t in_done = nullptr, comp_done = nullptr, out_done = nullptr;
    } pipe[2];
    cudaStream_t s_in = nullptr, s_out = nullptr;
    int fast_cap = 0;  // 0 = choose by environment / default
};

static std::string g_init_error;

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);               \
            return PO_E_CUDA;                                                             \
        }                                                                                 \
    } while (0)

static int ensure(po_ctx* ctx, DevBuf& buf, size_t bytes) {
    if (bytes <= buf.cap) return PO_OK;
    if (buf.p) CK(cudaFree(buf.p));
    buf.p = nullptr;
    buf.cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    CK(cudaMalloc(&buf.p, want));
    buf.cap = want;
    return PO_OK;
}

static void resolve_conditions(PoConsts* K, double mv, double dv, double dntp, double dna, double temp_c,
                               int max_loop, int max_nn) {
    // thal saltCorrectS(): if (dv <= 0) dntp = dv
    double d = dntp;
    if (dv <= 0) d = dv;
    K->salt_corr = 0.368 * ((log((mv + 120 * (sqrt(fmax(0.0, dv - d)))) / 1000)));
    K->rc_sym = 1.9872 * log(dna / 1000000000.0);
    K->rc_nonsym = 1.9872 * log(dna / 4000000000.0);
    K->temp_k = temp_c + 273.15;
    // oligotm divalent_to_monovalent()
    double dv2 = dv, dn2 = dntp;
    if (dv2 == 0) dn2 = 0;
    if (dv2 < dn2) dv2 = dn2;
    const double mv_eff = mv + 120 * (sqrt(dv2 - dn2));
    K->tm_log_mv = log(mv_eff / 1000.0);
    K->tm_log10_mv = log10(mv_eff / 1000.0);
    K->tm_rlog_sym = 1.987 * log(dna / 1000000000.0);
    K->tm_rlog_non = 1.987 * log(dna / 4000000000.0);
    K->max_loop = max_loop;
    K->max_nn_length = max_nn;
}

extern "C" int po_init(int device, const char* param_dir, po_ctx** out) {
    if (!out) return PO_E_INVALID;
    *out = nullptr;
    PoTables host;
    std::string err;
    if (!po_load_tables(param_dir, &host, &err)) {
        g_init_error = err;
        return PO_E_PARAMS;
    }
    po_ctx* ctx = new po_ctx();
    ctx->device = device;
    auto fail = [&](const char* what, cudaError_t e) {
        g_init_error = std::string(what) + ": " + cudaGetErrorString(e);
        delete ctx;
        return PO_E_CUDA;
    };
    cudaError_t e;
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail("cudaSetDevice", e);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail("cudaGetDeviceProperties", e);
    ctx->num_sms = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess)
        return fail("cudaStreamCreate", e);
    if ((e = cudaMalloc(&ctx->d_tables, sizeof(PoTables))) != cudaSuccess) return fail("cudaMalloc tables", e);
    if ((e = cudaMemcpy(ctx->d_tables, &host, sizeof(PoTables), cudaMemcpyHostToDevice)) != cudaSuccess)
        return fail("cudaMemcpy tables", e);
    if ((e = cudaMalloc(&ctx->d_ctr, 8 * sizeof(unsigned long long))) != cudaSuccess) return fail("cudaMalloc ctr", e);
    resolve_conditions(&ctx->K, 50.0, 0.0, 0.8, 50.0, 37.0, 30, 60);
    const char* fc = getenv("PO_FAST_CAP");
    if (fc) ctx->fast_cap = atoi(fc);
    *out = ctx;
    return PO_OK;
}

extern "C" void po_destroy(po_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    DevBuf* bufs[] = {&ctx->a, &ctx->b, &ctx->out, &ctx->out2, &ctx->out3, &ctx->tmv, &ctx->gcv, &ctx->props, &ctx->overflow,
                      &ctx->hp_cells};
    for (DevBuf* b : bufs)
        if (b->p) cudaFree(b->p);
    for (DevBuf& sb : ctx->scratch)
        if (sb.p) cudaFree(sb.p);
    for (DevBuf& sb : ctx->p3)
        if (sb.p) cudaFree(sb.p);
    for (DevBuf& sb : ctx->kd)
        if (sb.p) cudaFree(sb.p);
    if (ctx->h_in.p) cudaFreeHost(ctx->h_in.p);
    if (ctx->h_out.p) cudaFreeHost(ctx->h_out.p);
    for (auto& ps : ctx->pipe) {
        for (DevBuf& sb : ps.buf)
            if (sb.p) cudaFree(sb.p);
        if (ps.in_done) cudaEventDestroy(ps.in_done);
        if (ps.comp_done) cudaEventDestroy(ps.comp_done);
        if (ps.out_done) cudaEventDestroy(ps.out_done);
    }
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
    if (ctx->d_tables) cudaFree(ctx->d_tables);
    if (ctx->d_ctr) cudaFree(ctx->d_ctr);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char* po_last_error(const po_ctx* ctx) { return ctx ? ctx->err.c_str() : g_init_error.c_str(); }

extern "C" int po_set_conditions(po_ctx* ctx, double mv, double dv, double dntp, double dna_nM, double temp_c,
                                 int max_loop, int max_nn_length) {
    if (!ctx) return PO_E_INVALID;
    if (mv < 0 || dv < 0 || dntp < 0 || dna_nM <= 0 || max_loop < 0 || max_loop > 30) {
        ctx->err = "po_set_conditions: argument out of range (max_loop must be 0..30)";
        return PO_E_INVALID;
    }
    resolve_conditions(&ctx->K, mv, dv, dntp, dna_nM, temp_c, max_loop, max_nn_length);
    return PO_OK;
}

extern "C" int64_t po_thal_item_count(const po_ctx* ctx) { return ctx ? ctx->thal_items : 0; }

extern "C" int64_t po_launch_count(const po_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ============================================================================
// packing (host)
// ============================================================================
extern "C" int po_pack(const char* ascii, const int64_t* offsets, int64_t n, po_oligo* packed) {
    if (!ascii || !offsets || !packed || n < 0) return PO_E_INVALID;
    for (int64_t k = 0; k < n; ++k) {
        const int64_t len = offsets[k + 1] - offsets[k];
        const char* s = ascii + offsets[k];
        uint32_t w[4] = {0, 0, 0, 0};
        uint32_t meta = 0;
        const int64_t m = len > PO_MAX_OLIGO_LEN ? PO_MAX_OLIGO_LEN : len;
        for (int64_t p = 0; p < m; ++p) {
            uint32_t code = 0;
            switch (s[p]) {
                case 'A': case 'a': code = 0; break;
                case 'C': case 'c': code = 1; break;
                case 'G': case 'g': code = 2; break;
                case 'T': case 't': code = 3; break;
                case 'N': case 'n': case 'R': case 'r': case 'Y': case 'y': case 'S': case 's':
                case 'W': case 'w': case 'K': case 'k': case 'M': case 'm': case 'B': case 'b':
                case 'D': case 'd': case 'H': case 'h': case 'V': case 'v': case 'U': case 'u':
                    meta |= PO_OLIGO_HAS_N; break;
                default: meta |= PO_OLIGO_BAD_CHAR; break;
            }
            w[p >> 4] |= code << (2 * (p & 15));
        }
        if (len > PO_MAX_OLIGO_LEN) meta |= PO_OLIGO_BAD_CHAR;  // too long for a record: never evaluated
        meta |= (uint32_t)m;
        w[3] |= meta << 24;
        packed[k].w[0] = w[0];
        packed[k].w[1] = w[1];
        packed[k].w[2] = w[2];
        packed[k].w[3] = w[3];
    }
    return PO_OK;
}

extern "C" int po_unpack(const po_oligo* packed, int64_t n, char* out) {
    if (!packed || !out || n < 0) return PO_E_INVALID;
    static const char L[4] = {'A', 'C', 'G', 'T'};
    for (int64_t k = 0; k < n; ++k) {
        const uint32_t* w = packed[k].w;
        const int len = (w[3] >> 24) & PO_OLIGO_LEN_MASK;
        char* o = out + 61 * k;
        for (int p = 0; p < len; ++p) o[p] = L[(w[p >> 4] >> (2 * (p & 15))) & 3];
        o[len] = 0;
    }
    return PO_OK;
}

// ============================================================================
// launchers
// ============================================================================
template <int CAP, int WARPS, bool COUNT, bool HAIRPIN, int PHASE = 0>
static int launch_thal(po_ctx* ctx, cudaStream_t st, const uint4* A, const uint4* B, long long n, po_thal_out* out,
                       unsigned long long* work, const long long* list, const unsigned long long* list_n,
                       long long* overflow, unsigned long long* overflow_n, unsigned long long* relax,
                       int end_mode = 0, po_thal_out* out2 = nullptr, long long first = 0, double2* cells = nullptr) {
    using Ws = typename ThalWs<CAP, COUNT, HAIRPIN>::type;
    const size_t smem = sizeof(SmemTables) + sizeof(SmemExtra) + WARPS * sizeof(Ws);
    auto kern = k_thal<CAP, WARPS, COUNT, HAIRPIN, PHASE>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem));
    if (per_sm < 1) {
        ctx->err = "thal kernel does not fit on an SM";
        return PO_E_CUDA;
    }
    long long blocks = (long long)ctx->num_sms * per_sm;
    if (!list) {
        const long long need = (n + WARPS - 1) / WARPS;
        if (need < blocks) blocks = need;
    }
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, WARPS * 32, smem, st>>>(A, B, n, out, ctx->d_tables, ctx->K, work, list, list_n, overflow,
                                                     overflow_n, relax, end_mode, out2, first, cells);
    ctx->launches++;
    CK(cudaGetLastError());
    return PO_OK;
}

// Capacity tiers.  The first pass keeps the per-warp shared-memory workspace
// small (random 18-30 nt pairs have ~144 finite cells, <0.02% exceed 256) so
// that 26-28 warps are resident per SM; items that exceed a tier's capacity are
// queued for the next one (reporter-tailed KASP pairs, then low-complexity /
// 60-mer inputs whose DP table is dense).  Counters: ctr[0] work1, ctr[1]
// list1_n, ctr[2] work2, ctr[3] list2_n, ctr[4] work3, ctr[5] list3_n, ctr[6] work4, ctr[7] relaxations.
// d_out2 (dimer types only): ANY result to d_out and END1 result to d_out2 from ONE table fill
// (libprimer3 asks for both on every oligo and pair; the fill is > 90 % of the work).
template <bool COUNT>
static int thal_dev(po_ctx* ctx, int type, const void* d_a, const void* d_b, int64_t n, void* d_out, cudaStream_t st,
                    void* d_out2 = nullptr) {
    if (n == 0) return PO_OK;
    ctx->thal_items += d_out2 ? 2 * n : n;
    int rc = ensure(ctx, ctx->overflow, 2 * sizeof(long long) * (size_t)n);
    if (rc) return rc;
    unsigned long long* ctr = ctx->d_ctr;
    CK(cudaMemsetAsync(ctr, 0, 7 * sizeof(unsigned long long), st));
    const uint4* A = (const uint4*)d_a;
    const uint4* B = (const uint4*)d_b;
    po_thal_out* out = (po_thal_out*)d_out;
    long long* l1 = (long long*)ctx->overflow.p;
    long long* l2 = l1 + n;
    unsigned long long* rx = ctr + 7;
    if (type == PO_THAL_HAIRPIN) {
        if constexpr (COUNT) {
            rc = launch_thal<128, PO_HAIRPIN_WARPS, COUNT, true>(ctx, st, A, nullptr, n, out, ctr + 0, nullptr, nullptr, l1, ctr + 1, rx);
        } else {
            // first tier in two kernels per sub-batch (fill, finish): the cell values cross HBM in between
            const long long chunk = PO_HAIRPIN_CHUNK;
            const long long m0 = n < chunk ? n : chunk;
            if ((rc = ensure(ctx, ctx->hp_cells, sizeof(double2) * (size_t)PO_HAIRPIN_XFER(PO_HAIRPIN32_CAP) * (size_t)m0))) return rc;
            double2* cells = (double2*)ctx->hp_cells.p;
            for (long long s = 0; s < n; s += chunk) {
                const long long m = (n - s) < chunk ? (n - s) : chunk;
                CK(cudaMemsetAsync(ctr + 0, 0, sizeof(unsigned long long), st));
                if ((rc = launch_thal<128, PO_HAIRPIN_WARPS, false, true, 1>(ctx, st, A, nullptr, m, out, ctr + 0, nullptr, nullptr,
                                                                             l1, ctr + 1, rx, 0, nullptr, s, cells))) return rc;
                CK(cudaMemsetAsync(ctr + 0, 0, sizeof(unsigned long long), st));
                if ((rc = launch_thal<128, PO_HAIRPIN_WARPS, false, true, 2>(ctx, st, A, nullptr, m, out, ctr + 0, nullptr, nullptr,
                                                                             l1, ctr + 1, rx, 0, nullptr, s, cells))) return rc;
            }
        }
        if (rc) return rc;
        return launch_thal<1024, 8, COUNT, true>(ctx, st, A, nullptr, n, out, ctr + 2, l1, ctr + 1, l2, ctr + 3, rx);
    }
    if (type != PO_THAL_ANY && type != PO_THAL_END1 && type != PO_THAL_END2) {
        ctx->err = "po_thal: unknown alignment type";
        return PO_E_INVALID;
    }
    const int em = type == PO_THAL_END1 ? 1 : type == PO_THAL_END2 ? 2 : 0;
    po_thal_out* out2 = (po_thal_out*)d_out2;
    if (out2 && type != PO_THAL_ANY) {
        ctx->err = "po_thal: the fused ANY + END1 form needs type ANY";
        return PO_E_INVALID;
    }
    rc = launch_thal<256, PO_DIMER_WARPS, COUNT, false>(ctx, st, A, B, n, out, ctr + 0, nullptr, nullptr, l1, ctr + 1, rx, em, out2);
    if (rc) return rc;
    // reporter-tailed KASP pairs (<= 47 x 26 nt, ~230-400 finite cells) mostly land in the second tier
    if ((rc = launch_thal<384, 18, COUNT, false>(ctx, st, A, B, n, out, ctr + 2, l1, ctr + 1, l2, ctr + 3, rx, em, out2))) return rc;
    // l1 is fully consumed by now: its storage takes the third tier's overflow (own counter)
    if ((rc = launch_thal<768, 10, COUNT, false>(ctx, st, A, B, n, out, ctr + 4, l2, ctr + 3, l1, ctr + 5, rx, em, out2))) return rc;
    return launch_thal<3600, 2, COUNT, false>(ctx, st, A, B, n, out, ctr + 6, l1, ctr + 5, l2, ctr + 3, rx, em, out2);
}

extern "C" int po_thal_batch_dev(po_ctx* ctx, int type, const void* d_a, const void* d_b, int64_t n, void* d_out,
                                 void* stream) {
    if (!ctx || !d_a || !d_out || n < 0) return PO_E_INVALID;
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    return thal_dev<false>(ctx, type, d_a, d_b, n, d_out, st);
}

extern "C" int po_tm_batch_dev(po_ctx* ctx, const void* d_oligos, int64_t n, void* d_tm, void* d_gc, void* stream) {
    if (!ctx || !d_oligos || !d_tm || n < 0) return PO_E_INVALID;
    if (n == 0) return PO_OK;
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    const int threads = 256;
    const long long blocks = (n + threads - 1) / threads;
    k_tm<<<(unsigned)blocks, threads, 0, st>>>((const uint4*)d_oligos, n, (double*)d_tm, (double*)d_gc, ctx->K);
    ctx->launches++;
    CK(cudaGetLastError());
    return PO_OK;
}

extern "C" int po_sync(po_ctx* ctx, void* stream) {
    if (!ctx) return PO_E_INVALID;
    CK(cudaStreamSynchronize(stream ? (cudaStream_t)stream : ctx->stream));
    return PO_OK;
}

// ---- host-buffer entry points -------------------------------------------------

extern "C" int po_tm_batch(po_ctx* ctx, const po_oligo* oligos, int64_t n, double* tm, double* gc) {
    if (!ctx || !oligos || !tm || n < 0) return PO_E_INVALID;
    if (n == 0) return PO_OK;
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure(ctx, ctx->a, sizeof(po_oligo) * (size_t)n))) return rc;
    if ((rc = ensure(ctx, ctx->tmv, sizeof(double) * (size_t)n))) return rc;
    if (gc && (rc = ensure(ctx, ctx->gcv, sizeof(double) * (size_t)n))) return rc;
    CK(cudaMemcpyAsync(ctx->a.p, oligos, sizeof(po_oligo) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = po_tm_batch_dev(ctx, ctx->a.p, n, ctx->tmv.p, gc ? ctx->gcv.p : nullptr, ctx->stream))) return rc;
    CK(cudaMemcpyAsync(tm, ctx->tmv.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (gc) CK(cudaMemcpyAsync(gc, ctx->gcv.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PO_OK;
}

extern "C" int po_thal_batch(po_ctx* ctx, int type, const po_oligo* a, const po_oligo* b, int64_t n,
                             po_thal_out* out) {
    if (!ctx || !a || !out || n < 0) return PO_E_INVALID;
    if (n == 0) return PO_OK;
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure(ctx, ctx->a, sizeof(po_oligo) * (size_t)n))) return rc;
    if (b && (rc = ensure(ctx, ctx->b, sizeof(po_oligo) * (size_t)n))) return rc;
    if ((rc = ensure(ctx, ctx->out, sizeof(po_thal_out) * (size_t)n))) return rc;
    CK(cudaMemcpyAsync(ctx->a.p, a, sizeof(po_oligo) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    if (b) CK(cudaMemcpyAsync(ctx->b.p, b, sizeof(po_oligo) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = thal_dev<false>(ctx, type, ctx->a.p, b ? ctx->b.p : nullptr, n, ctx->out.p, ctx->stream))) return rc;
    CK(cudaMemcpyAsync(out, ctx->out.p, sizeof(po_thal_out) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PO_OK;
}

extern "C" int po_score_pairs_batch_dev(po_ctx* ctx, const void* d_a, const void* d_b, int64_t n, void* d_het,
                                        void* d_hairpin_a, void* d_hairpin_b, void* d_tm_a, void* d_tm_b,
                                        void* stream) {
    if (!ctx || !d_a || !d_b || n < 0) return PO_E_INVALID;
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    int rc;
    if (d_het && (rc = thal_dev<false>(ctx, PO_THAL_ANY, d_a, d_b, n, d_het, st))) return rc;
    if (d_hairpin_a && (rc = thal_dev<false>(ctx, PO_THAL_HAIRPIN, d_a, nullptr, n, d_hairpin_a, st))) return rc;
    if (d_hairpin_b && (rc = thal_dev<false>(ctx, PO_THAL_HAIRPIN, d_b, nullptr, n, d_hairpin_b, st))) return rc;
    if (d_tm_a && (rc = po_tm_batch_dev(ctx, d_a, n, d_tm_a, nullptr, st))) return rc;
    if (d_tm_b && (rc = po_tm_batch_dev(ctx, d_b, n, d_tm_b, nullptr, st))) return rc;
    return PO_OK;
}

extern "C" int po_score_pairs_batch(po_ctx* ctx, const po_oligo* a, const po_oligo* b, int64_t n, po_thal_out* het,
                                    po_thal_out* hairpin_a, po_thal_out* hairpin_b, double* tm_a, double* tm_b) {
    if (!ctx || !a || !b || n < 0) return PO_E_INVALID;
    if (n == 0) return PO_OK;
    CK(cudaSetDevice(ctx->device));
    // Chunks of 2^20 pairs through two buffer sets: the H2D copy of chunk k+1 (copy-in stream) and the
    // D2H copy of chunk k-1 (copy-out stream) run while chunk k is scored (context stream).  With pinned
    // host buffers the copies disappear behind the kernels; pageable buffers still give correct results.
    const int64_t chunk = 1 << 20;
    const int64_t m0 = n < chunk ? n : chunk;
    int rc;
    if (!ctx->s_in) {
        CK(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
        for (auto& ps : ctx->pipe) {
            CK(cudaEventCreateWithFlags(&ps.in_done, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ps.comp_done, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ps.out_done, cudaEventDisableTiming));
        }
    }
    void* const host_out[5] = {het, hairpin_a, hairpin_b, tm_a, tm_b};
    const size_t out_size[5] = {sizeof(po_thal_out), sizeof(po_thal_out), sizeof(po_thal_out), sizeof(double), sizeof(double)};
    const int n_slots = n > chunk ? 2 : 1;
    for (int sl = 0; sl < n_slots; ++sl) {
        auto& ps = ctx->pipe[sl];
        if ((rc = ensure(ctx, ps.buf[0], sizeof(po_oligo) * (size_t)m0))) return rc;
        if ((rc = ensure(ctx, ps.buf[1], sizeof(po_oligo) * (size_t)m0))) return rc;
        for (int o = 0; o < 5; ++o)
            if (host_out[o] && (rc = ensure(ctx, ps.buf[2 + o], out_size[o] * (size_t)m0))) return rc;
    }
    // the scratch thal_dev grows on demand, sized up front: no allocation (= device-wide sync) inside the pipeline
    if ((rc = ensure(ctx, ctx->overflow, 2 * sizeof(long long) * (size_t)m0))) return rc;
    {
        const long long hm = m0 < (long long)PO_HAIRPIN_CHUNK ? m0 : (long long)PO_HAIRPIN_CHUNK;
        if ((hairpin_a || hairpin_b) &&
            (rc = ensure(ctx, ctx->hp_cells, sizeof(double2) * (size_t)PO_HAIRPIN_XFER(PO_HAIRPIN32_CAP) * (size_t)hm)))
            return rc;
    }
    cudaStream_t st = ctx->stream;
    int64_t k = 0;
    for (int64_t s = 0; s < n; s += chunk, ++k) {
        const int64_t m = (n - s) < chunk ? (n - s) : chunk;
        auto& ps = ctx->pipe[k & 1];
        if (k >= 2) CK(cudaStreamWaitEvent(ctx->s_in, ps.comp_done, 0));  // the slot's previous chunk has been read
        CK(cudaMemcpyAsync(ps.buf[0].p, a + s, sizeof(po_oligo) * (size_t)m, cudaMemcpyHostToDevice, ctx->s_in));
        CK(cudaMemcpyAsync(ps.buf[1].p, b + s, sizeof(po_oligo) * (size_t)m, cudaMemcpyHostToDevice, ctx->s_in));
        CK(cudaEventRecord(ps.in_done, ctx->s_in));
        CK(cudaStreamWaitEvent(st, ps.in_done, 0));
        if (k >= 2) CK(cudaStreamWaitEvent(st, ps.out_done, 0));          // ... and its results have left the slot
        if ((rc = po_score_pairs_batch_dev(ctx, ps.buf[0].p, ps.buf[1].p, m, het ? ps.buf[2].p : nullptr,
                                           hairpin_a ? ps.buf[3].p : nullptr, hairpin_b ? ps.buf[4].p : nullptr,
                                           tm_a ? ps.buf[5].p : nullptr, tm_b ? ps.buf[6].p : nullptr, st)))
            return rc;
        CK(cudaEventRecord(ps.comp_done, st));
        CK(cudaStreamWaitEvent(ctx->s_out, ps.comp_done, 0));
        for (int o = 0; o < 5; ++o)
            if (host_out[o])
                CK(cudaMemcpyAsync((char*)host_out[o] + out_size[o] * (size_t)s, ps.buf[2 + o].p, out_size[o] * (size_t)m,
                                   cudaMemcpyDeviceToHost, ctx->s_out));
        CK(cudaEventRecord(ps.out_done, ctx->s_out));
    }
    CK(cudaStreamSynchronize(ctx->s_out));
    CK(cudaStreamSynchronize(st));
    return PO_OK;
}

extern "C" int po_thal_count_relaxations(po_ctx* ctx, int type, const po_oligo* a, const po_oligo* b, int64_t n,
                                         int64_t* relaxations) {
    if (!ctx || !a || !relaxations || n < 0) return PO_E_INVALID;
    *relaxations = 0;
    if (n == 0) return PO_OK;
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure(ctx, ctx->a, sizeof(po_oligo) * (size_t)n))) return rc;
    if (b && (rc = ensure(ctx, ctx->b, sizeof(po_oligo) * (size_t)n))) return rc;
    if ((rc = ensure(ctx, ctx->out, sizeof(po_thal_out) * (size_t)n))) return rc;
    CK(cudaMemcpyAsync(ctx->a.p, a, sizeof(po_oligo) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    if (b) CK(cudaMemcpyAsync(ctx->b.p, b, sizeof(po_oligo) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_ctr + 7, 0, sizeof(unsigned long long), ctx->stream));
    if ((rc = thal_dev<true>(ctx, type, ctx->a.p, b ? ctx->b.p : nullptr, n, ctx->out.p, ctx->stream))) return rc;
    unsigned long long v = 0;
    CK(cudaMemcpyAsync(&v, ctx->d_ctr + 7, sizeof(v), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *relaxations = (int64_t)v;
    return PO_OK;
}

extern "C" int po_oligo_props_batch(po_ctx* ctx, const po_oligo* oligos, int64_t n, const po_valid_params* valid,
                                    po_oligo_props* out) {
    if (!ctx || !oligos || !out || n < 0) return PO_E_INVALID;
    if (n == 0) return PO_OK;
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure(ctx, ctx->a, sizeof(po_oligo) * (size_t)n))) return rc;
    if ((rc = ensure(ctx, ctx->out, sizeof(po_thal_out) * (size_t)n))) return rc;
    if ((rc = ensure(ctx, ctx->out2, sizeof(po_thal_out) * (size_t)n))) return rc;
    if ((rc = ensure(ctx, ctx->props, sizeof(po_oligo_props) * (size_t)n))) return rc;
    CK(cudaMemcpyAsync(ctx->a.p, oligos, sizeof(po_oligo) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = thal_dev<false>(ctx, PO_THAL_HAIRPIN, ctx->a.p, nullptr, n, ctx->out.p, ctx->stream))) return rc;
    if ((rc = thal_dev<false>(ctx, PO_THAL_ANY, ctx->a.p, nullptr, n, ctx->out2.p, ctx->stream))) return rc;
    po_valid_params V;
    memset(&V, 0, sizeof V);
    if (valid) V = *valid;
    const int threads = 256;
    const long long blocks = (n + threads - 1) / threads;
    k_props<<<(unsigned)blocks, threads, 0, ctx->stream>>>((const uint4*)ctx->a.p, n, (const po_thal_out*)ctx->out.p,
                                                          (const po_thal_out*)ctx->out2.p, ctx->K, V, valid ? 1 : 0,
                                                          (po_oligo_props*)ctx->props.p);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, ctx->props.p, sizeof(po_oligo_props) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PO_OK;
}

extern "C" int po_measure_fp64_peak(po_ctx* ctx, double* tflops) {
    if (!ctx || !tflops) return PO_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    const int threads = 256, blocks = ctx->num_sms * 8, iters = 1 << 16;
    int rc;
    if ((rc = ensure(ctx, ctx->tmv, sizeof(double) * (size_t)threads * blocks))) return rc;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0, ctx->stream));
        k_fp64_peak<<<blocks, threads, 0, ctx->stream>>>((double*)ctx->tmv.p, iters);
        ctx->launches++;
        CK(cudaEventRecord(e1, ctx->stream));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 8.0 * (double)iters * threads * blocks;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    return PO_OK;
}

// ============================================================================
// KASP enumeration and pair ranking
// ============================================================================
extern "C" int po_kasp_enumerate_batch(po_ctx* ctx, const po_kasp_marker* markers, int64_t n, int min_size,
                                       int max_size, po_oligo* ref_out, po_oligo* alt_out) {
    if (!ctx || !markers || !ref_out || !alt_out || n < 0) return PO_E_INVALID;
    if (min_size < 1 || max_size < min_size || max_size > PO_MAX_OLIGO_LEN) {
        ctx->err = "po_kasp_enumerate_batch: need 1 <= min_size <= max_size <= 60";
        return PO_E_INVALID;
    }
    if (n == 0) return PO_OK;
    CK(cudaSetDevice(ctx->device));
    const int slots = 2 * (max_size - min_size + 1);
    const size_t total = (size_t)n * slots;
    int rc;
    if ((rc = ensure(ctx, ctx->scratch[0], sizeof(po_kasp_marker) * (size_t)n))) return rc;
    if ((rc = ensure(ctx, ctx->a, sizeof(po_oligo) * total))) return rc;
    if ((rc = ensure(ctx, ctx->b, sizeof(po_oligo) * total))) return rc;
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(ctx->scratch[0].p, markers, sizeof(po_kasp_marker) * (size_t)n, cudaMemcpyHostToDevice, st));
    const int threads = 256;
    const long long blocks = ((long long)total + threads - 1) / threads;
    k_kasp_enumerate<<<(unsigned)blocks, threads, 0, st>>>((const po_kasp_marker*)ctx->scratch[0].p, n, min_size,
                                                          max_size, (uint4*)ctx->a.p, (uint4*)ctx->b.p);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ref_out, ctx->a.p, sizeof(po_oligo) * total, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(alt_out, ctx->b.p, sizeof(po_oligo) * total, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return PO_OK;
}

extern "C" int po_rank_pairs_batch(po_ctx* ctx, const po_rank_in* in, const po_rank_out* out) {
    if (!ctx || !in || !out) return PO_E_INVALID;
    const int64_t n = in->n_pairs, nm = in->n_markers;
    const bool kasp = in->assay == PO_ASSAY_KASP, hrm = in->assay == PO_ASSAY_HRM;
    if (n < 0 || nm < 0 || in->top_n < 1 || !in->f || !in->r || !in->seg || !in->n_offtargets || !in->max_aaf ||
        !in->max_indel || !out->goodness || !out->qbits || !out->sel || !out->n_sel || (kasp && (!in->a || !in->dir)) ||
        (hrm && !in->hrm_delta) || in->assay < 0 || in->assay > 2) {
        ctx->err = "po_rank_pairs_batch: missing or inconsistent arguments";
        return PO_E_INVALID;
    }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int threads = 256;
    int rc;
    enum { S_F = 0, S_R, S_A, S_TF, S_TR, S_TA, S_MISC, S_AAF, S_SEL, S_GOOD, S_TAILA, S_TAILB };
    DevBuf* sb = ctx->scratch;
    const size_t nn = (size_t)(n > 0 ? n : 1);
    if ((rc = ensure(ctx, sb[S_F], sizeof(po_oligo) * nn))) return rc;
    if ((rc = ensure(ctx, sb[S_R], sizeof(po_oligo) * nn))) return rc;
    if ((rc = ensure(ctx, sb[S_A], sizeof(po_oligo) * nn))) return rc;
    if ((rc = ensure(ctx, sb[S_TF], sizeof(double) * nn))) return rc;
    if ((rc = ensure(ctx, sb[S_TR], sizeof(double) * nn))) return rc;
    if ((rc = ensure(ctx, sb[S_TA], sizeof(double) * nn))) return rc;
    // misc: n_off (i32), indel (i32), hrm_delta (f64), dir/sel (i8 x 3), seg (i64 x (nm+1)), n_sel (i32 x nm)
    const size_t off_noff = 0, off_ind = off_noff + 4 * nn, off_hd = off_ind + 4 * nn, off_sel8 = off_hd + 8 * nn,
                 off_seg = (off_sel8 + 3 * nn + 15) / 16 * 16, off_nsel = off_seg + 8 * (size_t)(nm + 1),
                 misc_bytes = off_nsel + 4 * (size_t)(nm > 0 ? nm : 1);
    if ((rc = ensure(ctx, sb[S_MISC], misc_bytes))) return rc;
    if ((rc = ensure(ctx, sb[S_AAF], sizeof(double) * 3 * nn))) return rc;
    if ((rc = ensure(ctx, sb[S_SEL], sizeof(long long) * (size_t)(nm > 0 ? nm : 1) * in->top_n))) return rc;
    if ((rc = ensure(ctx, sb[S_GOOD], sizeof(int32_t) * 2 * nn))) return rc;
    if ((rc = ensure(ctx, sb[S_TAILA], sizeof(po_oligo) * nn))) return rc;
    if ((rc = ensure(ctx, sb[S_TAILB], sizeof(po_oligo) * nn))) return rc;
    if ((rc = ensure(ctx, ctx->out, sizeof(po_thal_out) * nn))) return rc;
    if ((rc = ensure(ctx, ctx->out2, sizeof(po_thal_out) * nn))) return rc;
    char* misc = (char*)sb[S_MISC].p;
    if (n > 0) {
        CK(cudaMemcpyAsync(sb[S_F].p, in->f, sizeof(po_oligo) * n, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(sb[S_R].p, in->r, sizeof(po_oligo) * n, cudaMemcpyHostToDevice, st));
        if (kasp) CK(cudaMemcpyAsync(sb[S_A].p, in->a, sizeof(po_oligo) * n, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(misc + off_noff, in->n_offtargets, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(misc + off_ind, in->max_indel, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
        if (hrm) CK(cudaMemcpyAsync(misc + off_hd, in->hrm_delta, 8 * (size_t)n, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(sb[S_AAF].p, in->max_aaf, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, st));
    }
    CK(cudaMemcpyAsync(misc + off_seg, in->seg, 8 * (size_t)(nm + 1), cudaMemcpyHostToDevice, st));
    const uint4* dF = (const uint4*)sb[S_F].p;
    const uint4* dR = (const uint4*)sb[S_R].p;
    const uint4* dA = (const uint4*)sb[S_A].p;
    const long long blocks = (n + threads - 1) / threads;
    if (n > 0) {
        // calcTm of the untailed primers (Primer.calc_thermals inside check_heterodimerization)
        if ((rc = po_tm_batch_dev(ctx, dF, n, sb[S_TF].p, nullptr, st))) return rc;
        if ((rc = po_tm_batch_dev(ctx, dR, n, sb[S_TR].p, nullptr, st))) return rc;
        if (kasp && (rc = po_tm_batch_dev(ctx, dA, n, sb[S_TA].p, nullptr, st))) return rc;
        const po_thal_out* het1 = (const po_thal_out*)ctx->out.p;
        const po_thal_out* het2 = (const po_thal_out*)ctx->out2.p;
        if (kasp) {
            // add_reporter_dyes: dir F -> F gets dye1, A dye2, R none; dir R -> R gets dye1, A dye2, F none.
            // heterodimers: (F', R') and, by direction, (A', R') or (F', A').
            std::vector<int8_t> sel(3 * (size_t)n);
            for (int64_t k = 0; k < n; ++k) {
                const bool dirF = in->dir[k] == 0;
                sel[k] = dirF ? 0 : -1;                  // tail of F
                sel[(size_t)n + k] = dirF ? -1 : 0;      // tail of R
                sel[2 * (size_t)n + k] = 1;              // tail of A
            }
            CK(cudaMemcpyAsync(misc + off_sel8, sel.data(), 3 * (size_t)n, cudaMemcpyHostToDevice, st));
            CK(cudaStreamSynchronize(st));  // sel is a stack-lifetime staging buffer
            uint4 t0, t1;
            memcpy(&t0, &in->reporters[0], 16);
            memcpy(&t1, &in->reporters[1], 16);
            uint4* Ft = (uint4*)sb[S_TAILA].p;
            uint4* Rt = (uint4*)sb[S_TAILB].p;
            uint4* At = (uint4*)ctx->a.p;
            if ((rc = ensure(ctx, ctx->a, sizeof(po_oligo) * nn))) return rc;
            if ((rc = ensure(ctx, ctx->b, sizeof(po_oligo) * 2 * nn))) return rc;
            At = (uint4*)ctx->a.p;
            const int8_t* s8 = (const int8_t*)(misc + off_sel8);
            k_concat_tail<<<(unsigned)blocks, threads, 0, st>>>(dF, s8, t0, t1, n, Ft);
            k_concat_tail<<<(unsigned)blocks, threads, 0, st>>>(dR, s8 + n, t0, t1, n, Rt);
            k_concat_tail<<<(unsigned)blocks, threads, 0, st>>>(dA, s8 + 2 * n, t0, t1, n, At);
            ctx->launches += 3;
            CK(cudaGetLastError());
            if ((rc = thal_dev<false>(ctx, PO_THAL_ANY, Ft, Rt, n, ctx->out.p, st))) return rc;
            // second pair: first strand A' (dir F) or F' (dir R); second strand R' (dir F) or A' (dir R)
            std::vector<po_oligo> dummy;
            uint4* X1 = (uint4*)ctx->b.p;
            uint4* X2 = X1 + n;
            k_select_pair<<<(unsigned)blocks, threads, 0, st>>>(Ft, Rt, At, s8, n, X1, X2);
            ctx->launches++;
            CK(cudaGetLastError());
            if ((rc = thal_dev<false>(ctx, PO_THAL_ANY, X1, X2, n, ctx->out2.p, st))) return rc;
        } else {
            if ((rc = thal_dev<false>(ctx, PO_THAL_ANY, dF, dR, n, ctx->out.p, st))) return rc;
        }
        RankArgs ra;
        ra.assay = in->assay;
        ra.n = n;
        ra.tmF = (const double*)sb[S_TF].p;
        ra.tmR = (const double*)sb[S_TR].p;
        ra.tmA = (const double*)sb[S_TA].p;
        ra.het1 = het1;
        ra.het2 = het2;
        ra.n_off = (const int32_t*)(misc + off_noff);
        ra.aaf = (const double*)sb[S_AAF].p;
        ra.indel = (const int32_t*)(misc + off_ind);
        ra.hrm_delta = (const double*)(misc + off_hd);
        ra.tm_delta = in->tm_delta;
        ra.goodness = (int32_t*)sb[S_GOOD].p;
        ra.qbits = ra.goodness + n;
        k_score_pairs<<<(unsigned)blocks, threads, 0, st>>>(ra);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    if (nm > 0) {
        const long long tb = (nm * 32 + threads - 1) / threads;
        k_top_n<<<(unsigned)tb, threads, 0, st>>>((const int32_t*)sb[S_GOOD].p, (const double*)(misc + off_hd), hrm ? 1 : 0,
                                                  (const long long*)(misc + off_seg), nm, in->top_n,
                                                  (long long*)sb[S_SEL].p, (int32_t*)(misc + off_nsel));
        ctx->launches++;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(out->sel, sb[S_SEL].p, sizeof(long long) * (size_t)nm * in->top_n, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(out->n_sel, misc + off_nsel, 4 * (size_t)nm, cudaMemcpyDeviceToHost, st));
    }
    if (n > 0) {
        CK(cudaMemcpyAsync(out->goodness, sb[S_GOOD].p, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(out->qbits, (int32_t*)sb[S_GOOD].p + n, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
        if (out->tm) {
            k_gather_tm<<<(unsigned)blocks, threads, 0, st>>>((const double*)sb[S_TF].p, (const double*)sb[S_TR].p,
                                                             kasp ? (const double*)sb[S_TA].p : nullptr,
                                                             (const po_thal_out*)ctx->out.p,
                                                             kasp ? (const po_thal_out*)ctx->out2.p : nullptr, n,
                                                             (double*)sb[S_AAF].p, (double*)sb[S_TAILA].p);
            ctx->launches++;
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(out->tm, sb[S_AAF].p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, st));
            if (out->het_tm)
                CK(cudaMemcpyAsync(out->het_tm, sb[S_TAILA].p, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost, st));
        }
    }
    CK(cudaStreamSynchronize(st));
    return PO_OK;
}

// ============================================================================
// "next" rows: HRM product Tm and mutation annotation
// ============================================================================
extern "C" int po_tm_nn_batch(po_ctx* ctx, const char* ascii, const int64_t* offsets, int64_t n, double* tm) {
    if (!ctx || !ascii || !offsets || !tm || n < 0) return PO_E_INVALID;
    if (n == 0) return PO_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)offsets[n];
    int rc;
    if ((rc = ensure(ctx, ctx->scratch[0], bytes + 16))) return rc;
    if ((rc = ensure(ctx, ctx->scratch[1], sizeof(int64_t) * (size_t)(n + 1)))) return rc;
    if ((rc = ensure(ctx, ctx->tmv, sizeof(double) * (size_t)n))) return rc;
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(ctx->scratch[0].p, ascii, bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->scratch[1].p, offsets, sizeof(int64_t) * (size_t)(n + 1), cudaMemcpyHostToDevice, st));
    TmNNConsts K;
    K.log_mon = log((50 + 0 + 0 / 2.0) * 1e-3);        // salt_correction(Na=50, K=0, Tris=0), method 5
    K.r_log_k = 1.987 * (log((25 - (25 / 2.0)) * 1e-9));  // dnac1 = dnac2 = 25 nM, selfcomp False
    const int threads = 128;
    const long long blocks = (n + threads - 1) / threads;
    k_tm_nn<<<(unsigned)blocks, threads, 0, st>>>((const char*)ctx->scratch[0].p, (const long long*)ctx->scratch[1].p, n,
                                                 K, (double*)ctx->tmv.p);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(tm, ctx->tmv.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return PO_OK;
}

extern "C" int po_annotate_mutations_batch(po_ctx* ctx, const int64_t* mut_pos, const double* mut_aaf,
                                           const int32_t* mut_indel, const int64_t* mut_seg, const int64_t* pair_seg,
                                           int64_t n_markers, const int64_t* primer_start, const int64_t* primer_stop,
                                           int n_primers, double* max_aaf, int32_t* max_indel) {
    if (!ctx || !mut_seg || !pair_seg || !primer_start || !primer_stop || !max_aaf || !max_indel || n_markers < 0 ||
        (n_primers != 2 && n_primers != 3))
        return PO_E_INVALID;
    if (n_markers == 0) return PO_OK;
    const int64_t n_mut = mut_seg[n_markers], n_pairs = pair_seg[n_markers];
    if (n_pairs == 0) return PO_OK;
    if (n_mut > 0 && (!mut_pos || !mut_aaf || !mut_indel)) return PO_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int rc;
    DevBuf* sb = ctx->scratch;
    const size_t nm1 = (size_t)(n_markers + 1), nmu = (size_t)(n_mut > 0 ? n_mut : 1), np_ = (size_t)n_pairs;
    if ((rc = ensure(ctx, sb[0], 8 * nmu))) return rc;
    if ((rc = ensure(ctx, sb[1], 8 * nmu))) return rc;
    if ((rc = ensure(ctx, sb[2], 4 * nmu))) return rc;
    if ((rc = ensure(ctx, sb[3], 8 * nm1))) return rc;
    if ((rc = ensure(ctx, sb[4], 8 * nm1))) return rc;
    if ((rc = ensure(ctx, sb[5], 8 * 3 * np_))) return rc;
    if ((rc = ensure(ctx, sb[6], 8 * 3 * np_))) return rc;
    if ((rc = ensure(ctx, sb[7], 8 * 3 * np_))) return rc;
    if ((rc = ensure(ctx, sb[8], 4 * np_))) return rc;
    if (n_mut > 0) {
        CK(cudaMemcpyAsync(sb[0].p, mut_pos, 8 * (size_t)n_mut, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(sb[1].p, mut_aaf, 8 * (size_t)n_mut, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(sb[2].p, mut_indel, 4 * (size_t)n_mut, cudaMemcpyHostToDevice, st));
    }
    CK(cudaMemcpyAsync(sb[3].p, mut_seg, 8 * nm1, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(sb[4].p, pair_seg, 8 * nm1, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(sb[5].p, primer_start, 8 * 3 * np_, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(sb[6].p, primer_stop, 8 * 3 * np_, cudaMemcpyHostToDevice, st));
    const int threads = 128;
    const long long blocks = (n_pairs + threads - 1) / threads;
    k_annotate_mutations<<<(unsigned)blocks, threads, 0, st>>>(
        (const long long*)sb[0].p, (const double*)sb[1].p, (const int32_t*)sb[2].p, (const long long*)sb[3].p,
        (const long long*)sb[4].p, n_markers, (const long long*)sb[5].p, (const long long*)sb[6].p, n_primers, n_pairs,
        (double*)sb[7].p, (int32_t*)sb[8].p);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(max_aaf, sb[7].p, 8 * 3 * np_, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(max_indel, sb[8].p, 4 * np_, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return PO_OK;
}

#include "po_p3.cuh"
#include "po_kasp_design.cuh"
