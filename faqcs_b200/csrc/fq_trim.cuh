// fq_trim.cuh -- the per-read trim / filter / statistics kernel (trim_read, trim.cpp:225-551).
//
// Persistent grid, one 1024-thread CTA per SM.  A warp takes 32 reads at a time.  Phase 1 walks them one by one
// with the lanes striped over consecutive base positions (lane l holds positions l, l+32, ...): reads of up to
// 320 bases are register resident (phase1<K>, K chunks of 32 bases, branch-free: predicated byte loads, two
// shared-memory LUT reads and two shared-memory reductions per base); longer reads take process_generic, which
// re-reads global memory chunk by chunk.  The per-read scalar work (window, quality trimming, filters, bins,
// verdict) then runs one lane per read, and the rare per-base follow-ups (bases trimmed away, discarded reads,
// dinucleotide counts, exact N runs) are ballot-driven cooperative passes.
//
// Statistics go to shared-memory privatised histograms stored transposed ([column][position], rows % 32 == 0) so
// that a warp's 32 consecutive positions always fall into 32 distinct banks; they are merged into the global u64
// block once per CTA.  The post-trim matrices are accumulated as "pre minus removed" (see StatsLayout), so an
// untrimmed surviving read costs one histogram update per base and matrix.  The kernel is instantiated per phase-1
// width and for the default option set (k_trim<KSEL, PLAIN>): it is issue-bound and instruction-cache sensitive.
#pragma once
#include "fq_common.cuh"

namespace fq {

struct TrimArgs {
    const uint8_t *raw[2];
    const Rec *rec[2];
    const uint2 *adp[2];        // adapter verdict {start, length} or nullptr
    const int32_t *adp_best[2];
    uint2 *res[2];              // {offset_5, length | flags << 24}
    fq_read_result *dbg[2];     // optional
    uint32_t n_rec;
    uint32_t n_mates;
    unsigned long long *stats;
    StatsLayout L;
    StatsRows *rows;
    BatchInfo *info;
    uint32_t smem_rows;         // rows held in shared memory (multiple of 32, <= L.rows)
    uint32_t comp_key_len;      // reads of exactly this length use the shared composition tables (0xffffffff: none)
};

// Base classes: A/a 0, T/t 1, C/c 2, G/g 3, N/n 4 (FaQCs.h:35-42), anything else 5.
__host__ __device__ __forceinline__ int base_code_slow(uint32_t c)
{
    c |= 0x20u;
    return c == 'a' ? 0 : c == 't' ? 1 : c == 'c' ? 2 : c == 'g' ? 3 : c == 'n' ? 4 : 5;
}
// LUT entry: bits 28..31 = class, bits 0..24 = 1 << (5 * class) for classes 0..4 (packed per-lane counters).
__host__ __device__ __forceinline__ uint32_t lut_entry(uint32_t c)
{
    const int code = base_code_slow(c);
    return ((uint32_t)code << 28) | (code < 5 ? (1u << (5 * code)) : 0u);
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) { return __reduce_add_sync(0xffffffffu, v); }
__device__ __forceinline__ int warp_sum_i(int v) { return __reduce_add_sync(0xffffffffu, v); }

// trim.cpp:553-576 with the reference's C types: float(int)/float(size_t) - float(char), clamped at 0.
__device__ __forceinline__ float average_quality(int total, uint32_t len, int offset)
{
    if (len == 0) return 0.0f;
    return fmaxf(0.0f, __fsub_rn(__fdiv_rn((float)total, (float)len), (float)offset));
}

// trim.cpp:860-874: bin = size_t(float(10000)/len * count), float arithmetic.
__device__ __forceinline__ float composition_norm(uint32_t len) { return len ? __fdiv_rn(10000.0f, (float)len) : 0.0f; }
__device__ __forceinline__ uint32_t composition_bin_n(float norm, uint32_t count) { return __float2uint_rz(__fmul_rn(norm, (float)count)); }

// Quality value of absolute position p with terminal-N masking applied
// (mask_quality_terminal_N, trim.cpp:1191-1216; quality_score, fastq.h:17-36).
struct QualAt {
    const signed char *q;
    uint32_t lead, trail;
    int in_off;
    __device__ __forceinline__ int operator()(uint32_t p) const
    {
        if (p < lead || p >= trail) return 0;
        return max(0, (int)q[p] - in_off);
    }
};

// hard_trim (trim.cpp:629-672).  Window-relative; returns new length, f5 = 5' cut.
__device__ __forceinline__ uint32_t hard_trim(const QualAt &qa, uint32_t lo, int len, int Q, bool protect_5, uint32_t &f5)
{
    int pos_3 = len - 1, final_pos_5 = 0, final_pos_3 = pos_3;
    while (pos_3 > 0) {
        if (Q < qa(lo + pos_3)) { final_pos_3 = pos_3; break; }
        --pos_3;
    }
    if (!protect_5) {
        int pos_5 = 0;
        while (pos_5 < pos_3) {
            if (Q < qa(lo + pos_5)) { final_pos_5 = pos_5; break; }
            ++pos_5;
        }
    }
    f5 = (uint32_t)final_pos_5;
    return (uint32_t)(final_pos_3 - final_pos_5 + 1);
}

// BWA_trim (trim.cpp:675-709).
__device__ __forceinline__ uint32_t bwa_trim(const QualAt &qa, uint32_t lo, int len, int Q, uint32_t &f5)
{
    int pos_3 = len - 1, final_pos_3 = pos_3, area = 0, max_area = 0;
    while (pos_3 > 0 && area >= 0) {
        area += Q - qa(lo + pos_3);
        if (area > max_area) { max_area = area; final_pos_3 = pos_3 - 1; }
        --pos_3;
    }
    f5 = 0;
    return (uint32_t)(final_pos_3 + 1);
}

// BWA_plus_trim (trim.cpp:714-793).
__device__ __forceinline__ uint32_t bwa_plus_trim(const QualAt &qa, uint32_t lo, int len, int Q, bool protect_5, uint32_t &f5)
{
    const int nan = min(2, len);
    int als = min(5, len), pos_3 = len - 1, final_pos_5 = 0, final_pos_3 = pos_3, area = 0, max_area = 0;
    while (als) {
        --als;
        if (pos_3 > nan && area >= 0) als = nan;
        area += Q - qa(lo + pos_3);
        if (area > max_area) { max_area = area; final_pos_3 = pos_3 - 1; }
        --pos_3;
    }
    if (!protect_5) {
        int pos_5 = 0;
        max_area = 0;
        area = 0;
        als = min(5, len);
        while (als) {
            --als;
            if (pos_5 < final_pos_3 - nan && area >= 0) als = nan;
            area += Q - qa(lo + pos_5);
            if (area > max_area) { max_area = area; final_pos_5 = pos_5 + 1; }
            ++pos_5;
        }
    }
    f5 = (uint32_t)final_pos_5;
    if (final_pos_3 <= final_pos_5) return 0;
    return (uint32_t)(final_pos_3 - final_pos_5 + 1);
}

// Shared-memory block of one CTA, addressed by word offsets from the dynamic shared base
// (offsets are functions of `rows` alone, so the layout costs two registers, not twelve pointers).
extern __shared__ uint32_t g_smem[];
constexpr uint32_t kTrashWords = 320;      // phase 1 handles reads of up to 320 bases; every position needs a dump slot
struct SmemHist {
    uint32_t rows, key;
    // Tables at fixed offsets first (their addresses are compile-time constants in the hot loop), then the histograms:
    //   [256] class LUT, [256] + [258] phase-1 LUTs (uint2; the quality table has a pad entry 256 for lanes past the end of a
    //   read), [320] trash row, [12] composition bin 0, [4][42] avg-Q hists, [32] filter counters,
    //   [42][rows] pre / removed quality, [5][rows] pre / removed base, [rows] g2n, [rows+1] x2 length,
    //   [2][7][key+1] composition by count
    static constexpr uint32_t kLut = 0, kLutBase = 256, kLutQual = kLutBase + 2 * 256, kTrash = kLutQual + 2 * 258, kZero = kTrash + kTrashWords,
                              kQh = kZero + 12, kFilt = kQh + 4 * kQualCols, kHist = kFilt + 32;
    __host__ __device__ static size_t words(uint32_t rows, uint32_t key)
    {
        return kHist + (size_t)rows * (2 * kQualCols + 2 * kBaseCols + 1) + 2 * ((size_t)rows + 1) + (key == 0xffffffffu ? 0 : 14 * ((size_t)key + 1));
    }
    __device__ __forceinline__ uint32_t *preq() const { return g_smem + kHist; }
    __device__ __forceinline__ uint32_t *remq() const { return preq() + kQualCols * rows; }
    __device__ __forceinline__ uint32_t *preb() const { return preq() + 2 * kQualCols * rows; }
    __device__ __forceinline__ uint32_t *remb() const { return preq() + (2 * kQualCols + kBaseCols) * rows; }
    __device__ __forceinline__ uint32_t *g2n() const { return preq() + (2 * kQualCols + 2 * kBaseCols) * rows; }
    __device__ __forceinline__ uint32_t *prelen() const { return preq() + (2 * kQualCols + 2 * kBaseCols + 1) * rows; }
    __device__ __forceinline__ uint32_t *postlen() const { return prelen() + rows + 1; }
    __device__ __forceinline__ uint32_t *compk() const { return postlen() + rows + 1; }
    __device__ __forceinline__ uint32_t *qh() const { return g_smem + kQh; }
    __device__ __forceinline__ uint32_t *filt() const { return g_smem + kFilt; }
    __device__ __forceinline__ uint32_t *lut() const { return g_smem + kLut; }
    // phase-1 tables: per byte value {byte offset (from the start of the block) of the histogram row to bump, payload}
    //   base:    row = pre_b[class] (trash row for non-ACGTN), payload = 1 << (8 * class) for A, T, C, G (packed per-lane
    //            counters, one byte each), 0 for N and everything else
    //   quality: row = pre_q[max(0, (signed char)ch - in_off)], payload = (int)(signed char)ch; a score above 41 goes to the
    //            trash row with kBadQual added to the payload (the read's quality sum then exposes it); entry 256 is the pad
    //            of lanes past the end of the read: trash row, payload 0
    __device__ __forceinline__ uint2 *lut_base() const { return reinterpret_cast<uint2 *>(g_smem + kLutBase); }
    __device__ __forceinline__ uint2 *lut_qual() const { return reinterpret_cast<uint2 *>(g_smem + kLutQual); }
    __device__ __forceinline__ uint32_t *trash() const { return g_smem + kTrash; }
    __device__ __forceinline__ uint32_t trash_bytes() const { return kTrash * 4u; }
    __device__ __forceinline__ uint32_t *zero() const { return g_smem + kZero; }
};

__device__ __forceinline__ void gadd(unsigned long long *p, unsigned long long v) { atomicAdd(p, v); }

struct KernelCtx {
    const TrimArgs &a;
    const DevOpts &o;
    const SmemHist &H;
    unsigned long long *S;
    uint32_t lane;
    uint32_t col;       // shared-memory byte address of this lane's position column (phase 1)
    uint32_t one;       // a run-time 1 (see red_shared_add)
};

// One increment in each of the six composition histograms (trim.cpp:860-874).  `cnt` holds the
// count of class `lane` in lanes 0..4.  which = 0 (pre) / 1 (post).
__device__ __forceinline__ void composition_update(const KernelCtx &kc, int which, uint32_t len, uint32_t cnt)
{
    const uint32_t lane = kc.lane;
    const float norm = composition_norm(len);
    const uint32_t bin = composition_bin_n(norm, cnt);
    const uint32_t iC = __shfl_sync(0xffffffffu, bin, 2), iG = __shfl_sync(0xffffffffu, bin, 3);
    const size_t comp = which ? kc.a.L.post_comp : kc.a.L.pre_comp;
    if (len == kc.H.key) {
        const uint32_t cC = __shfl_sync(0xffffffffu, cnt, 2), cG = __shfl_sync(0xffffffffu, cnt, 3);
        uint32_t *tab = kc.H.compk() + (size_t)which * 7 * (kc.H.key + 1);
        if (lane < 5) atomicAdd(&tab[lane * (kc.H.key + 1) + cnt], 1u);
        else if (lane == 5) {
            const uint32_t s = cC + cG;
            const uint32_t delta = composition_bin_n(norm, s) - (iC + iG);      // bin(G)+bin(C) vs bin(G+C): 0 or 1
            if (delta < 2) atomicAdd(&tab[(5 + delta) * (kc.H.key + 1) + s], 1u);
            else gadd(&kc.S[comp + 5 * (size_t)kCompBins + iC + iG], 1);
        }
    } else if (lane < 6) {
        const uint32_t b = lane < 5 ? bin : iC + iG;
        if (b == 0) atomicAdd(&kc.H.zero()[which * 6 + lane], 1u);
        else gadd(&kc.S[comp + (size_t)lane * kCompBins + b], 1);
    }
}

// Lane-parallel scalar statistics of one read: composition, length histogram, avg-Q histograms.
__device__ __forceinline__ void scalar_stats(const KernelCtx &kc, int which, uint32_t len, uint32_t cnt, int qbin)
{
    composition_update(kc, which, len, cnt);
    const uint32_t lane = kc.lane;
    qbin = min(max(qbin, 0), 41);
    if (lane == 6) {
        uint32_t *h = which ? kc.H.postlen() : kc.H.prelen();
        if (len <= kc.H.rows) atomicAdd(&h[len], 1u);
        else gadd(&kc.S[(which ? kc.a.L.post_len : kc.a.L.pre_len) + len], 1);
    } else if (lane == 7) atomicAdd(&kc.H.qh()[(2 * which) * kQualCols + qbin], 1u);
    else if (lane == 8) atomicAdd(&kc.H.qh()[(2 * which + 1) * kQualCols + qbin], len);
}

// Window after adapter clip, 5'/3' clip, length filter and quality trim (trim.cpp:270-360).
struct Window {
    uint32_t lo, wl, off5, flags;
    bool ret;
    int best_adapter;
};

__device__ __forceinline__ Window determine_window(const KernelCtx &kc, uint32_t mate, uint32_t r, uint32_t len, const QualAt &qa)
{
    const DevOpts &o = kc.o;
    const SmemHist &H = kc.H;
    const uint32_t lane = kc.lane;
    Window w{0, len, 0, 0, true, -1};
    if (o.filter_adapter && kc.a.adp[mate]) {
        const uint2 v = kc.a.adp[mate][r];
        w.best_adapter = kc.a.adp_best[mate][r];
        if (w.best_adapter >= 0) w.flags |= FQ_RR_ADAPTER;
        if (len != v.y) {
            w.lo = v.x;
            w.wl = v.y;
            w.off5 += (v.y == 0) ? len : v.x;
        }
    }
    if (o.trim_5 && !o.qc_only) {
        if (o.trim_5 > w.wl) w.wl = 0;                          // offset_5 += len after len = 0 (Q14)
        else { w.lo += o.trim_5; w.wl -= o.trim_5; w.off5 += o.trim_5; }
    }
    if (o.trim_3 && !o.qc_only) {
        if (o.trim_3 > w.wl) w.wl = 0;
        else w.wl -= o.trim_3;
    }
    if (w.wl < o.min_len || w.wl == 0) {                        // trim.cpp:317-323
        if (lane == 0) { atomicAdd(&H.filt()[FQ_READ_LENGTH], 1u); atomicAdd(&H.filt()[FQ_BASE_LENGTH], w.wl); }
        w.flags |= FQ_RR_F_LENGTH;
        w.ret = false;
    }
    if (!o.qc_only && w.ret) {                                  // trim.cpp:325-360
        const uint32_t init_len = w.wl;
        uint32_t f5 = 0;
        if (o.mode == FQ_MODE_HARD) w.wl = hard_trim(qa, w.lo, (int)w.wl, o.quality, o.protect_5 != 0, f5);
        else if (o.mode == FQ_MODE_BWA) w.wl = bwa_trim(qa, w.lo, (int)w.wl, o.quality, f5);
        else w.wl = bwa_plus_trim(qa, w.lo, (int)w.wl, o.quality, o.protect_5 != 0, f5);
        w.off5 += f5;
        w.lo += f5;
        if (init_len != w.wl) {
            if (lane == 0) { atomicAdd(&H.filt()[FQ_READ_QUAL_TRIM], 1u); atomicAdd(&H.filt()[FQ_BASE_QUAL_TRIM], init_len - w.wl); }
            w.flags |= FQ_RR_QUAL_TRIMMED;
        }
        if (w.wl < o.min_len || w.wl == 0) {
            if (lane == 0) { atomicAdd(&H.filt()[FQ_READ_LENGTH], 1u); atomicAdd(&H.filt()[FQ_BASE_LENGTH], w.wl); }
            w.flags |= FQ_RR_F_LENGTH;
            w.ret = false;
        }
    }
    return w;
}

// Dinucleotide counts of the window straight from global memory (rare path, trim.cpp:426-481).
__device__ __forceinline__ bool dinucleotide_low_complexity(const KernelCtx &kc, const uint8_t *sp, const QualAt &qa, uint32_t lo, uint32_t wl, float norm2)
{
    const DevOpts &o = kc.o;
    const uint32_t lane = kc.lane;
    uint32_t dc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) dc[k] = 0;
    int prev_carry = 4;
    for (uint32_t b = 0; b < wl; b += 32) {
        const uint32_t i = b + lane;
        const bool in = i < wl;
        const uint32_t p = lo + i;
        const uint32_t c = in ? sp[p] : 0;
        int cur = in ? (int)(kc.H.lut()[c] >> 28) : 4;
        if (cur > 3) cur = 4;
        if (in && o.replace_q > 0 && c == 'G' && qa(p) < (int)o.replace_q) cur = 4;
        int prev = __shfl_up_sync(0xffffffffu, cur, 1);
        if (lane == 0) prev = prev_carry;
        prev_carry = __shfl_sync(0xffffffffu, cur, 31);
        const int code = (in && cur != 4 && prev != 4 && cur != prev) ? ((prev << 2) | cur) : -1;
#pragma unroll
        for (int k = 0; k < 16; ++k) dc[k] += __popc(__ballot_sync(0xffffffffu, code == k));
    }
    bool lowc = false;
#pragma unroll
    for (int k = 0; k < 16; ++k) lowc |= __fmul_rn((float)dc[k], norm2) > o.lc;
    return lowc;
}

// Filters on the window (trim.cpp:363-513).  Counts are of the window after G->N replacement.
__device__ __forceinline__ void apply_filters(const KernelCtx &kc, Window &w, uint32_t r, uint32_t max_run, int sum_w,
                                              uint32_t wA, uint32_t wT, uint32_t wC, uint32_t wG, int max_qv, const uint8_t *sp,
                                              const QualAt &qa, float &ave_q)
{
    const DevOpts &o = kc.o;
    const SmemHist &H = kc.H;
    const uint32_t lane = kc.lane;
    if (max_run >= o.max_poly_n) {                              // trim.cpp:363-371
        if (lane == 0) { atomicAdd(&H.filt()[FQ_READ_NN], 1u); atomicAdd(&H.filt()[FQ_BASE_NN], w.wl); }
        w.flags |= FQ_RR_F_NN;
        if (!o.qc_only) w.ret = false;
    }
    ave_q = average_quality(sum_w, w.wl, o.in_off);
    if (w.ret && ave_q < o.avg_q) {                             // trim.cpp:374-382
        if (lane == 0) { atomicAdd(&H.filt()[FQ_READ_AVG_Q], 1u); atomicAdd(&H.filt()[FQ_BASE_AVG_Q], w.wl); }
        w.flags |= FQ_RR_F_AVGQ;
        w.ret = false;
    }
    if (w.ret) {                                                // low complexity, trim.cpp:405-513
        const float norm = (float)(1.0 / (double)w.wl);          // trim.cpp:483
        bool lowc = __fmul_rn((float)wA, norm) > o.lc || __fmul_rn((float)wT, norm) > o.lc ||
                    __fmul_rn((float)wG, norm) > o.lc || __fmul_rn((float)wC, norm) > o.lc;
        if (!lowc) {
            const float norm2 = norm * 2.0f;                    // trim.cpp:499
            // a dinucleotide count can not exceed the second largest base count
            const uint32_t second = max(max(min(wA, wT), min(wC, wG)), min(max(wA, wT), max(wC, wG)));
            if (__fmul_rn((float)second, norm2) > o.lc) lowc = dinucleotide_low_complexity(kc, sp, qa, w.lo, w.wl, norm2);
        }
        if (lowc) {
            if (lane == 0) { atomicAdd(&H.filt()[FQ_READ_LOW_COMPLEXITY], 1u); atomicAdd(&H.filt()[FQ_BASE_LOW_COMPLEXITY], w.wl); }
            w.flags |= FQ_RR_F_LOWCOMP;
            w.ret = false;
        }
    }
    if (w.ret && o.in_off != o.out_off && max_qv + o.out_off > 127 && lane == 0) {      // trim.cpp:516-525
        atomicOr(&kc.a.info->err, kErrReencode);
        atomicMin(&kc.a.info->err_record, r);
    }
}

// Longest run of set bits in a multi-word bitmap fed word by word (count_poly_n, trim.cpp:578-597).
struct RunTracker {
    uint32_t carry = 0, best = 0;
    __device__ __forceinline__ void feed(uint32_t m)
    {
        if (m) {
            uint32_t x = m;
            const int f = __ffs(~x);
            const uint32_t head = f ? (uint32_t)(f - 1) : 32u;
            best = max(best, carry + head);
            uint32_t k = 0;
            while (x) { x &= x >> 1; ++k; }
            best = max(best, k);
            carry = (m == 0xffffffffu) ? carry + 32 : (uint32_t)__clz(~m);
        } else carry = 0;
    }
};

__device__ __forceinline__ uint32_t unpack5(uint32_t packed, int field) { return (packed >> (5 * field)) & 31u; }

// ---------------------------------------------------------------------------------------------
// Batched fast path.  A warp takes 32 reads at a time.  Per-BASE work (loads, histogram
// updates, class counts) is done cooperatively, one read after the other, lanes striped over
// positions.  Per-READ scalar work (window, quality trim, filters, bins, verdict) is done by one
// lane per read, so its cost is amortised over 32 reads.  Rare per-base follow-ups (bases that
// were trimmed away, discarded reads, dinucleotide counts, exact N runs) are cooperative again,
// driven by ballots.
// ---------------------------------------------------------------------------------------------

// What phase 1 leaves in the lane that owns the read.
struct LaneRead {
    Rec rc;
    int sum_q;              // sum of (masked) raw quality chars of the whole read
    uint32_t cnt_ac;        // 16-bit fields A,C  (phase 1), then 10-bit fields A,T,C
    uint32_t cnt_tg;        // 16-bit fields T,G  (phase 1), then 10-bit fields G,N
    uint32_t cnt_n;         // N / n
    uint32_t lead, trail;   // terminal-N mask bounds
    uint32_t run_whole;     // longest 'N' run of the whole read (0 if fewer than -n 'N's)
    uint32_t lowg;          // 'G' below --replace_to_N_q in the whole read
    bool done;              // already fully processed (generic path) or out of range
};

// `one` must be a run-time 1: with an immediate the assembler picks the warp-aggregating form, which needs a
// convergence region (three more instructions) around every single increment.
__device__ __forceinline__ void red_shared_add(uint32_t addr, uint32_t v)
{
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

__device__ __forceinline__ uint32_t f10(uint32_t packed, int field) { return (packed >> (10 * field)) & 1023u; }

constexpr uint32_t kQualPad = 256;          // index of the pad entry of the quality table
constexpr int kBadQual = 1 << 20;           // added to the quality sum by every score above 41 (fastq.h:31-33)

// Loads of one read for phase 1: raw bytes per chunk of 32 bases, lanes striped over positions.  A lane past the end of the
// read holds a non-base and the quality table's pad index, whose table entries point at the trash row and add nothing to
// the per-read sums -- so the histogram increment is the same run-time 1 for every lane.  Reads that fill all but the last
// chunk (the fixed-length case) load those chunks without a predicate.
template <int K>
__device__ __forceinline__ void phase1_load(const KernelCtx &kc, const uint8_t *raw, uint32_t seq, uint32_t qual, uint32_t len,
                                            uint32_t (&c)[K], uint32_t (&q)[K])
{
    const uint32_t lane = kc.lane;
    const uint8_t *const spl = raw + (seq + lane), *const qpl = raw + (qual + lane);
    if (K <= 5 && len > (uint32_t)(32 * (K - 1))) {
#pragma unroll
        for (int k = 0; k < K - 1; ++k) {
            c[k] = __ldg(spl + k * 32);
            q[k] = __ldg(qpl + k * 32);
        }
        const bool in = (uint32_t)(32 * (K - 1)) + lane < len;
        c[K - 1] = in ? (uint32_t)__ldg(in ? spl + 32 * (K - 1) : spl) : 0u;
        q[K - 1] = in ? (uint32_t)__ldg(in ? qpl + 32 * (K - 1) : qpl) : kQualPad;
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const bool in = (uint32_t)(k * 32) + lane < len;
            c[k] = 0;
            q[k] = kQualPad;
            if (in) {
                c[k] = __ldg(spl + k * 32);
                q[k] = __ldg(qpl + k * 32);
            }
        }
    }
}

// Bounds of the terminal 'N' runs of a read given its per-chunk 'N' ballots (mask_quality_terminal_N, trim.cpp:1191-1216):
// returns {lead, trail}.  Rare (first or last base is 'N'), kept out of line and out of the hot loop's registers.
template <int K>
struct NBallots {
    uint32_t w[K];
};
template <int K>
__device__ __noinline__ uint2 terminal_n_bounds(const NBallots<K> nm, uint32_t len)
{
    uint32_t lead = 0, trail = len;
    bool open = true;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        if (open && (uint32_t)(k * 32) < len) {
            const uint32_t nb = min(32u, len - k * 32);
            const uint32_t valid = nb == 32 ? 0xffffffffu : ((1u << nb) - 1u);
            const uint32_t inv = ~nm.w[k] & valid;
            if (inv) { lead += __ffs(inv) - 1; open = false; }
            else lead += nb;
        }
    }
    open = true;
#pragma unroll
    for (int k = K - 1; k >= 0; --k) {
        if (open && (uint32_t)(k * 32) < len) {
            const uint32_t nb = min(32u, len - k * 32);
            const uint32_t valid = nb == 32 ? 0xffffffffu : ((1u << nb) - 1u);
            const uint32_t inv = ~nm.w[k] & valid;
            if (inv) { trail = k * 32 + (31 - __clz(inv)) + 1; open = false; }
            else trail = k * 32;
        }
    }
    if (lead >= len) trail = 0;
    return make_uint2(lead, trail);
}

// Phase 1 for one read: PRE matrices + per-read summaries.  Branch-free in the common case: every lane looks both bytes up
// (two LDS.64 from tables at constant addresses), bumps two histogram cells (two RED.shared) and accumulates the payloads;
// the A/T/C/G counts of the read are ONE warp reduction of byte-packed counters.  Only a read that holds something else
// than A, C, G, T takes the ballot path ('N' counts, terminal-N mask, longest run); a quality above 41 shows up in the sum.
struct ReadSummary {
    int sum_q;
    uint32_t ac, tg;        // 16-bit fields: A | C << 16, T | G << 16
    uint32_t n;             // N / n
    uint32_t lead, trail, run;
    uint32_t lowg;          // 'G' below --replace_to_N_q in the whole read (0 when the option is off)
};

template <int K>
__device__ __forceinline__ void phase1_body(const KernelCtx &kc, uint32_t (&c)[K], uint32_t (&q)[K], uint32_t len, ReadSummary &out)
{
    const DevOpts &o = kc.o;
    const uint32_t lane = kc.lane;
    const uint2 *const lb = reinterpret_cast<const uint2 *>(g_smem + SmemHist::kLutBase);
    const uint2 *const lq = reinterpret_cast<const uint2 *>(g_smem + SmemHist::kLutQual);
    const uint32_t col = kc.col;                    // shared address of this lane's position column; chunk k adds 128 bytes
#ifdef FQ_EXP_ONE_RUNTIME
    const uint32_t one = kc.one;                    // a run-time 1 keeps the plain ATOMS.ADD
#else
    const uint32_t one = 1u;                        // ptxas picks ATOMS.POPC.INC; every lane takes part, so no convergence region
#endif
    constexpr int KH = K <= 5 ? K : 5;              // byte-packed counters hold at most 5 chunks per lane
    uint32_t packed = 0, packed_hi = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        if (K <= 5 || (uint32_t)(k * 32) < len) {
            const uint2 eb = lb[c[k]];
            if (k < KH) packed += eb.y; else packed_hi += eb.y;
#ifndef FQ_EXP_NOATOM
            red_shared_add(col + eb.x + k * 128, one);
#endif
        }
    }
    uint32_t atcg = warp_sum(packed);               // A | T << 8 | C << 16 | G << 24: at most 160 each
    out.ac = atcg & 0x00ff00ffu;
    out.tg = (atcg >> 8) & 0x00ff00ffu;
    uint32_t acgt = __dp4a(atcg, 0x01010101u, 0u);
    if (K > 5) {
        atcg = warp_sum(packed_hi);
        out.ac += atcg & 0x00ff00ffu;
        out.tg += (atcg >> 8) & 0x00ff00ffu;
        acgt = __dp4a(atcg, 0x01010101u, acgt);
    }
    out.n = 0;
    out.lead = 0;
    out.trail = len;
    out.run = 0;
    out.lowg = 0;
    if (o.replace_q > 0) {                          // candidates of the G -> N replacement (trim.cpp:389-403) in the whole read
        const int below = o.in_off + (int)o.replace_q;          // quality_score(q) < replace_q  <=>  q < in_off + replace_q
        uint32_t cnt = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) cnt += __popc(__ballot_sync(0xffffffffu, c[k] == 'G' && (int)(signed char)q[k] < below));
        out.lowg = cnt;
    }
    if (acgt != len) {                              // something else than A, C, G, T (any case) in the read (warp-uniform)
        uint32_t n_any = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) n_any += __popc(__ballot_sync(0xffffffffu, (c[k] | 0x20u) == 'n'));
        out.n = n_any;
        if (n_any) {
            NBallots<K> nm;
            uint32_t n_count = 0;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                nm.w[k] = __ballot_sync(0xffffffffu, c[k] == 'N');
                n_count += __popc(nm.w[k]);
            }
            uint32_t nm_last = nm.w[0];             // ballot of the chunk that holds the last base (no dynamic indexing)
#pragma unroll
            for (int k = 1; k < K; ++k)
                if (((len - 1) >> 5) == (uint32_t)k) nm_last = nm.w[k];
            if (len && ((nm.w[0] & 1u) | ((nm_last >> ((len - 1) & 31)) & 1u))) {   // terminal 'N' runs (trim.cpp:1191-1216)
                const uint2 lt = terminal_n_bounds<K>(nm, len);
                out.lead = lt.x;
                out.trail = lt.y;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const uint32_t p = k * 32 + lane;
                    if (p < len && (p < lt.x || p >= lt.y)) q[k] = (uint32_t)o.in_off & 0xffu;
                }
            }
            if (n_count >= o.max_poly_n) {          // candidate for the N filter: longest run of the whole read
                RunTracker rt;
#pragma unroll
                for (int k = 0; k < K; ++k)
                    if ((uint32_t)(k * 32) < len) rt.feed(nm.w[k]);
                out.run = rt.best;
            }
        }
    }
    int sum_q = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        if (K <= 5 || (uint32_t)(k * 32) < len) {
            const uint2 eq = lq[q[k]];
            sum_q += (int)eq.y;
#ifndef FQ_EXP_NOATOM
            red_shared_add(col + eq.x + k * 128, one);
#endif
        }
    }
    out.sum_q = warp_sum_i(sum_q);
}

// Cooperative pass over one read for the "removed" histograms.
//   mode 0: positions OUTSIDE the window -> rem hists; returns their class counts / quality sum and, with
//           --replace_to_N_q, the G->N candidates among them (phase 1 counted those of the whole read); chunks that
//           lie inside the window are skipped unless the largest quality inside is asked for (max_qv >= 0).
//   mode 1: positions INSIDE the window -> rem hists (read turned out invalid).
//   mode 2: G->N positions inside the window of a valid read: leave column G, enter column N.
__device__ __forceinline__ void removed_pass(const KernelCtx &kc, int mode, const uint8_t *sp, const signed char *qp, uint32_t len,
                                             uint32_t lo, uint32_t wl, uint32_t lead, uint32_t trail, uint32_t &r_atc, uint32_t &r_gn,
                                             int &r_sum, uint32_t &n_lowg, int &max_qv)
{
    const DevOpts &o = kc.o;
    const SmemHist &H = kc.H;
    const StatsLayout &L = kc.a.L;
    const uint32_t lane = kc.lane, R = H.rows;
    uint32_t pk = 0, lowg_cnt = 0;
    int sum = 0, mq = 0;
    for (uint32_t b = 0; b < len; b += 32) {
        const uint32_t p = b + lane;
        // skip chunks that cannot contain work
        const bool chunk_inside = b >= lo && b + 32 <= lo + wl;
        if (mode == 0 && chunk_inside && max_qv < 0) continue;
        if (mode != 0 && (b + 32 <= lo || b >= lo + wl)) continue;
        if (p >= len) continue;
        const bool inside = p >= lo && p < lo + wl;
        const uint32_t c = sp[p];
        int qc = (int)qp[p];
        if (p < lead || p >= trail) qc = o.in_off;
        const int qv = max(0, qc - o.in_off);
        const uint32_t code = H.lut()[c] >> 28;
        const bool lowg_any = o.replace_q > 0 && c == 'G' && qv < (int)o.replace_q;
        const bool lowg = inside && lowg_any;
        if (mode == 0) {
            if (!inside) {
                sum += qc;
                if (code < 5) pk += 1u << (5 * code);
                lowg_cnt += lowg_any;               // candidates OUTSIDE the window: the caller knows the read's total
            } else mq = max(mq, qv);
        }
        const bool to_rem = (mode == 0 && !inside) || (mode == 1 && inside);
        if (to_rem) {
            if (qv <= FQ_MAX_QUALITY_SCORE) {
                if (p < R) atomicAdd(&H.remq()[qv * R + p], 1u);
                else gadd(&kc.S[L.rem_q + (size_t)qv * L.rows + p], 1);
            }
            if (code < 5) {
                if (p < R) atomicAdd(&H.remb()[code * R + p], 1u);
                else gadd(&kc.S[L.rem_b + (size_t)code * L.rows + p], 1);
            }
        } else if (mode == 2 && lowg) {
            if (p < R) { atomicAdd(&H.remb()[3 * R + p], 1u); atomicAdd(&H.g2n()[p], 1u); }
            else { gadd(&kc.S[L.rem_b + (size_t)3 * L.rows + p], 1); gadd(&kc.S[L.g2n + p], 1); }
        }
    }
    if (mode == 0) {
        r_sum = warp_sum_i(sum);
        r_atc = warp_sum(unpack5(pk, 0) | (unpack5(pk, 1) << 10) | (unpack5(pk, 2) << 20));
        r_gn = warp_sum(unpack5(pk, 3) | (unpack5(pk, 4) << 10));
        n_lowg = warp_sum(lowg_cnt);
        max_qv = __reduce_max_sync(0xffffffffu, mq);
    }
}

// Longest 'N' run inside [lo, lo+wl), cooperative, from global memory (rare).
__device__ __forceinline__ uint32_t window_n_run(const KernelCtx &kc, const uint8_t *sp, uint32_t lo, uint32_t wl)
{
    RunTracker rt;
    for (uint32_t b = 0; b < wl; b += 32) {
        const uint32_t i = b + kc.lane;
        rt.feed(__ballot_sync(0xffffffffu, i < wl && sp[lo + i] == 'N'));
    }
    return rt.best;
}

// One lane: the six composition increments of one read (trim.cpp:860-874).
__device__ __forceinline__ void lane_composition(const KernelCtx &kc, int which, uint32_t len, uint32_t atc, uint32_t gn)
{
    const SmemHist &H = kc.H;
    const float norm = composition_norm(len);
    const uint32_t cnt[5] = {f10(atc, 0), f10(atc, 1), f10(atc, 2), f10(gn, 0), f10(gn, 1)};
    const size_t comp = which ? kc.a.L.post_comp : kc.a.L.pre_comp;
    const uint32_t iC = composition_bin_n(norm, cnt[2]), iG = composition_bin_n(norm, cnt[3]);
    if (len == H.key) {
        uint32_t *tab = H.compk() + (size_t)which * 7 * (H.key + 1);
#pragma unroll
        for (int h = 0; h < 5; ++h) atomicAdd(&tab[h * (H.key + 1) + cnt[h]], 1u);
        const uint32_t s = cnt[2] + cnt[3];
        const uint32_t delta = composition_bin_n(norm, s) - (iC + iG);          // bin(G)+bin(C) vs bin(G+C): 0 or 1
        if (delta < 2) atomicAdd(&tab[(5 + delta) * (H.key + 1) + s], 1u);
        else gadd(&kc.S[comp + 5 * (size_t)kCompBins + iC + iG], 1);
    } else {
#pragma unroll
        for (int h = 0; h < 6; ++h) {
            const uint32_t b = h == 2 ? iC : h == 3 ? iG : h == 5 ? iC + iG : composition_bin_n(norm, cnt[h == 5 ? 0 : h]);
            if (b == 0) atomicAdd(&H.zero()[which * 6 + h], 1u);
            else gadd(&kc.S[comp + (size_t)h * kCompBins + b], 1);
        }
    }
}

__device__ __forceinline__ void lane_scalar_stats(const KernelCtx &kc, int which, uint32_t len, uint32_t atc, uint32_t gn, int qbin)
{
    const SmemHist &H = kc.H;
    lane_composition(kc, which, len, atc, gn);
    qbin = min(max(qbin, 0), 41);
    uint32_t *h = which ? H.postlen() : H.prelen();
    if (len <= H.rows) atomicAdd(&h[len], 1u);
    else gadd(&kc.S[(which ? kc.a.L.post_len : kc.a.L.pre_len) + len], 1);
    atomicAdd(&H.qh()[(2 * which) * kQualCols + qbin], 1u);
    atomicAdd(&H.qh()[(2 * which + 1) * kQualCols + qbin], len);
}

// Per-lane accumulators (reduced over the warp at the end of the kernel).
struct LaneAcc {
    uint32_t reads = 0, trimmed = 0;
    unsigned long long len = 0, trimmed_len = 0;
    uint32_t max_pre_rows = 0, max_post_rows = 0, max_post_len1 = 0;
};

// Window determination by ONE lane for its own read (same logic as determine_window, lane-private counters).
__device__ __forceinline__ Window lane_window(const KernelCtx &kc, uint32_t mate, uint32_t r, uint32_t len, const QualAt &qa)
{
    const DevOpts &o = kc.o;
    const SmemHist &H = kc.H;
    Window w{0, len, 0, 0, true, -1};
    if (o.filter_adapter && kc.a.adp[mate]) {
        const uint2 v = kc.a.adp[mate][r];
        w.best_adapter = kc.a.adp_best[mate][r];
        if (w.best_adapter >= 0) w.flags |= FQ_RR_ADAPTER;
        if (len != v.y) {
            w.lo = v.x;
            w.wl = v.y;
            w.off5 += (v.y == 0) ? len : v.x;
        }
    }
    if (o.trim_5 && !o.qc_only) {
        if (o.trim_5 > w.wl) w.wl = 0;
        else { w.lo += o.trim_5; w.wl -= o.trim_5; w.off5 += o.trim_5; }
    }
    if (o.trim_3 && !o.qc_only) {
        if (o.trim_3 > w.wl) w.wl = 0;
        else w.wl -= o.trim_3;
    }
    if (w.wl < o.min_len || w.wl == 0) {
        atomicAdd(&H.filt()[FQ_READ_LENGTH], 1u);
        atomicAdd(&H.filt()[FQ_BASE_LENGTH], w.wl);
        w.flags |= FQ_RR_F_LENGTH;
        w.ret = false;
    }
    if (!o.qc_only && w.ret) {
        const uint32_t init_len = w.wl;
        uint32_t f5 = 0;
        if (o.mode == FQ_MODE_HARD) w.wl = hard_trim(qa, w.lo, (int)w.wl, o.quality, o.protect_5 != 0, f5);
        else if (o.mode == FQ_MODE_BWA) w.wl = bwa_trim(qa, w.lo, (int)w.wl, o.quality, f5);
        else w.wl = bwa_plus_trim(qa, w.lo, (int)w.wl, o.quality, o.protect_5 != 0, f5);
        w.off5 += f5;
        w.lo += f5;
        if (init_len != w.wl) {
            atomicAdd(&H.filt()[FQ_READ_QUAL_TRIM], 1u);
            atomicAdd(&H.filt()[FQ_BASE_QUAL_TRIM], init_len - w.wl);
            w.flags |= FQ_RR_QUAL_TRIMMED;
        }
        if (w.wl < o.min_len || w.wl == 0) {
            atomicAdd(&H.filt()[FQ_READ_LENGTH], 1u);
            atomicAdd(&H.filt()[FQ_BASE_LENGTH], w.wl);
            w.flags |= FQ_RR_F_LENGTH;
            w.ret = false;
        }
    }
    return w;
}

// ---------------------------------------------------------------------------------------------
// Generic path: any length, chunk loops over global memory (L1 resident after the first pass).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void process_generic(const KernelCtx &kc, uint32_t mate, uint32_t r, const Rec &rc, uint32_t &v_off5, uint32_t &v_lenflags)
{
    const DevOpts &o = kc.o;
    const SmemHist &H = kc.H;
    const StatsLayout &L = kc.a.L;
    unsigned long long *const S = kc.S;
    const uint32_t lane = kc.lane, R = H.rows;
    const uint8_t *sp = kc.a.raw[mate] + rc.seq;
    const signed char *qp = reinterpret_cast<const signed char *>(kc.a.raw[mate] + rc.qual);
    const uint32_t len = rc.len;

    uint32_t lead = 0, trail = len;
    if (len) {
        if (sp[0] == 'N') {
            lead = len;
            for (uint32_t b = 0; b < len; b += 32) {
                const uint32_t p = b + lane;
                const uint32_t m = __ballot_sync(0xffffffffu, p < len && sp[p] != 'N');
                if (m) { lead = b + __ffs(m) - 1; break; }
            }
        }
        if (sp[len - 1] == 'N') {
            trail = 0;
            for (int b = (int)((len - 1) & ~31u); b >= 0; b -= 32) {
                const uint32_t p = b + lane;
                const uint32_t m = __ballot_sync(0xffffffffu, p < len && sp[p] != 'N');
                if (m) { trail = b + (31 - __clz(m)) + 1; break; }
            }
        }
    }
    const QualAt qa{qp, lead, trail, o.in_off};

    int sum_q = 0;
    uint32_t nA = 0, nT = 0, nC = 0, nG = 0, nN = 0;
    bool bad_q = false;
    for (uint32_t b = 0; b < len; b += 32) {
        const uint32_t p = b + lane;
        const bool in = p < len;
        const uint32_t c = in ? sp[p] : 0;
        int qc = in ? (int)qp[p] : o.in_off;
        if (p < lead || p >= trail) qc = o.in_off;
        if (in) sum_q += qc;
        const int qv = max(0, qc - o.in_off);
        bad_q |= in && (qv > FQ_MAX_QUALITY_SCORE);
        const int bc = in ? (int)(H.lut()[c] >> 28) : 5;
        if (in && qv <= FQ_MAX_QUALITY_SCORE) {
            if (p < R) atomicAdd(&H.preq()[qv * R + p], 1u);
            else gadd(&S[L.pre_q + (size_t)qv * L.rows + p], 1);
        }
        if (bc < 5) {
            if (p < R) atomicAdd(&H.preb()[bc * R + p], 1u);
            else gadd(&S[L.pre_b + (size_t)bc * L.rows + p], 1);
        }
        nA += __popc(__ballot_sync(0xffffffffu, bc == 0));
        nT += __popc(__ballot_sync(0xffffffffu, bc == 1));
        nC += __popc(__ballot_sync(0xffffffffu, bc == 2));
        nG += __popc(__ballot_sync(0xffffffffu, bc == 3));
        nN += __popc(__ballot_sync(0xffffffffu, bc == 4));
    }
    if (__any_sync(0xffffffffu, bad_q) && lane == 0) { atomicOr(&kc.a.info->err, kErrQualGt41); atomicMin(&kc.a.info->err_record, r); }
    sum_q = warp_sum_i(sum_q);
    {
        const uint32_t mine = lane == 0 ? nA : lane == 1 ? nT : lane == 2 ? nC : lane == 3 ? nG : nN;
        scalar_stats(kc, 0, len, mine, (int)average_quality(sum_q, len, o.in_off));
    }
    if (lane == 0) {            // long reads are rare: counters go straight to the global block
        gadd(&S[L.filter + FQ_TOTAL_COUNT], 1);
        gadd(&S[L.filter + FQ_TOTAL_NUMBER], 1);
        gadd(&S[L.filter + FQ_TOTAL_LENGTH], len);
        atomicMax(&kc.a.rows->pre_rows, len);
        atomicMax(&kc.a.rows->pre_len_size, len + 1);
    }

    Window w = determine_window(kc, mate, r, len, qa);

    float ave_q = 0.0f;
    uint32_t wA = 0, wT = 0, wC = 0, wG = 0, wN = 0, n_lowg = 0;
    if (w.ret) {
        int sum_w = 0, max_qv = 0;
        RunTracker rt;
        for (uint32_t b = 0; b < w.wl; b += 32) {
            const uint32_t i = b + lane;
            const bool in = i < w.wl;
            const uint32_t p = w.lo + i;
            const uint32_t c = in ? sp[p] : 0;
            int qc = in ? (int)qp[p] : o.in_off;
            if (p < lead || p >= trail) qc = o.in_off;
            if (in) sum_w += qc;
            const int qv = max(0, qc - o.in_off);
            max_qv = max(max_qv, in ? qv : 0);
            const int bc = in ? (int)(H.lut()[c] >> 28) : 5;
            const bool lowg = in && o.replace_q > 0 && c == 'G' && qv < (int)o.replace_q;
            rt.feed(__ballot_sync(0xffffffffu, c == 'N'));
            n_lowg += __popc(__ballot_sync(0xffffffffu, lowg));
            wA += __popc(__ballot_sync(0xffffffffu, bc == 0));
            wT += __popc(__ballot_sync(0xffffffffu, bc == 1));
            wC += __popc(__ballot_sync(0xffffffffu, bc == 2));
            wG += __popc(__ballot_sync(0xffffffffu, bc == 3 && !lowg));
            wN += __popc(__ballot_sync(0xffffffffu, bc == 4 || lowg));
        }
        sum_w = warp_sum_i(sum_w);
        max_qv = __reduce_max_sync(0xffffffffu, max_qv);
        apply_filters(kc, w, r, rt.best, sum_w, wA, wT, wC, wG, max_qv, sp, qa, ave_q);
    }
    if (w.ret) {
        w.flags |= FQ_RR_VALID;
        if (lane == 0) {
            gadd(&S[L.filter + FQ_TOTAL_TRIMMED_NUMBER], 1);
            gadd(&S[L.filter + FQ_TOTAL_TRIMMED_LENGTH], w.wl);
            atomicMax(&kc.a.rows->post_rows, w.off5 + w.wl);
            atomicMax(&kc.a.rows->post_len_size, w.wl + 1);
        }
        const uint32_t mine = lane == 0 ? wA : lane == 1 ? wT : lane == 2 ? wC : lane == 3 ? wG : wN;
        scalar_stats(kc, 1, w.wl, mine, (int)ave_q);
    }
    const bool whole_removed = !w.ret;
    if (whole_removed || w.lo > 0 || w.lo + w.wl < len || n_lowg) {
        for (uint32_t b = 0; b < len; b += 32) {
            const uint32_t p = b + lane;
            if (p >= len) continue;
            const bool inside = !whole_removed && p >= w.lo && p < w.lo + w.wl;
            const uint32_t c = sp[p];
            const int qv = qa(p);
            if (!inside) {
                const int bc = (int)(H.lut()[c] >> 28);
                if (qv <= FQ_MAX_QUALITY_SCORE) {
                    if (p < R) atomicAdd(&H.remq()[qv * R + p], 1u);
                    else gadd(&S[L.rem_q + (size_t)qv * L.rows + p], 1);
                }
                if (bc < 5) {
                    if (p < R) atomicAdd(&H.remb()[bc * R + p], 1u);
                    else gadd(&S[L.rem_b + (size_t)bc * L.rows + p], 1);
                }
            } else if (n_lowg && o.replace_q > 0 && c == 'G' && qv < (int)o.replace_q) {
                if (p < R) { atomicAdd(&H.remb()[3 * R + p], 1u); atomicAdd(&H.g2n()[p], 1u); }
                else { gadd(&S[L.rem_b + (size_t)3 * L.rows + p], 1); gadd(&S[L.g2n + p], 1); }
            }
        }
    }
    v_off5 = w.off5;
    v_lenflags = pack_len_flags(w.ret ? w.wl : 0, w.flags | ((lead > 0 || trail < len) ? kFlagMasked : 0u));
    if (lane == 0) {
        if (kc.a.dbg[mate]) {
            fq_read_result d;
            d.offset_5 = w.off5;
            d.length = w.ret ? w.wl : 0;
            d.flags = (uint16_t)w.flags;
            d.adapter = (int16_t)w.best_adapter;
            d.avg_q = ave_q;
            kc.a.dbg[mate][r] = d;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// trim_group: trim_read (trim.cpp:225-551) for the 32 reads r0 .. r0+31 of one mate, by one warp.
// Returns, in the lane that owns read r0 + lane, the verdict {offset_5, length | flags << 24}
// (length 0 = invalid).  KSEL / PLAIN: see k_trim below.
// ---------------------------------------------------------------------------------------------
#ifndef FQ_TRIM_THREADS
#define FQ_TRIM_THREADS 1024
#endif
// width-specialised instances request the bytes of read j+1 before they process read j (C2: 1.27 -> 1.15 ms per 4 M reads)
#ifndef FQ_TRIM_PIPE
#define FQ_TRIM_PIPE 1
#endif
struct GroupErr {
    uint32_t bits = 0, rec = 0xffffffffu;
};

template <int KSEL, bool PLAIN>
__device__ __forceinline__ void trim_group(const KernelCtx &kc, uint32_t mate, uint32_t r0, const Rec &my_rc, LaneAcc &acc, GroupErr &ge,
                                           uint32_t &v_off5, uint32_t &v_lenflags)
{
    const TrimArgs &a = kc.a;
    const DevOpts &o = kc.o;
    const SmemHist &H = kc.H;
    const uint32_t lane = kc.lane, R = H.rows;
    const uint8_t *const raw = mate ? a.raw[1] : a.raw[0];
    const bool need_max = o.in_off != o.out_off && o.out_off + FQ_MAX_QUALITY_SCORE > 127;   // re-encode can overflow

    const uint32_t r = r0 + lane;
    const uint32_t n_here = r0 < a.n_rec ? min(32u, a.n_rec - r0) : 0u;
    LaneRead me;
    me.done = lane >= n_here;
    me.rc = my_rc;
    me.sum_q = 0; me.cnt_ac = 0; me.cnt_tg = 0; me.cnt_n = 0; me.lead = 0; me.trail = me.rc.len; me.run_whole = 0; me.lowg = 0;
    uint32_t g_off5 = 0, g_lenflags = 0;          // verdict of a read that took the generic path

    // ---- phase 1: cooperative per-base pass, read by read
    int max_sum = 0;
#if FQ_TRIM_PIPE
    // width-specialised instances request the bytes of read j+1 before they process read j
    constexpr int KP = KSEL == 4 ? 4 : 5;
    uint32_t nc[KP], nq[KP], nlen = 0;
    if (KSEL != 0 && n_here) {
        nlen = __shfl_sync(0xffffffffu, me.rc.len, 0);
        phase1_load<KP>(kc, raw, __shfl_sync(0xffffffffu, me.rc.seq, 0), __shfl_sync(0xffffffffu, me.rc.qual, 0), nlen, nc, nq);
    }
#endif
    for (uint32_t j = 0; j < n_here; ++j) {
#if FQ_TRIM_PIPE
        if (KSEL != 0) {
            uint32_t c[KP], q[KP];
#pragma unroll
            for (int k = 0; k < KP; ++k) { c[k] = nc[k]; q[k] = nq[k]; }
            const uint32_t len = nlen;
            if (j + 1 < n_here) {
                nlen = __shfl_sync(0xffffffffu, me.rc.len, j + 1);
                phase1_load<KP>(kc, raw, __shfl_sync(0xffffffffu, me.rc.seq, j + 1), __shfl_sync(0xffffffffu, me.rc.qual, j + 1), nlen, nc, nq);
            }
            ReadSummary rs;
            phase1_body<KP>(kc, c, q, len, rs);
            max_sum = max(max_sum, rs.sum_q);
            if (lane == j) { me.sum_q = rs.sum_q; me.cnt_ac = rs.ac; me.cnt_tg = rs.tg; }
            if (o.replace_q > 0 && lane == j) me.lowg = rs.lowg;
            if (rs.n && lane == j) { me.cnt_n = rs.n; me.lead = rs.lead; me.trail = rs.trail; me.run_whole = rs.run; }
            continue;
        }
#endif
        const uint32_t len = __shfl_sync(0xffffffffu, me.rc.len, j);
        const uint32_t seq = __shfl_sync(0xffffffffu, me.rc.seq, j);
        const uint32_t qual = __shfl_sync(0xffffffffu, me.rc.qual, j);
        ReadSummary rs;
        // KSEL 4 / 5: the host launches these instances only for batches whose longest read fits the width, so
        // the generic path (and its code) is not part of them
        if (KSEL == 4) {
            uint32_t c[4], q[4];
            phase1_load<4>(kc, raw, seq, qual, len, c, q);
            phase1_body<4>(kc, c, q, len, rs);
        } else if (KSEL == 5 || (len <= 160 && len <= R)) {
            uint32_t c[5], q[5];
            phase1_load<5>(kc, raw, seq, qual, len, c, q);
            phase1_body<5>(kc, c, q, len, rs);
        } else if (len <= 320 && len <= R) {
            uint32_t c[10], q[10];
            phase1_load<10>(kc, raw, seq, qual, len, c, q);
            phase1_body<10>(kc, c, q, len, rs);
        } else {
            const Rec rcj{__shfl_sync(0xffffffffu, me.rc.hdr, j), seq, qual, len};
            uint32_t go = 0, gl = 0;
            process_generic(kc, mate, r0 + j, rcj, go, gl);
            if (lane == j) { g_off5 = go; g_lenflags = gl; me.done = true; }
            continue;
        }
        max_sum = max(max_sum, rs.sum_q);
        if (lane == j) { me.sum_q = rs.sum_q; me.cnt_ac = rs.ac; me.cnt_tg = rs.tg; }
            if (o.replace_q > 0 && lane == j) me.lowg = rs.lowg;
        if (rs.n && lane == j) { me.cnt_n = rs.n; me.lead = rs.lead; me.trail = rs.trail; me.run_whole = rs.run; }
    }
    const bool bad_sum = max_sum >= kBadQual / 2;
    {   // counts in the layout the scalar phases use: 10-bit fields A,T,C and G,N
        const uint32_t A = me.cnt_ac & 0xffffu, C = me.cnt_ac >> 16, T = me.cnt_tg & 0xffffu, G = me.cnt_tg >> 16;
        me.cnt_ac = A | (T << 10) | (C << 20);
        me.cnt_tg = G | (me.cnt_n << 10);
    }

    if (bad_sum) {
        // a quality score above 41 (fastq.h:31-33): find the first offending read of the group
        for (uint32_t j = 0; j < n_here; ++j) {
            const uint32_t lenj = __shfl_sync(0xffffffffu, me.rc.len, j), qual = __shfl_sync(0xffffffffu, me.rc.qual, j);
            const uint32_t lead = __shfl_sync(0xffffffffu, me.lead, j), trail = __shfl_sync(0xffffffffu, me.trail, j);
            const bool skip = __shfl_sync(0xffffffffu, (int)me.done, j) != 0;      // generic path reports its own errors
            const signed char *qp = reinterpret_cast<const signed char *>(raw + qual);
            bool bad = false;
            for (uint32_t p = lane; p < lenj && !skip; p += 32)
                bad |= p >= lead && p < trail && (int)qp[p] - o.in_off > FQ_MAX_QUALITY_SCORE;
            if (__any_sync(0xffffffffu, bad)) {
                ge.bits |= kErrQualGt41;
                ge.rec = min(ge.rec, r0 + j);
                break;
            }
        }
    }

    // ---- phase 2a: one lane per read: PRE scalar statistics and the window
    const signed char *qp_mine = reinterpret_cast<const signed char *>(raw + me.rc.qual);
    const uint32_t len = me.rc.len;
    const QualAt qa{qp_mine, me.lead, me.trail, o.in_off};
    Window w{0, len, 0, 0, false, -1};
    if (!me.done) {
        lane_scalar_stats(kc, 0, len, me.cnt_ac, me.cnt_tg, (int)average_quality(me.sum_q, len, o.in_off));
        acc.reads += 1;
        acc.len += len;
        acc.max_pre_rows = max(acc.max_pre_rows, len);
        w = lane_window(kc, mate, r, len, qa);
    }

    // ---- phase 3a: cooperative pass over the bases outside the window (and, when a re-encoding can overflow, the max quality inside)
    uint32_t w_atc = me.cnt_ac, w_gn = me.cnt_tg, n_lowg = me.lowg;      // G -> N candidates: those of the whole read, minus the cut ones
    int sum_w = me.sum_q, max_qv = 0;
    {
        const bool partial = !me.done && (w.lo > 0 || w.lo + w.wl < len);
        const bool want = !me.done && (partial || (need_max && w.ret));
        uint32_t todo = __ballot_sync(0xffffffffu, want);
        while (todo) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            const uint32_t lenj = __shfl_sync(0xffffffffu, len, j);
            const uint8_t *sp = raw + __shfl_sync(0xffffffffu, me.rc.seq, j);
            const signed char *qp = reinterpret_cast<const signed char *>(raw + __shfl_sync(0xffffffffu, me.rc.qual, j));
            const uint32_t lo = __shfl_sync(0xffffffffu, w.lo, j), wl = __shfl_sync(0xffffffffu, w.wl, j);
            const uint32_t lead = __shfl_sync(0xffffffffu, me.lead, j), trail = __shfl_sync(0xffffffffu, me.trail, j);
            uint32_t r_atc = 0, r_gn = 0, nl = 0;
            int r_sum = 0, mq = need_max ? 0 : -1;
            removed_pass(kc, 0, sp, qp, lenj, lo, wl, lead, trail, r_atc, r_gn, r_sum, nl, mq);
            if ((int)lane == j) {
                w_atc -= r_atc;            // fields never borrow: removed counts <= totals per class
                w_gn -= r_gn;
                sum_w -= r_sum;
                n_lowg -= nl;
                max_qv = mq;
            }
        }
    }

    // ---- phase 2b: one lane per read: filters (trim.cpp:363-513)
    float ave_q = 0.0f;
    bool want_dinuc = false, want_run = false;
    float norm2 = 0.0f;
    uint32_t wA = f10(w_atc, 0), wT = f10(w_atc, 1), wC = f10(w_atc, 2), wG = f10(w_gn, 0), wN = f10(w_gn, 1);
    if (!me.done && w.ret) {
        wG -= n_lowg;                                       // G -> N replacement happens before the complexity filter
        wN += n_lowg;
        if (me.run_whole >= o.max_poly_n) want_run = true;  // exact run inside the window needed
    }
    {   // exact 'N' run of the window (rare: the whole read has a long enough run)
        uint32_t todo = __ballot_sync(0xffffffffu, want_run);
        uint32_t run_w = 0;
        while (todo) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            const uint8_t *sp = raw + __shfl_sync(0xffffffffu, me.rc.seq, j);
            const uint32_t rw = window_n_run(kc, sp, __shfl_sync(0xffffffffu, w.lo, j), __shfl_sync(0xffffffffu, w.wl, j));
            if ((int)lane == j) run_w = rw;
        }
        if (want_run && run_w >= o.max_poly_n) {            // trim.cpp:363-371
            atomicAdd(&H.filt()[FQ_READ_NN], 1u);
            atomicAdd(&H.filt()[FQ_BASE_NN], w.wl);
            w.flags |= FQ_RR_F_NN;
            if (!o.qc_only) w.ret = false;
        }
    }
    if (!me.done && w.ret) {
        ave_q = average_quality(sum_w, w.wl, o.in_off);
        if (ave_q < o.avg_q) {                              // trim.cpp:374-382
            atomicAdd(&H.filt()[FQ_READ_AVG_Q], 1u);
            atomicAdd(&H.filt()[FQ_BASE_AVG_Q], w.wl);
            w.flags |= FQ_RR_F_AVGQ;
            w.ret = false;
        }
    }
    bool lowc = false;
    if (!me.done && w.ret) {                                // low complexity, trim.cpp:405-513
        const float norm = (float)(1.0 / (double)w.wl);     // trim.cpp:483
        lowc = __fmul_rn((float)wA, norm) > o.lc || __fmul_rn((float)wT, norm) > o.lc ||
               __fmul_rn((float)wG, norm) > o.lc || __fmul_rn((float)wC, norm) > o.lc;
        if (!lowc) {
            norm2 = norm * 2.0f;                            // trim.cpp:499
            const uint32_t second = max(max(min(wA, wT), min(wC, wG)), min(max(wA, wT), max(wC, wG)));
            want_dinuc = __fmul_rn((float)second, norm2) > o.lc;   // a dinucleotide count <= second largest base count
        }
    }
    {   // dinucleotide counts (rare), cooperative
        uint32_t todo = __ballot_sync(0xffffffffu, want_dinuc);
        while (todo) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            const uint8_t *sp = raw + __shfl_sync(0xffffffffu, me.rc.seq, j);
            const QualAt qaj{reinterpret_cast<const signed char *>(raw + __shfl_sync(0xffffffffu, me.rc.qual, j)),
                             __shfl_sync(0xffffffffu, me.lead, j), __shfl_sync(0xffffffffu, me.trail, j), o.in_off};
            const bool res = dinucleotide_low_complexity(kc, sp, qaj, __shfl_sync(0xffffffffu, w.lo, j), __shfl_sync(0xffffffffu, w.wl, j),
                                                         __shfl_sync(0xffffffffu, norm2, j));
            if ((int)lane == j) lowc = res;
        }
    }
    if (!me.done && w.ret) {
        if (lowc) {
            atomicAdd(&H.filt()[FQ_READ_LOW_COMPLEXITY], 1u);
            atomicAdd(&H.filt()[FQ_BASE_LOW_COMPLEXITY], w.wl);
            w.flags |= FQ_RR_F_LOWCOMP;
            w.ret = false;
        } else if (need_max && max_qv + o.out_off > 127) {  // trim.cpp:516-525
            ge.bits |= kErrReencode;
            ge.rec = min(ge.rec, r);
        }
    }

    // ---- phase 2c: one lane per read: POST scalar statistics and the verdict (trim.cpp:527-548)
    // third line of the record: with LF line ends and a one-character line it sits two bytes behind the bases (same sector
    // as the end of the bases or the start of the qualities); anything but '+' bars the record from being block-copied
    uint32_t plus_bad = 0;
    if (lane < n_here && me.rc.qual == me.rc.seq + me.rc.len + 3 && raw[me.rc.seq + me.rc.len + 1] != '+') plus_bad = kResPlusBad;
    v_off5 = g_off5 | plus_bad;
    v_lenflags = g_lenflags;
    if (!me.done) {
        if (w.ret) {
            w.flags |= FQ_RR_VALID;
            acc.trimmed += 1;
            acc.trimmed_len += w.wl;
            acc.max_post_rows = max(acc.max_post_rows, w.off5 + w.wl);
            acc.max_post_len1 = max(acc.max_post_len1, w.wl + 1);
            const uint32_t p_atc = wA | (wT << 10) | (wC << 20), p_gn = wG | (wN << 10);
            lane_scalar_stats(kc, 1, w.wl, p_atc, p_gn, (int)ave_q);
        }
        const uint32_t masked = (me.lead > 0 || me.trail < len) ? kFlagMasked : 0u;
        v_off5 = w.off5 | plus_bad;
        v_lenflags = pack_len_flags(w.ret ? w.wl : 0, w.flags | masked);
        fq_read_result *dbg = mate ? a.dbg[1] : a.dbg[0];
        if (dbg) {
            fq_read_result d;
            d.offset_5 = w.off5;
            d.length = w.ret ? w.wl : 0;
            d.flags = (uint16_t)w.flags;
            d.adapter = (int16_t)w.best_adapter;
            d.avg_q = ave_q;
            dbg[r] = d;
        }
    }

    // ---- phase 3c: cooperative: the window of reads that turned out invalid (mode 1), surviving G->N (mode 2)
    {
        const bool inv = !me.done && !w.ret && w.wl > 0;
        const bool g2n = !me.done && w.ret && n_lowg > 0;
        uint32_t todo = __ballot_sync(0xffffffffu, inv || g2n);
        const uint32_t inv_mask = __ballot_sync(0xffffffffu, inv);
        while (todo) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            const uint8_t *sp = raw + __shfl_sync(0xffffffffu, me.rc.seq, j);
            const signed char *qp = reinterpret_cast<const signed char *>(raw + __shfl_sync(0xffffffffu, me.rc.qual, j));
            uint32_t d0 = 0, d1 = 0, d3 = 0;
            int d2 = 0, d4 = -1;
            removed_pass(kc, ((inv_mask >> j) & 1u) ? 1 : 2, sp, qp, __shfl_sync(0xffffffffu, len, j), __shfl_sync(0xffffffffu, w.lo, j),
                         __shfl_sync(0xffffffffu, w.wl, j), __shfl_sync(0xffffffffu, me.lead, j), __shfl_sync(0xffffffffu, me.trail, j),
                         d0, d1, d2, d3, d4);
        }
    }
}

// Fill the CTA's shared tables (call with all threads, then __syncthreads()).
__device__ __forceinline__ void init_shared_tables(const SmemHist &H, const TrimArgs &a, const DevOpts &o)
{
    const size_t n_words = SmemHist::words(a.smem_rows, a.comp_key_len);
    for (size_t i = threadIdx.x; i < n_words; i += blockDim.x) g_smem[i] = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < 258; i += blockDim.x) {
        if (i < 256) {
            H.lut()[i] = lut_entry(i);
            const int code = base_code_slow(i);
            H.lut_base()[i] = code < 5 ? make_uint2((SmemHist::kHist + (uint32_t)((2 * kQualCols + code) * a.smem_rows)) * 4u, code < 4 ? 1u << (8 * code) : 0u)
                                       : make_uint2(H.trash_bytes(), 0u);
            const int ch = (int)(signed char)i, qv = max(0, ch - o.in_off);
            H.lut_qual()[i] = qv <= FQ_MAX_QUALITY_SCORE ? make_uint2((SmemHist::kHist + (uint32_t)(qv * a.smem_rows)) * 4u, (uint32_t)ch)
                                                         : make_uint2(H.trash_bytes(), (uint32_t)(ch + kBadQual));
        } else H.lut_qual()[i] = make_uint2(H.trash_bytes(), 0u);       // pad entry (kQualPad) and its alignment filler
    }
}

// Merge at the end of the kernel: lane accumulators -> warp -> global; shared -> global (matrix.h:111-142 / trim.cpp:120-154).
__device__ __forceinline__ void flush_accumulators(const KernelCtx &kc, const LaneAcc &acc, const GroupErr &ge)
{
    const TrimArgs &a = kc.a;
    const SmemHist &H = kc.H;
    const StatsLayout &L = a.L;
    unsigned long long *const S = kc.S;
    const uint32_t lane = kc.lane, R = H.rows;
    {
        const uint32_t reads = warp_sum(acc.reads), trimmed = warp_sum(acc.trimmed);
        unsigned long long len_sum = acc.len, tlen_sum = acc.trimmed_len;
#pragma unroll
        for (int k = 16; k; k >>= 1) {
            len_sum += __shfl_xor_sync(0xffffffffu, len_sum, k);
            tlen_sum += __shfl_xor_sync(0xffffffffu, tlen_sum, k);
        }
        const uint32_t pre_rows = __reduce_max_sync(0xffffffffu, acc.max_pre_rows);
        const uint32_t post_rows = __reduce_max_sync(0xffffffffu, acc.max_post_rows);
        const uint32_t post_len1 = __reduce_max_sync(0xffffffffu, acc.max_post_len1);
        const uint32_t err_all = __reduce_or_sync(0xffffffffu, ge.bits);
        const uint32_t err_min = __reduce_min_sync(0xffffffffu, ge.rec);
        if (lane == 0) {
            if (reads) {
                gadd(&S[L.filter + FQ_TOTAL_COUNT], reads);
                gadd(&S[L.filter + FQ_TOTAL_NUMBER], reads);
                gadd(&S[L.filter + FQ_TOTAL_LENGTH], len_sum);
                atomicMax(&a.rows->pre_rows, pre_rows);
                atomicMax(&a.rows->pre_len_size, pre_rows + 1);
            }
            if (trimmed) {
                gadd(&S[L.filter + FQ_TOTAL_TRIMMED_NUMBER], trimmed);
                gadd(&S[L.filter + FQ_TOTAL_TRIMMED_LENGTH], tlen_sum);
                atomicMax(&a.rows->post_rows, post_rows);
                atomicMax(&a.rows->post_len_size, post_len1);
            }
            if (err_all) {
                atomicOr(&a.info->err, err_all);
                atomicMin(&a.info->err_record, err_min);
            }
        }
    }
    __syncthreads();
    auto flush = [&](const uint32_t *src, size_t dst, size_t cols, bool plus1) {
        // src is [cols][R (+1)], dst is [cols][L.rows (+1)]
        // every CTA walks the table from its own starting cell, so the 148 CTAs' atomics on one global cell do not
        // arrive together (32-bit index math: the tables hold at most 42 x 1024 cells)
        const uint32_t w = R + (plus1 ? 1u : 0u), W = L.rows + (plus1 ? 1u : 0u), total = (uint32_t)cols * w;
        const uint32_t rot = (uint32_t)(((unsigned long long)blockIdx.x * total) / gridDim.x);
        for (uint32_t k = threadIdx.x; k < total; k += blockDim.x) {
            uint32_t i = k + rot;
            if (i >= total) i -= total;
            const uint32_t v = src[i];
            if (v) gadd(&S[dst + (size_t)(i / w) * W + (i % w)], v);
        }
    };
    flush(H.preq(), L.pre_q, kQualCols, false);
    flush(H.remq(), L.rem_q, kQualCols, false);
    flush(H.preb(), L.pre_b, kBaseCols, false);
    flush(H.remb(), L.rem_b, kBaseCols, false);
    flush(H.g2n(), L.g2n, 1, false);
    flush(H.prelen(), L.pre_len, 1, true);
    flush(H.postlen(), L.post_len, 1, true);
    for (uint32_t i = threadIdx.x; i < 4 * kQualCols; i += blockDim.x) {
        const uint32_t v = H.qh()[i];
        if (v) gadd(&S[(i < kQualCols ? L.pre_rq : i < 2 * kQualCols ? L.pre_bq : i < 3 * kQualCols ? L.post_rq : L.post_bq) + (i % kQualCols)], v);
    }
    for (uint32_t i = threadIdx.x; i < FQ_NUM_STAT; i += blockDim.x) {
        const uint32_t v = H.filt()[i];
        if (v) gadd(&S[L.filter + i], v);
    }
    for (uint32_t i = threadIdx.x; i < 12; i += blockDim.x) {
        const uint32_t v = H.zero()[i];
        if (v) gadd(&S[(i < 6 ? L.pre_comp : L.post_comp) + (size_t)(i % 6) * kCompBins], v);
    }
    if (H.key != 0xffffffffu) {
        const uint32_t kw = H.key + 1;
        const float norm = composition_norm(H.key);
        for (uint32_t i = threadIdx.x; i < 14 * kw; i += blockDim.x) {
            const uint32_t v = H.compk()[i];
            if (!v) continue;
            const uint32_t which = i / (7 * kw), row = (i / kw) % 7, cnt = i % kw;
            const uint32_t hist = row < 5 ? row : 5;
            const uint32_t bin = composition_bin_n(norm, cnt) - (row == 6 ? 1u : 0u);
            gadd(&S[(which ? L.post_comp : L.pre_comp) + (size_t)hist * kCompBins + bin], v);
        }
    }
}

constexpr int kTrimThreads = FQ_TRIM_THREADS;

// KSEL: which register-resident phase-1 widths this instance carries (the host picks it from the batch's longest read):
//   4: reads <= 128 bases, 5: reads <= 160 bases, 0: all widths (5 and 10 chunks of 32 bases) + the chunked generic path for
//   longer reads.  Specialised instances keep the hot loop small: the kernel is issue-bound and instruction-cache sensitive.
// PLAIN: the option set of a default run (BWA_plus, no 5'/3' clip, no adapters, no G->N replacement, no re-encoding,
//   not --qc_only) is baked in, so every branch on those options disappears from the instance.
// Persistent grid, one CTA per SM; a warp takes groups of 32 consecutive reads of one mate (mate 1's groups first).
template <int KSEL, bool PLAIN>
__global__ void __launch_bounds__(kTrimThreads, 1) k_trim(const TrimArgs a, const DevOpts o_in)
{
    DevOpts o = o_in;
    if (PLAIN) {
        o.mode = FQ_MODE_BWA_PLUS;
        o.trim_5 = 0;
        o.trim_3 = 0;
        o.replace_q = 0;
        o.qc_only = 0;
        o.filter_adapter = 0;
        o.out_off = o.in_off;
    }
    SmemHist H{a.smem_rows, a.comp_key_len};
    init_shared_tables(H, a, o);
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    const KernelCtx kc{a, o, H, a.stats, lane, (uint32_t)__cvta_generic_to_shared(g_smem) + 4 * lane, min(a.n_mates, 1u)};
    const uint32_t groups_per_mate = (a.n_rec + 31) / 32, n_groups = groups_per_mate * a.n_mates;
    LaneAcc acc;
    GroupErr ge;
    for (uint32_t g = warp_global; g < n_groups; g += n_warps) {
        const uint32_t mate = g >= groups_per_mate ? 1u : 0u;
        const uint32_t r0 = (g - mate * groups_per_mate) * 32, r = r0 + lane;
        Rec rc{0, 0, 0, 0};
        if (r < a.n_rec) rc = (mate ? a.rec[1] : a.rec[0])[r];
        uint32_t off5 = 0, lenflags = 0;
        trim_group<KSEL, PLAIN>(kc, mate, r0, rc, acc, ge, off5, lenflags);
        if (r < a.n_rec) (mate ? a.res[1] : a.res[0])[r] = make_uint2(off5, lenflags);
    }
    flush_accumulators(kc, acc, ge);
}

}  // namespace fq
