// fq_trim.cuh -- the per-read trim / filter / statistics kernel (trim_read, trim.cpp:225-551).
//
// One warp per read, lanes striped over consecutive base positions, persistent
// grid.  Statistics go to shared-memory privatised histograms stored transposed
// ([column][position], rows % 32 == 0) so that a warp's 32 consecutive positions
// always fall into 32 distinct banks; they are merged into the global u64 block
// once per CTA.  post-trim matrices are accumulated as "pre minus removed" (see
// StatsLayout), so an untrimmed surviving read costs one histogram update per base.
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
};

__device__ __forceinline__ int base_code(uint32_t c)
{
    c |= 0x20u;
    return c == 'a' ? 0 : c == 't' ? 1 : c == 'c' ? 2 : c == 'g' ? 3 : c == 'n' ? 4 : 5;
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// trim.cpp:553-576 with the reference's C types: float(int)/float(size_t) - float(char), clamped at 0.
__device__ __forceinline__ float average_quality(int total, uint32_t len, int offset)
{
    if (len == 0) return 0.0f;
    return fmaxf(0.0f, __fsub_rn(__fdiv_rn((float)total, (float)len), (float)offset));
}

// trim.cpp:860-874: bin = size_t(float(10000)/len * count), float arithmetic.
__device__ __forceinline__ uint32_t composition_bin(uint32_t len, uint32_t count)
{
    const float norm = len ? __fdiv_rn(10000.0f, (float)len) : 0.0f;
    return __float2uint_rz(__fmul_rn(norm, (float)count));
}

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

// Shared-memory histogram block of one CTA.
struct SmemHist {
    uint32_t *preq, *remq;   // [42][rows]
    uint32_t *preb, *remb;   // [5][rows]
    uint32_t *g2n;           // [rows]
    uint32_t *prelen, *postlen;   // [rows + 1]
    uint32_t *qh;            // [4][42]: pre_rq, pre_bq, post_rq, post_bq
    uint32_t *filt;          // [32]
    uint32_t rows;
    __device__ static size_t words(uint32_t rows) { return (size_t)rows * (2 * kQualCols + 2 * kBaseCols + 1) + 2 * (rows + 1) + 4 * kQualCols + 32; }
    __device__ void carve(uint32_t *base, uint32_t r)
    {
        rows = r;
        preq = base; base += (size_t)kQualCols * r;
        remq = base; base += (size_t)kQualCols * r;
        preb = base; base += (size_t)kBaseCols * r;
        remb = base; base += (size_t)kBaseCols * r;
        g2n = base; base += r;
        prelen = base; base += r + 1;
        postlen = base; base += r + 1;
        qh = base; base += 4 * kQualCols;
        filt = base;
    }
};

__device__ __forceinline__ void gadd(unsigned long long *p, unsigned long long v) { atomicAdd(p, v); }

__global__ void __launch_bounds__(512, 1) k_trim(const TrimArgs a, const DevOpts o)
{
    extern __shared__ uint32_t smem[];
    SmemHist H;
    H.carve(smem, a.smem_rows);
    const size_t n_words = SmemHist::words(a.smem_rows);
    for (size_t i = threadIdx.x; i < n_words; i += blockDim.x) smem[i] = 0;
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t total = a.n_rec * a.n_mates;
    const StatsLayout &L = a.L;
    unsigned long long *const S = a.stats;
    const uint32_t R = H.rows;

    // warp-uniform accumulators for the counters every read touches
    uint32_t acc_reads = 0, acc_trimmed = 0;
    unsigned long long acc_len = 0, acc_trimmed_len = 0;
    uint32_t max_pre_rows = 0, max_post_rows = 0, max_post_len1 = 0;
    uint32_t err = 0, err_rec = 0xffffffffu;

    for (uint32_t g = warp_global; g < total; g += n_warps) {
        const uint32_t mate = g >= a.n_rec ? 1 : 0;
        const uint32_t r = g - mate * a.n_rec;
        const Rec rc = a.rec[mate][r];
        const uint8_t *sp = a.raw[mate] + rc.seq;
        const signed char *qp = reinterpret_cast<const signed char *>(a.raw[mate] + rc.qual);
        const uint32_t len = rc.len;
        uint32_t flags = 0;

        // ---- terminal 'N' runs (trim.cpp:1191-1216), uppercase only
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

        // ---- PRE statistics over the whole (masked) read (trim.cpp:247-258)
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
            const int bc = in ? base_code(c) : 5;
            if (in && qv <= FQ_MAX_QUALITY_SCORE) {
                if (p < R) atomicAdd(&H.preq[qv * R + p], 1u);
                else gadd(&S[L.pre_q + (size_t)qv * L.rows + p], 1);
            }
            if (bc < 5) {
                if (p < R) atomicAdd(&H.preb[bc * R + p], 1u);
                else gadd(&S[L.pre_b + (size_t)bc * L.rows + p], 1);
            }
            nA += __popc(__ballot_sync(0xffffffffu, bc == 0));
            nT += __popc(__ballot_sync(0xffffffffu, bc == 1));
            nC += __popc(__ballot_sync(0xffffffffu, bc == 2));
            nG += __popc(__ballot_sync(0xffffffffu, bc == 3));
            nN += __popc(__ballot_sync(0xffffffffu, bc == 4));
        }
        if (__any_sync(0xffffffffu, bad_q)) { err |= kErrQualGt41; err_rec = min(err_rec, r); }
        sum_q = warp_sum_i(sum_q);
        {
            const int qbin = (int)average_quality(sum_q, len, o.in_off);
            uint32_t cnt = 0, bin = 0;
            // lanes 0..5: composition bins (A,T,C,G,N,GC); lane 6: length hist; lanes 7,8: avg-Q hists
            const uint32_t iC = composition_bin(len, nC), iG = composition_bin(len, nG);
            if (lane == 0) bin = composition_bin(len, nA);
            else if (lane == 1) bin = composition_bin(len, nT);
            else if (lane == 2) bin = iC;
            else if (lane == 3) bin = iG;
            else if (lane == 4) bin = composition_bin(len, nN);
            else if (lane == 5) bin = iG + iC;
            if (lane < 6) gadd(&S[L.pre_comp + (size_t)lane * kCompBins + bin], 1);
            else if (lane == 6) {
                if (len <= R) atomicAdd(&H.prelen[len], 1u);
                else gadd(&S[L.pre_len + len], 1);
            } else if (lane == 7) atomicAdd(&H.qh[0 * kQualCols + min(max(qbin, 0), 41)], 1u);
            else if (lane == 8) atomicAdd(&H.qh[1 * kQualCols + min(max(qbin, 0), 41)], len);
            (void)cnt;
        }
        acc_reads += 1;
        acc_len += len;
        max_pre_rows = max(max_pre_rows, len);

        // ---- window after adapter clip, 5'/3' clip (trim.cpp:270-314)
        uint32_t lo = 0, wl = len, off5 = 0;
        bool ret = true;
        int best_adapter = -1;
        if (o.filter_adapter && a.adp[mate]) {
            const uint2 v = a.adp[mate][r];
            best_adapter = a.adp_best[mate][r];
            if (best_adapter >= 0) flags |= FQ_RR_ADAPTER;
            if (len != v.y) {
                lo = v.x;
                wl = v.y;
                off5 += (v.y == 0) ? len : v.x;
            }
        }
        if (o.trim_5 && !o.qc_only) {
            if (o.trim_5 > wl) wl = 0;
            else { lo += o.trim_5; wl -= o.trim_5; off5 += o.trim_5; }
        }
        if (o.trim_3 && !o.qc_only) {
            if (o.trim_3 > wl) wl = 0;
            else wl -= o.trim_3;
        }
        if (wl < o.min_len || wl == 0) {                       // trim.cpp:317-323
            if (lane == 0) { atomicAdd(&H.filt[FQ_READ_LENGTH], 1u); atomicAdd(&H.filt[FQ_BASE_LENGTH], wl); }
            flags |= FQ_RR_F_LENGTH;
            ret = false;
        }
        if (!o.qc_only && ret) {                               // trim.cpp:325-360
            const uint32_t init_len = wl;
            uint32_t f5 = 0;
            if (o.mode == FQ_MODE_HARD) wl = hard_trim(qa, lo, (int)wl, o.quality, o.protect_5 != 0, f5);
            else if (o.mode == FQ_MODE_BWA) wl = bwa_trim(qa, lo, (int)wl, o.quality, f5);
            else wl = bwa_plus_trim(qa, lo, (int)wl, o.quality, o.protect_5 != 0, f5);
            off5 += f5;
            lo += f5;
            if (init_len != wl) {
                if (lane == 0) { atomicAdd(&H.filt[FQ_READ_QUAL_TRIM], 1u); atomicAdd(&H.filt[FQ_BASE_QUAL_TRIM], init_len - wl); }
                flags |= FQ_RR_QUAL_TRIMMED;
            }
            if (wl < o.min_len || wl == 0) {
                if (lane == 0) { atomicAdd(&H.filt[FQ_READ_LENGTH], 1u); atomicAdd(&H.filt[FQ_BASE_LENGTH], wl); }
                flags |= FQ_RR_F_LENGTH;
                ret = false;
            }
        }

        // ---- filters on the window (trim.cpp:363-513)
        float ave_q = 0.0f;
        uint32_t wA = 0, wT = 0, wC = 0, wG = 0, wN = 0, n_lowg = 0;
        if (ret) {
            int sum_w = 0;
            uint32_t run_carry = 0, max_run = 0;
            int max_qv = 0;
            for (uint32_t b = 0; b < wl; b += 32) {
                const uint32_t i = b + lane;
                const bool in = i < wl;
                const uint32_t p = lo + i;
                const uint32_t c = in ? sp[p] : 0;
                int qc = in ? (int)qp[p] : o.in_off;
                if (p < lead || p >= trail) qc = o.in_off;
                if (in) sum_w += qc;
                const int qv = max(0, qc - o.in_off);
                max_qv = max(max_qv, in ? qv : 0);
                const int bc = in ? base_code(c) : 5;
                const bool lowg = in && o.replace_q > 0 && c == 'G' && qv < (int)o.replace_q;
                const uint32_t mN = __ballot_sync(0xffffffffu, c == 'N');
                if (mN) {                                      // longest run of 'N' (count_poly_n, trim.cpp:578-597)
                    uint32_t x = mN;
                    const uint32_t head = __ffs(~x) ? (uint32_t)(__ffs(~x) - 1) : 32u;   // ones from bit 0
                    max_run = max(max_run, run_carry + head);
                    uint32_t k = 0;
                    while (x) { x &= x >> 1; ++k; }
                    max_run = max(max_run, k);
                    run_carry = (mN == 0xffffffffu) ? run_carry + 32 : (uint32_t)__clz(~mN);
                } else run_carry = 0;
                const uint32_t mLG = __ballot_sync(0xffffffffu, lowg);
                n_lowg += __popc(mLG);
                wA += __popc(__ballot_sync(0xffffffffu, bc == 0));
                wT += __popc(__ballot_sync(0xffffffffu, bc == 1));
                wC += __popc(__ballot_sync(0xffffffffu, bc == 2));
                wG += __popc(__ballot_sync(0xffffffffu, bc == 3 && !lowg));
                wN += __popc(__ballot_sync(0xffffffffu, bc == 4 || lowg));
            }
            sum_w = warp_sum_i(sum_w);
            if (max_run >= o.max_poly_n) {                     // trim.cpp:363-371
                if (lane == 0) { atomicAdd(&H.filt[FQ_READ_NN], 1u); atomicAdd(&H.filt[FQ_BASE_NN], wl); }
                flags |= FQ_RR_F_NN;
                if (!o.qc_only) ret = false;
            }
            ave_q = average_quality(sum_w, wl, o.in_off);
            if (ret && ave_q < o.avg_q) {                      // trim.cpp:374-382
                if (lane == 0) { atomicAdd(&H.filt[FQ_READ_AVG_Q], 1u); atomicAdd(&H.filt[FQ_BASE_AVG_Q], wl); }
                flags |= FQ_RR_F_AVGQ;
                ret = false;
            }
            if (ret) {                                         // low complexity, trim.cpp:405-513
                float norm = (float)(1.0 / (double)wl);
                bool lowc = __fmul_rn((float)wA, norm) > o.lc || __fmul_rn((float)wT, norm) > o.lc ||
                            __fmul_rn((float)wG, norm) > o.lc || __fmul_rn((float)wC, norm) > o.lc;
                if (!lowc) {
                    norm = norm * 2.0f;
                    // a dinucleotide count can not exceed the second largest base count
                    const uint32_t m1 = max(max(wA, wT), max(wC, wG));
                    const uint32_t second = max(max(min(wA, wT), min(wC, wG)), min(max(wA, wT), max(wC, wG)));
                    (void)m1;
                    if (__fmul_rn((float)second, norm) > o.lc) {
                        uint32_t dc[16];
#pragma unroll
                        for (int k = 0; k < 16; ++k) dc[k] = 0;
                        int prev_carry = 4;
                        for (uint32_t b = 0; b < wl; b += 32) {
                            const uint32_t i = b + lane;
                            const bool in = i < wl;
                            const uint32_t p = lo + i;
                            const uint32_t c = in ? sp[p] : 0;
                            int cur = in ? base_code(c) : 4;
                            if (cur > 3) cur = 4;
                            if (in && o.replace_q > 0 && c == 'G' && qa(p) < (int)o.replace_q) cur = 4;
                            int prev = __shfl_up_sync(0xffffffffu, cur, 1);
                            if (lane == 0) prev = prev_carry;
                            prev_carry = __shfl_sync(0xffffffffu, cur, 31);
                            const int code = (in && cur != 4 && prev != 4 && cur != prev) ? ((prev << 2) | cur) : -1;
#pragma unroll
                            for (int k = 0; k < 16; ++k) dc[k] += __popc(__ballot_sync(0xffffffffu, code == k));
                        }
#pragma unroll
                        for (int k = 0; k < 16; ++k) lowc |= __fmul_rn((float)dc[k], norm) > o.lc;
                    }
                }
                if (lowc) {
                    if (lane == 0) { atomicAdd(&H.filt[FQ_READ_LOW_COMPLEXITY], 1u); atomicAdd(&H.filt[FQ_BASE_LOW_COMPLEXITY], wl); }
                    flags |= FQ_RR_F_LOWCOMP;
                    ret = false;
                }
            }
            if (ret && o.in_off != o.out_off) {                // re-encode overflow, trim.cpp:516-525
                if (o.out_off + 41 > 127) {
                    max_qv = max(max_qv, __shfl_xor_sync(0xffffffffu, max_qv, 16));
                    max_qv = max(max_qv, __shfl_xor_sync(0xffffffffu, max_qv, 8));
                    max_qv = max(max_qv, __shfl_xor_sync(0xffffffffu, max_qv, 4));
                    max_qv = max(max_qv, __shfl_xor_sync(0xffffffffu, max_qv, 2));
                    max_qv = max(max_qv, __shfl_xor_sync(0xffffffffu, max_qv, 1));
                    if (max_qv + o.out_off > 127) { err |= kErrReencode; err_rec = min(err_rec, r); }
                }
            }
        }

        // ---- POST statistics (trim.cpp:527-548) as "removed" updates
        if (ret) {
            flags |= FQ_RR_VALID;
            acc_trimmed += 1;
            acc_trimmed_len += wl;
            max_post_rows = max(max_post_rows, off5 + wl);
            max_post_len1 = max(max_post_len1, wl + 1);
            const int qbin = min(max((int)ave_q, 0), 41);
            uint32_t bin = 0;
            const uint32_t iC = composition_bin(wl, wC), iG = composition_bin(wl, wG);
            if (lane == 0) bin = composition_bin(wl, wA);
            else if (lane == 1) bin = composition_bin(wl, wT);
            else if (lane == 2) bin = iC;
            else if (lane == 3) bin = iG;
            else if (lane == 4) bin = composition_bin(wl, wN);
            else if (lane == 5) bin = iG + iC;
            if (lane < 6) gadd(&S[L.post_comp + (size_t)lane * kCompBins + bin], 1);
            else if (lane == 6) {
                if (wl <= R) atomicAdd(&H.postlen[wl], 1u);
                else gadd(&S[L.post_len + wl], 1);
            } else if (lane == 7) atomicAdd(&H.qh[2 * kQualCols + qbin], 1u);
            else if (lane == 8) atomicAdd(&H.qh[3 * kQualCols + qbin], wl);
        }
        const bool whole = !ret;
        if (whole || lo > 0 || lo + wl < len || n_lowg) {
            for (uint32_t b = 0; b < len; b += 32) {
                const uint32_t p = b + lane;
                if (p >= len) continue;
                const bool inside = !whole && p >= lo && p < lo + wl;
                const uint32_t c = sp[p];
                const int qv = qa(p);
                if (!inside) {
                    const int bc = base_code(c);
                    if (qv <= FQ_MAX_QUALITY_SCORE) {
                        if (p < R) atomicAdd(&H.remq[qv * R + p], 1u);
                        else gadd(&S[L.rem_q + (size_t)qv * L.rows + p], 1);
                    }
                    if (bc < 5) {
                        if (p < R) atomicAdd(&H.remb[bc * R + p], 1u);
                        else gadd(&S[L.rem_b + (size_t)bc * L.rows + p], 1);
                    }
                } else if (n_lowg && o.replace_q > 0 && c == 'G' && qv < (int)o.replace_q) {
                    // surviving G -> N replacement: leaves column G, enters column N
                    if (p < R) { atomicAdd(&H.remb[3 * R + p], 1u); atomicAdd(&H.g2n[p], 1u); }
                    else { gadd(&S[L.rem_b + (size_t)3 * L.rows + p], 1); gadd(&S[L.g2n + p], 1); }
                }
            }
        }
        if (lane == 0) {
            a.res[mate][r] = make_uint2(off5, pack_len_flags(ret ? wl : 0, flags));
            if (a.dbg[mate]) {
                fq_read_result d;
                d.offset_5 = off5;
                d.length = ret ? wl : 0;
                d.flags = (uint16_t)flags;
                d.adapter = (int16_t)best_adapter;
                d.avg_q = ave_q;
                a.dbg[mate][r] = d;
            }
        }
    }

    // ---- merge: warp accumulators -> shared -> global (matrix.h:111-142 / trim.cpp:120-154)
    if (lane == 0) {
        if (acc_reads) {
            gadd(&S[L.filter + FQ_TOTAL_COUNT], acc_reads);
            gadd(&S[L.filter + FQ_TOTAL_NUMBER], acc_reads);
            gadd(&S[L.filter + FQ_TOTAL_LENGTH], acc_len);
        }
        if (acc_trimmed) {
            gadd(&S[L.filter + FQ_TOTAL_TRIMMED_NUMBER], acc_trimmed);
            gadd(&S[L.filter + FQ_TOTAL_TRIMMED_LENGTH], acc_trimmed_len);
        }
        if (acc_reads) {
            atomicMax(&a.rows->pre_rows, max_pre_rows);
            atomicMax(&a.rows->pre_len_size, max_pre_rows + 1);
        }
        if (acc_trimmed) {
            atomicMax(&a.rows->post_rows, max_post_rows);
            atomicMax(&a.rows->post_len_size, max_post_len1);
        }
        if (err) {
            atomicOr(&a.info->err, err);
            atomicMin(&a.info->err_record, err_rec);
        }
    }
    __syncthreads();
    auto flush = [&](const uint32_t *src, size_t dst, size_t cols, bool plus1) {
        // src is [cols][R (+1)], dst is [cols][L.rows (+1)]
        const size_t w = R + (plus1 ? 1 : 0), W = L.rows + (plus1 ? 1 : 0);
        for (size_t i = threadIdx.x; i < cols * w; i += blockDim.x) {
            const uint32_t v = src[i];
            if (v) gadd(&S[dst + (i / w) * W + (i % w)], v);
        }
    };
    flush(H.preq, L.pre_q, kQualCols, false);
    flush(H.remq, L.rem_q, kQualCols, false);
    flush(H.preb, L.pre_b, kBaseCols, false);
    flush(H.remb, L.rem_b, kBaseCols, false);
    flush(H.g2n, L.g2n, 1, false);
    flush(H.prelen, L.pre_len, 1, true);
    flush(H.postlen, L.post_len, 1, true);
    for (uint32_t i = threadIdx.x; i < 4 * kQualCols; i += blockDim.x) {
        const uint32_t v = H.qh[i];
        if (v) gadd(&S[(i < kQualCols ? L.pre_rq : i < 2 * kQualCols ? L.pre_bq : i < 3 * kQualCols ? L.post_rq : L.post_bq) + (i % kQualCols)], v);
    }
    for (uint32_t i = threadIdx.x; i < FQ_NUM_STAT; i += blockDim.x) {
        const uint32_t v = H.filt[i];
        if (v) gadd(&S[L.filter + i], v);
    }
}

}  // namespace fq
