// fq_api.cu -- C ABI (include/faqcs_b200.h) over the sm_100a kernels.
//
// One fq_ctx = one CUDA device + one stream + the run's accumulators (the
// filter_stats / adaptor_stats / PlotInfo trio main() owns, FaQCs.cpp:67-69).
// A batch goes through: frame -> [pair-id check] -> [adapter pass] -> trim ->
// route -> scan -> emit.  There is no CPU implementation of any of these steps
// in this library: without a CUDA device fq_create fails.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>          // types and enums only: the library is loaded with dlopen, not linked

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "fq_adapter.cuh"
#include "fq_common.cuh"
#include "fq_emit.cuh"
#include "fq_frame.cuh"
#include "fq_kmer.cuh"
#include "fq_trim.cuh"

using namespace fq;

#define FQ_STR2(x) #x
#define FQ_STR(x) FQ_STR2(x)

namespace {

std::string g_create_error;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        const size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct PinnedBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        const size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

const char *kPhiX = "__PhiX174_NC_001422__";                       // FaQCs.h:29-30
const char *kPhiXComplement = "__PhiX174_NC_001422_complement__";

int na_bits_host(char c)
{
    switch (c | 0x20) {
        case 'a': return 1;  case 'c': return 2;  case 'g': return 4;  case 't': return 8;
        case 'm': return 3;  case 'r': return 5;  case 's': return 6;  case 'v': return 7;
        case 'w': return 9;  case 'y': return 10; case 'h': return 11; case 'k': return 12;
        case 'd': return 13; case 'b': return 14; case 'n': return 15;
        default: return c == '-' ? 16 : -1;
    }
}

}  // namespace

struct fq_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    fq_options opt{};
    DevOpts dopt{};
    std::vector<std::string> adapter_names, adapter_seqs;
    AdapterSet aset{};
    DevBuf d_adp_codes, d_adp_off, d_adp_or;

    DevBuf d_raw[2], d_chunk[2], d_nl[2], d_rec[2], d_canon[2], d_adp[2], d_adp_best[2], d_res[2], d_dbg[2], d_tile, d_out[2][4];
    DevBuf d_raw_slot[2][2];           // pipelined submit: raw input slots
    DevBuf d_info, d_stats, d_rows;
    BatchInfo *h_info = nullptr;       // pinned
    StatsLayout L{};
    PinnedBuf h_out[2][4], h_dbg[2], h_stats, h_pieces[2][4];
    DevBuf d_pieces[2][4];
    bool pieces = false;               // fq_set_output_pieces
    int out_slot = 0;                   // which d_out / h_out set process_common fills
    bool async_out = false;            // D2H on s_out, caller waits on ev_out
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    // pipelined path: an input slot is busy from submit until run returns, an output slot from run until wait
    struct Submitted { bool busy = false; size_t n1 = 0, n2 = 0; bool paired = false; uint64_t first = 0; int is_final = 0; uint64_t ticket = 0; } sub[2];
    struct Finished { bool busy = false; uint64_t ticket = 0; fq_batch_out out{}; } fin[2];
    uint64_t next_ticket = 0;
    bool debug_results = false;
    bool check_pair_ids = true;
    bool frame_exact[2] = {false, false};      // this mate's input is not plain LF text: skip the fast framing instance

    // k-mer rarefaction (fq_kmer_*): Options::kmer_rarefaction / kmer / split_size / num_subsample, PlotInfo::kmer_*
    struct Kmer {
        bool enabled = false, collecting = false;
        uint32_t k = 31, num_subsample = 0;
        uint64_t split_size = 1;
        uint64_t total_reads = 0;                  // TOTAL_NUMBER
        std::vector<fq_rarefaction> samples;       // finished points
        std::map<uint64_t, uint64_t> freq;         // kmer_frequency_histogram
        // current pass
        uint32_t pass_calls = 0;                   // trim() calls of this pass so far
        uint32_t pass_counted = 0;                 // calls [0, pass_counted) fed the table
        std::vector<std::pair<uint32_t, uint64_t>> pass_points;   // (call, TOTAL_NUMBER): distinct / total filled in at the end of the pass
        unsigned long long cap = 0, occupied_bound = 0;
        DevBuf d_slots, d_call_total, d_call_distinct, d_small, d_big, d_scalars;
        std::vector<uint64_t> flat;                // fq_kmer_view::frequency
    } kmer;

    // host-side stats views handed out by fq_stats
    std::vector<uint64_t> v_adapter_reads, v_adapter_bases, v_pre_q, v_post_q, v_pre_b, v_post_b, v_hist[4], v_pre_comp,
        v_post_comp, v_pre_len, v_post_len;

    std::string error;
    uint64_t launches = 0;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    float t_seg[5] = {0, 0, 0, 0, 0};
    int sm_count = 0;
    size_t smem_optin = 0;
    float allreduce_ms = 0.0f;
    const void *last_dev_out[4] = {nullptr, nullptr, nullptr, nullptr};
};

namespace {

fq_status fail(fq_ctx *c, fq_status code, const std::string &msg)
{
    if (c) c->error = msg; else g_create_error = msg;
    return code;
}

#define CK(call)                                                                                           \
    do {                                                                                                   \
        cudaError_t e__ = (call);                                                                          \
        if (e__ != cudaSuccess)                                                                            \
            return fail(ctx, FQ_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));           \
    } while (0)

uint32_t round_up(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }

// (Re)allocate the stats block for at least `rows` position rows, keeping the contents.
// The dynamic shared-memory limit of a kernel is state of the device's (primary) CUDA context, shared by every fq_ctx on
// that device -- and contexts on one device run on different host threads.  It is therefore raised ONCE per kernel and
// device to the opt-in maximum and never lowered: a per-launch value would race with the other context's launch.
cudaError_t allow_dynamic_smem(fq_ctx *ctx, const void *kernel)
{
    static std::mutex mu;
    static std::map<std::pair<int, const void *>, bool> done;
    std::lock_guard<std::mutex> lk(mu);
    bool &d = done[std::make_pair(ctx->device, kernel)];
    if (d) return cudaSuccess;
    cudaFuncAttributes fa{};
    cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(ctx->smem_optin - fa.sharedSizeBytes));      // static + dynamic <= opt-in
    if (e == cudaSuccess) d = true;
    return e;
}

__global__ void __launch_bounds__(256) k_merge_stats(unsigned long long *dst, unsigned long long *src, size_t n, uint32_t *dst_rows, uint32_t *src_rows)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        dst[i] += src[i];
        src[i] = 0;
    }
    if (blockIdx.x == 0 && threadIdx.x < 4) {
        dst_rows[threadIdx.x] = max(dst_rows[threadIdx.x], src_rows[threadIdx.x]);
        src_rows[threadIdx.x] = 0;
    }
}

fq_status ensure_stats_rows(fq_ctx *ctx, uint32_t rows)
{
    rows = std::max(64u, round_up(rows, 64));
    if (ctx->d_stats.p && rows <= ctx->L.rows) return FQ_OK;
    const StatsLayout oldL = ctx->L;
    const StatsLayout newL = StatsLayout::make(rows, ctx->opt.n_adapters);
    std::vector<unsigned long long> oldv, newv(newL.total, 0);
    if (ctx->d_stats.p) {
        oldv.resize(oldL.total);
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaMemcpy(oldv.data(), ctx->d_stats.p, oldL.total * 8, cudaMemcpyDeviceToHost));
        auto copy = [&](size_t so, size_t dso, size_t n) { std::copy(oldv.begin() + so, oldv.begin() + so + n, newv.begin() + dso); };
        copy(oldL.filter, newL.filter, 32);
        copy(oldL.adapter_reads, newL.adapter_reads, oldL.n_adapters);
        copy(oldL.adapter_bases, newL.adapter_bases, oldL.n_adapters);
        copy(oldL.pre_rq, newL.pre_rq, 4 * kQualCols);
        copy(oldL.pre_comp, newL.pre_comp, 12 * (size_t)kCompBins);
        for (int c = 0; c < kQualCols; ++c) {
            copy(oldL.pre_q + (size_t)c * oldL.rows, newL.pre_q + (size_t)c * newL.rows, oldL.rows);
            copy(oldL.rem_q + (size_t)c * oldL.rows, newL.rem_q + (size_t)c * newL.rows, oldL.rows);
        }
        for (int c = 0; c < kBaseCols; ++c) {
            copy(oldL.pre_b + (size_t)c * oldL.rows, newL.pre_b + (size_t)c * newL.rows, oldL.rows);
            copy(oldL.rem_b + (size_t)c * oldL.rows, newL.rem_b + (size_t)c * newL.rows, oldL.rows);
        }
        copy(oldL.g2n, newL.g2n, oldL.rows);
        copy(oldL.pre_len, newL.pre_len, oldL.rows + 1);
        copy(oldL.post_len, newL.post_len, oldL.rows + 1);
    }
    DevBuf nb;
    CK(nb.ensure(newL.total * 8));
    CK(cudaMemcpy(nb.p, newv.data(), newL.total * 8, cudaMemcpyHostToDevice));
    ctx->d_stats.release();
    ctx->d_stats = nb;
    ctx->L = newL;
    return FQ_OK;
}

fq_status upload_adapters(fq_ctx *ctx)
{
    const uint32_t n = ctx->opt.n_adapters;
    std::vector<uint8_t> codes, orb(n ? n : 1, 0);
    std::vector<uint32_t> off(n + 1, 0);
    for (uint32_t j = 0; j < n; ++j) {
        off[j] = (uint32_t)codes.size();
        for (char ch : ctx->adapter_seqs[j]) {
            const int b = na_bits_host(ch);
            if (b < 0) return fail(ctx, FQ_ERR_BASE, "seq_overlap.cpp:na_to_bits: Unknown base!");
            codes.push_back((uint8_t)b);
            orb[j] |= (uint8_t)b;
        }
    }
    off[n] = (uint32_t)codes.size();
    CK(ctx->d_adp_codes.ensure(codes.size() + 16));
    CK(ctx->d_adp_off.ensure(off.size() * 4));
    CK(ctx->d_adp_or.ensure(orb.size()));
    if (!codes.empty()) CK(cudaMemcpy(ctx->d_adp_codes.p, codes.data(), codes.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_adp_off.p, off.data(), off.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_adp_or.p, orb.data(), orb.size(), cudaMemcpyHostToDevice));
    ctx->aset.codes = ctx->d_adp_codes.as<uint8_t>();
    ctx->aset.offset = ctx->d_adp_off.as<uint32_t>();
    ctx->aset.or_bits = ctx->d_adp_or.as<uint8_t>();
    ctx->aset.n = n;
    ctx->aset.total = (uint32_t)codes.size();
    ctx->aset.any_bits = 0;
    return FQ_OK;
}

void refresh_dev_opts(fq_ctx *ctx)
{
    const fq_options &o = ctx->opt;
    DevOpts &d = ctx->dopt;
    d.mode = o.mode;
    d.quality = (int)(signed char)o.quality;
    d.trim_5 = o.trim_5;
    d.trim_3 = o.trim_3;
    d.min_len = o.min_read_length;
    d.max_poly_n = o.max_num_poly_N;
    d.avg_q = o.average_quality;
    d.lc = o.low_complexity_cutoff_ratio;
    d.match_rate = (float)(1.0 - (double)o.adapter_mismatch_rate);     // trim.cpp:969
    d.in_off = (int)(signed char)o.input_quality_offset;
    d.out_off = (int)(signed char)o.output_quality_offset;
    d.replace_q = o.replace_to_N_q;
    d.qc_only = o.qc_only;
    d.protect_5 = o.protect_5;
    d.filter_adapter = o.filter_adapter && o.n_adapters > 0;
    d.discard = o.discard_output;
    d.num_thread = o.num_thread;
    d.n_adapters = o.n_adapters;
}

// Framing of one mate: count -> scan -> (sync) -> scatter -> records.
fq_status frame_mate(fq_ctx *ctx, int m, const uint8_t *d_raw, size_t n, uint32_t *n_rec_out)
{
    *n_rec_out = 0;
    if (n == 0) return FQ_OK;
    if (n >= (1ull << 30)) return fail(ctx, FQ_ERR_ARG, "batch larger than 1 GiB per mate: split it");
    if ((reinterpret_cast<uintptr_t>(d_raw) & 15) != 0) return fail(ctx, FQ_ERR_ARG, "device input must be 16-byte aligned");
    BatchInfo *info = ctx->d_info.as<BatchInfo>();
    // single-pass segmented line index: one contiguous segment per resident warp, no cross-warp dependency
    int frame_ctas = 4;                                     // resident CTAs per SM -> exactly one wave of segments
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&frame_ctas, k_frame_lines<false>, kFrameThreads, 0));
    const uint32_t want_warps = (uint32_t)ctx->sm_count * (uint32_t)std::max(frame_ctas, 1) * (kFrameThreads / 32);
    uint32_t seg_bytes = (uint32_t)((n + want_warps - 1) / want_warps);
    seg_bytes = std::max<uint32_t>(kChunkBytes, (seg_bytes + kChunkBytes - 1) / kChunkBytes * kChunkBytes);
    const uint32_t n_seg = (uint32_t)((n + seg_bytes - 1) / seg_bytes);
    CK(ctx->d_chunk[m].ensure(((size_t)n_seg * 2 + 2) * 4));
    uint32_t *seg_count = ctx->d_chunk[m].as<uint32_t>();
    uint32_t *seg_base = seg_count + n_seg;
    uint32_t n_lines = 0;
    uint32_t seg_cap = seg_bytes / 24 + 64;                 // lines per segment region: >= 24 bytes per line on average
    for (int attempt = 0; attempt < 2; ++attempt) {
        CK(ctx->d_nl[m].ensure((size_t)n_seg * seg_cap * 4));
        const int grid = (n_seg + kFrameThreads / 32 - 1) / (kFrameThreads / 32);
        if (attempt) {
            CK(cudaMemsetAsync(&info->n_cr[m], 0, 4, ctx->stream));
            CK(cudaMemsetAsync(&info->n_cr_eol[m], 0, 4, ctx->stream));
            CK(cudaMemsetAsync(&info->seg_overflow, 0, 4, ctx->stream));
        }
        // the fast instance first (plain LF text); the exact instance redoes the mate only if the fast one met a control
        // character other than '\n' (CRLF input: remembered, later batches go straight to the exact instance)
        for (int pass = 0; pass < 2; ++pass) {
            const bool exact = ctx->frame_exact[m];
            if (exact) k_frame_lines<false><<<grid, kFrameThreads, 0, ctx->stream>>>(d_raw, n, seg_bytes, n_seg, ctx->d_nl[m].as<uint32_t>(), seg_cap, seg_count, info, m, 0);
            else k_frame_lines<true><<<grid, kFrameThreads, 0, ctx->stream>>>(d_raw, n, seg_bytes, n_seg, ctx->d_nl[m].as<uint32_t>(), seg_cap, seg_count, info, m, 0);
            k_scan_segments<<<1, 1024, 0, ctx->stream>>>(seg_count, n_seg, seg_base, info, m);
            ctx->launches += 2;
            CK(cudaMemcpyAsync(ctx->h_info, info, sizeof(BatchInfo), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            if (exact || !ctx->h_info->frame_exact[m]) break;
            ctx->frame_exact[m] = true;                   // the fast instance gave up: run the exact one
        }
        n_lines = ctx->h_info->n_lines[m];
        if (ctx->h_info->seg_overflow == 0) break;
        seg_cap = ctx->h_info->seg_overflow + 64;           // very short lines: repeat once with the exact need
    }
    const bool had_cr = ctx->h_info->n_cr[m] != 0;
    if (had_cr) {
        // CRLF input (or stray CRs): exact CR count, to be compared with the CRs that sit right before a '\n'
        CK(cudaMemsetAsync(&info->n_cr[m], 0, 4, ctx->stream));
        k_count_byte<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(d_raw, n, '\r', &info->n_cr[m]);
        ctx->launches++;
    }
    if (n_lines % 4 != 0) {
        static const char *msg[4] = {"", "fastq.cpp:next_read: Unable to read sequence", "fastq.cpp:next_read: Unable to read '+'",
                                     "fastq.cpp:next_read: Unable to read quality"};
        return fail(ctx, FQ_ERR_FORMAT, msg[n_lines % 4]);
    }
    const uint32_t n_rec = n_lines / 4;
    if (n_rec == 0) return FQ_OK;
    CK(ctx->d_rec[m].ensure((size_t)n_rec * sizeof(Rec)));
    CK(ctx->d_canon[m].ensure((size_t)n_rec));
    const LineIndex li{ctx->d_nl[m].as<uint32_t>(), seg_base, seg_cap, n_seg};
    k_build_records<<<(n_rec + 255) / 256, 256, 0, ctx->stream>>>(li, n_rec, ctx->d_rec[m].as<Rec>(), ctx->d_canon[m].as<uint8_t>(), info, m);
    ctx->launches++;
    if (had_cr) {
        // exact CR count against the CRs that close a line: any other one cuts its line short (strpbrk, fastq.cpp:44)
        CK(cudaMemcpyAsync(ctx->h_info, info, sizeof(BatchInfo), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->h_info->n_cr[m] != ctx->h_info->n_cr_eol[m]) {
            // the lengths (and a possible |seq| != |qual| verdict) of k_build_records are void for such lines: redo them
            if ((ctx->h_info->err & ~kErrLenMismatch) == 0) {
                const uint32_t zero = 0, none = 0xffffffffu;
                CK(cudaMemcpyAsync(&info->err, &zero, 4, cudaMemcpyHostToDevice, ctx->stream));
                CK(cudaMemcpyAsync(&info->err_record, &none, 4, cudaMemcpyHostToDevice, ctx->stream));
            }
            k_fix_lone_cr<<<(n_rec + 255) / 256, 256, 0, ctx->stream>>>(d_raw, ctx->d_rec[m].as<Rec>(), ctx->d_canon[m].as<uint8_t>(), n_rec, info, m);
            ctx->launches++;
        }
    }
    *n_rec_out = n_rec;
    return FQ_OK;
}

fq_status map_device_error(fq_ctx *ctx, const BatchInfo &hi)
{
    if (!hi.err) return FQ_OK;
    char where[64];
    snprintf(where, sizeof where, " (record %u)", hi.err_record);
    if (hi.err & kErrLenMismatch) return fail(ctx, FQ_ERR_FORMAT, std::string("fastq.cpp:next_read: |Sequence| != |Quality|") + where);
    if (hi.err & kErrPairId) return fail(ctx, FQ_ERR_FORMAT, std::string("FaQCs.cpp:trim: I/O error") + where);
    if (hi.err & kErrUnknownBase) return fail(ctx, FQ_ERR_BASE, std::string("seq_overlap.cpp:na_to_bits: Unknown base!") + where);
    if (hi.err & kErrQualGt41)
        return fail(ctx, FQ_ERR_QUALITY,
                    std::string("fastq.h:quality_score: Found a quality score value that is greater than the maximum allowed quality score") + where);
    if (hi.err & kErrReencode) return fail(ctx, FQ_ERR_QUALITY, std::string("trim.cpp: quality error!") + where);
    if (hi.err & kErrInternal) return fail(ctx, FQ_ERR_CUDA, "internal: a bulk copy did not complete");
    return fail(ctx, FQ_ERR_STATE, "unknown device error");
}

// ---- k-mer rarefaction: host side of fq_kmer.cuh ---------------------------------------------------------
KmerTable kmer_table_of(fq_ctx::Kmer &K)
{
    return KmerTable{K.d_slots.as<KmerSlot>(), K.cap - 1};
}

// (Re)allocate the table for at least `slots` slots (a power of two), keeping its entries.
fq_status kmer_resize(fq_ctx *ctx, unsigned long long slots)
{
    fq_ctx::Kmer &K = ctx->kmer;
    unsigned long long cap = 1ull << 20;
    while (cap < slots) cap <<= 1;
    if (cap <= K.cap) return FQ_OK;
    DevBuf ns;
    CK(ns.ensure(cap * sizeof(KmerSlot)));
    KmerTable to{ns.as<KmerSlot>(), cap - 1};
    k_kmer_clear<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(to);
    if (K.cap) k_kmer_rehash<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(kmer_table_of(K), to);
    ctx->launches += K.cap ? 2 : 1;
    CK(cudaStreamSynchronize(ctx->stream));
    K.d_slots.release();
    K.d_slots = ns;
    K.cap = cap;
    return FQ_OK;
}

// One batch: which trim() calls of the reference it spans, which of them still feed the table, where curve points fall
// (trim.cpp:157-185), then the counting kernel.
fq_status kmer_batch(fq_ctx *ctx, const uint8_t *d_r1, const uint8_t *d_r2, uint32_t n, int n_mates, uint32_t max_len, uint64_t first_record_index, int is_final, bool trimmed)
{
    fq_ctx::Kmer &K = ctx->kmer;
    if ((first_record_index % FQ_REF_BATCH) != 0 || (!is_final && (n % FQ_REF_BATCH) != 0))
        return fail(ctx, FQ_ERR_ARG, "k-mer rarefaction needs batches that start on a 32768-record boundary and hold a multiple of 32768 records");
    const uint32_t first_call = K.pass_calls;
    uint64_t counted_reads = 0;
    for (uint32_t b0 = 0; b0 < n; b0 += FQ_REF_BATCH) {
        const uint32_t reads = std::min<uint32_t>(FQ_REF_BATCH, n - b0);
        for (int m = 0; m < n_mates; ++m) {
            const uint32_t call = K.pass_calls++;
            if (K.collecting) {                              // this call's reads go through update_kmer (trim.cpp:260-262)
                if (call >= kKmerMaxCalls) return fail(ctx, FQ_ERR_ARG, "k-mer rarefaction: more than 65536 trim() batches while the curve is collected");
                K.pass_counted = call + 1;
                counted_reads += reads;
            }
            K.total_reads += reads;
            if (K.collecting) {                              // end of trim(): trim.cpp:157-185
                const uint64_t index = K.total_reads / K.split_size;
                const uint64_t n_points = K.samples.size() + K.pass_points.size();
                if (index > n_points && n_points < K.num_subsample) K.pass_points.push_back(std::make_pair(call, K.total_reads));
                if (n_points >= K.num_subsample) K.collecting = false;
            }
        }
    }
    if (K.pass_counted <= first_call || counted_reads == 0) return FQ_OK;
    // room for every k-mer this batch can add (load factor <= 1/2); the bound on the occupied slots is made exact when it matters
    const unsigned long long per_read = max_len >= K.k ? (unsigned long long)(max_len - K.k + 1) : 0ull;
    unsigned long long need = K.occupied_bound + counted_reads * per_read;
    if (2 * need > K.cap && K.cap) {
        CK(cudaMemsetAsync(K.d_scalars.p, 0, 16, ctx->stream));
        k_kmer_count<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(kmer_table_of(K), K.d_scalars.as<unsigned long long>());
        ctx->launches++;
        unsigned long long occ = 0;
        CK(cudaMemcpyAsync(&occ, K.d_scalars.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        K.occupied_bound = occ;
        need = occ + counted_reads * per_read;
    }
    if (2 * need > K.cap) {
        fq_status st = kmer_resize(ctx, 2 * need);
        if (st != FQ_OK) return st;
    }
    K.occupied_bound = need;
    KmerArgs ka{};
    ka.raw[0] = d_r1; ka.raw[1] = d_r2;
    ka.rec[0] = ctx->d_rec[0].as<Rec>(); ka.rec[1] = ctx->d_rec[1].as<Rec>();
    ka.res[0] = trimmed ? ctx->d_res[0].as<uint2>() : nullptr;
    ka.res[1] = trimmed ? ctx->d_res[1].as<uint2>() : nullptr;
    ka.n_rec = n; ka.n_mates = (uint32_t)n_mates; ka.k = K.k;
    ka.replace_q = ctx->dopt.replace_q; ka.in_off = ctx->dopt.in_off;
    ka.first_call = first_call; ka.stop_call = K.pass_counted;
    ka.T = kmer_table_of(K);
    ka.call_total = K.d_call_total.as<unsigned long long>();
    k_kmer<<<(n * n_mates + 255) / 256, 256, 0, ctx->stream>>>(ka);
    ctx->launches++;
    return FQ_OK;
}

fq_status process_common(fq_ctx *ctx, const uint8_t *d_r1, size_t n1, const uint8_t *d_r2, size_t n2, bool paired,
                         uint64_t first_record_index, int is_final, int copy_out, fq_batch_out *out)
{
    memset(out, 0, sizeof(*out));
    for (auto &p : ctx->last_dev_out) p = nullptr;
    if (ctx->opt.input_quality_offset == FQ_OFFSET_AUTO)
        return fail(ctx, FQ_ERR_STATE, "quality offset not set; call fq_autodetect on the first batch");
    refresh_dev_opts(ctx);
    ctx->dopt.paired = paired ? 1 : 0;
    const DevOpts o = ctx->dopt;
    if (o.filter_adapter && o.num_thread && (first_record_index % FQ_REF_BATCH) != 0)
        return fail(ctx, FQ_ERR_ARG, "thread-count emulation needs batches that start on a 32768-record boundary");
    BatchInfo *info = ctx->d_info.as<BatchInfo>();
    BatchInfo init{};
    init.err_record = 0xffffffffu;
    init.detect_key = ~0ull;
    *ctx->h_info = init;
    CK(cudaMemcpyAsync(info, ctx->h_info, sizeof(BatchInfo), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));

    uint32_t n_rec[2] = {0, 0};
    fq_status st = frame_mate(ctx, 0, d_r1, n1, &n_rec[0]);
    if (st != FQ_OK) return st;
    if (paired) {
        st = frame_mate(ctx, 1, d_r2, n2, &n_rec[1]);
        if (st != FQ_OK) return st;
        if (n_rec[0] != n_rec[1]) return fail(ctx, FQ_ERR_FORMAT, "FaQCs.cppI/O error");   // FaQCs.cpp:370-380 (sic)
    }
    const uint32_t n = n_rec[0];
    const int n_mates = paired ? 2 : 1;
    out->n_records = n;
    if (o.filter_adapter && o.num_thread && !is_final && (n % FQ_REF_BATCH) != 0)
        return fail(ctx, FQ_ERR_ARG, "thread-count emulation needs non-final batches of a multiple of 32768 records");
    if (n == 0) {
        CK(cudaStreamSynchronize(ctx->stream));
        return FQ_OK;
    }
    // second look at the framing scalars: max length, CR accounting, |seq| != |qual|
    CK(cudaMemcpyAsync(ctx->h_info, info, sizeof(BatchInfo), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    {
        BatchInfo hi = *ctx->h_info;
        st = map_device_error(ctx, hi);
        if (st != FQ_OK) return st;
    }
    CK(cudaEventRecord(ctx->ev[4], ctx->stream));
    const uint32_t max_len = std::max(ctx->h_info->max_len[0], ctx->h_info->max_len[1]);
    if (max_len > kResLenMask) return fail(ctx, FQ_ERR_ARG, "read longer than 16 Mi bases");
    st = ensure_stats_rows(ctx, max_len);
    if (st != FQ_OK) return st;
    unsigned long long *S = ctx->d_stats.as<unsigned long long>();

    // read ids of the two mates (FaQCs.cpp:383-389): k_emit compares them while both headers pass through it;
    // --qc_only emits nothing, so it keeps the stand-alone kernel
    if (paired && ctx->check_pair_ids && o.qc_only) {
        k_check_pair_ids<<<(n + 255) / 256, 256, 0, ctx->stream>>>(d_r1, ctx->d_rec[0].as<Rec>(), d_r2, ctx->d_rec[1].as<Rec>(), n, info);
        ctx->launches++;
    }
    for (int m = 0; m < n_mates; ++m) {
        CK(ctx->d_res[m].ensure((size_t)n * sizeof(uint2)));
        if (ctx->debug_results) CK(ctx->d_dbg[m].ensure((size_t)n * sizeof(fq_read_result)));
    }

    // ---- adapter pass
    if (o.filter_adapter) {
        AdapterArgs aa{};
        for (int m = 0; m < n_mates; ++m) {
            CK(ctx->d_adp[m].ensure((size_t)n * sizeof(uint2)));
            CK(ctx->d_adp_best[m].ensure((size_t)n * sizeof(int32_t)));
            aa.raw[m] = m ? d_r2 : d_r1;
            aa.rec[m] = ctx->d_rec[m].as<Rec>();
            aa.adp[m] = ctx->d_adp[m].as<uint2>();
            aa.adp_best[m] = ctx->d_adp_best[m].as<int32_t>();
        }
        aa.n_rec = n;
        aa.n_mates = n_mates;
        aa.max_len = max_len;
        aa.stats = S;
        aa.L = ctx->L;
        aa.info = info;
        aa.first_index = first_record_index;
        aa.end_index = is_final ? first_record_index + n : ~0ull;
        // shared memory plan: offsets, codes, [adapter bit planes], per warp read codes + mask (+ read planes)
        const size_t smem_cap = ctx->smem_optin - 2048;          // the kernel also has 1 KiB of static shared memory
        const uint32_t mask_words = (max_len + 31) >> 5;
        uint32_t plane_words = 0, has_gap = 0, rpad = 1;
        for (uint32_t j = 0; j < ctx->aset.n; ++j) {
            const uint32_t wt = ((uint32_t)ctx->adapter_seqs[j].size() + 31) / 32;
            plane_words += 5 * wt;
            rpad = std::max(rpad, wt + 1);
            if (ctx->adapter_seqs[j].find('-') != std::string::npos) has_gap = 1;
        }
        aa.rpad = rpad;
        const size_t base_fixed = (size_t)(ctx->aset.n + 1) * 8 + ((ctx->aset.total + 3) & ~3u);
        const size_t per_warp_plain = ((max_len + 3) & ~3u) + 4 * (size_t)mask_words;
        const size_t per_warp_planes = per_warp_plain + 20 * ((size_t)mask_words + 2 * rpad);
        // segment sweep: four more read masks, the candidate bits and per-read bounds per warp; two 16-byte mask sets + 8 bytes per segment of the adapters
        uint32_t n_seg = 0, sweep_pure = 1;
        for (uint32_t j = 0; j < ctx->aset.n; ++j) {
            const size_t T = ctx->adapter_seqs[j].size();
            if (T && T <= kSweepMaxLen) {
                n_seg += (uint32_t)((T + 31) / 32);
                if (ctx->adapter_seqs[j].find_first_not_of("ACGTacgt") != std::string::npos) sweep_pure = 0;
            }
        }
        const size_t per_warp_sweep = 16 * ((size_t)mask_words + 2 * rpad) + 4 * (((size_t)ctx->aset.n + 31) / 32) + (((size_t)n_seg + 3) & ~(size_t)3);
        const size_t fixed_sweep = 40 * (size_t)n_seg + 16 + 8 + 4 * (((size_t)ctx->aset.n + 31) / 32);          // + alignment, + unswept bits, + origin / longest
        aa.use_planes = base_fixed + (size_t)plane_words * 4 + 4 * per_warp_planes <= smem_cap ? 1 : 0;
        aa.sweep = aa.use_planes && n_seg && base_fixed + (size_t)plane_words * 4 + fixed_sweep + 4 * (per_warp_planes + per_warp_sweep) <= smem_cap ? 1 : 0;
        aa.n_seg = aa.sweep ? n_seg : 0;
        aa.sweep_pure = sweep_pure;
        aa.plane_words = aa.use_planes ? plane_words : 0;
        aa.has_gap = has_gap;
        const size_t fixed = base_fixed + (size_t)aa.plane_words * 4 + (aa.sweep ? fixed_sweep : 0);
        const size_t per_warp = (aa.use_planes ? per_warp_planes : per_warp_plain) + (aa.sweep ? per_warp_sweep : 0);
        int warps = 8;
        while (warps > 1 && fixed + warps * per_warp > smem_cap) warps >>= 1;
        const size_t smem = fixed + warps * per_warp;
        if (smem > smem_cap) return fail(ctx, FQ_ERR_ARG, "adapter set + read length exceed shared memory");
        CK(allow_dynamic_smem(ctx, reinterpret_cast<const void *>(k_adapter)));
        int per_sm = 1;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_adapter, warps * 32, smem));
        const int grid = std::max(1, std::min<int>((n * n_mates + warps - 1) / warps, ctx->sm_count * std::max(per_sm, 1)));
        k_adapter<<<grid, warps * 32, smem, ctx->stream>>>(aa, o, ctx->aset);
        ctx->launches++;
    }

    // ---- trim / filter / statistics
    CK(cudaEventRecord(ctx->ev[1], ctx->stream));
    {
        TrimArgs ta{};
        for (int m = 0; m < n_mates; ++m) {
            ta.raw[m] = m ? d_r2 : d_r1;
            ta.rec[m] = ctx->d_rec[m].as<Rec>();
            ta.adp[m] = o.filter_adapter ? ctx->d_adp[m].as<uint2>() : nullptr;
            ta.adp_best[m] = o.filter_adapter ? ctx->d_adp_best[m].as<int32_t>() : nullptr;
            ta.res[m] = ctx->d_res[m].as<uint2>();
            ta.dbg[m] = ctx->debug_results ? ctx->d_dbg[m].as<fq_read_result>() : nullptr;
        }
        ta.n_rec = n;
        ta.n_mates = n_mates;
        ta.stats = S;
        ta.L = ctx->L;
        ta.rows = ctx->d_rows.as<StatsRows>();
        ta.info = info;
        // shared histogram rows: cover the longest read if it fits next to the composition tables
        ta.comp_key_len = max_len <= 1023 ? max_len : 0xffffffffu;
        const size_t budget = std::min<size_t>(ctx->smem_optin, 200 * 1024);
        uint32_t rows = round_up(std::max(max_len, 1u), 32);
        while (rows > 32 && SmemHist::words(rows, ta.comp_key_len) * 4 > budget) rows -= 32;
        rows = std::min(rows, ctx->L.rows);
        ta.smem_rows = rows;
        const size_t smem = SmemHist::words(rows, ta.comp_key_len) * 4;
        // instance with only the phase-1 width(s) this batch needs (smaller hot loop)
        using TrimKernel = void (*)(const TrimArgs, const DevOpts);
        const bool plain = o.mode == FQ_MODE_BWA_PLUS && !o.trim_5 && !o.trim_3 && !o.replace_q && !o.qc_only && !o.filter_adapter && o.in_off == o.out_off;
        // width-specialised instances have no generic path: every read of the batch must fit the shared-memory rows
        const int ksel = rows < max_len ? 0 : max_len <= 128 ? 4 : max_len <= 160 ? 5 : 0;
        const TrimKernel kern = plain ? (ksel == 4 ? (TrimKernel)k_trim<4, true> : ksel == 5 ? (TrimKernel)k_trim<5, true> : (TrimKernel)k_trim<0, true>)
                                      : (ksel == 4 ? (TrimKernel)k_trim<4, false> : ksel == 5 ? (TrimKernel)k_trim<5, false> : (TrimKernel)k_trim<0, false>);
        CK(allow_dynamic_smem(ctx, reinterpret_cast<const void *>(kern)));
        const int threads = kTrimThreads;
        int per_sm = 1;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
        per_sm = std::max(per_sm, 1);
        {   // tuning knob: cap the resident k_trim CTAs per SM (room for another context's kernels on the same device)
            static const int cap = [] { const char *e = getenv("FAQCS_B200_TRIM_CTAS_PER_SM"); return e ? atoi(e) : 0; }();
            if (cap > 0) per_sm = std::min(per_sm, cap);
        }
        const uint32_t groups = ((n + 31) / 32) * n_mates;
        const int grid = std::max(1, std::min<int>((groups + threads / 32 - 1) / (threads / 32), ctx->sm_count * per_sm));
        kern<<<grid, threads, smem, ctx->stream>>>(ta, o);
        ctx->launches++;
    }
    CK(cudaEventRecord(ctx->ev[2], ctx->stream));

    // ---- k-mer rarefaction: raw reads under --qc_only (trim.cpp:260-262), trimmed survivors otherwise (trim.cpp:545-547)
    if (ctx->kmer.enabled) {
        st = kmer_batch(ctx, d_r1, d_r2, n, n_mates, max_len, first_record_index, is_final, !o.qc_only);
        if (st != FQ_OK) return st;
    }

    // ---- route / scan / emit
    EmitArgs ea{};
    ea.n_rec = n;
    ea.n_tiles = (n + kTile - 1) / kTile;
    ea.check_ids = paired && ctx->check_pair_ids && !o.qc_only;
    {   // rounds in which a warp stages its 32 records: the smallest split whose average slab leaves 8 % headroom
        const double avg = (double)std::max(n1, n2) / (double)std::max<uint32_t>(n, 1u);
        ea.parts = 1;
        while (ea.parts < 8 && (32.0 / ea.parts) * avg * 1.08 + 32.0 > (double)kEmitSlab) ea.parts *= 2;
    }
    const bool pieces = ctx->pieces && !o.qc_only;
    ea.prefetch_tiles = (uint32_t)(ctx->sm_count * FQ_EMIT_MIN_CTAS);      // one wave of k_emit CTAs
    if (const char *pf = getenv("FAQCS_B200_EMIT_PREFETCH_TILES")) ea.prefetch_tiles = (uint32_t)atoi(pf);
    CK(ctx->d_tile.ensure((size_t)ea.n_tiles * (pieces ? 12 : 4) * 4));
    ea.tile_sum = ctx->d_tile.as<uint32_t>();
    ea.info = info;
    ea.stats = S;
    ea.filter_off = ctx->L.filter;
    for (int m = 0; m < n_mates; ++m) {
        ea.raw[m] = m ? d_r2 : d_r1;
        ea.raw_bytes[m] = m ? n2 : n1;
        ea.rec[m] = ctx->d_rec[m].as<Rec>();
        ea.res[m] = ctx->d_res[m].as<uint2>();
        ea.canon[m] = ctx->d_canon[m].as<uint8_t>();
    }
    if (!o.qc_only) {
        // a trimmed record is never longer than its raw record, so the inputs bound the outputs; the unpaired stream of a
        // paired run takes mate 1 OR mate 2 of each pair, i.e. up to sum_i max(rec1_i, rec2_i) <= n1 + n2 bytes
        const size_t cap[4] = {paired ? n1 : 0, paired ? n2 : 0, paired ? n1 + n2 : n1, o.discard ? n1 + n2 : 0};
        for (int s = 0; s < 4; ++s) {
            if (cap[s]) CK(ctx->d_out[ctx->out_slot][s].ensure(cap[s] + 16));
            ea.out[s] = ctx->d_out[ctx->out_slot][s].as<uint8_t>();
            if (pieces && cap[s]) {       // at most one piece per record and mate
                CK(ctx->d_pieces[ctx->out_slot][s].ensure(((size_t)n * (s == 3 && paired ? 2 : 1) + 1) * sizeof(fq_out_piece)));
                ea.pieces[s] = ctx->d_pieces[ctx->out_slot][s].as<fq_out_piece>();
            }
        }
    }
    if (pieces) {
        k_route_pieces<<<ea.n_tiles, kTile, 0, ctx->stream>>>(ea, o);
        k_scan_tiles<<<12, 1024, 0, ctx->stream>>>(ea.tile_sum, ea.n_tiles, info);
        using EmitKernel = void (*)(const EmitArgs, const DevOpts);
        const EmitKernel kern = (!o.replace_q && o.in_off == o.out_off) ? (EmitKernel)k_emit_pieces<true> : (EmitKernel)k_emit_pieces<false>;
        kern<<<ea.n_tiles, kTile, 0, ctx->stream>>>(ea, o);
        ctx->launches += 3;
    } else {
    k_route<<<ea.n_tiles, kTile, 0, ctx->stream>>>(ea, o);
    k_scan_tiles<<<4, 1024, 0, ctx->stream>>>(ea.tile_sum, ea.n_tiles, info);
    ctx->launches += 2;
    }
    if (!o.qc_only && !pieces) {
        using EmitKernel = void (*)(const EmitArgs, const DevOpts);
        const EmitKernel kern = (!o.replace_q && o.in_off == o.out_off) ? (EmitKernel)k_emit<true> : (EmitKernel)k_emit<false>;
        CK(allow_dynamic_smem(ctx, reinterpret_cast<const void *>(kern)));
        kern<<<ea.n_tiles, kTile, kEmitSmem, ctx->stream>>>(ea, o);
        ctx->launches++;
    }
    CK(cudaEventRecord(ctx->ev[3], ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_info, info, sizeof(BatchInfo), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    cudaEventElapsedTime(&ctx->t_seg[0], ctx->ev[0], ctx->ev[3]);
    cudaEventElapsedTime(&ctx->t_seg[1], ctx->ev[0], ctx->ev[4]);
    cudaEventElapsedTime(&ctx->t_seg[2], ctx->ev[4], ctx->ev[1]);
    cudaEventElapsedTime(&ctx->t_seg[3], ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&ctx->t_seg[4], ctx->ev[2], ctx->ev[3]);
    const BatchInfo hi = *ctx->h_info;
    st = map_device_error(ctx, hi);
    if (st != FQ_OK) return st;
    out->n_valid[0] = hi.n_valid[0];
    out->n_valid[1] = hi.n_valid[1];
    out->paired_read_number = hi.paired_reads;
    out->paired_base_length = hi.paired_bases;
    for (int s = 0; s < 4; ++s) {
        out->bytes[s] = o.qc_only ? 0 : hi.out_bytes[s];
        ctx->last_dev_out[s] = out->bytes[s] ? ctx->d_out[ctx->out_slot][s].p : nullptr;
        out->n_pieces[s] = pieces ? hi.out_pieces[s] : 0;
        out->literal_bytes[s] = pieces ? hi.out_literal[s] : 0;
    }
    if (copy_out) {
        // D2H of the four streams: on the compute stream (synchronous API) or, pipelined, on the copy-out
        // stream behind an event so that it overlaps the next batch's kernels
        cudaStream_t cs = ctx->stream;
        if (ctx->async_out) {
            CK(cudaEventRecord(ctx->ev_comp[ctx->out_slot], ctx->stream));
            CK(cudaStreamWaitEvent(ctx->s_out, ctx->ev_comp[ctx->out_slot], 0));
            cs = ctx->s_out;
        }
        for (int s = 0; s < 4; ++s) {
            if (!out->bytes[s]) continue;
            // pieces mode: only the literal bytes and the piece list travel
            const size_t nb = pieces ? out->literal_bytes[s] : out->bytes[s];
            if (nb) {
                CK(ctx->h_out[ctx->out_slot][s].ensure(nb));
                CK(cudaMemcpyAsync(ctx->h_out[ctx->out_slot][s].p, ctx->d_out[ctx->out_slot][s].p, nb, cudaMemcpyDeviceToHost, cs));
                out->data[s] = static_cast<const uint8_t *>(ctx->h_out[ctx->out_slot][s].p);
            }
            if (pieces && out->n_pieces[s]) {
                const size_t pb = out->n_pieces[s] * sizeof(fq_out_piece);
                CK(ctx->h_pieces[ctx->out_slot][s].ensure(pb));
                CK(cudaMemcpyAsync(ctx->h_pieces[ctx->out_slot][s].p, ctx->d_pieces[ctx->out_slot][s].p, pb, cudaMemcpyDeviceToHost, cs));
                out->pieces[s] = static_cast<const fq_out_piece *>(ctx->h_pieces[ctx->out_slot][s].p);
            }
        }
        if (ctx->async_out) CK(cudaEventRecord(ctx->ev_out[ctx->out_slot], ctx->s_out));
    }
    if (ctx->debug_results) {
        for (int m = 0; m < n_mates; ++m) {
            CK(ctx->h_dbg[m].ensure((size_t)n * sizeof(fq_read_result)));
            CK(cudaMemcpyAsync(ctx->h_dbg[m].p, ctx->d_dbg[m].p, (size_t)n * sizeof(fq_read_result), cudaMemcpyDeviceToHost, ctx->stream));
            out->results[m] = static_cast<const fq_read_result *>(ctx->h_dbg[m].p);
        }
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return FQ_OK;
}

}  // namespace

extern "C" {

int fq_abi_version(void) { return FQ_ABI_VERSION; }

const char *fq_build_info(void) { return "faqcs_b200 " __DATE__ " sm_100a (nvcc " FQ_STR(__CUDACC_VER_MAJOR__) "." FQ_STR(__CUDACC_VER_MINOR__) ")"; }

fq_status fq_create(const fq_options *opt, int device, fq_ctx **out)
{
    fq_ctx *ctx = nullptr;
    if (!opt || !out) return fail(nullptr, FQ_ERR_ARG, "fq_create: null argument");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
        return fail(nullptr, FQ_ERR_NO_DEVICE, "no CUDA device: faqcs_b200 has no CPU fallback");
    if (device < 0 || device >= n_dev) return fail(nullptr, FQ_ERR_ARG, "fq_create: bad device index");
    if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, FQ_ERR_CUDA, "cudaSetDevice failed");
    ctx = new fq_ctx();
    ctx->device = device;
    ctx->opt = *opt;
    for (uint32_t i = 0; i < opt->n_adapters; ++i) {
        ctx->adapter_names.push_back(opt->adapters[i].name ? opt->adapters[i].name : "");
        ctx->adapter_seqs.push_back(opt->adapters[i].seq ? opt->adapters[i].seq : "");
    }
    ctx->opt.adapters = nullptr;
    cudaDeviceProp prop{};
    fq_status st = FQ_OK;
    auto boot = [&]() -> fq_status {
        CK(cudaGetDeviceProperties(&prop, device));
        ctx->sm_count = prop.multiProcessorCount;
        ctx->smem_optin = prop.sharedMemPerBlockOptin;
        CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
        for (int k = 0; k < 2; ++k) {
            CK(cudaEventCreateWithFlags(&ctx->ev_in[k], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ctx->ev_comp[k], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ctx->ev_out[k], cudaEventDisableTiming));
        }
        for (auto &e : ctx->ev) CK(cudaEventCreate(&e));
        CK(cudaMallocHost(reinterpret_cast<void **>(&ctx->h_info), sizeof(BatchInfo)));
        CK(ctx->d_info.ensure(sizeof(BatchInfo)));
        CK(ctx->d_rows.ensure(sizeof(StatsRows)));
        CK(cudaMemset(ctx->d_rows.p, 0, sizeof(StatsRows)));
        fq_status s2 = upload_adapters(ctx);
        if (s2 != FQ_OK) return s2;
        return ensure_stats_rows(ctx, 320);
    };
    st = boot();
    if (st != FQ_OK) {
        g_create_error = ctx->error;
        fq_destroy(ctx);
        return st;
    }
    refresh_dev_opts(ctx);
    *out = ctx;
    return FQ_OK;
}

void fq_destroy(fq_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (int m = 0; m < 2; ++m) {
        ctx->d_raw[m].release(); ctx->d_chunk[m].release(); ctx->d_nl[m].release(); ctx->d_rec[m].release(); ctx->d_canon[m].release();
        ctx->d_adp[m].release(); ctx->d_adp_best[m].release(); ctx->d_res[m].release(); ctx->d_dbg[m].release();
        ctx->h_dbg[m].release();
    }
    for (int k = 0; k < 2; ++k) {
        for (int s = 0; s < 4; ++s) { ctx->d_out[k][s].release(); ctx->h_out[k][s].release(); ctx->d_pieces[k][s].release(); ctx->h_pieces[k][s].release(); }
        for (int m = 0; m < 2; ++m) ctx->d_raw_slot[k][m].release();
        if (ctx->ev_in[k]) cudaEventDestroy(ctx->ev_in[k]);
        if (ctx->ev_comp[k]) cudaEventDestroy(ctx->ev_comp[k]);
        if (ctx->ev_out[k]) cudaEventDestroy(ctx->ev_out[k]);
    }
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
    ctx->d_tile.release(); ctx->d_info.release(); ctx->d_stats.release(); ctx->d_rows.release();
    ctx->d_adp_codes.release(); ctx->d_adp_off.release(); ctx->d_adp_or.release();
    for (DevBuf *b : {&ctx->kmer.d_slots, &ctx->kmer.d_call_total, &ctx->kmer.d_call_distinct, &ctx->kmer.d_small,
                      &ctx->kmer.d_big, &ctx->kmer.d_scalars}) b->release();
    ctx->h_stats.release();
    if (ctx->h_info) cudaFreeHost(ctx->h_info);
    for (auto &e : ctx->ev) if (e) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *fq_last_error(const fq_ctx *ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

fq_status fq_set_debug_results(fq_ctx *ctx, int enable)
{
    if (!ctx) return FQ_ERR_ARG;
    ctx->debug_results = enable != 0;
    return FQ_OK;
}

fq_status fq_set_output_pieces(fq_ctx *ctx, int enable)
{
    if (!ctx) return FQ_ERR_ARG;
    ctx->pieces = enable != 0;
    return FQ_OK;
}

fq_status fq_set_check_pair_ids(fq_ctx *ctx, int enable)
{
    if (!ctx) return FQ_ERR_ARG;
    ctx->check_pair_ids = enable != 0;
    return FQ_OK;
}

fq_status fq_set_quality(fq_ctx *ctx, int32_t quality)
{
    if (!ctx) return FQ_ERR_ARG;
    ctx->opt.quality = quality;
    refresh_dev_opts(ctx);
    return FQ_OK;
}

void *fq_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}

void fq_host_free(void *p) { if (p) cudaFreeHost(p); }

fq_status fq_autodetect(fq_ctx *ctx, const uint8_t *r1, size_t n1, const uint8_t *r2, size_t n2,
                        int32_t *input_quality_offset, int32_t *quality)
{
    if (!ctx || (!r1 && n1) || (!r2 && n2)) return FQ_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const bool paired = r2 != nullptr;
    if (ctx->opt.input_quality_offset == FQ_OFFSET_AUTO) {
        int detected[2] = {0, 0};
        for (int m = 0; m < (paired ? 2 : 1); ++m) {
            const uint8_t *h = m ? r2 : r1;
            const size_t n = m ? n2 : n1;
            BatchInfo init{};
            init.err_record = 0xffffffffu;
            init.detect_key = ~0ull;
            *ctx->h_info = init;
            CK(cudaMemcpyAsync(ctx->d_info.p, ctx->h_info, sizeof(BatchInfo), cudaMemcpyHostToDevice, ctx->stream));
            uint32_t n_rec = 0;
            if (n) {
                CK(ctx->d_raw[m].ensure(n + 16));
                CK(cudaMemcpyAsync(ctx->d_raw[m].p, h, n, cudaMemcpyHostToDevice, ctx->stream));
                fq_status st = frame_mate(ctx, m, ctx->d_raw[m].as<uint8_t>(), n, &n_rec);
                if (st != FQ_OK) return st;
            }
            const uint32_t look = std::min<uint32_t>(n_rec, FQ_REF_BATCH);
            if (look) {
                k_detect_offset<<<(look * 32 + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_raw[m].as<uint8_t>(), ctx->d_rec[m].as<Rec>(), look,
                                                                                 ctx->d_info.as<BatchInfo>());
                ctx->launches++;
            }
            CK(cudaMemcpyAsync(ctx->h_info, ctx->d_info.p, sizeof(BatchInfo), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            fq_status st = map_device_error(ctx, *ctx->h_info);
            if (st != FQ_OK) return st;
            if (ctx->h_info->detect_key == ~0ull)
                return fail(ctx, FQ_ERR_OFFSET, "trim.cpp:auto_detect_quality_offset: Unknown quality format!");
            detected[m] = (int)(ctx->h_info->detect_key & 0xff);
        }
        if (paired && detected[0] != detected[1])
            return fail(ctx, FQ_ERR_OFFSET, "FaQCs.cpp:process_paired: I/O Error (inconsistent quality offset detection between reads one and two)");
        ctx->opt.input_quality_offset = detected[0];
    }
    // auto_detect_next_seq (trim.cpp:619-626): header of the first read starts with "@NS"
    if (ctx->opt.quality < 20 && n1 >= 3 && r1[0] == '@' && r1[1] == 'N' && r1[2] == 'S') {
        bool first_line = true;     // "@NS" must sit inside the first header line
        for (int i = 0; i < 3; ++i) first_line &= (r1[i] != '\n' && r1[i] != '\r');
        if (first_line) ctx->opt.quality = 20;
    }
    refresh_dev_opts(ctx);
    if (input_quality_offset) *input_quality_offset = ctx->opt.input_quality_offset;
    if (quality) *quality = ctx->opt.quality;
    return FQ_OK;
}

fq_status fq_process_host(fq_ctx *ctx, const uint8_t *r1, size_t n1, const uint8_t *r2, size_t n2,
                          uint64_t first_record_index, int is_final, fq_batch_out *out)
{
    if (!ctx || !out || (!r1 && n1) || (!r2 && n2)) return FQ_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const bool paired = r2 != nullptr;
    CK(ctx->d_raw[0].ensure(n1 + 16));
    if (n1) CK(cudaMemcpyAsync(ctx->d_raw[0].p, r1, n1, cudaMemcpyHostToDevice, ctx->stream));
    if (paired) {
        CK(ctx->d_raw[1].ensure(n2 + 16));
        if (n2) CK(cudaMemcpyAsync(ctx->d_raw[1].p, r2, n2, cudaMemcpyHostToDevice, ctx->stream));
    }
    ctx->out_slot = 0;
    ctx->async_out = false;
    return process_common(ctx, ctx->d_raw[0].as<uint8_t>(), n1, paired ? ctx->d_raw[1].as<uint8_t>() : nullptr, n2, paired,
                          first_record_index, is_final, 1, out);
}

// ---- pipelined host path: submit (H2D on its own stream) / run (kernels) / wait (D2H on its own stream) ----
fq_status fq_submit_host(fq_ctx *ctx, const uint8_t *r1, size_t n1, const uint8_t *r2, size_t n2,
                         uint64_t first_record_index, int is_final, uint64_t *ticket)
{
    if (!ctx || !ticket || (!r1 && n1) || (!r2 && n2)) return FQ_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const uint64_t t = ctx->next_ticket;
    const int slot = (int)(t & 1);
    if (ctx->sub[slot].busy) return fail(ctx, FQ_ERR_STATE, "fq_submit_host: the batch submitted two tickets ago has not been run yet");
    fq_ctx::Submitted &p = ctx->sub[slot];
    p = fq_ctx::Submitted{};
    p.busy = true;
    p.n1 = n1; p.n2 = n2; p.paired = r2 != nullptr; p.first = first_record_index; p.is_final = is_final; p.ticket = t;
    CK(ctx->d_raw_slot[slot][0].ensure(n1 + 16));
    if (n1) CK(cudaMemcpyAsync(ctx->d_raw_slot[slot][0].p, r1, n1, cudaMemcpyHostToDevice, ctx->s_in));
    if (p.paired) {
        CK(ctx->d_raw_slot[slot][1].ensure(n2 + 16));
        if (n2) CK(cudaMemcpyAsync(ctx->d_raw_slot[slot][1].p, r2, n2, cudaMemcpyHostToDevice, ctx->s_in));
    }
    CK(cudaEventRecord(ctx->ev_in[slot], ctx->s_in));
    ctx->next_ticket++;
    *ticket = t;
    return FQ_OK;
}

fq_status fq_run(fq_ctx *ctx, uint64_t ticket)
{
    if (!ctx) return FQ_ERR_ARG;
    const int slot = (int)(ticket & 1);
    fq_ctx::Submitted &p = ctx->sub[slot];
    if (!p.busy || p.ticket != ticket) return fail(ctx, FQ_ERR_STATE, "fq_run: unknown or already processed ticket");
    if (ctx->fin[slot].busy) return fail(ctx, FQ_ERR_STATE, "fq_run: the outputs of the ticket two submissions ago have not been collected with fq_wait");
    // the per-read debug verdicts live in one buffer per mate, not one per output slot
    if (ctx->debug_results) return fail(ctx, FQ_ERR_STATE, "fq_run: per-read debug results are only available through the synchronous entry points");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_in[slot], 0));
    ctx->out_slot = slot;
    ctx->async_out = true;
    fq_ctx::Finished &f = ctx->fin[slot];
    fq_status st = process_common(ctx, ctx->d_raw_slot[slot][0].as<uint8_t>(), p.n1, p.paired ? ctx->d_raw_slot[slot][1].as<uint8_t>() : nullptr,
                                  p.n2, p.paired, p.first, p.is_final, 1, &f.out);
    ctx->async_out = false;
    p.busy = false;                       // the kernels are done with the raw input slot
    if (st != FQ_OK) return st;
    f.busy = true;
    f.ticket = ticket;
    return FQ_OK;
}

fq_status fq_wait(fq_ctx *ctx, uint64_t ticket, fq_batch_out *out)
{
    if (!ctx || !out) return FQ_ERR_ARG;
    const int slot = (int)(ticket & 1);
    fq_ctx::Finished &f = ctx->fin[slot];
    if (!f.busy || f.ticket != ticket) return fail(ctx, FQ_ERR_STATE, "fq_wait: ticket has not been run");
    CK(cudaSetDevice(ctx->device));
    bool any = false;
    for (int s = 0; s < 4; ++s) any |= f.out.bytes[s] != 0;
    if (any) CK(cudaEventSynchronize(ctx->ev_out[slot]));
    *out = f.out;
    f.busy = false;
    return FQ_OK;
}

fq_status fq_process_device(fq_ctx *ctx, const void *d_r1, size_t n1, const void *d_r2, size_t n2,
                            uint64_t first_record_index, int is_final, int copy_out, fq_batch_out *out)
{
    if (!ctx || !out || (!d_r1 && n1) || (!d_r2 && n2)) return FQ_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    ctx->out_slot = 0;
    ctx->async_out = false;
    return process_common(ctx, static_cast<const uint8_t *>(d_r1), n1, static_cast<const uint8_t *>(d_r2), n2, d_r2 != nullptr,
                          first_record_index, is_final, copy_out, out);
}

fq_status fq_device_outputs(fq_ctx *ctx, const void *d_out[FQ_NUM_STREAM])
{
    if (!ctx || !d_out) return FQ_ERR_ARG;
    for (int s = 0; s < 4; ++s) d_out[s] = ctx->last_dev_out[s];
    return FQ_OK;
}

fq_status fq_last_timing(fq_ctx *ctx, float *ms, int n)
{
    if (!ctx || !ms) return FQ_ERR_ARG;
    for (int i = 0; i < n && i < 5; ++i) ms[i] = ctx->t_seg[i];
    return FQ_OK;
}

uint64_t fq_launch_count(const fq_ctx *ctx) { return ctx ? ctx->launches : 0; }

void *fq_stream(fq_ctx *ctx) { return ctx ? static_cast<void *>(ctx->stream) : nullptr; }

fq_status fq_stats_reserve_rows(fq_ctx *ctx, uint32_t rows)
{
    if (!ctx) return FQ_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    return ensure_stats_rows(ctx, rows);
}

fq_status fq_stats_device_buffer(fq_ctx *ctx, void **d_u64, size_t *n_u64, void **d_rows_u32x4)
{
    if (!ctx) return FQ_ERR_ARG;
    CK(cudaStreamSynchronize(ctx->stream));
    if (d_u64) *d_u64 = ctx->d_stats.p;
    if (n_u64) *n_u64 = ctx->L.total;
    if (d_rows_u32x4) *d_rows_u32x4 = ctx->d_rows.p;
    return FQ_OK;
}

fq_status fq_kmer_enable(fq_ctx *ctx, uint32_t k, uint64_t split_size, uint32_t num_subsample)
{
    if (!ctx || k < 2 || k > 31 || split_size == 0 || num_subsample == 0) return FQ_ERR_ARG;      // options.cpp:537-568
    CK(cudaSetDevice(ctx->device));
    fq_ctx::Kmer &K = ctx->kmer;
    K.enabled = K.collecting = true;
    K.k = k; K.split_size = split_size; K.num_subsample = num_subsample;
    CK(K.d_call_total.ensure((size_t)kKmerMaxCalls * 8));
    CK(K.d_call_distinct.ensure((size_t)kKmerMaxCalls * 8));
    CK(K.d_small.ensure((size_t)kKmerSmallCounts * 8));
    CK(K.d_big.ensure((size_t)(1u << 20) * 4));
    CK(K.d_scalars.ensure(64));
    CK(cudaMemset(K.d_call_total.p, 0, (size_t)kKmerMaxCalls * 8));
    return kmer_resize(ctx, 1ull << 24);
}

fq_status fq_kmer_end_pass(fq_ctx *ctx)
{
    if (!ctx) return FQ_ERR_ARG;
    fq_ctx::Kmer &K = ctx->kmer;
    if (!K.enabled) return FQ_OK;
    CK(cudaSetDevice(ctx->device));
    // the pass's table -> distinct k-mers per first call, histogram of counts
    const uint32_t big_cap = 1u << 20;
    CK(cudaMemsetAsync(K.d_call_distinct.p, 0, (size_t)kKmerMaxCalls * 8, ctx->stream));
    CK(cudaMemsetAsync(K.d_small.p, 0, (size_t)kKmerSmallCounts * 8, ctx->stream));
    CK(cudaMemsetAsync(K.d_scalars.p, 0, 64, ctx->stream));
    k_kmer_summarize<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(kmer_table_of(K), K.d_call_distinct.as<unsigned long long>(), K.d_small.as<unsigned long long>(),
                                                              K.d_big.as<uint32_t>(), big_cap, reinterpret_cast<uint32_t *>(K.d_scalars.as<uint8_t>() + 16),
                                                              K.d_scalars.as<unsigned long long>());
    ctx->launches++;
    const uint32_t nc = std::max(K.pass_calls, 1u);
    std::vector<unsigned long long> call_total(nc), call_distinct(nc), small(kKmerSmallCounts);
    unsigned long long n_distinct = 0;
    uint32_t n_big = 0;
    CK(cudaMemcpyAsync(call_total.data(), K.d_call_total.p, (size_t)std::min(nc, kKmerMaxCalls) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(call_distinct.data(), K.d_call_distinct.p, (size_t)std::min(nc, kKmerMaxCalls) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(small.data(), K.d_small.p, (size_t)kKmerSmallCounts * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&n_distinct, K.d_scalars.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&n_big, K.d_scalars.as<uint8_t>() + 16, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (n_big > big_cap) return fail(ctx, FQ_ERR_ARG, "k-mer rarefaction: more than 2^20 k-mers with counts above 65535");
    std::vector<uint32_t> big(n_big);
    if (n_big) CK(cudaMemcpy(big.data(), K.d_big.p, (size_t)n_big * 4, cudaMemcpyDeviceToHost));
    // points of the curve taken in this pass (trim.cpp:165-178): table size and instances after their call
    unsigned long long run_d = 0, run_t = 0, all_t = 0;
    size_t next = 0;
    for (uint32_t c = 0; c < std::min(nc, kKmerMaxCalls); ++c) {
        run_d += call_distinct[c];
        run_t += call_total[c];
        while (next < K.pass_points.size() && K.pass_points[next].first == c) {
            K.samples.push_back(fq_rarefaction{K.pass_points[next].second, run_d, run_t});
            ++next;
        }
    }
    all_t = run_t;
    // FaQCs.cpp:518-537 / 737-756: frequency histogram of the table, and one point if the curve is still empty
    for (uint32_t c = 1; c < kKmerSmallCounts; ++c)
        if (small[c]) K.freq[c] += small[c];
    for (uint32_t c : big) K.freq[c] += 1;
    if (K.collecting && K.samples.empty()) K.samples.push_back(fq_rarefaction{K.total_reads, n_distinct, all_t});
    // a new table for the next pass
    K.pass_calls = K.pass_counted = 0;
    K.pass_points.clear();
    K.occupied_bound = 0;
    k_kmer_clear<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(kmer_table_of(K));
    ctx->launches++;
    CK(cudaMemsetAsync(K.d_call_total.p, 0, (size_t)kKmerMaxCalls * 8, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return FQ_OK;
}

fq_status fq_kmer_results(fq_ctx *ctx, fq_kmer_view *view)
{
    if (!ctx || !view) return FQ_ERR_ARG;
    fq_ctx::Kmer &K = ctx->kmer;
    K.flat.clear();
    for (const auto &kv : K.freq) { K.flat.push_back(kv.first); K.flat.push_back(kv.second); }
    view->n_rarefaction = (uint32_t)K.samples.size();
    view->rarefaction = K.samples.data();
    view->n_frequency = K.flat.size() / 2;
    view->frequency = K.flat.data();
    return FQ_OK;
}

fq_status fq_reset_stats(fq_ctx *ctx)
{
    if (!ctx) return FQ_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemset(ctx->d_stats.p, 0, ctx->L.total * 8));
    CK(cudaMemset(ctx->d_rows.p, 0, sizeof(StatsRows)));
    return FQ_OK;
}

fq_status fq_stats(fq_ctx *ctx, fq_stats_view *v)
{
    if (!ctx || !v) return FQ_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    const StatsLayout &L = ctx->L;
    CK(ctx->h_stats.ensure(L.total * 8 + sizeof(StatsRows)));
    uint64_t *S = static_cast<uint64_t *>(ctx->h_stats.p);
    StatsRows rows{};
    CK(cudaMemcpy(S, ctx->d_stats.p, L.total * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&rows, ctx->d_rows.p, sizeof(StatsRows), cudaMemcpyDeviceToHost));
    memset(v, 0, sizeof(*v));
    for (int i = 0; i < FQ_NUM_STAT; ++i) v->filter_stats[i] = S[L.filter + i];
    const uint32_t na = L.n_adapters;
    ctx->v_adapter_reads.assign(S + L.adapter_reads, S + L.adapter_reads + na);
    ctx->v_adapter_bases.assign(S + L.adapter_bases, S + L.adapter_bases + na);
    // main(): phiX pseudo-adapters -> READ_PHIX/BASE_PHIX, the rest -> READ_ADAPTER/BASE_ADAPTER (FaQCs.cpp:92-127)
    for (uint32_t j = 0; j < na; ++j) {
        const bool phix = ctx->adapter_names[j] == kPhiX || ctx->adapter_names[j] == kPhiXComplement;
        v->filter_stats[phix ? FQ_READ_PHIX : FQ_READ_ADAPTER] += ctx->v_adapter_reads[j];
        v->filter_stats[phix ? FQ_BASE_PHIX : FQ_BASE_ADAPTER] += ctx->v_adapter_bases[j];
    }
    v->n_adapters = na;
    v->adapter_reads = ctx->v_adapter_reads.data();
    v->adapter_bases = ctx->v_adapter_bases.data();
    const uint32_t pr = std::min(rows.pre_rows, L.rows), qr = std::min(rows.post_rows, L.rows);
    v->pre_rows = pr;
    v->post_rows = qr;
    ctx->v_pre_q.assign((size_t)pr * kQualCols, 0);
    ctx->v_post_q.assign((size_t)qr * kQualCols, 0);
    ctx->v_pre_b.assign((size_t)pr * kBaseCols, 0);
    ctx->v_post_b.assign((size_t)qr * kBaseCols, 0);
    for (uint32_t p = 0; p < pr; ++p) {
        for (int c = 0; c < kQualCols; ++c) ctx->v_pre_q[(size_t)p * kQualCols + c] = S[L.pre_q + (size_t)c * L.rows + p];
        for (int c = 0; c < kBaseCols; ++c) ctx->v_pre_b[(size_t)p * kBaseCols + c] = S[L.pre_b + (size_t)c * L.rows + p];
    }
    for (uint32_t p = 0; p < qr; ++p) {          // post = pre - removed (+ G->N into column N)
        for (int c = 0; c < kQualCols; ++c)
            ctx->v_post_q[(size_t)p * kQualCols + c] = S[L.pre_q + (size_t)c * L.rows + p] - S[L.rem_q + (size_t)c * L.rows + p];
        for (int c = 0; c < kBaseCols; ++c)
            ctx->v_post_b[(size_t)p * kBaseCols + c] = S[L.pre_b + (size_t)c * L.rows + p] - S[L.rem_b + (size_t)c * L.rows + p] +
                                                       (c == 4 ? S[L.g2n + p] : 0);
    }
    v->pre_quality_matrix = ctx->v_pre_q.data();
    v->post_quality_matrix = ctx->v_post_q.data();
    v->pre_base_matrix = ctx->v_pre_b.data();
    v->post_base_matrix = ctx->v_post_b.data();
    const size_t hist_off[4] = {L.pre_rq, L.pre_bq, L.post_rq, L.post_bq};
    for (int h = 0; h < 4; ++h) ctx->v_hist[h].assign(S + hist_off[h], S + hist_off[h] + kQualCols);
    v->pre_read_quality_hist = ctx->v_hist[0].data();
    v->pre_base_quality_hist = ctx->v_hist[1].data();
    v->post_read_quality_hist = ctx->v_hist[2].data();
    v->post_base_quality_hist = ctx->v_hist[3].data();
    ctx->v_pre_comp.assign(S + L.pre_comp, S + L.pre_comp + 6 * (size_t)kCompBins);
    ctx->v_post_comp.assign(S + L.post_comp, S + L.post_comp + 6 * (size_t)kCompBins);
    v->pre_composition = ctx->v_pre_comp.data();
    v->post_composition = ctx->v_post_comp.data();
    v->pre_len_size = std::min(rows.pre_len_size, L.rows + 1);
    v->post_len_size = std::min(rows.post_len_size, L.rows + 1);
    ctx->v_pre_len.assign(S + L.pre_len, S + L.pre_len + v->pre_len_size);
    ctx->v_post_len.assign(S + L.post_len, S + L.post_len + v->post_len_size);
    v->pre_length_hist = ctx->v_pre_len.data();
    v->post_length_hist = ctx->v_post_len.data();
    return FQ_OK;
}


// ---- multi-GPU merge: NCCL, loaded at run time --------------------------------------------------------
namespace {
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi &nccl()
{
    static NcclApi api = [] {
        NcclApi a;
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (a.handle) break;
        }
        if (!a.handle) return a;
        auto sym = [&](const char *n) { return dlsym(a.handle, n); };
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
        a.CommInitAll = reinterpret_cast<decltype(a.CommInitAll)>(sym("ncclCommInitAll"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
        a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
        a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
        a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommInitAll && a.CommDestroy && a.AllReduce && a.GroupStart && a.GroupEnd && a.GetErrorString;
        return a;
    }();
    return api;
}
}  // namespace

struct fq_comm {
    ncclComm_t comm = nullptr;
    int device = 0;
};

#define NK(ctx_, call)                                                                                          \
    do {                                                                                                        \
        ncclResult_t r__ = (call);                                                                              \
        if (r__ != ncclSuccess) return fail(ctx_, FQ_ERR_CUDA, std::string(#call) + ": " + nccl().GetErrorString(r__)); \
    } while (0)

fq_status fq_comm_unique_id(uint8_t id[FQ_COMM_ID_BYTES])
{
    static_assert(sizeof(ncclUniqueId) == FQ_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    if (!id) return FQ_ERR_ARG;
    if (!nccl().ok) return fail(nullptr, FQ_ERR_STATE, "NCCL (libnccl.so.2) could not be loaded");
    ncclUniqueId u;
    NK(nullptr, nccl().GetUniqueId(&u));
    memcpy(id, &u, FQ_COMM_ID_BYTES);
    return FQ_OK;
}

fq_status fq_comm_init_rank(fq_ctx *ctx, int n_ranks, int rank, const uint8_t id[FQ_COMM_ID_BYTES], fq_comm **out)
{
    if (!ctx || !id || !out || n_ranks < 1 || rank < 0 || rank >= n_ranks) return FQ_ERR_ARG;
    if (!nccl().ok) return fail(ctx, FQ_ERR_STATE, "NCCL (libnccl.so.2) could not be loaded");
    CK(cudaSetDevice(ctx->device));
    ncclUniqueId u;
    memcpy(&u, id, FQ_COMM_ID_BYTES);
    fq_comm *c = new fq_comm();
    c->device = ctx->device;
    ncclResult_t r = nccl().CommInitRank(&c->comm, n_ranks, u, rank);
    if (r != ncclSuccess) { delete c; return fail(ctx, FQ_ERR_CUDA, std::string("ncclCommInitRank: ") + nccl().GetErrorString(r)); }
    *out = c;
    return FQ_OK;
}

fq_status fq_comm_init_all(fq_ctx *const *ctxs, int n, fq_comm **out)
{
    if (!ctxs || !out || n < 1) return FQ_ERR_ARG;
    fq_ctx *ctx = ctxs[0];
    if (!nccl().ok) return fail(ctx, FQ_ERR_STATE, "NCCL (libnccl.so.2) could not be loaded");
    std::vector<int> devs(n);
    std::vector<ncclComm_t> comms(n);
    for (int i = 0; i < n; ++i) devs[i] = ctxs[i]->device;
    NK(ctx, nccl().CommInitAll(comms.data(), n, devs.data()));
    for (int i = 0; i < n; ++i) {
        out[i] = new fq_comm();
        out[i]->comm = comms[i];
        out[i]->device = devs[i];
    }
    return FQ_OK;
}

void fq_comm_destroy(fq_comm *c)
{
    if (!c) return;
    if (c->comm && nccl().ok) {
        cudaSetDevice(c->device);
        nccl().CommDestroy(c->comm);
    }
    delete c;
}

fq_status fq_allreduce_stats(fq_ctx *const *ctxs, int n, fq_comm *const *comms)
{
    if (!ctxs || !comms || n < 1) return FQ_ERR_ARG;
    fq_ctx *ctx = ctxs[0];
    if (!nccl().ok) return fail(ctx, FQ_ERR_STATE, "NCCL (libnccl.so.2) could not be loaded");
    for (int i = 0; i < n; ++i) {
        if (!ctxs[i] || !comms[i] || comms[i]->device != ctxs[i]->device) return fail(ctx, FQ_ERR_ARG, "fq_allreduce_stats: communicator and context are on different devices");
        CK(cudaSetDevice(ctxs[i]->device));
        CK(cudaStreamSynchronize(ctxs[i]->stream));
    }
    // 1. every rank must use the same layout: agree on the row capacity (MAX), grow where needed
    std::vector<DevBuf> d_cap(n);
    std::vector<uint32_t> cap(n);
    for (int i = 0; i < n; ++i) {
        CK(cudaSetDevice(ctxs[i]->device));
        CK(d_cap[i].ensure(4));
        cap[i] = ctxs[i]->L.rows;
        CK(cudaMemcpyAsync(d_cap[i].p, &cap[i], 4, cudaMemcpyHostToDevice, ctxs[i]->stream));
    }
    NK(ctx, nccl().GroupStart());
    for (int i = 0; i < n; ++i) NK(ctx, nccl().AllReduce(d_cap[i].p, d_cap[i].p, 1, ncclUint32, ncclMax, comms[i]->comm, ctxs[i]->stream));
    NK(ctx, nccl().GroupEnd());
    for (int i = 0; i < n; ++i) {
        CK(cudaSetDevice(ctxs[i]->device));
        CK(cudaMemcpyAsync(&cap[i], d_cap[i].p, 4, cudaMemcpyDeviceToHost, ctxs[i]->stream));
        CK(cudaStreamSynchronize(ctxs[i]->stream));
        fq_status st = ensure_stats_rows(ctxs[i], cap[i]);
        if (st != FQ_OK) return st;
        if (ctxs[i]->L.rows != cap[i] && ctxs[i]->L.rows != std::max(64u, round_up(cap[i], 64)))
            return fail(ctx, FQ_ERR_STATE, "fq_allreduce_stats: row capacities disagree");
    }
    // 2. the merge: SUM of the flat u64 block, MAX of the four row counters (timed: the first collective above also carries
    //    NCCL's one-off connection set-up)
    for (int i = 0; i < n; ++i) {
        CK(cudaSetDevice(ctxs[i]->device));
        CK(cudaEventRecord(ctxs[i]->ev[0], ctxs[i]->stream));
    }
    NK(ctx, nccl().GroupStart());
    for (int i = 0; i < n; ++i) {
        NK(ctx, nccl().AllReduce(ctxs[i]->d_stats.p, ctxs[i]->d_stats.p, ctxs[i]->L.total, ncclUint64, ncclSum, comms[i]->comm, ctxs[i]->stream));
        NK(ctx, nccl().AllReduce(ctxs[i]->d_rows.p, ctxs[i]->d_rows.p, 4, ncclUint32, ncclMax, comms[i]->comm, ctxs[i]->stream));
    }
    NK(ctx, nccl().GroupEnd());
    for (int i = 0; i < n; ++i) {
        CK(cudaSetDevice(ctxs[i]->device));
        CK(cudaEventRecord(ctxs[i]->ev[1], ctxs[i]->stream));
        CK(cudaStreamSynchronize(ctxs[i]->stream));
        cudaEventElapsedTime(&ctxs[i]->allreduce_ms, ctxs[i]->ev[0], ctxs[i]->ev[1]);
        d_cap[i].release();
    }
    return FQ_OK;
}

// Contexts that share a device (one in-flight batch each, so that consecutive batches overlap on the device): dst += src,
// src = 0.  The same merge as the all-reduce (sum of the block, max of the row counters) without a communicator.
fq_status fq_merge_stats(fq_ctx *dst, fq_ctx *src)
{
    if (!dst || !src || dst == src) return FQ_ERR_ARG;
    if (dst->device != src->device) return fail(dst, FQ_ERR_ARG, "fq_merge_stats: contexts on different devices (use fq_allreduce_stats)");
    if (dst->opt.n_adapters != src->opt.n_adapters) return fail(dst, FQ_ERR_ARG, "fq_merge_stats: contexts with different adapter lists");
    fq_ctx *const ctx = dst;
    CK(cudaSetDevice(dst->device));
    CK(cudaStreamSynchronize(src->stream));
    CK(cudaStreamSynchronize(dst->stream));
    const uint32_t rows = std::max(dst->L.rows, src->L.rows);
    fq_status st = ensure_stats_rows(dst, rows);
    if (st == FQ_OK) st = ensure_stats_rows(src, rows);
    if (st != FQ_OK) return st;
    if (dst->L.total != src->L.total) return fail(dst, FQ_ERR_STATE, "fq_merge_stats: row capacities disagree");
    k_merge_stats<<<dst->sm_count * 4, 256, 0, dst->stream>>>(dst->d_stats.as<unsigned long long>(), src->d_stats.as<unsigned long long>(), dst->L.total,
                                                           dst->d_rows.as<uint32_t>(), src->d_rows.as<uint32_t>());
    dst->launches++;
    CK(cudaStreamSynchronize(dst->stream));
    return FQ_OK;
}

float fq_last_allreduce_ms(const fq_ctx *ctx) { return ctx ? ctx->allreduce_ms : 0.0f; }

}  // extern "C"
