// fq_common.cuh -- shared device/host definitions for the sm_100a FaQCs hot path.
//
// Data layout in HBM:
//   raw[m]      the FASTQ bytes of mate m exactly as inflated (never repacked)
//   nl_pos[m]   u32 byte offset of every '\n'                (framing, pass 1-2)
//   rec[m]      16-byte record descriptors {hdr, seq, qual offsets, length}
//   adp[m]      8-byte adapter verdict {start, length|adapter<<..}   (adapter pass)
//   res[m]      8-byte trim verdict {offset_5, length|flags}         (trim kernel)
//   stats       flat u64 accumulator block (StatsLayout)
//   out[s]      the four emitted FASTQ byte streams
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/faqcs_b200.h"

namespace fq {

constexpr int kWarp = 32;
constexpr int kQualCols = FQ_NUM_QUAL;        // 42
constexpr int kBaseCols = FQ_NUM_BASE;        // 5
constexpr int kCompBins = FQ_NUM_COMPOSITION_BIN;

// Device-side error bits (d_info.err); the host maps them to the reference's messages.
enum : uint32_t {
    kErrQualGt41 = 1u << 0,     // fastq.h:31-33
    kErrLenMismatch = 1u << 1,  // fastq.cpp:118-122
    kErrReencode = 1u << 2,     // trim.cpp:521-523
    kErrUnknownBase = 1u << 3,  // seq_overlap.cpp:409
    kErrPairId = 1u << 4,       // FaQCs.cpp:383-389
    kErrInternal = 1u << 6,     // a bulk copy did not complete (should never happen; reported instead of hanging)
};

struct Rec {
    uint32_t hdr;   // offset of the header line ('@...')
    uint32_t seq;   // offset of the sequence line
    uint32_t qual;  // offset of the quality line
    uint32_t len;   // bases (== quality chars)
};

// Options as the kernels see them (fq_options after host-side preprocessing).
struct DevOpts {
    int32_t mode;
    int32_t quality;          // (int)(char)Options::quality
    uint32_t trim_5, trim_3;
    uint32_t min_len;
    uint32_t max_poly_n;
    float avg_q;
    float lc;
    float match_rate;         // float(1.0 - rate), trim.cpp:969
    int32_t in_off, out_off;
    uint32_t replace_q;
    int32_t qc_only, protect_5, filter_adapter, discard, paired;
    uint32_t num_thread;      // Q3 emulation (0 = off)
    uint32_t n_adapters;
};

// Batch-level scalars written by the framing kernels and read back by the host.
struct BatchInfo {
    uint32_t n_lines[2];
    uint32_t max_len[2];
    uint32_t n_cr[2];          // '\r' bytes seen
    uint32_t n_cr_eol[2];      // '\r' immediately before a line's '\n'
    uint32_t err;              // kErr* bits
    uint32_t err_record;       // smallest record index that raised an error
    uint32_t seg_overflow;     // largest per-segment line count that did not fit its index region (0 = none)
    uint32_t frame_exact[2];   // the fast framing kernel met a control character that is not '\n': the exact kernel must run
    uint32_t seg_overflow_fast;
    uint32_t pad0;
    unsigned long long detect_key;   // autodetect: (record << 8 | offset), min over decisive reads
    unsigned long long out_bytes[4];
    unsigned long long out_pieces[4];      // pieces mode: pieces / literal bytes per stream
    unsigned long long out_literal[4];
    unsigned long long n_valid[2];
    unsigned long long paired_reads, paired_bases;
};

// Flat u64 statistics block.  Position-indexed arrays have `rows` capacity and are
// stored TRANSPOSED ([column][row]) so that a warp whose lanes hold consecutive
// positions touches consecutive words.  post = pre - rem (+ g2n for base column N):
// the quality value of a surviving base is the same before and after trimming, so
// only the bases that were trimmed away or belong to discarded reads are counted
// a second time.
struct StatsLayout {
    uint32_t rows;         // capacity of position-indexed arrays (multiple of 32)
    uint32_t n_adapters;
    // offsets in u64 units
    size_t filter, adapter_reads, adapter_bases;
    size_t pre_q, rem_q;   // [42][rows]
    size_t pre_b, rem_b;   // [5][rows]
    size_t g2n;            // [rows]   G->N replacements that survived (trim.cpp:390-403)
    size_t pre_rq, pre_bq, post_rq, post_bq;   // [42] each
    size_t pre_comp, post_comp;                // [6][10001]
    size_t pre_len, post_len;                  // [rows + 1]
    size_t total;

    __host__ __device__ static StatsLayout make(uint32_t rows, uint32_t n_adapters)
    {
        StatsLayout L;
        L.rows = rows;
        L.n_adapters = n_adapters;
        size_t o = 0;
        L.filter = o; o += 32;
        L.adapter_reads = o; o += n_adapters;
        L.adapter_bases = o; o += n_adapters;
        L.pre_rq = o; o += kQualCols;
        L.pre_bq = o; o += kQualCols;
        L.post_rq = o; o += kQualCols;
        L.post_bq = o; o += kQualCols;
        L.pre_comp = o; o += 6 * (size_t)kCompBins;
        L.post_comp = o; o += 6 * (size_t)kCompBins;
        L.pre_q = o; o += (size_t)kQualCols * rows;
        L.rem_q = o; o += (size_t)kQualCols * rows;
        L.pre_b = o; o += (size_t)kBaseCols * rows;
        L.rem_b = o; o += (size_t)kBaseCols * rows;
        L.g2n = o; o += rows;
        L.pre_len = o; o += (size_t)rows + 1;
        L.post_len = o; o += (size_t)rows + 1;
        L.total = o;
        return L;
    }
};

// rows actually used, kept next to the stats block (all-reduced with MAX).
struct StatsRows {
    uint32_t pre_rows;       // max raw length
    uint32_t post_rows;      // max over surviving reads of offset_5 + length
    uint32_t pre_len_size;   // size() of pre length histogram  (max len + 1, 0 if no read)
    uint32_t post_len_size;
};

// internal verdict flag (not part of the public FQ_RR_* set): terminal-N quality masking touched this read
constexpr uint32_t kFlagMasked = 0x80u;

// trim verdict packing: res.x = offset_5 | kResPlusBad (the one-character third line of the record is not "+": the record
// must not be block-copied), res.y = length | flags << 24
constexpr uint32_t kResPlusBad = 1u << 31;
constexpr uint32_t kResLenBits = 24;            // reads < 16 Mi bases
constexpr uint32_t kResLenMask = (1u << kResLenBits) - 1;

// One instruction asks the L2 for a whole byte range (Blackwell / Hopper bulk prefetch): base 16-byte aligned, size a multiple of 16.
__device__ __forceinline__ void prefetch_l2_bulk(const void *p, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

__host__ __device__ inline uint32_t pack_len_flags(uint32_t len, uint32_t flags) { return (len & kResLenMask) | (flags << kResLenBits); }

}  // namespace fq
