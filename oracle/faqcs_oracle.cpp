/*
 * faqcs_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A from-scratch CPU restatement of the FaQCs v2.10 per-read trim / filter /
 * statistics path, used ONLY as the checker for the CUDA implementation
 * (tests/, __graft_entry__.smoke(), bench.py cpu_baseline / --impl reference).
 * The product (faqcs_b200/csrc) never links, loads or calls this file.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py runs the unmodified
 * reference binary (oracle/_ref/FaQCs, compiled by oracle/Makefile from
 * /root/reference) on seeded inputs and compares every emitted FASTQ byte,
 * QC.stats.txt and the ten --debug matrix/histogram files with what this
 * restatement produces; committed fixtures of those runs live in tests/golden/.
 * The reference tree holds no known-answer tests of its own.
 *
 * Each function cites the reference file:line it restates.  Everything is
 * scalar and single threaded on purpose: it is the slow, obviously-literal
 * statement of the semantics.  Floating point expressions keep the reference's
 * C types (float vs double) so results are bit-identical on x86-64 SSE2;
 * build with -ffp-contract=off (oracle/Makefile).
 */
#include "faqcs_oracle.h"

#include <algorithm>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

struct OracleError {
    fq_status code;
    const char *msg;
};

// ---------------------------------------------------------------------------
// fastq.h:17-36  quality_score
// ---------------------------------------------------------------------------
inline int quality_score(char c, int offset)
{
    const char ret = (char)std::max(0, (int)c - offset);
    if (ret > FQ_MAX_QUALITY_SCORE) {
        throw OracleError{FQ_ERR_QUALITY,
            "fastq.h:quality_score: Found a quality score value that is greater than the maximum allowed quality score"};
    }
    return ret;
}

// ---------------------------------------------------------------------------
// trim.cpp:553-576  average_quality
// ---------------------------------------------------------------------------
float average_quality(const char *q, size_t len, int offset)
{
    int total = 0;
    for (size_t i = 0; i < len; ++i) total += q[i];   // plain (signed) char, trim.cpp:565
    if (len != 0) {
        // float / size_t -> float division; float(offset) subtraction (trim.cpp:572)
        return std::max(0.0f, float(total) / float(len) - float((char)offset));
    }
    return 0.0f;
}

// ---------------------------------------------------------------------------
// trim.cpp:629-672  hard_trim;  :675-709  BWA_trim;  :714-793  BWA_plus_trim
// Operate on q[0..len); return the new length, *f5 = bases cut from the 5' end.
// ---------------------------------------------------------------------------
uint32_t hard_trim(const char *q, int len, int Qopt, int off, bool protect_5, uint32_t *f5)
{
    int pos_3 = len - 1;
    int final_pos_5 = 0;
    int final_pos_3 = pos_3;
    while (pos_3 > 0) {
        if (Qopt < quality_score(q[pos_3], off)) { final_pos_3 = pos_3; break; }
        --pos_3;
    }
    if (!protect_5) {
        int pos_5 = final_pos_5;
        while (pos_5 < pos_3) {       // bounded by the loop variable, not final_pos_3 (Q1)
            if (Qopt < quality_score(q[pos_5], off)) { final_pos_5 = pos_5; break; }
            ++pos_5;
        }
    }
    *f5 = (uint32_t)final_pos_5;
    return (uint32_t)(final_pos_3 - final_pos_5 + 1);
}

uint32_t bwa_trim(const char *q, int len, int Q, int off, uint32_t *f5)
{
    int pos_3 = len - 1;
    int final_pos_3 = pos_3;
    int area = 0, maxArea = 0;
    while (pos_3 > 0 && area >= 0) {
        area += Q - quality_score(q[pos_3], off);
        if (area > maxArea) { maxArea = area; final_pos_3 = pos_3 - 1; }
        --pos_3;
    }
    *f5 = 0;
    return (uint32_t)(final_pos_3 + 1);
}

uint32_t bwa_plus_trim(const char *q, int len, int Q, int off, bool protect_5, uint32_t *f5)
{
    int at_least_scan = std::min(5, len);
    const int num_after_neg = std::min(2, len);
    int pos_3 = len - 1;
    int final_pos_5 = 0;
    int final_pos_3 = pos_3;
    int area = 0, maxArea = 0;
    while (at_least_scan) {
        --at_least_scan;
        if (pos_3 > num_after_neg && area >= 0) at_least_scan = num_after_neg;
        area += Q - quality_score(q[pos_3], off);
        if (area > maxArea) { maxArea = area; final_pos_3 = pos_3 - 1; }
        --pos_3;
    }
    if (!protect_5) {
        int pos_5 = 0;
        maxArea = 0;
        area = 0;
        at_least_scan = std::min(5, len);
        while (at_least_scan) {
            --at_least_scan;
            if (pos_5 < (final_pos_3 - num_after_neg) && area >= 0) at_least_scan = num_after_neg;
            area += Q - quality_score(q[pos_5], off);
            if (area > maxArea) { maxArea = area; final_pos_5 = pos_5 + 1; }
            ++pos_5;
        }
    }
    *f5 = (uint32_t)final_pos_5;
    if (final_pos_3 <= final_pos_5) return 0;      // 1-base survivors die too (Q14)
    return (uint32_t)(final_pos_3 - final_pos_5 + 1);
}

// ---------------------------------------------------------------------------
// seq_overlap.cpp:372-411  na_to_bits;  seq_overlap.h:133-150 bit codes
// ---------------------------------------------------------------------------
int na_to_bits(char c)
{
    enum { A = 1, C = 2, G = 4, T = 8, GAP = 16 };
    switch (c) {
        case 'A': case 'a': return A;
        case 'C': case 'c': return C;
        case 'G': case 'g': return G;
        case 'T': case 't': return T;
        case 'M': case 'm': return A | C;
        case 'R': case 'r': return G | A;
        case 'S': case 's': return G | C;
        case 'V': case 'v': return G | C | A;
        case 'W': case 'w': return A | T;
        case 'Y': case 'y': return T | C;
        case 'H': case 'h': return A | C | T;
        case 'K': case 'k': return G | T;
        case 'D': case 'd': return G | A | T;
        case 'B': case 'b': return G | T | C;
        case 'N': case 'n': return A | C | G | T;
        case '-': return GAP;
    }
    throw OracleError{FQ_ERR_BASE, "seq_overlap.cpp:na_to_bits: Unknown base!"};
}

// ---------------------------------------------------------------------------
// seq_overlap.cpp:46-370  SeqOverlap::align_smith_waterman, one SIMD lane.
// `lane` carries (start, stop) across calls exactly like max_elem.M_start_i /
// stop_i do in the reference (only max_elem.M is reset, seq_overlap.cpp:104).
// ---------------------------------------------------------------------------
struct Lane {
    int start = 0;   // max_elem.M_start_i.v[lane]
    int stop = 0;    // stop_i.v[lane]
    int score = 0;   // max_elem.M.v[lane]
};

struct Cell { short M; short start_i; };

void align_lane(const std::vector<uint8_t> &q, const std::vector<uint8_t> &t, Lane &lane,
                std::vector<Cell> &last_row, std::vector<Cell> &curr_row)
{
    const int L = (int)q.size();
    const int T = (int)t.size();
    last_row.assign(T + 1, Cell{0, 0});            // :86-101
    curr_row.assign(T + 1, Cell{0, 0});
    short max_M = 0;                               // :104 (start/stop are NOT reset)
    for (int i = 0; i < L; ++i) {                  // :106
        curr_row[0].M = 0;                         // :111
        curr_row[0].start_i = (short)(i + 1);      // :117
        for (int j = 0; j < T; ++j) {              // :147
            const Cell A = last_row[j];
            const short s = ((q[i] & t[j]) > 0) ? 1 : -1;             // :157-161
            Cell X;
            X.M = (short)(std::max<short>(A.M, 0) + s);               // :185-188
            X.start_i = (0 > A.M) ? (short)i : A.start_i;             // :255,272-275
            curr_row[j + 1] = X;
            if (!(X.M < max_M)) {                                     // :342 (>=; i<L and j<T hold)
                max_M = X.M;
                lane.start = X.start_i;
                lane.stop = i;
            }
        }
        std::swap(last_row, curr_row);             // :368
    }
    lane.score = max_M;
}

// ---------------------------------------------------------------------------
// trim.cpp:1144-1189  find_mask_range (with the run_length reset quirk, Q2)
// ---------------------------------------------------------------------------
void find_mask_range(const std::vector<uint8_t> &mask, uint32_t *start, uint32_t *length)
{
    const uint32_t len = (uint32_t)mask.size();
    uint32_t longest_run_start = 0, longest_run_length = 0, run_start = 0, run_length = 0;
    for (uint32_t i = 0; i < len; ++i) {
        if (!mask[i]) {
            if (run_length > longest_run_length) {
                longest_run_length = run_length;
                longest_run_start = run_start;
                run_length = 0;
            }
        } else {
            if (run_length == 0) run_start = i;
            ++run_length;
        }
    }
    if (run_length > longest_run_length) {
        longest_run_length = run_length;
        longest_run_start = run_start;
    }
    if (longest_run_length == 0) { *start = 0; *length = 0; return; }
    *start = longest_run_start;
    *length = longest_run_length;
}

// trim.cpp:969 + :1007-1008 / :1082
int32_t match_threshold(float rate, uint64_t n)
{
    const float filter_adapter_match_rate = (float)(1.0 - (double)rate);
    return (int32_t)(filter_adapter_match_rate * (float)n);
}

// trim.cpp:860-874
uint32_t composition_bin(uint32_t len, uint32_t count)
{
    const float norm = (len > 0) ? float(FQ_NUM_COMPOSITION_BIN - 1) / float(len) : 0.0f;
    return (uint32_t)(norm * float(count));
}

struct Read {
    std::string def, seq, qual;
    uint32_t sl_start = 0, sl_len = 0;   // Read::start_length (FaQCs.h:154)
    int16_t adapter = -1;
    bool adapter_hit = false;
};

struct Stats {
    uint64_t filter[FQ_NUM_STAT] = {0};
    std::vector<uint64_t> adapter_reads, adapter_bases;
    std::vector<uint64_t> pre_q, post_q;     // rows x 42
    std::vector<uint64_t> pre_b, post_b;     // rows x 5
    uint32_t pre_q_rows = 0, post_q_rows = 0, pre_b_rows = 0, post_b_rows = 0;
    uint64_t pre_rq[FQ_NUM_QUAL] = {0}, pre_bq[FQ_NUM_QUAL] = {0};
    uint64_t post_rq[FQ_NUM_QUAL] = {0}, post_bq[FQ_NUM_QUAL] = {0};
    std::vector<uint64_t> pre_comp, post_comp;   // 6 x 10001
    std::vector<uint64_t> pre_len, post_len;
    Stats() : pre_comp(FQ_NUM_COMPOSITION * FQ_NUM_COMPOSITION_BIN, 0),
              post_comp(FQ_NUM_COMPOSITION * FQ_NUM_COMPOSITION_BIN, 0) {}
};

} // namespace

struct fqo_ctx {
    fq_options opt;
    std::vector<std::string> adapter_names, adapter_seqs;
    std::vector<std::vector<uint8_t>> adapter_bits;
    Stats st;
    std::string error;
    bool debug_results = false;
    // batch outputs (owned here)
    std::string out[FQ_NUM_STREAM];
    std::vector<fq_read_result> results[2];
    // k-mer rarefaction (Options::kmer_rarefaction / kmer / split_size / num_subsample, PlotInfo::kmer_*)
    struct Kmer {
        bool enabled = false, collecting = false;         // collecting = m_opt.kmer_rarefaction (cleared at trim.cpp:180-184)
        uint32_t k = 31;
        uint64_t split_size = 1000000, num_subsample = 10;
        uint64_t total_number = 0;                        // FilterStat::TOTAL_NUMBER as trim() sees it after each call
        std::unordered_map<uint64_t, uint64_t> table;     // one per pass (FaQCs.cpp:235, 588)
        uint64_t table_total = 0;                         // sum of the table's counts
        std::vector<fq_rarefaction> samples;              // PlotInfo::kmer_rarefaction
        std::map<uint64_t, uint64_t> freq;                // PlotInfo::kmer_frequency_histogram
        std::vector<uint64_t> flat;
    } kmer;
};

namespace {

std::string g_create_error;

// ---------------------------------------------------------------------------
// fastq.cpp:8-125  next_read, restated over an in-memory buffer.  A "gzgets
// line" is the bytes up to and including the next '\n' (or to EOF); its
// content ends at the first '\n' or '\r' (strpbrk, fastq.cpp:44,70,100).
// ---------------------------------------------------------------------------
struct LineReader {
    const uint8_t *p;
    size_t n, pos = 0;
    // returns false at EOF (gzgets == NULL)
    bool gets(size_t *b, size_t *e, bool *has_eol)
    {
        if (pos >= n) return false;
        const uint8_t *nl = (const uint8_t *)memchr(p + pos, '\n', n - pos);
        const size_t end = nl ? (size_t)(nl - p) + 1 : n;
        *b = pos;
        size_t k = pos;
        while (k < end && p[k] != '\n' && p[k] != '\r') ++k;
        *has_eol = (k < end);
        *e = k;
        pos = end;
        return true;
    }
};

void append(std::string &s, const uint8_t *p, size_t b, size_t e) { s.append((const char *)p + b, e - b); }

bool next_read(LineReader &in, Read &r)
{
    size_t b, e;
    bool eol;
    r.def.clear();
    while (true) {
        if (!in.gets(&b, &e, &eol)) return false;                  // fastq.cpp:34-41 (gzeof)
        append(r.def, in.p, b, e);
        if (eol) break;
    }
    r.seq.clear();
    while (true) {
        if (!in.gets(&b, &e, &eol)) throw OracleError{FQ_ERR_FORMAT, "fastq.cpp:next_read: Unable to read sequence"};
        append(r.seq, in.p, b, e);
        if (eol) break;
    }
    if (!in.gets(&b, &e, &eol)) throw OracleError{FQ_ERR_FORMAT, "fastq.cpp:next_read: Unable to read '+'"};
    if (!eol) throw OracleError{FQ_ERR_FORMAT, "fastq.cpp:next_read: Error reading '+' delimiter"};
    r.qual.clear();
    while (true) {
        if (!in.gets(&b, &e, &eol)) throw OracleError{FQ_ERR_FORMAT, "fastq.cpp:next_read: Unable to read quality"};
        append(r.qual, in.p, b, e);
        if (eol) break;
    }
    if (r.seq.size() != r.qual.size()) throw OracleError{FQ_ERR_FORMAT, "fastq.cpp:next_read: |Sequence| != |Quality|"};
    return true;
}

void parse_all(const uint8_t *p, size_t n, std::vector<Read> &out)
{
    LineReader in{p, n};
    while (true) {
        Read r;
        if (!next_read(in, r)) break;
        out.push_back(std::move(r));
    }
}

// trim.cpp:188-222  parse_id
std::string parse_id(const std::string &def)
{
    size_t loc = def.find(' ');
    if (loc == std::string::npos) loc = def.size();
    if (loc > 1 && isdigit((unsigned char)def[loc - 1])) {
        if (def[loc - 2] == '.' || def[loc - 2] == '/') loc -= 2;
    }
    return def.substr(0, loc);
}

// ---------------------------------------------------------------------------
// Statistics updaters: trim.cpp:795-808, :810-875, :877-885
// ---------------------------------------------------------------------------
void update_quality_matrix(std::vector<uint64_t> &m, uint32_t &rows, const char *q, uint32_t len,
                           uint32_t offset_5, int qoff)
{
    const uint32_t full_len = len + offset_5;
    if (rows < full_len) { rows = full_len; m.resize((size_t)rows * FQ_NUM_QUAL, 0); }
    for (uint32_t i = 0; i < len; ++i) ++m[(size_t)(i + offset_5) * FQ_NUM_QUAL + quality_score(q[i], qoff)];
}

void update_base_statistics(std::vector<uint64_t> &m, uint32_t &rows, std::vector<uint64_t> &comp,
                            const char *s, uint32_t len, uint32_t offset_5)
{
    const uint32_t full_len = len + offset_5;
    if (rows < full_len) { rows = full_len; m.resize((size_t)rows * FQ_NUM_BASE, 0); }
    unsigned nA = 0, nT = 0, nC = 0, nG = 0, nN = 0;
    uint32_t index = offset_5;
    for (uint32_t i = 0; i < len; ++i, ++index) {
        switch (s[i]) {
            case 'A': case 'a': ++nA; ++m[(size_t)index * FQ_NUM_BASE + 0]; break;
            case 'T': case 't': ++nT; ++m[(size_t)index * FQ_NUM_BASE + 1]; break;
            case 'C': case 'c': ++nC; ++m[(size_t)index * FQ_NUM_BASE + 2]; break;
            case 'G': case 'g': ++nG; ++m[(size_t)index * FQ_NUM_BASE + 3]; break;
            case 'N': case 'n': ++nN; ++m[(size_t)index * FQ_NUM_BASE + 4]; break;
        }
    }
    const uint32_t iA = composition_bin(len, nA), iT = composition_bin(len, nT);
    const uint32_t iC = composition_bin(len, nC), iG = composition_bin(len, nG);
    const uint32_t iN = composition_bin(len, nN);
    const size_t B = FQ_NUM_COMPOSITION_BIN;
    ++comp[0 * B + iA];
    ++comp[1 * B + iT];
    ++comp[2 * B + iC];
    ++comp[3 * B + iG];
    ++comp[4 * B + iN];
    ++comp[5 * B + iG + iC];       // GC bin = sum of the truncated indices (trim.cpp:874)
}

void update_length_histogram(std::vector<uint64_t> &h, uint32_t len)
{
    if (h.size() <= len) h.resize((size_t)len + 1, 0);
    ++h[len];
}

// ---------------------------------------------------------------------------
// trim.cpp:225-551  trim_read
// On return r.seq / r.qual hold what write_read would emit (empty if invalid).
// ---------------------------------------------------------------------------
bool trim_read(fqo_ctx &c, Read &r, fq_read_result *res)
{
    const fq_options &o = c.opt;
    Stats &st = c.st;
    const int in_off = o.input_quality_offset;
    const int out_off = o.output_quality_offset;
    bool ret = true;
    uint32_t len = (uint32_t)r.seq.size();
    uint32_t offset_5 = 0;
    uint16_t flags = 0;

    ++st.filter[FQ_TOTAL_COUNT];
    ++st.filter[FQ_TOTAL_NUMBER];
    st.filter[FQ_TOTAL_LENGTH] += len;

    // mask_quality_terminal_N, trim.cpp:1191-1216 (uppercase 'N' only)
    for (uint32_t i = 0; i < len && r.seq[i] == 'N'; ++i) r.qual[i] = (char)in_off;
    for (uint32_t i = len; i > 0 && r.seq[i - 1] == 'N'; --i) r.qual[i - 1] = (char)in_off;

    update_quality_matrix(st.pre_q, st.pre_q_rows, r.qual.data(), len, 0, in_off);
    update_base_statistics(st.pre_b, st.pre_b_rows, st.pre_comp, r.seq.data(), len, 0);
    update_length_histogram(st.pre_len, len);
    int quality_bin = (int)average_quality(r.qual.data(), len, in_off);   // trim.cpp:254
    ++st.pre_rq[quality_bin];
    st.pre_bq[quality_bin] += len;

    // Window [lo, lo+len) into r.seq / r.qual instead of the reference's substr copies.
    uint32_t lo = 0;

    if (o.filter_adapter) {                                   // trim.cpp:270-277, :934-954
        const uint32_t full = len;
        if (full != r.sl_len) {
            lo = r.sl_start;
            len = r.sl_len;
            offset_5 += (r.sl_len == 0) ? full : r.sl_start;
        }
    }
    if (o.trim_5 && !o.qc_only) {                             // trim.cpp:279-297
        if (o.trim_5 > len) {
            len = 0;                                          // offset_5 += len (already 0), Q14
        } else {
            lo += o.trim_5;
            len -= o.trim_5;
            offset_5 += o.trim_5;
        }
    }
    if (o.trim_3 && !o.qc_only) {                             // trim.cpp:299-314
        if (o.trim_3 > len) len = 0;
        else len -= o.trim_3;
    }
    if (len < o.min_read_length || len == 0) {                // trim.cpp:317-323
        st.filter[FQ_BASE_LENGTH] += len;
        ++st.filter[FQ_READ_LENGTH];
        flags |= FQ_RR_F_LENGTH;
        ret = false;
    }
    if (!o.qc_only && ret) {                                  // trim.cpp:325-360
        const uint32_t init_len = len;
        uint32_t f5 = 0;
        const char *q = r.qual.data() + lo;
        const int Q = (int)(char)o.quality;
        switch (o.mode) {
            case FQ_MODE_HARD: len = hard_trim(q, (int)len, Q, in_off, o.protect_5 != 0, &f5); break;
            case FQ_MODE_BWA: len = bwa_trim(q, (int)len, Q, in_off, &f5); break;
            case FQ_MODE_BWA_PLUS: len = bwa_plus_trim(q, (int)len, Q, in_off, o.protect_5 != 0, &f5); break;
            default: throw OracleError{FQ_ERR_ARG, "trim.cpp:trim_read: Undefined trimming mode!"};
        }
        offset_5 += f5;
        lo += f5;
        if (init_len != len) {
            st.filter[FQ_BASE_QUAL_TRIM] += init_len - len;
            ++st.filter[FQ_READ_QUAL_TRIM];
            flags |= FQ_RR_QUAL_TRIMMED;
        }
        if (len < o.min_read_length || len == 0) {
            st.filter[FQ_BASE_LENGTH] += len;
            ++st.filter[FQ_READ_LENGTH];
            flags |= FQ_RR_F_LENGTH;
            ret = false;
        }
    }
    if (len == 0) lo = 0;
    char *seq = &r.seq[0] + lo;
    char *qual = &r.qual[0] + lo;

    if (ret) {                                                // trim.cpp:363-371, :578-597
        unsigned max_poly_n = 0, curr = 0;
        for (uint32_t i = 0; i < len; ++i) {
            if (seq[i] == 'N') { ++curr; max_poly_n = std::max(max_poly_n, curr); }
            else curr = 0;
        }
        if (max_poly_n >= o.max_num_poly_N) {
            st.filter[FQ_BASE_NN] += len;
            ++st.filter[FQ_READ_NN];
            flags |= FQ_RR_F_NN;
            if (!o.qc_only) ret = false;
        }
    }
    const float ave_Q = average_quality(qual, len, in_off);   // trim.cpp:374
    if (ret && ave_Q < o.average_quality) {
        st.filter[FQ_BASE_AVG_Q] += len;
        ++st.filter[FQ_READ_AVG_Q];
        flags |= FQ_RR_F_AVGQ;
        ret = false;
    }
    if (ret && len != 0) {                                    // trim.cpp:388-513
        if (o.replace_to_N_q > 0) {
            for (uint32_t i = 0; i < len; ++i) {
                if (seq[i] == 'G' && quality_score(qual[i], in_off) < (int)o.replace_to_N_q) seq[i] = 'N';
            }
        }
        unsigned nA = 0, nT = 0, nG = 0, nC = 0;
        unsigned dc[16] = {0};
        unsigned char last = 4;   // INVALID_BASE; A=0,T=1,C=2,G=3 (trim.cpp:419-424)
        for (uint32_t i = 0; i < len; ++i) {
            unsigned char cur;
            switch (seq[i]) {
                case 'A': case 'a': cur = 0; ++nA; break;
                case 'T': case 't': cur = 1; ++nT; break;
                case 'C': case 'c': cur = 2; ++nC; break;
                case 'G': case 'g': cur = 3; ++nG; break;
                default: cur = 4; break;
            }
            if (cur != 4 && cur != last && last != 4) ++dc[(last << 2) | cur];
            last = cur;
        }
        float norm = (float)(1.0 / (double)len);              // trim.cpp:483 (double division, narrowed)
        const float lc = o.low_complexity_cutoff_ratio;
        if (float(nA) * norm > lc || float(nT) * norm > lc || float(nG) * norm > lc || float(nC) * norm > lc) {
            st.filter[FQ_BASE_LOW_COMPLEXITY] += len;
            ++st.filter[FQ_READ_LOW_COMPLEXITY];
            flags |= FQ_RR_F_LOWCOMP;
            ret = false;
        } else {
            norm = (float)((double)norm * 2.0);               // trim.cpp:499
            for (int i = 0; i < 16; ++i) {
                if (float(dc[i]) * norm > lc) {
                    st.filter[FQ_BASE_LOW_COMPLEXITY] += len;
                    ++st.filter[FQ_READ_LOW_COMPLEXITY];
                    flags |= FQ_RR_F_LOWCOMP;
                    ret = false;
                    break;
                }
            }
        }
    }
    if (ret && in_off != out_off) {                           // trim.cpp:516-525
        for (uint32_t i = 0; i < len; ++i) {
            qual[i] = (char)(quality_score(qual[i], in_off) + out_off);
            if (qual[i] < 0) throw OracleError{FQ_ERR_QUALITY, "trim.cpp: quality error!"};
        }
    }
    if (ret) {                                                // trim.cpp:527-548
        st.filter[FQ_TOTAL_TRIMMED_LENGTH] += len;
        ++st.filter[FQ_TOTAL_TRIMMED_NUMBER];
        update_quality_matrix(st.post_q, st.post_q_rows, qual, len, offset_5, out_off);
        update_base_statistics(st.post_b, st.post_b_rows, st.post_comp, seq, len, offset_5);
        update_length_histogram(st.post_len, len);
        quality_bin = (int)ave_Q;
        ++st.post_rq[quality_bin];
        st.post_bq[quality_bin] += len;
        flags |= FQ_RR_VALID;
    }
    if (r.adapter_hit) flags |= FQ_RR_ADAPTER;
    if (res) {
        res->offset_5 = offset_5;
        res->length = ret ? len : 0;
        res->flags = flags;
        res->adapter = r.adapter;
        res->avg_q = ave_Q;
    }
    if (ret) {
        r.seq = r.seq.substr(lo, len);
        r.qual = r.qual.substr(lo, len);
    } else {
        r.seq.clear();                                        // trim.cpp:103-105
        r.qual.clear();
    }
    return ret;
}

// ---------------------------------------------------------------------------
// trim.cpp:961-1142  trim_adapters_and_phiX(vector<Read>&, ...)
// One call = one OpenMP thread's chunk [begin, end) of one trim() batch.
// emulate == true reproduces the 8-read grouping (Q3) and the per-lane stale
// alignment state (Q5); emulate == false treats every read on its own with
// threshold int(rate * min(own length, |adapter|)) and a zeroed stale range.
// ---------------------------------------------------------------------------
void adapter_pass_chunk(fqo_ctx &c, std::vector<Read> &reads, size_t begin, size_t end, bool emulate)
{
    const fq_options &o = c.opt;
    const size_t n_adapter = c.adapter_bits.size();
    std::vector<Cell> row_a, row_b;
    Lane lanes[8];
    std::vector<uint8_t> qbits[8];
    std::vector<uint8_t> mask[8];
    size_t slot_read[8];
    unsigned current_slot = 0;

    auto finish_group = [&](unsigned n_slot, bool full_group, size_t thr_read_len) {
        short best_score[8] = {0};
        unsigned best_adapter[8] = {0};
        for (size_t j = 0; j < n_adapter; ++j) {
            const size_t T = c.adapter_bits[j].size();
            for (unsigned slot = 0; slot < n_slot; ++slot) {
                size_t thr_len;
                if (!emulate) thr_len = std::min(qbits[slot].size(), T);
                else thr_len = full_group ? std::min(thr_read_len, T) : T;   // :1007-1008 vs :1082
                const int threshold = match_threshold(o.adapter_mismatch_rate, thr_len);
                align_lane(qbits[slot], c.adapter_bits[j], lanes[slot], row_a, row_b);
                const int score = lanes[slot].score;
                const int match_length = lanes[slot].stop - lanes[slot].start + 1;
                const int num_match = (match_length + score) / 2;             // :1024-1025
                if (num_match >= threshold) {
                    for (int k = lanes[slot].start; k <= lanes[slot].stop; ++k) {
                        if (k >= 0 && (size_t)k < mask[slot].size()) mask[slot][k] = 0;   // reference would be UB outside
                    }
                    if (score > best_score[slot]) { best_score[slot] = (short)score; best_adapter[slot] = (unsigned)j; }
                }
            }
        }
        for (unsigned slot = 0; slot < n_slot; ++slot) {
            if (best_score[slot] > 0) {                                       // :1048
                Read &r = reads[slot_read[slot]];
                uint32_t s, l;
                find_mask_range(mask[slot], &s, &l);
                r.sl_start = s;
                r.sl_len = l;
                r.adapter = (int16_t)best_adapter[slot];
                r.adapter_hit = true;
                ++c.st.adapter_reads[best_adapter[slot]];
                c.st.adapter_bases[best_adapter[slot]] += mask[slot].size() - l;   // :1064
            }
        }
    };

    for (size_t i = begin; i < end; ++i) {
        Read &r = reads[i];
        const size_t read_len = r.seq.size();
        r.sl_start = 0;
        r.sl_len = (uint32_t)read_len;
        r.adapter = -1;
        r.adapter_hit = false;
        mask[current_slot].assign(read_len, 1);
        qbits[current_slot].resize(read_len);
        for (size_t k = 0; k < read_len; ++k) qbits[current_slot][k] = (uint8_t)na_to_bits(r.seq[k]);
        slot_read[current_slot] = i;
        if (!emulate) {
            lanes[0] = Lane();
            finish_group(1, false, 0);
            continue;
        }
        ++current_slot;
        if (current_slot == 8) {
            finish_group(8, true, read_len);
            current_slot = 0;
        }
    }
    if (emulate && current_slot > 0) finish_group(current_slot, false, 0);
}

// trim.cpp:67-186  trim(): adapter pass + per-read pass for one mate's reads.
// With num_thread >= 1 the reads are cut into the reference's 32768-read
// batches (FaQCs.cpp:232) and each batch into libgomp static chunks.
void trim_mate(fqo_ctx &c, std::vector<Read> &reads, uint64_t first_record_index, std::vector<fq_read_result> *results)
{
    const fq_options &o = c.opt;
    const size_t n = reads.size();
    if (results) results->assign(n, fq_read_result{0, 0, 0, -1, 0.0f});
    if (o.filter_adapter) {
        if (o.num_thread == 0) {
            adapter_pass_chunk(c, reads, 0, n, false);
        } else {
            const size_t nt = o.num_thread;
            size_t pos = 0;
            while (pos < n) {
                const uint64_t g = first_record_index + pos;
                const size_t in_batch = (size_t)(g % FQ_REF_BATCH);
                const size_t N = std::min<size_t>(FQ_REF_BATCH - in_batch, n - pos);
                // libgomp static schedule: thread t gets q (+1 if t < r) consecutive iterations
                const size_t q = N / nt, rr = N % nt;
                size_t s = 0;
                for (size_t t = 0; t < nt; ++t) {
                    const size_t sz = q + (t < rr ? 1 : 0);
                    if (sz) adapter_pass_chunk(c, reads, pos + s, pos + s + sz, true);
                    s += sz;
                }
                pos += N;
            }
        }
    }
    for (size_t i = 0; i < n; ++i) trim_read(c, reads[i], results ? &(*results)[i] : nullptr);
}

// trim.cpp:887-931 update_kmer: canonical k-mers of one sequence, two bits per base (FaQCs.h BASE_A=0, T=1, C=2, G=3 --
// pinned by the reference's .kmerH.txt / .Kmercount.txt only through the min() of the two strands)
void update_kmer(fqo_ctx::Kmer &K, const std::string &seq)
{
    const uint64_t comp_shift = 2 * (K.k - 1), mask = (1ull << (2 * K.k)) - 1;
    uint64_t w = 0, comp = 0;
    uint32_t word_len = 0;
    for (char ch : seq) {
        ++word_len;
        switch (ch) {
            case 'A': case 'a': w = (w << 2) | 0; comp = (comp >> 2) | (1ull << comp_shift); break;
            case 'T': case 't': w = (w << 2) | 1; comp = (comp >> 2) | (0ull << comp_shift); break;
            case 'G': case 'g': w = (w << 2) | 3; comp = (comp >> 2) | (2ull << comp_shift); break;
            case 'C': case 'c': w = (w << 2) | 2; comp = (comp >> 2) | (3ull << comp_shift); break;
            default: word_len = 0; break;
        }
        if (word_len >= K.k) {
            ++K.table[std::min(w & mask, comp & mask)];
            ++K.table_total;
        }
    }
}

// The k-mer side of the trim() calls a batch stands for: FaQCs.cpp reads 32768 records, calls trim() on mate 1, then on
// mate 2 (FaQCs.cpp:287-291, 424-428, 628-629, 692-693).  `raw` = the reads as parsed, `done` = after trim_read.
void kmer_calls(fqo_ctx &c, const std::vector<Read> *raw, const std::vector<Read> *done, int n_mates)
{
    fqo_ctx::Kmer &K = c.kmer;
    const size_t n = raw[0].size();
    for (size_t b0 = 0; b0 < n; b0 += FQ_REF_BATCH) {
        const size_t b1 = std::min<size_t>(n, b0 + FQ_REF_BATCH);
        for (int m = 0; m < n_mates; ++m) {
            if (K.collecting) {
                for (size_t i = b0; i < b1; ++i) {
                    if (c.opt.qc_only) update_kmer(K, raw[m][i].seq);                  // trim.cpp:260-262 (before any trimming)
                    else if (!done[m][i].seq.empty()) update_kmer(K, done[m][i].seq);  // trim.cpp:527, 545-547 (survivors, trimmed)
                }
            }
            K.total_number += b1 - b0;
            if (K.collecting) {                                                        // trim.cpp:157-185
                const uint64_t index = K.total_number / K.split_size;
                const uint64_t num_rarefaction = K.samples.size();
                if (index > num_rarefaction && num_rarefaction < K.num_subsample)
                    K.samples.push_back(fq_rarefaction{K.total_number, (uint64_t)K.table.size(), K.table_total});
                if (num_rarefaction >= K.num_subsample) K.collecting = false;
            }
        }
    }
}

void write_read(std::string &out, const std::string &def, const std::string &seq, const std::string &qual)
{   // fastq.cpp:127-138
    out += def; out += '\n'; out += seq; out += "\n+\n"; out += qual; out += '\n';
}

fq_status fail(fqo_ctx *c, fq_status code, const char *msg)
{
    if (c) c->error = msg; else g_create_error = msg;
    return code;
}

} // namespace

extern "C" {

fq_status fqo_create(const fq_options *opt, fqo_ctx **out)
{
    if (!opt || !out) return fail(nullptr, FQ_ERR_ARG, "fqo_create: null argument");
    fqo_ctx *c = new fqo_ctx();
    c->opt = *opt;
    try {
        for (uint32_t i = 0; i < opt->n_adapters; ++i) {
            c->adapter_names.push_back(opt->adapters[i].name ? opt->adapters[i].name : "");
            c->adapter_seqs.push_back(opt->adapters[i].seq ? opt->adapters[i].seq : "");
            std::vector<uint8_t> bits;
            for (char ch : c->adapter_seqs.back()) bits.push_back((uint8_t)na_to_bits(ch));
            c->adapter_bits.push_back(bits);
        }
    } catch (const OracleError &e) {
        delete c;
        return fail(nullptr, e.code, e.msg);
    }
    c->opt.adapters = nullptr;
    c->st.adapter_reads.assign(opt->n_adapters, 0);
    c->st.adapter_bases.assign(opt->n_adapters, 0);
    *out = c;
    return FQ_OK;
}

void fqo_destroy(fqo_ctx *ctx) { delete ctx; }

const char *fqo_last_error(const fqo_ctx *ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

fq_status fqo_set_debug_results(fqo_ctx *ctx, int enable)
{
    if (!ctx) return FQ_ERR_ARG;
    ctx->debug_results = enable != 0;
    return FQ_OK;
}

// the drivers' "looks like NextSeq data" adjustment between two trim() calls (FaQCs.cpp:272-277, 406-411, 613-618, 675-680)
fq_status fqo_set_quality(fqo_ctx *ctx, int32_t quality)
{
    if (!ctx) return FQ_ERR_ARG;
    ctx->opt.quality = quality;
    return FQ_OK;
}

// trim.cpp:599-617 + :619-626 and their call sites FaQCs.cpp:261-277,393-414,609-619,669-683
fq_status fqo_autodetect(fqo_ctx *ctx, const uint8_t *r1, size_t n1, const uint8_t *r2, size_t n2,
                         int32_t *input_quality_offset, int32_t *quality)
{
    if (!ctx || (!r1 && n1)) return FQ_ERR_ARG;
    try {
        std::vector<Read> b1, b2;
        parse_all(r1, n1, b1);
        if (r2) parse_all(r2, n2, b2);
        if (b1.size() > FQ_REF_BATCH) b1.resize(FQ_REF_BATCH);
        if (b2.size() > FQ_REF_BATCH) b2.resize(FQ_REF_BATCH);
        auto detect = [](const std::vector<Read> &b) -> int {
            for (const Read &r : b)
                for (char ch : r.qual) {
                    if (ch > 74) return 64;
                    if (ch < 59) return 33;
                }
            throw OracleError{FQ_ERR_OFFSET, "trim.cpp:auto_detect_quality_offset: Unknown quality format!"};
        };
        if (ctx->opt.input_quality_offset == FQ_OFFSET_AUTO) {
            const int o1 = detect(b1);
            if (r2 && o1 != detect(b2))
                throw OracleError{FQ_ERR_OFFSET, "FaQCs.cpp:process_paired: I/O Error"};
            ctx->opt.input_quality_offset = o1;
        }
        if (ctx->opt.quality < 20 && !b1.empty() && b1[0].def.find("@NS") == 0) ctx->opt.quality = 20;
    } catch (const OracleError &e) {
        return fail(ctx, e.code, e.msg);
    }
    if (input_quality_offset) *input_quality_offset = ctx->opt.input_quality_offset;
    if (quality) *quality = ctx->opt.quality;
    return FQ_OK;
}

fq_status fqo_process_host(fqo_ctx *ctx, const uint8_t *r1, size_t n1, const uint8_t *r2, size_t n2,
                           uint64_t first_record_index, int is_final, fq_batch_out *out)
{
    (void)is_final;
    if (!ctx || !out || (!r1 && n1)) return FQ_ERR_ARG;
    if (ctx->opt.input_quality_offset == FQ_OFFSET_AUTO)
        return fail(ctx, FQ_ERR_STATE, "fqo_process_host: quality offset not set; call fqo_autodetect first");
    const bool paired = r2 != nullptr;
    for (auto &s : ctx->out) s.clear();
    memset(out, 0, sizeof(*out));
    try {
        std::vector<Read> b1, b2, raw1, raw2;
        parse_all(r1, n1, b1);
        if (paired) {
            parse_all(r2, n2, b2);
            if (b1.size() != b2.size())
                throw OracleError{FQ_ERR_FORMAT, "FaQCs.cppI/O error"};   // FaQCs.cpp:370-380 (sic)
            for (size_t i = 0; i < b1.size(); ++i)
                if (parse_id(b1[i].def) != parse_id(b2[i].def))
                    throw OracleError{FQ_ERR_FORMAT, "FaQCs.cpp:trim: I/O error"};   // FaQCs.cpp:383-389
        }
        const fq_options &o = ctx->opt;
        if (o.discard_output || ctx->kmer.enabled) { raw1 = b1; raw2 = b2; }       // FaQCs.cpp:279-285
        trim_mate(*ctx, b1, first_record_index, ctx->debug_results ? &ctx->results[0] : nullptr);
        if (paired) trim_mate(*ctx, b2, first_record_index, ctx->debug_results ? &ctx->results[1] : nullptr);
        if (ctx->kmer.enabled) {
            if (first_record_index % FQ_REF_BATCH) throw OracleError{FQ_ERR_ARG, "k-mer rarefaction: batch does not start on a 32768-record boundary"};
            const std::vector<Read> raws[2] = {raw1, raw2}, dones[2] = {b1, b2};
            kmer_calls(*ctx, raws, dones, paired ? 2 : 1);
        }
        const size_t n = b1.size();
        out->n_records = n;
        for (size_t i = 0; i < n; ++i) {
            const bool v1 = !b1[i].seq.empty();               // Read::valid, FaQCs.h:161-164
            if (!paired) {                                    // FaQCs.cpp:634-659
                if (v1) ++out->n_valid[0];
                if (o.qc_only) continue;
                if (v1) write_read(ctx->out[FQ_OUT_UNPAIRED], b1[i].def, b1[i].seq, b1[i].qual);
                else if (o.discard_output) write_read(ctx->out[FQ_OUT_DISCARD], raw1[i].def, raw1[i].seq, raw1[i].qual);
                continue;
            }
            const bool v2 = !b2[i].seq.empty();
            if (v1) ++out->n_valid[0];
            if (v2) ++out->n_valid[1];
            if (v1 && v2) {                                   // FaQCs.cpp:304-308
                out->paired_read_number += 2;
                out->paired_base_length += b1[i].seq.size() + b2[i].seq.size();
            }
            if (o.qc_only) continue;
            if (v1 && v2) {
                write_read(ctx->out[FQ_OUT_R1], b1[i].def, b1[i].seq, b1[i].qual);
                write_read(ctx->out[FQ_OUT_R2], b2[i].def, b2[i].seq, b2[i].qual);
            } else {
                if (v1) write_read(ctx->out[FQ_OUT_UNPAIRED], b1[i].def, b1[i].seq, b1[i].qual);
                else if (v2) write_read(ctx->out[FQ_OUT_UNPAIRED], b2[i].def, b2[i].seq, b2[i].qual);
                if (o.discard_output) {
                    if (!v1) write_read(ctx->out[FQ_OUT_DISCARD], raw1[i].def, raw1[i].seq, raw1[i].qual);
                    if (!v2) write_read(ctx->out[FQ_OUT_DISCARD], raw2[i].def, raw2[i].seq, raw2[i].qual);
                }
            }
        }
        ctx->st.filter[FQ_PAIRED_READ_NUMBER] += out->paired_read_number;
        ctx->st.filter[FQ_PAIRED_BASE_LENGTH] += out->paired_base_length;
    } catch (const OracleError &e) {
        return fail(ctx, e.code, e.msg);
    }
    for (int s = 0; s < FQ_NUM_STREAM; ++s) {
        out->bytes[s] = ctx->out[s].size();
        out->data[s] = ctx->out[s].empty() ? nullptr : (const uint8_t *)ctx->out[s].data();
    }
    if (ctx->debug_results) {
        out->results[0] = ctx->results[0].data();
        out->results[1] = r2 ? ctx->results[1].data() : nullptr;
    }
    return FQ_OK;
}

fq_status fqo_kmer_enable(fqo_ctx *ctx, uint32_t k, uint64_t split_size, uint32_t num_subsample)
{
    if (!ctx || k < 2 || k > 31 || !split_size || !num_subsample) return FQ_ERR_ARG;
    ctx->kmer.enabled = ctx->kmer.collecting = true;
    ctx->kmer.k = k; ctx->kmer.split_size = split_size; ctx->kmer.num_subsample = num_subsample;
    return FQ_OK;
}

fq_status fqo_kmer_end_pass(fqo_ctx *ctx)
{   // FaQCs.cpp:518-537, 737-756
    if (!ctx) return FQ_ERR_ARG;
    fqo_ctx::Kmer &K = ctx->kmer;
    if (!K.enabled) return FQ_OK;
    for (const auto &kv : K.table) ++K.freq[kv.second];
    if (K.collecting && K.samples.empty()) K.samples.push_back(fq_rarefaction{K.total_number, (uint64_t)K.table.size(), K.table_total});
    K.table.clear();
    K.table_total = 0;
    return FQ_OK;
}

fq_status fqo_kmer_results(fqo_ctx *ctx, fq_kmer_view *view)
{
    if (!ctx || !view) return FQ_ERR_ARG;
    fqo_ctx::Kmer &K = ctx->kmer;
    K.flat.clear();
    for (const auto &kv : K.freq) { K.flat.push_back(kv.first); K.flat.push_back(kv.second); }
    view->n_rarefaction = (uint32_t)K.samples.size();
    view->rarefaction = K.samples.data();
    view->n_frequency = K.flat.size() / 2;
    view->frequency = K.flat.data();
    return FQ_OK;
}

fq_status fqo_stats(fqo_ctx *ctx, fq_stats_view *v)
{
    if (!ctx || !v) return FQ_ERR_ARG;
    Stats &s = ctx->st;
    memset(v, 0, sizeof(*v));
    memcpy(v->filter_stats, s.filter, sizeof(s.filter));
    // main(): phiX pseudo-adapters -> READ/BASE_PHIX, the others -> READ/BASE_ADAPTER (FaQCs.cpp:92-127)
    for (size_t j = 0; j < s.adapter_reads.size(); ++j) {
        const bool phix = ctx->adapter_names[j] == "__PhiX174_NC_001422__" ||
                          ctx->adapter_names[j] == "__PhiX174_NC_001422_complement__";
        v->filter_stats[phix ? FQ_READ_PHIX : FQ_READ_ADAPTER] += s.adapter_reads[j];
        v->filter_stats[phix ? FQ_BASE_PHIX : FQ_BASE_ADAPTER] += s.adapter_bases[j];
    }
    v->n_adapters = (uint32_t)s.adapter_reads.size();
    v->adapter_reads = s.adapter_reads.data();
    v->adapter_bases = s.adapter_bases.data();
    // quality and base matrices grow together in the reference (same len + offset_5)
    v->pre_rows = std::max(s.pre_q_rows, s.pre_b_rows);
    v->post_rows = std::max(s.post_q_rows, s.post_b_rows);
    s.pre_q.resize((size_t)v->pre_rows * FQ_NUM_QUAL, 0);
    s.pre_b.resize((size_t)v->pre_rows * FQ_NUM_BASE, 0);
    s.post_q.resize((size_t)v->post_rows * FQ_NUM_QUAL, 0);
    s.post_b.resize((size_t)v->post_rows * FQ_NUM_BASE, 0);
    v->pre_len_size = (uint32_t)s.pre_len.size();
    v->post_len_size = (uint32_t)s.post_len.size();
    v->pre_quality_matrix = s.pre_q.data();
    v->post_quality_matrix = s.post_q.data();
    v->pre_base_matrix = s.pre_b.data();
    v->post_base_matrix = s.post_b.data();
    v->pre_read_quality_hist = s.pre_rq;
    v->pre_base_quality_hist = s.pre_bq;
    v->post_read_quality_hist = s.post_rq;
    v->post_base_quality_hist = s.post_bq;
    v->pre_composition = s.pre_comp.data();
    v->post_composition = s.post_comp.data();
    v->pre_length_hist = s.pre_len.data();
    v->post_length_hist = s.post_len.data();
    return FQ_OK;
}

uint32_t fqo_quality_trim(int mode, int quality, int in_offset, int protect_5, const char *qual, uint32_t len, uint32_t *f5)
{
    uint32_t cut = 0, out = 0;
    try {
        if (mode == FQ_MODE_HARD) out = hard_trim(qual, (int)len, quality, in_offset, protect_5 != 0, &cut);
        else if (mode == FQ_MODE_BWA) out = bwa_trim(qual, (int)len, quality, in_offset, &cut);
        else out = bwa_plus_trim(qual, (int)len, quality, in_offset, protect_5 != 0, &cut);
    } catch (const OracleError &) {
        out = 0xffffffffu;
    }
    if (f5) *f5 = cut;
    return out;
}

int fqo_align(const char *read, uint32_t read_len, const char *target, uint32_t target_len,
              int32_t *score, int32_t *start, int32_t *stop)
{
    std::vector<uint8_t> q(read_len), t(target_len);
    try {
        for (uint32_t i = 0; i < read_len; ++i) q[i] = (uint8_t)na_to_bits(read[i]);
        for (uint32_t i = 0; i < target_len; ++i) t[i] = (uint8_t)na_to_bits(target[i]);
    } catch (const OracleError &) {
        return -1;
    }
    Lane lane;
    lane.start = start ? *start : 0;     // caller may seed the stale state
    lane.stop = stop ? *stop : 0;
    std::vector<Cell> a, b;
    align_lane(q, t, lane, a, b);
    if (score) *score = lane.score;
    if (start) *start = lane.start;
    if (stop) *stop = lane.stop;
    return lane.score > 0 ? 0 : 1;
}

void fqo_find_mask_range(const uint8_t *mask, uint32_t len, uint32_t *start, uint32_t *length)
{
    std::vector<uint8_t> m(mask, mask + len);
    find_mask_range(m, start, length);
}

int32_t fqo_match_threshold(float rate, uint64_t n) { return match_threshold(rate, n); }
uint32_t fqo_composition_bin(uint32_t len, uint32_t count) { return composition_bin(len, count); }
float fqo_average_quality(const char *qual, uint32_t len, int offset) { return average_quality(qual, len, offset); }

} // extern "C"
