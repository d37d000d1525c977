// fq_frame.cuh -- FASTQ record framing on the device.
//
// Replaces the line scanning half of next_read (fastq.cpp:32-122): the byte scan for '\n' (gzgets) and '\r'
// (strpbrk), four lines per record, |seq| == |qual|.  Byte work: every resident warp streams its own contiguous
// segment of the input 4 KiB at a time, each lane scanning 128 contiguous bytes with exact SWAR zero-byte masks;
// one warp scan per chunk ranks the newlines, and the result is a segmented line index that k_build_records turns
// into record descriptors without touching the raw bytes again.
#pragma once
#include "fq_common.cuh"

namespace fq {

constexpr uint32_t kChunkBytes = 4096;                 // bytes per warp-chunk
constexpr uint32_t kVecPerChunk = kChunkBytes / 16;    // 256 uint4 per chunk -> 8 per lane

__device__ __forceinline__ uint4 ld_stream16(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// 16-bit mask of bytes equal to c in a 16-byte vector (bit i = byte i).
__device__ __forceinline__ uint32_t match16(const uint4 &v, uint32_t c4)
{
    // __vcmpeq4 -> 0xff per equal byte; gather the top bit of each byte.
    auto m4 = [&](uint32_t w) -> uint32_t {
        uint32_t e = __vcmpeq4(w, c4) & 0x80808080u;          // bit 7,15,23,31
        return ((e >> 7) | (e >> 14) | (e >> 21) | (e >> 28)) & 0xfu;
    };
    return m4(v.x) | (m4(v.y) << 4) | (m4(v.z) << 8) | (m4(v.w) << 12);
}

// ---------------------------------------------------------------------------------------------
// Single-pass line index: every '\n' offset recorded in order, one read of the input.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kFrameThreads = 256;
constexpr uint32_t kFrameTile = (kFrameThreads / 32) * kChunkBytes;   // 32 KiB
// A line-index entry is the byte offset of a '\n' plus two facts about the byte before it, so
// that record building never has to touch the raw bytes again (a batch is < 1 GiB per mate).
constexpr uint32_t kNlCr = 1u << 31;      // preceded by '\r'  (CRLF line end)
constexpr uint32_t kNlPosMask = (1u << 30) - 1;

// Exact zero-byte mask: 0x80 in every byte of x that is zero (no cross-byte borrows).
__device__ __forceinline__ uint32_t zero_bytes(uint32_t x)
{
    const uint32_t t = (x & 0x7f7f7f7fu) + 0x7f7f7f7fu;
    return ~(t | x | 0x7f7f7f7fu);
}
// 4-bit mask (bit b = byte b) of the bytes of w equal to the byte replicated in c4.
__device__ __forceinline__ uint32_t eq_nibble(uint32_t w, uint32_t c4)
{
    // 0x80 flags sit at bits 7,15,23,31; times (1 + 2^7 + 2^14 + 2^21) they meet in bits 28..31: the ten partial
    // products land on ten different bits, so nothing carries
    return (zero_bytes(w ^ c4) * 0x00204081u) >> 28;
}
__device__ __forceinline__ uint32_t eq_mask16(const uint4 &v, uint32_t c4)
{
    return eq_nibble(v.x, c4) | (eq_nibble(v.y, c4) << 4) | (eq_nibble(v.z, c4) << 8) | (eq_nibble(v.w, c4) << 12);
}
__device__ __forceinline__ uint32_t has_byte(uint32_t w, uint32_t c4)      // non-zero iff some byte of w equals c (may over-flag bytes, never misses)
{
    const uint32_t x = w ^ c4;
    return (x - 0x01010101u) & ~x & 0x80808080u;
}

// 16-bit mask (bit i = byte i), shifted left by 7, of the bytes of a 16-byte vector whose value lies in 0x08..0x0f ('\n' and
// its class): per word one exact zero-byte test on the value with the low three bits cleared (three instructions: the two
// constants sit in registers so that (w & c1) ^ c2 is one LOP3) and one dot product that drops the four 0x80 flags into
// consecutive bits.
__device__ __forceinline__ uint32_t class_flags(uint32_t w, uint32_t c_f8, uint32_t c_08)     // 0x80 in every byte of w that lies in 0x08..0x0f
{
    uint32_t x;
    asm("lop3.b32 %0, %1, %2, %3, 0x6a;" : "=r"(x) : "r"(w), "r"(c_f8), "r"(c_08));           // (w & c_f8) ^ c_08
    return (x - 0x01010101u) & ~x & 0x80808080u;
}
__device__ __forceinline__ uint32_t class_mask16_shl7(const uint4 &v, uint32_t c_f8, uint32_t c_08)
{
    const uint32_t lo = __dp4a(class_flags(v.y, c_f8, c_08), 0x80402010u, __dp4a(class_flags(v.x, c_f8, c_08), 0x08040201u, 0u));
    const uint32_t hi = __dp4a(class_flags(v.w, c_f8, c_08), 0x80402010u, __dp4a(class_flags(v.z, c_f8, c_08), 0x08040201u, 0u));
    return lo + (hi << 8);
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t x, uint32_t lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= (uint32_t)o) x += y;
    }
    return x;
}

// Line index, one pass, no cross-warp dependency.  Warp w owns the contiguous byte segment
// [w * seg_bytes, (w+1) * seg_bytes) and streams it 4 KiB at a time (coalesced 16-byte loads);
// it writes the offsets of its '\n' bytes, in order, to its own region nl_seg[w * seg_cap ...]
// and its line count to seg_count[w].  A tiny scan (k_scan_segments) then turns the counts into
// global line bases and k_build_records maps global line numbers to (segment, local index).
//
// Two instances.  FAST (the first one launched): one exact 3-instruction SWAR test per 32-bit word for the byte class
// 0x08..0x0f -- with the low three bits masked off, a byte of the word is zero exactly for that class and the classic
// (x - 0x01..) & ~x & 0x80.. test has no false positives because a masked byte is never 0x01 -- then one dot-product
// instruction per word turns the four flags into address-ordered mask bits.  Plain FASTQ text holds no other member of that
// class than '\n', so every flagged byte is checked to BE '\n' when its index entry is written; the first one that is not
// (a '\r' of a CRLF file, a tab) raises info->frame_exact[mate] and the EXACT instance, launched right behind, redoes the
// mate with separate masks for '\n' and '\r' (it returns at once when the flag is not set).
#ifndef FQ_FRAME_MIN_CTAS
#define FQ_FRAME_MIN_CTAS 4
#endif
#ifndef FQ_FRAME_PREFETCH
#define FQ_FRAME_PREFETCH 1
#endif
template <bool FAST>
__global__ void __launch_bounds__(kFrameThreads, FQ_FRAME_MIN_CTAS) k_frame_lines(const uint8_t *__restrict__ raw, uint64_t n, uint32_t seg_bytes, uint32_t n_seg,
                                                               uint32_t *__restrict__ nl_seg, uint32_t seg_cap, uint32_t *__restrict__ seg_count,
                                                               BatchInfo *info, int mate, int only_if_flagged)
{
    if (!FAST && only_if_flagged && *reinterpret_cast<volatile uint32_t *>(&info->frame_exact[mate]) == 0) return;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n_seg) return;
    const uint64_t seg_lo = (uint64_t)w * seg_bytes;
    const uint64_t seg_hi = min((uint64_t)n, seg_lo + seg_bytes);
    uint32_t *out = nl_seg + (size_t)w * seg_cap;
    uint32_t rank0 = 0, any_cr = 0, cr_eol = 0, not_nl = 0;
    const bool aligned32 = (reinterpret_cast<uintptr_t>(raw) & 31u) == 0;
    bool cr_before = true;        // the chunk in front of the segment belongs to another warp: assume it may end in CR
    for (uint64_t chunk_base = seg_lo; chunk_base < seg_hi; chunk_base += kChunkBytes) {
        // lane l owns the 128 contiguous bytes chunk_base + l*128 ..: its eight 16-byte vectors are one cache line,
        // its newline mask is 128 contiguous bits, and ranks follow from ONE warp scan of the per-lane counts
        uint32_t m16[4] = {0, 0, 0, 0};                 // two 16-bit newline masks per register, memory order
        uint32_t cr_chunk = 0;                          // non-zero: this lane's bytes may hold a CR
        const uint32_t lane_base = (uint32_t)chunk_base + lane * 128;         // a batch is < 1 GiB per mate
        uint4 v[8];
#if FQ_FRAME_PREFETCH
        // one instruction asks the L2 for the chunk this warp reads FQ_FRAME_PREFETCH iterations from now
        if (lane == 0 && aligned32 && chunk_base + (uint64_t)(FQ_FRAME_PREFETCH + 1) * kChunkBytes <= seg_hi)
            prefetch_l2_bulk(raw + chunk_base + (uint64_t)FQ_FRAME_PREFETCH * kChunkBytes, kChunkBytes);
#endif
        if (chunk_base + kChunkBytes <= seg_hi && aligned32) {
            // whole chunk inside the segment (warp-uniform): four 256-bit loads per lane (sm_100 LDG.256): each touches one
            // 32-byte sector once -- with 128-bit loads every sector is requested twice and the kernel is L1TEX-bound
#pragma unroll
            for (int k = 0; k < 4; ++k)
                asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=r"(v[2 * k].x), "=r"(v[2 * k].y), "=r"(v[2 * k].z), "=r"(v[2 * k].w), "=r"(v[2 * k + 1].x), "=r"(v[2 * k + 1].y),
                               "=r"(v[2 * k + 1].z), "=r"(v[2 * k + 1].w)
                             : "l"(raw + lane_base + 32 * k));
        } else if (chunk_base + kChunkBytes <= seg_hi) {
            const uint4 *src = reinterpret_cast<const uint4 *>(raw + lane_base);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = __ldg(src + k);
        } else {
            // bytes past the end of the segment read as 0xff: no member of any class the masks look for
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint64_t off = (uint64_t)lane_base + (uint64_t)k * 16;
                v[k] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
                if (off + 16 <= seg_hi) v[k] = __ldg(reinterpret_cast<const uint4 *>(raw + off));
                else if (off < seg_hi) {
                    uint32_t x[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
                    for (uint32_t b = 0; b < 16 && off + b < seg_hi; ++b) x[b >> 2] = (x[b >> 2] & ~(0xffu << (8 * (b & 3)))) | ((uint32_t)raw[off + b] << (8 * (b & 3)));
                    v[k] = make_uint4(x[0], x[1], x[2], x[3]);
                }
            }
        }
        if (FAST) {
            uint32_t c_f8 = 0xf8f8f8f8u, c_08 = 0x08080808u;
            asm volatile("" : "+r"(c_f8), "+r"(c_08));                      // keep the two constants in registers
#pragma unroll
            for (int k = 0; k < 4; ++k)
                m16[k] = (class_mask16_shl7(v[2 * k], c_f8, c_08) >> 7) | (class_mask16_shl7(v[2 * k + 1], c_f8, c_08) << 9);
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t m = eq_mask16(v[k], 0x0a0a0a0au);
                cr_chunk |= has_byte(v[k].x, 0x0d0d0d0du) | has_byte(v[k].y, 0x0d0d0d0du) | has_byte(v[k].z, 0x0d0d0d0du) | has_byte(v[k].w, 0x0d0d0d0du);
                m16[k >> 1] |= m << (16 * (k & 1));
            }
        }
        any_cr |= cr_chunk;
        const uint32_t cnt = __popc(m16[0]) + __popc(m16[1]) + __popc(m16[2]) + __popc(m16[3]);
        const uint32_t incl = warp_incl_scan(cnt, lane);
        uint32_t rank = rank0 + incl - cnt;
        if (FAST) {
            // one entry per flagged byte, lowest address first; the flagged byte must BE '\n' (it sits in the cache line this
            // lane has just loaded).  Which byte precedes it is not recorded here: LF text has no CR line ends, and the '+' of
            // a bare "+" line is looked at by the kernel that decides whether a record can be block-copied (k_emit).
            uint32_t m0 = m16[0], m1 = m16[1], m2 = m16[2], m3 = m16[3];
            const uint8_t *mine = raw + lane_base;
            for (;;) {
                const uint32_t mw = m0 ? m0 : m1 ? m1 : m2 ? m2 : m3;
                if (!mw) break;
                const uint32_t off = (m0 ? 0u : m1 ? 32u : m2 ? 64u : 96u) + (uint32_t)(__ffs(mw) - 1);
                not_nl |= (uint32_t)mine[off] ^ 0x0au;
                if (rank < seg_cap) out[rank] = lane_base + off;
                ++rank;
                const uint32_t cleared = mw & (mw - 1);
                if (m0) m0 = cleared; else if (m1) m1 = cleared; else if (m2) m2 = cleared; else m3 = cleared;
            }
        } else {
        // The byte in front of a newline decides the flag "preceded by CR".  Loading it for every newline costs a third of the
        // kernel's L1 requests, so it is only looked at when this chunk or the one before holds a CR at all (warp-uniform).
        const bool cr_here = __any_sync(0xffffffffu, cr_chunk != 0);
        const bool look_cr = cr_here || cr_before;
        cr_before = cr_here;
#pragma unroll
        for (int w4 = 0; w4 < 4; ++w4) {
            uint32_t m = m16[w4];
            while (m) {
                const uint32_t b = (uint32_t)(__ffs(m) - 1);
                const uint32_t pos = lane_base + (uint32_t)(w4 * 32) + b;
                uint32_t prev = 0;
                if (pos && look_cr) prev = raw[pos - 1];                                   // L1 resident: just loaded
                const uint32_t e = pos | (prev == '\r' ? kNlCr : 0u);
                cr_eol += prev == '\r';
                if (rank < seg_cap) out[rank] = e;
                ++rank;
                m &= m - 1;
            }
        }
        }
        rank0 += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (FAST) {
        not_nl = __reduce_or_sync(0xffffffffu, not_nl);
        if (lane == 0) {
            seg_count[w] = rank0;
            if (rank0 > seg_cap) atomicMax(&info->seg_overflow_fast, rank0);
            if (not_nl) atomicOr(&info->frame_exact[mate], 1u);
        }
        return;
    }
    any_cr = __reduce_or_sync(0xffffffffu, any_cr);
    cr_eol = __reduce_add_sync(0xffffffffu, cr_eol);
    if (lane == 0) {
        seg_count[w] = rank0;
        if (rank0 > seg_cap) atomicMax(&info->seg_overflow, rank0);
        if (any_cr) atomicOr(&info->n_cr[mate], 1u);      // "some CR exists": the host then asks for the exact count
        if (cr_eol) atomicAdd(&info->n_cr_eol[mate], cr_eol);
    }
}

// Exclusive scan of the segment line counts (n_seg <= a few thousand): one CTA.  seg_base gets n_seg + 1 entries.
__global__ void __launch_bounds__(1024) k_scan_segments(const uint32_t *__restrict__ seg_count, uint32_t n_seg, uint32_t *__restrict__ seg_base,
                                                        BatchInfo *info, int mate)
{
    __shared__ uint32_t warp_sum[32];
    __shared__ uint32_t carry;
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_seg; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n_seg ? seg_count[i] : 0;
        const uint32_t x = warp_incl_scan(v, lane);
        if (lane == 31) warp_sum[wid] = x;
        __syncthreads();
        if (wid == 0) warp_sum[lane] = warp_incl_scan(warp_sum[lane], lane);
        __syncthreads();
        const uint32_t before = carry + (wid ? warp_sum[wid - 1] : 0) + (x - v);
        if (i < n_seg) seg_base[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        seg_base[n_seg] = carry;
        info->n_lines[mate] = carry;
        // the index was written by the fast framing instance unless it asked for the exact one
        if (info->frame_exact[mate] == 0 && info->seg_overflow_fast > info->seg_overflow) info->seg_overflow = info->seg_overflow_fast;
        info->seg_overflow_fast = 0;
    }
}

// Exact count of one byte value (only launched when k_frame_lines saw a '\r': CRLF input).
__global__ void __launch_bounds__(256) k_count_byte(const uint8_t *__restrict__ raw, uint64_t n, uint32_t c, uint32_t *out)
{
    uint32_t cnt = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) cnt += raw[i] == c;
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(out, cnt);
}

// Record descriptors from 4 consecutive lines of the segmented line index.  One thread per record; the
// raw bytes are not touched.  Grammar per fastq.cpp: the content of a line ends at the '\r' of a CRLF
// line end; |seq| must equal |qual| (fastq.cpp:118-122).
struct LineIndex {
    const uint32_t *nl_seg;
    const uint32_t *seg_base;     // [n_seg + 1] global line number of each segment's first line
    uint32_t seg_cap, n_seg;
    // entry of global line g; s is a hint that is advanced (lines are looked up in increasing order)
    __device__ __forceinline__ uint32_t at(uint32_t g, uint32_t &s) const
    {
        while (g >= seg_base[s + 1]) ++s;
        return nl_seg[(size_t)s * seg_cap + (g - seg_base[s])];
    }
    __device__ __forceinline__ uint32_t find(uint32_t g) const      // segment holding global line g
    {
        uint32_t lo = 0, hi = n_seg;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (seg_base[mid] <= g) lo = mid; else hi = mid;
        }
        return lo;
    }
};

__global__ void __launch_bounds__(256) k_build_records(const LineIndex li, uint32_t n_rec, Rec *__restrict__ rec,
                                                       uint8_t *__restrict__ canon, BatchInfo *info, int mate)
{
    __shared__ uint32_t s_seg;
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (threadIdx.x == 0) {
        // the CTA's 1024 lines sit in this segment or the next few.  The segments hold (nearly) equal numbers of bytes, so
        // the line's share of all lines is a good first guess; a short walk fixes it, the binary search (a dozen dependent
        // L2 round trips in front of every CTA: it was what this kernel spent its time on) is only the fallback
        const uint32_t r0 = blockIdx.x * blockDim.x, g = r0 ? 4 * r0 - 1 : 0, n_lines = 4 * n_rec;
        uint32_t s = (uint32_t)(((unsigned long long)g * li.n_seg) / n_lines);
        if (s >= li.n_seg) s = li.n_seg - 1;
        int steps = 0;
        while (steps < 8 && s > 0 && li.seg_base[s] > g) { --s; ++steps; }
        while (steps < 8 && s + 1 < li.n_seg && li.seg_base[s + 1] <= g) { ++s; ++steps; }
        if (steps >= 8) s = li.find(g);
        s_seg = s;
    }
    __syncthreads();
    uint32_t len = 0;
    bool bad = false;
    if (r < n_rec) {
        uint32_t s = s_seg;
        uint32_t hdr, ex, ey, ez, ew;
        {
            // the five line-index entries of a record (the end of the previous record's last line and its own four
            // line ends) are consecutive inside one segment except at segment boundaries: locate the segment once,
            // then five independent loads
            const uint32_t g0 = r ? 4 * r - 1 : 0, g4 = 4 * r + 3;
            while (g0 >= li.seg_base[s + 1]) ++s;
            const uint32_t b0 = li.seg_base[s], b1 = li.seg_base[s + 1];
            if (g4 < b1) {
                const uint32_t *p = li.nl_seg + (size_t)s * li.seg_cap + (g0 - b0);
                if (r) { hdr = (p[0] & kNlPosMask) + 1; ++p; } else hdr = 0;
                ex = p[0]; ey = p[1]; ez = p[2]; ew = p[3];
            } else {
                hdr = r ? (li.at(4 * r - 1, s) & kNlPosMask) + 1 : 0;
                ex = li.at(4 * r, s); ey = li.at(4 * r + 1, s); ez = li.at(4 * r + 2, s); ew = li.at(4 * r + 3, s);
            }
        }
        const uint32_t p0 = ex & kNlPosMask, p1 = ey & kNlPosMask, p2 = ez & kNlPosMask, p3 = ew & kNlPosMask;
        const uint32_t seq = p0 + 1, plus = p1 + 1, qual = p2 + 1;
        // a '\r' right before the '\n' belongs to the line end only if the line is not empty
        const uint32_t c0 = (ex & kNlCr) && p0 > hdr, c1 = (ey & kNlCr) && p1 > seq;
        const uint32_t c2 = (ez & kNlCr) && p2 > plus, c3 = (ew & kNlCr) && p3 > qual;
        len = p1 - seq - c1;
        const uint32_t qlen = p3 - qual - c3;
        bad = (len != qlen);
        rec[r] = Rec{hdr, seq, qual, len};
        // canonical record: LF line ends and a one-character third line; when that character is '+' (k_trim looks at it
        // and clears the claim in its verdict otherwise) the raw bytes ARE what write_read (fastq.cpp:127-138) prints for an
        // untouched read, so emission can be a block copy
        canon[r] = (uint8_t)(!(c0 | c1 | c2 | c3) && p2 == plus + 1);
    }
    uint32_t m = len;
#pragma unroll
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    // one atomic per warp on ONE address is what this kernel would spend its time on: look first, almost every warp then skips it
    if ((threadIdx.x & 31) == 0 && m > *reinterpret_cast<volatile uint32_t *>(&info->max_len[mate])) atomicMax(&info->max_len[mate], m);
    if (bad) {
        atomicOr(&info->err, kErrLenMismatch);
        atomicMin(&info->err_record, r);
    }
}

// A '\r' that is not the byte in front of a line's '\n' (rare: never in files written by sequencers or by FaQCs).  The
// reference cuts the content of every line at its FIRST '\r' or '\n' (strpbrk, fastq.cpp:44,70,100) and drops the rest of
// the line, so the lengths k_build_records derived from the line ends are too long for such lines.  One thread per record
// re-reads the header, base and quality lines up to their first '\r' / '\n', rewrites the length, re-checks
// |seq| == |qual| and marks the record with code 2 (the emitters then look for the first '\r' of the header as well).
// Only launched when the batch holds such a '\r' (the exact framing instance counts them).
__global__ void __launch_bounds__(256) k_fix_lone_cr(const uint8_t *__restrict__ raw, Rec *__restrict__ rec, uint8_t *__restrict__ canon, uint32_t n_rec,
                                                     BatchInfo *info, int mate)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    Rec rc = rec[r];
    auto content = [&](uint32_t from, bool &cr) -> uint32_t {       // bytes of the line starting at `from` up to its first '\r' or '\n'
        uint32_t k = from;
        while (raw[k] != '\n' && raw[k] != '\r') ++k;
        cr = raw[k] == '\r' && raw[k + 1] != '\n';                  // a CR that is not part of a CRLF line end
        return k - from;
    };
    bool cr_h, cr_s, cr_q;
    content(rc.hdr, cr_h);
    const uint32_t ls = content(rc.seq, cr_s), lq = content(rc.qual, cr_q);
    if (ls != lq) {
        atomicOr(&info->err, kErrLenMismatch);
        atomicMin(&info->err_record, r);
    }
    if (cr_h || cr_s || cr_q) canon[r] = 2;
    if (ls != rc.len) {
        rc.len = ls;
        rec[r] = rc;
    }
}

// auto_detect_quality_offset (trim.cpp:599-617): the first quality char, in read
// order then position order, that is > 74 (=> 64) or < 59 (=> 33).  One warp per read;
// each decisive read contributes key = record << 8 | offset, the minimum key wins.
__global__ void __launch_bounds__(256) k_detect_offset(const uint8_t *__restrict__ raw, const Rec *__restrict__ rec,
                                                       uint32_t n_rec, BatchInfo *info)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n_rec) return;
    const Rec rc = rec[r];
    const signed char *q = reinterpret_cast<const signed char *>(raw + rc.qual);
    for (uint32_t b = 0; b < rc.len; b += 32) {
        const uint32_t p = b + lane;
        const int c = p < rc.len ? (int)q[p] : 60;
        const uint32_t hi = __ballot_sync(0xffffffffu, c > 74);
        const uint32_t lo = __ballot_sync(0xffffffffu, c < 59);
        if (hi | lo) {
            const int first = __ffs(hi | lo) - 1;
            const uint32_t off = ((hi >> first) & 1u) ? 64u : 33u;     // '> 74' is tested first (trim.cpp:605)
            if (lane == 0) atomicMin(&info->detect_key, ((unsigned long long)r << 8) | off);
            return;
        }
    }
}

// parse_id(r1.def) == parse_id(r2.def) (trim.cpp:188-222, FaQCs.cpp:383-389).  One thread per pair.
// Headers are read as aligned 32-bit words (funnel-shifted to the header start): the common
// case -- ids equal up to the first space -- is decided after ~|id|/4 word compares.
__device__ __forceinline__ uint32_t load_word_at(const uint8_t *p)       // generic pointer: global or shared memory
{
    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3u) * 8u;
    const uint32_t *w = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)3);
    const uint32_t lo = w[0];
    const uint32_t hi = sh ? w[1] : 0u;
    return __funnelshift_r(lo, hi, sh);
}
// parse_id (trim.cpp:188-222): the id ends at the first space; a trailing "/1", ".2" ... is not part of it.
__device__ __forceinline__ uint32_t id_length(const uint8_t *h, uint32_t n)
{
    uint32_t loc = 0;
    while (loc < n && h[loc] != ' ') ++loc;
    if (loc > 1 && h[loc - 1] >= '0' && h[loc - 1] <= '9' && (h[loc - 2] == '.' || h[loc - 2] == '/')) loc -= 2;
    return loc;
}
// parse_id equality (trim.cpp:188-222, FaQCs.cpp:383-389) of two header lines of na / nb bytes (line end excluded,
// a CR of a CRLF end possibly included).  Fast path: both headers agree word by word up to and including a space
// (no '/1' '.1' suffix before it); the exact byte path only when a word differs before the first space.
__device__ __forceinline__ bool pair_ids_differ(const uint8_t *ha, uint32_t na, const uint8_t *hb, uint32_t nb)
{
    {
        const uint32_t m = min(na, nb);
        bool decided = false, same = false;
        for (uint32_t i = 0; i + 4 <= m && !decided; i += 4) {
            const uint32_t wa = load_word_at(ha + i), wb = load_word_at(hb + i);
            const uint32_t sp = __vcmpeq4(wa, 0x20202020u);      // 0xff where wa holds a space
            const uint32_t diff = wa ^ wb;
            const uint32_t fs = sp ? (uint32_t)((__ffs(sp) - 1) >> 3) : 4u;      // first space (little endian byte order)
            const uint32_t fd = diff ? (uint32_t)((__ffs(diff) - 1) >> 3) : 4u;  // first differing byte
            if (fs < fd) { same = true; decided = true; }        // identical up to and including the first space
            else if (fd < 4) decided = true;                     // differ before a space: let the exact path decide
        }
        if (decided && same) return false;
    }
    // the content of a header line ends at its first '\r' (fastq.cpp:44)
    for (uint32_t i = 0; i < na; ++i) if (ha[i] == '\r') { na = i; break; }
    for (uint32_t i = 0; i < nb; ++i) if (hb[i] == '\r') { nb = i; break; }
    const uint32_t la = id_length(ha, na), lb = id_length(hb, nb);
    bool same = (la == lb);
    for (uint32_t i = 0; same && i < la; ++i) same = (ha[i] == hb[i]);
    return !same;
}
__global__ void __launch_bounds__(256) k_check_pair_ids(const uint8_t *__restrict__ raw1, const Rec *__restrict__ rec1,
                                                        const uint8_t *__restrict__ raw2, const Rec *__restrict__ rec2,
                                                        uint32_t n_rec, BatchInfo *info)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    const Rec a = rec1[r], b = rec2[r];
    if (pair_ids_differ(raw1 + a.hdr, a.seq - a.hdr - 1, raw2 + b.hdr, b.seq - b.hdr - 1)) {
        atomicOr(&info->err, kErrPairId);
        atomicMin(&info->err_record, r);
    }
}

}  // namespace fq
