// fq_emit.cuh -- pair routing and in-order stream compaction
// (FaQCs.cpp:296-361,431-496 paired; :634-659,696-720 unpaired; write_read, fastq.cpp:127-138).
//
// Three steps: k_route sums the emitted bytes of each 256-record tile for the four streams, k_scan_tiles turns
// the tile sums into tile bases, k_emit recomputes the per-record sizes and scans them inside the tile.  Each
// k_emit warp owns 32 consecutive records: it stages their contiguous raw slab in shared memory (cp.async), then
// copies runs of untouched records as block copies and writes the other records `def\nseq\n+\nqual\n` piecewise,
// applying the mutations trim_read leaves in the read (terminal-N quality mask, G->N replacement, quality
// re-encoding).  Discarded reads are copied raw.  The read-id comparison of the two mates (FaQCs.cpp:383-389)
// rides along while both headers pass through the kernel.
#pragma once
#include "fq_common.cuh"
#include "fq_frame.cuh"
#include "fq_trim.cuh"

namespace fq {

constexpr uint32_t kTile = 256;     // records per route/emit tile

struct EmitArgs {
    const uint8_t *raw[2];
    uint64_t raw_bytes[2];
    const Rec *rec[2];
    const uint2 *res[2];
    const uint8_t *canon[2];     // record code: 1 = LF line ends + one-character third line (canonical: raw bytes == write_read
                                 // output once k_trim has seen the '+'), 0 = CRLF line ends or a longer third line, 2 = a '\r'
                                 // somewhere inside a line (the content of a line ends at its FIRST '\r', fastq.cpp:44)
    uint32_t n_rec;
    uint32_t parts;              // k_emit stages the 32 records of a warp in this many rounds (1, 2, 4, 8)
    uint32_t prefetch_tiles;     // k_emit asks the L2 for the slabs of the tile this many tiles ahead (0: off)
    uint32_t check_ids;          // k_emit also compares the read ids of the two mates (FaQCs.cpp:383-389)
    uint32_t n_tiles;
    uint32_t *tile_sum;          // [4][n_tiles] -> exclusive bases after k_scan_tiles (u32: < 4 GiB per stream per batch)
                                 // pieces mode: [12][n_tiles]: stream bytes, pieces, literal bytes
    uint8_t *out[4];             // pieces mode: the literal bytes of each stream
    fq_out_piece *pieces[4];     // pieces mode: the streams as lists of pieces
    BatchInfo *info;
    unsigned long long *stats;   // global stats block (PAIRED_* counters are added here too)
    size_t filter_off;
};

__device__ __forceinline__ uint32_t header_len(const uint8_t *raw, const Rec &rc, uint32_t ccode = 0)
{
    uint32_t n = rc.seq - rc.hdr - 1;
    if (ccode == 1 || n == 0) return n;
    if (ccode == 2) {                                   // strpbrk semantics: the header ends at its first '\r'
        for (uint32_t i = 0; i < n; ++i)
            if (raw[rc.hdr + i] == '\r') return i;
        return n;
    }
    if (raw[rc.hdr + n - 1] == '\r') --n;
    return n;
}

// Sizes of what record r contributes to each stream.
__device__ __forceinline__ void route_sizes(const EmitArgs &a, const DevOpts &o, uint32_t r, uint32_t sz[4], bool valid[2],
                                            uint32_t wl[2])
{
    sz[0] = sz[1] = sz[2] = sz[3] = 0;
    valid[0] = valid[1] = false;
    wl[0] = wl[1] = 0;
    if (r >= a.n_rec) return;
    uint32_t trimmed[2] = {0, 0}, rawsz[2] = {0, 0};
    const int n_mates = o.paired ? 2 : 1;
    for (int m = 0; m < n_mates; ++m) {
        const Rec rc = a.rec[m][r];
        const uint2 v = a.res[m][r];
        const uint32_t hl = header_len(a.raw[m], rc, a.canon[m][r]);
        wl[m] = v.y & kResLenMask;
        valid[m] = ((v.y >> kResLenBits) & FQ_RR_VALID) != 0;
        trimmed[m] = hl + 2 * wl[m] + 5;
        rawsz[m] = hl + 2 * rc.len + 5;
    }
    if (o.qc_only) return;
    if (o.paired) {
        if (valid[0] && valid[1]) { sz[0] = trimmed[0]; sz[1] = trimmed[1]; }
        else {
            if (valid[0]) sz[2] = trimmed[0];
            else if (valid[1]) sz[2] = trimmed[1];
            if (o.discard) sz[3] = (valid[0] ? 0 : rawsz[0]) + (valid[1] ? 0 : rawsz[1]);
        }
    } else {
        if (valid[0]) sz[2] = trimmed[0];
        else if (o.discard) sz[3] = rawsz[0];
    }
}

__global__ void __launch_bounds__(kTile) k_route(const EmitArgs a, const DevOpts o)
{
    __shared__ uint32_t s_sum[8][8];
    const uint32_t r = blockIdx.x * kTile + threadIdx.x;
    uint32_t sz[4], wl[2];
    bool valid[2];
    route_sizes(a, o, r, sz, valid, wl);
    uint32_t v[8];
    v[0] = sz[0]; v[1] = sz[1]; v[2] = sz[2]; v[3] = sz[3];
    v[4] = valid[0]; v[5] = valid[1];
    const bool both = o.paired && valid[0] && valid[1];
    v[6] = both ? 2u : 0u;                       // PAIRED_READ_NUMBER (FaQCs.cpp:304-308)
    v[7] = both ? wl[0] + wl[1] : 0u;            // PAIRED_BASE_LENGTH
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = warp_sum(v[k]);
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) s_sum[wid][k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        uint32_t t = 0;
        for (int w = 0; w < (int)(kTile / 32); ++w) t += s_sum[w][threadIdx.x];
        if (threadIdx.x < 4) a.tile_sum[threadIdx.x * a.n_tiles + blockIdx.x] = t;
        else if (t) {
            unsigned long long *dst = threadIdx.x == 4 ? &a.info->n_valid[0] : threadIdx.x == 5 ? &a.info->n_valid[1]
                                    : threadIdx.x == 6 ? &a.info->paired_reads : &a.info->paired_bases;
            atomicAdd(dst, (unsigned long long)t);
            if (threadIdx.x == 6) atomicAdd(&a.stats[a.filter_off + FQ_PAIRED_READ_NUMBER], (unsigned long long)t);
            if (threadIdx.x == 7) atomicAdd(&a.stats[a.filter_off + FQ_PAIRED_BASE_LENGTH], (unsigned long long)t);
        }
    }
}

// Exclusive scan of the four tile-sum rows; totals go to info->out_bytes.  One 1024-thread CTA per stream,
// 1024 tiles per round (warp scans + a scan of the 32 warp totals).
__global__ void __launch_bounds__(1024) k_scan_tiles(uint32_t *tile_sum, uint32_t n_tiles, BatchInfo *info)
{
    __shared__ uint32_t s_warp[32];
    __shared__ unsigned long long s_carry;
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5, s = blockIdx.x;
    uint32_t *row = tile_sum + (size_t)s * n_tiles;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_tiles; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n_tiles ? row[i] : 0;
        uint32_t x = v;
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, k);
            if (lane >= (uint32_t)k) x += y;
        }
        if (lane == 31) s_warp[wid] = x;
        __syncthreads();
        if (wid == 0) {
            uint32_t t = s_warp[lane];
#pragma unroll
            for (int k = 1; k < 32; k <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, t, k);
                if (lane >= (uint32_t)k) t += y;
            }
            s_warp[lane] = t;
        }
        __syncthreads();
        const unsigned long long before = s_carry + (wid ? s_warp[wid - 1] : 0u) + (x - v);
        if (i < n_tiles) row[i] = (uint32_t)before;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (s < 4) info->out_bytes[s] = s_carry;
        else if (s < 8) info->out_pieces[s - 4] = s_carry;
        else info->out_literal[s - 8] = s_carry;
    }
}

// Warp-cooperative copy of n bytes between arbitrarily aligned addresses.  The body is written
// as 16-byte aligned stores; each store gathers its bytes from five aligned 32-bit source words
// with funnel shifts.  The source is a generic pointer: normally the warp's staged slab in shared
// memory (k_emit), global memory when a slab does not fit.
__device__ __forceinline__ uint4 gather16(const uint8_t *sc)
{
    // 16 bytes starting at the arbitrarily aligned address sc, from five aligned 32-bit words
    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(sc) & 3u) * 8u;
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(sc) & ~(uintptr_t)3);
    const uint32_t w0 = sw[0], w1 = sw[1], w2 = sw[2], w3 = sw[3];
    const uint32_t w4 = sh ? sw[4] : 0u;                  // only bytes below src + n are ever consumed from it
    uint4 v;
    v.x = __funnelshift_r(w0, w1, sh);
    v.y = __funnelshift_r(w1, w2, sh);
    v.z = __funnelshift_r(w2, w3, sh);
    v.w = __funnelshift_r(w3, w4, sh);
    return v;
}

// `lane` / W: position and size of the cooperating lane group (a whole warp, or an 8-lane sub-group).
template <uint32_t W = 32>
__device__ __forceinline__ void copy_span(uint8_t *__restrict__ dst, const uint8_t *__restrict__ src, uint32_t n, uint32_t lane)
{
    if (n < 48) {
        for (uint32_t i = lane; i < n; i += W) dst[i] = src[i];
        return;
    }
    const uint32_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u;
    for (uint32_t i = lane; i < head; i += W) dst[i] = src[i];
    const uint32_t body = (n - head) >> 4;
    const uint8_t *s = src + head;
    uint8_t *d = dst + head;
    uint32_t c = lane;
    // four 16-byte chunks per lane per round: all loads are issued before the first store, so a warp keeps
    // 2 KiB in flight (the compiler cannot hoist loads over stores through possibly aliasing pointers)
    for (; c + 3 * W < body; c += 4 * W) {
        const uint4 v0 = gather16(s + 16 * c), v1 = gather16(s + 16 * (c + W)), v2 = gather16(s + 16 * (c + 2 * W)), v3 = gather16(s + 16 * (c + 3 * W));
        *reinterpret_cast<uint4 *>(d + 16 * c) = v0;
        *reinterpret_cast<uint4 *>(d + 16 * (c + W)) = v1;
        *reinterpret_cast<uint4 *>(d + 16 * (c + 2 * W)) = v2;
        *reinterpret_cast<uint4 *>(d + 16 * (c + 3 * W)) = v3;
    }
    {
        // up to three more chunks for this lane
        const bool p0 = c < body, p1 = c + W < body, p2 = c + 2 * W < body;
        uint4 v0 = make_uint4(0, 0, 0, 0), v1 = v0, v2 = v0;
        if (p0) v0 = gather16(s + 16 * c);
        if (p1) v1 = gather16(s + 16 * (c + W));
        if (p2) v2 = gather16(s + 16 * (c + 2 * W));
        if (p0) *reinterpret_cast<uint4 *>(d + 16 * c) = v0;
        if (p1) *reinterpret_cast<uint4 *>(d + 16 * (c + W)) = v1;
        if (p2) *reinterpret_cast<uint4 *>(d + 16 * (c + 2 * W)) = v2;
    }
    for (uint32_t i = head + 16 * body + lane; i < n; i += W) dst[i] = src[i];
}

// copy_span with a byte-wise transformation applied on the way (16 bytes at a time through SIMD-in-word video instructions,
// single bytes at the unaligned ends).  F::vec(a, b) / F::one(a, b): a = bytes of `src`, b = bytes of `src2` at the same offsets.
struct XfRequal {           // quality re-encoding (trim.cpp:516-525): max(0, q - in_off) + out_off
    uint32_t in4, out4;
    int in_off, out_off;
    __device__ __forceinline__ XfRequal(int i, int o) : in4(0x01010101u * (uint32_t)(i & 0xff)), out4(0x01010101u * (uint32_t)(o & 0xff)), in_off(i), out_off(o) {}
    __device__ __forceinline__ uint32_t vec(uint32_t q, uint32_t) const { return __vadd4(__vmaxs4(__vsubss4(q, in4), 0u), out4); }
    __device__ __forceinline__ uint8_t one(uint8_t q, uint8_t) const { return (uint8_t)(max(0, (int)(signed char)q - in_off) + out_off); }
};
struct XfLowG {             // G -> N below --replace_to_N_q (trim.cpp:389-403); a = base, b = its quality character
    uint32_t in4, rq4;
    int in_off, rq;
    __device__ __forceinline__ XfLowG(int i, uint32_t r) : in4(0x01010101u * (uint32_t)(i & 0xff)), rq4(0x01010101u * min(r, 255u)), in_off(i), rq((int)r) {}
    __device__ __forceinline__ uint32_t vec(uint32_t s, uint32_t q) const
    {
        const uint32_t qv = __vmaxs4(__vsubss4(q, in4), 0u);                      // 0..127 per byte
        const uint32_t m = __vcmpeq4(s, 0x47474747u) & __vcmpltu4(qv, rq4);       // 0xff where a 'G' is below the score
        return (s & ~m) | (0x4e4e4e4eu & m);
    }
    __device__ __forceinline__ uint8_t one(uint8_t s, uint8_t q) const
    {
        return (s == 'G' && max(0, (int)(signed char)q - in_off) < rq) ? (uint8_t)'N' : s;
    }
};

template <uint32_t W, class F>
__device__ __forceinline__ void xf_span(uint8_t *__restrict__ dst, const uint8_t *__restrict__ src, const uint8_t *__restrict__ src2, uint32_t n,
                                        uint32_t lane, const F f)
{
    const bool two = src2 != nullptr;
    if (n < 48) {
        for (uint32_t i = lane; i < n; i += W) dst[i] = f.one(src[i], two ? src2[i] : (uint8_t)0);
        return;
    }
    const uint32_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u;
    for (uint32_t i = lane; i < head; i += W) dst[i] = f.one(src[i], two ? src2[i] : (uint8_t)0);
    const uint32_t body = (n - head) >> 4;
    for (uint32_t c = lane; c < body; c += W) {
        const uint4 a = gather16(src + head + 16 * c);
        uint4 b = make_uint4(0, 0, 0, 0);
        if (two) b = gather16(src2 + head + 16 * c);
        *reinterpret_cast<uint4 *>(dst + head + 16 * c) = make_uint4(f.vec(a.x, b.x), f.vec(a.y, b.y), f.vec(a.z, b.z), f.vec(a.w, b.w));
    }
    for (uint32_t i = head + 16 * body + lane; i < n; i += W) dst[i] = f.one(src[i], two ? src2[i] : (uint8_t)0);
}

// Emit one surviving read: def \n seq \n + \n qual \n (write_read, fastq.cpp:127-138) with the
// mutations trim_read leaves behind.  `plain`: the record is canonical, untrimmed and untouched,
// so the output is its raw bytes.
template <uint32_t W = 32>
__device__ __forceinline__ void write_trimmed(uint8_t *dst, const uint8_t *raw, const Rec &rc, uint32_t ccode, uint32_t lo, uint32_t wl,
                                              uint32_t flags, const DevOpts &o, uint32_t lane)
{
    const bool canon = ccode == 1;
    const uint32_t hl = header_len(raw, rc, ccode);
    const bool masked = (flags & kFlagMasked) != 0;
    const bool requal = o.in_off != o.out_off;
    if (canon && lo == 0 && wl == rc.len && !masked && !requal && o.replace_q == 0) {
        copy_span<W>(dst, raw + rc.hdr, hl + 2 * wl + 5, lane);
        return;
    }
    const uint8_t *sp = raw + rc.seq;
    const signed char *qp = reinterpret_cast<const signed char *>(raw + rc.qual);
    uint32_t lead = 0, trail = rc.len;
    if (masked) {                                               // terminal-N mask bounds (trim.cpp:1191-1216)
        while (lead < rc.len && sp[lead] == 'N') ++lead;
        while (trail > 0 && sp[trail - 1] == 'N') --trail;
        if (lead >= rc.len) trail = 0;
    }
    const uint32_t s0 = hl + 1, s1 = s0 + wl, q0 = s1 + 3, q1 = q0 + wl;
    // header (+ its '\n' and, when nothing was cut at the 5' end, the bases: one contiguous source run)
    if (canon && lo == 0 && o.replace_q == 0) copy_span<W>(dst, raw + rc.hdr, s1, lane);
    else {
        copy_span<W>(dst, raw + rc.hdr, hl, lane);
        if (lane == 0) dst[hl] = '\n';
        // G -> N below --replace_to_N_q (trim.cpp:390-403); the terminal-N mask only covers 'N' bases, never a 'G'
        if (o.replace_q == 0) copy_span<W>(dst + s0, sp + lo, wl, lane);
        else xf_span<W>(dst + s0, sp + lo, reinterpret_cast<const uint8_t *>(qp) + lo, wl, lane, XfLowG(o.in_off, o.replace_q));
    }
    if (lane < 3) dst[s1 + lane] = lane == 1 ? '+' : '\n';
    if (!masked && !requal) copy_span<W>(dst + q0, raw + rc.qual + lo, wl, lane);
    else if (!masked) xf_span<W>(dst + q0, raw + rc.qual + lo, nullptr, wl, lane, XfRequal(o.in_off, o.out_off));
    else {
        for (uint32_t i = lane; i < wl; i += W) {
            const uint32_t p = lo + i;
            int qc = (p < lead || p >= trail) ? o.in_off : (int)qp[p];
            if (requal) qc = max(0, qc - o.in_off) + o.out_off;  // trim.cpp:516-525
            dst[q0 + i] = (uint8_t)qc;
        }
    }
    if (lane == 3) dst[q1] = '\n';
}

// Emit one discarded read: the raw, unmasked record (copy taken before trim(), FaQCs.cpp:279-285).
template <uint32_t W = 32>
__device__ __forceinline__ void write_raw(uint8_t *dst, const uint8_t *raw, const Rec &rc, uint32_t ccode, uint32_t lane)
{
    const bool canon = ccode == 1;
    const uint32_t hl = header_len(raw, rc, ccode);
    if (canon) {
        copy_span<W>(dst, raw + rc.hdr, hl + 2 * rc.len + 5, lane);
        return;
    }
    const uint32_t s0 = hl + 1, s1 = s0 + rc.len, q0 = s1 + 3, q1 = q0 + rc.len;
    copy_span<W>(dst, raw + rc.hdr, hl, lane);
    if (lane == 0) dst[hl] = '\n';
    copy_span<W>(dst + s0, raw + rc.seq, rc.len, lane);
    if (lane < 3) dst[s1 + lane] = lane == 1 ? '+' : '\n';
    copy_span<W>(dst + q0, raw + rc.qual, rc.len, lane);
    if (lane == 3) dst[q1] = '\n';
}

// Everything one lane knows about its record (pair) for emission.
struct LaneRec {
    Rec rc[2];
    uint2 res[2];
    uint32_t ccode[2];      // record code (EmitArgs::canon) with the '+' check folded in: 1 only if the raw bytes are the canonical record
    bool valid[2], plain[2];
    uint32_t tsize[2];      // bytes of the trimmed record
};

__device__ __forceinline__ LaneRec lane_record(const EmitArgs &a, const DevOpts &o, uint32_t r)
{
    LaneRec L{};
    if (r >= a.n_rec) return L;
    const int n_mates = o.paired ? 2 : 1;
    const bool requal = o.in_off != o.out_off;
    for (int m = 0; m < n_mates; ++m) {
        L.rc[m] = a.rec[m][r];
        const uint2 rv = a.res[m][r];
        L.res[m] = make_uint2(rv.x & ~kResPlusBad, rv.y);
        L.ccode[m] = a.canon[m][r];
        if (L.ccode[m] == 1 && (rv.x & kResPlusBad)) L.ccode[m] = 0;
        const uint32_t fl = L.res[m].y >> kResLenBits, wl = L.res[m].y & kResLenMask;
        L.valid[m] = (fl & FQ_RR_VALID) != 0;
        L.tsize[m] = header_len(a.raw[m], L.rc[m], L.ccode[m]) + 2 * wl + 5;
        // untouched canonical record: the emitted bytes are the raw bytes
        L.plain[m] = L.valid[m] && L.ccode[m] == 1 && L.res[m].x == 0 && wl == L.rc[m].len && !(fl & kFlagMasked) && !requal && o.replace_q == 0;
    }
    return L;
}

// Copy every run of consecutive lanes flagged in `mask` as ONE contiguous span: consecutive canonical
// records are adjacent in the raw input and, when they go to the same stream, adjacent in the output.
__device__ __forceinline__ void copy_runs(uint32_t mask, uint8_t *out, const uint8_t *raw, uint32_t src_mine, uint32_t size_mine,
                                          uint32_t dst_mine, uint32_t lane)
{
    while (mask) {
        const int first = __ffs(mask) - 1;
        const uint32_t rest = ~(mask >> first);
        const int run = rest ? __ffs(rest) - 1 : 32 - first;          // lanes first .. first+run-1
        const int last = first + run - 1;
        const uint32_t src = __shfl_sync(0xffffffffu, src_mine, first);
        const uint32_t dst = __shfl_sync(0xffffffffu, dst_mine, first);
        const uint32_t end = __shfl_sync(0xffffffffu, src_mine + size_mine, last);
        copy_span<32>(out + dst, raw + src, end - src, lane);
        mask &= (run + first >= 32) ? 0u : ~((1u << (first + run)) - 1u);
    }
}

// ---- k_emit -------------------------------------------------------------------------------------
// One warp owns 32 consecutive records.  Their raw bytes are one contiguous slab of the input
// (~10 KiB for 150-base reads), which the warp first copies asynchronously (cp.async, 16 bytes per
// request, no register staging) into its private shared-memory buffer.  Everything after that reads
// shared memory: the chains "descriptor -> header bytes -> bases -> qualities" that made the copy
// latency-bound on global memory now cost tens of cycles per link, and the only global traffic left
// is the coalesced slab load and the 16-byte stores.
#ifndef FQ_EMIT_SLAB
#define FQ_EMIT_SLAB (12 * 1024)
#endif
constexpr uint32_t kEmitSlab = FQ_EMIT_SLAB;            // staging bytes per warp
constexpr uint32_t kEmitSmem = (kTile / 32) * kEmitSlab;

__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Bulk (TMA) copy global -> shared of a 16-byte aligned range, completion on an mbarrier (SASS: UBLKCP + SYNCS).
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t smem_dst, const void *gsrc, uint32_t bytes, uint32_t bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(bar)
                 : "memory");
}
// Waits for the phase with the given parity; gives up after a bounded number of polls (returns false) instead of hanging.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity)
{
    for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
        uint32_t done;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return true;
    }
    return false;
}

extern __shared__ __align__(16) uint8_t g_emit_smem[];

#ifndef FQ_EMIT_MIN_CTAS
#define FQ_EMIT_MIN_CTAS 2
#endif
// stage the slabs with one bulk (TMA) copy per stage instead of 16-byte cp.async requests from every lane
#ifndef FQ_EMIT_BULK
#define FQ_EMIT_BULK 1
#endif
#ifndef FQ_EMIT_PREFETCH
#define FQ_EMIT_PREFETCH 0      // measured: 0.656 -> 0.746 ms with it (C2); the staging loads are not what bounds the kernel
#endif
// PLAIN: no quality re-encoding and no G->N replacement (the default run): those branches of write_trimmed are compiled out.
template <bool PLAIN>
__global__ void __launch_bounds__(kTile, FQ_EMIT_MIN_CTAS) k_emit(const EmitArgs a, const DevOpts o_in)
{
    DevOpts o = o_in;
    if (PLAIN) {
        o.replace_q = 0;
        o.out_off = o.in_off;
        o.qc_only = 0;
    }
    __shared__ uint32_t s_wsum[4][kTile / 32];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t r = blockIdx.x * kTile + threadIdx.x;
#if FQ_EMIT_PREFETCH
    // The tile one wave of CTAs ahead will be staged by the CTA that takes this one's place: ask the L2 for its slabs now
    // (lanes 0 / 1: mate 1 / mate 2), so that its staging loads find them there instead of waiting for DRAM.
    uint32_t pf_lo = 0, pf_hi = 0;
    {
        const uint32_t r_first = (blockIdx.x + a.prefetch_tiles) * kTile + wid * 32;
        if (lane < (o.paired ? 2u : 1u) && a.prefetch_tiles && r_first < a.n_rec) {
            const Rec *rec = lane ? a.rec[1] : a.rec[0];
            const uint32_t r_last = min(r_first + 31u, a.n_rec - 1u);
            pf_lo = rec[r_first].hdr & ~15u;
            const Rec last = rec[r_last];
            pf_hi = (uint32_t)min((uint64_t)((last.qual + last.len + 1 + 15) & ~15u), (lane ? a.raw_bytes[1] : a.raw_bytes[0]) & ~(uint64_t)15);
        }
    }
#endif
    uint32_t sz[4], wl[2];
    bool valid[2];
    route_sizes(a, o, r, sz, valid, wl);
#if FQ_EMIT_PREFETCH
    if (pf_hi > pf_lo) prefetch_l2_bulk((lane ? a.raw[1] : a.raw[0]) + pf_lo, pf_hi - pf_lo);
#endif
    uint32_t inc[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        uint32_t x = sz[s];
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, k);
            if (lane >= (uint32_t)k) x += y;
        }
        inc[s] = x;
        if (lane == 31) s_wsum[s][wid] = x;
    }
    __syncthreads();
    uint32_t off[4];        // where this lane's record starts in each stream
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        uint32_t before = a.tile_sum[(size_t)s * a.n_tiles + blockIdx.x];
        for (uint32_t w = 0; w < wid; ++w) before += s_wsum[s][w];
        off[s] = before + inc[s] - sz[s];
    }
    if (o.qc_only) return;
    const LaneRec L = lane_record(a, o, r);
    const bool in = r < a.n_rec;
    const uint32_t in_mask = __ballot_sync(0xffffffffu, in);
    if (in_mask == 0) return;
    const bool both = o.paired && in && L.valid[0] && L.valid[1];
    uint8_t *const slab = g_emit_smem + (size_t)wid * kEmitSlab;
    const uint32_t slab_s = (uint32_t)__cvta_generic_to_shared(slab);
#if FQ_EMIT_BULK
    __shared__ __align__(8) unsigned long long s_bar[kTile / 32];
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar[wid]);
    uint32_t bar_phase = 0;
    if (lane == 0) mbar_init(bar, 1);
    __syncwarp();
#endif
    const int n_mates = o.paired ? 2 : 1;

    // n_mates * P stages (mate, part): the slab of a stage is the raw bytes of 32 / P consecutive records; the host
    // picks P (1, 2, 4 or 8) so that a stage normally fits the buffer, a stage that does not reads global memory.
    // The loop is deliberately NOT unrolled (one copy of the body: instruction cache).  Splitting a slab that fits
    // into pipelined parts was measured and is slower (per-stage overhead), so P is 1 whenever possible.
    const int P = (int)a.parts;
    const int n_stages = n_mates * P;
    // per-stage view of this lane's record
    struct Mate { Rec rc; uint2 res; uint32_t ccode; bool valid, plain; uint32_t tsize; };
    auto mate_of = [&](int m) -> Mate {
        return m ? Mate{L.rc[1], L.res[1], L.ccode[1], L.valid[1], L.plain[1], L.tsize[1]} : Mate{L.rc[0], L.res[0], L.ccode[0], L.valid[0], L.plain[0], L.tsize[0]};
    };
    auto extent = [&](int t, const Rec &rc, uint32_t &lo, uint32_t &hi, uint32_t &part_mask) -> bool {
        const int m = t / P, part = t % P;
        const uint32_t per = 32u / (uint32_t)P;
        part_mask = (P == 1 ? 0xffffffffu : ((1u << per) - 1u) << (part * per)) & in_mask;
        lo = hi = 0;
        if (part_mask == 0) return false;
        const int first_lane = __ffs(part_mask) - 1, last_lane = 31 - __clz(part_mask);
        lo = __shfl_sync(0xffffffffu, rc.hdr, first_lane) & ~15u;
        const uint64_t end = min((uint64_t)__shfl_sync(0xffffffffu, rc.qual + rc.len, last_lane) + 1, m ? a.raw_bytes[1] : a.raw_bytes[0]);
        hi = (uint32_t)((end + 15) & ~(uint64_t)15);
        return true;
    };
#pragma unroll 1
    for (int t = 0; t < n_stages; ++t) {
        const int m = t / P;
        const Mate M = mate_of(m);
        uint32_t lo, hi, part_mask;
        if (extent(t, M.rc, lo, hi, part_mask)) {
            const bool mine = (part_mask >> lane) & 1u;
            const uint8_t *src = m ? a.raw[1] : a.raw[0];      // src[offset] = byte at raw offset `offset`
            if (hi - lo <= kEmitSlab) {
                __syncwarp();                                  // everyone is done reading the previous slab
#if FQ_EMIT_BULK
                if (lane == 0) bulk_g2s(slab_s, src + lo, hi - lo, bar);
                if (!mbar_wait(bar, bar_phase)) { atomicOr(&a.info->err, kErrInternal); return; }
                bar_phase ^= 1u;
#else
                for (uint32_t c = lo + 16 * lane; c < hi; c += 16 * 32) cp_async16(slab_s + (c - lo), src + c);
                cp_async_wait_all();
#endif
                __syncwarp();
                src = slab - lo;
            }
            if (m == 1 && a.check_ids && mine) {
                // mate 2's header sits in shared memory now; mate 1's was staged a moment ago and is read back from L2
                if (pair_ids_differ(a.raw[0] + L.rc[0].hdr, L.rc[0].seq - L.rc[0].hdr - 1, src + M.rc.hdr, M.rc.seq - M.rc.hdr - 1)) {
                    atomicOr(&a.info->err, kErrPairId);
                    atomicMin(&a.info->err_record, r);
                }
            }
            // ---- runs of untouched records of a surviving pair (or, unpaired input, of surviving reads): block copies
            const int s_main = o.paired ? m : 2;
            const bool main = mine && (o.paired ? both : M.valid);
            uint8_t *const out_main = s_main == 0 ? a.out[0] : s_main == 1 ? a.out[1] : a.out[2];
            const uint32_t off_main = s_main == 0 ? off[0] : s_main == 1 ? off[1] : off[2];
            copy_runs(__ballot_sync(0xffffffffu, main && M.plain), out_main, src, M.rc.hdr, M.tsize, off_main, lane);
            // ---- everything else of this mate, four records at a time (one per 8-lane group)
            const bool trimmed = mine && M.valid && !(main && M.plain);                // needs write_trimmed
            const bool disc = mine && o.discard && !M.valid;                           // raw copy to the discard stream
            uint8_t *dstp = nullptr;                                                    // where this mate's record goes
            if (trimmed) dstp = main ? out_main + off_main : a.out[2] + off[2];
            else if (disc) {
                dstp = a.out[3] + off[3];
                if (m == 1 && !L.valid[0]) dstp += header_len(a.raw[0], L.rc[0], L.ccode[0]) + 2 * L.rc[0].len + 5;
            }
            uint32_t todo = __ballot_sync(0xffffffffu, trimmed || disc);
            const uint32_t sub = lane & 7, grp = lane >> 3;
            while (todo) {
                int j = -1;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int b = todo ? __ffs(todo) - 1 : -1;
                    if (b >= 0) todo &= todo - 1;
                    if ((int)grp == g) j = b;
                }
                const int js = j < 0 ? 0 : j;
                Rec rc;
                rc.hdr = __shfl_sync(0xffffffffu, M.rc.hdr, js);
                rc.seq = __shfl_sync(0xffffffffu, M.rc.seq, js);
                rc.qual = __shfl_sync(0xffffffffu, M.rc.qual, js);
                rc.len = __shfl_sync(0xffffffffu, M.rc.len, js);
                const uint32_t ex = __shfl_sync(0xffffffffu, M.res.x, js), ey = __shfl_sync(0xffffffffu, M.res.y, js);
                const uint32_t cn = __shfl_sync(0xffffffffu, M.ccode, js);
                const bool tr = __shfl_sync(0xffffffffu, (int)trimmed, js) != 0;
                const unsigned long long dp = __shfl_sync(0xffffffffu, (unsigned long long)dstp, js);
                if (j < 0) continue;
                uint8_t *const outp = reinterpret_cast<uint8_t *>(dp);
                if (tr) write_trimmed<8>(outp, src, rc, cn, ex, ey & kResLenMask, ey >> kResLenBits, o, sub);
                else write_raw<8>(outp, src, rc, cn, sub);
            }
        }
    }
}

// ---- pieces mode ----------------------------------------------------------------------------------
// The caller keeps its input buffers, and almost every surviving record of a real run is emitted exactly as it came in
// (98.9 % of the C2 records).  In pieces mode a stream is therefore returned as a list of PIECES in stream order: a piece is
// a byte range of one of the caller's input buffers (runs of consecutive untouched records of a warp are ONE piece; a
// canonical discarded record is a piece of its own) or a range of the stream's literal bytes (trimmed, masked, re-encoded
// or re-formatted records, written exactly as byte mode writes them).  Device -> host traffic drops from the size of the
// output to the literal bytes plus 16 bytes per piece.
struct LaneItems {
    LaneRec L;
    int stream[2];          // stream the mate goes to (-1: nowhere)
    uint32_t bytes[2];      // its bytes in that stream
    bool copy[2];           // the bytes are a range of the input (no literal bytes)
    bool main[2];           // the mate sits in its main stream (R1 / R2 of a surviving pair, the only stream of single-end input)
};

__device__ __forceinline__ LaneItems lane_items(const EmitArgs &a, const DevOpts &o, uint32_t r)
{
    LaneItems I;
    I.L = lane_record(a, o, r);
    I.stream[0] = I.stream[1] = -1;
    I.bytes[0] = I.bytes[1] = 0;
    I.copy[0] = I.copy[1] = I.main[0] = I.main[1] = false;
    if (r >= a.n_rec || o.qc_only) return I;
    const LaneRec &L = I.L;
    const int n_mates = o.paired ? 2 : 1;
    const bool both = o.paired && L.valid[0] && L.valid[1];
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        if (m >= n_mates) break;
        if (L.valid[m]) {
            I.main[m] = o.paired ? both : true;
            I.stream[m] = both ? m : 2;
            I.bytes[m] = L.tsize[m];
            I.copy[m] = L.plain[m];
        } else if (o.discard) {
            I.stream[m] = 3;
            I.bytes[m] = header_len(a.raw[m], L.rc[m], L.ccode[m]) + 2 * L.rc[m].len + 5;
            I.copy[m] = L.ccode[m] == 1;
        }
    }
    return I;
}

// Per lane and stream: pieces this record starts and literal bytes it adds.  Runs of copies in a main stream count once.
__device__ __forceinline__ void piece_counts(const LaneItems &I, uint32_t lane, uint32_t (&pc)[4], uint32_t (&lb)[4], bool (&starts)[2], uint32_t (&run_mask)[2])
{
#pragma unroll
    for (int s = 0; s < 4; ++s) pc[s] = lb[s] = 0;
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        const bool in_run = I.stream[m] >= 0 && I.main[m] && I.copy[m];
        run_mask[m] = __ballot_sync(0xffffffffu, in_run);
        const bool run_start = in_run && !(lane && ((run_mask[m] >> (lane - 1)) & 1u));
        starts[m] = I.stream[m] >= 0 && (!in_run || run_start);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            if (I.stream[m] == s) {
                pc[s] += starts[m] ? 1u : 0u;
                lb[s] += I.copy[m] ? 0u : I.bytes[m];
            }
        }
    }
}

__global__ void __launch_bounds__(kTile) k_route_pieces(const EmitArgs a, const DevOpts o)
{
    __shared__ uint32_t s_sum[8][16];
    const uint32_t r = blockIdx.x * kTile + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const LaneItems I = lane_items(a, o, r);
    uint32_t pc[4], lb[4], run_mask[2];
    bool starts[2];
    piece_counts(I, lane, pc, lb, starts, run_mask);
    uint32_t v[16];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        v[s] = (I.stream[0] == s ? I.bytes[0] : 0u) + (I.stream[1] == s ? I.bytes[1] : 0u);
        v[4 + s] = pc[s];
        v[8 + s] = lb[s];
    }
    const bool in = r < a.n_rec;
    const bool both = o.paired && in && I.L.valid[0] && I.L.valid[1];
    v[12] = in && I.L.valid[0];
    v[13] = in && o.paired && I.L.valid[1];
    v[14] = both ? 2u : 0u;                                                                   // PAIRED_READ_NUMBER (FaQCs.cpp:304-308)
    v[15] = both ? (I.L.res[0].y & kResLenMask) + (I.L.res[1].y & kResLenMask) : 0u;          // PAIRED_BASE_LENGTH
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = warp_sum(v[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 16; ++k) s_sum[wid][k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        uint32_t t = 0;
        for (int w = 0; w < (int)(kTile / 32); ++w) t += s_sum[w][threadIdx.x];
        if (threadIdx.x < 12) a.tile_sum[threadIdx.x * a.n_tiles + blockIdx.x] = t;
        else if (t) {
            unsigned long long *dst = threadIdx.x == 12 ? &a.info->n_valid[0] : threadIdx.x == 13 ? &a.info->n_valid[1]
                                    : threadIdx.x == 14 ? &a.info->paired_reads : &a.info->paired_bases;
            atomicAdd(dst, (unsigned long long)t);
            if (threadIdx.x == 14) atomicAdd(&a.stats[a.filter_off + FQ_PAIRED_READ_NUMBER], (unsigned long long)t);
            if (threadIdx.x == 15) atomicAdd(&a.stats[a.filter_off + FQ_PAIRED_BASE_LENGTH], (unsigned long long)t);
        }
    }
}

// One warp owns 32 consecutive records.  Copies become piece descriptors; the other records are written into the stream's
// literal bytes with the same writers byte mode uses (reading the raw bytes straight from global memory: they are few).
template <bool PLAIN>
__global__ void __launch_bounds__(kTile) k_emit_pieces(const EmitArgs a, const DevOpts o_in)
{
    DevOpts o = o_in;
    if (PLAIN) {
        o.replace_q = 0;
        o.out_off = o.in_off;
        o.qc_only = 0;
    }
    __shared__ uint32_t s_wsum[8][kTile / 32];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t r = blockIdx.x * kTile + threadIdx.x;
    const LaneItems I = lane_items(a, o, r);
    uint32_t pc[4], lb[4], run_mask[2];
    bool starts[2];
    piece_counts(I, lane, pc, lb, starts, run_mask);
    uint32_t inc[8], val[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        val[k] = k < 4 ? pc[k] : lb[k - 4];
        inc[k] = warp_incl_scan(val[k], lane);
        if (lane == 31) s_wsum[k][wid] = inc[k];
    }
    __syncthreads();
    uint32_t poff[4], loff[4];      // first piece index / literal byte offset of this lane's record in each stream
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        uint32_t before = a.tile_sum[(size_t)(4 + k) * a.n_tiles + blockIdx.x];
        for (uint32_t w = 0; w < wid; ++w) before += s_wsum[k][w];
        const uint32_t x = before + inc[k] - val[k];
        if (k < 4) poff[k] = x; else loff[k - 4] = x;
    }
    if (o.qc_only) return;
    const bool in = r < a.n_rec;
    if (__ballot_sync(0xffffffffu, in) == 0) return;
    const LaneRec &L = I.L;
    if (a.check_ids && in && o.paired) {
        if (pair_ids_differ(a.raw[0] + L.rc[0].hdr, L.rc[0].seq - L.rc[0].hdr - 1, a.raw[1] + L.rc[1].hdr, L.rc[1].seq - L.rc[1].hdr - 1)) {
            atomicOr(&a.info->err, kErrPairId);
            atomicMin(&a.info->err_record, r);
        }
    }
    const int n_mates = o.paired ? 2 : 1;
#pragma unroll 1
    for (int m = 0; m < n_mates; ++m) {
        const Rec rc = m ? L.rc[1] : L.rc[0];
        const uint2 res = m ? L.res[1] : L.res[0];
        const uint32_t cn = m ? L.ccode[1] : L.ccode[0];
        const bool ok = m ? L.valid[1] : L.valid[0];
        const int st = m ? I.stream[1] : I.stream[0];
        const uint32_t nbytes = m ? I.bytes[1] : I.bytes[0];
        const bool is_copy = m ? I.copy[1] : I.copy[0], is_main = m ? I.main[1] : I.main[0], start = m ? starts[1] : starts[0];
        const uint32_t mask = m ? run_mask[1] : run_mask[0];
        // where this mate's piece / literal bytes go: behind mate 1's when both mates feed the same stream
        uint32_t pidx = 0, lpos = 0;
        if (st >= 0) {
            pidx = st == 0 ? poff[0] : st == 1 ? poff[1] : st == 2 ? poff[2] : poff[3];
            lpos = st == 0 ? loff[0] : st == 1 ? loff[1] : st == 2 ? loff[2] : loff[3];
            if (m == 1 && I.stream[0] == st) {
                pidx += starts[0] ? 1u : 0u;
                lpos += I.copy[0] ? 0u : I.bytes[0];
            }
        }
        fq_out_piece *plist = st == 0 ? a.pieces[0] : st == 1 ? a.pieces[1] : st == 2 ? a.pieces[2] : a.pieces[3];
        // ---- copies: one piece per run of consecutive lanes in the main stream, one per record elsewhere
        const bool in_run = st >= 0 && is_main && is_copy;
        uint32_t run_len = nbytes;
        {
            const uint32_t rest = ~(mask >> lane);                         // this lane's run ends in front of the first clear bit above it
            const int last = (int)lane + (rest ? __ffs(rest) - 1 : 32 - (int)lane) - 1;
            const uint32_t end = __shfl_sync(0xffffffffu, rc.hdr + nbytes, in_run ? last : (int)lane);
            if (in_run) run_len = end - rc.hdr;
        }
        if (st >= 0 && is_copy && start) plist[pidx] = fq_out_piece{(uint64_t)rc.hdr, run_len, (uint32_t)m};
        // ---- literal records, four at a time (one per 8-lane group)
        const bool lit = st >= 0 && !is_copy;
        uint8_t *dstp = nullptr;
        if (lit) {
            dstp = (st == 0 ? a.out[0] : st == 1 ? a.out[1] : st == 2 ? a.out[2] : a.out[3]) + lpos;
            plist[pidx] = fq_out_piece{(uint64_t)lpos, nbytes, 2u};
        }
        const uint8_t *src = m ? a.raw[1] : a.raw[0];
        uint32_t todo = __ballot_sync(0xffffffffu, lit);
        const uint32_t sub = lane & 7, grp = lane >> 3;
        while (todo) {
            int j = -1;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int b = todo ? __ffs(todo) - 1 : -1;
                if (b >= 0) todo &= todo - 1;
                if ((int)grp == g) j = b;
            }
            const int js = j < 0 ? 0 : j;
            Rec rj;
            rj.hdr = __shfl_sync(0xffffffffu, rc.hdr, js);
            rj.seq = __shfl_sync(0xffffffffu, rc.seq, js);
            rj.qual = __shfl_sync(0xffffffffu, rc.qual, js);
            rj.len = __shfl_sync(0xffffffffu, rc.len, js);
            const uint32_t ex = __shfl_sync(0xffffffffu, res.x, js), ey = __shfl_sync(0xffffffffu, res.y, js);
            const uint32_t cj = __shfl_sync(0xffffffffu, cn, js);
            const bool tr = __shfl_sync(0xffffffffu, (int)ok, js) != 0;
            const unsigned long long dp = __shfl_sync(0xffffffffu, (unsigned long long)dstp, js);
            if (j < 0) continue;
            uint8_t *const outp = reinterpret_cast<uint8_t *>(dp);
            if (tr) write_trimmed<8>(outp, src, rj, cj, ex, ey & kResLenMask, ey >> kResLenBits, o, sub);
            else write_raw<8>(outp, src, rj, cj, sub);
        }
    }
}

}  // namespace fq
