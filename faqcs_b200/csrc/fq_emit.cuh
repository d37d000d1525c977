// fq_emit.cuh -- pair routing and in-order stream compaction
// (FaQCs.cpp:296-361,431-496 paired; :634-659,696-720 unpaired; write_read, fastq.cpp:127-138).
//
// Three steps: k_route sums the emitted bytes of each 256-record tile for the
// four streams, k_scan_tiles turns the tile sums into tile bases, k_emit
// recomputes the per-record sizes, scans them inside the tile and has one warp
// per record copy `def\nseq\n+\nqual\n` to its final position, applying the
// mutations trim_read leaves in the read (terminal-N quality mask, G->N
// replacement, quality re-encoding).  Discarded reads are copied raw.
#pragma once
#include "fq_common.cuh"
#include "fq_trim.cuh"

namespace fq {

constexpr uint32_t kTile = 256;     // records per route/emit tile

struct EmitArgs {
    const uint8_t *raw[2];
    const Rec *rec[2];
    const uint2 *res[2];
    const uint8_t *canon[2];     // 1 = LF line ends + bare '+' line (raw bytes == write_read output)
    uint32_t n_rec;
    uint32_t n_tiles;
    uint32_t *tile_sum;          // [4][n_tiles] -> exclusive bases after k_scan_tiles (u32: < 4 GiB per stream per batch)
    uint8_t *out[4];
    BatchInfo *info;
    unsigned long long *stats;   // global stats block (PAIRED_* counters are added here too)
    size_t filter_off;
};

__device__ __forceinline__ uint32_t header_len(const uint8_t *raw, const Rec &rc, bool canon = false)
{
    uint32_t n = rc.seq - rc.hdr - 1;
    if (!canon && n && raw[rc.hdr + n - 1] == '\r') --n;
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
        const uint32_t hl = header_len(a.raw[m], rc, a.canon[m][r] != 0);
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

// Exclusive scan of the four tile-sum rows; totals go to info->out_bytes.  One CTA, warp w scans stream w.
__global__ void __launch_bounds__(128) k_scan_tiles(uint32_t *tile_sum, uint32_t n_tiles, BatchInfo *info)
{
    const uint32_t lane = threadIdx.x & 31, s = threadIdx.x >> 5;
    uint32_t *row = tile_sum + (size_t)s * n_tiles;
    unsigned long long carry = 0;
    for (uint32_t base = 0; base < n_tiles; base += 32) {
        const uint32_t i = base + lane;
        const uint32_t v = i < n_tiles ? row[i] : 0;
        uint32_t x = v;
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, k);
            if (lane >= (uint32_t)k) x += y;
        }
        if (i < n_tiles) row[i] = (uint32_t)(carry + x - v);
        carry += __shfl_sync(0xffffffffu, x, 31);
    }
    if (lane == 0) info->out_bytes[s] = carry;
}

// Warp-cooperative copy of n bytes between arbitrarily aligned addresses.  The body is written
// as 16-byte aligned stores; each store gathers its bytes from five aligned 32-bit source words
// with funnel shifts (the source is L1/L2 resident raw input, neighbouring lanes share words).
__device__ __forceinline__ uint4 gather16(const uint8_t *sc)
{
    // 16 bytes starting at the arbitrarily aligned address sc, from five aligned 32-bit words
    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(sc) & 3u) * 8u;
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(sc) & ~(uintptr_t)3);
    const uint32_t w0 = __ldg(sw), w1 = __ldg(sw + 1), w2 = __ldg(sw + 2), w3 = __ldg(sw + 3);
    const uint32_t w4 = sh ? __ldg(sw + 4) : 0u;          // only bytes below src + n are ever consumed from it
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

// Emit one surviving read: def \n seq \n + \n qual \n (write_read, fastq.cpp:127-138) with the
// mutations trim_read leaves behind.  `plain`: the record is canonical, untrimmed and untouched,
// so the output is its raw bytes.
template <uint32_t W = 32>
__device__ __forceinline__ void write_trimmed(uint8_t *dst, const uint8_t *raw, const Rec &rc, bool canon, uint32_t lo, uint32_t wl,
                                              uint32_t flags, const DevOpts &o, uint32_t lane)
{
    const uint32_t hl = header_len(raw, rc, canon);
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
        if (o.replace_q == 0) copy_span<W>(dst + s0, sp + lo, wl, lane);
        else {
            for (uint32_t i = lane; i < wl; i += W) {          // G -> N below --replace_to_N_q (trim.cpp:390-403)
                const uint32_t p = lo + i;
                uint32_t ch = sp[p];
                if (ch == 'G') {
                    const int qc = (p < lead || p >= trail) ? o.in_off : (int)qp[p];
                    if (max(0, qc - o.in_off) < (int)o.replace_q) ch = 'N';
                }
                dst[s0 + i] = (uint8_t)ch;
            }
        }
    }
    if (lane < 3) dst[s1 + lane] = lane == 1 ? '+' : '\n';
    if (!masked && !requal) copy_span<W>(dst + q0, raw + rc.qual + lo, wl, lane);
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
__device__ __forceinline__ void write_raw(uint8_t *dst, const uint8_t *raw, const Rec &rc, bool canon, uint32_t lane)
{
    const uint32_t hl = header_len(raw, rc, canon);
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
    bool canon[2], valid[2], plain[2];
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
        L.res[m] = a.res[m][r];
        L.canon[m] = a.canon[m][r] != 0;
        const uint32_t fl = L.res[m].y >> kResLenBits, wl = L.res[m].y & kResLenMask;
        L.valid[m] = (fl & FQ_RR_VALID) != 0;
        L.tsize[m] = header_len(a.raw[m], L.rc[m], L.canon[m]) + 2 * wl + 5;
        // untouched canonical record: the emitted bytes are the raw bytes
        L.plain[m] = L.valid[m] && L.canon[m] && L.res[m].x == 0 && wl == L.rc[m].len && !(fl & kFlagMasked) && !requal && o.replace_q == 0;
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

#ifndef FQ_EMIT_MIN_CTAS
#define FQ_EMIT_MIN_CTAS 4
#endif
__global__ void __launch_bounds__(kTile, FQ_EMIT_MIN_CTAS) k_emit(const EmitArgs a, const DevOpts o)
{
    __shared__ uint32_t s_wsum[4][kTile / 32];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t r = blockIdx.x * kTile + threadIdx.x;
    uint32_t sz[4], wl[2];
    bool valid[2];
    route_sizes(a, o, r, sz, valid, wl);
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

    if (o.paired) {
        const bool both = in && L.valid[0] && L.valid[1];
        // block copies: runs of untouched pairs, mate by mate
        copy_runs(__ballot_sync(0xffffffffu, both && L.plain[0]), a.out[0], a.raw[0], L.rc[0].hdr, L.tsize[0], off[0], lane);
        copy_runs(__ballot_sync(0xffffffffu, both && L.plain[1]), a.out[1], a.raw[1], L.rc[1].hdr, L.tsize[1], off[1], lane);
        // everything else, record by record
        const bool single0 = in && ((both && !L.plain[0]) || (!both && L.valid[0]));
        const bool single1 = in && ((both && !L.plain[1]) || (!both && L.valid[1]));
        const bool disc = in && o.discard && !both && (!L.valid[0] || !L.valid[1]);
        uint32_t todo = __ballot_sync(0xffffffffu, single0 || single1 || disc);
        const uint32_t sub = lane & 7, grp = lane >> 3;
        while (todo) {
            // four records at a time, one per 8-lane group: each record is a short chain of dependent
            // load -> store round trips, so concurrency across records is what hides the latency
            int j = -1;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int b = todo ? __ffs(todo) - 1 : -1;
                if (b >= 0) todo &= todo - 1;
                if ((int)grp == g) j = b;
            }
            const int js = j < 0 ? 0 : j;
            Rec rc[2];
            uint2 e[2];
            bool cn[2], v[2], pl[2];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                rc[m].hdr = __shfl_sync(0xffffffffu, L.rc[m].hdr, js);
                rc[m].seq = __shfl_sync(0xffffffffu, L.rc[m].seq, js);
                rc[m].qual = __shfl_sync(0xffffffffu, L.rc[m].qual, js);
                rc[m].len = __shfl_sync(0xffffffffu, L.rc[m].len, js);
                e[m].x = __shfl_sync(0xffffffffu, L.res[m].x, js);
                e[m].y = __shfl_sync(0xffffffffu, L.res[m].y, js);
                cn[m] = __shfl_sync(0xffffffffu, (int)L.canon[m], js) != 0;
                v[m] = __shfl_sync(0xffffffffu, (int)L.valid[m], js) != 0;
                pl[m] = __shfl_sync(0xffffffffu, (int)L.plain[m], js) != 0;
            }
            uint32_t o4[4];
#pragma unroll
            for (int s = 0; s < 4; ++s) o4[s] = __shfl_sync(0xffffffffu, off[s], js);
            if (j < 0) continue;
            if (v[0] && v[1]) {
                if (!pl[0]) write_trimmed<8>(a.out[0] + o4[0], a.raw[0], rc[0], cn[0], e[0].x, e[0].y & kResLenMask, e[0].y >> kResLenBits, o, sub);
                if (!pl[1]) write_trimmed<8>(a.out[1] + o4[1], a.raw[1], rc[1], cn[1], e[1].x, e[1].y & kResLenMask, e[1].y >> kResLenBits, o, sub);
            } else {
                if (v[0]) write_trimmed<8>(a.out[2] + o4[2], a.raw[0], rc[0], cn[0], e[0].x, e[0].y & kResLenMask, e[0].y >> kResLenBits, o, sub);
                else if (v[1]) write_trimmed<8>(a.out[2] + o4[2], a.raw[1], rc[1], cn[1], e[1].x, e[1].y & kResLenMask, e[1].y >> kResLenBits, o, sub);
                if (o.discard) {
                    uint32_t d = o4[3];
                    if (!v[0]) { write_raw<8>(a.out[3] + d, a.raw[0], rc[0], cn[0], sub); d += header_len(a.raw[0], rc[0], cn[0]) + 2 * rc[0].len + 5; }
                    if (!v[1]) write_raw<8>(a.out[3] + d, a.raw[1], rc[1], cn[1], sub);
                }
            }
        }
    } else {
        copy_runs(__ballot_sync(0xffffffffu, in && L.plain[0]), a.out[2], a.raw[0], L.rc[0].hdr, L.tsize[0], off[2], lane);
        const bool single = in && L.valid[0] && !L.plain[0];
        const bool disc = in && o.discard && !L.valid[0];
        uint32_t todo = __ballot_sync(0xffffffffu, single || disc);
        while (todo) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            Rec rc;
            rc.hdr = __shfl_sync(0xffffffffu, L.rc[0].hdr, j);
            rc.seq = __shfl_sync(0xffffffffu, L.rc[0].seq, j);
            rc.qual = __shfl_sync(0xffffffffu, L.rc[0].qual, j);
            rc.len = __shfl_sync(0xffffffffu, L.rc[0].len, j);
            const uint32_t ex = __shfl_sync(0xffffffffu, L.res[0].x, j), ey = __shfl_sync(0xffffffffu, L.res[0].y, j);
            const bool cn = __shfl_sync(0xffffffffu, (int)L.canon[0], j) != 0, v = __shfl_sync(0xffffffffu, (int)L.valid[0], j) != 0;
            const uint32_t o2 = __shfl_sync(0xffffffffu, off[2], j), o3 = __shfl_sync(0xffffffffu, off[3], j);
            if (v) write_trimmed(a.out[2] + o2, a.raw[0], rc, cn, ex, ey & kResLenMask, ey >> kResLenBits, o, lane);
            else write_raw(a.out[3] + o3, a.raw[0], rc, cn, lane);
        }
    }
}

}  // namespace fq
