// fq_frame.cuh -- FASTQ record framing on the device.
//
// Replaces the line scanning half of next_read (fastq.cpp:32-122): the byte scan
// for '\n' (gzgets) and '\r' (strpbrk), four lines per record, |seq| == |qual|.
// HBM-bound byte work: 16-byte vector loads, one warp per 4 KiB chunk, SIMD byte
// compares (__vcmpeq4), warp-shuffle scans.  No shared memory needed.
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

// Pass 1: newline (and CR) count per 4 KiB chunk.  `n` bytes at `raw` (16-byte aligned).
__global__ void __launch_bounds__(256) k_count_lines(const uint8_t *__restrict__ raw, uint64_t n,
                                                     uint32_t *__restrict__ chunk_count, uint32_t n_chunks,
                                                     BatchInfo *info, int mate)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warps_per_grid = (gridDim.x * blockDim.x) >> 5;
    uint32_t cr_total = 0;
    for (uint32_t chunk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; chunk < n_chunks; chunk += warps_per_grid) {
        const uint64_t base = (uint64_t)chunk * kChunkBytes;
        uint32_t cnt = 0;
        if (base + kChunkBytes <= n) {
            const uint4 *p = reinterpret_cast<const uint4 *>(raw + base);
            uint4 v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = ld_stream16(p + k * 32 + lane);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                cnt += __popc(match16(v[k], 0x0a0a0a0au));
                cr_total += __popc(match16(v[k], 0x0d0d0d0du));
            }
        } else {                                       // ragged tail chunk: byte loop
            for (uint64_t i = base + lane; i < n; i += 32) {
                const uint8_t c = raw[i];
                cnt += (c == '\n');
                cr_total += (c == '\r');
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) chunk_count[chunk] = cnt;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) cr_total += __shfl_xor_sync(0xffffffffu, cr_total, o);
    if (lane == 0 && cr_total) atomicAdd(&info->n_cr[mate], cr_total);
}

// Exclusive scan of the chunk counts (single CTA, 1024 threads, sequential tiles).
__global__ void __launch_bounds__(1024) k_scan_chunks(uint32_t *__restrict__ chunk_count, uint32_t n_chunks,
                                                      BatchInfo *info, int mate)
{
    __shared__ uint32_t warp_sum[32];
    __shared__ uint32_t carry;
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_chunks; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n_chunks ? chunk_count[i] : 0;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= (uint32_t)o) x += y;
        }
        if (lane == 31) warp_sum[wid] = x;
        __syncthreads();
        if (wid == 0) {
            uint32_t s = warp_sum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= (uint32_t)o) s += y;
            }
            warp_sum[lane] = s;     // inclusive over warps
        }
        __syncthreads();
        const uint32_t before = carry + (wid ? warp_sum[wid - 1] : 0) + (x - v);
        if (i < n_chunks) chunk_count[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) info->n_lines[mate] = carry;
}

// Pass 2: write the byte offset of every '\n' at its global rank.
__global__ void __launch_bounds__(256) k_scatter_lines(const uint8_t *__restrict__ raw, uint64_t n,
                                                       const uint32_t *__restrict__ chunk_base, uint32_t n_chunks,
                                                       uint32_t *__restrict__ nl_pos)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warps_per_grid = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t chunk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; chunk < n_chunks; chunk += warps_per_grid) {
        const uint64_t base = (uint64_t)chunk * kChunkBytes;
        uint32_t rank = chunk_base[chunk];
        if (base + kChunkBytes <= n) {
            const uint4 *p = reinterpret_cast<const uint4 *>(raw + base);
            uint4 v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = ld_stream16(p + k * 32 + lane);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                uint32_t m = match16(v[k], 0x0a0a0a0au);
                const uint32_t c = __popc(m);
                uint32_t x = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
                    if (lane >= (uint32_t)o) x += y;
                }
                uint32_t r = rank + x - c;
                const uint32_t pos0 = (uint32_t)(base + (uint64_t)(k * 32 + lane) * 16);
                while (m) {
                    const int b = __ffs(m) - 1;
                    nl_pos[r++] = pos0 + b;
                    m &= m - 1;
                }
                rank += __shfl_sync(0xffffffffu, x, 31);
            }
        } else {
            for (uint64_t i0 = base; i0 < n; i0 += 32) {
                const uint64_t i = i0 + lane;
                const bool is_nl = i < n && raw[i] == '\n';
                const uint32_t m = __ballot_sync(0xffffffffu, is_nl);
                if (is_nl) nl_pos[rank + __popc(m & ((1u << lane) - 1))] = (uint32_t)i;
                rank += __popc(m);
            }
        }
    }
}

// Record descriptors from 4 consecutive line ends.  One thread per record.
// Grammar per fastq.cpp: content of a line ends at '\r' when it is followed by the
// '\n' (CRLF input); |seq| must equal |qual| (fastq.cpp:118-122).
__global__ void __launch_bounds__(256) k_build_records(const uint8_t *__restrict__ raw,
                                                       const uint32_t *__restrict__ nl_pos, uint32_t n_rec,
                                                       Rec *__restrict__ rec, uint8_t *__restrict__ canon, BatchInfo *info, int mate)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t len = 0, cr = 0;
    bool bad = false;
    if (r < n_rec) {
        const uint32_t p0 = nl_pos[4 * r], p1 = nl_pos[4 * r + 1], p2 = nl_pos[4 * r + 2], p3 = nl_pos[4 * r + 3];
        const uint32_t hdr = r ? nl_pos[4 * r - 1] + 1 : 0;
        const uint32_t seq = p0 + 1, plus = p1 + 1, qual = p2 + 1;
        const uint32_t c0 = (p0 > hdr && raw[p0 - 1] == '\r');
        const uint32_t c1 = (p1 > seq && raw[p1 - 1] == '\r');
        const uint32_t c2 = (p2 > plus && raw[p2 - 1] == '\r');
        const uint32_t c3 = (p3 > qual && raw[p3 - 1] == '\r');
        len = p1 - seq - c1;
        const uint32_t qlen = p3 - qual - c3;
        bad = (len != qlen);
        cr = c0 + c1 + c2 + c3;
        rec[r] = Rec{hdr, seq, qual, len};
        // canonical record: LF line ends and a bare "+" line, i.e. the raw bytes ARE what write_read
        // (fastq.cpp:127-138) prints for an untouched read, so emission can be a block copy
        canon[r] = (uint8_t)(cr == 0 && p2 == plus + 1 && raw[plus] == '+');
    }
    // block-level reductions: max length, CR count, first bad record
    uint32_t m = len;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        cr += __shfl_xor_sync(0xffffffffu, cr, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (m) atomicMax(&info->max_len[mate], m);
        if (cr) atomicAdd(&info->n_cr_eol[mate], cr);
    }
    if (bad) {
        atomicOr(&info->err, kErrLenMismatch);
        atomicMin(&info->err_record, r);
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
__device__ __forceinline__ uint32_t id_length(const uint8_t *h, uint32_t n)
{
    uint32_t loc = 0;
    while (loc < n && h[loc] != ' ') ++loc;
    if (loc > 1 && h[loc - 1] >= '0' && h[loc - 1] <= '9' && (h[loc - 2] == '.' || h[loc - 2] == '/')) loc -= 2;
    return loc;
}

__global__ void __launch_bounds__(256) k_check_pair_ids(const uint8_t *__restrict__ raw1, const Rec *__restrict__ rec1,
                                                        const uint8_t *__restrict__ raw2, const Rec *__restrict__ rec2,
                                                        uint32_t n_rec, BatchInfo *info)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    const Rec a = rec1[r], b = rec2[r];
    uint32_t na = a.seq - a.hdr - 1, nb = b.seq - b.hdr - 1;
    const uint8_t *ha = raw1 + a.hdr, *hb = raw2 + b.hdr;
    if (na && ha[na - 1] == '\r') --na;
    if (nb && hb[nb - 1] == '\r') --nb;
    const uint32_t la = id_length(ha, na), lb = id_length(hb, nb);
    bool same = (la == lb);
    for (uint32_t i = 0; same && i < la; ++i) same = (ha[i] == hb[i]);
    if (!same) {
        atomicOr(&info->err, kErrPairId);
        atomicMin(&info->err_record, r);
    }
}

}  // namespace fq
