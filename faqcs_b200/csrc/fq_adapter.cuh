// fq_adapter.cuh -- adapter / artifact matching (trim_adapters_and_phiX, trim.cpp:961-1142;
// SeqOverlap::align_smith_waterman, seq_overlap.cpp:46-370; find_mask_range, trim.cpp:1144-1189).
//
// The reference's "Smith-Waterman" has gaps compiled out, so every diagonal of
// the read x adapter matrix is an independent recurrence
//     M_k = max(M_{k-1}, 0) + (match ? +1 : -1),   start_k = (M_{k-1} < 0) ? i_k : start_{k-1}
// and the reported alignment is the cell with the largest M, ties going to the
// LAST cell in (i outer, j inner) order (seq_overlap.cpp:342 uses >=).  Here one
// warp owns a read, lanes own diagonals (32 at a time), the adapter set lives in
// shared memory for the whole CTA, and the winner is picked with a warp max over
// the key (M, i, j).  Integer ALU bound, not HBM bound.
#pragma once
#include "fq_common.cuh"
#include "fq_trim.cuh"

namespace fq {

struct AdapterSet {
    const uint8_t *codes;     // concatenated NA bit codes (seq_overlap.h:133-150)
    const uint32_t *offset;   // [n + 1]
    uint32_t n;
    uint32_t total;           // bytes in codes
    uint32_t any_bits;        // per-adapter OR of codes is in or_bits
    const uint8_t *or_bits;   // [n]
};

struct AdapterArgs {
    const uint8_t *raw[2];
    const Rec *rec[2];
    uint2 *adp[2];
    int32_t *adp_best[2];
    uint32_t n_rec, n_mates;
    uint32_t max_len;            // longest read in the batch (sizes the per-warp buffers)
    unsigned long long *stats;
    StatsLayout L;
    BatchInfo *info;
    unsigned long long first_index;   // global index of record 0 (Q3 emulation)
    unsigned long long end_index;     // first_index + n_rec if this is the final batch, else ~0
    uint32_t use_planes;              // bit-plane prefilter enabled (planes fit in shared memory)
    uint32_t plane_words;             // total words of the adapter bit planes (incl. padding)
    uint32_t has_gap;                 // some adapter contains '-'
    uint32_t rpad;                    // zero words on both sides of every read bit plane (longest adapter in words + 1)
    uint32_t sweep;                   // segment sweep enabled (needs use_planes)
    uint32_t n_seg;                   // number of 32-base segments of the adapters of up to kSweepMaxLen bases
    uint32_t sweep_pure;              // every swept adapter is plain A/C/G/T
};

// ---- segment sweep ------------------------------------------------------------------------------------
// Every adapter of up to 96 bases is cut into segments of (up to) 32 bases.  A segment is four position masks: x: the base
// can be a purine (A, G), y: a pyrimidine (C, T), u: strong (C, G), v: weak (A, T); IUPAC unions, N and '-' set all that
// apply.  Two bases match iff they agree in the purine/pyrimidine AND in the strong/weak split, so for 32 read positions
// with masks (wx, wy, wu, wv) the match mask of the segment is ((x & wx) | (y & wy)) & ((u & wu) | (v & wv)) -- exact for
// A, C, G, T, N and never missing a match for the other codes.
// If an adapter of T bases has `threshold` matching positions on some diagonal, one of its segments (len_k bases) has at
// least threshold * len_k / T of them against the 32 read positions it faces there (pigeonhole).  The sweep checks that
// necessary condition for ALL segments at once: one lane owns one segment (masks and bound in registers), the warp walks
// the windows of 32 read positions one position at a time (four funnel shifts of warp-uniform words), and each step
// costs every lane five logic operations, a popcount and a max.  Only the adapters it flags are looked at any further:
// one-segment adapters go straight to the exact alignment (their count was exact), longer ones through the exact
// per-diagonal plane count first.
constexpr uint32_t kSweepMaxLen = 96;
struct SegInfo {
    uint16_t owner;     // adapter index
    uint8_t len;        // bases in this segment (an adapter is cut into ceil(T / 32) segments of equal length, +-1)
    uint8_t total;      // bases in the adapter
    uint8_t start;      // first adapter base of the segment
    uint8_t bound_full; // matches the segment needs when the threshold is taken from the whole adapter (the read is not shorter)
    uint16_t pad;
};
#ifndef FQ_ADAPTER_MIN_CTAS
#define FQ_ADAPTER_MIN_CTAS 3
#endif
constexpr int kSweepWindows = 6;        // windows of 32 diagonals held in registers per pass over the segments

// seq_overlap.cpp:372-411; 0xff = unknown base (the reference throws)
__device__ __forceinline__ uint32_t na_to_bits(uint32_t c)
{
    switch (c | 0x20u) {
        case 'a': return 1;  case 'c': return 2;  case 'g': return 4;  case 't': return 8;
        case 'm': return 3;  case 'r': return 5;  case 's': return 6;  case 'v': return 7;
        case 'w': return 9;  case 'y': return 10; case 'h': return 11; case 'k': return 12;
        case 'd': return 13; case 'b': return 14; case 'n': return 15;
        default: return c == '-' ? 16u : 0xffu;
    }
}

// Exact alignment of the warp's read (codes in s_read) against one adapter: the cell with the largest M,
// ties to the last cell in (i outer, j inner) order.  Returns the score; start/stop only change if score > 0.
__device__ __forceinline__ int exact_align(const uint8_t *s_read, uint32_t L, const uint8_t *t, uint32_t T, uint32_t lane, int &st_start, int &st_stop)
{
    int bM = 0, bi = 0, bj = 0, bst = 0;      // per-lane best over the diagonals this lane walks: (M, i, jj) lexicographic max
    for (int d0 = -(int)(L - 1); d0 <= (int)T - 1; d0 += 32) {
        const int d = d0 + (int)lane;
        const int i_lo = max(0, -(d0 + 31));
        const int i_hi = min((int)L - 1, (int)T - 1 - d0);
        int M = 0, st = i_lo, gM = 0, gi = 0, gst = 0;
        for (int i = i_lo; i <= i_hi; ++i) {
            const int jj = i + d;
            const uint32_t qi = s_read[i];
            if ((unsigned)jj < T) {
                const bool match = (qi & t[jj]) != 0;
                st = (M < 0) ? i : st;
                M = max(M, 0) + (match ? 1 : -1);
                if (M >= gM && M > 0) { gM = M; gi = i; gst = st; }
            } else {
                M = 0;
                st = i + 1;
            }
        }
        const int gj = gi + d;
        if (gM > bM || (gM == bM && gM > 0 && (gi > bi || (gi == bi && gj > bj)))) { bM = gM; bi = gi; bj = gj; bst = gst; }
    }
    const unsigned long long key = bM > 0 ? (((unsigned long long)bM << 48) | ((unsigned long long)bi << 24) | (unsigned long long)bj) : 0ull;
    unsigned long long kmax = key;
#pragma unroll
    for (int k = 16; k; k >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, kmax, k);
        kmax = other > kmax ? other : kmax;
    }
    if (!kmax) return 0;
    const uint32_t win = __ballot_sync(0xffffffffu, key == kmax);
    st_stop = (int)((kmax >> 24) & 0xffffffu);
    st_start = __shfl_sync(0xffffffffu, bst, __ffs(win) - 1);
    return (int)(kmax >> 48);
}

// One pass of the segment sweep: NW windows of 32 diagonals (lane = diagonal) against every segment.
struct SweepCtx {
    const uint32_t *rx;         // read masks x, y, u, v: 4 planes of `rstride` words, zero outside the read
    uint32_t rstride;
    const uint4 *seg;           // [n_seg] {x, y, u, v}
    const uint4 *segp;          // [n_seg] {x, valid, u, bound from the whole adapter}: plain A/C/G/T segments
    const SegInfo *info;
    const uint8_t *wbound;      // per-read bounds (read shorter than an adapter)
    uint32_t *cand;
    uint32_t n_seg, lane;
    int S, n_windows;
    bool full, pure;
};

template <int NW>
__device__ __forceinline__ void sweep_windows(const SweepCtx &c, int w0, uint32_t *s_segflag)
{
    uint32_t wx[NW], wy[NW], wu[NW], wv[NW];
#pragma unroll
    for (int i = 0; i < NW; ++i) {
        const int p0 = -c.S + 32 * (w0 + i) + (int)c.lane;
        const uint32_t sh = (uint32_t)p0 & 31u;
        const uint32_t *p = c.rx + (p0 >> 5);
        wx[i] = wy[i] = wu[i] = wv[i] = 0;
        if (w0 + i < c.n_windows) {
            wx[i] = __funnelshift_r(p[0], p[1], sh);
            wy[i] = __funnelshift_r(p[c.rstride], p[c.rstride + 1], sh);
            wu[i] = __funnelshift_r(p[2 * c.rstride], p[2 * c.rstride + 1], sh);
            wv[i] = __funnelshift_r(p[3 * c.rstride], p[3 * c.rstride + 1], sh);
        }
    }
    if (c.full && c.pure) {
        // plain A/C/G/T segments, bounds known up front: y = ~x and v = ~u inside the segment, so each pair of planes is a bit
        // select; a lane keeps one bit per segment ("no diagonal of mine reached the bound"), merged once per 32 segments
        for (uint32_t k0 = 0; k0 < c.n_seg; k0 += 32) {
            const uint32_t n_here = min(32u, c.n_seg - k0);
            uint32_t miss = 0;
#pragma unroll 2
            for (uint32_t j = 0; j < n_here; ++j) {
                const uint4 sm = c.segp[k0 + j];
                uint32_t best = 0;
#pragma unroll
                for (int i = 0; i < NW; ++i)
                    best = max(best, (uint32_t)__popc(((sm.x & wx[i]) | (~sm.x & wy[i])) & ((sm.z & wu[i]) | (~sm.z & wv[i])) & sm.y));
                miss = __funnelshift_l(best - sm.w, miss, 1);        // (miss << 1) | sign(best - bound)
            }
            uint32_t hit = ~miss & (n_here == 32 ? 0xffffffffu : ((1u << n_here) - 1u));     // bit n_here-1-j: segment k0 + j
            hit = __reduce_or_sync(0xffffffffu, hit);
            if (c.lane == 0 && hit) s_segflag[k0 >> 5] |= hit;
        }
        return;
    }
    for (uint32_t k = 0; k < c.n_seg; ++k) {
        const uint4 sm = c.seg[k];
        const SegInfo si = c.info[k];
        const uint32_t bound = c.full ? (uint32_t)si.bound_full : (uint32_t)c.wbound[k];
        const uint32_t va = sm.x | sm.y;
        uint32_t best = 0;
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            uint32_t m;
            if (c.pure) m = ((sm.x & wx[i]) | (~sm.x & wy[i])) & ((sm.z & wu[i]) | (~sm.z & wv[i])) & va;
            else m = ((sm.x & wx[i]) | (sm.y & wy[i])) & ((sm.z & wu[i]) | (sm.w & wv[i]));
            best = max(best, (uint32_t)__popc(m));
        }
        if (__any_sync(0xffffffffu, best >= bound) && c.lane == 0) c.cand[si.owner >> 5] |= 1u << (si.owner & 31);
    }
}

// Shared memory: [offsets n+1][plane offsets n+1][adapter codes][adapter bit planes][segment masks][segment infos]
//                [per warp: read codes, mask words, read planes (5 + 4 sweep masks), candidate bits]
//
// Prefilter (exact-safe): an adapter can only pass the threshold test num_match >= threshold (trim.cpp:1024-1027)
// if SOME diagonal holds at least `threshold` matching positions, because num_match counts the matches inside
// one diagonal segment.  Matches per diagonal are popcounts over base bit planes (A,C,G,T,gap: a read base
// matches an adapter base iff their IUPAC bit sets intersect), 32 diagonals per round, one per lane.  Only
// adapters that survive get the exact alignment; the stale range an all-mismatch adapter inherits (Q5) is
// produced lazily by aligning the nearest earlier adapter that shares a base with the read.
__global__ void __launch_bounds__(256, FQ_ADAPTER_MIN_CTAS) k_adapter(const AdapterArgs a, const DevOpts o, const AdapterSet A)
{
    extern __shared__ __align__(16) uint32_t smem_u32[];
    uint32_t *s_off = smem_u32;
    uint32_t *s_poff = s_off + A.n + 1;
    uint8_t *s_codes = reinterpret_cast<uint8_t *>(s_poff + A.n + 1);
    const uint32_t codes_pad = (A.total + 3) & ~3u;
    const uint32_t read_pad = (a.max_len + 3) & ~3u;
    const uint32_t mask_words = (a.max_len + 31) >> 5;         // also the number of words of a read bit plane
    const uint32_t rstride = mask_words + 2 * a.rpad;          // words of one (zero padded) read bit plane
    uint32_t *s_planes = reinterpret_cast<uint32_t *>(s_codes + codes_pad);
    const uint32_t warp_in_cta = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cand_words = (A.n + 31) >> 5;
    const uint32_t bound_words = (a.n_seg + 3) >> 2;           // per-read segment bounds (bytes) of a read shorter than an adapter
    const uint32_t per_warp_words = read_pad / 4 + mask_words + (a.use_planes ? 5 * rstride : 0) + (a.sweep ? 4 * rstride + cand_words + bound_words : 0);
    // [n_seg] masks {x, y, u, v}, 16-byte aligned (an offset from the array keeps the pointer in the shared address space)
    uint4 *s_seg = reinterpret_cast<uint4 *>(smem_u32 + (((uint32_t)(s_planes + a.plane_words - smem_u32) + 3u) & ~3u));
    uint4 *s_segp = s_seg + a.n_seg;                                                  // [n_seg] {x, valid, u, whole-adapter bound}
    SegInfo *s_seginfo = reinterpret_cast<SegInfo *>(s_segp + a.n_seg);               // [n_seg]
    uint32_t *s_unswept = reinterpret_cast<uint32_t *>(s_seginfo + a.n_seg);          // [cand_words] adapters the sweep does not cover (longer than its limit)
    uint32_t *s_meta = s_unswept + cand_words;                 // [2] sweep origin S (diagonals start S bases left of the read), longest swept adapter
    uint32_t *warp_base = (a.sweep ? s_meta + 2 : s_planes + a.plane_words) + (size_t)warp_in_cta * per_warp_words;
    uint8_t *s_read = reinterpret_cast<uint8_t *>(warp_base);
    uint32_t *s_mask = warp_base + read_pad / 4;
    uint32_t *s_rp = s_mask + mask_words;                      // read planes [5][rstride], data at word offset rpad
    uint32_t *s_rxy = s_rp + 5 * rstride;                      // sweep masks of the read x, y, u, v: [4][rstride]
    uint32_t *s_cand = s_rxy + 4 * rstride;                    // adapters flagged by the sweep
    uint8_t *s_wbound = reinterpret_cast<uint8_t *>(s_cand + cand_words);
    uint32_t *s_segflag = s_cand + cand_words;                 // plain segments with whole-adapter bounds: flagged segments (same words)

    for (uint32_t i = threadIdx.x; i <= A.n; i += blockDim.x) s_off[i] = A.offset[i];
    for (uint32_t i = threadIdx.x; i < A.total; i += blockDim.x) s_codes[i] = A.codes[i];
    __syncthreads();
    if (a.use_planes) {
        // plane block of adapter j: 5 planes x Wt_j words; offsets by a serial prefix (n is small)
        if (threadIdx.x == 0) {
            uint32_t acc = 0;
            for (uint32_t j = 0; j < A.n; ++j) {
                s_poff[j] = acc;
                acc += 5 * (((s_off[j + 1] - s_off[j]) + 31) / 32);
            }
            s_poff[A.n] = acc;
        }
        for (uint32_t i = threadIdx.x; i < a.plane_words; i += blockDim.x) s_planes[i] = 0;
        for (uint32_t i = lane; i < 5 * rstride; i += 32) s_rp[i] = 0;       // the padding stays zero for the whole kernel
        if (a.sweep) {
            for (uint32_t i = lane; i < 4 * rstride; i += 32) s_rxy[i] = 0;
            if (threadIdx.x == 0) {                // segment -> (adapter, part): serial walk, the set is small
                uint32_t k = 0, origin = 0, longest = 0;
                for (uint32_t w = 0; w < cand_words; ++w) s_unswept[w] = 0;
                for (uint32_t j = 0; j < A.n; ++j) {
                    const uint32_t T = s_off[j + 1] - s_off[j];
                    if (T > kSweepMaxLen) s_unswept[j >> 5] |= 1u << (j & 31);
                    if (T == 0 || T > kSweepMaxLen) continue;
                    longest = max(longest, T);
                    // equal parts: a short tail segment would reach its (tiny) bound on almost every read
                    const uint32_t parts = (T + 31) / 32, base = T / parts, rem = T % parts;
                    const int threshold = __float2int_rz(__fmul_rn(o.match_rate, (float)T));
                    uint32_t start = 0;
                    for (uint32_t i = 0; i < parts; ++i) {
                        const uint32_t len = base + (i < rem ? 1u : 0u);
                        const uint32_t bound = threshold <= 0 ? 0u : ((uint32_t)threshold * len + T - 1) / T;      // ceil(threshold * len / T)
                        s_seginfo[k++] = SegInfo{(uint16_t)j, (uint8_t)len, (uint8_t)T, (uint8_t)start, (uint8_t)bound, 0};
                        origin = max(origin, len - min(len, bound));      // a diagonal further left cannot hold `bound` matches
                        start += len;
                    }
                }
                s_meta[0] = min(origin, 32u);
                s_meta[1] = longest;
            }
        }
        __syncthreads();
        if (a.sweep) {
            for (uint32_t k = threadIdx.x; k < a.n_seg; k += blockDim.x) {
                const SegInfo si = s_seginfo[k];
                const uint8_t *codes = s_codes + s_off[si.owner] + si.start;
                uint32_t m[4] = {0, 0, 0, 0};
                for (uint32_t p = 0; p < si.len; ++p) {
                    const uint32_t code = codes[p], bit = 1u << p;
                    if (code & (1u | 4u | 16u)) m[0] |= bit;
                    if (code & (2u | 8u | 16u)) m[1] |= bit;
                    if (code & (2u | 4u | 16u)) m[2] |= bit;
                    if (code & (1u | 8u | 16u)) m[3] |= bit;
                }
                s_seg[k] = make_uint4(m[0], m[1], m[2], m[3]);
                s_segp[k] = make_uint4(m[0], m[0] | m[1], m[2], si.bound_full);
            }
        }
        __syncthreads();
        for (uint32_t j = 0; j < A.n; ++j) {
            const uint32_t T = s_off[j + 1] - s_off[j], stride = (T + 31) / 32;
            for (uint32_t p = threadIdx.x; p < T; p += blockDim.x) {
                const uint32_t code = s_codes[s_off[j] + p];
#pragma unroll
                for (int b = 0; b < 5; ++b)
                    if ((code >> b) & 1u) atomicOr(&s_planes[s_poff[j] + b * stride + (p >> 5)], 1u << (p & 31));
            }
        }
        __syncthreads();
    }

    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t total = a.n_rec * a.n_mates;

    for (uint32_t g = warp_global; g < total; g += n_warps) {
        const uint32_t mate = g >= a.n_rec ? 1 : 0;
        const uint32_t r = g - mate * a.n_rec;
        const Rec rc = a.rec[mate][r];
        const uint32_t L = rc.len;
        const uint8_t *sp = a.raw[mate] + rc.seq;
        const uint32_t read_words = (L + 31) >> 5;

        // pack_query (seq_overlap.h:370-411) + bit planes of the read
        bool unknown = false;
        uint32_t read_or = 0;
        for (uint32_t b0 = 0; b0 < L; b0 += 32) {
            const uint32_t p = b0 + lane;
            uint32_t code = 0;
            if (p < L) {
                code = na_to_bits(sp[p]);
                unknown |= (code == 0xffu);
                s_read[p] = (uint8_t)code;
                if (code == 0xffu) code = 0;
            }
            read_or |= code;
            if (a.use_planes) {
#pragma unroll
                for (int b = 0; b < 5; ++b) {
                    const uint32_t m = __ballot_sync(0xffffffffu, (code >> b) & 1u);
                    if (lane == 0) s_rp[b * rstride + a.rpad + (b0 >> 5)] = m;
                }
                if (a.sweep) {
                    const uint32_t mx = __ballot_sync(0xffffffffu, (code & (1u | 4u | 16u)) != 0), my = __ballot_sync(0xffffffffu, (code & (2u | 8u | 16u)) != 0);
                    const uint32_t mu = __ballot_sync(0xffffffffu, (code & (2u | 4u | 16u)) != 0), mv = __ballot_sync(0xffffffffu, (code & (1u | 8u | 16u)) != 0);
                    if (lane == 0) {
                        uint32_t *w = s_rxy + a.rpad + (b0 >> 5);
                        w[0] = mx; w[rstride] = my; w[2 * rstride] = mu; w[3 * rstride] = mv;
                    }
                }
            }
        }
        if (a.use_planes) {     // a shorter read after a longer one: clear the plane words it does not own
            for (uint32_t i = lane; i < 5 * (mask_words - read_words); i += 32) {
                const uint32_t b = i / (mask_words - read_words), w = read_words + i % (mask_words - read_words);
                s_rp[b * rstride + a.rpad + w] = 0;
            }
            if (a.sweep)
                for (uint32_t i = lane; i < 4 * (mask_words - read_words); i += 32) {
                    const uint32_t b = i / (mask_words - read_words), w = read_words + i % (mask_words - read_words);
                    s_rxy[b * rstride + a.rpad + w] = 0;
                }
        }
        for (uint32_t w = lane; w < mask_words; w += 32) s_mask[w] = 0;     // 1 = masked
        if (__any_sync(0xffffffffu, unknown)) {
            if (lane == 0) { atomicOr(&a.info->err, kErrUnknownBase); atomicMin(&a.info->err_record, r); }
        }
        read_or = __reduce_or_sync(0xffffffffu, read_or);
        __syncwarp();

        // Q3: which length feeds the match threshold (trim.cpp:985,996-1008,1074-1082)
        uint32_t thr_len = L;
        bool thr_min = true;            // threshold uses min(thr_len, T); false: T alone (tail group)
        if (o.num_thread) {
            const unsigned long long gidx = a.first_index + r;
            const unsigned long long bstart = gidx - (gidx % FQ_REF_BATCH);
            unsigned long long N = FQ_REF_BATCH;
            if (a.end_index != ~0ull && a.end_index - bstart < N) N = a.end_index - bstart;
            const uint32_t ib = (uint32_t)(gidx - bstart), nt = o.num_thread;
            const uint32_t q = (uint32_t)N / nt, rr = (uint32_t)N % nt;
            uint32_t cs, sz;
            if (ib < rr * (q + 1)) { const uint32_t t = ib / (q + 1); cs = t * (q + 1); sz = q + 1; }
            else { const uint32_t t = rr + (q ? (ib - rr * (q + 1)) / q : 0); cs = rr * (q + 1) + (t - rr) * q; sz = q; }
            const uint32_t pc = ib - cs, grp = pc >> 3;
            if (grp * 8 + 8 <= sz) thr_len = a.rec[mate][r + (grp * 8 + 7 - pc)].len;
            else thr_min = false;
        }

        bool all_shared = false;                  // every adapter shares a base with the read (the usual case): see below
        if (a.sweep && L) {
            for (uint32_t w = lane; w < cand_words; w += 32) s_cand[w] = 0;
            bool sh_ok = true;
            for (uint32_t j = lane; j < A.n; j += 32) sh_ok = sh_ok && (read_or & A.or_bits[j]) && (s_off[j + 1] > s_off[j]);
            all_shared = __all_sync(0xffffffffu, sh_ok);
            __syncwarp();
            const uint32_t *rx = s_rxy + a.rpad;
            // Bounds: from the whole adapter (precomputed) unless this read's threshold length is shorter than some adapter
            const bool full = !thr_min || thr_len >= s_meta[1];
            if (!full) {
                for (uint32_t k = lane; k < a.n_seg; k += 32) {
                    const SegInfo si = s_seginfo[k];
                    const uint32_t T = si.total;
                    const int threshold = __float2int_rz(__fmul_rn(o.match_rate, (float)min(thr_len, T)));
                    s_wbound[k] = (uint8_t)(threshold <= 0 ? 0u : ((uint32_t)threshold * si.len + T - 1) / T);
                }
                __syncwarp();
            }
            // Lane = diagonal.  Window w, lane l: segment bit i faces read position p0 + i with p0 = -S + 32 w + l; the planes are
            // zero outside the read.  The windows of one pass stay in registers while the segments stream past them (their masks
            // are warp-uniform shared-memory reads), one popcount per (diagonal, segment).
            const int S = full ? (int)s_meta[0] : 32;
            const int n_windows = ((int)L + S + 31) >> 5;
            const SweepCtx sc{rx, rstride, s_seg, s_segp, s_seginfo, s_wbound, s_cand, a.n_seg, lane, S, n_windows, full, a.sweep_pure != 0};
            const bool fast = full && a.sweep_pure;
            if (fast) {
                for (uint32_t w = lane; w < (a.n_seg + 31) >> 5; w += 32) s_segflag[w] = 0;
                __syncwarp();           // lane 0 ORs into these words
            }
            if (n_windows <= 5) sweep_windows<5>(sc, 0, s_segflag);
            else
                for (int w0 = 0; w0 < n_windows; w0 += 6) sweep_windows<6>(sc, w0, s_segflag);
            if (fast) {             // segment flags -> adapters
                __syncwarp();
                for (uint32_t k = lane; k < a.n_seg; k += 32) {
                    const uint32_t n_here = min(32u, a.n_seg - (k & ~31u));
                    if ((s_segflag[k >> 5] >> (n_here - 1 - (k & 31))) & 1u) {
                        const uint32_t owner = s_seginfo[k].owner;
                        atomicOr(&s_cand[owner >> 5], 1u << (owner & 31));
                    }
                }
            }
            __syncwarp();
        }

        int best_score = 0, best_adapter = -1;
        int st_start = 0, st_stop = 0;            // max_elem.M_start_i / stop_i survive across align() calls (Q5)
        int pending = -1;                         // last adapter that shares a base with the read but was not aligned yet

        // When every adapter shares a base with the read, an adapter the sweep did not flag cannot pass its threshold and its
        // stale range (Q5) is never asked for: only the flagged adapters are visited, in order.
        uint32_t j = 0;
        const bool flagged_only = a.sweep && all_shared && a.n_seg > 0;
        for (;; ++j) {
            if (flagged_only) {
                // next adapter the sweep flagged, or one it does not cover (longer than its limit: visited always)
                uint32_t w = j >> 5, bits = w < cand_words ? (s_cand[w] | s_unswept[w]) & (0xffffffffu << (j & 31)) : 0u;
                while (!bits && ++w < cand_words) bits = s_cand[w] | s_unswept[w];
                j = bits ? (w << 5) + (uint32_t)__ffs(bits) - 1 : A.n;
            }
            if (j >= A.n) break;
            const uint32_t T = s_off[j + 1] - s_off[j];
            const uint8_t *t = s_codes + s_off[j];
            const uint32_t tl = thr_min ? min(thr_len, T) : T;
            const int threshold = __float2int_rz(__fmul_rn(o.match_rate, (float)tl));
            int score = 0;
            const bool shared = (read_or & A.or_bits[j]) && L && T;
            if (shared) {
                bool candidate = true;
                const bool swept = a.sweep && T <= kSweepMaxLen;
                if (swept && !((s_cand[j >> 5] >> (j & 31)) & 1u)) candidate = false;      // the sweep: no diagonal can reach the threshold
                else if (swept && T <= 32) candidate = true;                                // one segment: the sweep's count was exact
                else if (a.use_planes) {
                    // matches on every diagonal that is long enough to reach the threshold, 32 diagonals per round.
                    // Diagonal d pairs adapter position j with read position j - d: for each adapter word the
                    // facing 32 read bits are pulled out of the (zero padded) read planes with a funnel shift.
                    const uint32_t Wt = (T + 31) / 32;
                    const uint32_t *tp = s_planes + s_poff[j];
                    const uint32_t *rp = s_rp + a.rpad;
                    const int need = max(threshold, 1);
                    const int d_lo = max(-(int)(L - 1), need - (int)L), d_hi = min((int)T - 1, (int)T - need);
                    int cmax = 0;
                    for (int d0 = d_lo; d0 <= d_hi; d0 += 32) {
                        const int d = d0 + (int)lane;
                        int cnt = 0;
                        for (uint32_t wa = 0; wa < Wt; ++wa) {
                            const int ofs = 32 * (int)wa - d;            // read bit that faces adapter bit 32*wa
                            const int idx = ofs >> 5;                    // floor
                            const uint32_t sh = (uint32_t)ofs & 31u;
                            uint32_t m = 0;
#pragma unroll
                            for (int b = 0; b < 5; ++b) {
                                if (b == 4 && !a.has_gap) break;
                                const uint32_t *pl = rp + b * rstride;
                                m |= tp[b * Wt + wa] & __funnelshift_r(pl[idx], pl[idx + 1], sh);
                            }
                            cnt += __popc(m);
                        }
                        if (d > d_hi) cnt = 0;
                        cmax = max(cmax, cnt);
                    }
                    cmax = __reduce_max_sync(0xffffffffu, cmax);
                    candidate = cmax >= threshold;
                }
                if (candidate) {
                    score = exact_align(s_read, L, t, T, lane, st_start, st_stop);
                    pending = -1;
                } else pending = (int)j;            // cannot pass; its range matters only to a later all-mismatch adapter
            } else if (pending >= 0) {
                // all-mismatch adapter: it reports the range left by the previous alignment (Q5): produce it now
                exact_align(s_read, L, s_codes + s_off[pending], s_off[pending + 1] - s_off[pending], lane, st_start, st_stop);
                pending = -1;
            }
            if (shared && score == 0 && pending >= 0) continue;      // skipped by the prefilter: cannot pass
            // trim.cpp:1021-1041
            const int match_length = st_stop - st_start + 1;
            const int num_match = (match_length + score) / 2;
            if (num_match >= threshold) {
                for (uint32_t w = lane; w < mask_words; w += 32) {
                    const int lo_b = max(st_start, (int)(w << 5)), hi_b = min(st_stop, (int)(w << 5) + 31);
                    if (lo_b <= hi_b) {
                        const uint32_t nb = (uint32_t)(hi_b - lo_b + 1);
                        const uint32_t bits = (nb == 32 ? 0xffffffffu : ((1u << nb) - 1u)) << (lo_b & 31);
                        s_mask[w] |= bits;
                    }
                }
                if (score > best_score) { best_score = score; best_adapter = (int)j; }
            }
        }
        __syncwarp();

        uint32_t out_start = 0, out_len = L;
        if (best_score > 0) {
            // find_mask_range (trim.cpp:1144-1189) over runs instead of bits; every lane runs it redundantly
            uint32_t run_start = 0, run_len = 0, longest = 0, longest_start = 0;
            for (uint32_t w = 0; w < read_words; ++w) {      // only the words this read owns
                const uint32_t nbits = min(32u, L - (w << 5));
                uint32_t keep = ~s_mask[w];
                if (nbits < 32) keep &= (1u << nbits) - 1u;
                uint32_t bp = 0;
                while (bp < nbits) {
                    const uint32_t x = keep >> bp;
                    if (x & 1u) {
                        const uint32_t ones = min((uint32_t)(__ffs(~x) ? __ffs(~x) - 1 : 32), nbits - bp);
                        if (run_len == 0) run_start = (w << 5) + bp;
                        run_len += ones;
                        bp += ones;
                    } else {
                        const uint32_t zeros = x ? (uint32_t)(__ffs(x) - 1) : (nbits - bp);
                        if (run_len > longest) { longest = run_len; longest_start = run_start; run_len = 0; }
                        bp += zeros;
                    }
                }
            }
            if (run_len > longest) { longest = run_len; longest_start = run_start; }
            if (longest == 0) longest_start = 0;
            out_start = longest_start;
            out_len = longest;
            if (lane == 0) {
                atomicAdd(&a.stats[a.L.adapter_reads + best_adapter], 1ull);
                atomicAdd(&a.stats[a.L.adapter_bases + best_adapter], (unsigned long long)(L - longest));
            }
        }
        if (lane == 0) {
            a.adp[mate][r] = make_uint2(out_start, out_len);
            a.adp_best[mate][r] = best_score > 0 ? best_adapter : -1;
        }
        __syncwarp();
    }
}

}  // namespace fq
