// Probe: can the TMA's tensor form move byte spans between ARBITRARILY aligned global addresses?  One-byte elements, 1-D
// tensor maps with boxes of 256 / 64 / 16 bytes; a warp copies one span: tensor loads at byte coordinate `src` into 16-byte
// aligned shared memory, tensor stores from there to byte coordinate `dst`; the last < 16 bytes go by lanes.
// Prints whether the bytes arrive and the copy rate against cudaMemcpy D2D.   nvcc -arch=sm_100a -o tma_unaligned tma_unaligned.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct Span { unsigned long long src, dst; uint32_t len, pad; };
struct Maps { CUtensorMap in256, in64, in16, out256, out64, out16; };

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const CUtensorMap *map, int c0, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2}], [%3];"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_1d(const CUtensorMap *map, int c0, uint32_t src)
{
    asm volatile("cp.async.bulk.tensor.1d.global.shared::cta.bulk_group [%0, {%1}], [%2];" :: "l"(map), "r"(c0), "r"(src) : "memory");
}

constexpr int kWarps = 4;
constexpr uint32_t kSlab = 24 * 1024;      // spans up to ~22 KB (+ the 128-byte cells of the small boxes)

__global__ void __launch_bounds__(kWarps * 32) k_copy(const __grid_constant__ Maps M, const Span *spans, uint32_t n_spans, const uint8_t *in, uint8_t *out,
                                                    int *err)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) unsigned long long bars[kWarps];
    const uint32_t wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *slab = smem + (size_t)wid * kSlab;
    const uint32_t bar = smem_u32(&bars[wid]);
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    uint32_t phase = 0;
    for (uint32_t s = blockIdx.x * kWarps + wid; s < n_spans; s += gridDim.x * kWarps) {
        const Span sp = spans[s];
        const uint32_t body = sp.len & ~15u;
        if (lane == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(body) : "memory");
            // the shared-memory side of a tensor copy must be 128-byte aligned: boxes below 128 bytes get a 128-byte cell each
            uint32_t o = 0, so = 0;
            for (; o + 256 <= body; o += 256, so += 256) tma_load_1d(smem_u32(slab + so), &M.in256, (int)(sp.src + o), bar);
            for (; o + 64 <= body; o += 64, so += 128) tma_load_1d(smem_u32(slab + so), &M.in64, (int)(sp.src + o), bar);
            for (; o + 16 <= body; o += 16, so += 128) tma_load_1d(smem_u32(slab + so), &M.in16, (int)(sp.src + o), bar);
        }
        // bounded wait: a copy that never completes is reported, not waited for
        uint32_t done = 0;
        for (int it = 0; it < (1 << 22) && !done; ++it)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(phase) : "memory");
        if (!done) { if (lane == 0) atomicExch(err, 1); return; }
        phase ^= 1;
        if (lane == 0) {
            uint32_t o = 0, so = 0;
            for (; o + 256 <= body; o += 256, so += 256) tma_store_1d(&M.out256, (int)(sp.dst + o), smem_u32(slab + so));
            for (; o + 64 <= body; o += 64, so += 128) tma_store_1d(&M.out64, (int)(sp.dst + o), smem_u32(slab + so));
            for (; o + 16 <= body; o += 16, so += 128) tma_store_1d(&M.out16, (int)(sp.dst + o), smem_u32(slab + so));
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        for (uint32_t i = body + lane; i < sp.len; i += 32) out[sp.dst + i] = in[sp.src + i];      // the last < 16 bytes
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");              // the slab is free again
        __syncwarp();
    }
}

static bool encode(EncodeTiledFn enc, CUtensorMap *m, void *base, uint64_t n, uint32_t box)
{
    cuuint64_t dims[1] = {n};
    cuuint64_t strides[1] = {0};          // rank - 1 entries are read: none
    cuuint32_t boxd[1] = {box}, estr[1] = {1};
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, base, dims, strides, boxd, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) printf("cuTensorMapEncodeTiled failed: %d (box %u)\n", (int)r, box);
    return r == CUDA_SUCCESS;
}

int main(int argc, char **argv)
{
    const size_t total = argc > 1 ? (size_t)atoll(argv[1]) : (size_t)1300 << 20;
    const uint32_t span = argc > 2 ? (uint32_t)atoi(argv[2]) : 10600;
    const uint32_t align = argc > 3 ? (uint32_t)atoi(argv[3]) : 1;          // 16: control run with aligned spans
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
    if (!fn) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    auto enc = reinterpret_cast<EncodeTiledFn>(fn);
    uint8_t *d_in, *d_out, *d_ref;
    CK(cudaMalloc(&d_in, total + 4096)); CK(cudaMalloc(&d_out, total + 4096)); CK(cudaMalloc(&d_ref, total + 4096));
    std::vector<uint8_t> h(total);
    uint32_t x = 12345;
    for (size_t i = 0; i < total; ++i) { x = x * 1664525u + 1013904223u; h[i] = (uint8_t)(x >> 24); }
    CK(cudaMemcpy(d_in, h.data(), total, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_out, 0, total + 4096));
    // spans: back to back in the input; in the output every span is shifted by a few more bytes (1 % of the input is skipped),
    // so source and destination alignments drift against each other through all 16 x 16 combinations
    std::vector<Span> spans;
    size_t src = align > 1 ? align : 3, dst = 0;
    std::vector<uint8_t> want(total, 0);
    while (src + span + 64 < total) {
        x = x * 1664525u + 1013904223u;
        uint32_t len = span - 300 + (x >> 24) % 600, skip = (x >> 8) % 7;
        if (align > 1) { len = len / align * align; skip = skip * align; }
        spans.push_back(Span{src, dst, len, 0});
        memcpy(want.data() + dst, h.data() + src, len);
        src += len + skip;
        dst += len;
    }
    const size_t out_bytes = dst;
    Span *d_spans; int *d_err;
    CK(cudaMalloc(&d_spans, spans.size() * sizeof(Span))); CK(cudaMalloc(&d_err, 4)); CK(cudaMemset(d_err, 0, 4));
    CK(cudaMemcpy(d_spans, spans.data(), spans.size() * sizeof(Span), cudaMemcpyHostToDevice));
    Maps M;
    if (!encode(enc, &M.in256, d_in, total, 256) || !encode(enc, &M.in64, d_in, total, 64) || !encode(enc, &M.in16, d_in, total, 16) ||
        !encode(enc, &M.out256, d_out, total, 256) || !encode(enc, &M.out64, d_out, total, 64) || !encode(enc, &M.out16, d_out, total, 16)) return 1;
    const size_t smem = (size_t)kWarps * kSlab;
    CK(cudaFuncSetAttribute(k_copy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1, sms = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_copy, kWarps * 32, smem));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int grid = sms * per_sm;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e9f, best_cpy = 1e9f;
    for (int it = 0; it < 5; ++it) {
        CK(cudaEventRecord(e0));
        k_copy<<<grid, kWarps * 32, smem>>>(M, d_spans, (uint32_t)spans.size(), d_in, d_out, d_err);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = ms < best ? ms : best;
        CK(cudaEventRecord(e0));
        CK(cudaMemcpyAsync(d_ref, d_in, out_bytes, cudaMemcpyDeviceToDevice));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1)); best_cpy = ms < best_cpy ? ms : best_cpy;
    }
    CK(cudaGetLastError());
    int err = 0; CK(cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost));
    std::vector<uint8_t> got(out_bytes);
    CK(cudaMemcpy(got.data(), d_out, out_bytes, cudaMemcpyDeviceToHost));
    size_t bad = 0, first = 0;
    for (size_t i = 0; i < out_bytes; ++i) if (got[i] != want[i]) { if (!bad) first = i; ++bad; }
    printf("spans %zu of ~%u bytes, %zu bytes out, grid %d x %d threads (%d CTAs/SM): timeout flag %d, wrong bytes %zu (first at %zu)\n", spans.size(), span, out_bytes,
           grid, kWarps * 32, per_sm, err, bad, first);
    printf("tensor-TMA copy %.3f ms = %.0f GB/s (read + write); cudaMemcpy D2D of the same volume %.3f ms = %.0f GB/s\n", best, 2.0 * out_bytes / best / 1e6, best_cpy,
           2.0 * out_bytes / best_cpy / 1e6);
    return bad || err ? 2 : 0;
}
