/*
 * faqcs_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * C ABI of the CPU restatement of FaQCs v2.10's trim()/routing path
 * (oracle/faqcs_oracle.cpp).  It deliberately mirrors include/faqcs_b200.h
 * (same POD structs, fqo_ prefix) so that parity tests drive the CUDA library
 * and the oracle through the same harness.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 */
#ifndef FAQCS_ORACLE_H
#define FAQCS_ORACLE_H

#include "../include/faqcs_b200.h"   /* POD option / result / stats structs only */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fqo_ctx fqo_ctx;

fq_status   fqo_create(const fq_options *opt, fqo_ctx **out);
void        fqo_destroy(fqo_ctx *ctx);
const char *fqo_last_error(const fqo_ctx *ctx);
fq_status   fqo_set_debug_results(fqo_ctx *ctx, int enable);
fq_status   fqo_set_quality(fqo_ctx *ctx, int32_t quality);       /* m_opt.quality = ... between trim() calls, FaQCs.cpp:272-277 */
fq_status   fqo_autodetect(fqo_ctx *ctx, const uint8_t *r1, size_t n1,
                           const uint8_t *r2, size_t n2,
                           int32_t *input_quality_offset, int32_t *quality);
fq_status   fqo_process_host(fqo_ctx *ctx, const uint8_t *r1, size_t n1,
                             const uint8_t *r2, size_t n2,
                             uint64_t first_record_index, int is_final,
                             fq_batch_out *out);
fq_status   fqo_stats(fqo_ctx *ctx, fq_stats_view *view);
/* k-mer rarefaction: update_kmer (trim.cpp:887-931), the sampling at the end of trim() (trim.cpp:157-185), the end of a
 * pass (FaQCs.cpp:518-537, 737-756) -- see fq_kmer_* in include/faqcs_b200.h */
fq_status   fqo_kmer_enable(fqo_ctx *ctx, uint32_t k, uint64_t split_size, uint32_t num_subsample);
fq_status   fqo_kmer_end_pass(fqo_ctx *ctx);
fq_status   fqo_kmer_results(fqo_ctx *ctx, fq_kmer_view *view);

/* Single-function probes used by unit tests of the micro-semantics. */
/* BWA_plus / BWA / HARD trim of one quality string: returns new length, *f5 = 5' cut. */
uint32_t fqo_quality_trim(int mode, int quality, int in_offset, int protect_5,
                          const char *qual, uint32_t len, uint32_t *f5);
/* Ungapped local alignment of one read against one target: score, query start/stop.
 * Returns 0 when score > 0, 1 when the range is stale (score == 0). */
int fqo_align(const char *read, uint32_t read_len, const char *target, uint32_t target_len,
              int32_t *score, int32_t *start, int32_t *stop);
/* find_mask_range on a 0/1 byte mask (1 = keep). */
void fqo_find_mask_range(const uint8_t *mask, uint32_t len, uint32_t *start, uint32_t *length);
/* int(float(1.0 - rate) * n) exactly as trim.cpp:969,1007-1008 computes it. */
int32_t fqo_match_threshold(float rate, uint64_t n);
/* composition bin of `count` bases out of `len` (trim.cpp:860-874). */
uint32_t fqo_composition_bin(uint32_t len, uint32_t count);
float    fqo_average_quality(const char *qual, uint32_t len, int offset);

#ifdef __cplusplus
}
#endif
#endif
