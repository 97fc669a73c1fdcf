/* nanocaller_b200.h — C-ABI of libnanocaller_b200.so (sm_100a CUDA, one context per GPU).
 *
 * The reference (WGLab/NanoCaller) is pure Python and has no FFI; the drop-in boundary for the
 * hot path is four Python call sites.  Each entry point below names the reference code it
 * replaces (paths under nanocaller_src/):
 *
 *   get_snp_testing_candidates   generate_SNP_pileups.py:103-278, called at snpCaller.py:86
 *   get_cnd_pos                  generate_SNP_pileups.py:6-101
 *   coverage scaling (N1)        snpCaller.py:90-96, 167-173
 *   SNP_model.call               model_architect.py:36-64,       called at snpCaller.py:111
 *   haploid_SNP_model.call       model_architect_SNP_haploid.py:33-53, called at snpCaller.py:183
 *
 * Conventions: every function returns 0 on success and a negative NC_E* code on failure;
 * nc_last_error(ctx) returns the message of the last failure on that context.  No exceptions
 * cross the boundary.  A context is bound to one CUDA device and one stream and is NOT
 * thread-safe (one host thread per context, like one reference worker process).  All pointer
 * arguments are HOST pointers unless the name ends in _dev.  There is no CPU fallback: with no
 * usable CUDA device nc_create fails.
 */
#ifndef NANOCALLER_B200_H
#define NANOCALLER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NC_OK            0
#define NC_ECUDA        -1   /* CUDA runtime error                          */
#define NC_EINVAL       -2   /* bad argument                                */
#define NC_ESTATE       -3   /* call order violated (e.g. scan before stage) */
#define NC_ENOMEM       -4
#define NC_EOVERFLOW    -5   /* a count does not fit the device encoding    */

#define NC_ABI_VERSION   1

/* --sequencing (NanoCaller:96) -> neighbour-window rule of get_cnd_pos */
#define NC_SEQ_ONT             0
#define NC_SEQ_SHORT_ONT       1
#define NC_SEQ_UL_ONT          2
#define NC_SEQ_UL_ONT_EXTREME  3
#define NC_SEQ_PACBIO          4

/* Tensor geometry (generate_SNP_pileups.py:254): int [5][41][5] per site.  Device and fetch
 * buffers use a per-site stride of NC_SNP_SITE_STRIDE int16 (16-byte aligned rows). */
#define NC_SNP_ROWS          5
#define NC_SNP_COLS          41
#define NC_SNP_CH            5
#define NC_SNP_SITE_ELEMS    1025
#define NC_SNP_SITE_STRIDE   1032

typedef struct nc_ctx nc_ctx;

/* Subset of the reference's `dct` that the SNP path reads (generate_SNP_pileups.py:113-132,
 * 170-183,202,215,244) plus region['ploidy']. */
typedef struct NcSnpParams {
    double  thr_lo, thr_hi;      /* dct['threshold'] = --neighbor_threshold               */
    double  min_allele_freq;     /* dct['min_allele_freq']                                */
    int32_t mincov, maxcov;      /* dct['mincov'], dct['maxcov']                          */
    int32_t min_nbr_sites;       /* dct['min_nbr_sites']                                  */
    int32_t seq;                 /* NC_SEQ_* from dct['seq']                              */
    int32_t supplementary;       /* dct['supplementary'] (keeps 0x800 reads when non-zero) */
    int32_t haploid;             /* region['ploidy'] == 'haploid'                         */
} NcSnpParams;

/* One chunk of utils.get_chunks (utils.py:67-83): 1-based, both ends inclusive. */
typedef struct NcChunk {
    int32_t start, end;
} NcChunk;

/* Per emitted candidate: everything get_snp_testing_candidates returns besides the tensor. */
typedef struct NcSiteMeta {
    int32_t  pos;            /* 1-based v_pos                                                  */
    int32_t  chunk;          /* index into the chunk array of the scan call                   */
    int32_t  dp;             /* get_num_aligned() of the column (:164, :259)                  */
    int32_t  alt;            /* max non-reference A/G/T/C count; freq = alt / dp (:166, :260) */
    uint16_t fwd[4];         /* forward-strand depth of A,G,T,C before down-sampling (:212)   */
    uint16_t rev[4];         /* reverse-strand depth (:213)                                   */
    uint8_t  ref_code;       /* A0 G1 T2 C3 (:104)                                            */
    uint8_t  n_left;         /* len(ls_total_1)                                               */
    uint8_t  n_right;        /* len(ls_total_2)                                               */
    uint8_t  reserved;
    int32_t  sample_depth;   /* len(sample) after the maxcov cut (:215-216, :263)            */
} NcSiteMeta;

/* Cumulative device-side timings of the last nc_snp_scan / nc_snp_forward (CUDA events on the
 * context stream), in milliseconds, and launch counters.  */
typedef struct NcTimings {
    float    decode_ms;      /* K0: CIGAR + 4-bit bases -> reference-aligned code rows        */
    float    scan_ms;        /* K1: column counts, site selection, column store               */
    float    tensor_ms;      /* K2: neighbour choice + tensor build                           */
    float    cnn_ms;         /* K3: CNN forward                                               */
    uint64_t launches;       /* kernels launched by this library on this context since create */
    uint64_t tensor_bytes;   /* algorithmic bytes of the last K2 launch        */
    uint64_t scan_bytes;     /* SURVEY.md 8(d) B1 of the last K0+K1 pass: 4 B per CIGAR op + 4-bit bases + 16 B per read + 1 B per
                              * piled reference position + 5 B per (read, neighbour site); the candidates' (read, site) pairs
                              * are added by the caller from NcSiteMeta.dp */
    float    cnn_a_ms;       /* part of cnn_ms spent in the conv1+conv2 kernel (tensor-core path), else 0 */
    float    reserved;
} NcTimings;

int         nc_abi_version(void);
int         nc_create(int device, nc_ctx** out);
void        nc_destroy(nc_ctx* ctx);
const char* nc_last_error(const nc_ctx* ctx);
int         nc_sync(nc_ctx* ctx);
/* on != 0: host waits of this context sleep (cudaEventBlockingSync) instead of spinning — for hosts where several contexts /
 * ranks share few cores (the reference's analogue is one worker process per --cpu, snpCaller.py:238).  Default: spin. */
int         nc_set_blocking_sync(nc_ctx* ctx, int on);
int         nc_get_timings(nc_ctx* ctx, NcTimings* out);
int         nc_device_sm_count(nc_ctx* ctx);

/* Measurement helpers (bench.py): CUDA events on the context stream — the only stream this library
 * launches on — in slots 0..3, and a switch that drops the decoded aligned rows so the next scan runs
 * K0 again on the already staged (HBM-resident) reads. */
int         nc_event_record(nc_ctx* ctx, int slot);
int         nc_event_elapsed_ms(nc_ctx* ctx, int slot_start, int slot_end, float* ms);   /* synchronises */
int         nc_invalidate_decode(nc_ctx* ctx);

/* Replaces pysam.Samfile/FastaFile opening at generate_SNP_pileups.py:134-137: uploads one
 * contig's coordinate-sorted alignments in BAM-native encoding (SAM spec 4.2: 0-based pos,
 * CIGAR as len<<4|op, 4-bit bases two per byte, each read starting on a byte boundary at
 * seq_off[i]) and the reference characters ref[0..ref_len) whose first base has 0-based
 * coordinate ref_start.  Arrays may be pageable or pinned; the copy is asynchronous on the
 * context stream when pinned.  Invalidates earlier scan results. */
int nc_stage_reads(nc_ctx* ctx, int64_t n_reads, const int32_t* pos, const uint16_t* flag,
                   const int64_t* cigar_off, const uint32_t* cigar, const int64_t* seq_off,
                   const int32_t* l_seq, const uint8_t* seq4, const uint8_t* ref,
                   int64_t ref_start, int64_t ref_len);

/* ---- device-side BAM input (replaces pysam.Samfile / htslib BGZF + record decoding at generate_SNP_pileups.py:134-156 and
 * generate_indel_pileups.py:147-185 for file-backed runs): the COMPRESSED file bytes cross PCIe, BGZF blocks are inflated on the GPU
 * (one thread per block), the record chain is walked there and every mapped record is decoded into the same device arrays
 * nc_stage_reads fills (long CIGARs are restored from CG:B,I, HP / PS tags are staged as by nc_stage_tags). ---- */
typedef struct NcBamDeviceContig {
    char    name[256];
    int32_t length;          /* @SQ LN of the BAM header                                   */
    int32_t reserved;
    int64_t n_reads;         /* records with this refID                                    */
    int64_t n_tagged;        /* of which carry an HP tag of 1 or 2                         */
} NcBamDeviceContig;

/* Reads `path` (a BGZF-compressed BAM; the whole file is inflated into device memory, so it has to fit), parses the header and walks
 * the records.  *n_contigs = references of the header.  Replaces any BAM opened on this context before. */
int nc_bam_device_open(nc_ctx* ctx, const char* path, int32_t* n_contigs);
int nc_bam_device_contig(nc_ctx* ctx, int32_t i, NcBamDeviceContig* out);
/* nc_stage_reads + nc_stage_tags for contig i of the opened BAM, from the inflated stream on the device; `ref` as in nc_stage_reads. */
int nc_bam_device_stage(nc_ctx* ctx, int32_t i, const uint8_t* ref, int64_t ref_start, int64_t ref_len);
/* Frees the inflated stream (staged contigs stay valid). */
int nc_bam_device_close(nc_ctx* ctx);
/* Milliseconds of the last nc_bam_device_open: host (mmap + block table + copy into pinned memory), H2D, inflate kernel, record walk + fields. */
int nc_bam_device_timings(nc_ctx* ctx, float ms[4], int64_t* compressed_bytes, int64_t* inflated_bytes);
/* 1 when the record chain of the open BAM was followed in parallel from the record starts its BAI index (`path`.bai or the path with
 * .bai for .bam) lists, 0 when one thread walked it (no index, or an index that does not match the file). */
int nc_bam_device_walk_mode(nc_ctx* ctx);

/* K0 — htslib pileup-engine replacement (generate_SNP_pileups.py:156):
 * turns every staged read into a reference-aligned row of 4-bit codes (A0 G1 T2 C3, 4 for a
 * deletion / N / anything else).  Called implicitly by nc_snp_scan when needed. */
int nc_decode_reads(nc_ctx* ctx);

/* get_snp_testing_candidates for a batch of chunks of the staged contig
 * (generate_SNP_pileups.py:103-278).  `bed` holds n_bed (start,end) pairs of the exclude BED for
 * this contig (NULL/0 for none), interpreted like the reference does: a site is skipped iff
 * start <= v_pos < end for some pair (:116-119,161).  On return *n_sites is the number of
 * emitted candidates over all chunks, ordered by (chunk, position).  Tensors, metadata and the
 * per-chunk mean sampled depth stay on the device until fetched. */
int nc_snp_scan(nc_ctx* ctx, const NcSnpParams* params, const NcChunk* chunks, int32_t n_chunks,
                const int32_t* bed, int32_t n_bed, int64_t* n_sites);

/* Copies results of the last scan to the host.  Any pointer may be NULL.
 *   mat         int16 [n_sites][NC_SNP_SITE_STRIDE]  (first 1025 of each row = [5][41][5])
 *   meta        NcSiteMeta [n_sites]
 *   chunk_depth double [n_chunks]   mean(len(sample)) per chunk (:274), 0 for empty chunks
 *   chunk_count int64  [n_chunks]   candidates emitted per chunk */
int nc_snp_fetch(nc_ctx* ctx, int16_t* mat, NcSiteMeta* meta, double* chunk_depth, int64_t* chunk_count);

/* Tensors of `count` consecutive sites starting at site `first` of the last scan (sites are in (chunk, position) order, so
 * a chunk is one range): int16 [count][NC_SNP_SITE_STRIDE]. */
int nc_snp_fetch_range(nc_ctx* ctx, int64_t first, int64_t count, int16_t* mat);

/* Model weights: packed fp32 blob in the canonical tensor order of
 * nanocaller_b200/host/weights.py (`pack_snp_blob`): conv1_1 k,b  conv1_2 k,b  conv1_3 k,b
 * conv2 k,b  conv3 k,b  fc1 k,b  then diploid: fa A G T C fc2 fc3 GT (k,b each)
 *                                     haploid: fc2 fc3 (k,b each).
 * Kernels keep Keras layout (HWIO / [in,out]).  train_coverage is the `.coverage` value
 * (snpCaller.py:48-53; 30 for the haploid model, :73). */
int nc_load_snp_weights(nc_ctx* ctx, const float* blob, size_t n_floats, double train_coverage, int haploid);

/* snp_model / hap_snp_model forward on the tensors of the last scan, with the coverage scaling
 * of snpCaller.py:90-96 fused into the input load:
 *   normalize != 0 : x[:,1:,:,:4] *= fp32(train_coverage / chunk_depth)
 *   normalize == 0 : --disable_coverage_normalization, per-site train_coverage / dp (float64 product)
 * probs (host, may be NULL) receives float32 [n_sites][4] = P(A),P(G),P(T),P(C): the [:,1]
 * softmax columns of the four heads (snpCaller.py:115) or the haploid 4-way softmax (:183).
 * impl: 0 = tcgen05 tensor-core kernel (default), 1 = fp32 CUDA-core kernel. */
int nc_snp_forward(nc_ctx* ctx, int normalize, int impl, float* probs);

/* Copies the probabilities of the last nc_snp_forward to the host: float32 [n_sites][4]. */
int nc_snp_fetch_probs(nc_ctx* ctx, float* probs);

/* Drop-in for snp_model([x, A_ref, G_ref, T_ref, C_ref]) (snpCaller.py:111) and
 * hap_snp_model([x, ref]) (:183) on host tensors: x float32 [n][5][41][5] (already scaled),
 * ref_onehot float32 [n][4].  out: diploid float32 [n][10] = out_A[2] out_G[2] out_T[2] out_C[2]
 * out_GT[2]; haploid float32 [n][4]. */
int nc_snp_model_forward(nc_ctx* ctx, const float* x, const float* ref_onehot, int64_t n, int haploid, int impl, float* out);

/* Device pointers of the last scan/forward for callers that keep the data on the GPU
 * (torch / NCCL plumbing); valid until the next stage/scan on this context. */
int nc_snp_device_buffers(nc_ctx* ctx, void** mat_dev, void** meta_dev, void** probs_dev, int64_t* n_sites);

/* Indel CNNs (indelCaller.py:85 `indel_model(batch_x_all)`, :171 `hap_indel_model(batch_x)`;
 * model_architect_indel.py:28-48, model_architect_indels_haploid.py:29-48).  Blob order:
 * conv1_1 k,b conv1_2 k,b conv1_3 k,b conv2 k,b conv3 k,b fc1 k,b fc2 k,b fc3 k,b (Keras layouts). */
int nc_load_indel_weights(nc_ctx* ctx, const float* blob, size_t n_floats, int haploid);

/* x float32 [n][15][128][2] (diploid: hstack of hap0, hap1, all — indelCaller.py:83) or
 * [n][5][128][2] (haploid).  out: float32 [n][4] softmax / [n][1] sigmoid. */
int nc_indel_model_forward(nc_ctx* ctx, const float* x, int64_t n, int haploid, int impl, float* out);

/* ---- indel feature path (diploid): generate_indel_pileups.get_indel_testing_candidates (generate_indel_pileups.py:129-370,
 * called at indelCaller.py:69).  MUSCLE (:30) and parasail (:79) are external; the alignment used instead is this
 * library's own star alignment / affine NW, specified in oracle/star_msa.py and DESIGN.md. ---- */
#define NC_INDEL_CNS_MAX 544

typedef struct NcIndelParams {
    double  ins_t, del_t;            /* dct['ins_t'], dct['del_t']                                  */
    int32_t mincov, maxcov;
    int32_t win_size, small_win_size;
    int32_t window_after;            /* 160, or 260 for dct['seq'] == 'pacbio' (:136-139)           */
    int32_t supplementary;
    int32_t haploid;                 /* generate_indel_pileups_haploid.py: one window set / one MSA over all reads */
    int32_t impute_indel_phase;      /* dct['impute_indel_phase'] (:278-304); ignored by the haploid caller */
} NcIndelParams;

/* variants[key] = type (:268,:274,:301).  src = 0, or for a candidate found by impute_indel_phase the 1-based column whose
 * pileup strings define the two read sets (extra_variants[key], :302); nc_indel_build recomputes the sets from that column. */
typedef struct NcIndelVariant { int32_t key, type, chunk, src; } NcIndelVariant;

typedef struct NcIndelSiteMeta {
    int32_t pos, chunk, type, phase, ref_len;   /* phase: PS of the first hap0 read; -1 = that read has no HP tag (imputed sites) */
    int32_t n[3];                    /* reads used per group: HP1, HP2, all                         */
    int32_t cns_len[3];
    int32_t ok[3];                   /* msa() flag per group (:48)                                  */
} NcIndelSiteMeta;

/* HP / PS tags of the staged reads (hp: 0 untagged, 1, 2), read by generate_indel_pileups.py:180-188. */
int nc_stage_tags(nc_ctx* ctx, const int8_t* hp, const int32_t* ps);

/* Pass 1 (:213-275) over a batch of chunks: *n_variants = number of (key, type, chunk) triples, unordered. */
int nc_indel_scan(nc_ctx* ctx, const NcIndelParams* params, const NcChunk* chunks, int32_t n_chunks, const int32_t* bed,
                  int32_t n_bed, int64_t* n_variants);
int nc_indel_fetch_variants(nc_ctx* ctx, NcIndelVariant* out);

/* Pass 2 + msa (:306-348, :12-73) for key positions given in (chunk, key) order (the host applies the dict semantics of
 * `variants`).  Results stay on the device until fetched:
 *   meta     NcIndelSiteMeta [n_sites]
 *   tensors  float32 [n_sites][3][5][128][2]  (group 0 HP1, 1 HP2, 2 all; zero when !ok)
 *   cns      uint8 [n_sites][3][NC_INDEL_CNS_MAX] consensus codes A0 G1 T2 C3 */
int nc_indel_build(nc_ctx* ctx, const NcIndelParams* params, const NcChunk* chunks, int32_t n_chunks,
                   const NcIndelVariant* sites, int64_t n_sites);
int nc_indel_fetch(nc_ctx* ctx, NcIndelSiteMeta* meta, float* tensors, uint8_t* cns);

/* allele_prediction (generate_indel_pileups.py:77-127) of the last build, computed on the device right after msa: int32
 * [n_sites][3][2] = (length of the reference allele string ref_seq[:r], length of the alternative allele string alt[:a]) per read
 * group, -1 / -1 where the reference returns (None, None) or the site was not kept (:342-348), -2 / -2 if an item did not fit the
 * device scratch (the caller then uses nc_allele_predict_batch for it).  Same alignment as nc_nw_trace. */
int nc_indel_fetch_alleles(nc_ctx* ctx, int32_t* out);

/* Tensors of `count` consecutive sites of the last build starting at site `first`: float32 [count][3][5][128][2]. */
int nc_indel_fetch_range(nc_ctx* ctx, int64_t first, int64_t count, float* tensors);

/* indel_model(batch_x_all) (indelCaller.py:83-85) / hap_indel_model(batch_x) (:171) on the tensors of the last nc_indel_build,
 * which stay on the device: [n_sites][3][5][128][2] IS np.hstack([x0, x1, x2]) = [n_sites][15][128][2] (diploid); the haploid
 * model reads group 2.  Every built site gets a row (the host drops the sites whose msa() flags are not all set, :342-348).
 * probs (host, may be NULL): float32 [n_sites][4] softmax / [n_sites][1] sigmoid.  impl as in nc_snp_forward. */
int nc_indel_forward(nc_ctx* ctx, int impl, float* probs);
int nc_indel_fetch_probs(nc_ctx* ctx, float* probs);

/* Device-side timings (CUDA events on the context stream, host synchronisation gaps included) of the last nc_indel_scan /
 * nc_indel_build / nc_indel_forward, and the unit counts their rooflines are computed from. */
typedef struct NcIndelTimings {
    float    scan_ms;        /* I1: depth per haplotype, indel events, window sets, decision, greedy `prev` pass */
    float    reads_ms;       /* I2: reads and query slices of every key position                               */
    float    align_ms;       /* I3: slice x reference-window alignment (indel_align_kernel)                    */
    float    msa_ms;         /* I3: column merge, frequencies, tensors, consensus (indel_msa_kernel)           */
    float    cnn_ms;         /* M3 / M4: nc_indel_forward                                                      */
    float    allele_ms;      /* I4: consensus x reference alignment + allele lengths (indel_allele_kernel)     */
    float    reserved[2];
    uint64_t n_sites;        /* key positions of the last build                                               */
    uint64_t n_entries;      /* (site, read group member) slices aligned by the last build                    */
    uint64_t scan_bytes;     /* algorithmic bytes of the scan: aligned rows + CIGARs once, 6 B of depth + 1 B flag per column */
    uint64_t build_bytes;    /* algorithmic bytes of the build: slices + reference windows read, tensors + consensus written */
} NcIndelTimings;
int nc_get_indel_timings(nc_ctx* ctx, NcIndelTimings* out);

/* Host-side global affine alignment with traceback, replaces parasail.nw_trace(query, ref, open, extend, matrix)
 * (generate_indel_pileups.py:79): codes A0 G1 T2 C3; cigar_out receives (len << 4 | op) words, op '='7 'X'8 'I'1 'D'2.
 * Returns the number of words, or a negative NC_E* code (NC_EOVERFLOW if cap is too small). */
int nc_nw_trace(const uint8_t* query, int32_t n, const uint8_t* ref, int32_t m, int32_t gap_open, int32_t gap_extend,
                int32_t match, int32_t mismatch, uint32_t* cigar_out, int32_t cap);

/* allele_prediction(alt, ref_seq, max_range) (generate_indel_pileups.py:77-127) for a batch of consensus / reference-window pairs on
 * `threads` host threads (<= 0: all cores): nc_nw_trace, then the reference's CIGAR walk.  Item i reads alt_len[i] codes at
 * alt_codes + alt_off[i] and ref_len[i] codes at ref_codes + ref_off[i]; it returns the lengths of the allele strings
 * ref_seq[:ref_out_len[i]] / alt[:alt_out_len[i]], or -1 / -1 where the reference returns (None, None). */
int nc_allele_predict_batch(int64_t n_items, const uint8_t* alt_codes, const int64_t* alt_off, const int32_t* alt_len,
                            const uint8_t* ref_codes, const int64_t* ref_off, const int32_t* ref_len, const int32_t* max_range,
                            int32_t gap_open, int32_t gap_extend, int32_t match, int32_t mismatch, int32_t threads,
                            int32_t* ref_out_len, int32_t* alt_out_len);

/* SNP genotype decision + VCF record text for n call records (snpCaller.py:113-163; haploid: :183-198) on `threads` host threads.
 * probs4 = P(A), P(G), P(T), P(C) per site (the order of nc_snp_forward), alt_cnt / dp = the FQ value (generate_SNP_pileups.py:166),
 * fwd4 / rev4 = strand depths by base code.  Writes the records back to back into `out` (capacity cap), line_off[n + 1] = start of
 * every record (a candidate the reference writes no record for has an empty range), is_pass[n] = FILTER is PASS.
 * Returns the number of bytes written, NC_EOVERFLOW when cap is too small (512 bytes per record always suffice). */
int64_t nc_format_snp_records(const char* chrom, int64_t n, const int32_t* pos, const uint8_t* ref_code, const float* probs4,
                              const int32_t* dp, const int32_t* alt_cnt, const uint16_t* fwd4, const uint16_t* rev4, int32_t haploid,
                              int32_t threads, char* out, int64_t cap, int64_t* line_off, uint8_t* is_pass);

#ifdef __cplusplus
}
#endif
#endif /* NANOCALLER_B200_H */
