/* nanocaller_b200_io.h — C ABI of libnc_bamio.so: native BAM reader into the staging arrays of libnanocaller_b200.
 *
 * Replaces, for the hot path's inputs, what the reference gets from pysam / htslib when it opens the alignment file inside
 * every call:
 *     pysam.Samfile(sam_path, "rb") + fetch / pileup     nanocaller_src/generate_SNP_pileups.py:134-156
 *                                                         nanocaller_src/generate_indel_pileups.py:147-185 (HP / PS tags :181-185)
 *     sam_file.references / get_reference_length          nanocaller_src/utils.py:9-50
 * The device path consumes BAM's own encodings (CIGAR words len<<4|op, 4-bit bases), so records are copied, not decoded.
 * Plain pointers and sizes; every function returns 0 or a negative NC_IO_* code; nc_bam_error(handle) gives the text.
 *
 * Use: nc_bam_open -> nc_bam_n_contigs / nc_bam_contig (sizes) -> allocate -> nc_bam_fill per contig -> nc_bam_close.
 * The filled arrays are exactly the arguments of nc_stage_reads (include/nanocaller_b200.h).
 */
#ifndef NANOCALLER_B200_IO_H
#define NANOCALLER_B200_IO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NC_IO_OK 0
#define NC_IO_EINVAL (-1)   /* bad argument */
#define NC_IO_EOPEN (-2)    /* file cannot be opened / read */
#define NC_IO_EFORMAT (-3)  /* not BGZF / not BAM / truncated / not coordinate-sorted */

typedef struct nc_bam nc_bam;

typedef struct NcBamContig {
    char name[256];
    int32_t length;      /* @SQ LN */
    int32_t reserved;
    int64_t n_reads;     /* mapped records (refID = this contig), coordinate-sorted */
    int64_t n_cigar;     /* total CIGAR words */
    int64_t n_seq;       /* total packed sequence bytes, every read starting on a byte boundary */
} NcBamContig;

/* Reads and inflates the whole file with `threads` workers (<= 0: all cores) and indexes the records.  *out is set even on
 * failure (for nc_bam_error) and must be closed. */
int nc_bam_open(const char* path, int threads, nc_bam** out);
/* The same for a subset of contigs of a BAM that has a BAI index next to it: only the BGZF blocks holding the named contigs'
 * records are inflated (samfile.fetch(chrom, ...), generate_SNP_pileups.py:141,156).  Every reference of the header is listed by
 * nc_bam_contig; contigs that were not asked for (or have no reads) report n_reads = 0. */
int nc_bam_open_region(const char* path, const char* bai_path, const char* const* names, int n_names, int threads, nc_bam** out);
const char* nc_bam_error(const nc_bam* b);
int nc_bam_n_contigs(const nc_bam* b);
const char* nc_bam_header_text(const nc_bam* b, int64_t* len);
int nc_bam_contig(const nc_bam* b, int i, NcBamContig* out);

/* Fills caller-owned arrays for contig i: pos[n], flag[n], cigar_off[n+1], cigar[n_cigar], seq_off[n+1], l_seq[n],
 * seq4[n_seq], and (optional, both or neither) hp[n], ps[n] from the integer HP / PS aux tags (0 when absent). */
int nc_bam_fill(const nc_bam* b, int i, int threads, int32_t* pos, uint16_t* flag, int64_t* cigar_off, uint32_t* cigar,
                int64_t* seq_off, int32_t* l_seq, uint8_t* seq4, int8_t* hp, int32_t* ps);

/* Query name of read k of contig i (debugging / duplicate-name checks; not needed by the device path). */
int nc_bam_qname(const nc_bam* b, int i, int64_t k, char* out, int cap);

/* Haplotagged copy of contig i's records into a new BAM file (the reference's intermediate_phase_files/{contig}.phased.bam,
 * written there by `whatshap haplotag | samtools view -b`, indelCaller.py:244): records are copied whole (names, qualities,
 * aux fields) except that existing HP / PS fields are dropped and, where hp[k] > 0, `HP:C:hp[k]` and `PS:i:ps[k]` are appended
 * (k counts the contig's mapped records in file order, as nc_bam_fill does).  The header keeps all references.  No index is
 * written.  level: zlib level 0..9; threads <= 0: all cores. */
int nc_bam_write_tagged(const nc_bam* b, int i, const int8_t* hp, const int32_t* ps, const char* out_path, int level, int threads);

void nc_bam_close(nc_bam* b);

#ifdef __cplusplus
}
#endif
#endif
