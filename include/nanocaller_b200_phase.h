/* nanocaller_b200 — read-based phasing and haplotagging on the host (libnc_phase.so).
 *
 * Replaces what the reference delegates to an external program between its two stages:
 *   `whatshap phase`    indelCaller.py:237   heterozygous SNP calls (QUAL >= --phase_qual_score, :232) -> phased genotypes + PS
 *   `whatshap haplotag` indelCaller.py:244   reads -> HP / PS tags, which generate_indel_pileups.py:180-188 then reads
 * WhatsHap itself is not part of the reference repository (environment.yml:15) and cannot be had in this image, so this
 * is a separately specified algorithm, validated against the synthetic generator's true haplotypes —
 * "parity unpinned" against WhatsHap.  Plain C ABI, host memory only, no CUDA: it runs between the GPU stages.
 *
 * Sites are heterozygous SNPs sorted by position; a site has two alleles A and B given as BAM base nibbles
 * (A=1 C=2 G=4 T=8).  `0/1` calls pass A = REF, B = ALT; `1/2` calls pass A = ALT1, B = ALT2.
 */
#ifndef NANOCALLER_B200_PHASE_H
#define NANOCALLER_B200_PHASE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NC_PHASE_NONE 255   /* the read shows neither allele (other base, deletion, reference skip, beyond the sequence) */

/* Allele every read shows at every site it spans.  Read r spans the sites first_site[r] .. first_site[r] + n_r - 1 with
 * n_r = pair_off[r+1] - pair_off[r] (the caller computes both from two binary searches of site_pos against pos and the
 * reference end); allele[pair_off[r] + k] receives 0 (A), 1 (B) or NC_PHASE_NONE.  Reads are BAM-native: 0-based pos,
 * CIGAR words len<<4|op, 4-bit bases starting on a byte boundary per read.  threads <= 0: all cores.  Returns 0. */
int nc_phase_read_alleles(int64_t n_reads, const int32_t* pos, const int64_t* cigar_off, const uint32_t* cigar,
                          const int64_t* seq_off, const int32_t* l_seq, const uint8_t* seq4,
                          int64_t n_sites, const int32_t* site_pos, const uint8_t* nib_a, const uint8_t* nib_b,
                          const int64_t* first_site, const int64_t* pair_off, uint8_t* allele, int32_t threads);

/* Phasing and haplotagging from the allele table.
 *   read_use[r]   1: the read takes part in phasing (primary, passes the flag filter); every read is haplotagged
 *   site_hap[j]   out: 0 = haplotype 1 carries A (GT A|B), 1 = haplotype 1 carries B (GT B|A), -1 = unphased
 *   site_block[j] out: index of the first site of j's phase block (-1 when unphased); PS = site_pos[site_block[j]] + 1
 *   read_hp[r]    out: 1, 2, or 0 (no phased site, or a tie)
 *   read_block[r] out: the phase block the tag refers to (-1 when read_hp is 0)
 * `iterations` = refinement sweeps after the left-to-right pass (2 is enough in practice).  Returns 0. */
int nc_phase_sites(int64_t n_reads, const uint8_t* read_use, const int64_t* first_site, const int64_t* pair_off,
                   const uint8_t* allele, int64_t n_sites, int32_t iterations,
                   int8_t* site_hap, int32_t* site_block, int8_t* read_hp, int32_t* read_block);

#ifdef __cplusplus
}
#endif
#endif
