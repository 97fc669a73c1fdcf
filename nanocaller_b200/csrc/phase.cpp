// libnc_phase.so — read-based phasing of heterozygous SNP calls and haplotagging of reads on the host
// (include/nanocaller_b200_phase.h; stands in for `whatshap phase` / `whatshap haplotag`, indelCaller.py:237,:244).
//
// Algorithm:
//   1. allele table: one CIGAR walk per read against the sorted site list (threads over reads);
//   2. phase blocks: connected components of sites linked by a phasing read that is informative at both;
//   3. left-to-right pass: every read carries a score (> 0: believed to come from haplotype 1); a site's orientation is the
//      vote of its reads, weighted by their clamped scores; the reads' scores are then updated with the decision;
//   4. refinement sweeps: every site is re-decided against the scores the reads have WITHOUT that site (a local search on the
//      minimum-error-correction objective); a flip updates the scores at once;
//   5. haplotag: sign of a read's agreement count over the phased sites of its best-supported block.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <numeric>
#include <thread>
#include <vector>

#include "nanocaller_b200_phase.h"

namespace {

inline int ref_len(uint32_t w) { return ((0x18Du >> (w & 15)) & 1u) ? (int)(w >> 4) : 0; }     // M D N = X
inline int qry_len(uint32_t w) { return ((0x193u >> (w & 15)) & 1u) ? (int)(w >> 4) : 0; }     // M I S = X
inline bool is_match(uint32_t w) { return ((0x181u >> (w & 15)) & 1u) != 0; }                  // M = X

inline int clamp2(int v) { return v > 2 ? 2 : (v < -2 ? -2 : v); }

struct Dsu {
    std::vector<int32_t> p;
    explicit Dsu(size_t n) : p(n) { std::iota(p.begin(), p.end(), 0); }
    int32_t find(int32_t x) { while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; } return x; }
    void unite(int32_t a, int32_t b) { a = find(a); b = find(b); if (a != b) { if (a < b) p[b] = a; else p[a] = b; } }   // root = smallest index
};

}  // namespace

extern "C" int nc_phase_read_alleles(int64_t n_reads, const int32_t* pos, const int64_t* cigar_off, const uint32_t* cigar,
                                     const int64_t* seq_off, const int32_t* l_seq, const uint8_t* seq4,
                                     int64_t n_sites, const int32_t* site_pos, const uint8_t* nib_a, const uint8_t* nib_b,
                                     const int64_t* first_site, const int64_t* pair_off, uint8_t* allele, int32_t threads) {
    if (n_reads < 0 || n_sites < 0) return -1;
    if (n_reads == 0 || n_sites == 0) return 0;
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min<int>(nt, (int)std::min<int64_t>(n_reads, 256)));
    auto work = [&](int64_t r0, int64_t r1) {
        for (int64_t r = r0; r < r1; r++) {
            const int64_t np = pair_off[r + 1] - pair_off[r];
            if (np <= 0) continue;
            uint8_t* out = allele + pair_off[r];
            int64_t j = first_site[r];
            const int64_t jend = j + np;
            const uint8_t* sq = seq4 + seq_off[r];
            int32_t x = pos[r], y = 0;
            for (int64_t k = cigar_off[r]; k < cigar_off[r + 1] && j < jend; k++) {
                const uint32_t w = cigar[k];
                const int rl = ref_len(w), ql = qry_len(w);
                if (rl > 0) {
                    while (j < jend && site_pos[j] < x + rl) {
                        uint8_t a = NC_PHASE_NONE;
                        if (site_pos[j] >= x && is_match(w)) {
                            const int32_t q = y + (site_pos[j] - x);
                            if (q < l_seq[r]) {
                                const uint8_t b = sq[q >> 1];
                                const uint8_t nib = (q & 1) ? (b & 15) : (b >> 4);
                                a = nib == nib_a[j] ? 0 : (nib == nib_b[j] ? 1 : NC_PHASE_NONE);
                            }
                        }
                        out[j - first_site[r]] = a;
                        j++;
                    }
                }
                x += rl; y += ql;
            }
            for (; j < jend; j++) out[j - first_site[r]] = NC_PHASE_NONE;
        }
    };
    if (nt == 1) { work(0, n_reads); return 0; }
    std::vector<std::thread> th;
    const int64_t per = (n_reads + nt - 1) / nt;
    for (int t = 0; t < nt; t++) {
        const int64_t r0 = t * per, r1 = std::min(n_reads, r0 + per);
        if (r0 < r1) th.emplace_back(work, r0, r1);
    }
    for (auto& t : th) t.join();
    return 0;
}

extern "C" int nc_phase_sites(int64_t n_reads, const uint8_t* read_use, const int64_t* first_site, const int64_t* pair_off,
                              const uint8_t* allele, int64_t n_sites, int32_t iterations,
                              int8_t* site_hap, int32_t* site_block, int8_t* read_hp, int32_t* read_block) {
    if (n_reads < 0 || n_sites < 0) return -1;
    for (int64_t j = 0; j < n_sites; j++) { site_hap[j] = -1; site_block[j] = -1; }
    for (int64_t r = 0; r < n_reads; r++) { read_hp[r] = 0; read_block[r] = -1; }
    if (n_reads == 0 || n_sites == 0) return 0;

    // site -> (read, allele) lists of the phasing reads, and the phase blocks
    std::vector<int64_t> soff((size_t)n_sites + 1, 0);
    Dsu dsu((size_t)n_sites);
    for (int64_t r = 0; r < n_reads; r++) {
        if (!read_use[r]) continue;
        int64_t prev = -1;
        for (int64_t e = pair_off[r]; e < pair_off[r + 1]; e++) {
            if (allele[e] == NC_PHASE_NONE) continue;
            const int64_t j = first_site[r] + (e - pair_off[r]);
            soff[(size_t)j + 1]++;
            if (prev >= 0) dsu.unite((int32_t)prev, (int32_t)j);
            prev = j;
        }
    }
    for (int64_t j = 0; j < n_sites; j++) soff[(size_t)j + 1] += soff[(size_t)j];
    std::vector<int32_t> sread((size_t)soff[(size_t)n_sites]);
    std::vector<uint8_t> sall((size_t)soff[(size_t)n_sites]);
    {
        std::vector<int64_t> fill(soff.begin(), soff.end() - 1);
        for (int64_t r = 0; r < n_reads; r++) {
            if (!read_use[r]) continue;
            for (int64_t e = pair_off[r]; e < pair_off[r + 1]; e++) {
                if (allele[e] == NC_PHASE_NONE) continue;
                const int64_t j = first_site[r] + (e - pair_off[r]);
                sread[(size_t)fill[(size_t)j]] = (int32_t)r; sall[(size_t)fill[(size_t)j]] = allele[e]; fill[(size_t)j]++;
            }
        }
    }
    std::vector<int32_t> bsize((size_t)n_sites, 0);
    for (int64_t j = 0; j < n_sites; j++) bsize[(size_t)dsu.find((int32_t)j)]++;

    // left-to-right pass
    std::vector<int32_t> score((size_t)n_reads, 0);
    std::vector<int8_t> h((size_t)n_sites, 0);
    for (int64_t j = 0; j < n_sites; j++) {
        int64_t vote = 0;
        for (int64_t e = soff[(size_t)j]; e < soff[(size_t)j + 1]; e++) {
            const int w = clamp2(score[(size_t)sread[(size_t)e]]);
            vote += sall[(size_t)e] == 1 ? w : -w;
        }
        h[(size_t)j] = vote > 0 ? 1 : 0;
        for (int64_t e = soff[(size_t)j]; e < soff[(size_t)j + 1]; e++)
            score[(size_t)sread[(size_t)e]] += sall[(size_t)e] == (uint8_t)h[(size_t)j] ? 1 : -1;
    }
    // refinement sweeps
    for (int it = 0; it < iterations; it++) {
        int64_t flips = 0;
        for (int64_t j = 0; j < n_sites; j++) {
            int64_t vote = 0;
            for (int64_t e = soff[(size_t)j]; e < soff[(size_t)j + 1]; e++) {
                const int own = sall[(size_t)e] == (uint8_t)h[(size_t)j] ? 1 : -1;
                const int w = clamp2(score[(size_t)sread[(size_t)e]] - own);
                vote += sall[(size_t)e] == 1 ? w : -w;
            }
            const int8_t want = vote > 0 ? 1 : (vote < 0 ? 0 : h[(size_t)j]);
            if (want != h[(size_t)j]) {
                for (int64_t e = soff[(size_t)j]; e < soff[(size_t)j + 1]; e++) {
                    const int own = sall[(size_t)e] == (uint8_t)h[(size_t)j] ? 1 : -1;
                    score[(size_t)sread[(size_t)e]] -= 2 * own;
                }
                h[(size_t)j] = want;
                flips++;
            }
        }
        if (flips == 0) break;
    }
    for (int64_t j = 0; j < n_sites; j++) {
        const int32_t root = dsu.find((int32_t)j);
        if (bsize[(size_t)root] >= 2 && soff[(size_t)j + 1] > soff[(size_t)j]) { site_hap[j] = h[(size_t)j]; site_block[j] = root; }
    }
    // haplotag every read against the phased sites; the block with the strongest evidence gives the tag
    for (int64_t r = 0; r < n_reads; r++) {
        int32_t best_block = -1, cur_block = -1;
        int64_t best = 0, cur = 0;
        auto close = [&]() { if (cur_block >= 0 && std::llabs(cur) > std::llabs(best)) { best = cur; best_block = cur_block; } };
        for (int64_t e = pair_off[r]; e < pair_off[r + 1]; e++) {
            if (allele[e] == NC_PHASE_NONE) continue;
            const int64_t j = first_site[r] + (e - pair_off[r]);
            if (site_hap[j] < 0) continue;
            if (site_block[j] != cur_block) { close(); cur_block = site_block[j]; cur = 0; }
            cur += allele[e] == (uint8_t)site_hap[j] ? 1 : -1;
        }
        close();
        if (best != 0) { read_hp[r] = best > 0 ? 1 : 2; read_block[r] = best_block; }
    }
    return 0;
}
