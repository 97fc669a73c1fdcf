// libnc_phase.so — read-based phasing of heterozygous SNP calls and haplotagging of reads on the host
// (include/nanocaller_b200_phase.h; stands in for `whatshap phase` / `whatshap haplotag`, indelCaller.py:237,:244).
//
// Algorithm:
//   1. allele table: one CIGAR walk per read against the sorted site list (threads over reads);
//   2. usable sites: both alleles seen in >= 2 phasing reads and the rarer one in >= 15 % of them;
//   3. pairwise linkage of every usable site with the next 16 sites: cis = reads showing the same allele index at both,
//      trans = different; a link is accepted when |cis - trans| >= 3 and >= 60 % of the reads that see both sites;
//   4. maximum spanning forest over the accepted links, strongest first (union-find with parity): the components with >= 2
//      sites are the phase blocks, the parity along the tree is the orientation; sites that link to nothing stay unphased;
//   5. refinement sweeps: every phased site is re-decided against the scores its reads have WITHOUT that site (a local search
//      on the minimum-error-correction objective); a flip updates the scores at once;
//   6. haplotag: sign of a read's agreement count over the phased sites of its best-supported block.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <utility>
#include <numeric>
#include <thread>
#include <vector>

#include "nanocaller_b200_phase.h"

namespace {

inline int ref_len(uint32_t w) { return ((0x18Du >> (w & 15)) & 1u) ? (int)(w >> 4) : 0; }     // M D N = X
inline int qry_len(uint32_t w) { return ((0x193u >> (w & 15)) & 1u) ? (int)(w >> 4) : 0; }     // M I S = X
inline bool is_match(uint32_t w) { return ((0x181u >> (w & 15)) & 1u) != 0; }                  // M = X

inline int clamp2(int v) { return v > 2 ? 2 : (v < -2 ? -2 : v); }

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
    constexpr int kPairSpan = 16;          // a site is compared with the next 16 sites
    constexpr int kMinMinor = 2;           // usable site: both alleles seen in >= 2 phasing reads ...
    constexpr double kMinMinorFrac = 0.15; // ... and the rarer one in >= 15 % of them
    constexpr int kMinLink = 3;            // accepted link: |cis - trans| >= 3 ...
    constexpr double kMinLinkFrac = 0.6;   // ... and >= 60 % of the reads that see both sites

    // usable sites: a call whose reads (almost) all show one allele carries no phase information and would tie everything to it
    std::vector<int32_t> cnt0((size_t)n_sites, 0), cnt1((size_t)n_sites, 0);
    for (int64_t r = 0; r < n_reads; r++) {
        if (!read_use[r]) continue;
        for (int64_t e = pair_off[r]; e < pair_off[r + 1]; e++) {
            if (allele[e] == NC_PHASE_NONE) continue;
            const int64_t j = first_site[r] + (e - pair_off[r]);
            (allele[e] ? cnt1 : cnt0)[(size_t)j]++;
        }
    }
    std::vector<uint8_t> usable((size_t)n_sites, 0);
    for (int64_t j = 0; j < n_sites; j++) {
        const int32_t mn = std::min(cnt0[(size_t)j], cnt1[(size_t)j]), tot = cnt0[(size_t)j] + cnt1[(size_t)j];
        usable[(size_t)j] = mn >= kMinMinor && (double)mn >= kMinMinorFrac * (double)tot;
    }

    // site -> (read, allele) lists of the phasing reads over usable sites
    std::vector<int64_t> soff((size_t)n_sites + 1, 0);
    for (int64_t j = 0; j < n_sites; j++) soff[(size_t)j + 1] = soff[(size_t)j] + (usable[(size_t)j] ? cnt0[(size_t)j] + cnt1[(size_t)j] : 0);
    std::vector<int32_t> sread((size_t)soff[(size_t)n_sites]);
    std::vector<uint8_t> sall((size_t)soff[(size_t)n_sites]);
    // pairwise linkage: cis = reads showing the same allele index at both sites, trans = different
    std::vector<int32_t> cis((size_t)n_sites * kPairSpan, 0), trans((size_t)n_sites * kPairSpan, 0);
    {
        std::vector<int64_t> fill(soff.begin(), soff.end() - 1);
        std::vector<std::pair<int64_t, uint8_t>> mine;
        for (int64_t r = 0; r < n_reads; r++) {
            if (!read_use[r]) continue;
            mine.clear();
            for (int64_t e = pair_off[r]; e < pair_off[r + 1]; e++) {
                if (allele[e] == NC_PHASE_NONE) continue;
                const int64_t j = first_site[r] + (e - pair_off[r]);
                if (!usable[(size_t)j]) continue;
                sread[(size_t)fill[(size_t)j]] = (int32_t)r; sall[(size_t)fill[(size_t)j]] = allele[e]; fill[(size_t)j]++;
                mine.emplace_back(j, allele[e]);
            }
            for (size_t a = 0; a < mine.size(); a++)
                for (size_t b = a + 1; b < mine.size() && mine[b].first - mine[a].first <= kPairSpan; b++) {
                    const size_t k = (size_t)mine[a].first * kPairSpan + (size_t)(mine[b].first - mine[a].first - 1);
                    if (mine[a].second == mine[b].second) cis[k]++; else trans[k]++;
                }
        }
    }
    // maximum spanning forest over the accepted links (strongest first), union-find with parity: par[j] = orientation of j
    // relative to its root
    struct Edge { int32_t s, i, j; uint8_t flip; };
    std::vector<Edge> edges;
    for (int64_t i = 0; i < n_sites; i++)
        for (int d = 0; d < kPairSpan && i + d + 1 < n_sites; d++) {
            const int32_t c = cis[(size_t)i * kPairSpan + d], t = trans[(size_t)i * kPairSpan + d];
            const int32_t s = std::abs(c - t);
            if (s >= kMinLink && (double)s >= kMinLinkFrac * (double)(c + t)) edges.push_back({s, (int32_t)i, (int32_t)(i + d + 1), (uint8_t)(t > c)});
        }
    std::stable_sort(edges.begin(), edges.end(), [](const Edge& a, const Edge& b) { return a.s > b.s; });
    std::vector<int32_t> root((size_t)n_sites);
    std::iota(root.begin(), root.end(), 0);
    std::vector<uint8_t> par((size_t)n_sites, 0);
    auto find = [&](int32_t x, uint8_t& p) {             // -> root of x; p = parity of x relative to that root (with path compression)
        int32_t r = x; uint8_t acc = 0;
        while (root[(size_t)r] != r) { acc ^= par[(size_t)r]; r = root[(size_t)r]; }
        int32_t y = x; uint8_t py = acc;
        while (root[(size_t)y] != y) { const int32_t nxt = root[(size_t)y]; const uint8_t pn = py ^ par[(size_t)y]; root[(size_t)y] = r; par[(size_t)y] = py; y = nxt; py = pn; }
        p = acc;
        return r;
    };
    for (const Edge& e : edges) {
        uint8_t pi, pj;
        const int32_t ri = find(e.i, pi), rj = find(e.j, pj);
        if (ri == rj) continue;
        const uint8_t rel = pi ^ pj ^ e.flip;            // parity between the two roots
        if (ri < rj) { root[(size_t)rj] = ri; par[(size_t)rj] = rel; } else { root[(size_t)ri] = rj; par[(size_t)ri] = rel; }   // root = smallest index
    }
    std::vector<int32_t> bsize((size_t)n_sites, 0), blk((size_t)n_sites, -1);
    std::vector<int8_t> h((size_t)n_sites, -1);
    for (int64_t j = 0; j < n_sites; j++) { uint8_t p; blk[(size_t)j] = find((int32_t)j, p); bsize[(size_t)blk[(size_t)j]]++; }
    for (int64_t j = 0; j < n_sites; j++) {
        if (!usable[(size_t)j] || bsize[(size_t)blk[(size_t)j]] < 2) continue;
        uint8_t p; find((int32_t)j, p);
        h[(size_t)j] = (int8_t)p;                        // the block's first site has orientation 0
    }
    // read scores over the phased sites, then refinement sweeps: every site is re-decided against the scores its reads have
    // WITHOUT that site (a local search on the minimum-error-correction objective); a flip updates the scores at once
    std::vector<int32_t> score((size_t)n_reads, 0);
    for (int64_t j = 0; j < n_sites; j++) {
        if (h[(size_t)j] < 0) continue;
        for (int64_t e = soff[(size_t)j]; e < soff[(size_t)j + 1]; e++)
            score[(size_t)sread[(size_t)e]] += sall[(size_t)e] == (uint8_t)h[(size_t)j] ? 1 : -1;
    }
    for (int it = 0; it < iterations; it++) {
        int64_t flips = 0;
        for (int64_t j = 0; j < n_sites; j++) {
            if (h[(size_t)j] < 0) continue;
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
    // keep the convention "first site of a block is A|B" after the sweeps
    std::vector<int8_t> first_h((size_t)n_sites, -1);
    for (int64_t j = 0; j < n_sites; j++)
        if (h[(size_t)j] >= 0 && first_h[(size_t)blk[(size_t)j]] < 0) first_h[(size_t)blk[(size_t)j]] = h[(size_t)j];
    std::vector<int32_t> first_site_of((size_t)n_sites, -1);
    for (int64_t j = 0; j < n_sites; j++) {
        if (h[(size_t)j] < 0) continue;
        const int32_t b = blk[(size_t)j];
        if (first_site_of[(size_t)b] < 0) first_site_of[(size_t)b] = (int32_t)j;
        site_hap[j] = (int8_t)(h[(size_t)j] ^ first_h[(size_t)b]);
        site_block[j] = first_site_of[(size_t)b];
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
