// Synthetic long-read world generator (test / bench infrastructure, host only).
//
// Produces a seeded reference contig, a diploid (or haploid) truth set and a coordinate-sorted
// read set in BAM-native encoding (CIGAR u32 words, 4-bit packed bases) following SURVEY.md
// §8(d): uniform ACGT contig, het SNP every ~het_every bp, hom-alt every ~hom_every bp,
// a fraction of "systematic error" positions whose alt fraction is unlinked to haplotype,
// log-normal read lengths, per-base sub/del/ins errors, CIGAR ops M/I/D/S only.
//
// Everything is a pure function of (seed, position) or (seed, read index), so generation is
// parallel over reads and bit-reproducible.  The same arrays feed the CPU oracle, the golden
// fixtures and the device stager.
#include <cstdint>
#include <cstring>
#include <cmath>
#include <thread>
#include <vector>
#include <algorithm>

extern "C" {

struct NcSynthParams {
    uint64_t seed;
    int64_t  contig_len;
    double   coverage;
    double   len_median, len_sigma;
    int32_t  len_min, len_max;
    double   sub_rate, del_rate, ins_rate;
    int32_t  het_every, hom_every;   // 0 disables
    int32_t  sys_per_10k;            // systematic-error positions per 10,000
    double   clip_prob;
    int32_t  clip_max;
    int32_t  ploidy;                 // 2 diploid, 1 haploid (all variants hom)
    int32_t  mask_every, mask_len;   // lower-case (soft-masked) runs in the reference; 0 disables
    double   junk_frac;              // reads flagged secondary/suppl/dup/qcfail/unmapped
    double   nbase_rate;             // read bases emitted as N
    int32_t  indel_every;            // truth indels (het/hom alternate) every ~indel_every bp; 0 disables
    int32_t  indel_maxlen;
    double   untagged_frac;          // reads without an HP tag
};

}  // extern "C"

namespace {

inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    inline uint64_t next() { s += 0x9E3779B97F4A7C15ULL; uint64_t z = s;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31); }
    inline double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

const char kBases[4] = {'A', 'C', 'G', 'T'};
const uint8_t kNib[4] = {1, 2, 4, 8};  // BAM 4-bit codes of A,C,G,T

inline int base_idx(uint8_t c) {
    switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1;
                 case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return -1; }
}

// var byte: bits0-1 kind (0 none, 1 het, 2 hom, 3 sys) | bits2-3 alt base idx | bits4-7 sys quantile
// indel truth lives in a second byte array: bits0-5 length (1..50, 0 none) | bit6 ins(1)/del(0) | bit7 hom
inline uint32_t thr32(double r) { return (uint32_t)std::min(4294967295.0, r * 4294967296.0); }

struct ReadPlan { int64_t start; int64_t span; int hap; uint16_t flag; int lclip, rclip; uint64_t stream; int tagged; };

int64_t num_reads(const NcSynthParams& P) {
    // expected clipped log-normal span from a fixed deterministic sample of the length stream
    Rng r(mix64(P.seed ^ 0x1234567ULL));
    double acc = 0; const int NS = 65536;
    for (int i = 0; i < NS; i++) {
        double u1 = r.uni(), u2 = r.uni();
        double z = std::sqrt(-2.0 * std::log(u1 + 1e-300)) * std::cos(6.283185307179586 * u2);
        double len = std::exp(std::log(P.len_median) + P.len_sigma * z);
        len = std::min((double)P.len_max, std::max((double)P.len_min, len));
        acc += len;
    }
    double mean = acc / NS;
    int64_t n = (int64_t)std::ceil(P.coverage * (double)P.contig_len / mean);
    return std::max<int64_t>(n, 1);
}

ReadPlan plan_read(const NcSynthParams& P, int64_t i, int64_t n) {
    ReadPlan rp;
    Rng r(mix64(P.seed ^ mix64((uint64_t)i * 2 + 1)));
    int64_t lo = (int64_t)(((__int128)i * P.contig_len) / n);
    int64_t hi = (int64_t)(((__int128)(i + 1) * P.contig_len) / n);
    int64_t w = std::max<int64_t>(1, hi - lo);
    rp.start = std::min<int64_t>(P.contig_len - 1, lo + (int64_t)(r.next() % (uint64_t)w));
    double u1 = r.uni(), u2 = r.uni();
    double z = std::sqrt(-2.0 * std::log(u1 + 1e-300)) * std::cos(6.283185307179586 * u2);
    double len = std::exp(std::log(P.len_median) + P.len_sigma * z);
    len = std::min((double)P.len_max, std::max((double)P.len_min, len));
    rp.span = std::min<int64_t>((int64_t)len, P.contig_len - rp.start);
    if (rp.span < 1) rp.span = 1;
    uint64_t b = r.next();
    rp.hap = (int)(b & 1);
    rp.flag = (b & 2) ? 0x10 : 0;
    double j = r.uni();
    if (j < P.junk_frac) {
        static const uint16_t junk[5] = {0x100, 0x800, 0x400, 0x200, 0x4};
        rp.flag |= junk[(b >> 8) % 5];
    }
    rp.lclip = rp.rclip = 0;
    if (P.clip_max > 0) {
        if (r.uni() < P.clip_prob) rp.lclip = 1 + (int)(r.next() % (uint64_t)P.clip_max);
        if (r.uni() < P.clip_prob) rp.rclip = 1 + (int)(r.next() % (uint64_t)P.clip_max);
    }
    rp.tagged = r.uni() >= P.untagged_frac;
    rp.stream = r.next();
    return rp;
}

// Sink that either counts or writes CIGAR words and packed bases.
struct Sink {
    uint32_t* cig; uint8_t* seq; int64_t ncig = 0, nseq = 0;
    int cur_op = -1; uint32_t cur_len = 0;
    Sink(uint32_t* c, uint8_t* s) : cig(c), seq(s) {}
    inline void flush() { if (cur_op >= 0 && cur_len) { if (cig) cig[ncig] = (cur_len << 4) | (uint32_t)cur_op; ncig++; } cur_op = -1; cur_len = 0; }
    inline void op(int o, uint32_t l) { if (!l) return; if (o == cur_op) cur_len += l; else { flush(); cur_op = o; cur_len = l; } }
    inline void base(uint8_t nib) {
        if (seq) { if (nseq & 1) seq[nseq >> 1] |= nib; else seq[nseq >> 1] = (uint8_t)(nib << 4); }
        nseq++;
    }
};

enum { OP_M = 0, OP_I = 1, OP_D = 2, OP_S = 4 };

void gen_read(const NcSynthParams& P, const uint8_t* ref, const uint8_t* var, const uint8_t* indel,
              const ReadPlan& rp, Sink& out) {
    Rng r(rp.stream);
    const uint32_t t_del = thr32(P.del_rate), t_ins = thr32(P.ins_rate), t_sub = thr32(P.sub_rate),
                   t_n = thr32(P.nbase_rate);
    for (int k = 0; k < rp.lclip; k++) out.base(kNib[r.next() & 3]);
    out.op(OP_S, (uint32_t)rp.lclip);
    const int64_t end = rp.start + rp.span;
    int64_t p = rp.start;
    while (p < end) {
        const bool edge = (p == rp.start) || (p == end - 1);
        uint64_t u = r.next();
        uint32_t ua = (uint32_t)u, ub = (uint32_t)(u >> 32);
        // truth indel anchored at p (applies after base p): handled below
        bool deleted = (!edge) && (ua < t_del);
        if (deleted) {
            out.op(OP_D, 1);
        } else {
            uint8_t v = var[p];
            int kind = v & 3, alt = (v >> 2) & 3;
            int b = base_idx(ref[p]);
            if (b < 0) b = (int)(u >> 20) & 3;  // N in reference: emit a random base
            if (kind == 2 || (kind == 1 && (rp.hap == 1 || P.ploidy == 1))) b = alt;
            else if (kind == 3) {
                uint32_t q = (v >> 4) & 15;
                double f = 0.15 + 0.20 * ((double)q + 0.5) / 16.0;
                if ((double)(r.next() >> 11) * (1.0 / 9007199254740992.0) < f) b = alt;
            }
            uint64_t u2 = r.next();
            if ((uint32_t)u2 < t_sub) b = (b + 1 + (int)((u2 >> 32) % 3)) & 3;
            uint8_t nib = kNib[b];
            if (t_n && (uint32_t)(u2 >> 16) < t_n && !edge) nib = 15;
            out.op(OP_M, 1);
            out.base(nib);
            if (!edge && ub < t_ins) {
                uint64_t u3 = r.next();
                int il = 1 + (int)((u3 & 7) == 0) + (int)((u3 & 63) == 0);
                for (int k = 0; k < il; k++) out.base(kNib[(u3 >> (8 + 2 * k)) & 3]);
                out.op(OP_I, (uint32_t)il);
            }
        }
        // truth indel after p
        if (indel && !edge) {
            uint8_t iv = indel[p];
            int L = iv & 63;
            if (L) {
                bool hom = (iv & 0x80) || P.ploidy == 1;
                if (hom || rp.hap == 1) {
                    if (iv & 0x40) {  // insertion of L bases derived from position hash
                        uint64_t h = mix64(P.seed ^ 0xABCDEFULL ^ (uint64_t)p);
                        for (int k = 0; k < L; k++) { if ((k & 31) == 31) h = mix64(h); out.base(kNib[(h >> (2 * (k & 31))) & 3]); }
                        out.op(OP_I, (uint32_t)L);
                    } else {          // deletion of the next L reference bases
                        int64_t dl = std::min<int64_t>(L, end - 1 - (p + 1));
                        if (dl > 0) { out.op(OP_D, (uint32_t)dl); p += dl; }
                    }
                }
            }
        }
        p++;
    }
    for (int k = 0; k < rp.rclip; k++) out.base(kNib[r.next() & 3]);
    out.op(OP_S, (uint32_t)rp.rclip);
    out.flush();
}

template <class F>
void parallel_for(int64_t n, int nthreads, F f) {
    nthreads = std::max(1, nthreads);
    if (nthreads == 1 || n < 64) { f(0, n); return; }
    std::vector<std::thread> th;
    int64_t per = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++) {
        int64_t a = t * per, b = std::min(n, a + per);
        if (a >= b) break;
        th.emplace_back([=]() { f(a, b); });
    }
    for (auto& x : th) x.join();
}

}  // namespace

extern "C" {

// Fill reference (ASCII), SNP truth bytes and indel truth bytes (indel may be NULL).
int nc_synth_world(const NcSynthParams* Pp, uint8_t* ref, uint8_t* var, uint8_t* indel, int nthreads) {
    const NcSynthParams P = *Pp;
    const int64_t L = P.contig_len;
    parallel_for(L, nthreads, [&](int64_t a, int64_t b) {
        for (int64_t p = a; p < b; p++) {
            uint64_t h = mix64(P.seed * 0x100000001B3ULL + (uint64_t)p);
            int rb = (int)(h & 3);
            char c = kBases[rb];
            if (P.mask_every > 0) {
                int64_t blk = p / P.mask_every;
                int64_t off = (int64_t)(mix64(P.seed ^ 0x5151ULL ^ (uint64_t)blk) % (uint64_t)P.mask_every);
                int64_t s = blk * P.mask_every + off;
                if (p >= s && p < s + P.mask_len) c = (char)(c + 32);
            }
            ref[p] = (uint8_t)c;
            uint8_t v = 0;
            int alt = (rb + 1 + (int)((h >> 8) % 3)) & 3;
            bool is_het = false, is_hom = false;
            if (P.het_every > 0) {
                int64_t k = p / P.het_every;
                is_het = (p == k * P.het_every + (int64_t)(mix64(P.seed ^ 0xA1A1ULL ^ (uint64_t)k) % (uint64_t)P.het_every));
            }
            if (P.hom_every > 0) {
                int64_t k = p / P.hom_every;
                is_hom = (p == k * P.hom_every + (int64_t)(mix64(P.seed ^ 0xB2B2ULL ^ (uint64_t)k) % (uint64_t)P.hom_every));
            }
            if (is_het) v = 1;
            else if (is_hom) v = 2;
            else if ((int)((h >> 16) % 10000) < P.sys_per_10k) v = 3;
            if (v) v |= (uint8_t)(alt << 2) | (uint8_t)(((h >> 40) & 15) << 4);
            var[p] = v;
            if (indel) {
                uint8_t iv = 0;
                if (P.indel_every > 0 && !v) {
                    int64_t k = p / P.indel_every;
                    uint64_t hk = mix64(P.seed ^ 0xC3C3ULL ^ (uint64_t)k);
                    if (p == k * P.indel_every + (int64_t)(hk % (uint64_t)P.indel_every)) {
                        int len = 1 + (int)((hk >> 20) % (uint64_t)std::max(1, P.indel_maxlen));
                        iv = (uint8_t)(len & 63) | (uint8_t)(((hk >> 32) & 1) ? 0x40 : 0) | (uint8_t)((k & 1) ? 0x80 : 0);
                    }
                }
                indel[p] = iv;
            }
        }
    });
    return 0;
}

int64_t nc_synth_num_reads(const NcSynthParams* P) { return num_reads(*P); }

// Pass 1: per-read header fields and sizes.
int nc_synth_count(const NcSynthParams* Pp, const uint8_t* ref, const uint8_t* var, const uint8_t* indel,
                   int64_t n, int32_t* pos, int32_t* n_cigar, int32_t* l_seq, uint16_t* flag,
                   int8_t* hap, int32_t* ref_span, int nthreads) {
    const NcSynthParams P = *Pp;
    parallel_for(n, nthreads, [&](int64_t a, int64_t b) {
        for (int64_t i = a; i < b; i++) {
            ReadPlan rp = plan_read(P, i, n);
            Sink s(nullptr, nullptr);
            gen_read(P, ref, var, indel, rp, s);
            pos[i] = (int32_t)rp.start; n_cigar[i] = (int32_t)s.ncig; l_seq[i] = (int32_t)s.nseq;
            flag[i] = rp.flag; hap[i] = (int8_t)(rp.tagged ? rp.hap + 1 : 0); ref_span[i] = (int32_t)rp.span;
        }
    });
    return 0;
}

// Pass 2: write CIGAR words and packed bases at the given offsets (cigar_off in words, seq_off in bytes).
int nc_synth_fill(const NcSynthParams* Pp, const uint8_t* ref, const uint8_t* var, const uint8_t* indel,
                  int64_t n, const int64_t* cigar_off, const int64_t* seq_off, uint32_t* cigar,
                  uint8_t* seq4, int nthreads) {
    const NcSynthParams P = *Pp;
    parallel_for(n, nthreads, [&](int64_t a, int64_t b) {
        for (int64_t i = a; i < b; i++) {
            ReadPlan rp = plan_read(P, i, n);
            Sink s(cigar + cigar_off[i], seq4 + seq_off[i]);
            gen_read(P, ref, var, indel, rp, s);
        }
    });
    return 0;
}

}  // extern "C"
