// Indel feature path (diploid): candidate scan, read slices, star alignment, [5][128][2] tensors, consensus.
//
// Reference being replaced: nanocaller_src/generate_indel_pileups.py
//   I1 scan      per-haplotype indel-event windows over the pileup columns, thresholds, `prev` suppression (:213-275)
//   I2 slices    query_sequence[query_position_or_next : +160|260] per read at a key position (:306-338)
//   I3 msa       MUSCLE is an external binary (:30); this library aligns every slice to the reference window with its
//                own, fully specified star alignment (oracle/star_msa.py states it) and builds the column-frequency
//                tensors and the consensus exactly as `msa` does (:54-71)
//   I1 impute   impute_indel_phase (:278-304): columns without phased coverage on both haplotypes whose pileup strings carry
//                enough indel marks are split into two read sets by grouping equal strings; pass 2 then uses those sets in
//                place of the HP tags (:309-313)
#pragma once
#include "nc_common.cuh"
#include "nc_pileup.cuh"

namespace nc {

struct IndelChunk { int32_t lo, hi; int64_t grank_lo; int64_t rank_off; int32_t n_em; int32_t pad; };   // 0-based [lo, hi)

// ------------------------------------------------------------------------------------------------
// I1.1 per-position read depth of haplotype 1, haplotype 2 and all admitted reads (aligned rows, like K1)
// ------------------------------------------------------------------------------------------------
struct DepthArgs {
    int64_t n_reads;
    const int32_t* pos; const int32_t* end; const uint16_t* flag; const int8_t* hp; const int32_t* pmaxend;
    const int64_t* rowoff; const int32_t* nwords; const uint32_t* rows;
    int32_t lo_al, lo, hi; uint32_t flag_filter;
    uint16_t* depth;           // [3][n_al]
    int64_t n_al;
};

__global__ void __launch_bounds__(kTileThreads) indel_depth_kernel(const DepthArgs a) {
    __shared__ int32_t s_posw[kTileThreads];
    __shared__ int32_t s_nw[kTileThreads];
    __shared__ int64_t s_rowoff[kTileThreads];
    __shared__ int8_t s_hp[kTileThreads];
    __shared__ uint16_t s_acc[3][kTilePos];
    __shared__ int32_t s_cnt;
    __shared__ int64_t s_ilo, s_ihi;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int32_t P0 = a.lo_al + kTilePos * (int32_t)blockIdx.x;
    const int32_t P1 = min(P0 + kTilePos, a.hi);
    for (int i = tid; i < 3 * kTilePos; i += kTileThreads) (&s_acc[0][0])[i] = 0;
    if (tid == 0) s_cnt = 0;
    if (w == 0) { const int64_t r = warp_bound_i32(a.pos, a.n_reads, P1 - 1, true, lane); if (lane == 0) s_ihi = r; }
    if (w == 1) { const int64_t r = warp_bound_i32(a.pmaxend, a.n_reads, P0, true, lane); if (lane == 0) s_ilo = r; }
    __syncthreads();
    const int64_t ilo = s_ilo, ihi = s_ihi;
    const int32_t Wbase = (P0 >> 3) + 32 * w, myw = Wbase + lane;
    uint32_t c[3][2] = {{0, 0}, {0, 0}, {0, 0}};
    int pending = 0;
    const int sbase = w * 256 + lane * 8;
    auto flush = [&]() {
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int k = 0; k < 3; k++) {
                s_acc[k][sbase + 2 * j] += (c[k][0] >> (8 * j)) & 255u;
                s_acc[k][sbase + 2 * j + 1] += (c[k][1] >> (8 * j)) & 255u;
            }
#pragma unroll
        for (int k = 0; k < 3; k++) c[k][0] = c[k][1] = 0;
        pending = 0;
    };
    for (int64_t base = ilo; base < ihi; base += kTileThreads) {
        const int64_t i = base + tid;
        if (i < ihi) {
            const int32_t p = __ldg(a.pos + i), e = __ldg(a.end + i);
            const uint32_t f = __ldg(a.flag + i);
            if ((f & a.flag_filter) == 0 && e > P0 && p < P1 && e > p) {
                const int slot = atomicAdd(&s_cnt, 1);
                s_posw[slot] = p >> 3; s_nw[slot] = __ldg(a.nwords + i); s_rowoff[slot] = __ldg(a.rowoff + i); s_hp[slot] = __ldg(a.hp + i);
            }
        }
        __syncthreads();
        const int cnt = s_cnt;
        for (int j = 0; j < cnt; j++) {
            const int32_t pw = s_posw[j], nw = s_nw[j];
            if (Wbase + 31 < pw || Wbase >= pw + nw) continue;
            const int32_t rel = myw - pw;
            uint32_t word = 0xFFFFFFFFu;
            if (rel >= 0 && rel < nw) word = __ldg(a.rows + s_rowoff[j] + rel);
            const uint32_t cov = ((word >> 3) & 0x11111111u) ^ 0x11111111u;
            const uint32_t e0 = cov & 0x01010101u, e1 = (cov >> 4) & 0x01010101u;
            c[2][0] += e0; c[2][1] += e1;
            const int h = s_hp[j];
            if (h == 1) { c[0][0] += e0; c[0][1] += e1; }
            else if (h == 2) { c[1][0] += e0; c[1][1] += e1; }
            if (++pending == 255) flush();
        }
        __syncthreads();
        if (tid == 0) s_cnt = 0;
        __syncthreads();
    }
    flush();
    const size_t o = (size_t)(P0 - a.lo_al) + sbase;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        uint16_t v[8];
#pragma unroll
        for (int t = 0; t < 8; t++) v[t] = (P0 + sbase + t >= a.lo && P0 + sbase + t < a.hi) ? s_acc[k][sbase + t] : (uint16_t)0;
        *reinterpret_cast<uint4*>(a.depth + (size_t)k * a.n_al + o) = *reinterpret_cast<const uint4*>(v);
    }
}

// emitted[p] = the pileup yields a column (depth > 0) that is not excluded (:218)
__global__ void indel_emitted_kernel(const uint16_t* __restrict__ depth_all, int64_t n, int32_t lo_al, const int32_t* __restrict__ bed,
                                     int32_t n_bed, int32_t* __restrict__ em) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int e = depth_all[i] > 0;
    if (e && n_bed > 0 && bed_excluded(bed, n_bed, lo_al + (int32_t)i + 1)) e = 0;
    em[i] = e;
}
// global rank -> position of the emitted column
__global__ void indel_empos_kernel(const int32_t* __restrict__ em, const int64_t* __restrict__ grank, int64_t n, int32_t lo_al,
                                   int32_t* __restrict__ em_pos) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && em[i]) em_pos[grank[i]] = lo_al + (int32_t)i;
}
__global__ void indel_chunk_kernel(const NcChunk* __restrict__ chunks, int32_t n_chunks, int32_t lo_al, int32_t ref_lo, int32_t ref_hi,
                                   const int64_t* __restrict__ grank, IndelChunk* __restrict__ out, int32_t* __restrict__ n_em1) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    IndelChunk ic;
    ic.lo = max(max(0, chunks[c].start - 1), ref_lo);
    ic.hi = max(ic.lo, min(chunks[c].end, ref_hi));
    ic.grank_lo = grank[ic.lo - lo_al];
    ic.n_em = (int32_t)(grank[ic.hi - lo_al] - ic.grank_lo);
    ic.rank_off = 0; ic.pad = 0;
    out[c] = ic;
    n_em1[c] = ic.n_em + 1;
}
__global__ void indel_chunk_off_kernel(IndelChunk* __restrict__ ch, int32_t n_chunks, const int64_t* __restrict__ off) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n_chunks) ch[c].rank_off = off[c];
}

// ------------------------------------------------------------------------------------------------
// I1.2 indel events of every haplotagged read -> difference arrays of the window unions.
// A read with a qualifying event at emitted column rank r is a member of the union of the last `win` columns at
// ranks r .. r+win-1; overlapping intervals of one read are merged so it is counted once (a set union).
// kinds: key = hap*4 + {0 del 2<L<=50, 1 del L<=10, 2 ins 2<L<=50, 3 ins L<=10}.
// ------------------------------------------------------------------------------------------------
// emitted flags + ranks in 8 bytes per 32 positions: {bit i = position 32 b + i is emitted, rank of position 32 b}.  The event kernel
// looks both up for every indel of every read — tens of millions of scattered reads; from the int32 flag and the int64 rank arrays
// that was two 32-byte sectors per event, from the pairs it is one sector per ~128 positions of a read.
__global__ void indel_empairs_kernel(const int32_t* __restrict__ em, const int64_t* __restrict__ grank, int64_t n, uint2* __restrict__ pairs) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t bits = __ballot_sync(0xffffffffu, i < n && em[i] != 0);
    if ((threadIdx.x & 31) == 0 && i < n) pairs[i >> 5] = make_uint2(bits, (uint32_t)grank[i]);
}

struct EventArgs {
    int64_t n_reads;
    const int32_t* pos; const int32_t* end; const uint16_t* flag; const int8_t* hp;
    const int64_t* cigar_off; const uint32_t* cigar; const int2* opstart;
    const IndelChunk* chunks; int32_t n_chunks;
    const uint2* empairs; int32_t lo_al;     // indel_empairs_kernel
    uint32_t flag_filter; int32_t win, small_win, haploid;
    int32_t* diff; int64_t R;          // [8][R]
};

// One WARP per read, lanes over its CIGAR operations (coalesced; the reference offsets of the operations come from K0's cigar scan).
// The union of a read's window intervals [r, min(n_em, r + win)] per kind is what one sequential pass would merge: with the events in
// column order the interval ends are monotone, so event i opens a new interval iff its rank exceeds the end of the event before it —
// a test against the previous event of the same kind (the lane below in the ballot, or the carry from the previous 32 operations).
// (The first version ran one THREAD per read: 5.5 of 32 lanes active, every CIGAR word its own 32-byte sector — 4.6 GB of DRAM traffic
// for a 20 Mb contig.)
__global__ void __launch_bounds__(128) indel_events_kernel(const EventArgs a) {
    const uint32_t full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= a.n_reads) return;
    const int h = a.haploid ? 0 : a.hp[r] - 1;           // haploid caller: one window set over all reads
    if (h < 0 || h > 1 || (a.flag[r] & a.flag_filter) != 0) return;
    const int32_t rp = a.pos[r], re = a.end[r];
    if (re <= rp) return;
    // chunks overlapping [rp, re): chunk los are ascending
    int c0;
    { int lo = 0, hi = a.n_chunks; while (lo < hi) { int mid = (lo + hi) >> 1; if (a.chunks[mid].hi <= rp) lo = mid + 1; else hi = mid; } c0 = lo; }
    const int64_t k0 = a.cigar_off[r], k1 = a.cigar_off[r + 1];
    for (int c = c0; c < a.n_chunks && a.chunks[c].lo < re; c++) {
        const IndelChunk ch = a.chunks[c];
        int32_t* dbase = a.diff + (int64_t)(h * 4) * a.R + ch.rank_off;
        bool have[4] = {false, false, false, false};
        int32_t last_b[4] = {0, 0, 0, 0};
        for (int64_t kb = k0; kb < k1; kb += 32) {
            const int64_t k = kb + lane;
            int kA = -1, kB = -1;
            int32_t rk = 0;
            bool past = false;
            if (k < k1) {
                const uint32_t cw = __ldg(a.cigar + k);
                const int32_t rl = cig_ref_len(cw);
                const int32_t x = rp + __ldg(&a.opstart[k].x);
                past = x >= ch.hi;
                const int32_t plast = x + rl - 1;
                if (rl > 0 && k + 1 < k1 && plast >= ch.lo && plast < ch.hi) {
                    const uint32_t op = cw & 15u, w1 = __ldg(a.cigar + k + 1), op2 = w1 & 15u;
                    int32_t tot = 0; bool is_del = false;
                    if (op2 == 2u && op != 2u) {
                        is_del = true;
                        tot = (int32_t)(w1 >> 4);
                        for (int64_t j = k + 2; j < k1; j++) {
                            const uint32_t w2 = __ldg(a.cigar + j), o2 = w2 & 15u;
                            if (o2 == 2u) tot += (int32_t)(w2 >> 4);
                            else if (o2 == 1u || o2 == 4u || o2 == 0u || o2 == 7u || o2 == 8u) break;
                        }
                    } else if (op2 == 1u || (op2 == 6u && k + 2 < k1)) {
                        for (int64_t j = k + 1; j < k1; j++) {
                            const uint32_t w2 = __ldg(a.cigar + j), o2 = w2 & 15u;
                            if (o2 == 1u) tot += (int32_t)(w2 >> 4);
                            else if (o2 != 6u) break;
                        }
                    }
                    if (tot > 0) {
                        const int64_t pi = (int64_t)plast - a.lo_al;
                        const uint2 pr = __ldg(a.empairs + (pi >> 5));
                        if ((pr.x >> (pi & 31)) & 1u) {
                            rk = (int32_t)((int64_t)pr.y + __popc(pr.x & ((1u << (pi & 31)) - 1u)) - ch.grank_lo);
                            const int base = is_del ? 0 : 2;
                            if (tot > 2 && tot <= 50) kA = base;
                            if (tot <= 10) kB = base + 1;
                        }
                    }
                }
            }
#pragma unroll
            for (int kind = 0; kind < 4; kind++) {
                const bool mine = kA == kind || kB == kind;
                const uint32_t bm = __ballot_sync(full, mine);
                if (bm == 0) continue;                                  // warp-uniform
                const int32_t win = (kind & 1) ? a.small_win : a.win;
                const int32_t b_mine = min(ch.n_em, rk + win);
                const uint32_t below = bm & ((1u << lane) - 1u);
                const int src = (mine && below) ? 31 - __clz(below) : lane;
                const int32_t b_below = __shfl_sync(full, b_mine, src);
                if (mine) {
                    const bool has_prev = below != 0u || have[kind];
                    const int32_t b_prev = below ? b_below : last_b[kind];
                    if (!has_prev) atomicAdd(dbase + (int64_t)kind * a.R + rk, 1);
                    else if (rk > b_prev) { atomicAdd(dbase + (int64_t)kind * a.R + b_prev, -1); atomicAdd(dbase + (int64_t)kind * a.R + rk, 1); }
                }
                last_b[kind] = __shfl_sync(full, b_mine, 31 - __clz(bm));
                have[kind] = true;
            }
            if (__all_sync(full, past || k >= k1)) break;             // every operation of this block starts behind the chunk
        }
        if (lane == 0) {
#pragma unroll
            for (int kind = 0; kind < 4; kind++)
                if (have[kind]) atomicAdd(dbase + (int64_t)kind * a.R + last_b[kind], -1);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// I1.3 threshold tests per emitted column (:252-275): hit = 1 large-window condition, 2 small-window condition
// ------------------------------------------------------------------------------------------------
struct DecideArgs {
    const IndelChunk* chunks; int32_t n_chunks; int64_t R;
    const int32_t* diff;       // [8][R] difference arrays of the window unions: every interval adds +1 and -1 inside its own chunk's rank range
    const int32_t* em_pos; const uint16_t* depth; int64_t n_al; int32_t lo_al;
    int32_t mincov, haploid; double ins_t, del_t;
    uint8_t* hit; unsigned long long* n_hits;
    // impute_indel_phase (:278-285): per-column indel marks of all reads; pending columns get hit 4 and are counted in n_hits[2]
    int32_t impute; const int32_t* cdel; const int32_t* cins;
};
// One CTA per chunk walks the chunk's emitted columns in order, 256 at a time, and turns the difference arrays into window counts on the
// way: four kinds ride in the 16-bit fields of one 64-bit word (packing is linear, and every complete prefix is a count >= 0, so the
// fields of the scanned word ARE the counts), i.e. two block scans per 256 columns with the running totals carried in registers.  The
// counts never exist in memory: a global scan of the eight arrays into 64-bit prefixes, read back by a thread per column, moved 24
// bytes per (kind, column) where this reads 4 (scan 14.6 + decide 9.8 ms for configs[2]).
__global__ void __launch_bounds__(256) indel_decide_kernel(const DecideArgs a) {
    __shared__ long long s_tot[2][8];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const IndelChunk ch = a.chunks[blockIdx.x];
    const int32_t n1 = ch.n_em + 1;                                   // the chunk's rank range has one slot past its last column (interval ends)
    long long carry0 = 0, carry1 = 0;
    // Four tiles of 256 columns per trip: all their loads are issued first (difference arrays and positions, then the depths behind the
    // positions), so a trip costs two memory round trips instead of eight — with one CTA per 100 kb chunk a small contig has only a few
    // CTAs per SM and the walk is bound by exactly that latency.
    for (int32_t r0 = 0; r0 < n1; r0 += 1024) {
        long long q0[4], q1[4];
        int64_t pi[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int32_t r = r0 + 256 * u + tid;
            const int64_t g = ch.rank_off + r;
            q0[u] = 0; q1[u] = 0; pi[u] = -1;
            if (r < n1) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    q0[u] += (long long)__ldg(a.diff + (int64_t)k * a.R + g) * (1ll << (16 * k));
                    q1[u] += (long long)__ldg(a.diff + (int64_t)(4 + k) * a.R + g) * (1ll << (16 * k));
                }
                if (r < ch.n_em) pi[u] = (int64_t)__ldg(a.em_pos + ch.grank_lo + r) - a.lo_al;
            }
        }
        int32_t l0[4], l1[4], lt[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            l0[u] = l1[u] = lt[u] = 0;
            if (pi[u] >= 0) {
                if (a.haploid) { l0[u] = l1[u] = a.depth[2 * a.n_al + pi[u]]; }
                else { l0[u] = a.depth[pi[u]]; l1[u] = a.depth[a.n_al + pi[u]]; if (a.impute) lt[u] = a.depth[2 * a.n_al + pi[u]]; }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int32_t r = r0 + 256 * u + tid;
            if (r0 + 256 * u >= n1) break;                            // uniform
            long long p0 = q0[u], p1 = q1[u];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const long long t0 = __shfl_up_sync(0xffffffffu, p0, d), t1 = __shfl_up_sync(0xffffffffu, p1, d);
                if (lane >= d) { p0 += t0; p1 += t1; }
            }
            __syncthreads();                                          // the previous tile's totals have been read
            if (lane == 31) { s_tot[0][wid] = p0; s_tot[1][wid] = p1; }
            __syncthreads();
            long long tile0 = 0, tile1 = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) {
                const long long t0 = s_tot[0][w], t1 = s_tot[1][w];
                if (w < wid) { p0 += t0; p1 += t1; }
                tile0 += t0; tile1 += t1;
            }
            p0 += carry0; p1 += carry1;
            carry0 += tile0; carry1 += tile1;
            if (r >= n1) continue;
            uint8_t hit = 0;
            if (pi[u] >= 0) {
                if (a.haploid) {
                    if (l0[u] >= a.mincov) {                   // generate_indel_pileups_haploid.py:224-241
                        double f[4];
#pragma unroll
                        for (int k = 0; k < 4; k++) f[k] = l0[u] > 0 ? (double)((p0 >> (16 * k)) & 0xFFFF) / (double)l0[u] : 0.0;
                        if (f[0] >= a.del_t || f[2] >= a.ins_t) hit = 1;
                        else if (f[1] >= a.del_t || f[3] >= a.ins_t || (f[1] + f[3]) >= 0.9) hit = 2;
                    }
                } else if (l0[u] >= a.mincov && l1[u] >= a.mincov) {
                    double f[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        const int32_t l = k < 4 ? l0[u] : l1[u];
                        const long long pk = k < 4 ? p0 : p1;
                        f[k] = l > 0 ? (double)((pk >> (16 * (k & 3))) & 0xFFFF) / (double)l : 0.0;
                    }
                    if (fmax(f[0], f[4]) >= a.del_t || fmax(f[2], f[6]) >= a.ins_t) hit = 1;
                    else if (fmax(f[1], f[5]) >= a.del_t || fmax(f[3], f[7]) >= a.ins_t || (f[1] + f[3]) >= 0.9 || (f[5] + f[7]) >= 0.9) hit = 2;
                } else if (a.impute) {
                    if (lt[u] > 0 && lt[u] >= 2 * a.mincov) {
                        const double fd = (double)a.cdel[pi[u]] / (double)lt[u], fi = (double)a.cins[pi[u]] / (double)lt[u];
                        if (a.del_t <= fd || a.ins_t <= fi) hit = 4;
                    }
                }
            }
            a.hit[ch.rank_off + r] = hit;
            if (hit == 4) atomicAdd(a.n_hits + 2, 1ull);
            else if (hit) atomicAdd(a.n_hits, 1ull);
        }
    }
}

// I1.4 greedy pass with `prev` (:249,267,273): one warp per chunk walks its columns in order.  Hits are sparse (tens per 100 kb chunk)
// and the walk is one dependent step per 32 columns, so the time is load latency: a lane fetches the flags of eight 32-column groups
// at once, positions are only read for the hits (the per-chunk time does not depend on the contig, 1.65 ms before, for any size).
__global__ void indel_greedy_kernel(const IndelChunk* __restrict__ chunks, int32_t n_chunks, const uint8_t* __restrict__ hit,
                                    const int32_t* __restrict__ em_pos, int32_t win, NcIndelVariant* __restrict__ out,
                                    unsigned long long* __restrict__ n_out) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= n_chunks) return;
    const IndelChunk ch = chunks[c];
    int32_t prev = 0;
    for (int32_t r0 = 0; r0 < ch.n_em; r0 += 256) {
        int h[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { const int32_t r = r0 + 32 * k + lane; h[k] = r < ch.n_em ? (int)__ldg(hit + ch.rank_off + r) : 0; }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            uint32_t m = __ballot_sync(0xffffffffu, h[k] != 0);
            while (m) {
                const int l = __ffs(m) - 1;
                m &= m - 1;
                const int hl = __shfl_sync(0xffffffffu, h[k], l);
                const int32_t vl = __ldg(em_pos + ch.grank_lo + r0 + 32 * k + l) + 1;
                if (vl <= prev) continue;
                const int32_t back = hl == 1 ? win : 10;
                prev = vl + back;
                if (lane == 0) {
                    const unsigned long long slot = atomicAdd(n_out, 1ull);
                    NcIndelVariant nv; nv.key = max(1, vl - back); nv.type = hl == 1 ? 0 : 1; nv.chunk = c;
                    nv.src = hl == 3 ? vl : 0;             // extra_variants: the read sets come from this column (:302)
                    out[slot] = nv;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// I2 / I3 — per key position: covering reads, slices, alignment, tensors
// ------------------------------------------------------------------------------------------------
struct SiteArgs {
    int64_t n_reads;
    const int32_t* pos; const int32_t* end; const uint16_t* flag; const int8_t* hp; const int32_t* ps; const int32_t* pmaxend;
    const int64_t* cigar_off; const uint32_t* cigar; const int2* opstart;
    const int64_t* seq_off; const int32_t* l_seq; const uint8_t* seq4;
    const uint8_t* ref; int64_t ref_start, ref_len; int32_t contig_len;
    const NcIndelVariant* sites; int64_t n_sites;
    const NcChunk* chunks;
    uint32_t flag_filter; int32_t wa, win, mincov, maxcov;
    int32_t* site_m;           // reference window length per site (0 = skipped)
    int32_t* site_cnt;         // covering reads per site
    const int64_t* site_off;   // [n_sites+1] entry offsets
    int32_t* e_read; int32_t* e_qpn;
    int8_t* e_grp;             // read group of the entry: 0 / 1 (HP 1 / 2, or the imputed read sets of the site), -1 neither
    // imputed sites (:309-313): site_imp[s] = index of the site's source column in the imputed-column tables, or -1
    const int32_t* site_imp; const int64_t* imp_off; const int32_t* imp_read; const int8_t* imp_label;
    // alignment outputs, stride per entry
    uint8_t* e_slice; uint8_t* e_acode; uint16_t* e_inslen; uint16_t* e_insfirst; int32_t* e_n;
    int32_t nmax, mmax;
    // msa outputs
    float* tensors;            // [n_sites][3][5][128][2]
    uint8_t* cns;              // [n_sites][3][cmax]
    int32_t cmax;
    NcIndelSiteMeta* meta;
};

// window of BAM indices of reads that can cover p
__device__ __forceinline__ void read_window(const int32_t* pos, const int32_t* pmaxend, int64_t n, int32_t p, int64_t& ilo, int64_t& ihi) {
    ihi = upper_bound_i32_64(pos, n, p);
    int64_t lo = 0, hi = ihi;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (__ldg(pmaxend + mid) <= p) lo = mid + 1; else hi = mid; }
    ilo = lo;
}

// Indel annotation of the pileup string on the LAST column of reference-consuming op k (appendix C.4): > 0 inserted bases
// (*qn = query index of the first one), < 0 deleted bases, 0 none.  k1 = end of the read's CIGAR words.
__device__ __forceinline__ int32_t op_indel(const uint32_t* __restrict__ cigar, const int2* __restrict__ opstart, int64_t k, int64_t k1, int32_t* qn) {
    if (k + 1 >= k1) return 0;
    const uint32_t op = __ldg(cigar + k) & 15u, w1 = __ldg(cigar + k + 1), op2 = w1 & 15u;
    if (op2 == 2u && op != 2u) {
        int32_t tot = (int32_t)(w1 >> 4);
        for (int64_t j = k + 2; j < k1; j++) {
            const uint32_t w2 = __ldg(cigar + j), o2 = w2 & 15u;
            if (o2 == 2u) tot += (int32_t)(w2 >> 4);
            else if (o2 == 1u || o2 == 4u || o2 == 0u || o2 == 7u || o2 == 8u) break;
        }
        return -tot;
    }
    if (op2 == 1u || (op2 == 6u && k + 2 < k1)) {
        int32_t tot = 0; int64_t first = -1;
        for (int64_t j = k + 1; j < k1; j++) {
            const uint32_t w2 = __ldg(cigar + j), o2 = w2 & 15u;
            if (o2 == 1u) { if (first < 0) first = j; tot += (int32_t)(w2 >> 4); }
            else if (o2 != 6u) break;
        }
        if (tot > 0 && qn && opstart) *qn = __ldg(&opstart[first].y);
        return tot;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// I1 impute (:278-304).  (a) per column, over all admitted reads: cdel = '*' and '-' marks, cins = '+' marks among the
// first two characters of the pileup strings (:281-285).
// ------------------------------------------------------------------------------------------------
struct ImputeCountArgs {
    int64_t n_reads;
    const int32_t* pos; const int32_t* end; const uint16_t* flag;
    const int64_t* cigar_off; const uint32_t* cigar;
    int32_t lo_al, lo, hi; uint32_t flag_filter;
    int32_t* cdel; int32_t* cins;        // [n_al], zeroed
};
__global__ void indel_impute_count_kernel(const ImputeCountArgs a) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n_reads) return;
    if ((a.flag[r] & a.flag_filter) != 0) return;
    const int32_t rp = a.pos[r], re = a.end[r];
    if (re <= rp || re <= a.lo || rp >= a.hi) return;
    const int64_t k0 = a.cigar_off[r], k1 = a.cigar_off[r + 1];
    int32_t x = rp;
    for (int64_t k = k0; k < k1; k++) {
        const uint32_t cw = __ldg(a.cigar + k);
        const int32_t rl = cig_ref_len(cw);
        if (rl == 0) continue;
        if ((cw & 15u) == 2u)
            for (int32_t q = max(x, a.lo); q < min(x + rl, a.hi); q++) atomicAdd(a.cdel + (q - a.lo_al), 1);
        const int32_t plast = x + rl - 1;
        x += rl;
        if (plast >= a.hi) break;
        if (plast < a.lo) continue;
        const int32_t ind = op_indel(a.cigar, nullptr, k, k1, nullptr);
        if (ind < 0) atomicAdd(a.cdel + (plast - a.lo_al), 1);
        else if (ind > 0) atomicAdd(a.cins + (plast - a.lo_al), 1);
    }
}

// (b) the columns the decide kernel left pending (hit == 4) -> compact list
__global__ void indel_impute_collect_kernel(const uint8_t* __restrict__ hit, int64_t R, const IndelChunk* __restrict__ chunks, int32_t n_chunks,
                                            const int32_t* __restrict__ em_pos, int64_t* __restrict__ out_g, int32_t* __restrict__ out_col,
                                            unsigned long long* __restrict__ n_out) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= R || hit[g] != 4) return;
    int c;
    { int lo = 0, hi = n_chunks; while (lo < hi) { int mid = (lo + hi) >> 1; if (chunks[mid].rank_off <= g) lo = mid + 1; else hi = mid; } c = lo - 1; }
    const unsigned long long slot = atomicAdd(n_out, 1ull);
    out_g[slot] = g;
    out_col[slot] = em_pos[chunks[c].grank_lo + (g - chunks[c].rank_off)];
}

// (c) reads of a column in pileup (BAM) order with their pileup-string descriptors: warp per column, count / fill
struct ImputeArgs {
    int64_t n_reads;
    const int32_t* pos; const int32_t* end; const uint16_t* flag; const int32_t* pmaxend;
    const int64_t* cigar_off; const uint32_t* cigar; const int2* opstart;
    const int64_t* seq_off; const int32_t* l_seq; const uint8_t* seq4;
    uint32_t flag_filter; int32_t mincov;
    const int32_t* cols; int64_t n_cols;   // 0-based column positions
    int32_t* col_cnt; const int64_t* col_off;
    int32_t* e_read; uint8_t* e_ch; int32_t* e_ind; int32_t* e_qn;     // per (column, read) entry
    int32_t* e_gid; int32_t* g_rep; int32_t* g_cnt;                    // scratch of the grouping, entry-sized
    int8_t* e_label;                                                   // 0 read_names_0, 1 read_names_1, -1 neither
    int32_t* col_ok;
};
__device__ __forceinline__ uint32_t read_nibble(const uint8_t* __restrict__ sq, int32_t q) {
    const uint32_t b = __ldg(sq + (q >> 1));
    return (q & 1) ? (b & 15u) : (b >> 4);
}
template <bool FILL>
__global__ void __launch_bounds__(128) indel_impute_reads_kernel(const ImputeArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= a.n_cols) return;
    const int32_t p = a.cols[c];
    int64_t ilo, ihi;
    read_window(a.pos, a.pmaxend, a.n_reads, p, ilo, ihi);
    const int64_t off = FILL ? a.col_off[c] : 0;
    int32_t cnt = 0;
    for (int64_t base = ilo; base < ihi; base += 32) {
        const int64_t i = base + lane;
        bool cover = false;
        if (i < ihi) {
            const int32_t rp = __ldg(a.pos + i), re = __ldg(a.end + i);
            cover = (__ldg(a.flag + i) & a.flag_filter) == 0 && rp <= p && p < re;
        }
        const uint32_t cm = __ballot_sync(0xffffffffu, cover);
        if (FILL && cover) {
            const int64_t e = off + cnt + __popc(cm & ((1u << lane) - 1u));
            const int64_t c0 = a.cigar_off[i];
            const int32_t nops = (int32_t)(a.cigar_off[i + 1] - c0), os = p - __ldg(a.pos + i);
            int32_t lo = 0, hi = nops;
            while (lo < hi) { int32_t mid = (lo + hi) >> 1; if (__ldg(&a.opstart[c0 + mid].x) <= os) lo = mid + 1; else hi = mid; }
            const int32_t k = lo - 1;
            const uint32_t cw = __ldg(a.cigar + c0 + k);
            const int2 st = __ldg(a.opstart + c0 + k);
            uint32_t ch;                                               // 0..15 base nibble, 16 '*', 17 '>', 18 '<'
            if (cig_is_match(cw)) {
                const int32_t q = st.y + (os - st.x);
                ch = q < __ldg(a.l_seq + i) ? read_nibble(a.seq4 + __ldg(a.seq_off + i), q) : 15u;
            } else ch = (cw & 15u) == 2u ? 16u : ((__ldg(a.flag + i) & 0x10u) ? 18u : 17u);
            int32_t qn = 0, ind = 0;
            if (os == st.x + cig_ref_len(cw) - 1) ind = op_indel(a.cigar, a.opstart, c0 + k, c0 + nops, &qn);
            a.e_read[e] = (int32_t)i; a.e_ch[e] = (uint8_t)ch; a.e_ind[e] = ind; a.e_qn[e] = qn;
        }
        cnt += __popc(cm);
    }
    if (!FILL && lane == 0) a.col_cnt[c] = cnt;
}

// (d) grouping of equal strings and the two read sets (:287-300): one thread per column, in pileup order like the reference
__device__ inline bool impute_same_token(const ImputeArgs& a, int64_t e1, int64_t e2) {
    if (a.e_ch[e1] != a.e_ch[e2] || a.e_ind[e1] != a.e_ind[e2]) return false;
    const int32_t L = a.e_ind[e1];
    if (L <= 0) return true;                                           // '-L' + 'N' * L carries no bases
    const int64_t r1 = a.e_read[e1], r2 = a.e_read[e2];
    const int32_t q1 = a.e_qn[e1], q2 = a.e_qn[e2];
    const int32_t n1 = max(0, min(L, __ldg(a.l_seq + r1) - q1)), n2 = max(0, min(L, __ldg(a.l_seq + r2) - q2));
    if (n1 != n2) return false;
    const uint8_t* s1 = a.seq4 + __ldg(a.seq_off + r1);
    const uint8_t* s2 = a.seq4 + __ldg(a.seq_off + r2);
    for (int32_t t = 0; t < n1; t++)
        if (read_nibble(s1, q1 + t) != read_nibble(s2, q2 + t)) return false;
    return true;
}
__global__ void indel_impute_group_kernel(const ImputeArgs a) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.n_cols) return;
    const int64_t off = a.col_off[c];
    const int32_t n = (int32_t)(a.col_off[c + 1] - off);
    int32_t G = 0;
    for (int32_t k = 0; k < n; k++) {
        int32_t found = -1;
        for (int32_t g = 0; g < G && found < 0; g++)
            if (impute_same_token(a, off + k, off + a.g_rep[off + g])) found = g;
        if (found < 0) { found = G++; a.g_rep[off + found] = k; a.g_cnt[off + found] = 0; }
        a.e_gid[off + k] = found;
        a.g_cnt[off + found]++;
    }
    int32_t n0 = 0, n1 = 0;
    if (n > 0) {
        int32_t g0 = 0;                                                // stable descending sort: first maximum, then first maximum of the rest
        for (int32_t g = 1; g < G; g++) if (a.g_cnt[off + g] > a.g_cnt[off + g0]) g0 = g;
        const int32_t c0 = a.g_cnt[off + g0];
        if ((double)c0 <= 0.8 * (double)n) {                           // :293 (then G >= 2)
            int32_t g1 = -1;
            for (int32_t g = 0; g < G; g++) if (g != g0 && (g1 < 0 || a.g_cnt[off + g] > a.g_cnt[off + g1])) g1 = g;
            const int32_t c1 = a.g_cnt[off + g1];
            const bool second = c1 >= a.mincov;                        // :295 else: all the other reads of the column
            for (int32_t k = 0; k < n; k++) {
                const int32_t g = a.e_gid[off + k];
                a.e_label[off + k] = g == g0 ? 0 : ((!second || g == g1) ? 1 : -1);
            }
            n0 = c0; n1 = second ? c1 : n - c0;
        } else {                                                       // :297-298 the top group is halved
            const int32_t half = c0 / 2;
            int32_t seen = 0;
            for (int32_t k = 0; k < n; k++) {
                if (a.e_gid[off + k] != g0) { a.e_label[off + k] = -1; continue; }
                a.e_label[off + k] = seen < half ? 0 : 1;
                seen++;
            }
            n0 = half; n1 = c0 - half;
        }
    }
    a.col_ok[c] = (n0 >= a.mincov && n1 >= a.mincov) ? 1 : 0;
}
// (e) pending columns -> hit 3 (imputed small-window candidate, :300-303) or 0
__global__ void indel_impute_apply_kernel(const int64_t* __restrict__ gl, const int32_t* __restrict__ ok, int64_t n, uint8_t* __restrict__ hit,
                                          unsigned long long* __restrict__ n_hits) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    hit[gl[i]] = ok[i] ? 3 : 0;
    if (ok[i]) atomicAdd(n_hits, 1ull);
}

// pass A (count) / pass B (fill): warp per site
template <bool FILL>
__global__ void __launch_bounds__(128) indel_site_reads_kernel(const SiteArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= a.n_sites) return;
    const NcIndelVariant sv = a.sites[s];
    const int32_t v = sv.key, p = v - 1;
    const NcChunk ck = a.chunks[sv.chunk];
    // pass-2 pileup range (:306) and reference window (:325-328)
    const int32_t plo = max(0, ck.start - 10 - a.win), phi = min(ck.end, a.contig_len);
    int32_t m = min(a.contig_len, v + a.wa + 1) - v;
    bool okref = p >= plo && p < phi && m > 0 && (int64_t)p >= a.ref_start && (int64_t)(p + m) <= a.ref_start + a.ref_len;
    if (okref) {
        bool bad = false;
        for (int j = lane; j < m; j += 32) { const uint8_t ch = __ldg(a.ref + ((int64_t)p + j - a.ref_start)); bad |= ref_code_of(ch) > 3; }
        okref = !__any_sync(0xffffffffu, bad);
    }
    int64_t ilo = 0, ihi = 0;
    if (okref) read_window(a.pos, a.pmaxend, a.n_reads, p, ilo, ihi);
    int32_t cnt = 0;
    const int64_t off = FILL ? a.site_off[s] : 0;
    for (int64_t base = ilo; base < ihi; base += 32) {
        const int64_t i = base + lane;
        bool cover = false;
        if (i < ihi) {
            const int32_t rp = __ldg(a.pos + i), re = __ldg(a.end + i);
            cover = (__ldg(a.flag + i) & a.flag_filter) == 0 && rp <= p && p < re;
        }
        const uint32_t cm = __ballot_sync(0xffffffffu, cover);
        if (FILL && cover) {
            const int64_t e = off + cnt + __popc(cm & ((1u << lane) - 1u));
            a.e_read[e] = (int32_t)i;
            // query_position_or_next (appendix C.4): op containing p
            const int64_t c0 = a.cigar_off[i];
            const int32_t nops = (int32_t)(a.cigar_off[i + 1] - c0), os = p - __ldg(a.pos + i);
            int32_t lo = 0, hi = nops;
            while (lo < hi) { int32_t mid = (lo + hi) >> 1; if (__ldg(&a.opstart[c0 + mid].x) <= os) lo = mid + 1; else hi = mid; }
            const int32_t k = lo - 1;
            const int2 st = __ldg(a.opstart + c0 + k);
            a.e_qpn[e] = cig_is_match(__ldg(a.cigar + c0 + k)) ? st.y + (os - st.x) : st.y;
            const int32_t imp = a.site_imp ? __ldg(a.site_imp + s) : -1;
            int grp = -1;
            if (imp < 0) { const int h = __ldg(a.hp + i); grp = (h == 1 || h == 2) ? h - 1 : -1; }
            else {                                  // the source column lists its reads in ascending BAM index
                int64_t lo2 = __ldg(a.imp_off + imp), hi2 = __ldg(a.imp_off + imp + 1);
                const int64_t end2 = hi2;
                while (lo2 < hi2) { const int64_t mid = (lo2 + hi2) >> 1; if (__ldg(a.imp_read + mid) < (int32_t)i) lo2 = mid + 1; else hi2 = mid; }
                if (lo2 < end2 && __ldg(a.imp_read + lo2) == (int32_t)i) grp = __ldg(a.imp_label + lo2);
            }
            a.e_grp[e] = (int8_t)grp;
        }
        cnt += __popc(cm);
    }
    if (!FILL && lane == 0) { a.site_cnt[s] = okref ? cnt : 0; a.site_m[s] = (okref && cnt > 0) ? m : 0; }
}

// Star alignment step 1: global linear-gap NW of one read slice against the reference window, one warp per entry.
// match +2, mismatch -4, gap -3; direction DIAG if the diagonal attains the max, else UP (read base unaligned), else LEFT.
// Lane l owns reference columns [l*CW, (l+1)*CW) (CW = 6 for windows up to 192 columns, 9 up to 288); rows are processed as a
// skewed wavefront (lane l works on row t-l).  Everything the wavefront and the traceback touch sits in shared memory — the
// read slice, the direction words, the aligned codes and the insertion tables — and leaves with coalesced stores at the end:
// the first version read the slice back from global memory in every step and ran the traceback (lane 0, ~n + m dependent steps)
// on read-modify-writes to global memory, 357 ms for configs[2]'s 4.1 M slices.
constexpr int kAlignWarps = 4;
constexpr int kRowsMax = 264;
constexpr int kAlignAux = 4 * 272 + 2 * 2 * 272;        // per warp: slice, aligned codes (bytes); insertion length / first (u16)
__host__ __device__ constexpr int align_smem_per_warp(int rows, int word_bytes) { return rows * 32 * word_bytes + kAlignAux; }

// WordT: direction word of one lane and row, 2 bits per column of its strip (u16 for CW = 6, u32 for CW = 9): 12.6 KB of shared memory per
// warp for the 161-column ONT window, i.e. 16 resident warps per SM to hide the sequential traceback of each
template <int CW, typename WordT>
__global__ void __launch_bounds__(kAlignWarps * 32) indel_align_kernel(const SiteArgs a, int64_t n_entries, int rows) {
    extern __shared__ __align__(16) uint8_t s_align_all[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint8_t* s_base = s_align_all + (size_t)wib * align_smem_per_warp(rows, (int)sizeof(WordT));
    WordT* s_dir = reinterpret_cast<WordT*>(s_base);             // [rows][32] direction words
    uint8_t* s_slice = s_base + (size_t)rows * 32 * sizeof(WordT);   // [272]
    uint8_t* s_ac = s_slice + 272;                                // [272] (+ 544 spare)
    uint16_t* s_il = reinterpret_cast<uint16_t*>(s_slice + 4 * 272);
    uint16_t* s_if = s_il + 272;
    const uint32_t full = 0xffffffffu;
    for (int64_t e = (int64_t)blockIdx.x * kAlignWarps + wib; e < n_entries; e += (int64_t)gridDim.x * kAlignWarps) {
        int64_t s;
        { int64_t lo = 0, hi = a.n_sites; while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (__ldg(a.site_off + mid) <= e) lo = mid + 1; else hi = mid; } s = lo - 1; }
        const int32_t m = a.site_m[s];
        const int32_t v = a.sites[s].key, p = v - 1;
        const int64_t ri = a.e_read[e];
        const int32_t q0 = a.e_qpn[e], lseq = __ldg(a.l_seq + ri);
        const int32_t n = max(0, min(a.wa, lseq - q0));
        const uint8_t* sq = a.seq4 + __ldg(a.seq_off + ri);
        uint8_t* o_slice = a.e_slice + e * a.nmax;
        uint8_t* o_ac = a.e_acode + e * a.mmax;
        uint16_t* o_il = a.e_inslen + e * (a.mmax + 1);
        uint16_t* o_if = a.e_insfirst + e * (a.mmax + 1);
        // slice codes -> shared (the DP's read sequence) and global (the msa kernel reads inserted bases from it)
        for (int i = lane; i < n; i += 32) {
            const int32_t q = q0 + i;
            const uint32_t b = __ldg(sq + (q >> 1));
            const uint32_t bn = (q & 1) ? (b & 15u) : (b >> 4);
            const uint8_t c = (uint8_t)((kNibToCode >> (4 * bn)) & 15u);
            s_slice[i] = c; o_slice[i] = c;
        }
        for (int j = lane; j <= m; j += 32) { s_il[j] = 0; s_if[j] = 0; }
        __syncwarp();
        // reference codes of the lane's strip
        int refc[CW];
#pragma unroll
        for (int k = 0; k < CW; k++) {
            const int j = lane * CW + k;         // 0-based reference column
            refc[k] = j < m ? ref_code_of(__ldg(a.ref + ((int64_t)p + j - a.ref_start))) : 7;
        }
        // H of the previous row for the strip: hp[k] = H[i-1][j0+k+1], hleft = H[i-1][j0] (column left of the strip)
        int32_t hp[CW];
#pragma unroll
        for (int k = 0; k < CW; k++) hp[k] = -3 * (lane * CW + k + 1);
        int32_t h_left_prev = -3 * (lane * CW);               // H[i-1][j0]
        int32_t last_out = 0;                                   // H[i][last column of strip] of the row finished in the previous step
        for (int t = 1; t <= n + 31; t++) {
            // value of the left neighbour's strip end for the row this lane works on now: lane-1 finished row i at step t-1
            const int32_t from_left = __shfl_up_sync(full, last_out, 1);
            const int i = t - lane;
            if (i >= 1 && i <= n) {
                const int rb = s_slice[i - 1];
                const int32_t h_left_cur = lane == 0 ? -3 * i : from_left;        // H[i][j0]
                int32_t diag_in = h_left_prev, left = h_left_cur;
                uint32_t dw = 0;
#pragma unroll
                for (int k = 0; k < CW; k++) {
                    const int32_t sc = (refc[k] == rb && rb < 4) ? 2 : -4;
                    const int32_t dg = diag_in + sc, up = hp[k] - 3, lf = left - 3;
                    const int32_t best = max(dg, max(up, lf));
                    const uint32_t d = best == dg ? 0u : (best == up ? 1u : 2u);
                    dw |= d << (2 * k);
                    diag_in = hp[k];
                    hp[k] = best;
                    left = best;
                }
                s_dir[i * 32 + lane] = (WordT)dw;
                h_left_prev = h_left_cur;
                last_out = left;
            }
        }
        __syncwarp();
        // traceback (lane 0) from (n, m), in shared memory
        if (lane == 0) {
            int i = n, j = m;
            while (i > 0 || j > 0) {
                uint32_t d;
                if (i == 0) d = 2u; else if (j == 0) d = 1u;
                else d = ((uint32_t)s_dir[i * 32 + (j - 1) / CW] >> (2 * ((j - 1) % CW))) & 3u;
                if (d == 0u) { s_ac[j - 1] = s_slice[i - 1]; i--; j--; }
                else if (d == 1u) { s_il[j]++; s_if[j] = (uint16_t)(i - 1); i--; }
                else { s_ac[j - 1] = 5; j--; }
            }
            a.e_n[e] = n;
        }
        __syncwarp();
        for (int j = lane; j < m; j += 32) o_ac[j] = s_ac[j];
        for (int j = lane; j <= m; j += 32) { o_il[j] = s_il[j]; o_if[j] = s_if[j]; }
        __syncwarp();
    }
}

// The same alignment, two slices of one site at a time in the halves of 32-bit registers (windows up to 8 columns per lane strip).
//   * A cell value is V = 4 H + tie with tie = 2 (DIAG), 1 (UP), 0 (LEFT) in the two low bits: max(V_diag, V_up, V_left) then IS the
//     reference's choice (DIAG if the diagonal attains the maximum, else UP, else LEFT) and its low bits are the direction — no
//     compares.  |4 H| <= 12 (n + m) < 2^15, so a cell fits 16 bits and VIADD.16x2 / VIADDMNMX.S16x2 do two slices per instruction.
//   * Both slices see the same reference window, so the substitution scores of a strip column are one register of four bytes
//     (score against A, G, T, C) and the scores of the two read bases of a row are ONE byte permute with a per-row selector
//     (sign-replicating prmt; selector 4 = the all-mismatch constant serves N and the rows past a slice's end).
//   * One warp per site walks the site's slices pair by pair (the reference strip is set up once per site); lanes 0 and 1 trace the
//     two paths back at the same time.
// Per pair and strip column: prmt, add, 2 x add-max, and-mask, shift, or-and = 7 instructions for two cells (the scalar kernel: ~12
// per cell).
constexpr int kAlign2Sel = 352;                                  // row selectors: rows + lane strips of the widest window, 32 entries of padding in front
constexpr int kAlign2Aux = 2 * kAlign2Sel + 2 * (272 + 2 * 272);   // row selectors (u16); per half: slice (u8), traceback record per reference column (u16)
// direction words: rows -2 .. wa (row 0 and the two rows above it are all LEFT / never read as directions), `rs` words per row = lane
// strips in use (27 for the 161-column ONT window): 17.6 KB per pair, 11 warps per SM.  Strips of nine columns (windows up to 288
// columns, the HiFi preset's 261) keep the ninth column's two bits per half in a byte array beside the words: 38 KB per pair, 5 warps.
__host__ __device__ constexpr int align2_dir_bytes(int wa, int rs) { return ((wa + 3) * rs * 4 + 15) & ~15; }
__host__ __device__ constexpr int align2_dirb_bytes(int wa, int rs, bool wide) { return wide ? (((wa + 3) * rs + 15) & ~15) : 0; }
__host__ __device__ constexpr int align2_smem_per_warp(int wa, int rs, bool wide = false) {
    return align2_dir_bytes(wa, rs) + align2_dirb_bytes(wa, rs, wide) + kAlign2Aux;
}
__device__ __forceinline__ uint32_t pack2_s16(int32_t x) { return ((uint32_t)x & 0xFFFFu) | ((uint32_t)x << 16); }
__device__ __forceinline__ uint32_t prmt_sx(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }

template <int CW, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) indel_align2_kernel(const SiteArgs a, int rs) {
    static_assert(CW <= 9, "two direction bits per column and half: eight columns in the word, a ninth in the byte array");
    constexpr bool WIDE = CW > 8;
    extern __shared__ __align__(16) uint8_t s_align_all[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint8_t* s_base = s_align_all + (size_t)wib * align2_smem_per_warp(a.wa, rs, WIDE);
    uint32_t* s_dir = reinterpret_cast<uint32_t*>(s_base) + 2 * rs;              // [row -2 .. wa][rs]: half h = slice h, 2 bits per strip column, column 0 highest
    uint8_t* s_dirb = s_base + align2_dir_bytes(a.wa, rs) + 2 * rs;              // WIDE: column 8 of the strip, bits 1:0 slice 0, 3:2 slice 1
    uint16_t* s_sel = reinterpret_cast<uint16_t*>(s_base + align2_dir_bytes(a.wa, rs) + align2_dirb_bytes(a.wa, rs, WIDE));   // [kAlign2Sel]
    uint8_t* s_half = reinterpret_cast<uint8_t*>(s_sel + kAlign2Sel);            // per half: slice[272], rec[272 u16]
    constexpr int kHalfBytes = 272 + 2 * 272;
    for (int i = lane; i < 3 * rs; i += 32) { s_dir[i - 2 * rs] = 0u; if (WIDE) s_dirb[i - 2 * rs] = 0; }   // rows <= 0 = LEFT everywhere: a path that has used up its read walks left
    const uint32_t full = 0xffffffffu;
    const uint32_t kMis = 0xF2F2F2F2u;                                          // 4 * (-4) + 2 in every byte
    const uint32_t kUp = 0xFFF5FFF5u, kLeft = 0xFFF4FFF4u;                      // 4 * (-3) + 1, 4 * (-3) + 0
#ifdef NC_ALIGN_PROFILE
    long long prof[5] = {0, 0, 0, 0, 0}, pl = clock64(); int prof_n = 0;
#define NC_AMARK(i) { const long long now_ = clock64(); prof[i] += now_ - pl; pl = now_; }
#else
#define NC_AMARK(i)
#endif
    for (int64_t s = (int64_t)blockIdx.x * WARPS + wib; s < a.n_sites; s += (int64_t)gridDim.x * WARPS) {
        const int64_t e0 = __ldg(a.site_off + s), e1 = __ldg(a.site_off + s + 1);
        const int32_t m = a.site_m[s];
        if (e1 <= e0 || m <= 0) continue;
        const int32_t p = a.sites[s].key - 1;
        // scores of the strip's reference columns against read base 0..3: match 4 * 2 + 2, mismatch 4 * (-4) + 2
        uint32_t S[CW];
#pragma unroll
        for (int k = 0; k < CW; k++) {
            const int j = lane * CW + k;
            const int rc = j < m ? ref_code_of(__ldg(a.ref + ((int64_t)p + j - a.ref_start))) : 7;
            S[k] = rc < 4 ? (kMis ^ (0xF8u << (8 * rc))) : kMis;                // 0xF2 ^ 0xF8 = 0x0A
        }
        const int last_lane = (m - 1) / CW;                                     // < rs
        for (int64_t e = e0; e < e1; e += 2) {
            const bool two = e + 1 < e1;
            NC_AMARK(0);
            int32_t nh[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                uint8_t* sl = s_half + h * kHalfBytes;
                int32_t n = 0;
                if (h == 0 || two) {
                    const int64_t ri = a.e_read[e + h];
                    const int32_t q0 = a.e_qpn[e + h], lseq = __ldg(a.l_seq + ri);
                    n = max(0, min(a.wa, lseq - q0));
                    const uint8_t* sq = a.seq4 + __ldg(a.seq_off + ri);
                    uint8_t* o_slice = a.e_slice + (e + h) * a.nmax;
                    for (int i = lane; i < n; i += 32) {
                        const int32_t q = q0 + i;
                        const uint32_t b = __ldg(sq + (q >> 1));
                        const uint32_t bn = (q & 1) ? (b & 15u) : (b >> 4);
                        const uint8_t c = (uint8_t)((kNibToCode >> (4 * bn)) & 15u);
                        sl[i] = c; o_slice[i] = c;
                    }
                }
                nh[h] = n;
            }
            const int32_t nmax = max(nh[0], nh[1]);
            __syncwarp();
            for (int i = lane; i < nmax; i += 32) {
                const uint32_t ra = i < nh[0] ? s_half[i] : 4u, rb = i < nh[1] ? s_half[kHalfBytes + i] : 4u;
                s_sel[32 + i] = (uint16_t)(ra | ((ra | 8u) << 4) | (rb << 8) | ((rb | 8u) << 12));
            }
            __syncwarp();
            NC_AMARK(1);
            // wavefront: lane l works on row t - l
            uint32_t hp[CW];
#pragma unroll
            for (int k = 0; k < CW; k++) hp[k] = pack2_s16(-12 * (lane * CW + k + 1));
            uint32_t h_left_prev = pack2_s16(-12 * (lane * CW));
            uint32_t last_out = 0;
            const int t_end = nmax + last_lane;
            // The row selector is loaded one step ahead (s_sel is padded by 32 entries in front: lanes that have not started yet read
            // unused values).  (A variant with a one-instruction serial chain per column — the tie bits cleared off the chain — was
            // slower: the kernel is bound by instruction issue, not by the chain's latency.)
            uint32_t sel = s_sel[32 - lane];                                    // row 1 - lane ... (index i - 1 + 32)
            for (int t = 1; t <= t_end; t++) {
                const uint32_t from_left = __shfl_up_sync(full, last_out, 1);
                const int i = t - lane;
                const uint32_t sel_next = s_sel[32 + i];
                if (i >= 1 && i <= nmax && lane <= last_lane) {
                    const uint32_t h_left_cur = lane == 0 ? pack2_s16(-12 * i) : from_left;
                    uint32_t diag_in = h_left_prev, left = h_left_cur, dw = 0, dwb = 0;
#pragma unroll
                    for (int k = 0; k < CW; k++) {
                        const uint32_t dg = __vadd2(diag_in, prmt_sx(S[k], kMis, sel));
                        const uint32_t b = __viaddmax_s16x2(left, kLeft, __viaddmax_s16x2(hp[k], kUp, dg));
                        if (k < 8) dw = (dw << 2) | (b & 0x00030003u);
                        else dwb = b & 0x00030003u;
                        diag_in = hp[k];
                        hp[k] = left = b & 0xFFFCFFFCu;
                    }
                    s_dir[i * rs + lane] = dw;
                    if (WIDE) s_dirb[i * rs + lane] = (uint8_t)((dwb | (dwb >> 14)) & 0xFu);
                    h_left_prev = h_left_cur;
                    last_out = left;
                }
                sel = sel_next;
            }
            __syncwarp();
            NC_AMARK(2);
            // traceback from (n, m): lane h follows slice h.  The sequential loop only records, per reference column j, the row at
            // which the path leaves it and whether it leaves diagonally; codes and insertion tables follow from that in parallel.
            // It is branch-free (the two lanes would diverge on every branch), and the direction words of the two rows above the
            // current one are loaded ahead, so that a step's dependent chain is shift, mask, compare, select instead of a
            // shared-memory round trip; only a new strip (every CW-th column) reloads.
            if (WIDE) {
                // nine-column strips: a cell's bits come from the word (columns 0-7: bit 16 h + 2 (7 - column)) or from the byte (column
                // 8: bit 2 h), so the three rows in flight are 64-bit values {word, byte << 32} and the shift walks 32 + 2 h, 16 h, 16 h + 2, ...
                if (lane < 2 && (lane == 0 || two)) {
                    const int n_me = lane == 0 ? nh[0] : nh[1];
                    const uint32_t rec0 = smem_u32(s_half + lane * kHalfBytes + 272);       // rec[0]
                    const uint32_t a0 = smem_u32(s_dir), b0 = smem_u32(s_dirb);
                    auto loadw = [&](int32_t idx) { return (uint64_t)lds_u32(a0 + 4u * (uint32_t)idx) | ((uint64_t)lds_u8(b0 + (uint32_t)idx) << 32); };
                    const int jl0 = (m - 1) / CW, kk0 = (m - 1) - jl0 * CW;
                    int32_t idx = n_me * rs + jl0;
                    uint32_t rp = rec0 + 2u * (uint32_t)m;
                    const uint32_t sh_b = 32u + 2u * lane, sh_a0 = 16u * lane, sh_end = sh_a0 + 16u;
                    uint32_t sh = kk0 == 8 ? sh_b : sh_a0 + 2u * (uint32_t)(7 - kk0);
                    uint32_t i2 = (uint32_t)n_me << 1;
                    uint64_t w_cur = loadw(idx), w_up = loadw(idx - rs), w_up2 = loadw(idx - 2 * rs);
#pragma unroll 2
                    while (rp != rec0) {
                        const uint32_t d = (uint32_t)(w_cur >> sh) & 3u;                    // 2 DIAG, 1 UP, 0 LEFT
                        const bool cm = d != 1u, rm = d != 0u;
                        if (cm) asm volatile("st.shared.u16 [%0], %1;" :: "r"(rp), "h"((uint16_t)(i2 | (d >> 1))) : "memory");
                        rp -= cm ? 2u : 0u;
                        const uint32_t nsh = sh == sh_b ? sh_a0 : sh + 2u;
                        sh = cm ? nsh : sh;
                        i2 -= rm ? 2u : 0u;
                        idx -= rm ? rs : 0;
                        const bool wrap = sh == sh_end && rp != rec0;                       // past column 0 of the strip: the strip to the left, its column 8
                        idx -= wrap ? 1 : 0;
                        sh = wrap ? sh_b : sh;
                        const uint64_t nw = loadw(idx - 2 * rs);
                        w_cur = rm ? w_up : w_cur;
                        w_up = rm ? w_up2 : w_up;
                        w_up2 = nw;
                        if (wrap) { w_cur = loadw(idx); w_up = loadw(idx - rs); }
                    }
                    a.e_n[e + lane] = n_me;
                }
            } else if (lane < 2 && (lane == 0 || two)) {
                const int n_me = lane == 0 ? nh[0] : nh[1];
                const uint32_t rec0 = smem_u32(s_half + lane * kHalfBytes + 272);           // rec[0]
                const uint32_t rsb = (uint32_t)rs * 4u;
                const int jl0 = (m - 1) / CW;
                uint32_t ad = smem_u32(s_dir + n_me * rs + jl0);
                uint32_t rp = rec0 + 2u * (uint32_t)m;
                const uint32_t sh_lo = 16u * lane, sh_wrap = sh_lo + 2u * CW;
                uint32_t sh = sh_lo + 2u * (uint32_t)(CW - 1 - ((m - 1) - jl0 * CW));
                uint32_t i2 = (uint32_t)n_me << 1;                                          // 2 i
                uint32_t w_cur = lds_u32(ad), w_up = lds_u32(ad - rsb), w_up2 = lds_u32(ad - 2u * rsb);
#pragma unroll 2
                while (rp != rec0) {
                    const uint32_t d = (w_cur >> sh) & 3u;                                  // 2 DIAG, 1 UP, 0 LEFT
                    const bool cm = d != 1u, rm = d != 0u;
                    if (cm) asm volatile("st.shared.u16 [%0], %1;" :: "r"(rp), "h"((uint16_t)(i2 | (d >> 1))) : "memory");
                    rp -= cm ? 2u : 0u;
                    sh += cm ? 2u : 0u;
                    i2 -= rm ? 2u : 0u;
                    ad -= rm ? rsb : 0u;
                    const bool wrap = sh == sh_wrap && rp != rec0;                        // (not past column 1: strip -1 does not exist)
                    ad -= wrap ? 4u : 0u;
                    sh = wrap ? sh_lo : sh;
                    const uint32_t nw = lds_u32(ad - 2u * rsb);                             // in use two row moves from now
                    w_cur = rm ? w_up : w_cur;
                    w_up = rm ? w_up2 : w_up;
                    w_up2 = nw;
                    if (wrap) { w_cur = lds_u32(ad); w_up = lds_u32(ad - rsb); }
                }
                a.e_n[e + lane] = n_me;
            }
            __syncwarp();
            NC_AMARK(3);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                if (h == 1 && !two) break;
                const uint8_t* sl = s_half + h * kHalfBytes;
                const uint16_t* rec = reinterpret_cast<const uint16_t*>(sl + 272);
                uint8_t* o_ac = a.e_acode + (e + h) * a.mmax;
                uint16_t* o_il = a.e_inslen + (e + h) * (a.mmax + 1);
                uint16_t* o_if = a.e_insfirst + (e + h) * (a.mmax + 1);
                for (int j = lane; j <= m; j += 32) {
                    // row at which the path arrives at column j; rows above the row it leaves at are the bases inserted after column j
                    const int32_t arrive = j == m ? nh[h] : (int32_t)(rec[j + 1] >> 1) - (int32_t)(rec[j + 1] & 1u);
                    int32_t leave = 0;
                    if (j > 0) {
                        const uint32_t r = rec[j];
                        leave = (int32_t)(r >> 1);
                        o_ac[j - 1] = (r & 1u) ? sl[leave - 1] : (uint8_t)5;
                    }
                    const int32_t il = arrive - leave;
                    o_il[j] = (uint16_t)il;
                    o_if[j] = (uint16_t)(il > 0 ? leave : 0);
                }
            }
            __syncwarp();
            NC_AMARK(4);
#ifdef NC_ALIGN_PROFILE
            prof_n++;
#endif
        }
    }
#ifdef NC_ALIGN_PROFILE
    if (blockIdx.x == 7 && threadIdx.x == 0 && prof_n > 0)
        printf("align2 profile pairs=%d cycles/pair: other %lld setup %lld wave %lld trace %lld out %lld\n", prof_n, prof[0] / prof_n, prof[1] / prof_n, prof[2] / prof_n, prof[3] / prof_n, prof[4] / prof_n);
#endif
#undef NC_AMARK
}

// Star alignment step 2 + `msa` tensor assembly (:54-71): one warp per (site, group); group 0 = HP1, 1 = HP2, 2 = all reads.
__global__ void __launch_bounds__(96) indel_msa_kernel(const SiteArgs a) {
    __shared__ uint16_t s_width[3][kRowsMax + 8];
    __shared__ uint16_t s_col[3][kRowsMax + 8];
    __shared__ uint8_t s_cns[3][2 * kRowsMax + 32];
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int64_t s = blockIdx.x;
    const uint32_t full = 0xffffffffu;
    const int32_t m = a.site_m[s];
    float* T = a.tensors + ((s * 3 + g) * 5) * 256;
    for (int i = lane; i < 5 * 256; i += 32) T[i] = 0.f;
    __syncwarp();
    int32_t n_g = 0, first_read = -1;
    const int64_t e0 = a.site_off[s];
    const int32_t cnt = m > 0 ? a.site_cnt[s] : 0;
    const int32_t mincov = g == 2 ? a.mincov : 2;
    // membership: first maxcov reads of the group in pileup order (deterministic rule for the unseeded random.sample, :19)
    auto member = [&](int32_t k) -> bool { return g == 2 || a.e_grp[e0 + k] == g; };
    // n_g and the first member
    for (int32_t k0 = 0; k0 < cnt; k0 += 32) {
        const int32_t k = k0 + lane;
        const bool mem = k < cnt && member(k);
        const uint32_t bm = __ballot_sync(full, mem);
        if (first_read < 0 && bm) first_read = a.e_read[e0 + k0 + __ffs(bm) - 1];
        n_g += __popc(bm);
    }
    const int32_t n_use = min(n_g, a.maxcov);
    const bool ok = m > 0 && n_use >= mincov;
    uint16_t* width = s_width[g];
    uint16_t* col = s_col[g];
    int32_t L = 0, cl = 0;
    if (ok) {
        // widths of the insertion blocks
        for (int j = lane; j <= m; j += 32) width[j] = 0;
        __syncwarp();
        int32_t seen = 0;
        for (int32_t k = 0; k < cnt && seen < n_use; k++) {
            if (!member(k)) continue;                      // warp-uniform
            seen++;
            const uint16_t* il = a.e_inslen + (e0 + k) * (a.mmax + 1);
            for (int j = lane; j <= m; j += 32) { const uint16_t w = il[j]; if (w > width[j]) width[j] = w; }
        }
        __syncwarp();
        // column of reference base j = j + sum_{j' <= j} width[j']
        int32_t carry = 0;
        for (int j0 = 0; j0 <= m; j0 += 32) {
            const int j = j0 + lane;
            const int32_t w = j <= m ? width[j] : 0;
            const int32_t inc = warp_incl_scan32(w, lane);
            if (j <= m) col[j] = (uint16_t)(j + carry + inc);
            carry += __shfl_sync(full, inc, 31);
        }
        __syncwarp();
        L = m + carry;
        const float nf = (float)n_use;
        uint8_t* cns_sym = s_cns[g];
        for (int i = lane; i < L; i += 32) cns_sym[i] = 4;
        __syncwarp();
        // match columns: lane-parallel over reference bases
        for (int j = lane; j < m; j += 32) {
            int32_t c[5] = {0, 0, 0, 0, 0};
            int32_t seen2 = 0;
            for (int32_t k = 0; k < cnt && seen2 < n_use; k++) {
                if (!member(k)) continue;
                seen2++;
                const uint8_t code = a.e_acode[(e0 + k) * a.mmax + j];
                c[code < 4 ? code : 4]++;
            }
            const int rc = ref_code_of(__ldg(a.ref + ((int64_t)(a.sites[s].key - 1) + j - a.ref_start)));
            const int cc = col[j];
            float best = -1.f; int bi = 0;
#pragma unroll
            for (int b = 0; b < 5; b++) {
                const float f = (float)c[b] / nf;
                const float tf = b == 4 ? f - 0.01f : f;
                if (tf > best) { best = tf; bi = b; }
                if (cc < 128) { T[(b * 128 + cc) * 2] = f - (b == rc ? 1.f : 0.f); T[(b * 128 + cc) * 2 + 1] = b == rc ? 1.f : 0.f; }
            }
            cns_sym[cc] = (uint8_t)bi;
        }
        // insertion columns: slot j (before reference base j, or after the last one for j = m), kk-th inserted base.  Lanes take
        // insertion COLUMNS (all slots' columns numbered consecutively: column q of the alignment is an insertion column iff it
        // is not col[j] of a reference base), not the 1-3 columns of one slot at a time — that loop ran at 6 of 32 lanes.
        {
            const int32_t n_ins = L - m;                            // = carry
            for (int32_t q0 = 0; q0 < n_ins; q0 += 32) {
                const int32_t q = q0 + lane;
                if (q < n_ins) {
                    // slot j = number of reference bases whose insertion blocks end at or before the q-th insertion column:
                    // the insertion columns of slots 0..j total col[j] - j (col[j] = j + sum of widths up to j); find the first j with col[j] - j > q
                    int lo = 0, hi = m;                             // slot m (after the last base) when no j < m qualifies
                    while (lo < hi) { const int mid = (lo + hi) >> 1; if ((int32_t)col[mid] - mid > q) hi = mid; else lo = mid + 1; }
                    const int j = lo;
                    const int w = width[j];
                    const int before = (j < m ? (int32_t)col[j] - j : n_ins) - w;      // insertion columns of the slots before j
                    const int kk = q - before;
                    const int cc = (j < m ? (int32_t)col[j] : L) - w + kk;
                    int32_t c[5] = {0, 0, 0, 0, 0};
                    int32_t seen2 = 0;
                    for (int32_t k = 0; k < cnt && seen2 < n_use; k++) {
                        if (!member(k)) continue;
                        seen2++;
                        const int64_t e = e0 + k;
                        const uint16_t il = a.e_inslen[e * (a.mmax + 1) + j];
                        int code = 4;
                        if (kk < il) { code = a.e_slice[e * a.nmax + a.e_insfirst[e * (a.mmax + 1) + j] + kk]; if (code > 3) code = 4; }
                        c[code]++;
                    }
                    float best = -1.f; int bi = 0;
#pragma unroll
                    for (int b = 0; b < 5; b++) {
                        const float f = (float)c[b] / nf;
                        const float tf = b == 4 ? f - 0.01f : f;
                        if (tf > best) { best = tf; bi = b; }
                        if (cc < 128) { T[(b * 128 + cc) * 2] = f - (b == 4 ? 1.f : 0.f); T[(b * 128 + cc) * 2 + 1] = b == 4 ? 1.f : 0.f; }
                    }
                    cns_sym[cc] = (uint8_t)bi;
                }
            }
        }
        __syncwarp();
        // consensus = column argmax without gaps (:63-64), ordered compaction
        uint8_t* out = a.cns + (s * 3 + g) * a.cmax;
        for (int i0 = 0; i0 < L; i0 += 32) {
            const int i = i0 + lane;
            const int sy = i < L ? cns_sym[i] : 4;
            const uint32_t bm = __ballot_sync(full, sy < 4);
            if (sy < 4) { const int o = cl + __popc(bm & ((1u << lane) - 1u)); if (o < a.cmax) out[o] = (uint8_t)sy; }
            cl += __popc(bm);
        }
    }
    if (lane == 0) {
        NcIndelSiteMeta* mt = a.meta + s;
        mt->n[g] = ok ? n_use : 0;
        mt->cns_len[g] = ok ? min(cl, a.cmax) : 0;
        mt->ok[g] = ok ? 1 : 0;
        if (g == 0) { mt->pos = a.sites[s].key; mt->chunk = a.sites[s].chunk; mt->type = a.sites[s].type; mt->ref_len = m;
                      const bool imputed = a.site_imp && a.site_imp[s] >= 0;      // phase_dict gives None for a read without HP: -1
                      mt->phase = (ok && first_read >= 0) ? ((imputed && __ldg(a.hp + first_read) <= 0) ? -1 : __ldg(a.ps + first_read)) : 0; }
    }
}


// ------------------------------------------------------------------------------------------------
// I4 — allele_prediction (generate_indel_pileups.py:77-127) on the device: global affine alignment of the consensus against the
// reference window with traceback (the reference calls parasail.nw_trace(alt, ref, 9, 1, matrix 20 / -10), :79; this is the same
// recurrence and the same tie rules as nc_nw_trace on the host: H prefers DIAG, then F ('I'), then E ('D'); E / F prefer extension),
// then the reference's walk over the CIGAR.  One warp per (site, read group); lane l owns reference columns [l*CW, (l+1)*CW), rows
// run as a skewed wavefront; four flag bits per cell go to a per-warp scratch in global memory (L2), the traceback and the walk are
// sequential on lane 0.  out[(site * 3 + group) * 2] = {ref_out_len, alt_out_len}, -1 / -1 where the reference returns (None, None).
// ------------------------------------------------------------------------------------------------
struct AlleleArgs {
    int64_t n_sites; const NcIndelVariant* sites; const NcIndelSiteMeta* meta; const uint8_t* cns; int32_t cmax;
    const uint8_t* ref; int64_t ref_start; int32_t win; int32_t haploid;
    int32_t go, ge, match, mismatch;
    void* scratch; int32_t rows_cap;           // per warp [rows_cap][32] flag words
    uint8_t* ops_scratch;                      // per warp [cmax + 272] reversed alignment ops
    int32_t* out;
};

template <int CW, typename WordT>
__global__ void __launch_bounds__(128) indel_allele_kernel(const AlleleArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    WordT* dirs = reinterpret_cast<WordT*>(a.scratch) + warp_global * (int64_t)a.rows_cap * 32;
    uint8_t* ops = a.ops_scratch + warp_global * (int64_t)(a.cmax + 272);
    const uint32_t full = 0xffffffffu;
    const int32_t NEG = -(1 << 28);                                       // a multiple of 16 (see the flag bits below)
    for (int64_t item = warp_global; item < a.n_sites * 3; item += n_warps) {
        const int64_t s = item / 3;
        const int g = (int)(item - s * 3);
        const NcIndelSiteMeta mt = a.meta[s];
        const bool kept = a.haploid ? (g == 2 && mt.ok[2] > 0) : (mt.ok[0] > 0 && mt.ok[1] > 0 && mt.ok[2] > 0);
        if (!kept) { if (lane == 0) { a.out[item * 2] = -1; a.out[item * 2 + 1] = -1; } continue; }
        const int32_t n = mt.cns_len[g], m = mt.ref_len;
        const uint8_t* q = a.cns + item * a.cmax;
        const int64_t p = (int64_t)mt.pos - 1 - a.ref_start;
        int32_t ref_out = -1, alt_out = -1;
        if (n + 1 <= a.rows_cap && m <= 32 * CW) {
            int refc[CW];
#pragma unroll
            for (int k = 0; k < CW; k++) { const int j = lane * CW + k; refc[k] = j < m ? ref_code_of(__ldg(a.ref + p + j)) : 7; }
            // previous row of the strip: H[i-1][j0+k+1], F[i-1][j0+k+1]; row 0: H = E = -go - ge * (j - 1), F = NEG.
            // Values are kept as 16 * score: the four low bits of a candidate carry its rank in the reference's tie rules, so that the
            // maximum IS the choice and the low bits of the three maxima are the cell's flags — H: diagonal 2 > F 1 > E 0 in bits 1:0
            // (the order in which the traceback asks), F: extension 4 > opening 0 in bit 2, E: extension 8 > opening 0 in bit 3 —
            // instead of four equality tests per cell (the kernel is bound by instruction issue: 188 -> ~125 per wavefront step).
            const int32_t go16 = 16 * a.go, ge16 = 16 * a.ge, m16 = 16 * a.match + 2, x16 = 16 * a.mismatch + 2;
            int32_t hp[CW], fp[CW];                                                      // clean H; F + 1 (its rank inside H's maximum)
#pragma unroll
            for (int k = 0; k < CW; k++) { hp[k] = -go16 - ge16 * (lane * CW + k); fp[k] = NEG + 1; }
            int32_t h_left_prev = lane == 0 ? 0 : -go16 - ge16 * (lane * CW - 1);        // H[i-1][j0]
            int32_t out_h = 0, out_e = 0;                                                // H / E at the strip's last column of the row just finished
            for (int t = 1; t <= n + 31; t++) {
                const int32_t in_h = __shfl_up_sync(full, out_h, 1), in_e = __shfl_up_sync(full, out_e, 1);
                const int i = t - lane;
                if (i >= 1 && i <= n) {
                    const int qb = __ldg(q + i - 1);
                    const int32_t h_left_cur = lane == 0 ? -go16 - ge16 * (i - 1) : in_h;    // H[i][j0]  (column 0: H = F = -go - ge (i - 1), E = NEG)
                    int32_t e_left = lane == 0 ? NEG : in_e;                                 // E[i][j0]
                    int32_t diag_in = h_left_prev, left = h_left_cur;
                    WordT dw = 0;
#pragma unroll
                    for (int k = 0; k < CW; k++) {
                        const int32_t fv = max(fp[k] - ge16 + 3, hp[k] - go16);              // fp holds F + 1: extension = F - ge + 4
                        const int32_t ev = max(e_left - ge16 + 8, left - go16);
                        const int32_t dg = diag_in + ((refc[k] == qb) ? m16 : x16);
                        const int32_t f1 = (fv & ~15) | 1, e = ev & ~15;
                        const int32_t best = max(dg, max(f1, e));                            // bits 1:0: 2 diagonal, 1 F, 0 E
                        const uint32_t fl = ((uint32_t)best & 3u) | ((uint32_t)fv & 4u) | ((uint32_t)ev & 8u);
                        dw |= (WordT)fl << (4 * k);
                        diag_in = hp[k];
                        hp[k] = left = best & ~15; fp[k] = f1;
                        e_left = e;
                    }
                    dirs[(int64_t)i * 32 + lane] = dw;
                    h_left_prev = h_left_cur;
                    out_h = left; out_e = e_left;
                }
            }
            __syncwarp();
            if (lane == 0) {
                // traceback (nc_nw_trace): ops in reverse order
                int i = n, j = m, state = 0, nops = 0;
                while (i > 0 || j > 0) {
                    uint32_t fl = 0;
                    if (i > 0 && j > 0) fl = (uint32_t)(dirs[(int64_t)i * 32 + (j - 1) / CW] >> (4 * ((j - 1) % CW))) & 15u;
                    if (state == 0) {
                        if (i > 0 && j > 0 && (fl & 3u) == 2u) { ops[nops++] = (__ldg(q + i - 1) == ref_code_of(__ldg(a.ref + p + j - 1))) ? 7 : 8; i--; j--; }
                        else if (i > 0 && (j == 0 || (fl & 3u) == 1u)) state = 1;
                        else state = 2;
                    } else if (state == 1) {
                        ops[nops++] = 1;
                        const bool ext = i > 1 && (j == 0 || (fl & 4u));
                        i--;
                        if (!ext) state = 0;
                    } else {
                        ops[nops++] = 2;
                        const bool ext = j > 1 && (i == 0 || (fl & 8u));
                        j--;
                        if (!ext) state = 0;
                    }
                }
                // the reference's walk over the run-length CIGAR, control flow kept line for line (allele_predict_one on the host)
                const int32_t max_range = mt.type == 0 ? max(10, a.win) : 10;
                bool indel = false, mis_before = false, done = false;
                int32_t rc7 = 0, rc8 = 0, rc2 = 0, ac7 = 0, ac8 = 0, ac1 = 0, ma0 = 0, ma1 = 0;
                int op = 0; int32_t cnt = 0;
                int k = nops - 1;
                while (k >= 0 && !done) {
                    op = ops[k]; cnt = 0;
                    while (k >= 0 && ops[k] == op) { cnt++; k--; }
                    if (op == 8 || op == 7) {
                        if (op == 7) { rc7 += cnt; ac7 += cnt; } else { rc8 += cnt; ac8 += cnt; }
                        if (indel) { if (op == 7) ma0 += cnt; else ma1 += cnt; } else mis_before = true;
                    }
                    if (op == 1) { ac1 += cnt; ma0 = ma1 = 0; indel = true; }
                    if (op == 2) { rc2 += cnt; ma0 = ma1 = 0; indel = true; }
                    const int32_t rsum = rc7 + rc8 + rc2;
                    if (!indel && rsum >= max_range + 10) {
                        if (rc8) { const int32_t ol = op == 8 ? rsum : rsum - cnt; ref_out = ol; alt_out = ol; }
                        done = true;
                        break;
                    }
                    if (indel && ma0 + ma1 > 20) break;
                }
                if (!done && nops > 0) {
                    const int32_t rsum = rc7 + rc8 + rc2, asum = ac7 + ac8 + ac1;
                    int32_t r = op == 8 ? rsum : rsum - cnt, al = op == 8 ? asum : asum - cnt;
                    if (!mis_before) { r += 1; al += 1; }
                    ref_out = r; alt_out = al;
                }
            }
        } else if (lane == 0) {
            ref_out = -2; alt_out = -2;                       // does not fit the scratch: the host aligns this one (nc_allele_predict_batch)
        }
        if (lane == 0) { a.out[item * 2] = ref_out; a.out[item * 2 + 1] = alt_out; }
        __syncwarp();
    }
}

}  // namespace nc
