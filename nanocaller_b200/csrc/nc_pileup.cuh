// Pileup kernels of libnanocaller_b200 (sm_100a): the htslib-pileup replacement (K0), the column
// scan (K1), the neighbour matrix (K1b) and the per-candidate tensor build (K2).
//
// Reference being replaced: nanocaller_src/generate_SNP_pileups.py
//   K0  decode        samfile.pileup column semantics (:156-162)
//   K1  scan          per-column counts, alt_freq, neighbour / candidate tests (:158-186)
//   K1b neighbour mat pileup_dict[nb_pos][name] restricted to neighbour sites (:175,:179)
//   K2  tensor        get_cnd_pos (:6-101) + tensor assembly (:200-263)
//
// HBM layout (all per staged contig):
//   rows     u32 words, one "aligned row" per read: 8 reference positions per word, nibble t of a
//            word = position 8*w+t (absolute position multiples of 8 start a word, so tiles of any
//            read line up).  Nibble codes: 0..3 = A,G,T,C (generate_SNP_pileups.py:104), 4 = '*'/'N'
//            (deletion, ref-skip, N or any IUPAC code), 0xF = read does not cover the position.
//   flags    u8 per scanned position: bit0 neighbour site, bit1 candidate-eligible.
//   nrows    per read, the codes of the read at the neighbour sites inside its span, two per byte.
#pragma once
#include "nc_common.cuh"

namespace nc {

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int32_t lower_bound_i32(const int32_t* __restrict__ a, int32_t n, int32_t key) {
    int32_t lo = 0, hi = n;                      // first index with a[i] >= key
    while (lo < hi) { int32_t mid = (lo + hi) >> 1; if (__ldg(a + mid) < key) lo = mid + 1; else hi = mid; }
    return lo;
}
__device__ __forceinline__ int64_t lower_bound_i32_64(const int32_t* __restrict__ a, int64_t n, int32_t key) {
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (__ldg(a + mid) < key) lo = mid + 1; else hi = mid; }
    return lo;
}
// first index with a[i] > key
__device__ __forceinline__ int64_t upper_bound_i32_64(const int32_t* __restrict__ a, int64_t n, int32_t key) {
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (__ldg(a + mid) <= key) lo = mid + 1; else hi = mid; }
    return lo;
}
__device__ __forceinline__ int ref_code_of(uint8_t c) {    // only UPPER-case AGTC count (:137)
    return c == 'A' ? 0 : c == 'G' ? 1 : c == 'T' ? 2 : c == 'C' ? 3 : 4;
}
__device__ __forceinline__ int64_t warp_incl_scan64(int64_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int64_t t = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += t; }
    return v;
}
__device__ __forceinline__ int32_t warp_incl_scan32(int32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int32_t t = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += t; }
    return v;
}
// Exclusive block scan; `sm` must hold 33 int64.  Returns the exclusive prefix of v, sets total.
__device__ __forceinline__ int64_t block_excl_scan64(int64_t v, int64_t* sm, int64_t& total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int64_t inc = warp_incl_scan64(v, lane);
    if (lane == 31) sm[w] = inc;
    __syncthreads();
    if (w == 0) {
        int64_t x = lane < nw ? sm[lane] : 0;
        int64_t xi = warp_incl_scan64(x, lane);
        sm[lane] = xi - x;
        if (lane == 31) sm[32] = xi;
    }
    __syncthreads();
    total = sm[32];
    int64_t r = inc - v + sm[w];
    __syncthreads();
    return r;
}

// Warp-cooperative search in a sorted global array: 32 probes per round, so ~log32(n) dependent loads instead of log2(n).
// Returns the first index with a[i] >= key (upper = false) or a[i] > key (upper = true); identical on all lanes.
__device__ __forceinline__ int64_t warp_bound_i32(const int32_t* __restrict__ a, int64_t n, int32_t key, bool upper, int lane) {
    const uint32_t full = 0xffffffffu;
    int64_t lo = 0, hi = n;                                            // the answer lies in [lo, hi]
    while (hi - lo > 32) {
        const int64_t step = (hi - lo + 31) >> 5;
        const int64_t idx = lo + (int64_t)lane * step;
        bool before = false;
        if (idx < hi) { const int32_t x = __ldg(a + idx); before = upper ? x <= key : x < key; }
        const int k = __popc(__ballot_sync(full, before));             // probes lo, lo+step, ... that lie before the answer
        if (k < 32) hi = min(hi, lo + (int64_t)k * step);
        if (k > 0) lo = lo + (int64_t)(k - 1) * step + 1;
    }
    const int64_t idx = lo + lane;
    bool before = false;
    if (idx < hi) { const int32_t x = __ldg(a + idx); before = upper ? x <= key : x < key; }
    return lo + __popc(__ballot_sync(full, before));
}

// ------------------------------------------------------------------------------------------------
// exclusive scan int32 -> int64 (out has n+1 entries, out[n] = total); three launches
// ------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const int32_t* __restrict__ in, int64_t n,
                                                                    int64_t* __restrict__ partial) {
    __shared__ int64_t sm[33];
    int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int64_t s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) if (base + i < n) s += in[base + i];
    int64_t tot;
    block_excl_scan64(s, sm, tot);
    if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}
// single block: in-place exclusive scan of the partials, total -> partial[nparts]
__global__ void __launch_bounds__(1024) scan_partials_kernel(int64_t* __restrict__ partial, int64_t nparts) {
    __shared__ int64_t sm[33];
    int64_t carry = 0;
    for (int64_t b = 0; b < nparts; b += blockDim.x) {
        int64_t i = b + threadIdx.x;
        int64_t v = i < nparts ? partial[i] : 0, tot;
        int64_t ex = block_excl_scan64(v, sm, tot);
        if (i < nparts) partial[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) partial[nparts] = carry;
}
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const int32_t* __restrict__ in, int64_t n,
                                                                   const int64_t* __restrict__ partial,
                                                                   int64_t nparts, int64_t* __restrict__ out) {
    __shared__ int64_t sm[33];
    int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int32_t v[kScanItems];
    int64_t s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) { v[i] = base + i < n ? in[base + i] : 0; s += v[i]; }
    int64_t tot;
    int64_t ex = block_excl_scan64(s, sm, tot) + partial[blockIdx.x];
#pragma unroll
    for (int i = 0; i < kScanItems; i++) { if (base + i < n) out[base + i] = ex; ex += v[i]; }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = partial[nparts];
}

// ------------------------------------------------------------------------------------------------
// K0a — CIGAR prefix scan: one warp per read.  Per op the (reference, query) offsets at which it
// starts; per read the exclusive reference end and the number of aligned-row words.
// CIGAR op codes (SAM spec): M0 I1 D2 N3 S4 H5 P6 =7 X8.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int32_t cig_ref_len(uint32_t w) { return ((0x18Du >> (w & 15)) & 1u) ? (int32_t)(w >> 4) : 0; }
__device__ __forceinline__ int32_t cig_qry_len(uint32_t w) { return ((0x193u >> (w & 15)) & 1u) ? (int32_t)(w >> 4) : 0; }
__device__ __forceinline__ bool cig_is_match(uint32_t w) { return ((0x181u >> (w & 15)) & 1u) != 0; }

__global__ void __launch_bounds__(256) cigar_scan_kernel(int64_t n_reads, const int32_t* __restrict__ pos,
                                                         const int64_t* __restrict__ cigar_off,
                                                         const uint32_t* __restrict__ cigar,
                                                         int32_t* __restrict__ end, int32_t* __restrict__ nwords,
                                                         int2* __restrict__ opstart) {
    const int lane = threadIdx.x & 31;
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_reads;
         r += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const int64_t c0 = cigar_off[r], c1 = cigar_off[r + 1];
        int32_t rr = 0, qq = 0;
        for (int64_t k = c0; k < c1; k += 32) {
            const int64_t kk = k + lane;
            const uint32_t w = kk < c1 ? __ldg(cigar + kk) : 0u;
            const int32_t rl = cig_ref_len(w), ql = cig_qry_len(w);
            const int32_t ri = warp_incl_scan32(rl, lane), qi = warp_incl_scan32(ql, lane);
            if (kk < c1) opstart[kk] = make_int2(rr + ri - rl, qq + qi - ql);
            rr += __shfl_sync(0xffffffffu, ri, 31);
            qq += __shfl_sync(0xffffffffu, qi, 31);
        }
        if (lane == 0) {
            const int32_t p = pos[r];
            if (p < 0) rr = 0;                       // unplaced record: covers nothing
            end[r] = p + rr;
            nwords[r] = rr > 0 ? ((p + rr + 7) >> 3) - (p >> 3) : 0;
        }
    }
}

// Inclusive prefix maximum of end[] (single block).  pmaxend[i] > p  <=>  some read j <= i ends after p,
// which bounds from below the BAM-index window of reads that can cover p.
__global__ void __launch_bounds__(1024) prefix_max_kernel(const int32_t* __restrict__ end, int64_t n,
                                                          int32_t* __restrict__ pmax) {
    __shared__ int32_t sm[33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int32_t carry = INT32_MIN;
    for (int64_t b = 0; b < n; b += blockDim.x) {
        const int64_t i = b + threadIdx.x;
        int32_t v = i < n ? end[i] : INT32_MIN;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int32_t t = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v = max(v, t); }
        if (lane == 31) sm[w] = v;
        __syncthreads();
        if (w == 0) {
            int32_t x = sm[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { int32_t t = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x = max(x, t); }
            sm[lane] = x;
        }
        __syncthreads();
        int32_t pre = w > 0 ? sm[w - 1] : INT32_MIN;
        v = max(max(v, pre), carry);
        if (i < n) pmax[i] = v;
        carry = max(carry, sm[31]);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// K0c — query codes: the packed BAM sequence (4 bits per base, first base of a byte in the HIGH
// nibble, codes '='0 A1 C2 M3 G4 R5 S6 V7 T8 W9 Y10 H11 K12 D13 B14 N15) -> tensor codes
// (A0 G1 T2 C3, all else 4) with base i in nibble i & 7 of word i >> 3.  One streaming pass at HBM
// speed; afterwards a run of aligned bases is a funnel shift of two words.
// ------------------------------------------------------------------------------------------------
constexpr uint64_t kNibToCode = 0x4444444244414304ull;   // the same map as a 16-entry nibble table (scalar users)

// 8 BAM base codes (one per nibble) -> 8 tensor codes
__device__ __forceinline__ uint32_t bam_codes_to_tensor_codes(uint32_t n) {
    const uint32_t M1 = 0x11111111u;
    const uint32_t b0 = n & M1, b1 = (n >> 1) & M1, b2 = (n >> 2) & M1, b3 = (n >> 3) & M1;
    const uint32_t isA = b0 & ~(b1 | b2 | b3), isC = b1 & ~(b0 | b2 | b3), isG = b2 & ~(b0 | b1 | b3), isT = b3 & ~(b0 | b1 | b2);
    return (isC | isG) | ((isC | isT) << 1) | (((isA | isC | isG | isT) ^ M1) << 2);
}
__device__ __forceinline__ uint32_t swap_nibbles(uint32_t w) { return ((w & 0x0F0F0F0Fu) << 4) | ((w >> 4) & 0x0F0F0F0Fu); }

__global__ void __launch_bounds__(256) seq_codes_kernel(const uint4* __restrict__ seq4, int64_t n_vec, uint4* __restrict__ codes) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (int64_t)gridDim.x * blockDim.x) {
        uint4 v = __ldg(seq4 + i);
        v.x = bam_codes_to_tensor_codes(swap_nibbles(v.x)); v.y = bam_codes_to_tensor_codes(swap_nibbles(v.y));
        v.z = bam_codes_to_tensor_codes(swap_nibbles(v.z)); v.w = bam_codes_to_tensor_codes(swap_nibbles(v.w));
        codes[i] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// K0b — aligned-row fill: one CTA per read (reads handed out dynamically, long reads dominate), a
// warp per block of 64 output words (512 reference positions), a lane per 16 positions.
//   1. one linear pass over the read's ops records, per block, the op containing the block's first
//      position (shared memory);
//   2. the warp stages the block's reference-consuming ops in order (one 32-bit entry each: the
//      read-relative query index that reference offset 0 would have under the op, and whether it is
//      a match) and sets one bit per op START inside the block in a 512-bit mask: the ordinal of
//      the op a position belongs to is then a population count of the mask up to that position;
//   3. a lane builds its own two words: the 16 mask bits of its positions (plus a cut between the
//      two words) split them into runs of one op each (3 on average for ONT); a match run is a
//      funnel shift of two words of query codes, a deletion is a constant.  No atomics and no
//      per-segment tables; all index arithmetic is 32-bit and relative to the read.
// History in DESIGN 4.2 (lanes as words, lanes as (op, word) segments, this).
// ------------------------------------------------------------------------------------------------
constexpr int kFillThreads = 128;
constexpr int kFillWarps = kFillThreads / 32;
constexpr int kFillIdx = 1024;           // blocks indexed in shared memory (reads up to 524 kb; longer reads: per-word global search)
constexpr int kFillOps = 192;            // ops staged per block (more: per-word global search for that block)

__device__ __forceinline__ uint32_t nib_mask(int n) { return n >= 8 ? 0xFFFFFFFFu : ((1u << (4 * n)) - 1u); }   // n low nibbles
__device__ __forceinline__ uint32_t nib_mask_1to8(int n) { return 0xFFFFFFFFu >> (32 - 4 * n); }                 // 1 <= n <= 8
// tensor codes of the 8 query bases starting at absolute base index Q
__device__ __forceinline__ uint32_t fetch_codes8(const uint32_t* __restrict__ codes, int64_t Q) {
    return __funnelshift_r(__ldg(codes + (Q >> 3)), __ldg(codes + (Q >> 3) + 1), 4 * (int)(Q & 7));
}

// Fallback, one thread per word: read-relative offset o of nibble 0 (in [-7, span)), ops [0, n) in global memory.
__device__ __forceinline__ uint32_t fill_word_global(const int2* __restrict__ st, const uint32_t* __restrict__ cig, int n, int32_t o,
                                                     int32_t span, int32_t lseq, const uint32_t* __restrict__ codes, int64_t q_base) {
    uint32_t word = 0xFFFFFFFFu;                                    // 0xF = not covered
    int t = o < 0 ? -o : 0;
    int32_t oo = o + t;
    if (oo >= span) return word;
    int lo = 0, hi = n;                                             // last op whose reference start is <= oo
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(&st[mid].x) <= oo) lo = mid + 1; else hi = mid; }
    int j = lo - 1;
    uint32_t cw = __ldg(cig + j);
    int32_t x = __ldg(&st[j].x), rl = cig_ref_len(cw);
    const int32_t stop = min(o + 8, span);
    while (oo < stop) {
        while (oo >= x + rl) { j++; cw = __ldg(cig + j); x = __ldg(&st[j].x); rl = cig_ref_len(cw); }   // terminates: oo < span
        const int L = min(x + rl, stop) - oo;                      // 1..8 positions under this op
        uint32_t vals = 0x44444444u;                                // '*': deletion, ref-skip, base beyond l_seq
        if (cig_is_match(cw)) {
            const int32_t q = __ldg(&st[j].y) + (oo - x);
            const int Lv = min(L, lseq - q);                        // bases that exist
            if (Lv > 0) { const uint32_t m = nib_mask(Lv); vals = (fetch_codes8(codes, q_base + q) & m) | (vals & ~m); }
        }
        const uint32_t m = nib_mask(L) << (4 * t);
        word = (word & ~m) | ((vals << (4 * t)) & m);
        t += L; oo += L;
    }
    return word;
}

__global__ void __launch_bounds__(kFillThreads) row_fill_kernel(const int32_t* __restrict__ pos, const int32_t* __restrict__ end,
                                                       const int64_t* __restrict__ cigar_off,
                                                       const uint32_t* __restrict__ cigar,
                                                       const int2* __restrict__ opstart,
                                                       const int64_t* __restrict__ seq_off,
                                                       const int32_t* __restrict__ l_seq,
                                                       const uint32_t* __restrict__ codes,
                                                       const int64_t* __restrict__ rowoff,
                                                       const int32_t* __restrict__ nwords,
                                                       uint32_t* __restrict__ rows, int64_t n_reads,
                                                       unsigned long long* __restrict__ next_read) {
    __shared__ int32_t s_idx[kFillIdx + 1];
    __shared__ int32_t s_q[kFillWarps][kFillOps];            // per reference-consuming op of the block: (query index of reference offset 0, read-relative) << 1 | not a match
    __shared__ uint32_t s_mark[kFillWarps][16];              // bit t: an op starts at block position t (t > 0)
    __shared__ long long s_read;
    const int tid = threadIdx.x, lane = tid & 31, wi = tid >> 5;
    for (;;) {
        __syncthreads();                                        // previous read's index and s_read no longer in use
        if (tid == 0) s_read = (long long)atomicAdd(next_read, 1ull);
        __syncthreads();
        const int64_t r = s_read;
        if (r >= n_reads) return;
        const int32_t nw = nwords[r];
        if (nw == 0) continue;
        const int32_t p0 = pos[r], span = end[r] - p0, d = p0 & 7;
        const int64_t c0 = cigar_off[r];
        const int32_t nops = (int32_t)(cigar_off[r + 1] - c0);
        const int32_t lseq = l_seq[r];
        const int64_t q_base = 2 * seq_off[r];                  // absolute index of the read's first base
        const uint32_t* __restrict__ cbase = codes + (q_base >> 3);
        const int32_t qb7 = (int32_t)(q_base & 7);
        uint32_t* __restrict__ out = rows + rowoff[r];
        const int32_t nblk = (nw + 63) >> 6;
        if (nblk > kFillIdx) {                                  // longer than the shared index: every word searches all ops
            for (int32_t w = tid; w < nw; w += kFillThreads) out[w] = fill_word_global(opstart + c0, cigar + c0, nops, 8 * w - d, span, lseq, codes, q_base);
            continue;
        }
        // ---- 1. per block b, the op containing its first position max(512 b - d, 0)
        for (int32_t k = tid; k < nops; k += kFillThreads) {
            const int32_t rl = cig_ref_len(__ldg(cigar + c0 + k));
            if (rl == 0) continue;
            const int32_t x = __ldg(&opstart[c0 + k].x);
            const int32_t b_lo = x == 0 ? 0 : (x + d + 511) >> 9, b_hi = min((x + rl - 1 + d) >> 9, nblk - 1);
            for (int32_t b = b_lo; b <= b_hi; b++) s_idx[b] = k;
        }
        __syncthreads();
        // ---- 2. a warp per block
        for (int32_t b = wi; b < nblk; b += kFillWarps) {
            const int32_t k0 = s_idx[b], k1 = b + 1 < nblk ? s_idx[b + 1] : nops - 1;
            const int32_t n = k1 - k0 + 1, w0 = 64 * b + 2 * lane;
            const int32_t o_blk = 512 * b - d;                  // read-relative offset of the block's first nibble
            if (n > kFillOps) {
                if (w0 < nw) out[w0] = fill_word_global(opstart + c0 + k0, cigar + c0 + k0, n, 8 * w0 - d, span, lseq, codes, q_base);
                if (w0 + 1 < nw) out[w0 + 1] = fill_word_global(opstart + c0 + k0, cigar + c0 + k0, n, 8 * (w0 + 1) - d, span, lseq, codes, q_base);
                continue;
            }
            __syncwarp();                                       // the previous block's tables are no longer in use
            if (lane < 16) s_mark[wi][lane] = 0u;
            __syncwarp();
            int32_t n_ref = 0;
            const int32_t first_cov = max(o_blk, 0);            // the op under the block's first covered position has ordinal 0: no mark
            for (int32_t j0 = 0; j0 < n; j0 += 32) {
                const int32_t j = j0 + lane;
                bool is_ref = false;
                uint32_t cw = 0; int2 st = make_int2(0, 0);
                if (j < n) {
                    cw = __ldg(cigar + c0 + k0 + j);
                    st = __ldg(opstart + c0 + k0 + j);
                    is_ref = cig_ref_len(cw) > 0;
                }
                const uint32_t bal = __ballot_sync(0xffffffffu, is_ref);
                if (is_ref) {
                    s_q[wi][n_ref + __popc(bal & ((1u << lane) - 1u))] = (st.y - st.x) * 2 + (cig_is_match(cw) ? 0 : 1);
                    const int32_t rel = st.x - o_blk;
                    if (st.x > first_cov && rel < 512) atomicOr(&s_mark[wi][rel >> 5], 1u << (rel & 31));
                }
                n_ref += __popc(bal);
            }
            __syncwarp();
            // ordinal of the op at the lane's first position = marks before it
            const uint32_t mw = s_mark[wi][lane >> 1];
            int32_t before = lane < 16 ? __popc(s_mark[wi][lane]) : 0;
            before = warp_incl_scan32(before, lane) - before;                    // marks in the mask words before word `lane`
            before = __shfl_sync(0xffffffffu, before, lane >> 1);
            const int sh16 = 16 * (lane & 1);
            const uint32_t m16 = (mw >> sh16) & 0xFFFFu;                          // op starts inside the lane's 16 positions
            const uint32_t cuts = m16 | 0x100u;                                   // ... and the cut between its two words
            int32_t ord = before + __popc(mw & ((1u << sh16) - 1u));
            const int32_t o_w = o_blk + 16 * lane;                                // read-relative offset of position 0
            const int t_lo = o_w < 0 ? -o_w : 0, t_hi = min(16, span - o_w);      // positions the read covers
            uint32_t lo = 0xFFFFFFFFu, hi = 0xFFFFFFFFu;                          // 0xF = not covered
            if (t_lo < t_hi) {
                ord += __popc(m16 & ((2u << t_lo) - 1u));
                int u = t_lo;
                while (u < t_hi) {
                    const uint32_t rest = cuts >> (u + 1);
                    const int nxt = min(rest ? u + __ffs(rest) : 16, t_hi);
                    const int L = nxt - u;                                        // 1..8 positions of one op inside one word
                    const int32_t e = s_q[wi][ord];
                    uint32_t vals = 0x44444444u;                                  // '*': deletion, ref-skip, base beyond l_seq
                    if (!(e & 1)) {
                        const int32_t q = (e >> 1) + o_w + u;                     // query index within the read
                        const int Lv = min(L, lseq - q);                          // bases that exist
                        if (Lv > 0) {
                            const int32_t Q = q + qb7;
                            const uint32_t m = nib_mask_1to8(Lv);
                            vals = (__funnelshift_r(__ldg(cbase + (Q >> 3)), __ldg(cbase + (Q >> 3) + 1), 4 * (Q & 7)) & m) | (vals & ~m);
                        }
                    }
                    const int un = u & 7;
                    const uint32_t m = nib_mask_1to8(L) << (4 * un);
                    const uint32_t ins = (vals << (4 * un)) & m;
                    if (u < 8) lo = (lo & ~m) | ins; else hi = (hi & ~m) | ins;
                    ord += (int32_t)((m16 >> nxt) & 1u);                          // the cut between the words is not an op start by itself
                    u = nxt;
                }
            }
            if (w0 < nw) out[w0] = lo;
            if (w0 + 1 < nw) out[w0 + 1] = hi;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K1 — column scan.  One CTA (4 warps) per tile of 1024 reference positions; lane owns 8
// consecutive positions (one aligned-row word per read).  Per position: n, A/G/T/C counts,
// alt_freq = max non-ref count / n as a correctly rounded float64 quotient (python int / int,
// generate_SNP_pileups.py:166), neighbour test (:170-179), candidate-eligibility (:183 without the
// chunk bounds, which are applied when chunks are cut).
// ------------------------------------------------------------------------------------------------
constexpr int kTilePos = 1024;
constexpr int kTileThreads = 128;

struct ScanArgs {
    int64_t n_reads;
    const int32_t* pos; const int32_t* end; const uint16_t* flag; const int32_t* pmaxend;
    const int64_t* rowoff; const int32_t* nwords; const uint32_t* rows;
    const uint8_t* ref; int64_t ref_start, ref_len;
    int32_t lo_al, lo, hi;             // scanned 0-based range [lo, hi); lo_al = lo & ~7
    int32_t mincov, haploid;
    double thr_lo, thr_hi, maf;
    uint32_t flag_filter;
    const int32_t* bed; int32_t n_bed; // merged, sorted (start, end) pairs; v excluded iff start <= v < end
    uint8_t* flags; int32_t* tile_nbr; int32_t* tile_cand;
};

__device__ __forceinline__ bool bed_excluded(const int32_t* __restrict__ bed, int32_t n_bed, int32_t v) {
    int32_t lo = 0, hi = n_bed;                               // last interval with start <= v
    while (lo < hi) { int32_t mid = (lo + hi) >> 1; if (__ldg(bed + 2 * mid) <= v) lo = mid + 1; else hi = mid; }
    return lo > 0 && v < __ldg(bed + 2 * (lo - 1) + 1);
}

__global__ void __launch_bounds__(kTileThreads) scan_kernel(const ScanArgs a) {
    __shared__ int32_t s_posw[kTileThreads];
    __shared__ int32_t s_nw[kTileThreads];
    __shared__ int64_t s_rowoff[kTileThreads];
    __shared__ uint8_t s_strand[kTileThreads];
    __shared__ uint16_t s_acc[9][kTilePos];        // A,G,T,C fwd ; A,G,T,C rev ; n
    __shared__ int32_t s_cnt, s_nbr, s_cand;
    __shared__ int64_t s_ilo, s_ihi;

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int32_t P0 = a.lo_al + kTilePos * (int32_t)blockIdx.x;
    const int32_t P1 = min(P0 + kTilePos, a.hi);
    for (int i = tid; i < 9 * kTilePos; i += kTileThreads) (&s_acc[0][0])[i] = 0;
    if (tid == 0) { s_cnt = 0; s_nbr = 0; s_cand = 0; }
    // BAM-index window of the tile, one warp-cooperative search each (32 probes per dependent load instead of 1):
    // first read starting after the tile .. first read whose prefix-max end exceeds P0
    if (w == 0) { const int64_t r = warp_bound_i32(a.pos, a.n_reads, P1 - 1, true, lane); if (lane == 0) s_ihi = r; }
    if (w == 1) { const int64_t r = warp_bound_i32(a.pmaxend, a.n_reads, P0, true, lane); if (lane == 0) s_ilo = r; }
    __syncthreads();
    const int64_t ilo = s_ilo, ihi = s_ihi;
    const int32_t Wbase = (P0 >> 3) + 32 * w;                 // absolute word index of lane 0
    const int32_t myw = Wbase + lane;

    uint32_t cf[4][2], cr[4][2], cn[2];
#pragma unroll
    for (int b = 0; b < 4; b++) { cf[b][0] = cf[b][1] = cr[b][0] = cr[b][1] = 0; }
    cn[0] = cn[1] = 0;
    int pending = 0;
    const int sbase = w * 256 + lane * 8;

    auto flush = [&]() {
#pragma unroll
        for (int j = 0; j < 4; j++) {
#pragma unroll
            for (int b = 0; b < 4; b++) {
                s_acc[b][sbase + 2 * j] += (cf[b][0] >> (8 * j)) & 255u;
                s_acc[b][sbase + 2 * j + 1] += (cf[b][1] >> (8 * j)) & 255u;
                s_acc[4 + b][sbase + 2 * j] += (cr[b][0] >> (8 * j)) & 255u;
                s_acc[4 + b][sbase + 2 * j + 1] += (cr[b][1] >> (8 * j)) & 255u;
            }
            s_acc[8][sbase + 2 * j] += (cn[0] >> (8 * j)) & 255u;
            s_acc[8][sbase + 2 * j + 1] += (cn[1] >> (8 * j)) & 255u;
        }
#pragma unroll
        for (int b = 0; b < 4; b++) { cf[b][0] = cf[b][1] = cr[b][0] = cr[b][1] = 0; }
        cn[0] = cn[1] = 0;
        pending = 0;
    };

    for (int64_t base = ilo; base < ihi; base += kTileThreads) {
        const int64_t i = base + tid;
        if (i < ihi) {
            const int32_t p = __ldg(a.pos + i), e = __ldg(a.end + i);
            const uint32_t f = __ldg(a.flag + i);
            if ((f & a.flag_filter) == 0 && e > P0 && p < P1 && e > p) {
                const int slot = atomicAdd(&s_cnt, 1);
                s_posw[slot] = p >> 3;
                s_nw[slot] = __ldg(a.nwords + i);
                s_rowoff[slot] = __ldg(a.rowoff + i);
                s_strand[slot] = (uint8_t)((f >> 4) & 1u);
            }
        }
        __syncthreads();
        const int cnt = s_cnt;
        for (int j = 0; j < cnt; j++) {
            const int32_t pw = s_posw[j], nw = s_nw[j];
            if (Wbase + 31 < pw || Wbase >= pw + nw) continue;            // warp-uniform
            const int32_t rel = myw - pw;
            uint32_t word = 0xFFFFFFFFu;
            if (rel >= 0 && rel < nw) word = __ldg(a.rows + s_rowoff[j] + rel);
            const uint32_t M1 = 0x11111111u;
            const uint32_t b0 = word & M1, b1 = (word >> 1) & M1, b2 = (word >> 2) & M1, b3 = (word >> 3) & M1;
            const uint32_t cov = b3 ^ M1;                                   // nibble <= 7: the read has a token here
            const uint32_t base4 = cov & ~b2;                               // codes 0..3
            const uint32_t e0 = base4 & ~b1 & ~b0, e1 = base4 & ~b1 & b0, e2 = base4 & b1 & ~b0, e3 = base4 & b1 & b0;
            const uint32_t M8 = 0x01010101u;
            cn[0] += cov & M8; cn[1] += (cov >> 4) & M8;
            if (s_strand[j]) {
                cr[0][0] += e0 & M8; cr[0][1] += (e0 >> 4) & M8;
                cr[1][0] += e1 & M8; cr[1][1] += (e1 >> 4) & M8;
                cr[2][0] += e2 & M8; cr[2][1] += (e2 >> 4) & M8;
                cr[3][0] += e3 & M8; cr[3][1] += (e3 >> 4) & M8;
            } else {
                cf[0][0] += e0 & M8; cf[0][1] += (e0 >> 4) & M8;
                cf[1][0] += e1 & M8; cf[1][1] += (e1 >> 4) & M8;
                cf[2][0] += e2 & M8; cf[2][1] += (e2 >> 4) & M8;
                cf[3][0] += e3 & M8; cf[3][1] += (e3 >> 4) & M8;
            }
            if (++pending == 255) flush();
        }
        __syncthreads();
        if (tid == 0) s_cnt = 0;
        __syncthreads();
    }
    flush();

    // finalize the lane's 8 positions
    uint32_t out_lo = 0, out_hi = 0;
    int my_nbr = 0, my_cand = 0;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        const int32_t p = P0 + sbase + t;
        uint32_t fl = 0;
        const int64_t ri = (int64_t)p - a.ref_start;
        if (p >= a.lo && p < a.hi && ri >= 0 && ri < a.ref_len) {
            const int rc = ref_code_of(__ldg(a.ref + ri));
            const int32_t n = s_acc[8][sbase + t];
            if (rc < 4 && n > 0 && n >= a.mincov) {
                int32_t alt = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const int32_t c = (int32_t)s_acc[b][sbase + t] + (int32_t)s_acc[4 + b][sbase + t];
                    if (b != rc) alt = max(alt, c);
                }
                const double fq = (double)alt / (double)n;
                const bool nbr = a.thr_lo <= fq && (a.haploid || fq < a.thr_hi);
                const bool cand = a.maf <= fq;
                if ((nbr || cand) && !(a.n_bed > 0 && bed_excluded(a.bed, a.n_bed, p + 1))) {
                    fl = (nbr ? 1u : 0u) | (cand ? 2u : 0u);
                    my_nbr += nbr; my_cand += cand;
                }
            }
        }
        if (t < 4) out_lo |= fl << (8 * t); else out_hi |= fl << (8 * (t - 4));
    }
    *reinterpret_cast<uint2*>(a.flags + (size_t)(P0 - a.lo_al) + sbase) = make_uint2(out_lo, out_hi);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        my_nbr += __shfl_xor_sync(0xffffffffu, my_nbr, d);
        my_cand += __shfl_xor_sync(0xffffffffu, my_cand, d);
    }
    if (lane == 0) { atomicAdd(&s_nbr, my_nbr); atomicAdd(&s_cand, my_cand); }
    __syncthreads();
    if (tid == 0) { a.tile_nbr[blockIdx.x] = s_nbr; a.tile_cand[blockIdx.x] = s_cand; }
}

// K1w — ordered compaction of the flagged positions into the neighbour list and the candidate list
// (1-based v_pos, ascending), using the scanned per-tile counts.
__global__ void __launch_bounds__(kTileThreads) site_list_kernel(const uint8_t* __restrict__ flags, int32_t lo_al,
                                                                 const int64_t* __restrict__ nbr_off,
                                                                 const int64_t* __restrict__ cand_off,
                                                                 int32_t* __restrict__ nbr_pos,
                                                                 int32_t* __restrict__ cand_pos) {
    __shared__ int64_t sm[33];
    const int tid = threadIdx.x;
    const size_t idx = (size_t)blockIdx.x * kTilePos + (size_t)tid * 8;
    const uint2 f = *reinterpret_cast<const uint2*>(flags + idx);
    const uint64_t bits = ((uint64_t)f.y << 32) | f.x;
    const int nn = __popcll(bits & 0x0101010101010101ull), nc_ = __popcll(bits & 0x0202020202020202ull);
    int64_t tot;
    int64_t on = block_excl_scan64(nn, sm, tot) + nbr_off[blockIdx.x];
    int64_t oc = block_excl_scan64(nc_, sm, tot) + cand_off[blockIdx.x];
    const int32_t v0 = lo_al + (int32_t)idx + 1;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        const uint32_t fl = (uint32_t)(bits >> (8 * t)) & 255u;
        if (fl & 1u) nbr_pos[on++] = v0 + t;
        if (fl & 2u) cand_pos[oc++] = v0 + t;
    }
}

// ------------------------------------------------------------------------------------------------
// K1b — neighbour matrix: for every read, its codes at the neighbour sites inside its span.
// ------------------------------------------------------------------------------------------------
__global__ void nmat_len_kernel(int64_t n_reads, const int32_t* __restrict__ pos, const int32_t* __restrict__ end,
                                const int32_t* __restrict__ nbr_pos, int32_t n_nbr, int32_t* __restrict__ nfirst,
                                int32_t* __restrict__ nlen, int32_t* __restrict__ nbytes) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const int32_t p = pos[r], e = end[r];
    int32_t jf = 0, len = 0;
    if (e > p) {
        jf = lower_bound_i32(nbr_pos, n_nbr, p + 1);          // v_pos - 1 >= pos
        len = lower_bound_i32(nbr_pos, n_nbr, e + 1) - jf;    // v_pos - 1 <  end
    }
    nfirst[r] = jf; nlen[r] = len; nbytes[r] = (len + 1) >> 1;
}
__global__ void nmat_fill_kernel(int64_t n_reads, const int32_t* __restrict__ pos, const int64_t* __restrict__ rowoff,
                                 const uint32_t* __restrict__ rows, const int32_t* __restrict__ nbr_pos,
                                 const int32_t* __restrict__ nfirst, const int32_t* __restrict__ nlen,
                                 const int64_t* __restrict__ noff, uint8_t* __restrict__ nrows) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const int32_t len = nlen[r];
    if (len == 0) return;
    const int32_t jf = nfirst[r], pw = pos[r] >> 3;
    const uint32_t* __restrict__ row = rows + rowoff[r];
    uint8_t* __restrict__ out = nrows + noff[r];
    for (int32_t j = 0; j < len; j += 2) {
        const int32_t p0 = __ldg(nbr_pos + jf + j) - 1;
        uint32_t b = (__ldg(row + ((p0 >> 3) - pw)) >> (4 * (p0 & 7))) & 15u;
        if (j + 1 < len) {
            const int32_t p1 = __ldg(nbr_pos + jf + j + 1) - 1;
            b |= ((__ldg(row + ((p1 >> 3) - pw)) >> (4 * (p1 & 7))) & 15u) << 4;
        } else {
            b |= 0xF0u;
        }
        out[j >> 1] = (uint8_t)b;
    }
}

// ------------------------------------------------------------------------------------------------
// Chunk cut: candidates of chunk c are the eligible sites with start <= v_pos <= end
// (generate_SNP_pileups.py:183); a site on a shared chunk boundary belongs to both chunks.
// ------------------------------------------------------------------------------------------------
__global__ void chunk_ranges_kernel(const NcChunk* __restrict__ chunks, int32_t n_chunks,
                                    const int32_t* __restrict__ cand_pos, int64_t n_cand,
                                    int32_t* __restrict__ chunk_lo, int32_t* __restrict__ chunk_cnt) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    const int64_t lo = lower_bound_i32_64(cand_pos, n_cand, chunks[c].start);
    const int64_t hi = upper_bound_i32_64(cand_pos, n_cand, chunks[c].end);
    chunk_lo[c] = (int32_t)lo;
    chunk_cnt[c] = hi > lo ? (int32_t)(hi - lo) : 0;
}

// ------------------------------------------------------------------------------------------------
// get_cnd_pos (generate_SNP_pileups.py:6-101) as distance bins: neighbours p with a < |p - v| <= b,
// keep k of them, `far` = the k farthest inside the bin instead of the k nearest.  The outermost
// bin is bounded by the strict radius test, hence b = R - 1.
// ------------------------------------------------------------------------------------------------
struct BinSpec { int32_t a, b, k, far; };
__constant__ BinSpec c_bins[5][7] = {
    /* ont            */ {{0, 2000, 2, 1}, {2000, 5000, 3, 0}, {5000, 10000, 4, 0}, {10000, 20000, 5, 0}, {20000, 49999, 6, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}},
    /* short_ont      */ {{0, 2000, 5, 0}, {2000, 5000, 10, 0}, {5000, 49999, 5, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}},
    /* ul_ont         */ {{0, 2000, 2, 1}, {2000, 5000, 2, 0}, {5000, 10000, 3, 0}, {10000, 20000, 3, 0}, {20000, 40000, 4, 0}, {40000, 50000, 3, 0}, {50000, 99999, 3, 0}},
    /* ul_ont_extreme */ {{0, 10000, 2, 1}, {10000, 20000, 2, 0}, {20000, 50000, 3, 0}, {50000, 75000, 3, 0}, {75000, 100000, 4, 0}, {100000, 200000, 4, 0}, {200000, 299999, 2, 0}},
    /* pacbio         */ {{0, 2000, 4, 1}, {2000, 5000, 5, 0}, {5000, 10000, 5, 0}, {10000, 19999, 6, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}},
};
__constant__ int c_nbins[5] = {5, 3, 7, 7, 4};

// Warp-cooperative neighbour choice.  Lanes 0..6 own the left bins, lanes 8..14 the right bins.
// On return every lane holds (sel_lo, sel_cnt) of its bin (cnt 0 elsewhere); n_left / n_right are
// warp-uniform.  wlo/whi = inclusive v_pos bounds of the chunk's pileup window (:156).
// `nbr_pos` may point to global or shared memory (plain loads): any sorted sub-range of the neighbour
// list that contains every neighbour within the search radius of v gives the same choice.
__device__ __forceinline__ int32_t lower_bound_plain(const int32_t* a, int32_t n, int32_t key) {
    int32_t lo = 0, hi = n;
    while (lo < hi) { const int32_t mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}
__device__ __forceinline__ void choose_neighbours(const int32_t* nbr_pos, int32_t n_nbr, int seq,
                                                  int32_t v, int32_t wlo, int32_t whi, int lane,
                                                  int32_t& sel_lo, int32_t& sel_cnt, int& n_left, int& n_right) {
    sel_lo = 0; sel_cnt = 0;
    const int side = lane >> 3, bin = lane & 7;
    if (lane < 16 && bin < c_nbins[seq]) {
        const BinSpec bs = c_bins[seq][bin];
        int32_t lo, hi;
        if (side == 0) {                                               // v-b <= p < v-a
            const int64_t from = max((int64_t)v - bs.b, (int64_t)wlo);
            lo = lower_bound_plain(nbr_pos, n_nbr, (int32_t)max(from, (int64_t)INT32_MIN + 1));
            hi = lower_bound_plain(nbr_pos, n_nbr, v - bs.a);
            if (hi > lo) { sel_cnt = min(hi - lo, bs.k); sel_lo = bs.far ? lo : hi - sel_cnt; }
        } else {                                                       // v+a < p <= v+b
            const int64_t to = min((int64_t)v + bs.b, (int64_t)whi);
            lo = lower_bound_plain(nbr_pos, n_nbr, v + bs.a + 1);
            hi = lower_bound_plain(nbr_pos, n_nbr, (int32_t)min(to + 1, (int64_t)INT32_MAX));
            if (hi > lo) { sel_cnt = min(hi - lo, bs.k); sel_lo = bs.far ? hi - sel_cnt : lo; }
        }
    }
    const uint32_t full = 0xffffffffu;
    int l = 0, r = 0;
#pragma unroll
    for (int b = 0; b < 7; b++) {
        l += __shfl_sync(full, sel_cnt, b);
        r += __shfl_sync(full, sel_cnt, 8 + b);
    }
    n_left = l; n_right = r;
}

// Neighbour-list index of tensor column c (0..40) or -1 for padding / the candidate column.
__device__ __forceinline__ int32_t column_neighbour(int c, int seq, int32_t sel_lo, int32_t sel_cnt, int n_left, int n_right) {
    const uint32_t full = 0xffffffffu;
    const int nb = c_nbins[seq];
    int32_t j = -1;
    int il = c - (20 - n_left);          // rank inside the sorted left list
    int ir = c - 21;                     // rank inside the sorted right list
    const bool is_l = c < 20 && il >= 0, is_r = c > 20 && c < 41 && ir < n_right;
#pragma unroll
    for (int b = 6; b >= 0; b--) {       // left list ascending = farthest bin first
        const int32_t lo = __shfl_sync(full, sel_lo, b), cnt = __shfl_sync(full, sel_cnt, b);
        if (b < nb && is_l && j < 0) { if (il < cnt) j = lo + il; else il -= cnt; }
    }
#pragma unroll
    for (int b = 0; b < 7; b++) {        // right list ascending = nearest bin first
        const int32_t lo = __shfl_sync(full, sel_lo, 8 + b), cnt = __shfl_sync(full, sel_cnt, 8 + b);
        if (b < nb && is_r && j < 0) { if (ir < cnt) j = lo + ir; else ir -= cnt; }
    }
    return j;
}

struct TensorArgs {
    int64_t n_reads;
    const int32_t* pos; const int32_t* end; const uint16_t* flag; const int32_t* pmaxend;
    const int64_t* rowoff; const uint32_t* rows;
    const int32_t* nfirst; const int32_t* nlen; const int64_t* noff; const uint8_t* nrows;
    const int32_t* nbr_pos; int32_t n_nbr;
    const int32_t* cand_pos;
    const NcChunk* chunks; int32_t n_chunks;
    const int64_t* chunk_off;          // [n_chunks+1] slot offsets (before the min_nbr_sites filter)
    const int32_t* chunk_lo;           // first candidate-list index of every chunk
    const int64_t* outidx;             // [n_slots+1] output row of every slot, or nullptr = identity
    const uint8_t* ref; int64_t ref_start, ref_len;
    int32_t seq, maxcov, min_nbr_sites;
    uint32_t flag_filter;
    int64_t n_slots;
    int16_t* mat; NcSiteMeta* meta;
    unsigned long long* chunk_depth_sum; unsigned long long* chunk_count;
    uint8_t* keep;                     // K2a only
};

__device__ __forceinline__ int slot_chunk(const int64_t* __restrict__ chunk_off, int32_t n_chunks, int64_t s) {
    int lo = 0, hi = n_chunks;           // last chunk with chunk_off[c] <= s
    while (lo < hi) { int mid = (lo + hi) >> 1; if (__ldg(chunk_off + mid) <= s) lo = mid + 1; else hi = mid; }
    return lo - 1;
}

// K2a — only when min_nbr_sites > 1: keep[s] = len(total_rlist) >= min_nbr_sites (:244).
__global__ void __launch_bounds__(128) keep_kernel(const TensorArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= a.n_slots) return;
    const int c = slot_chunk(a.chunk_off, a.n_chunks, s);
    const int32_t v = __ldg(a.cand_pos + a.chunk_lo[c] + (s - a.chunk_off[c]));
    const int32_t wlo = max(1, a.chunks[c].start - 50000), whi = a.chunks[c].end + 50000;
    int32_t sel_lo, sel_cnt; int nl, nr;
    choose_neighbours(a.nbr_pos, a.n_nbr, a.seq, v, wlo, whi, lane, sel_lo, sel_cnt, nl, nr);
    if (lane == 0) a.keep[s] = (nl + nr + 1 >= a.min_nbr_sites) ? 1 : 0;
}
__global__ void keep_to_i32_kernel(const uint8_t* __restrict__ keep, int64_t n, int32_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = keep[i];
}

// K2 — tensor build (generate_SNP_pileups.py:200-263), one warp per candidate site.
//   lanes as READS   walk the reads that can cover the candidate, 32 at a time: code at the candidate
//                    (one nibble of the aligned row), strand depths;
//   lanes as COLUMNS for every sampled read (first maxcov in BAM order, see DESIGN.md on :215-216) lane c
//                    looks up the code at column c in the read's neighbour-matrix row (columns 32..40
//                    ride on lanes 0..8) and counts into 4 x 4 packed fields.
// A CTA takes a RUN of consecutive slots (same region of the contig) and prepares once what the per-site
// version recomputed for every site: the admitted reads overlapping the run, compacted in BAM order in
// shared memory (~40 entries at 30x instead of a ~250-read BAM-index window per site), and the sub-range
// of the neighbour list within the search radius of the run (binary searches over ~100 shared-memory
// entries instead of the whole list).  Runs that do not fit (span > 64 kb, > 256 overlapping reads,
// maxcov > 255) take the per-site generic path.
constexpr int kTensorWarps = 4;
constexpr int kRunSlots = 32;
constexpr int kRunList = 256;
constexpr int kRunNbr = 1024;
constexpr int kMaskWords = 4;             // bit-parallel path: run lists of up to 128 reads as 4 x 32-bit masks
constexpr int kMaskNbr = 192;             // ... and up to 256 staged neighbour sites

struct SiteCols {                       // what a lane knows about its two tensor columns (lane, lane + 32)
    int32_t j0, j1;                     // neighbour-list index (global numbering) or -1
    int rc0, rc1, rc_v, nl, nr;
};
__device__ __forceinline__ uint64_t expand_fields_8_to_16(uint32_t x) {
    return (uint64_t)(x & 0xFFu) | ((uint64_t)(x & 0xFF00u) << 8) | ((uint64_t)(x & 0xFF0000u) << 16) | ((uint64_t)(x & 0xFF000000u) << 24);
}

// neighbour choice + the lane's columns; nbr = sorted neighbour positions [0, n), nb_base = global index of nbr[0]
__device__ __forceinline__ SiteCols site_columns(const TensorArgs& a, const int32_t* nbr, int32_t n, int32_t nb_base, int c, int32_t v, int lane) {
    SiteCols sc;
    const int32_t wlo = max(1, a.chunks[c].start - 50000), whi = a.chunks[c].end + 50000;
    int32_t sel_lo, sel_cnt;
    choose_neighbours(nbr, n, a.seq, v, wlo, whi, lane, sel_lo, sel_cnt, sc.nl, sc.nr);
    const int32_t r0 = column_neighbour(lane, a.seq, sel_lo, sel_cnt, sc.nl, sc.nr);
    const int32_t r1 = column_neighbour(lane + 32, a.seq, sel_lo, sel_cnt, sc.nl, sc.nr);
    sc.rc_v = ref_code_of(__ldg(a.ref + ((int64_t)(v - 1) - a.ref_start)));
    sc.rc0 = 4; sc.rc1 = 4;
    if (lane == 20) sc.rc0 = sc.rc_v;
    if (r0 >= 0) sc.rc0 = ref_code_of(__ldg(a.ref + ((int64_t)nbr[r0] - 1 - a.ref_start)));
    if (r1 >= 0) sc.rc1 = ref_code_of(__ldg(a.ref + ((int64_t)nbr[r1] - 1 - a.ref_start)));
    sc.j0 = r0 >= 0 ? r0 + nb_base : -1;
    sc.j1 = r1 >= 0 ? r1 + nb_base : -1;
    return sc;
}

// Same result as site_columns for the sites of a RUN whose neighbour range is staged in shared memory (nbr[0, n), reference codes
// nbr_rc): a warp visits its sites in ascending (chunk, position) order, so the 2 x 7 bin boundaries only move forward — every
// bin lane keeps its two list indices (st_lo, st_hi) from the previous site and advances them instead of running two binary
// searches per bin and site (first = true: search).  The tensor column of every chosen neighbour comes from a prefix sum over the
// bins' counts (left list = farthest bin first, right list = nearest bin first) written to a 41-entry table, instead of a walk
// over all bins per column; reference codes come from the staged copies instead of dependent global loads.
__device__ __forceinline__ SiteCols site_columns_run(const TensorArgs& a, const int32_t* nbr, const uint8_t* nbr_rc, int32_t n, int32_t nb_base, int c, int32_t v,
                                                     int rc_v, int lane, bool first, int32_t& st_lo, int32_t& st_hi, int8_t* colj) {
    const uint32_t full = 0xffffffffu;
    SiteCols sc;
    const int32_t wlo = max(1, a.chunks[c].start - 50000), whi = a.chunks[c].end + 50000;
    const int side = lane >> 3, bin = lane & 7, nb = c_nbins[a.seq];
    const bool owner = lane < 16 && bin < nb;
    int32_t sel_lo = 0, sel_cnt = 0;
    if (owner) {
        const BinSpec bs = c_bins[a.seq][bin];
        int32_t key_lo, key_hi;
        if (side == 0) { key_lo = (int32_t)max(max((int64_t)v - bs.b, (int64_t)wlo), (int64_t)INT32_MIN + 1); key_hi = v - bs.a; }
        else { key_lo = v + bs.a + 1; key_hi = (int32_t)min(min((int64_t)v + bs.b, (int64_t)whi) + 1, (int64_t)INT32_MAX); }
        if (first) { st_lo = lower_bound_plain(nbr, n, key_lo); st_hi = lower_bound_plain(nbr, n, key_hi); }
        else {
            while (st_lo < n && nbr[st_lo] < key_lo) st_lo++;
            while (st_hi < n && nbr[st_hi] < key_hi) st_hi++;
        }
        if (st_hi > st_lo) {
            sel_cnt = min(st_hi - st_lo, bs.k);
            sel_lo = side == 0 ? (bs.far ? st_lo : st_hi - sel_cnt) : (bs.far ? st_hi - sel_cnt : st_lo);
        }
    }
    // left list ascending = farthest bin first: bin b starts after the bins b' > b; right list: after the bins b' < b
    int before = 0, l = 0, r = 0;
#pragma unroll
    for (int b = 0; b < 7; b++) {
        const int cl = __shfl_sync(full, sel_cnt, b), cr = __shfl_sync(full, sel_cnt, 8 + b);
        l += cl; r += cr;
        if (side == 0 && b > bin) before += cl;
        if (side == 1 && b < bin) before += cr;
    }
    sc.nl = l; sc.nr = r;
    colj[lane] = -1;
    if (lane < 9) colj[32 + lane] = -1;
    __syncwarp();
    if (owner) {
        const int col0 = side == 0 ? 20 - l + before : 21 + before;
        int16_t* colw = reinterpret_cast<int16_t*>(colj + 48);
        for (int t = 0; t < sel_cnt; t++) { colj[col0 + t] = 0; colw[col0 + t] = (int16_t)(sel_lo + t); }
    }
    __syncwarp();
    const int16_t* colv = reinterpret_cast<const int16_t*>(colj + 48);
    const int32_t r0 = colj[lane] == 0 ? (int32_t)colv[lane] : -1;
    const int32_t r1 = (lane < 9 && colj[32 + lane] == 0) ? (int32_t)colv[32 + lane] : -1;
    sc.rc_v = rc_v;
    sc.rc0 = 4; sc.rc1 = 4;
    if (lane == 20) sc.rc0 = rc_v;
    if (r0 >= 0) sc.rc0 = nbr_rc[r0];
    if (r1 >= 0) sc.rc1 = nbr_rc[r1];
    sc.j0 = r0 >= 0 ? r0 + nb_base : -1;
    sc.j1 = r1 >= 0 ? r1 + nb_base : -1;
    __syncwarp();                                                      // the table is reused by the warp's next site
    return sc;
}

// assemble [5][41][5] in shared memory, 16-byte coalesced stores, site metadata, chunk depth sums
__device__ __forceinline__ void site_finish(const TensorArgs& a, int64_t orow, int c, int32_t v, int lane, int16_t* buf, const SiteCols& sc,
                                            const uint64_t* acc0, const uint64_t* acc1, uint64_t fwd, uint64_t rev, int32_t dp, int32_t sampled) {
    const uint32_t full = 0xffffffffu;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        fwd += __shfl_xor_sync(full, fwd, d);
        rev += __shfl_xor_sync(full, rev, d);
    }
    for (int i = lane; i < NC_SNP_SITE_STRIDE / 8; i += 32) reinterpret_cast<uint4*>(buf)[i] = make_uint4(0, 0, 0, 0);
    __syncwarp();
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const int col = lane + 32 * half;
        const int rc = half ? sc.rc1 : sc.rc0;
        const bool real = half ? (sc.j1 >= 0) : (sc.j0 >= 0 || lane == 20);
        if (real && col < NC_SNP_COLS) {
            if (rc < 4) buf[col * 5 + rc] = 1;                                          // :249-251 total_ref
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint64_t av = half ? acc1[i] : acc0[i];
                int16_t* o = buf + ((i + 1) * NC_SNP_COLS + col) * 5;
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const int32_t cnt = (int32_t)((av >> (16 * b)) & 0xFFFFull);
                    o[b] = (int16_t)(b == rc ? -cnt : cnt);                             // :253 mat * (1 - 2*total_ref)
                }
                o[4] = (i == sc.rc_v) ? 1 : 0;                                          // :252
            }
        }
    }
    __syncwarp();
    uint4* dst = reinterpret_cast<uint4*>(a.mat + orow * NC_SNP_SITE_STRIDE);
    for (int i = lane; i < NC_SNP_SITE_STRIDE / 8; i += 32) dst[i] = reinterpret_cast<const uint4*>(buf)[i];
    __syncwarp();
    if (lane == 0) {
        NcSiteMeta m;
        m.pos = v; m.chunk = c; m.dp = dp;
        int32_t alt = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const uint32_t fb = (uint32_t)((fwd >> (16 * b)) & 0xFFFFull), rb = (uint32_t)((rev >> (16 * b)) & 0xFFFFull);
            m.fwd[b] = (uint16_t)fb; m.rev[b] = (uint16_t)rb;
            if (b != sc.rc_v) alt = max(alt, (int32_t)(fb + rb));
        }
        m.alt = alt;
        m.ref_code = (uint8_t)sc.rc_v; m.n_left = (uint8_t)sc.nl; m.n_right = (uint8_t)sc.nr; m.reserved = 0;
        m.sample_depth = sampled;
        a.meta[orow] = m;
        atomicAdd(a.chunk_depth_sum + c, (unsigned long long)sampled);
        atomicAdd(a.chunk_count + c, 1ull);
    }
}

// keep the lowest `room` set bits of cm (sample = the first maxcov covering reads in BAM order)
__device__ __forceinline__ uint32_t take_first(uint32_t cm, int room) {
    if (__popc(cm) <= room) return cm;
    uint32_t m = cm, kept = 0;
    for (int q = 0; q < room; q++) { const uint32_t low = m & (0u - m); kept |= low; m ^= low; }
    return room > 0 ? kept : 0u;
}

// Generic per-site path: BAM-index window from the prefix-max of read ends, 16-bit count fields.
__device__ __forceinline__ void tensor_site_generic(const TensorArgs& a, int64_t orow, int c, int32_t v, int lane, int16_t* buf) {
    const uint32_t full = 0xffffffffu;
    const int32_t p = v - 1;
    const SiteCols sc = site_columns(a, a.nbr_pos, a.n_nbr, 0, c, v, lane);
    const int64_t ihi = upper_bound_i32_64(a.pos, a.n_reads, p);
    int64_t ilo;
    {
        int64_t lo = 0, hi = ihi;
        while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (__ldg(a.pmaxend + mid) <= p) lo = mid + 1; else hi = mid; }
        ilo = lo;
    }
    uint64_t acc0[4] = {0, 0, 0, 0}, acc1[4] = {0, 0, 0, 0};
    uint64_t fwd = 0, rev = 0;                                        // per-lane 4 x 16-bit strand depths
    int32_t dp = 0, sampled = 0;
    for (int64_t base = ilo; base < ihi; base += 32) {
        const int64_t i = base + lane;
        bool cover = false;
        uint32_t code = 4;
        int32_t nf = 0, nlen = 0;
        int64_t noff = 0;
        if (i < ihi) {
            const int32_t rp = __ldg(a.pos + i), re = __ldg(a.end + i);
            const uint32_t f = __ldg(a.flag + i);
            if ((f & a.flag_filter) == 0 && rp <= p && p < re) {
                cover = true;
                code = (__ldg(a.rows + __ldg(a.rowoff + i) + ((p >> 3) - (rp >> 3))) >> (4 * (p & 7))) & 15u;
                if (code < 4) { if (f & 0x10u) rev += 1ull << (16 * code); else fwd += 1ull << (16 * code); }
                nf = __ldg(a.nfirst + i); nlen = __ldg(a.nlen + i); noff = __ldg(a.noff + i);
            }
        }
        const uint32_t cm = __ballot_sync(full, cover);
        if (cm == 0) continue;
        dp += __popc(cm);
        uint32_t take = take_first(cm, a.maxcov - sampled);
        sampled += __popc(take);
        while (take) {
            const int k = __ffs(take) - 1;
            take &= take - 1;
            const uint32_t ci = __shfl_sync(full, code, k);
            const int32_t rnf = __shfl_sync(full, nf, k), rnl = __shfl_sync(full, nlen, k);
            const int64_t rno = __shfl_sync(full, noff, k);
            if (ci >= 4) continue;                                     // '*' / N at the candidate: counted in depth only
            uint32_t b0 = 4, b1 = 4;
            if (lane == 20) b0 = ci;
            if (sc.j0 >= 0) {
                const uint32_t rel = (uint32_t)(sc.j0 - rnf);
                if (rel < (uint32_t)rnl) { const uint32_t by = __ldg(a.nrows + rno + (rel >> 1)); b0 = (rel & 1u) ? (by >> 4) : (by & 15u); }
            }
            if (sc.j1 >= 0) {
                const uint32_t rel = (uint32_t)(sc.j1 - rnf);
                if (rel < (uint32_t)rnl) { const uint32_t by = __ldg(a.nrows + rno + (rel >> 1)); b1 = (rel & 1u) ? (by >> 4) : (by & 15u); }
            }
            const uint64_t i0 = b0 < 4 ? 1ull << (16 * b0) : 0ull, i1 = b1 < 4 ? 1ull << (16 * b1) : 0ull;
            switch (ci) {                                              // warp-uniform
                case 0: acc0[0] += i0; acc1[0] += i1; break;
                case 1: acc0[1] += i0; acc1[1] += i1; break;
                case 2: acc0[2] += i0; acc1[2] += i1; break;
                default: acc0[3] += i0; acc1[3] += i1; break;
            }
        }
    }
    site_finish(a, orow, c, v, lane, buf, sc, acc0, acc1, fwd, rev, dp, sampled);
}

struct RunList {                        // admitted reads overlapping the run, BAM order (shared memory)
    int32_t rp[kRunList], re[kRunList], nf[kRunList], nl[kRunList];
    int64_t rowoff[kRunList], noff[kRunList];
    uint8_t rev[kRunList];
};

// Fast per-site path over the run's compact read list; 8-bit count fields (maxcov <= 255).
__device__ __forceinline__ void tensor_site_run(const TensorArgs& a, int64_t orow, int c, int32_t v, int lane, int16_t* buf,
                                                const RunList& L, int cnt, const SiteCols& sc) {
    const uint32_t full = 0xffffffffu;
    const int32_t p = v - 1;
    uint32_t a0[4] = {0, 0, 0, 0}, a1[4] = {0, 0, 0, 0};                // per candidate code: 4 x 8-bit counts by column code
    uint64_t fwd = 0, rev = 0;
    int32_t dp = 0, sampled = 0;
    for (int base = 0; base < cnt; base += 32) {
        const int e = base + lane;
        bool cover = false;
        uint32_t code = 4;
        if (e < cnt) {
            const int32_t rp = L.rp[e];
            if (rp <= p && p < L.re[e]) {
                cover = true;
                code = (__ldg(a.rows + L.rowoff[e] + ((p >> 3) - (rp >> 3))) >> (4 * (p & 7))) & 15u;
                if (code < 4) { if (L.rev[e]) rev += 1ull << (16 * code); else fwd += 1ull << (16 * code); }
            }
        }
        const uint32_t cm = __ballot_sync(full, cover);
        if (cm == 0) continue;
        dp += __popc(cm);
        uint32_t take = take_first(cm, a.maxcov - sampled);
        sampled += __popc(take);
        while (take) {
            const int k = __ffs(take) - 1;
            take &= take - 1;
            const uint32_t ci = __shfl_sync(full, code, k);
            if (ci >= 4) continue;                                     // '*' / N at the candidate: counted in depth only
            const int e2 = base + k;
            const int32_t rnf = L.nf[e2], rnl = L.nl[e2];
            const uint8_t* __restrict__ row = a.nrows + L.noff[e2];
            uint32_t b0 = 4, b1 = 4;
            if (lane == 20) b0 = ci;
            if (sc.j0 >= 0) {
                const uint32_t rel = (uint32_t)(sc.j0 - rnf);
                if (rel < (uint32_t)rnl) { const uint32_t by = __ldg(row + (rel >> 1)); b0 = (rel & 1u) ? (by >> 4) : (by & 15u); }
            }
            if (sc.j1 >= 0) {
                const uint32_t rel = (uint32_t)(sc.j1 - rnf);
                if (rel < (uint32_t)rnl) { const uint32_t by = __ldg(row + (rel >> 1)); b1 = (rel & 1u) ? (by >> 4) : (by & 15u); }
            }
            const uint32_t i0 = b0 < 4 ? 1u << (8 * b0) : 0u, i1 = b1 < 4 ? 1u << (8 * b1) : 0u;
            switch (ci) {                                              // warp-uniform
                case 0: a0[0] += i0; a1[0] += i1; break;
                case 1: a0[1] += i0; a1[1] += i1; break;
                case 2: a0[2] += i0; a1[2] += i1; break;
                default: a0[3] += i0; a1[3] += i1; break;
            }
        }
    }
    uint64_t acc0[4], acc1[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { acc0[i] = expand_fields_8_to_16(a0[i]); acc1[i] = expand_fields_8_to_16(a1[i]); }
    site_finish(a, orow, c, v, lane, buf, sc, acc0, acc1, fwd, rev, dp, sampled);
}

// Bit-parallel per-site path.  With the run's reads numbered by their position in the compact list, "which reads have code b
// at neighbour site j" is a bit mask M[j][b] (built once per run, a ballot per code with lanes as reads), "which sampled reads
// have code i at the candidate" is a mask MS[i] (four ballots per site), and every tensor entry is
//     mat[i][column of j][b] = popcount(MS[i] & M[j][b])                      (generate_SNP_pileups.py:221-247)
// — 16 AND + POPC per column and mask word instead of a dependent lookup per (read, column).
template <int W>
__device__ __forceinline__ void tensor_site_masks(const TensorArgs& a, int64_t orow, int c, int32_t v, int lane, int16_t* buf,
                                                  const RunList& L, int cnt, int32_t nb_base, const SiteCols& sc,
                                                  const uint32_t* __restrict__ M) {
    const uint32_t full = 0xffffffffu;
    const int32_t p = v - 1;
    uint32_t cw[W], rv[W], ms[4][W];
#pragma unroll
    for (int w = 0; w < W; w++) {
        cw[w] = 0; rv[w] = 0; ms[0][w] = ms[1][w] = ms[2][w] = ms[3][w] = 0;
        if (32 * w < cnt) {                                            // warp-uniform
            const int e = 32 * w + lane;
            bool cover = false, rev = false;
            uint32_t code = 4;
            if (e < cnt) {
                const int32_t rp = L.rp[e];
                if (rp <= p && p < L.re[e]) {
                    cover = true;
                    rev = L.rev[e] != 0;
                    code = (__ldg(a.rows + L.rowoff[e] + ((p >> 3) - (rp >> 3))) >> (4 * (p & 7))) & 15u;
                }
            }
            cw[w] = __ballot_sync(full, cover);
            rv[w] = __ballot_sync(full, cover && rev);
            ms[0][w] = __ballot_sync(full, code == 0); ms[1][w] = __ballot_sync(full, code == 1);
            ms[2][w] = __ballot_sync(full, code == 2); ms[3][w] = __ballot_sync(full, code == 3);
        }
    }
    // depth, strand depths over ALL covering reads (:210-213), then the sample = first maxcov covering reads in BAM order (:215-216)
    int32_t dp = 0;
    uint64_t fwd = 0, rev = 0;
#pragma unroll
    for (int w = 0; w < W; w++) {
        dp += __popc(cw[w]);
#pragma unroll
        for (int b = 0; b < 4; b++) {
            fwd += (uint64_t)__popc(ms[b][w] & ~rv[w]) << (16 * b);
            rev += (uint64_t)__popc(ms[b][w] & rv[w]) << (16 * b);
        }
    }
    if (lane != 0) { fwd = 0; rev = 0; }                               // site_finish sums the lanes
    int32_t sampled = dp;
    if (dp > a.maxcov) {
        sampled = a.maxcov;
        int room = a.maxcov;
#pragma unroll
        for (int w = 0; w < W; w++) {
            const uint32_t keep = take_first(cw[w], room);
            room -= __popc(keep);
#pragma unroll
            for (int i = 0; i < 4; i++) ms[i][w] &= keep;
        }
    }
    // lanes as columns
    uint64_t acc0[4] = {0, 0, 0, 0}, acc1[4] = {0, 0, 0, 0};
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const int32_t j = half ? sc.j1 : sc.j0;
        uint64_t* acc = half ? acc1 : acc0;
        if (half == 0 && lane == 20) {                                 // the candidate column: a read's column code is its candidate code
#pragma unroll
            for (int i = 0; i < 4; i++) {
                int n = 0;
#pragma unroll
                for (int w = 0; w < W; w++) n += __popc(ms[i][w]);
                acc[i] = (uint64_t)n << (16 * i);
            }
        } else if (j >= 0) {
            const uint32_t* mj = M + (size_t)(j - nb_base) * 4 * kMaskWords;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                uint32_t m[W];
                if (W == 1) m[0] = mj[b * kMaskWords];
                else if (W == 2) { const uint2 t2 = *reinterpret_cast<const uint2*>(mj + b * kMaskWords); m[0] = t2.x; m[W - 1] = t2.y; }
                else { const uint4 t4 = *reinterpret_cast<const uint4*>(mj + b * kMaskWords); m[0] = t4.x; m[1] = t4.y; m[2] = t4.z; if (W == 4) m[W - 1] = t4.w; }
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    int n = 0;
#pragma unroll
                    for (int w = 0; w < W; w++) n += __popc(ms[i][w] & m[w]);
                    acc[i] += (uint64_t)n << (16 * b);
                }
            }
        }
    }
    site_finish(a, orow, c, v, lane, buf, sc, acc0, acc1, fwd, rev, dp, sampled);
}

__global__ void __launch_bounds__(kTensorWarps * 32, 6) tensor_kernel(const TensorArgs a) {
    __shared__ __align__(16) int16_t s_out[kTensorWarps][NC_SNP_SITE_STRIDE];
    __shared__ RunList s_list;
    __shared__ int32_t s_nbr[kRunNbr];
    __shared__ __align__(16) uint32_t s_mask[kMaskNbr * 4 * kMaskWords];     // M[neighbour][code][word]
    __shared__ int32_t s_v[kRunSlots], s_c[kRunSlots];
    __shared__ uint8_t s_nbr_rc[kRunNbr], s_rcv[kRunSlots];                   // reference codes of the staged neighbours / of the run's candidates
    __shared__ __align__(4) int8_t s_colj[kTensorWarps][48 + 2 * 48];         // per warp: column -> chosen neighbour (flag bytes, then int16 indices)
    __shared__ int32_t s_wcnt[kTensorWarps];
    __shared__ int64_t s_ilo, s_ihi;
    __shared__ int32_t s_pmin, s_pmax, s_nb_lo, s_nb_hi, s_fast;
    const uint32_t full = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    int16_t* buf = s_out[wib];

    for (int64_t s0 = (int64_t)blockIdx.x * kRunSlots; s0 < a.n_slots; s0 += (int64_t)gridDim.x * kRunSlots) {
        const int nslots = (int)min((int64_t)kRunSlots, a.n_slots - s0);
        __syncthreads();                                               // previous run's shared state no longer in use
        // ---- the run's sites and its position span
        if (wib == 0) {
            int32_t v = 0, c = 0, lo = INT32_MAX, hi = INT32_MIN;
            if (lane < nslots) {
                const int64_t s = s0 + lane;
                c = slot_chunk(a.chunk_off, a.n_chunks, s);
                v = __ldg(a.cand_pos + a.chunk_lo[c] + (s - a.chunk_off[c]));
                lo = hi = v - 1;
            }
            s_v[lane] = v; s_c[lane] = c;
            s_rcv[lane] = lane < nslots ? (uint8_t)ref_code_of(__ldg(a.ref + ((int64_t)(v - 1) - a.ref_start))) : (uint8_t)4;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) { lo = min(lo, __shfl_xor_sync(full, lo, d)); hi = max(hi, __shfl_xor_sync(full, hi, d)); }
            if (lane == 0) {
                s_pmin = lo; s_pmax = hi;
                s_fast = ((int64_t)hi - lo <= 65536 && a.maxcov <= 255) ? 1 : 0;
            }
        }
        __syncthreads();
        const int32_t pmin = s_pmin, pmax = s_pmax;
        {
            // ---- four searches, one per warp: the BAM-index window of reads that can overlap [pmin, pmax] (first read whose
            //      prefix-max end exceeds pmin .. first read starting after pmax) and the neighbours within the search radius
            const int32_t R = c_bins[a.seq][c_nbins[a.seq] - 1].b + 1;
            if (wib == 0) { const int64_t r = warp_bound_i32(a.pos, a.n_reads, pmax, true, lane); if (lane == 0) s_ihi = r; }
            else if (wib == 1) { const int64_t r = warp_bound_i32(a.pmaxend, a.n_reads, pmin, true, lane); if (lane == 0) s_ilo = r; }
            else if (wib == 2) { const int64_t r = warp_bound_i32(a.nbr_pos, a.n_nbr, (int32_t)max((int64_t)pmin + 1 - R, (int64_t)INT32_MIN + 1), false, lane); if (lane == 0) s_nb_lo = (int32_t)r; }
            else { const int64_t r = warp_bound_i32(a.nbr_pos, a.n_nbr, (int32_t)min((int64_t)pmax + 2 + R, (int64_t)INT32_MAX), false, lane); if (lane == 0) s_nb_hi = (int32_t)r; }
        }
        __syncthreads();
        int cnt = 0;
        if (s_fast) {
            // ---- ordered compaction of the admitted reads overlapping [pmin, pmax]
            const int64_t ilo = s_ilo, ihi = s_ihi;
            for (int64_t base = ilo; base < ihi; base += kTensorWarps * 32) {
                const int64_t i = base + tid;
                bool keep = false;
                int32_t rp = 0, re = 0; uint32_t f = 0;
                if (i < ihi) {
                    rp = __ldg(a.pos + i); re = __ldg(a.end + i); f = __ldg(a.flag + i);
                    keep = (f & a.flag_filter) == 0 && rp <= pmax && re > pmin;
                }
                const uint32_t bm = __ballot_sync(full, keep);
                if (lane == 0) s_wcnt[wib] = __popc(bm);
                __syncthreads();
                int off = cnt, tot = 0;
#pragma unroll
                for (int w = 0; w < kTensorWarps; w++) { const int n = s_wcnt[w]; if (w < wib) off += n; tot += n; }
                const int slot = off + __popc(bm & ((1u << lane) - 1u));
                if (keep && slot < kRunList) {
                    s_list.rp[slot] = rp; s_list.re[slot] = re; s_list.rev[slot] = (uint8_t)((f >> 4) & 1u);
                    s_list.rowoff[slot] = __ldg(a.rowoff + i);
                    s_list.nf[slot] = __ldg(a.nfirst + i); s_list.nl[slot] = __ldg(a.nlen + i); s_list.noff[slot] = __ldg(a.noff + i);
                }
                cnt += tot;
                __syncthreads();                                       // s_wcnt reused by the next round
            }
        }
        const bool fast = s_fast && cnt <= kRunList;
        const int32_t nb_lo = s_nb_lo, nb_n = s_nb_hi - s_nb_lo;
        const bool nbr_sm = nb_n <= kRunNbr;
        if (fast && nbr_sm)
            for (int i = tid; i < nb_n; i += kTensorWarps * 32) {
                const int32_t np = __ldg(a.nbr_pos + nb_lo + i);
                s_nbr[i] = np;
                s_nbr_rc[i] = (uint8_t)ref_code_of(__ldg(a.ref + ((int64_t)np - 1 - a.ref_start)));
            }
        const bool masks = fast && cnt <= 32 * kMaskWords && nb_n <= kMaskNbr;
        if (masks) {
            // ---- M[j][b]: lanes as reads, one ballot per code; a warp takes a contiguous quarter of the neighbours
            const int per = (nb_n + kTensorWarps - 1) / kTensorWarps;
            const int j_end = min(nb_n, (wib + 1) * per);
            for (int jr = wib * per; jr < j_end; jr++) {
#pragma unroll
                for (int w = 0; w < kMaskWords; w++) {
                    if (32 * w >= cnt) break;                          // warp-uniform: words past the list are never looked at
                    uint32_t code = 15;
                    const int e = 32 * w + lane;
                    if (e < cnt) {
                        const uint32_t rel = (uint32_t)(nb_lo + jr - s_list.nf[e]);
                        if (rel < (uint32_t)s_list.nl[e]) { const uint32_t by = __ldg(a.nrows + s_list.noff[e] + (rel >> 1)); code = (rel & 1u) ? (by >> 4) : (by & 15u); }
                    }
                    const uint32_t m0 = __ballot_sync(full, code == 0), m1 = __ballot_sync(full, code == 1);
                    const uint32_t m2 = __ballot_sync(full, code == 2), m3 = __ballot_sync(full, code == 3);
                    if (lane < 4) s_mask[(jr * 4 + lane) * kMaskWords + w] = lane == 0 ? m0 : lane == 1 ? m1 : lane == 2 ? m2 : m3;
                }
            }
        }
        __syncthreads();
        // ---- a warp per site; with the neighbours staged, the bin boundaries are carried from one site of the warp to the next
        int32_t st_lo = 0, st_hi = 0;
        int prev_c = -1;                                               // a new chunk may lie anywhere (and clips differently): search again
        for (int k = wib; k < nslots; k += kTensorWarps) {
            const int64_t s = s0 + k;
            int64_t orow = s;
            if (a.outidx) {
                orow = __ldg(a.outidx + s);
                if (__ldg(a.outidx + s + 1) == orow) continue;         // dropped by min_nbr_sites
            }
            const int c = s_c[k];
            const int32_t v = s_v[k];
            if (!fast) { tensor_site_generic(a, orow, c, v, lane, buf); continue; }
            SiteCols sc;
            if (nbr_sm) { sc = site_columns_run(a, s_nbr, s_nbr_rc, nb_n, nb_lo, c, v, s_rcv[k], lane, c != prev_c, st_lo, st_hi, s_colj[wib]); prev_c = c; }
            else sc = site_columns(a, a.nbr_pos + nb_lo, nb_n, nb_lo, c, v, lane);
            if (masks) {
                switch ((cnt + 31) >> 5) {
                    case 0: case 1: tensor_site_masks<1>(a, orow, c, v, lane, buf, s_list, cnt, nb_lo, sc, s_mask); break;
                    case 2: tensor_site_masks<2>(a, orow, c, v, lane, buf, s_list, cnt, nb_lo, sc, s_mask); break;
                    case 3: tensor_site_masks<3>(a, orow, c, v, lane, buf, s_list, cnt, nb_lo, sc, s_mask); break;
                    default: tensor_site_masks<4>(a, orow, c, v, lane, buf, s_list, cnt, nb_lo, sc, s_mask); break;
                }
            }
            else tensor_site_run(a, orow, c, v, lane, buf, s_list, cnt, sc);
        }
    }
}

// mean(current_depth) per chunk (generate_SNP_pileups.py:274): exact integer sum / count in float64.
__global__ void chunk_depth_kernel(const unsigned long long* __restrict__ sum, const unsigned long long* __restrict__ cnt,
                                   int32_t n_chunks, double* __restrict__ depth, int64_t* __restrict__ count) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    depth[c] = cnt[c] ? (double)sum[c] / (double)cnt[c] : 0.0;
    count[c] = (int64_t)cnt[c];
}

}  // namespace nc
