// Device-side BAM input: BGZF inflate and BAM record decoding on the GPU, so that a file-backed run moves
// the COMPRESSED bytes over PCIe and the host's zlib (0.94 GB/s of BAM on 16 threads, the bound of round 1's --from-bam figure)
// leaves the path.  Replaces what the reference gets from pysam / htslib when it opens the alignment file in every call
// (nanocaller_src/generate_SNP_pileups.py:134-156, generate_indel_pileups.py:147-185).
//
//   bgzf_inflate_kernel   one WARP per BGZF block (blocks are independent raw-DEFLATE members of <= 64 KB, RFC 1951): stored, fixed
//                         and dynamic Huffman blocks.  Lane 0 owns the bit stream (canonical-code tables as in Mark Adler's puff.c plus
//                         a 9-bit lookup table, all in shared memory) and decodes 32 symbols at a time; the warp writes them out:
//                         literals in one step, LZ77 matches as cooperative copies.  (A first version with one thread per block ran at
//                         4.5 GB/s: every byte of a match was a dependent store -> load round trip through L2.)
//   bam_walk_kernel       the record chain (every record starts with its own size, SAM spec 4.2): one thread follows it and writes
//                         the record offsets; everything after that is parallel over records.
//   bam_fields_kernel     thread per record: core fields, HP / PS aux tags, the CG:B,I long-CIGAR convention.
//   bam_copy_kernel       warp per record: CIGAR words and 4-bit bases into the staging arrays nc_stage_reads would have uploaded.
#pragma once
#include "nc_common.cuh"

namespace nc {

struct BgzfBlock { int64_t in_off; int32_t in_len; int32_t out_len; int64_t out_off; };     // = NcBgzfBlock of the C header

constexpr int kInflWarps = 4;                 // warps (= BGZF blocks in flight) per CTA
constexpr int kLutBits = 9;                   // primary lookup: codes of up to 9 bits resolve with one shared-memory read
// per warp, in shared memory: canonical-code tables (count per length, symbols by (length, value)) for the literal/length and the
// distance code, code lengths while a dynamic header is read, the two lookup tables, and a queue of 32 decoded symbols
struct InflWarp {
    uint16_t lcnt[16], lsym[288], dcnt[16], dsym[32];
    uint16_t llut[1 << kLutBits], dlut[1 << kLutBits];       // (symbol << 4) | code length, 0 = longer code: walk the canonical table
    uint8_t lens[320];
    uint16_t q_len[32], q_dist[32];                          // queue: dist == 0 -> literal byte in q_len
};

__constant__ uint16_t c_len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t c_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t c_dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t c_dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t c_clen_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

struct BitReader {
    const uint8_t* p; const uint8_t* end;
    uint64_t buf; int cnt;
    __device__ __forceinline__ void refill() {            // at least 32 valid bits afterwards (zeros past the end of the payload)
        while (cnt <= 32) {
            uint32_t w = 0;
            if (p + 4 <= end) { w = (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16) | ((uint32_t)__ldg(p + 3) << 24); p += 4; buf |= (uint64_t)w << cnt; cnt += 32; }
            else if (p < end) { buf |= (uint64_t)__ldg(p) << cnt; p++; cnt += 8; }
            else { cnt += 32; }                           // past the end: zero bits (a well-formed stream never consumes them)
        }
    }
    __device__ __forceinline__ uint32_t bits(int n) { const uint32_t v = (uint32_t)(buf & ((1ull << n) - 1ull)); buf >>= n; cnt -= n; return v; }
};

// canonical Huffman table from code lengths len[0..n): cnt[l] = codes of length l, sym = symbols ordered by (length, value).
// Returns 0 for a complete code, > 0 incomplete, < 0 over-subscribed (the construction of Mark Adler's puff.c).  One thread.
template <class LenAt>
__device__ __forceinline__ int huff_build(uint16_t* cnt, uint16_t* sym, LenAt len_at, int n) {
    for (int l = 0; l <= 15; l++) cnt[l] = 0;
    for (int s = 0; s < n; s++) cnt[len_at(s)]++;
    if (cnt[0] == n) return 0;
    int left = 1;
    for (int l = 1; l <= 15; l++) { left <<= 1; left -= cnt[l]; if (left < 0) return left; }
    uint16_t offs[16];
    offs[1] = 0;
    for (int l = 1; l < 15; l++) offs[l + 1] = offs[l] + cnt[l];
    for (int s = 0; s < n; s++) { const int l = len_at(s); if (l) sym[offs[l]++] = (uint16_t)s; }
    return left;
}
// lookup table of the codes of up to kLutBits bits (DEFLATE packs codes most significant bit first into an LSB-first stream, so
// the table is indexed by the bit-reversed code); all lanes of the warp fill it
__device__ __forceinline__ void huff_lut(const uint16_t* cnt, const uint16_t* sym, uint16_t* lut, int lane) {
    for (int i = lane; i < (1 << kLutBits); i += 32) lut[i] = 0;
    __syncwarp();
    int total = 0;
    for (int l = 1; l <= kLutBits; l++) total += cnt[l];
    for (int i = lane; i < total; i += 32) {
        int l = 1, first = 0, index = 0;                  // canonical code of the i-th symbol in (length, value) order
        while (i >= index + cnt[l]) { index += cnt[l]; first = (first + cnt[l]) << 1; l++; }
        const uint32_t code = (uint32_t)(first + (i - index));
        const uint32_t rev = __brev(code) >> (32 - l);
        const uint16_t e = (uint16_t)((sym[i] << 4) | l);
        for (uint32_t k = rev; k < (1u << kLutBits); k += 1u << l) lut[k] = e;
    }
    __syncwarp();
}
__device__ __forceinline__ int huff_decode(BitReader& br, const uint16_t* cnt, const uint16_t* sym, const uint16_t* lut) {
    const uint32_t e = lut[(uint32_t)br.buf & ((1u << kLutBits) - 1u)];
    if (e) { const int l = (int)(e & 15u); br.buf >>= l; br.cnt -= l; return (int)(e >> 4); }
    int code = 0, first = 0, index = 0;
#pragma unroll 1
    for (int l = 1; l <= 15; l++) {
        code |= (int)(br.buf & 1ull); br.buf >>= 1; br.cnt--;
        const int count = cnt[l];
        if (code - count < first) return sym[index + (code - first)];
        index += count; first += count; first <<= 1; code <<= 1;
    }
    return -1;
}

// One WARP per BGZF block: lane 0 reads the bit stream and decodes up to 32 symbols into a queue, then all lanes write them out —
// literals in one parallel step, every match as a cooperative copy (back-references with distance < length repeat the pattern).
// err[0] = number of blocks that failed to decode.
__global__ void __launch_bounds__(kInflWarps * 32) bgzf_inflate_kernel(const uint8_t* __restrict__ in, const BgzfBlock* __restrict__ blocks, int64_t n_blocks,
                                                                       uint8_t* __restrict__ out, int* __restrict__ err) {
    __shared__ InflWarp s_w[kInflWarps];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    InflWarp& W = s_w[wib];
    const uint32_t full = 0xffffffffu;
    const int64_t b = (int64_t)blockIdx.x * kInflWarps + wib;
    if (b >= n_blocks) return;
    const BgzfBlock bk = blocks[b];
    uint8_t* o = out + bk.out_off;
    const int out_len = bk.out_len;
    BitReader br;
    br.p = in + bk.in_off; br.end = br.p + bk.in_len; br.buf = 0; br.cnt = 0;
    int op = 0;                                            // warp-uniform output position
    int state = 0;                                         // lane 0: 0 = expect a block header, 1 = inside a Huffman block; broadcast below
    bool bad = false, last = false;
    for (;;) {
        // ---- lane 0: headers and up to 32 symbols
        int nq = 0, stored_len = -1;
        const uint8_t* stored_src = nullptr;
        int need_tables = 0;                               // 1: fixed code, 2: dynamic code lengths are in W.lens
        int nlen = 0, ndist = 0;
        if (lane == 0 && !bad) {
            if (state == 0) {
                if (last) state = 3;                       // done
                else {
                    br.refill();
                    last = br.bits(1) != 0;
                    const uint32_t type = br.bits(2);
                    if (type == 0) {
                        br.bits(br.cnt & 7);               // to the byte boundary
                        br.refill();
                        const uint32_t len = br.bits(16), nl = br.bits(16);
                        const uint8_t* src = br.p - (br.cnt >> 3);     // bytes still sitting in the bit buffer are given back
                        if ((len ^ 0xFFFFu) != nl || src + len > br.end || op + (int)len > out_len) bad = true;
                        else { stored_len = (int)len; stored_src = src; br.p = src + len; br.buf = 0; br.cnt = 0; }
                    } else if (type == 1) { need_tables = 1; state = 1; }
                    else if (type == 2) {
                        br.refill();
                        nlen = (int)br.bits(5) + 257; ndist = (int)br.bits(5) + 1;
                        const int ncode = (int)br.bits(4) + 4;
                        if (nlen > 286 || ndist > 30) bad = true;
                        else {
                            for (int i = 0; i < 19; i++) W.lens[i] = 0;
                            for (int i = 0; i < ncode; i++) { br.refill(); W.lens[c_clen_order[i]] = (uint8_t)br.bits(3); }
                            // the code-length code uses the distance table's slots while the lengths are read (no lookup table: 19 symbols)
                            if (huff_build(W.dcnt, W.dsym, [&](int s) { return (int)W.lens[s]; }, 19) != 0) bad = true;
                            int idx = 0;
                            while (idx < nlen + ndist && !bad) {
                                br.refill();
                                int code = 0, first = 0, index = 0, sy = -1;
                                for (int l = 1; l <= 7; l++) {
                                    code |= (int)(br.buf & 1ull); br.buf >>= 1; br.cnt--;
                                    const int count = W.dcnt[l];
                                    if (code - count < first) { sy = W.dsym[index + (code - first)]; break; }
                                    index += count; first += count; first <<= 1; code <<= 1;
                                }
                                if (sy < 0) { bad = true; break; }
                                if (sy < 16) { W.lens[idx++] = (uint8_t)sy; continue; }
                                int prev = 0, rep;
                                if (sy == 16) { if (idx == 0) { bad = true; break; } prev = W.lens[idx - 1]; rep = 3 + (int)br.bits(2); }
                                else if (sy == 17) rep = 3 + (int)br.bits(3);
                                else rep = 11 + (int)br.bits(7);
                                if (idx + rep > nlen + ndist) { bad = true; break; }
                                while (rep--) W.lens[idx++] = (uint8_t)prev;
                            }
                            if (!bad && W.lens[256] == 0) bad = true;
                            if (!bad) { need_tables = 2; state = 1; }
                        }
                    } else bad = true;
                }
            }
        }
        // ---- warp: broadcast what lane 0 found
        bad = __shfl_sync(full, (int)bad, 0) != 0;
        if (bad) break;
        state = __shfl_sync(full, state, 0);
        if (state == 3) break;
        stored_len = __shfl_sync(full, stored_len, 0);
        if (stored_len >= 0) {
            const uint64_t sp = __shfl_sync(full, (uint64_t)(uintptr_t)stored_src, 0);
            const uint8_t* src = reinterpret_cast<const uint8_t*>((uintptr_t)sp);
            for (int i = lane; i < stored_len; i += 32) o[op + i] = __ldg(src + i);
            op += stored_len;
            __syncwarp();
            continue;
        }
        need_tables = __shfl_sync(full, need_tables, 0);
        if (need_tables) {
            if (lane == 0) {
                if (need_tables == 1) {
                    huff_build(W.lcnt, W.lsym, [](int s) { return s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8; }, 288);
                    huff_build(W.dcnt, W.dsym, [](int) { return 5; }, 30);
                } else {
                    const int el = huff_build(W.lcnt, W.lsym, [&](int s) { return (int)W.lens[s]; }, nlen);
                    if (el < 0 || (el > 0 && nlen - W.lcnt[0] != 1)) bad = true;
                    const int ed = huff_build(W.dcnt, W.dsym, [&](int s) { return (int)W.lens[nlen + s]; }, ndist);
                    if (ed < 0 || (ed > 0 && ndist - W.dcnt[0] != 1)) bad = true;
                }
            }
            bad = __shfl_sync(full, (int)bad, 0) != 0;
            if (bad) break;
            __syncwarp();
            huff_lut(W.lcnt, W.lsym, W.llut, lane);
            huff_lut(W.dcnt, W.dsym, W.dlut, lane);
        }
        // ---- lane 0: decode up to 32 symbols of the current Huffman block
        if (lane == 0) {
            int room = out_len - op;
            while (nq < 32) {
                br.refill();
                int sy = huff_decode(br, W.lcnt, W.lsym, W.llut);
                if (sy < 0) { bad = true; break; }
                if (sy < 256) { if (room < 1) { bad = true; break; } W.q_len[nq] = (uint16_t)sy; W.q_dist[nq] = 0; nq++; room--; continue; }
                if (sy == 256) { state = 0; break; }
                sy -= 257;
                if (sy >= 29) { bad = true; break; }
                br.refill();
                const int len = c_len_base[sy] + (int)br.bits(c_len_extra[sy]);
                const int ds = huff_decode(br, W.dcnt, W.dsym, W.dlut);
                if (ds < 0 || ds >= 30) { bad = true; break; }
                br.refill();
                const int dist = c_dist_base[ds] + (int)br.bits(c_dist_extra[ds]);
                if (len > room) { bad = true; break; }
                W.q_len[nq] = (uint16_t)len; W.q_dist[nq] = (uint16_t)dist; nq++; room -= len;
            }
        }
        bad = __shfl_sync(full, (int)bad, 0) != 0;
        if (bad) break;
        state = __shfl_sync(full, state, 0);
        nq = __shfl_sync(full, nq, 0);
        __syncwarp();
        // ---- warp: write the queue.  Output offsets by a prefix sum over the symbols' lengths; literals first (independent bytes),
        //      then the matches in stream order (a match may read what an earlier symbol of the same batch wrote)
        const int my_dist = lane < nq ? (int)W.q_dist[lane] : 0;
        const int my_len = lane < nq ? (my_dist ? (int)W.q_len[lane] : 1) : 0;
        int incl = my_len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(full, incl, d); if (lane >= d) incl += v; }
        const int my_off = op + incl - my_len;
        if (lane < nq && my_dist == 0) o[my_off] = (uint8_t)W.q_len[lane];
        uint32_t mm = __ballot_sync(full, lane < nq && my_dist != 0);
        __syncwarp();
        while (mm) {
            const int k = __ffs(mm) - 1;
            mm &= mm - 1;
            const int len = __shfl_sync(full, my_len, k), dist = __shfl_sync(full, my_dist, k), at = __shfl_sync(full, my_off, k);
            if (dist > at) { bad = true; break; }             // warp-uniform
            const uint8_t* src = o + at - dist;
            if (dist >= len) { for (int i = lane; i < len; i += 32) o[at + i] = __ldcg(src + i); }
            else { for (int i = lane; i < len; i += 32) o[at + i] = __ldcg(src + i % dist); }
            __syncwarp();
        }
        if (bad) break;
        op += __shfl_sync(full, incl, 31);
    }
    if (lane == 0 && (bad || op != out_len)) atomicAdd(err, 1);
}

// ---- BAM records -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_u32(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
__device__ __forceinline__ uint32_t ld_u16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

// out[0] = number of records, out[1] = 1 if the chain ended exactly at `end`, rec_off[i] = offset of record i's block_size field
__global__ void bam_walk_kernel(const uint8_t* __restrict__ d, int64_t first, int64_t end, int64_t* __restrict__ rec_off, int64_t cap, int64_t* __restrict__ out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    int64_t off = first, n = 0;
    while (off + 4 <= end) {
        const int64_t bs = (int32_t)ld_u32(d + off);
        if (bs < 32 || off + 4 + bs > end) break;
        if (n < cap) rec_off[n] = off;
        n++;
        off += 4 + bs;
    }
    out[0] = n; out[1] = off == end ? 1 : 0;
}

// Parallel version of the walk when record starts are known (the BAI index lists the virtual offset of a record for every 16 kb
// window and every bin chunk): walker w follows the chain from start[w] and must land exactly on start[w + 1] (`end` for the last).
// FILL = false: cnt[w] = records of the segment; FILL = true: rec_off[off[w] ...] = their offsets.  err[3] counts segments that do not
// land on the next start (the caller then falls back to the single walker).
template <bool FILL>
__global__ void bam_walk_seg_kernel(const uint8_t* __restrict__ d, const int64_t* __restrict__ start, int64_t n_seg, int64_t end,
                                    int32_t* __restrict__ cnt, const int64_t* __restrict__ off, int64_t* __restrict__ rec_off, int* __restrict__ err) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_seg) return;
    int64_t p = start[w];
    const int64_t stop = w + 1 < n_seg ? start[w + 1] : end;
    int64_t n = 0, o = FILL ? off[w] : 0;
    while (p < stop) {
        if (p + 4 > end) break;
        const int64_t bs = (int32_t)ld_u32(d + p);
        if (bs < 32 || p + 4 + bs > end) break;
        if (FILL) rec_off[o + n] = p;
        n++;
        p += 4 + bs;
    }
    if (!FILL) { cnt[w] = (int32_t)n; if (p != stop) atomicAdd(err + 3, 1); }
}

struct BamFields {
    int32_t* rid; int32_t* pos; uint16_t* flag; int32_t* lseq; int32_t* ncig; int32_t* nseq;     // nseq = (l_seq + 1) / 2
    int64_t* cig_src; int64_t* seq_src; int8_t* hp; int32_t* ps;
};
__device__ __forceinline__ int aux_fixed_size(uint8_t ty) {
    switch (ty) { case 'A': case 'c': case 'C': return 1; case 's': case 'S': return 2; case 'i': case 'I': case 'f': return 4; }
    return -1;
}
__device__ __forceinline__ bool aux_int_value(uint8_t ty, const uint8_t* p, int32_t* v) {
    switch (ty) {
        case 'c': *v = (int8_t)p[0]; return true;
        case 'C': *v = p[0]; return true;
        case 's': *v = (int16_t)ld_u16(p); return true;
        case 'S': *v = (int32_t)ld_u16(p); return true;
        case 'i': case 'I': *v = (int32_t)ld_u32(p); return true;
    }
    return false;
}
// err[1] = records whose fields exceed their size
__global__ void bam_fields_kernel(const uint8_t* __restrict__ d, const int64_t* __restrict__ rec_off, int64_t n, BamFields f, int* __restrict__ err) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* base = d + rec_off[i];
    const int64_t bs = (int32_t)ld_u32(base);
    const uint8_t* r = base + 4;
    const int32_t l_name = r[8], n_cig = (int32_t)ld_u16(r + 12), l_seq = (int32_t)ld_u32(r + 16);
    f.rid[i] = (int32_t)ld_u32(r); f.pos[i] = (int32_t)ld_u32(r + 4); f.flag[i] = (uint16_t)ld_u16(r + 14);
    int64_t cig_at = rec_off[i] + 4 + 32 + l_name, ncig = n_cig;
    const int64_t seq_at = cig_at + 4ll * n_cig;
    int8_t hp = 0; int32_t ps = 0;
    if (l_seq < 0 || 32ll + l_name + 4ll * n_cig + (l_seq + 1) / 2 + l_seq > bs) {
        atomicAdd(err + 1, 1);
        f.lseq[i] = 0; f.ncig[i] = 0; f.nseq[i] = 0; f.cig_src[i] = cig_at; f.seq_src[i] = seq_at; f.hp[i] = 0; f.ps[i] = 0;
        return;
    }
    // long-CIGAR placeholder `<l_seq>S<ref_len>N` (SAM spec 4.2.2): the real operations are in CG:B,I
    bool placeholder = false;
    if (n_cig == 2) {
        const uint32_t c0 = ld_u32(d + cig_at), c1 = ld_u32(d + cig_at + 4);
        placeholder = (c0 & 15u) == 4u && (int64_t)(c0 >> 4) == (int64_t)l_seq && (c1 & 15u) == 3u;
    }
    const uint8_t* p = d + seq_at + (l_seq + 1) / 2 + l_seq;
    const uint8_t* end = r + bs;
    while (p + 3 <= end) {
        const uint8_t t0 = p[0], t1 = p[1], ty = p[2];
        p += 3;
        const int fs = aux_fixed_size(ty);
        if (fs > 0) {
            if (p + fs > end) break;
            int32_t v;
            if (t0 == 'H' && t1 == 'P' && aux_int_value(ty, p, &v)) hp = (int8_t)v;
            if (t0 == 'P' && t1 == 'S' && aux_int_value(ty, p, &v)) ps = v;
            p += fs;
        } else if (ty == 'Z' || ty == 'H') {
            while (p < end && *p) p++;
            p++;
        } else if (ty == 'B') {
            if (p + 5 > end) break;
            const int es = aux_fixed_size(p[0]);
            const int64_t cnt = (int32_t)ld_u32(p + 1);
            if (es < 0 || cnt < 0 || p + 5 + es * cnt > end) break;
            if (placeholder && t0 == 'C' && t1 == 'G' && p[0] == 'I') { cig_at = (p + 5) - d; ncig = cnt; }
            p += 5 + es * cnt;
        } else break;
    }
    f.lseq[i] = l_seq; f.ncig[i] = (int32_t)ncig; f.nseq[i] = (l_seq + 1) / 2; f.cig_src[i] = cig_at; f.seq_src[i] = seq_at; f.hp[i] = hp; f.ps[i] = ps;
}

// warp per record of the contig (records [first, first + n) of the field arrays): payload copies into the staging arrays
__global__ void __launch_bounds__(128) bam_copy_kernel(const uint8_t* __restrict__ d, int64_t first, int64_t n, BamFields f,
                                                       const int64_t* __restrict__ cigar_off, const int64_t* __restrict__ seq_off,
                                                       uint32_t* __restrict__ cigar, uint8_t* __restrict__ seq4) {
    const int lane = threadIdx.x & 31;
    const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (k >= n) return;
    const int64_t i = first + k;
    const uint8_t* cs = d + f.cig_src[i];
    uint32_t* co = cigar + cigar_off[k];
    const int nc = f.ncig[i];
    for (int j = lane; j < nc; j += 32) co[j] = ld_u32(cs + 4ll * j);
    const uint8_t* ss = d + f.seq_src[i];
    uint8_t* so = seq4 + seq_off[k];
    const int nb = f.nseq[i];
    for (int j = lane; j < nb; j += 32) so[j] = ss[j];
}

// err[2] = positions that decrease inside the contig (the file is not coordinate-sorted)
__global__ void bam_sorted_kernel(const int32_t* __restrict__ pos, int64_t n, int* __restrict__ err) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 < n && pos[i + 1] < pos[i]) atomicAdd(err + 2, 1);
}

}  // namespace nc
