// libnc_bamio — native BAM reader into the staging arrays of libnanocaller_b200.
//
// Replaces what the reference gets from pysam/htslib when it opens the alignment file in every call
// (nanocaller_src/generate_SNP_pileups.py:134-156, generate_indel_pileups.py:147-185): BGZF inflate and BAM
// record decoding.  The device path wants BAM's own encodings (CIGAR words, 4-bit bases), so "decoding" is
// mostly copying: per contig one set of arrays pos / flag / cigar_off / cigar / seq_off / l_seq / seq4 plus the
// integer HP / PS tags of the indel path.
//
//   1. the BGZF block table is read from the block headers (BSIZE) and trailers (ISIZE);
//   2. the blocks are inflated in parallel (raw deflate, zlib) into one contiguous buffer;
//   3. one pointer-chasing pass finds the record boundaries and per-contig totals;
//   4. records are copied into the caller's arrays in parallel.
// C ABI, two calls: nc_bam_open parses and reports sizes, nc_bam_fill writes into caller-owned buffers (numpy or
// pinned memory), nc_bam_close frees.  No htslib; zlib and pthreads only.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "nanocaller_b200_io.h"

namespace {

struct Contig {
    std::string name;
    int32_t length = 0;
    int64_t n_reads = 0, n_cigar = 0, n_seq = 0;
    int64_t first_rec = 0;          // index into rec_off of the contig's first record (records are coordinate sorted)
};

}  // namespace

// Uninitialised byte buffer: the inflate workers are the first to touch the pages, so the page faults are spread over the
// threads instead of being paid by one zero-filling constructor.
struct RawBuf {
    uint8_t* p = nullptr;
    size_t n = 0;
    ~RawBuf() { free(p); }
    bool alloc(size_t bytes) { free(p); p = (uint8_t*)malloc(bytes ? bytes : 1); n = p ? bytes : 0; return p != nullptr; }
    const uint8_t* data() const { return p; }
    uint8_t* data() { return p; }
    size_t size() const { return n; }
};

struct nc_bam {
    RawBuf data;                    // inflated BAM stream
    std::string text, err;
    std::vector<Contig> contigs;
    std::vector<int64_t> rec_off;   // byte offset of every mapped record's block_size field, grouped by contig
    std::vector<int32_t> rec_rid;
    bool sorted = true;
};

namespace {

template <class T> T rd(const uint8_t* p) { T v; memcpy(&v, p, sizeof(T)); return v; }

int fail(nc_bam* b, const char* msg) { b->err = msg; return NC_IO_EFORMAT; }

template <class F> void parallel_for(int64_t n, int threads, F f) {
    threads = (int)std::max<int64_t>(1, std::min<int64_t>(threads, n));
    if (threads == 1) { for (int64_t i = 0; i < n; i++) f(i); return; }
    std::atomic<int64_t> next{0};
    std::vector<std::thread> pool;
    const int64_t grain = std::max<int64_t>(1, n / (threads * 16));
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&]() {
            for (;;) {
                const int64_t a = next.fetch_add(grain);
                if (a >= n) return;
                const int64_t e = std::min(n, a + grain);
                for (int64_t i = a; i < e; i++) f(i);
            }
        });
    for (auto& th : pool) th.join();
}

struct Blk { size_t start, off, csize; uint32_t isize; size_t out; };     // block start, payload offset / size, inflated size, output offset

// Appends the BGZF blocks of file[cbeg, cend) to `blocks` (cend = file size: all of them); `total` = running inflated size.
int block_table(nc_bam* b, const uint8_t* file, size_t n, size_t cbeg, size_t cend, std::vector<Blk>& blocks, size_t& total) {
    size_t off = cbeg;
    while (off < cend) {
        if (off + 18 > n) return fail(b, "truncated or malformed BGZF block");
        const uint8_t* p = file + off;
        if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return fail(b, "not a BGZF stream (gzip member without the BC extra field)");
        const uint16_t xlen = rd<uint16_t>(p + 10);
        size_t x = 12, xend = 12 + xlen;
        int bsize = -1;
        while (x + 4 <= xend && off + x + 4 <= n) {
            const uint16_t slen = rd<uint16_t>(p + x + 2);
            if (p[x] == 'B' && p[x + 1] == 'C' && slen == 2) bsize = rd<uint16_t>(p + x + 4);
            x += 4 + slen;
        }
        if (bsize < 0 || off + (size_t)bsize + 1 > n) return fail(b, "truncated or malformed BGZF block");
        const size_t blen = (size_t)bsize + 1;
        if (blen < xend + 8) return fail(b, "truncated or malformed BGZF block");
        const uint32_t isize = rd<uint32_t>(p + blen - 4);
        blocks.push_back({off, off + xend, blen - xend - 8, isize, total});
        total += isize;
        off += blen;
    }
    return NC_IO_OK;
}

int inflate_blocks(nc_bam* b, const uint8_t* file, const std::vector<Blk>& blocks, size_t total, int threads) {
    if (!b->data.alloc(total)) return fail(b, "out of memory");
    std::atomic<int> bad{0};
    parallel_for((int64_t)blocks.size(), threads, [&](int64_t i) {
        const Blk& k = blocks[i];
        if (k.isize == 0) return;
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, -15) != Z_OK) { bad = 1; return; }
        zs.next_in = const_cast<Bytef*>(file + k.off);
        zs.avail_in = (uInt)k.csize;
        zs.next_out = b->data.data() + k.out;
        zs.avail_out = k.isize;
        const int rc = inflate(&zs, Z_FINISH);
        if (rc != Z_STREAM_END || zs.total_out != k.isize) bad = 1;
        inflateEnd(&zs);
    });
    if (bad) return fail(b, "inflate failed (corrupt BGZF block)");
    return NC_IO_OK;
}

int inflate_all(nc_bam* b, const uint8_t* file, size_t n, int threads) {
    std::vector<Blk> blocks;
    size_t total = 0;
    const int rc = block_table(b, file, n, 0, n, blocks, total);
    if (rc) return rc;
    return inflate_blocks(b, file, blocks, total, threads);
}

// BAM header at d[0, n): text, reference names and lengths.  Returns the offset of the first record, 0 when more bytes are
// needed, or -1 on a format error.
int64_t parse_header(nc_bam* b, const uint8_t* d, size_t n) {
    if (n < 12) return 0;
    if (memcmp(d, "BAM\1", 4) != 0) { fail(b, "not a BAM file (bad magic)"); return -1; }
    const int32_t l_text = rd<int32_t>(d + 4);
    if (l_text < 0) { fail(b, "truncated BAM header"); return -1; }
    if (8 + (size_t)l_text + 4 > n) return 0;
    size_t off = 8 + (size_t)l_text;
    const int32_t n_ref = rd<int32_t>(d + off);
    off += 4;
    if (n_ref < 0) { fail(b, "negative reference count"); return -1; }
    std::vector<Contig> refs((size_t)n_ref);
    for (int32_t i = 0; i < n_ref; i++) {
        if (off + 4 > n) return 0;
        const int32_t l_name = rd<int32_t>(d + off);
        if (l_name <= 0) { fail(b, "truncated reference list"); return -1; }
        if (off + 4 + (size_t)l_name + 4 > n) return 0;
        refs[i].name.assign((const char*)d + off + 4, (size_t)l_name - 1);
        refs[i].length = rd<int32_t>(d + off + 4 + l_name);
        off += 8 + (size_t)l_name;
    }
    b->text.assign((const char*)d + 8, (size_t)l_text);
    b->contigs.swap(refs);
    return (int64_t)off;
}

int64_t real_cigar_count(const uint8_t* r, int32_t bs);     // long-CIGAR aware (CG:B,I), defined with real_cigar below

// Walks the records of data[off, end): boundaries and per-contig totals.  only_rid >= 0: stop at the first record of another
// reference (region reads).
int scan_records(nc_bam* b, size_t off, size_t end, int32_t only_rid, std::vector<std::vector<int64_t>>& per) {
    const uint8_t* d = b->data.data();
    const int32_t n_ref = (int32_t)b->contigs.size();
    int32_t last_rid = -1, last_pos = -1;
    while (off + 4 <= end) {
        const int32_t bs = rd<int32_t>(d + off);
        if (bs < 32 || off + 4 + (size_t)bs > end) return only_rid >= 0 ? NC_IO_OK : fail(b, "truncated alignment record");
        const uint8_t* r = d + off + 4;
        const int32_t rid = rd<int32_t>(r), pos = rd<int32_t>(r + 4);
        if (only_rid >= 0 && rid != only_rid) break;
        const uint8_t l_name = r[8];
        const uint16_t n_cig = rd<uint16_t>(r + 12);
        const int32_t l_seq = rd<int32_t>(r + 16);
        if (l_seq < 0 || 32 + (size_t)l_name + 4 * (size_t)n_cig + (size_t)(l_seq + 1) / 2 + (size_t)l_seq > (size_t)bs) return fail(b, "alignment record fields exceed its size");
        if (rid >= 0 && rid < n_ref) {
            if (rid < last_rid || (rid == last_rid && pos < last_pos)) b->sorted = false;
            last_rid = rid; last_pos = pos;
            Contig& c = b->contigs[(size_t)rid];
            c.n_reads++; c.n_cigar += real_cigar_count(r, bs); c.n_seq += (l_seq + 1) / 2;
            per[(size_t)rid].push_back((int64_t)off);
        }
        off += 4 + (size_t)bs;
    }
    if (!b->sorted) return fail(b, "BAM is not coordinate-sorted");
    return NC_IO_OK;
}

void finish_index(nc_bam* b, std::vector<std::vector<int64_t>>& per) {
    for (size_t i = 0; i < b->contigs.size(); i++) {
        b->contigs[i].first_rec = (int64_t)b->rec_off.size();
        b->rec_off.insert(b->rec_off.end(), per[i].begin(), per[i].end());
    }
}

struct MapFile {                       // read-only mapping of a file: the compressed bytes come straight from the page cache
    const uint8_t* p = nullptr;
    size_t n = 0;
    int err = 0;
    explicit MapFile(const char* path) {
        const int fd = open(path, O_RDONLY);
        if (fd < 0) { err = 1; return; }
        struct stat st;
        if (fstat(fd, &st) != 0) { close(fd); err = 1; return; }
        n = (size_t)st.st_size;
        void* m = n ? mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;
        close(fd);
        if (n && m == MAP_FAILED) { err = 2; n = 0; return; }
        p = (const uint8_t*)m;
    }
    ~MapFile() { if (p) munmap((void*)p, n); }
};

// integer value of a fixed-width aux field, or false
bool aux_int(uint8_t type, const uint8_t* p, int32_t* out) {
    switch (type) {
        case 'c': *out = (int8_t)p[0]; return true;
        case 'C': *out = p[0]; return true;
        case 's': *out = rd<int16_t>(p); return true;
        case 'S': *out = rd<uint16_t>(p); return true;
        case 'i': *out = rd<int32_t>(p); return true;
        case 'I': *out = (int32_t)rd<uint32_t>(p); return true;
    }
    return false;
}
int aux_size(uint8_t type) {
    switch (type) {
        case 'A': case 'c': case 'C': return 1;
        case 's': case 'S': return 2;
        case 'i': case 'I': case 'f': return 4;
    }
    return -1;
}
void scan_tags(const uint8_t* p, const uint8_t* end, int8_t* hp, int32_t* ps) {
    *hp = 0; *ps = 0;
    while (p + 3 <= end) {
        const uint8_t t0 = p[0], t1 = p[1], type = p[2];
        p += 3;
        const int fs = aux_size(type);
        if (fs > 0) {
            if (p + fs > end) return;
            int32_t v;
            if (t0 == 'H' && t1 == 'P' && aux_int(type, p, &v)) *hp = (int8_t)v;
            if (t0 == 'P' && t1 == 'S' && aux_int(type, p, &v)) *ps = v;
            p += fs;
        } else if (type == 'Z' || type == 'H') {
            while (p < end && *p) p++;
            p++;
        } else if (type == 'B') {
            if (p + 5 > end) return;
            const int es = aux_size(p[0]);
            const int32_t cnt = rd<int32_t>(p + 1);
            if (es < 0 || cnt < 0) return;
            p += 5 + (size_t)es * (size_t)cnt;
        } else {
            return;
        }
    }
}

// Long CIGARs (SAM spec 4.2.2): a read with more than 65,535 operations — routine for the ultra-long ONT reads the ul_ont presets
// target — stores the placeholder `<l_seq>S<reference length>N` in the CIGAR field and the real operations in a CG:B,I aux
// array; htslib (and so pysam in the reference, generate_SNP_pileups.py:141,156) restores it transparently.  Returns the
// operations a consumer must see: the aux array when the record is such a placeholder, else the CIGAR field itself.
struct CigarView { const uint8_t* p; int64_t n; };
CigarView real_cigar(const uint8_t* r, int32_t bs) {
    const uint8_t l_name = r[8];
    const uint16_t n_cig = rd<uint16_t>(r + 12);
    const int32_t l_seq = rd<int32_t>(r + 16);
    const uint8_t* cg = r + 32 + l_name;
    CigarView v = {cg, n_cig};
    if (n_cig != 2) return v;
    const uint32_t c0 = rd<uint32_t>(cg), c1 = rd<uint32_t>(cg + 4);
    if ((c0 & 15u) != 4u || (int64_t)(c0 >> 4) != (int64_t)l_seq || (c1 & 15u) != 3u) return v;
    const uint8_t* p = cg + 8 + (size_t)(l_seq + 1) / 2 + (size_t)l_seq;
    const uint8_t* end = r + bs;
    while (p + 3 <= end) {
        const uint8_t t0 = p[0], t1 = p[1], type = p[2];
        p += 3;
        const int fs = aux_size(type);
        if (fs > 0) { p += fs; }
        else if (type == 'Z' || type == 'H') { while (p < end && *p) p++; p++; }
        else if (type == 'B') {
            if (p + 5 > end) return v;
            const int es = aux_size(p[0]);
            const int32_t cnt = rd<int32_t>(p + 1);
            if (es < 0 || cnt < 0 || p + 5 + (size_t)es * (size_t)cnt > end) return v;
            if (t0 == 'C' && t1 == 'G' && p[0] == 'I') { v.p = p + 5; v.n = cnt; return v; }
            p += 5 + (size_t)es * (size_t)cnt;
        } else return v;
    }
    return v;
}
int64_t real_cigar_count(const uint8_t* r, int32_t bs) { return real_cigar(r, bs).n; }

}  // namespace

extern "C" {

int nc_bam_open(const char* path, int threads, nc_bam** out) {
    if (!path || !out) return NC_IO_EINVAL;
    *out = nullptr;
    nc_bam* b = new nc_bam();
    *out = b;                                   // returned even on failure so that nc_bam_error can explain
    MapFile f(path);
    if (f.err) { b->err = std::string(f.err == 1 ? "cannot open " : "cannot map ") + path; return NC_IO_EOPEN; }
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    int rc = inflate_all(b, f.p, f.n, threads);
    if (rc) return rc;
    const int64_t first = parse_header(b, b->data.data(), b->data.size());
    if (first < 0) return NC_IO_EFORMAT;
    if (first == 0) return fail(b, b->data.size() < 4 || memcmp(b->data.data(), "BAM\1", 4) != 0 ? "not a BAM file (bad magic)" : "truncated BAM header");
    std::vector<std::vector<int64_t>> per(b->contigs.size());
    rc = scan_records(b, (size_t)first, b->data.size(), -1, per);
    if (rc) return rc;
    finish_index(b, per);
    return NC_IO_OK;
}

// Region read through a BAI index: only the BGZF blocks that hold the named contigs' records are inflated (what
// samfile.fetch(chrom, ...) does for the reference, generate_SNP_pileups.py:141,156).  The handle lists every reference of the
// header; contigs that were not asked for report zero reads.
int nc_bam_open_region(const char* path, const char* bai_path, const char* const* names, int n_names, int threads, nc_bam** out) {
    if (!path || !bai_path || !out || n_names < 0 || (n_names > 0 && !names)) return NC_IO_EINVAL;
    *out = nullptr;
    nc_bam* b = new nc_bam();
    *out = b;
    MapFile f(path), x(bai_path);
    if (f.err) { b->err = std::string("cannot open ") + path; return NC_IO_EOPEN; }
    if (x.err) { b->err = std::string("cannot open ") + bai_path; return NC_IO_EOPEN; }
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    // ---- header: inflate leading blocks one at a time until it parses
    std::vector<uint8_t> head;
    {
        size_t off = 0;
        for (;;) {
            std::vector<Blk> one;
            size_t tot = 0;
            if (off >= f.n) return fail(b, "truncated BAM header");
            // one block: its end is found by the table walker
            int rc = block_table(b, f.p, f.n, off, off + 1, one, tot);
            if (rc) return rc;
            const Blk& k = one[0];
            const size_t at = head.size();
            head.resize(at + k.isize);
            if (k.isize) {
                z_stream zs;
                memset(&zs, 0, sizeof(zs));
                if (inflateInit2(&zs, -15) != Z_OK) return fail(b, "inflate failed");
                zs.next_in = const_cast<Bytef*>(f.p + k.off); zs.avail_in = (uInt)k.csize;
                zs.next_out = head.data() + at; zs.avail_out = k.isize;
                const int zr = inflate(&zs, Z_FINISH);
                inflateEnd(&zs);
                if (zr != Z_STREAM_END) return fail(b, "inflate failed (corrupt BGZF block)");
            }
            off = k.off + k.csize + 8;
            const int64_t first = parse_header(b, head.data(), head.size());
            if (first < 0) return NC_IO_EFORMAT;
            if (first > 0) break;
        }
    }
    const int32_t n_ref = (int32_t)b->contigs.size();
    // ---- BAI: per reference the smallest chunk begin and the largest chunk end (virtual offsets)
    std::vector<uint64_t> vbeg((size_t)n_ref, UINT64_MAX), vend((size_t)n_ref, 0);
    {
        const uint8_t* p = x.p;
        const uint8_t* e = x.p + x.n;
        if (x.n < 8 || memcmp(p, "BAI\1", 4) != 0) return fail(b, "not a BAI index (bad magic)");
        const int32_t nr = rd<int32_t>(p + 4);
        p += 8;
        if (nr != n_ref) return fail(b, "BAI index does not match the BAM header (reference count)");
        for (int32_t i = 0; i < nr; i++) {
            if (p + 4 > e) return fail(b, "truncated BAI index");
            const int32_t n_bin = rd<int32_t>(p); p += 4;
            for (int32_t k = 0; k < n_bin; k++) {
                if (p + 8 > e) return fail(b, "truncated BAI index");
                const uint32_t bin = rd<uint32_t>(p);
                const int32_t n_chunk = rd<int32_t>(p + 4);
                p += 8;
                if (n_chunk < 0 || p + 16 * (size_t)n_chunk > e) return fail(b, "truncated BAI index");
                if (bin != 37450)                           // the pseudo-bin carries counts, not chunks
                    for (int32_t c = 0; c < n_chunk; c++) {
                        vbeg[i] = std::min(vbeg[i], rd<uint64_t>(p + 16 * c));
                        vend[i] = std::max(vend[i], rd<uint64_t>(p + 16 * c + 8));
                    }
                p += 16 * (size_t)n_chunk;
            }
            if (p + 4 > e) return fail(b, "truncated BAI index");
            const int32_t n_intv = rd<int32_t>(p); p += 4;
            if (n_intv < 0 || p + 8 * (size_t)n_intv > e) return fail(b, "truncated BAI index");
            p += 8 * (size_t)n_intv;
        }
    }
    // ---- block tables of the requested contigs, one inflate, one record walk per contig
    struct Seg { int32_t rid; size_t data_beg, data_end; };
    std::vector<Seg> segs;
    std::vector<Blk> blocks;
    size_t total = 0;
    for (int k = 0; k < n_names; k++) {
        int32_t rid = -1;
        for (int32_t i = 0; i < n_ref; i++) if (b->contigs[i].name == names[k]) { rid = i; break; }
        if (rid < 0 || vbeg[rid] == UINT64_MAX) continue;        // unknown name or no reads: the contig stays empty
        bool dup = false;
        for (const Seg& sg : segs) dup |= sg.rid == rid;
        if (dup) continue;
        const size_t cbeg = (size_t)(vbeg[rid] >> 16), cend_blk = (size_t)(vend[rid] >> 16);
        const uint32_t ubeg = (uint32_t)(vbeg[rid] & 0xFFFF), uend = (uint32_t)(vend[rid] & 0xFFFF);
        if (cbeg >= f.n || cend_blk > f.n) return fail(b, "BAI index points past the end of the BAM file");
        const size_t seg0 = total, nb0 = blocks.size();
        int rc = block_table(b, f.p, f.n, cbeg, uend > 0 ? cend_blk + 1 : cend_blk, blocks, total);
        if (rc) return rc;
        // data offset of the last record's end: start of the block at cend_blk (if it was included) + uend
        size_t end_data = total;
        if (uend > 0) {
            for (size_t q = nb0; q < blocks.size(); q++)
                if (blocks[q].start == cend_blk) { end_data = blocks[q].out + uend; break; }
        }
        segs.push_back({rid, seg0 + ubeg, std::min(end_data, total)});
    }
    int rc = inflate_blocks(b, f.p, blocks, total, threads);
    if (rc) return rc;
    std::vector<std::vector<int64_t>> per((size_t)n_ref);
    for (const Seg& sg : segs) {
        if (sg.data_beg > sg.data_end) return fail(b, "BAI index is inconsistent with the BAM file");
        rc = scan_records(b, sg.data_beg, sg.data_end, sg.rid, per);
        if (rc) return rc;
    }
    finish_index(b, per);
    return NC_IO_OK;
}

const char* nc_bam_error(const nc_bam* b) { return b ? b->err.c_str() : "null handle"; }
int nc_bam_n_contigs(const nc_bam* b) { return b ? (int)b->contigs.size() : NC_IO_EINVAL; }
const char* nc_bam_header_text(const nc_bam* b, int64_t* len) {
    if (!b) return nullptr;
    if (len) *len = (int64_t)b->text.size();
    return b->text.data();
}

int nc_bam_contig(const nc_bam* b, int i, NcBamContig* out) {
    if (!b || !out || i < 0 || i >= (int)b->contigs.size()) return NC_IO_EINVAL;
    const Contig& c = b->contigs[(size_t)i];
    memset(out, 0, sizeof(*out));
    snprintf(out->name, sizeof(out->name), "%s", c.name.c_str());
    out->length = c.length; out->n_reads = c.n_reads; out->n_cigar = c.n_cigar; out->n_seq = c.n_seq;
    return NC_IO_OK;
}

int nc_bam_fill(const nc_bam* b, int i, int threads, int32_t* pos, uint16_t* flag, int64_t* cigar_off, uint32_t* cigar, int64_t* seq_off,
                int32_t* l_seq, uint8_t* seq4, int8_t* hp, int32_t* ps) {
    if (!b || i < 0 || i >= (int)b->contigs.size() || !pos || !flag || !cigar_off || !seq_off || !l_seq) return NC_IO_EINVAL;
    const Contig& c = b->contigs[(size_t)i];
    if ((c.n_cigar > 0 && !cigar) || (c.n_seq > 0 && !seq4)) return NC_IO_EINVAL;
    const uint8_t* d = b->data.data();
    const int64_t* ro = b->rec_off.data() + c.first_rec;
    // offsets: one sequential pass (cheap), then the payload copies in parallel
    int64_t co = 0, so = 0;
    for (int64_t k = 0; k < c.n_reads; k++) {
        const uint8_t* r = d + ro[k] + 4;
        cigar_off[k] = co; seq_off[k] = so;
        co += real_cigar(r, rd<int32_t>(r - 4)).n;
        so += (rd<int32_t>(r + 16) + 1) / 2;
    }
    cigar_off[c.n_reads] = co; seq_off[c.n_reads] = so;
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    parallel_for(c.n_reads, threads, [&](int64_t k) {
        const uint8_t* base = d + ro[k];
        const int32_t bs = rd<int32_t>(base);
        const uint8_t* r = base + 4;
        pos[k] = rd<int32_t>(r + 4);
        const uint8_t l_name = r[8];
        const uint16_t n_cig = rd<uint16_t>(r + 12);
        flag[k] = rd<uint16_t>(r + 14);
        const int32_t ls = rd<int32_t>(r + 16);
        l_seq[k] = ls;
        const uint8_t* p = r + 32 + l_name;
        const CigarView cv = real_cigar(r, bs);
        if (cv.n) memcpy(cigar + cigar_off[k], cv.p, 4 * (size_t)cv.n);
        p += 4 * (size_t)n_cig;
        const size_t nb = (size_t)(ls + 1) / 2;
        if (nb) memcpy(seq4 + seq_off[k], p, nb);
        p += nb + (size_t)ls;
        if (hp && ps) scan_tags(p, r + bs, hp + k, ps + k);
    });
    return NC_IO_OK;
}

int nc_bam_qname(const nc_bam* b, int i, int64_t k, char* out, int cap) {
    if (!b || !out || cap <= 0 || i < 0 || i >= (int)b->contigs.size()) return NC_IO_EINVAL;
    const Contig& c = b->contigs[(size_t)i];
    if (k < 0 || k >= c.n_reads) return NC_IO_EINVAL;
    const uint8_t* r = b->data.data() + b->rec_off[(size_t)(c.first_rec + k)] + 4;
    snprintf(out, (size_t)cap, "%s", (const char*)r + 32);
    return NC_IO_OK;
}

// Haplotagged copy of one contig's records (what `whatshap haplotag ... | samtools view -b` leaves in
// intermediate_phase_files/{contig}.phased.bam, indelCaller.py:244): every record is copied as it is — name, qualities and all
// other aux fields — with its HP / PS fields removed and, where hp[k] > 0, `HP:C` and `PS:i` appended.  The header lists all
// references of the source file.  BGZF blocks of <= 0xff00 payload bytes are deflated in parallel.
int nc_bam_write_tagged(const nc_bam* b, int i, const int8_t* hp, const int32_t* ps, const char* out_path, int level, int threads) {
    if (!b || !out_path || i < 0 || i >= (int)b->contigs.size()) return NC_IO_EINVAL;
    const Contig& c = b->contigs[(size_t)i];
    if (c.n_reads > 0 && (!hp || !ps)) return NC_IO_EINVAL;
    std::vector<uint8_t> u;
    auto put32 = [&](int32_t v) { uint8_t t[4]; memcpy(t, &v, 4); u.insert(u.end(), t, t + 4); };
    u.insert(u.end(), {'B', 'A', 'M', 1});
    put32((int32_t)b->text.size());
    u.insert(u.end(), b->text.begin(), b->text.end());
    put32((int32_t)b->contigs.size());
    for (const Contig& k : b->contigs) {
        put32((int32_t)k.name.size() + 1);
        u.insert(u.end(), k.name.begin(), k.name.end());
        u.push_back(0);
        put32(k.length);
    }
    // BGZF output in batches: the stream is deflated and written every ~64 MB, so memory stays bounded for any contig size
    const size_t kPayload = 0xff00, kBatch = kPayload * 1024;
    if (level < 0 || level > 9) level = 4;
    FILE* f = fopen(out_path, "wb");
    if (!f) return NC_IO_EOPEN;
    bool ok = true;
    auto flush = [&](bool all) {                                // writes every full block of `u` (all = the tail as well)
        const size_t n_full = all ? u.size() : (u.size() / kPayload) * kPayload;
        if (n_full == 0) return;
        const int64_t nblk = (int64_t)((n_full + kPayload - 1) / kPayload);
        std::vector<std::vector<uint8_t>> out((size_t)nblk);
        std::atomic<int> bad{0};
        parallel_for(nblk, threads, [&](int64_t j) {
            const size_t o = (size_t)j * kPayload, n = std::min(kPayload, n_full - o);
            std::vector<uint8_t>& blk = out[(size_t)j];
            blk.resize(18 + compressBound((uLong)n) + 8);
            z_stream zs;
            memset(&zs, 0, sizeof(zs));
            if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { bad = 1; return; }
            zs.next_in = u.data() + o; zs.avail_in = (uInt)n;
            zs.next_out = blk.data() + 18; zs.avail_out = (uInt)(blk.size() - 18 - 8);
            const int rc = deflate(&zs, Z_FINISH);
            const size_t cs = zs.total_out;
            deflateEnd(&zs);
            if (rc != Z_STREAM_END || 18 + cs + 8 > 0x10000) { bad = 1; return; }
            static const uint8_t head[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
            memcpy(blk.data(), head, 16);
            const uint16_t bsize = (uint16_t)(18 + cs + 8 - 1);
            memcpy(blk.data() + 16, &bsize, 2);
            const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), u.data() + o, (uInt)n), isz = (uint32_t)n;
            memcpy(blk.data() + 18 + cs, &crc, 4);
            memcpy(blk.data() + 18 + cs + 4, &isz, 4);
            blk.resize(18 + cs + 8);
        });
        if (bad) { ok = false; return; }
        for (auto& blk : out) ok = ok && fwrite(blk.data(), 1, blk.size(), f) == blk.size();
        u.erase(u.begin(), u.begin() + (std::ptrdiff_t)n_full);
    };
    const uint8_t* d = b->data.data();
    for (int64_t k = 0; k < c.n_reads && ok; k++) {
        const size_t off = (size_t)b->rec_off[(size_t)(c.first_rec + k)];
        const int32_t bs = rd<int32_t>(d + off);
        const uint8_t* r = d + off + 4;
        const uint8_t l_name = r[8];
        const uint16_t n_cig = rd<uint16_t>(r + 12);
        const int32_t l_seq = rd<int32_t>(r + 16);
        const size_t fixed = 32 + (size_t)l_name + 4 * (size_t)n_cig + (size_t)(l_seq + 1) / 2 + (size_t)l_seq;
        const size_t at = u.size();
        put32(0);                                               // block_size, patched below
        u.insert(u.end(), r, r + fixed);
        const uint8_t* p = r + fixed;
        const uint8_t* end = r + bs;
        while (p + 3 <= end) {                                  // copy every aux field except HP / PS
            const uint8_t* f0 = p;
            const uint8_t t0 = p[0], t1 = p[1], type = p[2];
            p += 3;
            const int fs = aux_size(type);
            if (fs > 0) p += fs;
            else if (type == 'Z' || type == 'H') { while (p < end && *p) p++; p++; }
            else if (type == 'B') {
                if (p + 5 > end) { p = end; break; }
                const int es = aux_size(p[0]);
                const int32_t cnt = rd<int32_t>(p + 1);
                if (es < 0 || cnt < 0) { p = end; break; }
                p += 5 + (size_t)es * (size_t)cnt;
            } else { p = end; }                                 // unknown type: keep the rest verbatim
            if (p > end) p = end;
            const bool drop = (t0 == 'H' && t1 == 'P') || (t0 == 'P' && t1 == 'S');
            if (!drop) u.insert(u.end(), f0, p);
        }
        if (p < end) u.insert(u.end(), p, end);
        if (hp[k] > 0) {
            u.insert(u.end(), {'H', 'P', 'C', (uint8_t)hp[k], 'P', 'S', 'i'});
            put32(ps[k]);
        }
        const int32_t nbs = (int32_t)(u.size() - at - 4);
        memcpy(u.data() + at, &nbs, 4);
        if (u.size() >= kBatch) flush(false);
    }
    if (ok) flush(true);
    static const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    ok = ok && fwrite(eof, 1, 28, f) == 28;
    ok = (fclose(f) == 0) && ok;
    return ok ? NC_IO_OK : NC_IO_EOPEN;
}

void nc_bam_close(nc_bam* b) { delete b; }

}  // extern "C"
