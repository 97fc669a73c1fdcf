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

int inflate_all(nc_bam* b, const uint8_t* file, size_t n, int threads) {
    struct Blk { size_t off, csize; uint32_t isize; size_t out; };
    std::vector<Blk> blocks;
    size_t off = 0, total = 0;
    while (off + 18 <= n) {
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
        const uint32_t isize = rd<uint32_t>(p + blen - 4);
        blocks.push_back({off + xend, blen - xend - 8, isize, total});
        total += isize;
        off += blen;
    }
    if (off != n) return fail(b, "trailing bytes after the last BGZF block");
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

}  // namespace

extern "C" {

int nc_bam_open(const char* path, int threads, nc_bam** out) {
    if (!path || !out) return NC_IO_EINVAL;
    *out = nullptr;
    nc_bam* b = new nc_bam();
    *out = b;                                   // returned even on failure so that nc_bam_error can explain
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { b->err = std::string("cannot open ") + path; return NC_IO_EOPEN; }
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); b->err = "cannot stat the file"; return NC_IO_EOPEN; }
    const size_t fsize = (size_t)st.st_size;
    void* map = fsize ? mmap(nullptr, fsize, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;     // the compressed bytes are read straight from the page cache
    close(fd);
    if (fsize && map == MAP_FAILED) { b->err = "mmap failed"; return NC_IO_EOPEN; }
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    int rc = inflate_all(b, (const uint8_t*)map, fsize, threads);
    if (map) munmap(map, fsize);
    if (rc) return rc;
    const RawBuf& d = b->data;
    const size_t n = d.size();
    if (n < 12 || memcmp(d.data(), "BAM\1", 4) != 0) return fail(b, "not a BAM file (bad magic)");
    const int32_t l_text = rd<int32_t>(d.data() + 4);
    if (l_text < 0 || 8 + (size_t)l_text + 4 > n) return fail(b, "truncated BAM header");
    b->text.assign((const char*)d.data() + 8, (size_t)l_text);
    size_t off = 8 + (size_t)l_text;
    const int32_t n_ref = rd<int32_t>(d.data() + off);
    off += 4;
    if (n_ref < 0) return fail(b, "negative reference count");
    b->contigs.resize((size_t)n_ref);
    for (int32_t i = 0; i < n_ref; i++) {
        if (off + 4 > n) return fail(b, "truncated reference list");
        const int32_t l_name = rd<int32_t>(d.data() + off);
        if (l_name <= 0 || off + 4 + (size_t)l_name + 4 > n) return fail(b, "truncated reference list");
        b->contigs[i].name.assign((const char*)d.data() + off + 4, (size_t)l_name - 1);
        b->contigs[i].length = rd<int32_t>(d.data() + off + 4 + l_name);
        off += 8 + (size_t)l_name;
    }
    // record boundaries and per-contig totals
    std::vector<std::vector<int64_t>> per((size_t)n_ref);
    int32_t last_rid = -1, last_pos = -1;
    while (off + 4 <= n) {
        const int32_t bs = rd<int32_t>(d.data() + off);
        if (bs < 32 || off + 4 + (size_t)bs > n) return fail(b, "truncated alignment record");
        const uint8_t* r = d.data() + off + 4;
        const int32_t rid = rd<int32_t>(r), pos = rd<int32_t>(r + 4);
        const uint8_t l_name = r[8];
        const uint16_t n_cig = rd<uint16_t>(r + 12);
        const int32_t l_seq = rd<int32_t>(r + 16);
        if (l_seq < 0 || 32 + (size_t)l_name + 4 * (size_t)n_cig + (size_t)(l_seq + 1) / 2 + (size_t)l_seq > (size_t)bs) return fail(b, "alignment record fields exceed its size");
        if (rid >= 0 && rid < n_ref) {
            if (rid < last_rid || (rid == last_rid && pos < last_pos)) b->sorted = false;
            last_rid = rid; last_pos = pos;
            Contig& c = b->contigs[(size_t)rid];
            c.n_reads++; c.n_cigar += n_cig; c.n_seq += (l_seq + 1) / 2;
            per[(size_t)rid].push_back((int64_t)off);
        }
        off += 4 + (size_t)bs;
    }
    if (!b->sorted) return fail(b, "BAM is not coordinate-sorted");
    for (int32_t i = 0; i < n_ref; i++) {
        b->contigs[i].first_rec = (int64_t)b->rec_off.size();
        b->rec_off.insert(b->rec_off.end(), per[i].begin(), per[i].end());
    }
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
        co += rd<uint16_t>(r + 12);
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
        if (n_cig) memcpy(cigar + cigar_off[k], p, 4 * (size_t)n_cig);
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

void nc_bam_close(nc_bam* b) { delete b; }

}  // extern "C"
