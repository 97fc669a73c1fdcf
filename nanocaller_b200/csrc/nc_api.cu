// libnanocaller_b200.so — C-ABI (include/nanocaller_b200.h) over the sm_100a kernels.
// One context = one device + one stream; every entry point returns 0 or a negative NC_E* code.
#include <algorithm>
#include <cmath>
#include <atomic>
#include <thread>
#include <cstdarg>
#include <cstdlib>
#include <utility>

#include "nc_common.cuh"
#include "nc_pileup.cuh"
#include "nc_cnn.cuh"
#include "nc_cnn_tc.cuh"
#include "nc_cnn_tc_indel.cuh"
#include "nc_indel.cuh"
#include "nc_bgzf.cuh"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <chrono>

using namespace nc;

namespace {

struct LayerRef { int64_t k = 0, b = 0; int KH = 0, KW = 0, Cin = 0, Cout = 0; };

struct Model {
    bool loaded = false;
    int kind = 0;                 // 0 SNP diploid, 1 SNP haploid, 2 indel diploid, 3 indel haploid
    int Hin = 0, Win = 0, Cin = 0, C1 = 0, C2 = 0, C3 = 0, F1 = 0;
    int H2 = 0, W2 = 0, H3 = 0, W3 = 0, flat = 0;
    double train_cov = 0;
    size_t n_floats = 0;
    DevBuf w;
    LayerRef conv[5], fc1;        // conv1_1 conv1_2 conv1_3 conv2 conv3
    int64_t tail[16][2];          // (kernel, bias) offsets of the tail layers in blob order
    DevBuf ktab[5];
    TcModel tc;                   // tensor-core operand images (nc_cnn_tc.cuh)
};

}  // namespace

struct nc_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0;

    // staged contig
    bool staged = false, decoded = false, scanned = false;
    int64_t n_reads = 0, n_cigar = 0, n_seq = 0, ref_start = 0, ref_len = 0;
    uint64_t decode_bytes = 0;
    DevBuf d_pos, d_flag, d_cigar_off, d_cigar, d_seq_off, d_lseq, d_seq4, d_ref, d_fill_counter, d_seqc;
    // decode products
    DevBuf d_end, d_nwords, d_opstart, d_pmaxend, d_rowoff, d_rows;
    int64_t n_row_words = 0;
    // scan products
    DevBuf d_flags, d_tile_nbr, d_tile_cand, d_nbr_off, d_cand_off, d_nbr_pos, d_cand_pos, d_bed;
    DevBuf d_nfirst, d_nlen, d_nbytes, d_noff, d_nrows;
    DevBuf d_chunks, d_chunk_lo, d_chunk_cnt, d_chunk_off, d_keep, d_keep32, d_outidx;
    DevBuf d_mat, d_meta, d_depth_sum, d_depth_cnt, d_chunk_depth, d_chunk_count, d_probs;
    DevBuf d_scan_partial;
    PinBuf pin;
    int64_t n_nbr = 0, n_cand = 0, n_slots = 0, n_sites = 0;
    int32_t n_chunks = 0;
    bool have_probs = false;
    int scan_haploid = 0;
    // indel path
    bool tags_staged = false, indel_scanned = false, indel_built = false;
    DevBuf d_hp, d_ps, d_idepth, d_em, d_grank, d_empos, d_ichunks, d_nem1, d_rankoff, d_diff, d_empairs, d_hit, d_variants, d_icount;
    DevBuf d_isites, d_site_m, d_site_cnt, d_site_off, d_eread, d_eqpn, d_eslice, d_eacode, d_einslen, d_einsfirst, d_en, d_itensors, d_icns, d_imeta;
    // impute_indel_phase: per-column indel marks, pending / source columns and their grouped reads
    DevBuf d_cdel, d_cins, d_imp_g, d_imp_cols, d_imp_cnt, d_imp_off, d_imp_read, d_imp_ch, d_imp_ind, d_imp_qn, d_imp_gid, d_imp_rep, d_imp_gcnt,
           d_imp_label, d_imp_ok, d_site_imp, d_egrp;
    int64_t n_variants = 0, n_isites = 0, indel_R = 0;
    int32_t indel_lo_al = 0;
    // CNN
    Model snp[2], indel[2];
    DevBuf ws_c1, ws_c2, ws_c3, ws_f1, ws_sf, ws_sd, ws_x, ws_ref, ws_out;
    // timings
    cudaEvent_t ev_block = nullptr;   // blocking-sync event (NC_BLOCKING_SYNC=1), else spin on the stream
    cudaEvent_t ev[13] = {};     // 0-7 phase timings, 8-11 user slots (nc_event_record), 12 end of the conv1/conv2 kernel
    cudaEvent_t evi[9] = {};     // indel path: 0-1 scan, 2-5 build (start, slices, align, msa), 6-7 CNN, 8 allele prediction (after msa)
    NcIndelTimings tmi = {};
    bool tmi_scan = false, tmi_build = false, tmi_cnn = false, have_iprobs = false;
    int build_haploid = 0;
    DevBuf d_iprobs, d_ialleles, d_allele_dirs, d_allele_ops;
    // device-side BAM input
    struct BamContig { std::string name; int32_t length = 0; int64_t first = 0, n = 0, n_tagged = 0; };
    DevBuf d_bam_comp, d_bam_blocks, d_bam, d_bam_recoff, d_bam_rid, d_bam_pos, d_bam_flag, d_bam_lseq, d_bam_ncig, d_bam_nseq, d_bam_cigsrc, d_bam_seqsrc,
           d_bam_hp, d_bam_ps, d_bam_err, d_bam_out, d_bam_wst, d_bam_wcnt, d_bam_woff;
    PinBuf pin_bam;
    std::vector<BamContig> bam_contigs;
    int64_t bam_records = 0, bam_bytes = 0, bam_comp_bytes = 0;
    float bam_ms[4] = {0, 0, 0, 0};
    bool bam_open = false, bam_walk_parallel = false;
    NcTimings tm = {};
    bool tm_decode = false, tm_scan = false, tm_cnn = false, tm_cnn_a = false;
};

static inline cudaError_t nc_stream_wait(nc_ctx* c) {
    if (!c->ev_block) return cudaStreamSynchronize(c->stream);
    cudaError_t e = cudaEventRecord(c->ev_block, c->stream);
    return e == cudaSuccess ? cudaEventSynchronize(c->ev_block) : e;
}


namespace {

int fail(nc_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return code;
}

#define NC_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(c, e_ == cudaErrorMemoryAllocation ? NC_ENOMEM : NC_ECUDA, "%s:%d %s: %s", \
                        __FILE__, __LINE__, #call, cudaGetErrorString(e_));                        \
    } while (0)

// Wait for the context's stream.  After nc_set_blocking_sync(ctx, 1) (or with NC_BLOCKING_SYNC=1 in the environment) the host
// thread sleeps on an event created with cudaEventBlockingSync instead of spinning.  Measured with 4 ranks x 2 pipelined contexts
// on one host: 63.3 against 52.1 M sites/s end to end, at the price of 8 % of the device-resident figure (wake-up latency at
// every host sync of a scan); neutral at 2 ranks.  Off by default.
#define NC_LAUNCH_CHECK()                                                                          \
    do {                                                                                           \
        c->launches++;                                                                             \
        cudaError_t e_ = cudaGetLastError();                                                       \
        if (e_ != cudaSuccess)                                                                     \
            return fail(c, NC_ECUDA, "%s:%d kernel launch: %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
    } while (0)

int upload(nc_ctx* c, DevBuf& d, const void* h, size_t bytes) {
    NC_CUDA(d.reserve(bytes ? bytes : 16));
    if (bytes) NC_CUDA(cudaMemcpyAsync(d.p, h, bytes, cudaMemcpyHostToDevice, c->stream));
    return NC_OK;
}

// exclusive scan int32[n] -> int64[n+1]
int device_scan(nc_ctx* c, const int32_t* in, int64_t n, int64_t* out) {
    if (n <= 0) { NC_CUDA(cudaMemsetAsync(out, 0, sizeof(int64_t), c->stream)); return NC_OK; }
    const int64_t nparts = div_up(n, kScanTile);
    NC_CUDA(c->d_scan_partial.reserve((size_t)(nparts + 1) * sizeof(int64_t)));
    int64_t* part = c->d_scan_partial.as<int64_t>();
    scan_reduce_kernel<<<(unsigned)nparts, kScanThreads, 0, c->stream>>>(in, n, part); NC_LAUNCH_CHECK();
    scan_partials_kernel<<<1, 1024, 0, c->stream>>>(part, nparts); NC_LAUNCH_CHECK();
    scan_apply_kernel<<<(unsigned)nparts, kScanThreads, 0, c->stream>>>(in, n, part, nparts, out); NC_LAUNCH_CHECK();
    return NC_OK;
}

// read one int64 from the device (synchronises the stream)
int read_i64(nc_ctx* c, const int64_t* dev, int64_t* out) {
    NC_CUDA(c->pin.reserve(64));
    NC_CUDA(cudaMemcpyAsync(c->pin.p, dev, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    NC_CUDA(nc_stream_wait(c));
    *out = *c->pin.as<int64_t>();
    return NC_OK;
}

// ------------------------------------------------------------------------------------------------
// model tables
// ------------------------------------------------------------------------------------------------
int model_init(nc_ctx* c, Model& M, int kind, const float* blob, size_t n_floats, double train_cov) {
    M.loaded = false;
    M.kind = kind; M.train_cov = train_cov;
    const bool snp = kind < 2;
    if (snp) { M.Hin = 5; M.Win = 41; M.Cin = 5; M.C1 = 16; M.C2 = 32; M.C3 = 64; M.F1 = 48; }
    else { M.Hin = kind == 2 ? 15 : 5; M.Win = 128; M.Cin = 2; M.C1 = 8; M.C2 = 32; M.C3 = 48; M.F1 = 32; }
    M.H2 = M.Hin - 1; M.W2 = (M.Win - 3) / 2 + 1; M.H3 = M.H2 - 1; M.W3 = (M.W2 - 3) / 2 + 1;
    M.flat = M.H3 * M.W3 * M.C3;
    int64_t off = 0;
    auto take = [&](LayerRef& L, int KH, int KW, int Cin, int Cout) {
        L.KH = KH; L.KW = KW; L.Cin = Cin; L.Cout = Cout;
        L.k = off; off += (int64_t)KH * KW * Cin * Cout; L.b = off; off += Cout;
    };
    take(M.conv[0], 1, 5, M.Cin, M.C1);
    take(M.conv[1], 5, 1, M.Cin, M.C1);
    take(M.conv[2], 5, 5, M.Cin, M.C1);
    take(M.conv[3], 2, 3, 3 * M.C1, M.C2);
    take(M.conv[4], 2, 3, M.C2, M.C3);
    take(M.fc1, 1, 1, M.flat, M.F1);
    int nt = 0;
    auto tail = [&](int nin, int nout) { M.tail[nt][0] = off; off += (int64_t)nin * nout; M.tail[nt][1] = off; off += nout; nt++; };
    if (kind == 0) { tail(48, 16); for (int j = 0; j < 4; j++) tail(17, 2); tail(48, 16); tail(24, 8); tail(8, 2); }
    else if (kind == 1) { tail(48, 16); tail(20, 4); }
    else if (kind == 2) { tail(32, 24); tail(24, 4); }
    else { tail(32, 24); tail(24, 1); }
    if ((size_t)off != n_floats)
        return fail(c, NC_EINVAL, "weight blob has %zu floats, model kind %d needs %lld", n_floats, kind, (long long)off);
    M.n_floats = n_floats;
    int rc = upload(c, M.w, blob, n_floats * sizeof(float));
    if (rc) return rc;
    for (int l = 0; l < 5; l++) {
        const LayerRef& L = M.conv[l];
        std::vector<int32_t> tab((size_t)L.KH * L.KW * L.Cin);
        size_t k = 0;
        for (int kh = 0; kh < L.KH; kh++)
            for (int kw = 0; kw < L.KW; kw++)
                for (int ci = 0; ci < L.Cin; ci++) tab[k++] = (kh << 24) | (kw << 16) | ci;
        rc = upload(c, M.ktab[l], tab.data(), tab.size() * sizeof(int32_t));
        if (rc) return rc;
    }
    NC_CUDA(nc_stream_wait(c));     // `tab` is pageable host memory
    rc = kind < 2 ? tc_model_prepare(c->stream, M.tc, kind, blob, n_floats, &c->err) : tci_model_prepare(c->stream, M.tc, kind, blob, n_floats, &c->err);
    if (rc) return rc;
    M.loaded = true;
    return NC_OK;
}

template <int BN, int MODE>
void conv_launch(nc_ctx* c, const ConvArgs& a) {
    const unsigned grid = (unsigned)div_up(a.M, 64);
    conv_f32_kernel<BN, MODE><<<grid, 16 * (BN / 4), 0, c->stream>>>(a);
}

int conv_dispatch(nc_ctx* c, int in_mode, const ConvArgs& a) {
    if (a.M <= 0) return NC_OK;
    if (in_mode == 1 && a.Cout == 16) conv_launch<16, 1>(c, a);
    else if (in_mode == 2 && a.Cout == 16) conv_launch<16, 2>(c, a);
    else if (in_mode != 0) return fail(c, NC_EINVAL, "int16 input only feeds the 16-channel SNP first layer");
    else if (a.Cout == 8) conv_launch<8, 0>(c, a);
    else if (a.Cout == 16) conv_launch<16, 0>(c, a);
    else if (a.Cout == 32) conv_launch<32, 0>(c, a);
    else if (a.Cout == 48) conv_launch<48, 0>(c, a);
    else if (a.Cout == 64) conv_launch<64, 0>(c, a);
    else return fail(c, NC_EINVAL, "unsupported Cout %d", a.Cout);
    NC_LAUNCH_CHECK();
    return NC_OK;
}

TailW tail_weights(const Model& M) {
    TailW t = {};
    const float* w = M.w.as<float>();
    if (M.kind == 0) {
        t.fa_k = w + M.tail[0][0]; t.fa_b = w + M.tail[0][1];
        for (int j = 0; j < 4; j++) { t.hk[j] = w + M.tail[1 + j][0]; t.hb[j] = w + M.tail[1 + j][1]; }
        t.fc2_k = w + M.tail[5][0]; t.fc2_b = w + M.tail[5][1];
        t.fc3_k = w + M.tail[6][0]; t.fc3_b = w + M.tail[6][1];
        t.gt_k = w + M.tail[7][0]; t.gt_b = w + M.tail[7][1];
    } else {
        t.fc2_k = w + M.tail[0][0]; t.fc2_b = w + M.tail[0][1];
        t.fc3_k = w + M.tail[1][0]; t.fc3_b = w + M.tail[1][1];
    }
    return t;
}

// fp32 CUDA-core forward over n sites.  in_mode 0: fp32 NHWC input; 1/2: int16 SNP tensors + scale.
// outputs: SNP diploid -> out_full [n][10] and/or probs [n][4]; SNP haploid -> probs [n][4];
// indel -> out_full [n][4] / [n][1].
int cnn_forward_f32(nc_ctx* c, Model& M, int in_mode, const void* in_dev, int64_t in_site_stride, int64_t n,
                    const NcSiteMeta* meta, const float* ref4, const float* scale_f, const double* scale_d,
                    float* out_full, float* probs) {
    if (n <= 0) return NC_OK;
    const int64_t c1_site = (int64_t)M.Hin * M.Win * 3 * M.C1, c2_site = (int64_t)M.H2 * M.W2 * M.C2;
    // batch so that the widest activation stays L2-resident (~64 MB)
    int64_t NB = std::max<int64_t>(64, (64ll << 20) / (c1_site * 4));
    NB = std::min(NB, n);
    NC_CUDA(c->ws_c1.reserve((size_t)NB * c1_site * 4));
    NC_CUDA(c->ws_c2.reserve((size_t)NB * c2_site * 4));
    NC_CUDA(c->ws_c3.reserve((size_t)NB * M.flat * 4));
    NC_CUDA(c->ws_f1.reserve((size_t)NB * M.F1 * 4));
    const float* w = M.w.as<float>();
    const size_t in_elt = in_mode == 0 ? 4 : 2;
    const TailW tw = tail_weights(M);
    for (int64_t b0 = 0; b0 < n; b0 += NB) {
        const int64_t nb = std::min(NB, n - b0);
        const char* in_b = reinterpret_cast<const char*>(in_dev) + (size_t)b0 * in_site_stride * in_elt;
        for (int l = 0; l < 3; l++) {
            const LayerRef& L = M.conv[l];
            ConvArgs a = {};
            a.in = in_b; a.w = w + L.k; a.bias = w + L.b; a.out = c->ws_c1.as<float>();
            a.scale_f = scale_f ? scale_f + b0 : nullptr; a.scale_d = scale_d ? scale_d + b0 : nullptr;
            a.ktab = M.ktab[l].as<int32_t>();
            a.M = nb * M.Hin * M.Win; a.in_site_stride = in_site_stride;
            a.K = L.KH * L.KW * L.Cin; a.Hin = M.Hin; a.Win = M.Win; a.Cin = M.Cin; a.SW = 1;
            a.PH = L.KH / 2; a.PW = L.KW / 2; a.Hout = M.Hin; a.Wout = M.Win; a.Cout = M.C1;
            a.out_cstride = 3 * M.C1; a.out_coff = l * M.C1; a.selu = 1;
            int rc = conv_dispatch(c, in_mode, a);
            if (rc) return rc;
        }
        {
            const LayerRef& L = M.conv[3];
            ConvArgs a = {};
            a.in = c->ws_c1.p; a.w = w + L.k; a.bias = w + L.b; a.out = c->ws_c2.as<float>();
            a.ktab = M.ktab[3].as<int32_t>();
            a.M = nb * M.H2 * M.W2; a.in_site_stride = c1_site;
            a.K = L.KH * L.KW * L.Cin; a.Hin = M.Hin; a.Win = M.Win; a.Cin = 3 * M.C1; a.SW = 2;
            a.PH = 0; a.PW = 0; a.Hout = M.H2; a.Wout = M.W2; a.Cout = M.C2; a.out_cstride = M.C2; a.out_coff = 0; a.selu = 1;
            int rc = conv_dispatch(c, 0, a);
            if (rc) return rc;
        }
        {
            const LayerRef& L = M.conv[4];
            ConvArgs a = {};
            a.in = c->ws_c2.p; a.w = w + L.k; a.bias = w + L.b; a.out = c->ws_c3.as<float>();
            a.ktab = M.ktab[4].as<int32_t>();
            a.M = nb * M.H3 * M.W3; a.in_site_stride = c2_site;
            a.K = L.KH * L.KW * L.Cin; a.Hin = M.H2; a.Win = M.W2; a.Cin = M.C2; a.SW = 2;
            a.PH = 0; a.PW = 0; a.Hout = M.H3; a.Wout = M.W3; a.Cout = M.C3; a.out_cstride = M.C3; a.out_coff = 0; a.selu = 1;
            int rc = conv_dispatch(c, 0, a);
            if (rc) return rc;
        }
        {
            ConvArgs a = {};
            a.in = c->ws_c3.p; a.w = w + M.fc1.k; a.bias = w + M.fc1.b; a.out = c->ws_f1.as<float>();
            a.ktab = nullptr; a.M = nb; a.in_site_stride = M.flat;
            a.K = M.flat; a.Hin = 1; a.Win = 1; a.Cin = M.flat; a.SW = 1; a.PH = 0; a.PW = 0; a.Hout = 1; a.Wout = 1;
            a.Cout = M.F1; a.out_cstride = M.F1; a.out_coff = 0; a.selu = 1;
            int rc = conv_dispatch(c, 0, a);
            if (rc) return rc;
        }
        const unsigned tg = (unsigned)div_up(nb, 128);
        const float* f1 = c->ws_f1.as<float>();
        const NcSiteMeta* mb = meta ? meta + b0 : nullptr;
        const float* rb = ref4 ? ref4 + b0 * 4 : nullptr;
        if (M.kind == 0) snp_tail_kernel<<<tg, 128, 0, c->stream>>>(f1, nb, tw, mb, rb, out_full ? out_full + b0 * 10 : nullptr, probs ? probs + b0 * 4 : nullptr);
        else if (M.kind == 1) snp_hap_tail_kernel<<<tg, 128, 0, c->stream>>>(f1, nb, tw, mb, rb, probs + b0 * 4);
        else if (M.kind == 2) indel_tail_kernel<false><<<tg, 128, 0, c->stream>>>(f1, nb, tw, out_full + b0 * 4);
        else indel_tail_kernel<true><<<tg, 128, 0, c->stream>>>(f1, nb, tw, out_full + b0);
        NC_LAUNCH_CHECK();
    }
    return NC_OK;
}

int cnn_forward(nc_ctx* c, Model& M, int impl, int in_mode, const void* in_dev, int64_t in_site_stride, int64_t n,
                const NcSiteMeta* meta, const float* ref4, const float* scale_f, const double* scale_d,
                float* out_full, float* probs) {
    if (impl == 1)
        return cnn_forward_f32(c, M, in_mode, in_dev, in_site_stride, n, meta, ref4, scale_f, scale_d, out_full, probs);
    if (impl != 0) return fail(c, NC_EINVAL, "impl must be 0 (tcgen05) or 1 (fp32 CUDA cores)");
    if (!M.tc.ready) return fail(c, NC_ESTATE, "no tensor-core operand image for model kind %d", M.kind);
    uint64_t launches = 0;
    if (M.kind >= 2) {
        if (in_mode != 0) return fail(c, NC_EINVAL, "the indel models read fp32 tensors");
        int rci = tci_forward(c->stream, M.tc, reinterpret_cast<const float*>(in_dev), in_site_stride, n, tail_weights(M), out_full, c->sm_count,
                              &launches, &c->err, 0);
        c->launches += launches;
        return rci;
    }
    int rc = tc_forward_ex(c->stream, M.tc, in_mode, in_dev, in_site_stride, n, meta, ref4, scale_f, scale_d, tail_weights(M),
                           out_full, probs, c->sm_count, &launches, &c->err, 0, c->ev[12]);
    c->launches += launches;
    return rc;
}

// mbarrier waits in the tensor-core kernels are bounded; a timeout raises this flag instead of hanging the device
int tc_check(nc_ctx* c, Model& M) {
    if (!M.tc.ready) return NC_OK;
    NC_CUDA(c->pin.reserve(64));
    NC_CUDA(cudaMemcpyAsync(c->pin.p, M.tc.err.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    NC_CUDA(nc_stream_wait(c));
    if (*c->pin.as<int>() != 0) return fail(c, NC_ECUDA, "tensor-core CNN kernel: mbarrier wait timed out");
    return NC_OK;
}

}  // namespace

namespace {
struct MapRO {
    const uint8_t* p = nullptr; size_t n = 0; bool ok = false;
    explicit MapRO(const char* path) {
        const int fd = open(path, O_RDONLY);
        if (fd < 0) return;
        struct stat st;
        if (fstat(fd, &st) == 0) {
            n = (size_t)st.st_size;
            void* m = n ? mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;
            if (!n || m != MAP_FAILED) { p = (const uint8_t*)m; ok = true; }
        }
        close(fd);
    }
    ~MapRO() { if (p) munmap((void*)p, n); }
};
template <class T> T rd_le(const uint8_t* p) { T v; memcpy(&v, p, sizeof(T)); return v; }
double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}  // namespace

// ================================================================================================
// C-ABI
// ================================================================================================
extern "C" {

int nc_abi_version(void) { return NC_ABI_VERSION; }

int nc_set_blocking_sync(nc_ctx* c, int on) {
    if (!c) return NC_EINVAL;
    NC_CUDA(cudaSetDevice(c->device));
    if (on && !c->ev_block) { NC_CUDA(cudaEventCreateWithFlags(&c->ev_block, cudaEventBlockingSync | cudaEventDisableTiming)); }
    if (!on && c->ev_block) { NC_CUDA(nc_stream_wait(c)); cudaEventDestroy(c->ev_block); c->ev_block = nullptr; }
    return NC_OK;
}

int nc_create(int device, nc_ctx** out) {
    if (!out) return NC_EINVAL;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return NC_ECUDA;
    if (cudaSetDevice(device) != cudaSuccess) return NC_ECUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return NC_ECUDA;
    if (prop.major != 10) return NC_ECUDA;          // sm_100a cubin only: no fallback path exists
    nc_ctx* c = new (std::nothrow) nc_ctx();
    if (!c) return NC_ENOMEM;
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return NC_ECUDA; }
    for (auto& e : c->ev)
        if (cudaEventCreate(&e) != cudaSuccess) { delete c; return NC_ECUDA; }
    for (auto& e : c->evi)
        if (cudaEventCreate(&e) != cudaSuccess) { delete c; return NC_ECUDA; }
    const char* bs = getenv("NC_BLOCKING_SYNC");
    if (bs && bs[0] == '1' && cudaEventCreateWithFlags(&c->ev_block, cudaEventBlockingSync | cudaEventDisableTiming) != cudaSuccess) { delete c; return NC_ECUDA; }
    *out = c;
    return NC_OK;
}

void nc_destroy(nc_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    nc_stream_wait(c);
    DevBuf* bufs[] = {&c->d_pos, &c->d_flag, &c->d_cigar_off, &c->d_cigar, &c->d_seq_off, &c->d_lseq, &c->d_seq4, &c->d_ref, &c->d_fill_counter, &c->d_seqc,
                      &c->d_end, &c->d_nwords, &c->d_opstart, &c->d_pmaxend, &c->d_rowoff, &c->d_rows, &c->d_flags,
                      &c->d_tile_nbr, &c->d_tile_cand, &c->d_nbr_off, &c->d_cand_off, &c->d_nbr_pos, &c->d_cand_pos, &c->d_bed,
                      &c->d_nfirst, &c->d_nlen, &c->d_nbytes, &c->d_noff, &c->d_nrows, &c->d_chunks, &c->d_chunk_lo,
                      &c->d_chunk_cnt, &c->d_chunk_off, &c->d_keep, &c->d_keep32, &c->d_outidx, &c->d_mat, &c->d_meta,
                      &c->d_depth_sum, &c->d_depth_cnt, &c->d_chunk_depth, &c->d_chunk_count, &c->d_probs, &c->d_scan_partial,
                      &c->d_hp, &c->d_ps, &c->d_idepth, &c->d_em, &c->d_grank, &c->d_empos, &c->d_ichunks, &c->d_nem1, &c->d_rankoff, &c->d_diff,
                      &c->d_empairs, &c->d_hit, &c->d_variants, &c->d_icount, &c->d_isites, &c->d_site_m, &c->d_site_cnt, &c->d_site_off, &c->d_eread,
                      &c->d_eqpn, &c->d_eslice, &c->d_eacode, &c->d_einslen, &c->d_einsfirst, &c->d_en, &c->d_itensors, &c->d_icns, &c->d_imeta,
                      &c->d_cdel, &c->d_cins, &c->d_imp_g, &c->d_imp_cols, &c->d_imp_cnt, &c->d_imp_off, &c->d_imp_read, &c->d_imp_ch, &c->d_imp_ind,
                      &c->d_imp_qn, &c->d_imp_gid, &c->d_imp_rep, &c->d_imp_gcnt, &c->d_imp_label, &c->d_imp_ok, &c->d_site_imp, &c->d_egrp,
                      &c->ws_c1, &c->ws_c2, &c->ws_c3, &c->ws_f1, &c->ws_sf, &c->ws_sd, &c->ws_x, &c->ws_ref, &c->ws_out};
    for (DevBuf* b : bufs) b->release();
    for (Model* m : {&c->snp[0], &c->snp[1], &c->indel[0], &c->indel[1]}) {
        m->w.release();
        for (auto& t : m->ktab) t.release();
        tc_model_release(m->tc);
    }
    c->pin.release();
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    for (auto& e : c->evi) if (e) cudaEventDestroy(e);
    c->d_iprobs.release(); c->d_ialleles.release(); c->d_allele_dirs.release(); c->d_allele_ops.release();
    for (DevBuf* b : {&c->d_bam_comp, &c->d_bam_blocks, &c->d_bam, &c->d_bam_recoff, &c->d_bam_rid, &c->d_bam_pos, &c->d_bam_flag, &c->d_bam_lseq, &c->d_bam_ncig,
                      &c->d_bam_nseq, &c->d_bam_cigsrc, &c->d_bam_seqsrc, &c->d_bam_hp, &c->d_bam_ps, &c->d_bam_err, &c->d_bam_out, &c->d_bam_wst, &c->d_bam_wcnt,
                      &c->d_bam_woff}) b->release();
    c->pin_bam.release();
    if (c->ev_block) cudaEventDestroy(c->ev_block);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char* nc_last_error(const nc_ctx* c) { return c ? c->err.c_str() : "null context"; }

int nc_sync(nc_ctx* c) {
    if (!c) return NC_EINVAL;
    NC_CUDA(cudaSetDevice(c->device));
    NC_CUDA(nc_stream_wait(c));
    return NC_OK;
}

int nc_device_sm_count(nc_ctx* c) { return c ? c->sm_count : NC_EINVAL; }

int nc_get_timings(nc_ctx* c, NcTimings* out) {
    if (!c || !out) return NC_EINVAL;
    NC_CUDA(cudaSetDevice(c->device));
    NC_CUDA(nc_stream_wait(c));
    if (c->tm_decode) { NC_CUDA(cudaEventElapsedTime(&c->tm.decode_ms, c->ev[0], c->ev[1])); }
    if (c->tm_scan) {
        NC_CUDA(cudaEventElapsedTime(&c->tm.scan_ms, c->ev[2], c->ev[3]));
        NC_CUDA(cudaEventElapsedTime(&c->tm.tensor_ms, c->ev[4], c->ev[5]));
    }
    if (c->tm_cnn) { NC_CUDA(cudaEventElapsedTime(&c->tm.cnn_ms, c->ev[6], c->ev[7])); }
    c->tm.cnn_a_ms = 0.f;
    if (c->tm_cnn && c->tm_cnn_a) { NC_CUDA(cudaEventElapsedTime(&c->tm.cnn_a_ms, c->ev[6], c->ev[12])); }
    c->tm.launches = c->launches;
    *out = c->tm;
    return NC_OK;
}

int nc_event_record(nc_ctx* c, int slot) {
    if (!c || slot < 0 || slot > 3) return fail(c, NC_EINVAL, "nc_event_record: slot must be 0..3");
    NC_CUDA(cudaSetDevice(c->device));
    NC_CUDA(cudaEventRecord(c->ev[8 + slot], c->stream));
    return NC_OK;
}

int nc_event_elapsed_ms(nc_ctx* c, int a, int b, float* ms) {
    if (!c || !ms || a < 0 || a > 3 || b < 0 || b > 3) return fail(c, NC_EINVAL, "nc_event_elapsed_ms: bad argument");
    NC_CUDA(cudaSetDevice(c->device));
    NC_CUDA(cudaEventSynchronize(c->ev[8 + b]));
    NC_CUDA(cudaEventElapsedTime(ms, c->ev[8 + a], c->ev[8 + b]));
    return NC_OK;
}

int nc_invalidate_decode(nc_ctx* c) {
    if (!c) return NC_EINVAL;
    c->decoded = false; c->scanned = false; c->have_probs = false;
    return NC_OK;
}

int nc_stage_reads(nc_ctx* c, int64_t n_reads, const int32_t* pos, const uint16_t* flag, const int64_t* cigar_off,
                   const uint32_t* cigar, const int64_t* seq_off, const int32_t* l_seq, const uint8_t* seq4,
                   const uint8_t* ref, int64_t ref_start, int64_t ref_len) {
    if (!c) return NC_EINVAL;
    NC_CUDA(cudaSetDevice(c->device));
    c->staged = c->decoded = c->scanned = false;
    c->have_probs = false; c->tags_staged = c->indel_scanned = c->indel_built = false;
    if (n_reads < 0 || ref_len < 0 || ref_start < 0 || (n_reads > 0 && (!pos || !flag || !cigar_off || !seq_off || !l_seq)) || (ref_len > 0 && !ref))
        return fail(c, NC_EINVAL, "nc_stage_reads: null or negative argument");
    if (ref_start + ref_len > 0x7fffff00ll) return fail(c, NC_EOVERFLOW, "contig coordinates must fit 31 bits");
    int64_t n_cig = 0, n_seq = 0;
    if (n_reads > 0) {
        if (cigar_off[0] != 0 || seq_off[0] != 0) return fail(c, NC_EINVAL, "offset arrays must start at 0");
        for (int64_t i = 0; i < n_reads; i++) {
            if (i && pos[i] < pos[i - 1]) return fail(c, NC_EINVAL, "reads must be coordinate-sorted (read %lld)", (long long)i);
            if (cigar_off[i + 1] < cigar_off[i] || seq_off[i + 1] < seq_off[i] || l_seq[i] < 0 ||
                (int64_t)(l_seq[i] + 1) / 2 > seq_off[i + 1] - seq_off[i])
                return fail(c, NC_EINVAL, "inconsistent offsets at read %lld", (long long)i);
        }
        n_cig = cigar_off[n_reads]; n_seq = seq_off[n_reads];
        if ((n_cig > 0 && !cigar) || (n_seq > 0 && !seq4)) return fail(c, NC_EINVAL, "nc_stage_reads: null payload");
    }
    int rc;
    static const int64_t zero = 0;
    if ((rc = upload(c, c->d_pos, pos, (size_t)n_reads * 4))) return rc;
    if ((rc = upload(c, c->d_flag, flag, (size_t)n_reads * 2))) return rc;
    if ((rc = upload(c, c->d_cigar_off, n_reads ? cigar_off : &zero, (size_t)(n_reads + 1) * 8))) return rc;
    if ((rc = upload(c, c->d_cigar, cigar, (size_t)n_cig * 4))) return rc;
    if ((rc = upload(c, c->d_seq_off, n_reads ? seq_off : &zero, (size_t)(n_reads + 1) * 8))) return rc;
    if ((rc = upload(c, c->d_lseq, l_seq, (size_t)n_reads * 4))) return rc;
    if ((rc = upload(c, c->d_seq4, seq4, (size_t)n_seq))) return rc;
    if ((rc = upload(c, c->d_ref, ref, (size_t)ref_len))) return rc;
    c->n_reads = n_reads; c->n_cigar = n_cig; c->n_seq = n_seq; c->ref_start = ref_start; c->ref_len = ref_len;
    c->staged = true;
    return NC_OK;
}

int nc_decode_reads(nc_ctx* c) {
    if (!c) return NC_EINVAL;
    if (!c->staged) return fail(c, NC_ESTATE, "nc_decode_reads before nc_stage_reads");
    if (c->decoded) return NC_OK;
    NC_CUDA(cudaSetDevice(c->device));
    const int64_t n = c->n_reads;
    NC_CUDA(cudaEventRecord(c->ev[0], c->stream));
    NC_CUDA(c->d_end.reserve((size_t)std::max<int64_t>(n, 1) * 4));
    NC_CUDA(c->d_nwords.reserve((size_t)std::max<int64_t>(n, 1) * 4));
    NC_CUDA(c->d_pmaxend.reserve((size_t)std::max<int64_t>(n, 1) * 4));
    NC_CUDA(c->d_opstart.reserve((size_t)std::max<int64_t>(c->n_cigar, 1) * 8));
    NC_CUDA(c->d_rowoff.reserve((size_t)(n + 1) * 8));
    c->n_row_words = 0;
    if (n > 0) {
        const unsigned grid = (unsigned)std::min<int64_t>(div_up(n * 32, 256), (int64_t)c->sm_count * 32);
        cigar_scan_kernel<<<grid, 256, 0, c->stream>>>(n, c->d_pos.as<int32_t>(), c->d_cigar_off.as<int64_t>(), c->d_cigar.as<uint32_t>(),
                                                       c->d_end.as<int32_t>(), c->d_nwords.as<int32_t>(), c->d_opstart.as<int2>());
        NC_LAUNCH_CHECK();
        prefix_max_kernel<<<1, 1024, 0, c->stream>>>(c->d_end.as<int32_t>(), n, c->d_pmaxend.as<int32_t>());
        NC_LAUNCH_CHECK();
        int rc = device_scan(c, c->d_nwords.as<int32_t>(), n, c->d_rowoff.as<int64_t>());
        if (rc) return rc;
        if ((rc = read_i64(c, c->d_rowoff.as<int64_t>() + n, &c->n_row_words))) return rc;
        NC_CUDA(c->d_rows.reserve((size_t)std::max<int64_t>(c->n_row_words, 1) * 4));
        NC_CUDA(c->d_fill_counter.reserve(8));
        NC_CUDA(cudaMemsetAsync(c->d_fill_counter.p, 0, 8, c->stream));
        const int64_t n_vec = div_up(c->n_seq, 16);
        NC_CUDA(c->d_seqc.reserve((size_t)n_vec * 16 + 64));
        if (n_vec > 0) {
            seq_codes_kernel<<<(unsigned)std::min<int64_t>(div_up(n_vec, 256), (int64_t)c->sm_count * 16), 256, 0, c->stream>>>(
                c->d_seq4.as<uint4>(), n_vec, c->d_seqc.as<uint4>());
            NC_LAUNCH_CHECK();
        }
        const unsigned g2 = (unsigned)std::min<int64_t>(n, (int64_t)c->sm_count * 12);      // persistent CTAs, reads handed out dynamically
        row_fill_kernel<<<g2, kFillThreads, 0, c->stream>>>(c->d_pos.as<int32_t>(), c->d_end.as<int32_t>(), c->d_cigar_off.as<int64_t>(),
                                                   c->d_cigar.as<uint32_t>(), c->d_opstart.as<int2>(), c->d_seq_off.as<int64_t>(),
                                                   c->d_lseq.as<int32_t>(), c->d_seqc.as<uint32_t>(), c->d_rowoff.as<int64_t>(),
                                                   c->d_nwords.as<int32_t>(), c->d_rows.as<uint32_t>(), n,
                                                   c->d_fill_counter.as<unsigned long long>());
        NC_LAUNCH_CHECK();
    }
    NC_CUDA(cudaEventRecord(c->ev[1], c->stream));
    c->tm_decode = true;
    // SURVEY 8(d) B1, read side: BAM-native CIGAR words + 4-bit bases + a 16-byte header per read
    c->decode_bytes = (uint64_t)c->n_cigar * 4 + (uint64_t)c->n_seq + (uint64_t)n * 16;
    c->tm.scan_bytes = c->decode_bytes;
    c->decoded = true;
    return NC_OK;
}

int nc_snp_scan(nc_ctx* c, const NcSnpParams* P, const NcChunk* chunks, int32_t n_chunks, const int32_t* bed,
                int32_t n_bed, int64_t* n_sites_out) {
    if (!c || !P || !n_sites_out || n_chunks < 0 || (n_chunks > 0 && !chunks) || n_bed < 0 || (n_bed > 0 && !bed))
        return fail(c, NC_EINVAL, "nc_snp_scan: bad argument");
    *n_sites_out = 0;
    if (!c->staged) return fail(c, NC_ESTATE, "nc_snp_scan before nc_stage_reads");
    if (P->seq < 0 || P->seq > 4) return fail(c, NC_EINVAL, "unknown sequencing mode %d", P->seq);
    if (P->maxcov < 1 || P->maxcov > 32767) return fail(c, NC_EOVERFLOW, "maxcov %d does not fit the int16 tensor", P->maxcov);
    NC_CUDA(cudaSetDevice(c->device));
    int rc = nc_decode_reads(c);
    if (rc) return rc;
    c->scanned = false; c->have_probs = false; c->scan_haploid = P->haploid ? 1 : 0;
    c->n_sites = c->n_slots = c->n_nbr = c->n_cand = 0; c->n_chunks = n_chunks;

    // scanned range = union of the chunks' pileup windows (generate_SNP_pileups.py:156), clipped to the staged reference
    int64_t lo = INT64_MAX, hi = INT64_MIN;
    for (int i = 0; i < n_chunks; i++) {
        if (chunks[i].end < chunks[i].start) continue;
        lo = std::min<int64_t>(lo, std::max<int64_t>(0, (int64_t)chunks[i].start - 1 - 50000));
        hi = std::max<int64_t>(hi, (int64_t)chunks[i].end + 50000);
    }
    lo = std::max(lo, c->ref_start); hi = std::min(hi, c->ref_start + c->ref_len);
    NC_CUDA(c->d_chunk_depth.reserve((size_t)std::max(n_chunks, 1) * 8));
    NC_CUDA(c->d_chunk_count.reserve((size_t)std::max(n_chunks, 1) * 8));
    NC_CUDA(cudaMemsetAsync(c->d_chunk_depth.p, 0, (size_t)std::max(n_chunks, 1) * 8, c->stream));
    NC_CUDA(cudaMemsetAsync(c->d_chunk_count.p, 0, (size_t)std::max(n_chunks, 1) * 8, c->stream));
    NC_CUDA(cudaEventRecord(c->ev[2], c->stream));
    NC_CUDA(cudaEventRecord(c->ev[3], c->stream));
    NC_CUDA(cudaEventRecord(c->ev[4], c->stream));
    NC_CUDA(cudaEventRecord(c->ev[5], c->stream));
    c->tm_scan = true; c->tm.tensor_bytes = 0;
    if (n_chunks == 0 || hi <= lo || c->n_reads == 0) { c->scanned = true; return NC_OK; }

    // exclude BED: sort + merge so that "start <= v < end for some interval" is a binary search.
    // (The reference's IntervalTree raises on null intervals, which disables exclusion: the host mirror handles that.)
    std::vector<std::pair<int32_t, int32_t>> iv;
    for (int i = 0; i < n_bed; i++) if (bed[2 * i + 1] > bed[2 * i]) iv.emplace_back(bed[2 * i], bed[2 * i + 1]);
    std::sort(iv.begin(), iv.end());
    std::vector<int32_t> merged;
    for (auto& x : iv) {
        if (!merged.empty() && x.first <= merged[merged.size() - 1]) merged[merged.size() - 1] = std::max(merged[merged.size() - 1], x.second);
        else { merged.push_back(x.first); merged.push_back(x.second); }
    }
    const int32_t n_merged = (int32_t)(merged.size() / 2);
    if ((rc = upload(c, c->d_bed, merged.data(), merged.size() * 4))) return rc;
    if ((rc = upload(c, c->d_chunks, chunks, (size_t)n_chunks * sizeof(NcChunk)))) return rc;
    NC_CUDA(nc_stream_wait(c));     // host vectors above are pageable

    const uint32_t flag_filter = P->supplementary ? 0x704u : 0xF04u;
    const int32_t lo_al = (int32_t)lo & ~7;
    const int64_t n_tiles = div_up(hi - lo_al, kTilePos);
    NC_CUDA(c->d_flags.reserve((size_t)n_tiles * kTilePos));
    NC_CUDA(c->d_tile_nbr.reserve((size_t)n_tiles * 4));
    NC_CUDA(c->d_tile_cand.reserve((size_t)n_tiles * 4));
    NC_CUDA(c->d_nbr_off.reserve((size_t)(n_tiles + 1) * 8));
    NC_CUDA(c->d_cand_off.reserve((size_t)(n_tiles + 1) * 8));

    NC_CUDA(cudaEventRecord(c->ev[2], c->stream));
    ScanArgs sa = {};
    sa.n_reads = c->n_reads; sa.pos = c->d_pos.as<int32_t>(); sa.end = c->d_end.as<int32_t>(); sa.flag = c->d_flag.as<uint16_t>();
    sa.pmaxend = c->d_pmaxend.as<int32_t>(); sa.rowoff = c->d_rowoff.as<int64_t>(); sa.nwords = c->d_nwords.as<int32_t>();
    sa.rows = c->d_rows.as<uint32_t>(); sa.ref = c->d_ref.as<uint8_t>(); sa.ref_start = c->ref_start; sa.ref_len = c->ref_len;
    sa.lo_al = lo_al; sa.lo = (int32_t)lo; sa.hi = (int32_t)hi; sa.mincov = P->mincov; sa.haploid = P->haploid;
    sa.thr_lo = P->thr_lo; sa.thr_hi = P->thr_hi; sa.maf = P->min_allele_freq; sa.flag_filter = flag_filter;
    sa.bed = c->d_bed.as<int32_t>(); sa.n_bed = n_merged;
    sa.flags = c->d_flags.as<uint8_t>(); sa.tile_nbr = c->d_tile_nbr.as<int32_t>(); sa.tile_cand = c->d_tile_cand.as<int32_t>();
    scan_kernel<<<(unsigned)n_tiles, kTileThreads, 0, c->stream>>>(sa); NC_LAUNCH_CHECK();
    if ((rc = device_scan(c, sa.tile_nbr, n_tiles, c->d_nbr_off.as<int64_t>()))) return rc;
    if ((rc = device_scan(c, sa.tile_cand, n_tiles, c->d_cand_off.as<int64_t>()))) return rc;
    if ((rc = read_i64(c, c->d_nbr_off.as<int64_t>() + n_tiles, &c->n_nbr))) return rc;
    if ((rc = read_i64(c, c->d_cand_off.as<int64_t>() + n_tiles, &c->n_cand))) return rc;
    NC_CUDA(c->d_nbr_pos.reserve((size_t)std::max<int64_t>(c->n_nbr, 1) * 4));
    NC_CUDA(c->d_cand_pos.reserve((size_t)std::max<int64_t>(c->n_cand, 1) * 4));
    site_list_kernel<<<(unsigned)n_tiles, kTileThreads, 0, c->stream>>>(c->d_flags.as<uint8_t>(), lo_al, c->d_nbr_off.as<int64_t>(),
                                                                        c->d_cand_off.as<int64_t>(), c->d_nbr_pos.as<int32_t>(),
                                                                        c->d_cand_pos.as<int32_t>());
    NC_LAUNCH_CHECK();

    // neighbour matrix
    NC_CUDA(c->d_nfirst.reserve((size_t)c->n_reads * 4));
    NC_CUDA(c->d_nlen.reserve((size_t)c->n_reads * 4));
    NC_CUDA(c->d_nbytes.reserve((size_t)c->n_reads * 4));
    NC_CUDA(c->d_noff.reserve((size_t)(c->n_reads + 1) * 8));
    const unsigned rg = (unsigned)div_up(c->n_reads, 256);
    nmat_len_kernel<<<rg, 256, 0, c->stream>>>(c->n_reads, c->d_pos.as<int32_t>(), c->d_end.as<int32_t>(), c->d_nbr_pos.as<int32_t>(),
                                               (int32_t)c->n_nbr, c->d_nfirst.as<int32_t>(), c->d_nlen.as<int32_t>(), c->d_nbytes.as<int32_t>());
    NC_LAUNCH_CHECK();
    if ((rc = device_scan(c, c->d_nbytes.as<int32_t>(), c->n_reads, c->d_noff.as<int64_t>()))) return rc;
    int64_t n_nbytes = 0;
    if ((rc = read_i64(c, c->d_noff.as<int64_t>() + c->n_reads, &n_nbytes))) return rc;
    NC_CUDA(c->d_nrows.reserve((size_t)std::max<int64_t>(n_nbytes, 16)));
    // SURVEY 8(d) B1 = reads (CIGAR + bases + header) + one reference byte per piled position + 5 B per (read, kept site):
    // the neighbour matrix holds one nibble per (read, neighbour site), i.e. 2 * n_nbytes pairs; the host adds the candidates' reads
    c->tm.scan_bytes = c->decode_bytes + (uint64_t)(hi - lo) + (uint64_t)10 * (uint64_t)n_nbytes;
    nmat_fill_kernel<<<rg, 256, 0, c->stream>>>(c->n_reads, c->d_pos.as<int32_t>(), c->d_rowoff.as<int64_t>(), c->d_rows.as<uint32_t>(),
                                                c->d_nbr_pos.as<int32_t>(), c->d_nfirst.as<int32_t>(), c->d_nlen.as<int32_t>(),
                                                c->d_noff.as<int64_t>(), c->d_nrows.as<uint8_t>());
    NC_LAUNCH_CHECK();

    // chunk cut
    NC_CUDA(c->d_chunk_lo.reserve((size_t)n_chunks * 4));
    NC_CUDA(c->d_chunk_cnt.reserve((size_t)n_chunks * 4));
    NC_CUDA(c->d_chunk_off.reserve((size_t)(n_chunks + 1) * 8));
    chunk_ranges_kernel<<<(unsigned)div_up(n_chunks, 128), 128, 0, c->stream>>>(c->d_chunks.as<NcChunk>(), n_chunks, c->d_cand_pos.as<int32_t>(),
                                                                                c->n_cand, c->d_chunk_lo.as<int32_t>(), c->d_chunk_cnt.as<int32_t>());
    NC_LAUNCH_CHECK();
    if ((rc = device_scan(c, c->d_chunk_cnt.as<int32_t>(), n_chunks, c->d_chunk_off.as<int64_t>()))) return rc;
    if ((rc = read_i64(c, c->d_chunk_off.as<int64_t>() + n_chunks, &c->n_slots))) return rc;
    NC_CUDA(cudaEventRecord(c->ev[3], c->stream));

    TensorArgs ta = {};
    ta.n_reads = c->n_reads; ta.pos = sa.pos; ta.end = sa.end; ta.flag = sa.flag; ta.pmaxend = sa.pmaxend;
    ta.rowoff = sa.rowoff; ta.rows = sa.rows;
    ta.nfirst = c->d_nfirst.as<int32_t>(); ta.nlen = c->d_nlen.as<int32_t>(); ta.noff = c->d_noff.as<int64_t>(); ta.nrows = c->d_nrows.as<uint8_t>();
    ta.nbr_pos = c->d_nbr_pos.as<int32_t>(); ta.n_nbr = (int32_t)c->n_nbr; ta.cand_pos = c->d_cand_pos.as<int32_t>();
    ta.chunks = c->d_chunks.as<NcChunk>(); ta.n_chunks = n_chunks; ta.chunk_off = c->d_chunk_off.as<int64_t>();
    ta.chunk_lo = c->d_chunk_lo.as<int32_t>(); ta.outidx = nullptr;
    ta.ref = sa.ref; ta.ref_start = c->ref_start; ta.ref_len = c->ref_len;
    ta.seq = P->seq; ta.maxcov = P->maxcov; ta.min_nbr_sites = P->min_nbr_sites; ta.flag_filter = flag_filter;
    ta.n_slots = c->n_slots;

    NC_CUDA(cudaEventRecord(c->ev[4], c->stream));
    c->n_sites = c->n_slots;
    if (c->n_slots > 0 && P->min_nbr_sites > 1) {
        NC_CUDA(c->d_keep.reserve((size_t)c->n_slots));
        NC_CUDA(c->d_keep32.reserve((size_t)c->n_slots * 4));
        NC_CUDA(c->d_outidx.reserve((size_t)(c->n_slots + 1) * 8));
        ta.keep = c->d_keep.as<uint8_t>();
        keep_kernel<<<(unsigned)div_up(c->n_slots * 32, 128), 128, 0, c->stream>>>(ta); NC_LAUNCH_CHECK();
        keep_to_i32_kernel<<<(unsigned)div_up(c->n_slots, 256), 256, 0, c->stream>>>(ta.keep, c->n_slots, c->d_keep32.as<int32_t>()); NC_LAUNCH_CHECK();
        if ((rc = device_scan(c, c->d_keep32.as<int32_t>(), c->n_slots, c->d_outidx.as<int64_t>()))) return rc;
        if ((rc = read_i64(c, c->d_outidx.as<int64_t>() + c->n_slots, &c->n_sites))) return rc;
        ta.outidx = c->d_outidx.as<int64_t>();
    }
    NC_CUDA(c->d_depth_sum.reserve((size_t)n_chunks * 8));
    NC_CUDA(c->d_depth_cnt.reserve((size_t)n_chunks * 8));
    NC_CUDA(cudaMemsetAsync(c->d_depth_sum.p, 0, (size_t)n_chunks * 8, c->stream));
    NC_CUDA(cudaMemsetAsync(c->d_depth_cnt.p, 0, (size_t)n_chunks * 8, c->stream));
    if (c->n_sites > 0) {
        NC_CUDA(c->d_mat.reserve((size_t)c->n_sites * NC_SNP_SITE_STRIDE * 2));
        NC_CUDA(c->d_meta.reserve((size_t)c->n_sites * sizeof(NcSiteMeta)));
        ta.mat = c->d_mat.as<int16_t>(); ta.meta = c->d_meta.as<NcSiteMeta>();
        ta.chunk_depth_sum = c->d_depth_sum.as<unsigned long long>(); ta.chunk_count = c->d_depth_cnt.as<unsigned long long>();
        const unsigned tg = (unsigned)std::min<int64_t>(div_up(c->n_slots, kRunSlots), (int64_t)c->sm_count * 16);
        tensor_kernel<<<tg, kTensorWarps * 32, 0, c->stream>>>(ta); NC_LAUNCH_CHECK();
    }
    chunk_depth_kernel<<<(unsigned)div_up(n_chunks, 128), 128, 0, c->stream>>>(c->d_depth_sum.as<unsigned long long>(), c->d_depth_cnt.as<unsigned long long>(),
                                                                               n_chunks, c->d_chunk_depth.as<double>(), c->d_chunk_count.as<int64_t>());
    NC_LAUNCH_CHECK();
    NC_CUDA(cudaEventRecord(c->ev[5], c->stream));
    c->tm.tensor_bytes = (uint64_t)c->n_sites * (NC_SNP_SITE_STRIDE * 2 + sizeof(NcSiteMeta));   // + per-read terms added by the host from meta
    c->scanned = true;
    *n_sites_out = c->n_sites;
    return NC_OK;
}

int nc_snp_fetch(nc_ctx* c, int16_t* mat, NcSiteMeta* meta, double* chunk_depth, int64_t* chunk_count) {
    if (!c) return NC_EINVAL;
    if (!c->scanned) return fail(c, NC_ESTATE, "nc_snp_fetch before nc_snp_scan");
    NC_CUDA(cudaSetDevice(c->device));
    if (mat && c->n_sites) NC_CUDA(cudaMemcpyAsync(mat, c->d_mat.p, (size_t)c->n_sites * NC_SNP_SITE_STRIDE * 2, cudaMemcpyDeviceToHost, c->stream));
    if (meta && c->n_sites) NC_CUDA(cudaMemcpyAsync(meta, c->d_meta.p, (size_t)c->n_sites * sizeof(NcSiteMeta), cudaMemcpyDeviceToHost, c->stream));
    if (chunk_depth && c->n_chunks) NC_CUDA(cudaMemcpyAsync(chunk_depth, c->d_chunk_depth.p, (size_t)c->n_chunks * 8, cudaMemcpyDeviceToHost, c->stream));
    if (chunk_count && c->n_chunks) NC_CUDA(cudaMemcpyAsync(chunk_count, c->d_chunk_count.p, (size_t)c->n_chunks * 8, cudaMemcpyDeviceToHost, c->stream));
    NC_CUDA(nc_stream_wait(c));
    return NC_OK;
}

int nc_snp_fetch_range(nc_ctx* c, int64_t first, int64_t count, int16_t* mat) {
    if (!c) return NC_EINVAL;
    if (!c->scanned) return fail(c, NC_ESTATE, "nc_snp_fetch_range before nc_snp_scan");
    if (first < 0 || count < 0 || first + count > c->n_sites || (count > 0 && !mat)) return fail(c, NC_EINVAL, "nc_snp_fetch_range: bad range");
    NC_CUDA(cudaSetDevice(c->device));
    if (count) NC_CUDA(cudaMemcpyAsync(mat, c->d_mat.as<int16_t>() + first * NC_SNP_SITE_STRIDE, (size_t)count * NC_SNP_SITE_STRIDE * 2, cudaMemcpyDeviceToHost, c->stream));
    NC_CUDA(nc_stream_wait(c));
    return NC_OK;
}

int nc_load_snp_weights(nc_ctx* c, const float* blob, size_t n_floats, double train_coverage, int haploid) {
    if (!c || !blob) return fail(c, NC_EINVAL, "nc_load_snp_weights: null argument");
    NC_CUDA(cudaSetDevice(c->device));
    return model_init(c, c->snp[haploid ? 1 : 0], haploid ? 1 : 0, blob, n_floats, train_coverage);
}

int nc_load_indel_weights(nc_ctx* c, const float* blob, size_t n_floats, int haploid) {
    if (!c || !blob) return fail(c, NC_EINVAL, "nc_load_indel_weights: null argument");
    NC_CUDA(cudaSetDevice(c->device));
    return model_init(c, c->indel[haploid ? 1 : 0], haploid ? 3 : 2, blob, n_floats, 0.0);
}

int nc_snp_forward(nc_ctx* c, int normalize, int impl, float* probs) {
    if (!c) return NC_EINVAL;
    if (!c->scanned) return fail(c, NC_ESTATE, "nc_snp_forward before nc_snp_scan");
    NC_CUDA(cudaSetDevice(c->device));
    // haploid regions go through hap_snp_model (snpCaller.py:165-183), diploid ones through snp_model (:88-111)
    Model& M = c->snp[c->scan_haploid ? 1 : 0];
    if (!M.loaded) return fail(c, NC_ESTATE, "nc_snp_forward: no %s SNP weights loaded", c->scan_haploid ? "haploid" : "diploid");
    if (normalize && !(M.train_cov > 0)) return fail(c, NC_EINVAL, "coverage normalisation needs a positive train_coverage");
    const int64_t n = c->n_sites;
    NC_CUDA(cudaEventRecord(c->ev[6], c->stream));
    if (n > 0) {
        NC_CUDA(c->d_probs.reserve((size_t)n * 4 * sizeof(float)));
        NC_CUDA(c->ws_sf.reserve((size_t)n * sizeof(float)));
        NC_CUDA(c->ws_sd.reserve((size_t)n * sizeof(double)));
        site_scale_kernel<<<(unsigned)div_up(n, 256), 256, 0, c->stream>>>(c->d_meta.as<NcSiteMeta>(), n, c->d_chunk_depth.as<double>(),
                                                                          M.train_cov, normalize, c->ws_sf.as<float>(), c->ws_sd.as<double>());
        NC_LAUNCH_CHECK();
        int rc = cnn_forward(c, M, impl, normalize ? 1 : 2, c->d_mat.p, NC_SNP_SITE_STRIDE, n, c->d_meta.as<NcSiteMeta>(), nullptr,
                             c->ws_sf.as<float>(), c->ws_sd.as<double>(), nullptr, c->d_probs.as<float>());
        if (rc) return rc;
    }
    NC_CUDA(cudaEventRecord(c->ev[7], c->stream));
    c->tm_cnn = true;
    c->tm_cnn_a = n > 0 && impl == 0 && M.tc.ready;          // ev[12] was recorded after the conv1/conv2 kernel of THIS forward
    c->have_probs = true;
    if (probs && n > 0) {
        NC_CUDA(cudaMemcpyAsync(probs, c->d_probs.p, (size_t)n * 4 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        NC_CUDA(nc_stream_wait(c));
        if (impl == 0) return tc_check(c, M);
    }
    return NC_OK;
}

int nc_snp_fetch_probs(nc_ctx* c, float* probs) {
    if (!c || !probs) return fail(c, NC_EINVAL, "nc_snp_fetch_probs: null argument");
    if (!c->scanned || !c->have_probs) return fail(c, NC_ESTATE, "nc_snp_fetch_probs before nc_snp_forward");
    NC_CUDA(cudaSetDevice(c->device));
    if (c->n_sites) NC_CUDA(cudaMemcpyAsync(probs, c->d_probs.p, (size_t)c->n_sites * 4 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    NC_CUDA(nc_stream_wait(c));
    return tc_check(c, c->snp[c->scan_haploid ? 1 : 0]);
}

int nc_snp_model_forward(nc_ctx* c, const float* x, const float* ref_onehot, int64_t n, int haploid, int impl, float* out) {
    if (!c || n < 0 || (n > 0 && (!x || !ref_onehot || !out))) return fail(c, NC_EINVAL, "nc_snp_model_forward: bad argument");
    NC_CUDA(cudaSetDevice(c->device));
    Model& M = c->snp[haploid ? 1 : 0];
    if (!M.loaded) return fail(c, NC_ESTATE, "nc_snp_model_forward: no %s SNP weights loaded", haploid ? "haploid" : "diploid");
    if (n == 0) return NC_OK;
    const int64_t site = NC_SNP_SITE_ELEMS;
    const int nout = haploid ? 4 : 10;
    NC_CUDA(c->ws_x.reserve((size_t)n * site * 4));
    NC_CUDA(c->ws_ref.reserve((size_t)n * 4 * 4));
    NC_CUDA(c->ws_out.reserve((size_t)n * nout * 4));
    NC_CUDA(cudaMemcpyAsync(c->ws_x.p, x, (size_t)n * site * 4, cudaMemcpyHostToDevice, c->stream));
    NC_CUDA(cudaMemcpyAsync(c->ws_ref.p, ref_onehot, (size_t)n * 16, cudaMemcpyHostToDevice, c->stream));
    int rc = cnn_forward(c, M, impl, 0, c->ws_x.p, site, n, nullptr, c->ws_ref.as<float>(), nullptr, nullptr,
                         haploid ? nullptr : c->ws_out.as<float>(), haploid ? c->ws_out.as<float>() : nullptr);
    if (rc) return rc;
    NC_CUDA(cudaMemcpyAsync(out, c->ws_out.p, (size_t)n * nout * 4, cudaMemcpyDeviceToHost, c->stream));
    NC_CUDA(nc_stream_wait(c));
    return impl == 0 ? tc_check(c, M) : NC_OK;
}

int nc_indel_model_forward(nc_ctx* c, const float* x, int64_t n, int haploid, int impl, float* out) {
    if (!c || n < 0 || (n > 0 && (!x || !out))) return fail(c, NC_EINVAL, "nc_indel_model_forward: bad argument");
    NC_CUDA(cudaSetDevice(c->device));
    Model& M = c->indel[haploid ? 1 : 0];
    if (!M.loaded) return fail(c, NC_ESTATE, "nc_indel_model_forward: no %s indel weights loaded", haploid ? "haploid" : "diploid");
    if (n == 0) return NC_OK;
    const int64_t site = (int64_t)M.Hin * M.Win * M.Cin;
    const int nout = haploid ? 1 : 4;
    NC_CUDA(c->ws_x.reserve((size_t)n * site * 4));
    NC_CUDA(c->ws_out.reserve((size_t)n * nout * 4));
    NC_CUDA(cudaMemcpyAsync(c->ws_x.p, x, (size_t)n * site * 4, cudaMemcpyHostToDevice, c->stream));
    int rc = cnn_forward(c, M, impl, 0, c->ws_x.p, site, n, nullptr, nullptr, nullptr, nullptr, c->ws_out.as<float>(), nullptr);
    if (rc) return rc;
    NC_CUDA(cudaMemcpyAsync(out, c->ws_out.p, (size_t)n * nout * 4, cudaMemcpyDeviceToHost, c->stream));
    NC_CUDA(nc_stream_wait(c));
    return impl == 0 ? tc_check(c, M) : NC_OK;
}

int nc_snp_device_buffers(nc_ctx* c, void** mat_dev, void** meta_dev, void** probs_dev, int64_t* n_sites) {
    if (!c) return NC_EINVAL;
    if (!c->scanned) return fail(c, NC_ESTATE, "nc_snp_device_buffers before nc_snp_scan");
    if (mat_dev) *mat_dev = c->n_sites ? c->d_mat.p : nullptr;
    if (meta_dev) *meta_dev = c->n_sites ? c->d_meta.p : nullptr;
    if (probs_dev) *probs_dev = (c->have_probs && c->n_sites) ? c->d_probs.p : nullptr;
    if (n_sites) *n_sites = c->n_sites;
    return NC_OK;
}

// ---- indel feature path ------------------------------------------------------------------------------------
int nc_stage_tags(nc_ctx* c, const int8_t* hp, const int32_t* ps) {
    if (!c) return NC_EINVAL;
    if (!c->staged) return fail(c, NC_ESTATE, "nc_stage_tags before nc_stage_reads");
    if (c->n_reads > 0 && (!hp || !ps)) return fail(c, NC_EINVAL, "nc_stage_tags: null argument");
    NC_CUDA(cudaSetDevice(c->device));
    int rc;
    if ((rc = upload(c, c->d_hp, hp, (size_t)c->n_reads))) return rc;
    if ((rc = upload(c, c->d_ps, ps, (size_t)c->n_reads * 4))) return rc;
    c->tags_staged = true; c->indel_scanned = c->indel_built = false;
    return NC_OK;
}

// impute_indel_phase (generate_indel_pileups.py:287-300) for the columns listed in c->d_imp_cols: reads of every column in BAM
// order (d_imp_off / d_imp_read), their read-set labels (d_imp_label) and the `both sets >= mincov` flag (d_imp_ok).
static int impute_columns(nc_ctx* c, const NcIndelParams* P, int64_t n_cols) {
    int rc;
    NC_CUDA(c->d_imp_cnt.reserve((size_t)n_cols * 4));
    NC_CUDA(c->d_imp_off.reserve((size_t)(n_cols + 1) * 8));
    NC_CUDA(c->d_imp_ok.reserve((size_t)n_cols * 4));
    ImputeArgs ia = {};
    ia.n_reads = c->n_reads; ia.pos = c->d_pos.as<int32_t>(); ia.end = c->d_end.as<int32_t>(); ia.flag = c->d_flag.as<uint16_t>();
    ia.pmaxend = c->d_pmaxend.as<int32_t>(); ia.cigar_off = c->d_cigar_off.as<int64_t>(); ia.cigar = c->d_cigar.as<uint32_t>();
    ia.opstart = c->d_opstart.as<int2>(); ia.seq_off = c->d_seq_off.as<int64_t>(); ia.l_seq = c->d_lseq.as<int32_t>(); ia.seq4 = c->d_seq4.as<uint8_t>();
    ia.flag_filter = P->supplementary ? 0x704u : 0xF04u; ia.mincov = P->mincov;
    ia.cols = c->d_imp_cols.as<int32_t>(); ia.n_cols = n_cols; ia.col_cnt = c->d_imp_cnt.as<int32_t>(); ia.col_off = c->d_imp_off.as<int64_t>();
    ia.col_ok = c->d_imp_ok.as<int32_t>();
    const unsigned wg = (unsigned)div_up(n_cols * 32, 128);
    indel_impute_reads_kernel<false><<<wg, 128, 0, c->stream>>>(ia); NC_LAUNCH_CHECK();
    if ((rc = device_scan(c, ia.col_cnt, n_cols, c->d_imp_off.as<int64_t>()))) return rc;
    int64_t n_entries = 0;
    if ((rc = read_i64(c, c->d_imp_off.as<int64_t>() + n_cols, &n_entries))) return rc;
    const size_t ne = (size_t)std::max<int64_t>(n_entries, 1);
    NC_CUDA(c->d_imp_read.reserve(ne * 4)); NC_CUDA(c->d_imp_ch.reserve(ne)); NC_CUDA(c->d_imp_ind.reserve(ne * 4)); NC_CUDA(c->d_imp_qn.reserve(ne * 4));
    NC_CUDA(c->d_imp_gid.reserve(ne * 4)); NC_CUDA(c->d_imp_rep.reserve(ne * 4)); NC_CUDA(c->d_imp_gcnt.reserve(ne * 4)); NC_CUDA(c->d_imp_label.reserve(ne));
    ia.e_read = c->d_imp_read.as<int32_t>(); ia.e_ch = c->d_imp_ch.as<uint8_t>(); ia.e_ind = c->d_imp_ind.as<int32_t>(); ia.e_qn = c->d_imp_qn.as<int32_t>();
    ia.e_gid = c->d_imp_gid.as<int32_t>(); ia.g_rep = c->d_imp_rep.as<int32_t>(); ia.g_cnt = c->d_imp_gcnt.as<int32_t>(); ia.e_label = c->d_imp_label.as<int8_t>();
    indel_impute_reads_kernel<true><<<wg, 128, 0, c->stream>>>(ia); NC_LAUNCH_CHECK();
    indel_impute_group_kernel<<<(unsigned)div_up(n_cols, 64), 64, 0, c->stream>>>(ia); NC_LAUNCH_CHECK();
    return NC_OK;
}

int nc_indel_scan(nc_ctx* c, const NcIndelParams* P, const NcChunk* chunks, int32_t n_chunks, const int32_t* bed, int32_t n_bed,
                  int64_t* n_variants) {
    if (!c || !P || !n_variants || n_chunks < 0 || (n_chunks > 0 && !chunks) || n_bed < 0 || (n_bed > 0 && !bed))
        return fail(c, NC_EINVAL, "nc_indel_scan: bad argument");
    *n_variants = 0;
    if (!c->staged || (!c->tags_staged && !P->haploid)) return fail(c, NC_ESTATE, "nc_indel_scan needs nc_stage_reads and nc_stage_tags");
    if (!c->tags_staged) {        // haploid caller never looks at HP / PS: stage neutral tags
        std::vector<int8_t> z8((size_t)c->n_reads, 0); std::vector<int32_t> z32((size_t)c->n_reads, 0);
        int rc0 = nc_stage_tags(c, z8.data(), z32.data());
        if (rc0) return rc0;
        NC_CUDA(nc_stream_wait(c));
    }
    if (P->win_size < 1 || P->small_win_size < 1 || P->win_size > 190) return fail(c, NC_EINVAL, "win_size must be in 1..190");
    NC_CUDA(cudaSetDevice(c->device));
    int rc = nc_decode_reads(c);
    if (rc) return rc;
    c->indel_scanned = c->indel_built = false; c->n_variants = 0; c->tmi_scan = false;
    for (int i = 0; i + 1 < n_chunks; i++)
        if (chunks[i + 1].start < chunks[i].start) return fail(c, NC_EINVAL, "indel chunks must be sorted by start");
    int64_t lo = INT64_MAX, hi = INT64_MIN;
    for (int i = 0; i < n_chunks; i++) {
        if (chunks[i].end < chunks[i].start) continue;
        lo = std::min<int64_t>(lo, std::max<int64_t>(0, (int64_t)chunks[i].start - 1));
        hi = std::max<int64_t>(hi, (int64_t)chunks[i].end);
    }
    lo = std::max(lo, c->ref_start); hi = std::min(hi, c->ref_start + c->ref_len);
    if (n_chunks == 0 || hi <= lo || c->n_reads == 0) { c->indel_scanned = true; return NC_OK; }

    std::vector<std::pair<int32_t, int32_t>> iv;
    for (int i = 0; i < n_bed; i++) if (bed[2 * i + 1] > bed[2 * i]) iv.emplace_back(bed[2 * i], bed[2 * i + 1]);
    std::sort(iv.begin(), iv.end());
    std::vector<int32_t> merged;
    for (auto& x : iv) {
        if (!merged.empty() && x.first <= merged[merged.size() - 1]) merged[merged.size() - 1] = std::max(merged[merged.size() - 1], x.second);
        else { merged.push_back(x.first); merged.push_back(x.second); }
    }
    const int32_t n_merged = (int32_t)(merged.size() / 2);
    if ((rc = upload(c, c->d_bed, merged.data(), merged.size() * 4))) return rc;
    if ((rc = upload(c, c->d_chunks, chunks, (size_t)n_chunks * sizeof(NcChunk)))) return rc;
    NC_CUDA(nc_stream_wait(c));

    const uint32_t flag_filter = P->supplementary ? 0x704u : 0xF04u;
    const int32_t lo_al = (int32_t)lo & ~7;
    const int64_t n_tiles = div_up(hi - lo_al, kTilePos), n_al = n_tiles * kTilePos;
    c->indel_lo_al = lo_al;
    NC_CUDA(cudaEventRecord(c->evi[0], c->stream));
    NC_CUDA(c->d_idepth.reserve((size_t)3 * n_al * 2));
    NC_CUDA(c->d_em.reserve((size_t)n_al * 4));
    NC_CUDA(c->d_grank.reserve((size_t)(n_al + 1) * 8));
    DepthArgs da = {};
    da.n_reads = c->n_reads; da.pos = c->d_pos.as<int32_t>(); da.end = c->d_end.as<int32_t>(); da.flag = c->d_flag.as<uint16_t>();
    da.hp = c->d_hp.as<int8_t>(); da.pmaxend = c->d_pmaxend.as<int32_t>(); da.rowoff = c->d_rowoff.as<int64_t>();
    da.nwords = c->d_nwords.as<int32_t>(); da.rows = c->d_rows.as<uint32_t>(); da.lo_al = lo_al; da.lo = (int32_t)lo; da.hi = (int32_t)hi;
    da.flag_filter = flag_filter; da.depth = c->d_idepth.as<uint16_t>(); da.n_al = n_al;
    indel_depth_kernel<<<(unsigned)n_tiles, kTileThreads, 0, c->stream>>>(da); NC_LAUNCH_CHECK();
    indel_emitted_kernel<<<(unsigned)div_up(n_al, 256), 256, 0, c->stream>>>(c->d_idepth.as<uint16_t>() + 2 * n_al, n_al, lo_al, c->d_bed.as<int32_t>(),
                                                                           n_merged, c->d_em.as<int32_t>());
    NC_LAUNCH_CHECK();
    if ((rc = device_scan(c, c->d_em.as<int32_t>(), n_al, c->d_grank.as<int64_t>()))) return rc;
    int64_t n_em_total = 0;
    if ((rc = read_i64(c, c->d_grank.as<int64_t>() + n_al, &n_em_total))) return rc;
    NC_CUDA(c->d_empos.reserve((size_t)std::max<int64_t>(n_em_total, 1) * 4));
    indel_empos_kernel<<<(unsigned)div_up(n_al, 256), 256, 0, c->stream>>>(c->d_em.as<int32_t>(), c->d_grank.as<int64_t>(), n_al, lo_al, c->d_empos.as<int32_t>());
    NC_LAUNCH_CHECK();
    NC_CUDA(c->d_ichunks.reserve((size_t)n_chunks * sizeof(IndelChunk)));
    NC_CUDA(c->d_nem1.reserve((size_t)n_chunks * 4));
    NC_CUDA(c->d_rankoff.reserve((size_t)(n_chunks + 1) * 8));
    indel_chunk_kernel<<<(unsigned)div_up(n_chunks, 128), 128, 0, c->stream>>>(c->d_chunks.as<NcChunk>(), n_chunks, lo_al, (int32_t)lo, (int32_t)hi,
                                                                               c->d_grank.as<int64_t>(), c->d_ichunks.as<IndelChunk>(), c->d_nem1.as<int32_t>());
    NC_LAUNCH_CHECK();
    if ((rc = device_scan(c, c->d_nem1.as<int32_t>(), n_chunks, c->d_rankoff.as<int64_t>()))) return rc;
    int64_t R = 0;
    if ((rc = read_i64(c, c->d_rankoff.as<int64_t>() + n_chunks, &R))) return rc;
    c->indel_R = R;
    indel_chunk_off_kernel<<<(unsigned)div_up(n_chunks, 128), 128, 0, c->stream>>>(c->d_ichunks.as<IndelChunk>(), n_chunks, c->d_rankoff.as<int64_t>());
    NC_LAUNCH_CHECK();
    NC_CUDA(c->d_diff.reserve((size_t)8 * R * 4));
    NC_CUDA(cudaMemsetAsync(c->d_diff.p, 0, (size_t)8 * R * 4, c->stream));
    EventArgs ea = {};
    ea.n_reads = c->n_reads; ea.pos = da.pos; ea.end = da.end; ea.flag = da.flag; ea.hp = da.hp; ea.cigar_off = c->d_cigar_off.as<int64_t>();
    ea.cigar = c->d_cigar.as<uint32_t>(); ea.opstart = c->d_opstart.as<int2>(); ea.chunks = c->d_ichunks.as<IndelChunk>(); ea.n_chunks = n_chunks;
    NC_CUDA(c->d_empairs.reserve((size_t)div_up(n_al, 32) * 8));
    indel_empairs_kernel<<<(unsigned)div_up(n_al, 256), 256, 0, c->stream>>>(c->d_em.as<int32_t>(), c->d_grank.as<int64_t>(), n_al, c->d_empairs.as<uint2>());
    NC_LAUNCH_CHECK();
    ea.empairs = c->d_empairs.as<uint2>(); ea.lo_al = lo_al; ea.flag_filter = flag_filter; ea.win = P->win_size; ea.small_win = P->small_win_size; ea.haploid = P->haploid;
    ea.diff = c->d_diff.as<int32_t>(); ea.R = R;
    indel_events_kernel<<<(unsigned)div_up(c->n_reads * 32, 128), 128, 0, c->stream>>>(ea); NC_LAUNCH_CHECK();
    NC_CUDA(c->d_hit.reserve((size_t)R));
    NC_CUDA(c->d_icount.reserve(64));
    NC_CUDA(cudaMemsetAsync(c->d_icount.p, 0, 64, c->stream));
    DecideArgs dd = {};
    dd.chunks = c->d_ichunks.as<IndelChunk>(); dd.n_chunks = n_chunks; dd.R = R; dd.diff = c->d_diff.as<int32_t>(); dd.em_pos = c->d_empos.as<int32_t>();
    dd.depth = c->d_idepth.as<uint16_t>(); dd.n_al = n_al; dd.lo_al = lo_al; dd.mincov = P->mincov; dd.haploid = P->haploid; dd.ins_t = P->ins_t; dd.del_t = P->del_t;
    dd.hit = c->d_hit.as<uint8_t>(); dd.n_hits = c->d_icount.as<unsigned long long>();
    const bool impute = P->impute_indel_phase && !P->haploid;
    if (impute) {
        NC_CUDA(c->d_cdel.reserve((size_t)n_al * 4)); NC_CUDA(c->d_cins.reserve((size_t)n_al * 4));
        NC_CUDA(cudaMemsetAsync(c->d_cdel.p, 0, (size_t)n_al * 4, c->stream));
        NC_CUDA(cudaMemsetAsync(c->d_cins.p, 0, (size_t)n_al * 4, c->stream));
        ImputeCountArgs ic = {};
        ic.n_reads = c->n_reads; ic.pos = da.pos; ic.end = da.end; ic.flag = da.flag; ic.cigar_off = ea.cigar_off; ic.cigar = ea.cigar;
        ic.lo_al = lo_al; ic.lo = (int32_t)lo; ic.hi = (int32_t)hi; ic.flag_filter = flag_filter;
        ic.cdel = c->d_cdel.as<int32_t>(); ic.cins = c->d_cins.as<int32_t>();
        indel_impute_count_kernel<<<(unsigned)div_up(c->n_reads, 128), 128, 0, c->stream>>>(ic); NC_LAUNCH_CHECK();
        dd.impute = 1; dd.cdel = ic.cdel; dd.cins = ic.cins;
    }
    indel_decide_kernel<<<(unsigned)n_chunks, 256, 0, c->stream>>>(dd); NC_LAUNCH_CHECK();      // scans the difference arrays chunk by chunk on the way
    if (impute) {
        int64_t n_pending = 0;
        if ((rc = read_i64(c, c->d_icount.as<int64_t>() + 2, &n_pending))) return rc;
        if (n_pending > 0) {
            NC_CUDA(c->d_imp_g.reserve((size_t)n_pending * 8)); NC_CUDA(c->d_imp_cols.reserve((size_t)n_pending * 4));
            indel_impute_collect_kernel<<<(unsigned)div_up(R, 256), 256, 0, c->stream>>>(c->d_hit.as<uint8_t>(), R, c->d_ichunks.as<IndelChunk>(), n_chunks,
                                                                                       c->d_empos.as<int32_t>(), c->d_imp_g.as<int64_t>(), c->d_imp_cols.as<int32_t>(),
                                                                                       c->d_icount.as<unsigned long long>() + 3);
            NC_LAUNCH_CHECK();
            if ((rc = impute_columns(c, P, n_pending))) return rc;
            indel_impute_apply_kernel<<<(unsigned)div_up(n_pending, 256), 256, 0, c->stream>>>(c->d_imp_g.as<int64_t>(), c->d_imp_ok.as<int32_t>(), n_pending,
                                                                                             c->d_hit.as<uint8_t>(), c->d_icount.as<unsigned long long>());
            NC_LAUNCH_CHECK();
        }
    }
    int64_t n_hits = 0;
    if ((rc = read_i64(c, c->d_icount.as<int64_t>(), &n_hits))) return rc;
    NC_CUDA(c->d_variants.reserve((size_t)std::max<int64_t>(n_hits, 1) * sizeof(NcIndelVariant)));
    indel_greedy_kernel<<<(unsigned)div_up((int64_t)n_chunks * 32, 128), 128, 0, c->stream>>>(c->d_ichunks.as<IndelChunk>(), n_chunks, c->d_hit.as<uint8_t>(),
                                                                                            c->d_empos.as<int32_t>(), P->win_size, c->d_variants.as<NcIndelVariant>(),
                                                                                            c->d_icount.as<unsigned long long>() + 1);
    NC_LAUNCH_CHECK();
    if ((rc = read_i64(c, c->d_icount.as<int64_t>() + 1, &c->n_variants))) return rc;
    NC_CUDA(cudaEventRecord(c->evi[1], c->stream));
    c->tmi_scan = true;
    // aligned rows and CIGAR words once, three u16 depths written and read + the emitted flag per scanned column
    c->tmi.scan_bytes = (uint64_t)c->n_row_words * 4 + (uint64_t)c->n_cigar * 4 + (uint64_t)(hi - lo) * 13;
    c->indel_scanned = true;
    *n_variants = c->n_variants;
    return NC_OK;
}

int nc_indel_fetch_variants(nc_ctx* c, NcIndelVariant* out) {
    if (!c) return NC_EINVAL;
    if (!c->indel_scanned) return fail(c, NC_ESTATE, "nc_indel_fetch_variants before nc_indel_scan");
    NC_CUDA(cudaSetDevice(c->device));
    if (c->n_variants > 0) {
        if (!out) return fail(c, NC_EINVAL, "nc_indel_fetch_variants: null output");
        NC_CUDA(cudaMemcpyAsync(out, c->d_variants.p, (size_t)c->n_variants * sizeof(NcIndelVariant), cudaMemcpyDeviceToHost, c->stream));
    }
    NC_CUDA(nc_stream_wait(c));
    return NC_OK;
}

int nc_indel_build(nc_ctx* c, const NcIndelParams* P, const NcChunk* chunks, int32_t n_chunks, const NcIndelVariant* sites, int64_t n_sites) {
    if (!c || !P || n_sites < 0 || (n_sites > 0 && (!sites || !chunks)) || n_chunks < 0) return fail(c, NC_EINVAL, "nc_indel_build: bad argument");
    if (!c->staged || !c->tags_staged) return fail(c, NC_ESTATE, "nc_indel_build needs nc_stage_reads and nc_stage_tags");
    if (P->window_after < 1 || P->window_after > 260) return fail(c, NC_EINVAL, "window_after must be <= 260");
    NC_CUDA(cudaSetDevice(c->device));
    int rc = nc_decode_reads(c);
    if (rc) return rc;
    c->indel_built = false; c->n_isites = n_sites; c->have_iprobs = false; c->tmi_build = false; c->build_haploid = P->haploid ? 1 : 0;
    c->tmi.n_sites = (uint64_t)n_sites; c->tmi.n_entries = 0; c->tmi.build_bytes = 0;
    if (n_sites == 0) { c->indel_built = true; return NC_OK; }
    NC_CUDA(cudaEventRecord(c->evi[2], c->stream));
    for (int64_t i = 0; i < n_sites; i++)
        if (sites[i].chunk < 0 || sites[i].chunk >= n_chunks) return fail(c, NC_EINVAL, "site %lld names chunk %d of %d", (long long)i, sites[i].chunk, n_chunks);
    // sites found by impute_indel_phase: recompute the two read sets from the source column (generate_indel_pileups.py:309-313)
    std::vector<int32_t> site_imp((size_t)n_sites, -1), imp_cols;
    for (int64_t i = 0; i < n_sites; i++) {
        if (sites[i].src == 0) continue;
        const int64_t p = (int64_t)sites[i].src - 1;
        if (p < c->ref_start || p >= c->ref_start + c->ref_len) return fail(c, NC_EINVAL, "site %lld: source column %d outside the staged reference", (long long)i, sites[i].src);
        site_imp[(size_t)i] = (int32_t)imp_cols.size();
        imp_cols.push_back((int32_t)p);
    }
    const bool any_imp = !imp_cols.empty();
    if (any_imp) {
        if ((rc = upload(c, c->d_imp_cols, imp_cols.data(), imp_cols.size() * 4))) return rc;
        if ((rc = upload(c, c->d_site_imp, site_imp.data(), site_imp.size() * 4))) return rc;
        NC_CUDA(nc_stream_wait(c));                       // the host vectors go out of use before the next blocking call anyway; keep it simple
        if ((rc = impute_columns(c, P, (int64_t)imp_cols.size()))) return rc;
    }
    if ((rc = upload(c, c->d_chunks, chunks, (size_t)n_chunks * sizeof(NcChunk)))) return rc;
    if ((rc = upload(c, c->d_isites, sites, (size_t)n_sites * sizeof(NcIndelVariant)))) return rc;
    NC_CUDA(c->d_site_m.reserve((size_t)n_sites * 4));
    NC_CUDA(c->d_site_cnt.reserve((size_t)n_sites * 4));
    NC_CUDA(c->d_site_off.reserve((size_t)(n_sites + 1) * 8));
    SiteArgs sa = {};
    sa.n_reads = c->n_reads; sa.pos = c->d_pos.as<int32_t>(); sa.end = c->d_end.as<int32_t>(); sa.flag = c->d_flag.as<uint16_t>();
    sa.hp = c->d_hp.as<int8_t>(); sa.ps = c->d_ps.as<int32_t>(); sa.pmaxend = c->d_pmaxend.as<int32_t>();
    sa.cigar_off = c->d_cigar_off.as<int64_t>(); sa.cigar = c->d_cigar.as<uint32_t>(); sa.opstart = c->d_opstart.as<int2>();
    sa.seq_off = c->d_seq_off.as<int64_t>(); sa.l_seq = c->d_lseq.as<int32_t>(); sa.seq4 = c->d_seq4.as<uint8_t>();
    sa.ref = c->d_ref.as<uint8_t>(); sa.ref_start = c->ref_start; sa.ref_len = c->ref_len; sa.contig_len = (int32_t)(c->ref_start + c->ref_len);
    sa.sites = c->d_isites.as<NcIndelVariant>(); sa.n_sites = n_sites; sa.chunks = c->d_chunks.as<NcChunk>();
    sa.flag_filter = P->supplementary ? 0x704u : 0xF04u; sa.wa = P->window_after; sa.win = P->win_size; sa.mincov = P->mincov; sa.maxcov = P->maxcov;
    sa.site_m = c->d_site_m.as<int32_t>(); sa.site_cnt = c->d_site_cnt.as<int32_t>(); sa.site_off = c->d_site_off.as<int64_t>();
    sa.nmax = 264; sa.mmax = 264; sa.cmax = NC_INDEL_CNS_MAX;
    const unsigned sg = (unsigned)div_up(n_sites * 32, 128);
    indel_site_reads_kernel<false><<<sg, 128, 0, c->stream>>>(sa); NC_LAUNCH_CHECK();
    if ((rc = device_scan(c, sa.site_cnt, n_sites, c->d_site_off.as<int64_t>()))) return rc;
    int64_t n_entries = 0;
    if ((rc = read_i64(c, c->d_site_off.as<int64_t>() + n_sites, &n_entries))) return rc;
    const size_t ne = (size_t)std::max<int64_t>(n_entries, 1);
    NC_CUDA(c->d_eread.reserve(ne * 4)); NC_CUDA(c->d_eqpn.reserve(ne * 4)); NC_CUDA(c->d_en.reserve(ne * 4)); NC_CUDA(c->d_egrp.reserve(ne));
    NC_CUDA(c->d_eslice.reserve(ne * sa.nmax)); NC_CUDA(c->d_eacode.reserve(ne * sa.mmax));
    NC_CUDA(c->d_einslen.reserve(ne * (sa.mmax + 1) * 2)); NC_CUDA(c->d_einsfirst.reserve(ne * (sa.mmax + 1) * 2));
    NC_CUDA(c->d_itensors.reserve((size_t)n_sites * 3 * 1280 * 4));
    NC_CUDA(c->d_icns.reserve((size_t)n_sites * 3 * NC_INDEL_CNS_MAX));
    NC_CUDA(c->d_imeta.reserve((size_t)n_sites * sizeof(NcIndelSiteMeta)));
    sa.e_read = c->d_eread.as<int32_t>(); sa.e_qpn = c->d_eqpn.as<int32_t>(); sa.e_n = c->d_en.as<int32_t>(); sa.e_grp = c->d_egrp.as<int8_t>();
    if (any_imp) {
        sa.site_imp = c->d_site_imp.as<int32_t>(); sa.imp_off = c->d_imp_off.as<int64_t>();
        sa.imp_read = c->d_imp_read.as<int32_t>(); sa.imp_label = c->d_imp_label.as<int8_t>();
    }
    sa.e_slice = c->d_eslice.as<uint8_t>(); sa.e_acode = c->d_eacode.as<uint8_t>(); sa.e_inslen = c->d_einslen.as<uint16_t>();
    sa.e_insfirst = c->d_einsfirst.as<uint16_t>(); sa.tensors = c->d_itensors.as<float>(); sa.cns = c->d_icns.as<uint8_t>();
    sa.meta = c->d_imeta.as<NcIndelSiteMeta>();
    indel_site_reads_kernel<true><<<sg, 128, 0, c->stream>>>(sa); NC_LAUNCH_CHECK();
    NC_CUDA(cudaEventRecord(c->evi[3], c->stream));
    if (n_entries > 0) {
        // direction rows 1..n, n <= window_after; strips of 6 columns cover reference windows up to 192 columns (ONT: 161), 9 up to 288
        const int rows = (P->window_after + 2 + 1) & ~1;            // even: keeps the per-warp blocks 16-byte aligned with 2-byte words
        const bool narrow = P->window_after + 1 <= 192;
        // function attributes are per device and one process may open contexts on several: set it on every call (cheap, no sync)
        // NC_INDEL_ALIGN_SCALAR=1: the one-slice-per-warp kernel for every window (validation: tests compare the two kernels at scale)
        const char* scalar_env = getenv("NC_INDEL_ALIGN_SCALAR");
        const bool scalar = scalar_env && scalar_env[0] == '1';
        const int rs = P->window_after / (narrow ? 6 : 9) + 1;        // lane strips in use: reference windows have <= window_after + 1 columns
        const int smem_limit = 227 * 1024 - 1024;
        auto launch2 = [&](auto kernel, int warps, bool wide) -> cudaError_t {
            // two slices of a site per warp step (16-bit SIMD halves), one warp per site; as many warps as the direction words let fit
            const int smem = warps * align2_smem_per_warp(P->window_after, rs, wide);
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return e;
            const unsigned ag = (unsigned)std::min<int64_t>(div_up(n_sites, warps), (int64_t)c->sm_count);
            kernel<<<ag, warps * 32, smem, c->stream>>>(sa, rs);
            return cudaSuccess;
        };
        if (!scalar && narrow && 11 * align2_smem_per_warp(P->window_after, rs, false) <= smem_limit) {
            NC_CUDA(launch2(indel_align2_kernel<6, 11>, 11, false));
        } else if (!scalar && narrow) {
            NC_CUDA(launch2(indel_align2_kernel<6, 6>, 6, false));
        } else if (!scalar && 5 * align2_smem_per_warp(P->window_after, rs, true) <= smem_limit) {
            NC_CUDA(launch2(indel_align2_kernel<9, 5>, 5, true));
        } else if (!scalar && 4 * align2_smem_per_warp(P->window_after, rs, true) <= smem_limit) {
            NC_CUDA(launch2(indel_align2_kernel<9, 4>, 4, true));
        } else if (narrow) {
            const int smem = kAlignWarps * align_smem_per_warp(rows, 2);
            NC_CUDA(cudaFuncSetAttribute(indel_align_kernel<6, uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            const unsigned ag = (unsigned)std::min<int64_t>(div_up(n_entries, kAlignWarps), (int64_t)c->sm_count * 16);
            indel_align_kernel<6, uint16_t><<<ag, kAlignWarps * 32, smem, c->stream>>>(sa, n_entries, rows);
        } else {
            const int smem = kAlignWarps * align_smem_per_warp(rows, 4);
            NC_CUDA(cudaFuncSetAttribute(indel_align_kernel<9, uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            const unsigned ag = (unsigned)std::min<int64_t>(div_up(n_entries, kAlignWarps), (int64_t)c->sm_count * 16);
            indel_align_kernel<9, uint32_t><<<ag, kAlignWarps * 32, smem, c->stream>>>(sa, n_entries, rows);
        }
        NC_LAUNCH_CHECK();
    }
    NC_CUDA(cudaEventRecord(c->evi[4], c->stream));
    indel_msa_kernel<<<(unsigned)n_sites, 96, 0, c->stream>>>(sa); NC_LAUNCH_CHECK();
    NC_CUDA(cudaEventRecord(c->evi[5], c->stream));
    {
        // I4: consensus x reference window -> allele lengths (generate_indel_pileups.py:77-127), one warp per (site, group)
        const bool narrow = P->window_after + 1 <= 192;
        // 32 resident warps per SM (64 registers): an alignment is ~230 k cycles of mostly latency (wavefront steps, then the traceback and
        // the walk on one lane through the L2-resident scratch), so the kernel's rate is warps / latency
        const int64_t n_warps = std::min<int64_t>(n_sites * 3, (int64_t)c->sm_count * 32);
        const int rows_cap = NC_INDEL_CNS_MAX + 1;
        NC_CUDA(c->d_ialleles.reserve((size_t)n_sites * 3 * 2 * sizeof(int32_t)));
        NC_CUDA(c->d_allele_dirs.reserve((size_t)n_warps * rows_cap * 32 * (narrow ? 4 : 8)));
        NC_CUDA(c->d_allele_ops.reserve((size_t)n_warps * (NC_INDEL_CNS_MAX + 272)));
        AlleleArgs aa = {};
        aa.n_sites = n_sites; aa.sites = sa.sites; aa.meta = sa.meta; aa.cns = sa.cns; aa.cmax = sa.cmax; aa.ref = sa.ref; aa.ref_start = c->ref_start;
        aa.win = P->win_size; aa.haploid = P->haploid; aa.go = 9; aa.ge = 1; aa.match = 20; aa.mismatch = -10;     // parasail call of :79-80
        aa.scratch = c->d_allele_dirs.p; aa.rows_cap = rows_cap; aa.ops_scratch = c->d_allele_ops.as<uint8_t>(); aa.out = c->d_ialleles.as<int32_t>();
        const unsigned gb = (unsigned)div_up(n_warps * 32, 128);
        if (narrow) indel_allele_kernel<6, uint32_t><<<gb, 128, 0, c->stream>>>(aa);
        else indel_allele_kernel<9, uint64_t><<<gb, 128, 0, c->stream>>>(aa);
        NC_LAUNCH_CHECK();
    }
    NC_CUDA(cudaEventRecord(c->evi[8], c->stream));
    c->tmi_build = true; c->tmi.n_entries = (uint64_t)n_entries;
    // per aligned slice: the query slice and the reference window (one byte per base); per site: 3 x [5][128][2] fp32 tensors,
    // the consensus strings and the site record
    c->tmi.build_bytes = (uint64_t)n_entries * (uint64_t)(2 * (P->window_after + 1)) +
                         (uint64_t)n_sites * (3 * 1280 * 4 + 3 * NC_INDEL_CNS_MAX + sizeof(NcIndelSiteMeta));
    c->indel_built = true;
    return NC_OK;
}

int nc_indel_fetch(nc_ctx* c, NcIndelSiteMeta* meta, float* tensors, uint8_t* cns) {
    if (!c) return NC_EINVAL;
    if (!c->indel_built) return fail(c, NC_ESTATE, "nc_indel_fetch before nc_indel_build");
    NC_CUDA(cudaSetDevice(c->device));
    const size_t n = (size_t)c->n_isites;
    if (n) {
        if (meta) NC_CUDA(cudaMemcpyAsync(meta, c->d_imeta.p, n * sizeof(NcIndelSiteMeta), cudaMemcpyDeviceToHost, c->stream));
        if (tensors) NC_CUDA(cudaMemcpyAsync(tensors, c->d_itensors.p, n * 3 * 1280 * 4, cudaMemcpyDeviceToHost, c->stream));
        if (cns) NC_CUDA(cudaMemcpyAsync(cns, c->d_icns.p, n * 3 * NC_INDEL_CNS_MAX, cudaMemcpyDeviceToHost, c->stream));
    }
    NC_CUDA(nc_stream_wait(c));
    return NC_OK;
}

int nc_indel_fetch_alleles(nc_ctx* c, int32_t* out) {
    if (!c) return NC_EINVAL;
    if (!c->indel_built) return fail(c, NC_ESTATE, "nc_indel_fetch_alleles before nc_indel_build");
    NC_CUDA(cudaSetDevice(c->device));
    if (c->n_isites) {
        if (!out) return fail(c, NC_EINVAL, "nc_indel_fetch_alleles: null output");
        NC_CUDA(cudaMemcpyAsync(out, c->d_ialleles.p, (size_t)c->n_isites * 3 * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    }
    NC_CUDA(nc_stream_wait(c));
    return NC_OK;
}

int nc_indel_fetch_range(nc_ctx* c, int64_t first, int64_t count, float* tensors) {
    if (!c) return NC_EINVAL;
    if (!c->indel_built) return fail(c, NC_ESTATE, "nc_indel_fetch_range before nc_indel_build");
    if (first < 0 || count < 0 || first + count > c->n_isites || (count > 0 && !tensors)) return fail(c, NC_EINVAL, "nc_indel_fetch_range: bad range");
    NC_CUDA(cudaSetDevice(c->device));
    if (count) NC_CUDA(cudaMemcpyAsync(tensors, c->d_itensors.as<float>() + first * 3 * 1280, (size_t)count * 3 * 1280 * 4, cudaMemcpyDeviceToHost, c->stream));
    NC_CUDA(nc_stream_wait(c));
    return NC_OK;
}

int nc_indel_forward(nc_ctx* c, int impl, float* probs) {
    if (!c) return NC_EINVAL;
    if (!c->indel_built) return fail(c, NC_ESTATE, "nc_indel_forward before nc_indel_build");
    NC_CUDA(cudaSetDevice(c->device));
    const int hap = c->build_haploid;
    Model& M = c->indel[hap];
    if (!M.loaded) return fail(c, NC_ESTATE, "nc_indel_forward: no %s indel weights loaded", hap ? "haploid" : "diploid");
    const int64_t n = c->n_isites;
    const int nout = hap ? 1 : 4;
    NC_CUDA(cudaEventRecord(c->evi[6], c->stream));
    if (n > 0) {
        NC_CUDA(c->d_iprobs.reserve((size_t)n * nout * sizeof(float)));
        // [n][3][5][128][2] is the hstack of the three groups (indelCaller.py:83); the haploid model reads group 2 of every site
        const float* x = c->d_itensors.as<float>() + (hap ? 2 * 1280 : 0);
        int rc = cnn_forward(c, M, impl, 0, x, 3 * 1280, n, nullptr, nullptr, nullptr, nullptr, c->d_iprobs.as<float>(), nullptr);
        if (rc) return rc;
    }
    NC_CUDA(cudaEventRecord(c->evi[7], c->stream));
    c->tmi_cnn = true; c->have_iprobs = true;
    if (probs && n > 0) {
        NC_CUDA(cudaMemcpyAsync(probs, c->d_iprobs.p, (size_t)n * nout * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        NC_CUDA(nc_stream_wait(c));
        if (impl == 0) return tc_check(c, M);
    }
    return NC_OK;
}

int nc_indel_fetch_probs(nc_ctx* c, float* probs) {
    if (!c || !probs) return fail(c, NC_EINVAL, "nc_indel_fetch_probs: null argument");
    if (!c->indel_built || !c->have_iprobs) return fail(c, NC_ESTATE, "nc_indel_fetch_probs before nc_indel_forward");
    NC_CUDA(cudaSetDevice(c->device));
    if (c->n_isites) NC_CUDA(cudaMemcpyAsync(probs, c->d_iprobs.p, (size_t)c->n_isites * (c->build_haploid ? 1 : 4) * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    NC_CUDA(nc_stream_wait(c));
    return tc_check(c, c->indel[c->build_haploid]);
}

int nc_get_indel_timings(nc_ctx* c, NcIndelTimings* out) {
    if (!c || !out) return NC_EINVAL;
    NC_CUDA(cudaSetDevice(c->device));
    NC_CUDA(nc_stream_wait(c));
    if (c->tmi_scan) { NC_CUDA(cudaEventElapsedTime(&c->tmi.scan_ms, c->evi[0], c->evi[1])); }
    if (c->tmi_build) {
        NC_CUDA(cudaEventElapsedTime(&c->tmi.reads_ms, c->evi[2], c->evi[3]));
        NC_CUDA(cudaEventElapsedTime(&c->tmi.align_ms, c->evi[3], c->evi[4]));
        NC_CUDA(cudaEventElapsedTime(&c->tmi.msa_ms, c->evi[4], c->evi[5]));
        NC_CUDA(cudaEventElapsedTime(&c->tmi.allele_ms, c->evi[5], c->evi[8]));
    }
    if (c->tmi_cnn) { NC_CUDA(cudaEventElapsedTime(&c->tmi.cnn_ms, c->evi[6], c->evi[7])); }
    *out = c->tmi;
    return NC_OK;
}

// Global affine-gap alignment with traceback on the host (the reference calls parasail, a CPU library, here too).
// Same recurrences and tie rules as oracle/star_msa.nw_trace: H ties DIAG > F ('I') > E ('D'); E/F ties prefer extension.
int nc_nw_trace(const uint8_t* q, int32_t n, const uint8_t* r, int32_t m, int32_t go, int32_t ge, int32_t match, int32_t mismatch,
                uint32_t* out, int32_t cap) {
    if (n < 0 || m < 0 || (n > 0 && !q) || (m > 0 && !r) || !out || cap < 1) return NC_EINVAL;
    const int64_t NEG = -(1ll << 40);
    const size_t W = (size_t)m + 1;
    std::vector<int64_t> H((size_t)(n + 1) * W, NEG), E((size_t)(n + 1) * W, NEG), F((size_t)(n + 1) * W, NEG);
    H[0] = 0;
    for (int j = 1; j <= m; j++) { E[j] = -go - (int64_t)ge * (j - 1); H[j] = E[j]; }
    for (int i = 1; i <= n; i++) { F[i * W] = -go - (int64_t)ge * (i - 1); H[i * W] = F[i * W]; }
    for (int i = 1; i <= n; i++)
        for (int j = 1; j <= m; j++) {
            const int64_t s = q[i - 1] == r[j - 1] ? match : mismatch;
            const int64_t f = std::max(F[(i - 1) * W + j] - ge, H[(i - 1) * W + j] - go);
            const int64_t e = std::max(E[i * W + j - 1] - ge, H[i * W + j - 1] - go);
            F[i * W + j] = f; E[i * W + j] = e;
            H[i * W + j] = std::max(H[(i - 1) * W + j - 1] + s, std::max(f, e));
        }
    std::vector<uint8_t> ops;
    int i = n, j = m, state = 0;     // 0 H, 1 F, 2 E
    while (i > 0 || j > 0) {
        if (state == 0) {
            if (i > 0 && j > 0 && H[i * W + j] == H[(i - 1) * W + j - 1] + (q[i - 1] == r[j - 1] ? match : mismatch)) {
                ops.push_back(q[i - 1] == r[j - 1] ? 7 : 8); i--; j--;
            } else if (i > 0 && H[i * W + j] == F[i * W + j]) state = 1;
            else state = 2;
        } else if (state == 1) {
            ops.push_back(1);
            if (i > 1 && F[i * W + j] == F[(i - 1) * W + j] - ge) i--; else { i--; state = 0; }
        } else {
            ops.push_back(2);
            if (j > 1 && E[i * W + j] == E[i * W + j - 1] - ge) j--; else { j--; state = 0; }
        }
    }
    int32_t nw = 0;
    for (size_t k = ops.size(); k-- > 0;) {
        if (nw > 0 && (out[nw - 1] & 15u) == ops[k]) out[nw - 1] += 16u;
        else { if (nw >= cap) return NC_EOVERFLOW; out[nw++] = (1u << 4) | ops[k]; }
    }
    return nw;
}

// allele_prediction (generate_indel_pileups.py:77-127) for a batch of (consensus, reference window) pairs on host threads:
// alignment by nc_nw_trace, then the reference's walk over the CIGAR, control flow kept line for line (including its
// `sum(ref_cnt) - cnt` after a trailing insertion).  Outputs the prefix lengths of the reference / alternative allele
// strings, or -1 / -1 where the reference returns (None, None).
static void allele_predict_one(const uint8_t* alt, int32_t n, const uint8_t* ref, int32_t m, int32_t max_range, int32_t go, int32_t ge,
                               int32_t match, int32_t mismatch, std::vector<uint32_t>& cig, int32_t* ref_out, int32_t* alt_out) {
    *ref_out = -1; *alt_out = -1;
    cig.resize((size_t)n + m + 2);
    const int nw = nc_nw_trace(alt, n, ref, m, go, ge, match, mismatch, cig.data(), (int32_t)cig.size());
    if (nw <= 0) return;
    bool indel = false, mis_before = false;
    int64_t ref_cnt[10] = {0}, alt_cnt[10] = {0}, mis_after[2] = {0, 0};
    auto sum10 = [](const int64_t* a) { int64_t t = 0; for (int i = 0; i < 10; i++) t += a[i]; return t; };
    int op = 0; int64_t cnt = 0;
    for (int k = 0; k < nw; k++) {
        op = (int)(cig[k] & 15u); cnt = (int64_t)(cig[k] >> 4);
        if (op == 8 || op == 7) {
            ref_cnt[op] += cnt; alt_cnt[op] += cnt;
            if (indel) mis_after[op - 7] += cnt; else mis_before = true;
        }
        if (op == 1) { alt_cnt[op] += cnt; mis_after[0] = mis_after[1] = 0; indel = true; }
        if (op == 2) { ref_cnt[op] += cnt; mis_after[0] = mis_after[1] = 0; indel = true; }
        if (!indel && sum10(ref_cnt) >= (int64_t)max_range + 10) {
            if (ref_cnt[8]) {
                const int64_t out_len = op == 8 ? sum10(ref_cnt) : sum10(ref_cnt) - cnt;
                *ref_out = (int32_t)out_len; *alt_out = (int32_t)out_len;
            }
            return;
        }
        if (indel && mis_after[0] + mis_after[1] > 20) break;
    }
    int64_t r = op == 8 ? sum10(ref_cnt) : sum10(ref_cnt) - cnt;
    int64_t a = op == 8 ? sum10(alt_cnt) : sum10(alt_cnt) - cnt;
    if (!mis_before) { r += 1; a += 1; }
    *ref_out = (int32_t)r; *alt_out = (int32_t)a;
}

int nc_allele_predict_batch(int64_t n_items, const uint8_t* alt_codes, const int64_t* alt_off, const int32_t* alt_len, const uint8_t* ref_codes,
                            const int64_t* ref_off, const int32_t* ref_len, const int32_t* max_range, int32_t go, int32_t ge, int32_t match,
                            int32_t mismatch, int32_t threads, int32_t* ref_out_len, int32_t* alt_out_len) {
    if (n_items < 0 || (n_items > 0 && (!alt_codes || !alt_off || !alt_len || !ref_codes || !ref_off || !ref_len || !max_range || !ref_out_len || !alt_out_len)))
        return NC_EINVAL;
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    threads = (int)std::max<int64_t>(1, std::min<int64_t>(threads, n_items));
    std::atomic<int64_t> next{0};
    auto work = [&]() {
        std::vector<uint32_t> cig;
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= n_items) return;
            allele_predict_one(alt_codes + alt_off[i], alt_len[i], ref_codes + ref_off[i], ref_len[i], max_range[i], go, ge, match, mismatch, cig,
                               ref_out_len + i, alt_out_len + i);
        }
    };
    if (threads == 1) { work(); return NC_OK; }
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) pool.emplace_back(work);
    for (auto& th : pool) th.join();
    return NC_OK;
}

// SNP genotype decision + VCF record text (snpCaller.py:113-163 diploid, :183-198 haploid) for n call records on host threads.
// Same decisions and the same printf conversions as the reference's Python (% formatting of float64 values): QUAL from the
// float32 probability widened to float64, PR= in A,C,G,T order, depths printed as integers.
static int format_one_snp(char* o, const char* chrom, int32_t pos, int r, const float* p, int32_t dp, int32_t alt_cnt, const uint16_t* fw,
                          const uint16_t* rv, int haploid, uint8_t* is_pass) {
    static const char B[4] = {'A', 'G', 'T', 'C'};
    const double freq = (double)alt_cnt / (double)dp;
    char info[96];
    snprintf(info, sizeof(info), "PR=%.4f,%.4f,%.4f,%.4f;FQ=%.4f", (double)p[0], (double)p[3], (double)p[1], (double)p[2], freq);
    auto qual = [](float x, double cap, double mult) { return std::min(cap, -mult * log10(1e-10 + 1.0 - (double)x)); };
    *is_pass = 0;
    if (haploid) {
        int pred = 0;
        for (int i = 1; i < 4; i++) if (p[i] > p[pred]) pred = i;                      // np.argmax: first maximum
        const bool pass = pred != r;
        *is_pass = pass;
        return sprintf(o, "%s\t%d\t.\t%c\t%c\t%.3f\t%s\t%s\tGT:DP:VF:AD:ADF:ADR\t1/1:%d:%.4f:.:.:.\n", chrom, pos, B[r], B[pred],
                       qual(p[pred], 999.0, 100.0), pass ? "PASS" : "REF", info, dp, freq);
    }
    int order[4] = {0, 1, 2, 3};                                                       // stable ascending argsort
    for (int i = 1; i < 4; i++) { const int v = order[i]; int j = i; while (j > 0 && p[order[j - 1]] > p[v]) { order[j] = order[j - 1]; j--; } order[j] = v; }
    const int p1 = order[3], p2 = order[2];
    int k = 0;
    for (int i = 0; i < 4; i++) k += p[i] >= 0.5f;
    const double q1 = qual(p[p1], 99.0, 10.0), q2 = qual(p[p2], 99.0, 10.0);
    const int rf = fw[r], rr = rv[r];
    if (k >= 2) {
        int alt;
        if (p1 == r) alt = p2;
        else if (p2 == r && p[p2] >= 0.5f) alt = p1;
        else if (p2 != r && p1 != r && p[p2] >= 0.5f) {
            const int f1 = fw[p1], r1 = rv[p1], f2 = fw[p2], r2 = rv[p2];
            *is_pass = 1;
            return sprintf(o, "%s\t%d\t.\t%c\t%c,%c\t%.3f\tPASS\t%s\tGT:DP:VF:AD:ADF:ADR\t1/2:%d:%.4f,%.4f:%d,%d,%d:%d,%d,%d:%d,%d,%d\n", chrom, pos, B[r],
                           B[p1], B[p2], q2, info, dp, (double)(f1 + r1) / dp, (double)(f2 + r2) / dp, rf + rr, f1 + r1, f2 + r2, rf, f1, f2, rr, r1, r2);
        } else return 0;
        const int af = fw[alt], ar = rv[alt];
        *is_pass = 1;
        return sprintf(o, "%s\t%d\t.\t%c\t%c\t%.3f\tPASS\t%s\tGT:DP:VF:AD:ADF:ADR\t0/1:%d:%.4f:%d,%d:%d,%d:%d,%d\n", chrom, pos, B[r], B[alt], q2, info, dp,
                       (double)(af + ar) / dp, rf + rr, af + ar, rf, af, rr, ar);
    }
    if (k == 1 && r != p1) {
        const int af = fw[p1], ar = rv[p1];
        *is_pass = 1;
        return sprintf(o, "%s\t%d\t.\t%c\t%c\t%.3f\tPASS\t%s\tGT:DP:VF:AD:ADF:ADR\t1/1:%d:%.4f:%d,%d:%d,%d:%d,%d\n", chrom, pos, B[r], B[p1], q1, info, dp,
                       (double)(af + ar) / dp, rf + rr, af + ar, rf, af, rr, ar);
    }
    if (k == 1) return sprintf(o, "%s\t%d\t.\t%c\t.\t%.3f\tREF\t%s\tGT:DP:VF:AD:ADF:ADR\t./.:%d:.:.:.:.\n", chrom, pos, B[r], q1, info, dp);
    return sprintf(o, "%s\t%d\t.\t%c\t.\t%.3f\tLOW\t%s\tGT:DP:VF:AD:ADF:ADR\t./.:%d:.:.:.:.\n", chrom, pos, B[r], 0.0, info, dp);
}

int64_t nc_format_snp_records(const char* chrom, int64_t n, const int32_t* pos, const uint8_t* ref_code, const float* probs4, const int32_t* dp,
                              const int32_t* alt_cnt, const uint16_t* fwd4, const uint16_t* rev4, int32_t haploid, int32_t threads, char* out,
                              int64_t cap, int64_t* line_off, uint8_t* is_pass) {
    if (!chrom || n < 0 || strlen(chrom) > 200 || (n > 0 && (!pos || !ref_code || !probs4 || !dp || !alt_cnt || !fwd4 || !rev4 || !out || !line_off || !is_pass)))
        return NC_EINVAL;
    constexpr int SLOT = 512;                                                         // a record is < 200 characters + the contig name
    std::vector<char> tmp((size_t)std::max<int64_t>(n, 1) * SLOT);
    std::vector<int32_t> len((size_t)n);
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    threads = (int)std::max<int64_t>(1, std::min<int64_t>(threads, (n + 4095) / 4096));
    std::atomic<int64_t> next{0};
    std::atomic<int> bad{0};
    auto work = [&]() {
        for (;;) {
            const int64_t a = next.fetch_add(4096);
            if (a >= n) return;
            const int64_t e = std::min(n, a + 4096);
            for (int64_t i = a; i < e; i++) {
                if (ref_code[i] > 3 || dp[i] <= 0) { bad = 1; len[(size_t)i] = 0; is_pass[i] = 0; continue; }
                len[(size_t)i] = format_one_snp(tmp.data() + (size_t)i * SLOT, chrom, pos[i], ref_code[i], probs4 + 4 * i, dp[i], alt_cnt[i], fwd4 + 4 * i,
                                                rev4 + 4 * i, haploid, is_pass + i);
            }
        }
    };
    if (threads == 1) work();
    else { std::vector<std::thread> pool; for (int t = 0; t < threads; t++) pool.emplace_back(work); for (auto& th : pool) th.join(); }
    if (bad) return NC_EINVAL;
    int64_t total = 0;
    for (int64_t i = 0; i < n; i++) { line_off[i] = total; total += len[(size_t)i]; }
    line_off[n] = total;
    if (total > cap) return NC_EOVERFLOW;
    for (int64_t i = 0; i < n; i++) memcpy(out + line_off[i], tmp.data() + (size_t)i * SLOT, (size_t)len[(size_t)i]);
    return total;
}


// ---- device-side BAM input -----------------------------------------------------------------------------------------
int nc_bam_device_close(nc_ctx* c) {
    if (!c) return NC_EINVAL;
    NC_CUDA(cudaSetDevice(c->device));
    NC_CUDA(nc_stream_wait(c));
    for (DevBuf* b : {&c->d_bam_comp, &c->d_bam_blocks, &c->d_bam, &c->d_bam_recoff, &c->d_bam_rid, &c->d_bam_pos, &c->d_bam_flag, &c->d_bam_lseq, &c->d_bam_ncig,
                      &c->d_bam_nseq, &c->d_bam_cigsrc, &c->d_bam_seqsrc, &c->d_bam_hp, &c->d_bam_ps, &c->d_bam_wst, &c->d_bam_wcnt, &c->d_bam_woff}) b->release();
    c->bam_open = false; c->bam_contigs.clear(); c->bam_records = 0;
    return NC_OK;
}

int nc_bam_device_open(nc_ctx* c, const char* path, int32_t* n_contigs) {
    if (!c || !path || !n_contigs) return fail(c, NC_EINVAL, "nc_bam_device_open: null argument");
    *n_contigs = 0;
    NC_CUDA(cudaSetDevice(c->device));
    c->bam_open = false; c->bam_contigs.clear();
    const double t0 = now_ms();
    MapRO f(path);
    if (!f.ok) return fail(c, NC_EINVAL, "cannot open %s", path);
    // ---- BGZF block table from the block headers (BSIZE) and trailers (ISIZE)
    std::vector<BgzfBlock> blocks;
    std::vector<int64_t> block_start;                              // file offset of every listed block (virtual offsets name blocks by it)
    int64_t total = 0;
    for (size_t off = 0; off < f.n;) {
        if (off + 18 > f.n) return fail(c, NC_EINVAL, "truncated or malformed BGZF block");
        const uint8_t* p = f.p + off;
        if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return fail(c, NC_EINVAL, "not a BGZF stream");
        const uint16_t xlen = rd_le<uint16_t>(p + 10);
        size_t x = 12; const size_t xend = 12 + (size_t)xlen;
        int bsize = -1;
        while (x + 4 <= xend && off + x + 4 <= f.n) {
            const uint16_t slen = rd_le<uint16_t>(p + x + 2);
            if (p[x] == 'B' && p[x + 1] == 'C' && slen == 2) bsize = rd_le<uint16_t>(p + x + 4);
            x += 4 + slen;
        }
        const size_t blen = (size_t)bsize + 1;
        if (bsize < 0 || off + blen > f.n || blen < xend + 8) return fail(c, NC_EINVAL, "truncated or malformed BGZF block");
        const uint32_t isize = rd_le<uint32_t>(p + blen - 4);
        if (isize) { blocks.push_back({(int64_t)(off + xend), (int32_t)(blen - xend - 8), (int32_t)isize, total}); block_start.push_back((int64_t)off); }
        total += isize;
        off += blen;
    }
    if (blocks.empty() || total < 12) return fail(c, NC_EINVAL, "empty BAM file");
    size_t free_b = 0, total_b = 0;
    NC_CUDA(cudaMemGetInfo(&free_b, &total_b));
    if ((size_t)total + f.n + (size_t)total / 8 > free_b / 2) return fail(c, NC_ENOMEM, "inflated BAM (%lld bytes) does not fit the device budget", (long long)total);
    // ---- compressed bytes -> pinned -> device (threads: page-cache reads + first touch of the pinned pages)
    NC_CUDA(c->pin_bam.reserve(f.n + blocks.size() * sizeof(BgzfBlock) + 64));
    uint8_t* pin = c->pin_bam.as<uint8_t>();
    {
        const int nt = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        std::vector<std::thread> pool;
        const size_t per = (f.n + nt - 1) / nt;
        for (int t = 0; t < nt; t++) pool.emplace_back([&, t]() { const size_t a = t * per, e = std::min(f.n, a + per); if (a < e) memcpy(pin + a, f.p + a, e - a); });
        for (auto& th : pool) th.join();
    }
    const size_t blk_at = (f.n + 63) & ~(size_t)63;
    memcpy(pin + blk_at, blocks.data(), blocks.size() * sizeof(BgzfBlock));
    const double t1 = now_ms();
    NC_CUDA(c->d_bam_comp.reserve(blk_at + blocks.size() * sizeof(BgzfBlock)));
    NC_CUDA(c->d_bam.reserve((size_t)total + 64));
    NC_CUDA(c->d_bam_err.reserve(64));
    NC_CUDA(c->d_bam_out.reserve(64));
    NC_CUDA(cudaMemsetAsync(c->d_bam_err.p, 0, 64, c->stream));
    cudaEvent_t e0 = c->evi[0], e1 = c->evi[1], e2 = c->evi[2], e3 = c->evi[3];
    NC_CUDA(cudaEventRecord(e0, c->stream));
    NC_CUDA(cudaMemcpyAsync(c->d_bam_comp.p, pin, blk_at + blocks.size() * sizeof(BgzfBlock), cudaMemcpyHostToDevice, c->stream));
    NC_CUDA(cudaEventRecord(e1, c->stream));
    const int64_t nb = (int64_t)blocks.size();
    bgzf_inflate_kernel<<<(unsigned)div_up(nb, kInflWarps), kInflWarps * 32, 0, c->stream>>>(
        c->d_bam_comp.as<uint8_t>(), reinterpret_cast<const BgzfBlock*>(c->d_bam_comp.as<uint8_t>() + blk_at), nb, c->d_bam.as<uint8_t>(), c->d_bam_err.as<int>());
    NC_LAUNCH_CHECK();
    NC_CUDA(cudaEventRecord(e2, c->stream));
    // ---- header: magic, text, references (D2H of the head of the stream, more if the reference list is long)
    std::vector<uint8_t> head;
    size_t want = std::min<size_t>((size_t)total, 1 << 20);
    int64_t first = 0;
    std::vector<nc_ctx::BamContig> contigs;
    for (;;) {
        head.resize(want);
        NC_CUDA(cudaMemcpyAsync(head.data(), c->d_bam.p, want, cudaMemcpyDeviceToHost, c->stream));
        NC_CUDA(nc_stream_wait(c));
        int herr[4] = {0, 0, 0, 0};
        NC_CUDA(cudaMemcpy(herr, c->d_bam_err.p, sizeof(herr), cudaMemcpyDeviceToHost));
        if (herr[0]) return fail(c, NC_EINVAL, "BGZF inflate failed for %d block(s) (corrupt or unsupported DEFLATE stream)", herr[0]);
        const uint8_t* d = head.data();
        if (memcmp(d, "BAM\1", 4) != 0) return fail(c, NC_EINVAL, "not a BAM file (bad magic)");
        const int32_t l_text = rd_le<int32_t>(d + 4);
        bool more = false;
        size_t off = 8 + (size_t)std::max(l_text, 0);
        contigs.clear();
        if (l_text < 0) return fail(c, NC_EINVAL, "truncated BAM header");
        if (off + 4 > want) more = true;
        else {
            const int32_t n_ref = rd_le<int32_t>(d + off);
            off += 4;
            if (n_ref < 0) return fail(c, NC_EINVAL, "negative reference count");
            for (int32_t i = 0; i < n_ref && !more; i++) {
                if (off + 4 > want) { more = true; break; }
                const int32_t l_name = rd_le<int32_t>(d + off);
                if (l_name <= 0) return fail(c, NC_EINVAL, "truncated reference list");
                if (off + 8 + (size_t)l_name > want) { more = true; break; }
                nc_ctx::BamContig bc;
                bc.name.assign((const char*)d + off + 4, (size_t)l_name - 1);
                bc.length = rd_le<int32_t>(d + off + 4 + l_name);
                contigs.push_back(bc);
                off += 8 + (size_t)l_name;
            }
        }
        if (!more) { first = (int64_t)off; break; }
        if (want >= (size_t)total) return fail(c, NC_EINVAL, "truncated BAM header");
        want = std::min<size_t>((size_t)total, want * 8);
    }
    // ---- record chain, then fields (parallel).  With a BAI index next to the file its virtual offsets give record starts all over
    //      the stream, and the chain is followed from all of them at once; without one (or if a segment does not close) one thread walks it.
    const int64_t cap = total / 96 + 4096;
    NC_CUDA(c->d_bam_recoff.reserve((size_t)cap * 8));
    int64_t wo[2] = {0, 0};
    bool walked = false;
    {
        std::vector<int64_t> starts;
        std::string bai = std::string(path) + ".bai";
        MapRO x(bai.c_str());
        if (!x.ok) { bai = path; if (bai.size() > 4 && bai.compare(bai.size() - 4, 4, ".bam") == 0) { bai = bai.substr(0, bai.size() - 4) + ".bai"; } MapRO y(bai.c_str()); if (y.ok) { std::swap(x.p, y.p); std::swap(x.n, y.n); x.ok = true; y.ok = false; } }
        if (x.ok && x.n >= 8 && memcmp(x.p, "BAI\1", 4) == 0) {
            auto add = [&](uint64_t v) {
                if (!v) return;
                const int64_t co = (int64_t)(v >> 16), uo = (int64_t)(v & 0xffff);
                const auto it = std::lower_bound(block_start.begin(), block_start.end(), co);
                if (it == block_start.end() || *it != co) return;
                const BgzfBlock& bk = blocks[(size_t)(it - block_start.begin())];
                if (uo < bk.out_len && bk.out_off + uo >= first) starts.push_back(bk.out_off + uo);
            };
            size_t o = 8; bool okb = true;
            const int32_t n_ref = rd_le<int32_t>(x.p + 4);
            for (int32_t r = 0; r < n_ref && okb; r++) {
                if (o + 4 > x.n) { okb = false; break; }
                const int32_t n_bin = rd_le<int32_t>(x.p + o); o += 4;
                for (int32_t b = 0; b < n_bin && okb; b++) {
                    if (o + 8 > x.n) { okb = false; break; }
                    const uint32_t bin = rd_le<uint32_t>(x.p + o); const int32_t n_chunk = rd_le<int32_t>(x.p + o + 4); o += 8;
                    if (n_chunk < 0 || o + 16ull * (size_t)n_chunk > x.n) { okb = false; break; }
                    if (bin != 37450) for (int32_t k = 0; k < n_chunk; k++) add(rd_le<uint64_t>(x.p + o + 16ull * k));      // 37450: the metadata pseudo-bin
                    o += 16ull * (size_t)n_chunk;
                }
                if (!okb || o + 4 > x.n) { okb = false; break; }
                const int32_t n_intv = rd_le<int32_t>(x.p + o); o += 4;
                if (n_intv < 0 || o + 8ull * (size_t)n_intv > x.n) { okb = false; break; }
                for (int32_t k = 0; k < n_intv; k++) add(rd_le<uint64_t>(x.p + o + 8ull * k));
                o += 8ull * (size_t)n_intv;
            }
            if (okb) {
                starts.push_back(first);
                std::sort(starts.begin(), starts.end());
                starts.erase(std::unique(starts.begin(), starts.end()), starts.end());
            } else starts.clear();
        }
        if (starts.size() >= 64) {
            const int64_t ns = (int64_t)starts.size();
            // (context members: a cudaMalloc / cudaFree pair per open costs anything from microseconds to hundreds of milliseconds)
            DevBuf& d_st = c->d_bam_wst; DevBuf& d_cnt = c->d_bam_wcnt; DevBuf& d_off = c->d_bam_woff;
            cudaError_t e_ = d_st.reserve((size_t)ns * 8);
            if (e_ == cudaSuccess) e_ = d_cnt.reserve((size_t)ns * 4);
            if (e_ == cudaSuccess) e_ = d_off.reserve((size_t)(ns + 1) * 8);
            if (e_ == cudaSuccess) e_ = cudaMemcpyAsync(d_st.p, starts.data(), (size_t)ns * 8, cudaMemcpyHostToDevice, c->stream);
            int rcw = NC_OK;
            if (e_ == cudaSuccess) {
                bam_walk_seg_kernel<false><<<(unsigned)div_up(ns, 128), 128, 0, c->stream>>>(c->d_bam.as<uint8_t>(), d_st.as<int64_t>(), ns, total, d_cnt.as<int32_t>(), nullptr,
                                                                                           nullptr, c->d_bam_err.as<int>());
                c->launches++;
                rcw = device_scan(c, d_cnt.as<int32_t>(), ns, d_off.as<int64_t>());
                int64_t nrec_seg = 0;
                if (!rcw) rcw = read_i64(c, d_off.as<int64_t>() + ns, &nrec_seg);
                int herr4[4] = {0, 0, 0, 0};
                if (!rcw && cudaMemcpy(herr4, c->d_bam_err.p, sizeof(herr4), cudaMemcpyDeviceToHost) == cudaSuccess && herr4[3] == 0 && nrec_seg <= cap) {
                    bam_walk_seg_kernel<true><<<(unsigned)div_up(ns, 128), 128, 0, c->stream>>>(c->d_bam.as<uint8_t>(), d_st.as<int64_t>(), ns, total, nullptr, d_off.as<int64_t>(),
                                                                                              c->d_bam_recoff.as<int64_t>(), c->d_bam_err.as<int>());
                    c->launches++;
                    if (nc_stream_wait(c) == cudaSuccess) { wo[0] = nrec_seg; wo[1] = 1; walked = true; }
                }
            }
            cudaStreamSynchronize(c->stream);
            if (rcw) return rcw;
        }
    }
    if (!walked) {
        bam_walk_kernel<<<1, 32, 0, c->stream>>>(c->d_bam.as<uint8_t>(), first, total, c->d_bam_recoff.as<int64_t>(), cap, c->d_bam_out.as<int64_t>());
        NC_LAUNCH_CHECK();
        NC_CUDA(cudaMemcpyAsync(wo, c->d_bam_out.p, sizeof(wo), cudaMemcpyDeviceToHost, c->stream));
        NC_CUDA(nc_stream_wait(c));
    }
    c->bam_walk_parallel = walked;
    if (!wo[1]) return fail(c, NC_EINVAL, "truncated alignment record");
    if (wo[0] > cap) return fail(c, NC_EOVERFLOW, "%lld records: more than the device reader's table holds (records shorter than 96 bytes on average)", (long long)wo[0]);
    const int64_t nrec = wo[0];
    const size_t nr = (size_t)std::max<int64_t>(nrec, 1);
    NC_CUDA(c->d_bam_rid.reserve(nr * 4)); NC_CUDA(c->d_bam_pos.reserve(nr * 4)); NC_CUDA(c->d_bam_flag.reserve(nr * 2)); NC_CUDA(c->d_bam_lseq.reserve(nr * 4));
    NC_CUDA(c->d_bam_ncig.reserve(nr * 4)); NC_CUDA(c->d_bam_nseq.reserve(nr * 4)); NC_CUDA(c->d_bam_cigsrc.reserve(nr * 8)); NC_CUDA(c->d_bam_seqsrc.reserve(nr * 8));
    NC_CUDA(c->d_bam_hp.reserve(nr)); NC_CUDA(c->d_bam_ps.reserve(nr * 4));
    std::vector<int32_t> rid((size_t)nrec);
    std::vector<int8_t> hp((size_t)nrec);
    if (nrec > 0) {
        BamFields bf = {c->d_bam_rid.as<int32_t>(), c->d_bam_pos.as<int32_t>(), c->d_bam_flag.as<uint16_t>(), c->d_bam_lseq.as<int32_t>(), c->d_bam_ncig.as<int32_t>(),
                        c->d_bam_nseq.as<int32_t>(), c->d_bam_cigsrc.as<int64_t>(), c->d_bam_seqsrc.as<int64_t>(), c->d_bam_hp.as<int8_t>(), c->d_bam_ps.as<int32_t>()};
        bam_fields_kernel<<<(unsigned)div_up(nrec, 128), 128, 0, c->stream>>>(c->d_bam.as<uint8_t>(), c->d_bam_recoff.as<int64_t>(), nrec, bf, c->d_bam_err.as<int>());
        NC_LAUNCH_CHECK();
        NC_CUDA(cudaMemcpyAsync(rid.data(), c->d_bam_rid.p, (size_t)nrec * 4, cudaMemcpyDeviceToHost, c->stream));
        NC_CUDA(cudaMemcpyAsync(hp.data(), c->d_bam_hp.p, (size_t)nrec, cudaMemcpyDeviceToHost, c->stream));
    }
    NC_CUDA(cudaEventRecord(e3, c->stream));
    NC_CUDA(nc_stream_wait(c));
    int herr[4] = {0, 0, 0, 0};
    NC_CUDA(cudaMemcpy(herr, c->d_bam_err.p, sizeof(herr), cudaMemcpyDeviceToHost));
    if (herr[1]) return fail(c, NC_EINVAL, "alignment record fields exceed its size (%d records)", herr[1]);
    // records are grouped by reference in a coordinate-sorted file (unmapped, refID -1, last)
    int32_t last = INT32_MIN;
    for (int64_t i = 0; i < nrec; i++) {
        const int32_t r = rid[(size_t)i];
        if (r >= 0 && r < (int32_t)contigs.size()) {
            if (last != r) {
                if (contigs[(size_t)r].n > 0 || (last >= 0 && r < last)) return fail(c, NC_EINVAL, "BAM is not coordinate-sorted");
                contigs[(size_t)r].first = i;
            }
            contigs[(size_t)r].n++;
            if (hp[(size_t)i] == 1 || hp[(size_t)i] == 2) contigs[(size_t)r].n_tagged++;
        }
        last = r;
    }
    c->bam_contigs.swap(contigs);
    c->bam_records = nrec; c->bam_bytes = total; c->bam_comp_bytes = (int64_t)f.n;
    c->bam_ms[0] = (float)(t1 - t0);
    NC_CUDA(cudaEventElapsedTime(&c->bam_ms[1], e0, e1));
    NC_CUDA(cudaEventElapsedTime(&c->bam_ms[2], e1, e2));
    NC_CUDA(cudaEventElapsedTime(&c->bam_ms[3], e2, e3));
    c->bam_open = true;
    *n_contigs = (int32_t)c->bam_contigs.size();
    return NC_OK;
}

int nc_bam_device_contig(nc_ctx* c, int32_t i, NcBamDeviceContig* out) {
    if (!c || !out) return NC_EINVAL;
    if (!c->bam_open) return fail(c, NC_ESTATE, "nc_bam_device_contig before nc_bam_device_open");
    if (i < 0 || i >= (int32_t)c->bam_contigs.size()) return fail(c, NC_EINVAL, "contig index %d out of range", i);
    const auto& bc = c->bam_contigs[(size_t)i];
    memset(out, 0, sizeof(*out));
    snprintf(out->name, sizeof(out->name), "%s", bc.name.c_str());
    out->length = bc.length; out->n_reads = bc.n; out->n_tagged = bc.n_tagged;
    return NC_OK;
}

int nc_bam_device_walk_mode(nc_ctx* c) { return (c && c->bam_open) ? (c->bam_walk_parallel ? 1 : 0) : NC_ESTATE; }

int nc_bam_device_timings(nc_ctx* c, float ms[4], int64_t* compressed_bytes, int64_t* inflated_bytes) {
    if (!c || !ms) return NC_EINVAL;
    for (int i = 0; i < 4; i++) ms[i] = c->bam_ms[i];
    if (compressed_bytes) *compressed_bytes = c->bam_comp_bytes;
    if (inflated_bytes) *inflated_bytes = c->bam_bytes;
    return NC_OK;
}

int nc_bam_device_stage(nc_ctx* c, int32_t i, const uint8_t* ref, int64_t ref_start, int64_t ref_len) {
    if (!c) return NC_EINVAL;
    if (!c->bam_open) return fail(c, NC_ESTATE, "nc_bam_device_stage before nc_bam_device_open");
    if (i < 0 || i >= (int32_t)c->bam_contigs.size()) return fail(c, NC_EINVAL, "contig index %d out of range", i);
    if (ref_len < 0 || ref_start < 0 || (ref_len > 0 && !ref)) return fail(c, NC_EINVAL, "nc_bam_device_stage: bad reference argument");
    if (ref_start + ref_len > 0x7fffff00ll) return fail(c, NC_EOVERFLOW, "contig coordinates must fit 31 bits");
    NC_CUDA(cudaSetDevice(c->device));
    c->staged = c->decoded = c->scanned = false;
    c->have_probs = false; c->tags_staged = c->indel_scanned = c->indel_built = false;
    const auto& bc = c->bam_contigs[(size_t)i];
    const int64_t n = bc.n, first = bc.first;
    int rc;
    NC_CUDA(c->d_pos.reserve((size_t)std::max<int64_t>(n, 1) * 4)); NC_CUDA(c->d_flag.reserve((size_t)std::max<int64_t>(n, 1) * 2));
    NC_CUDA(c->d_lseq.reserve((size_t)std::max<int64_t>(n, 1) * 4)); NC_CUDA(c->d_hp.reserve((size_t)std::max<int64_t>(n, 1))); NC_CUDA(c->d_ps.reserve((size_t)std::max<int64_t>(n, 1) * 4));
    NC_CUDA(c->d_cigar_off.reserve((size_t)(n + 1) * 8)); NC_CUDA(c->d_seq_off.reserve((size_t)(n + 1) * 8));
    int64_t n_cig = 0, n_seq = 0;
    if (n > 0) {
        NC_CUDA(cudaMemcpyAsync(c->d_pos.p, c->d_bam_pos.as<int32_t>() + first, (size_t)n * 4, cudaMemcpyDeviceToDevice, c->stream));
        NC_CUDA(cudaMemcpyAsync(c->d_flag.p, c->d_bam_flag.as<uint16_t>() + first, (size_t)n * 2, cudaMemcpyDeviceToDevice, c->stream));
        NC_CUDA(cudaMemcpyAsync(c->d_lseq.p, c->d_bam_lseq.as<int32_t>() + first, (size_t)n * 4, cudaMemcpyDeviceToDevice, c->stream));
        NC_CUDA(cudaMemcpyAsync(c->d_hp.p, c->d_bam_hp.as<int8_t>() + first, (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
        NC_CUDA(cudaMemcpyAsync(c->d_ps.p, c->d_bam_ps.as<int32_t>() + first, (size_t)n * 4, cudaMemcpyDeviceToDevice, c->stream));
        NC_CUDA(cudaMemsetAsync(c->d_bam_err.as<int>() + 2, 0, sizeof(int), c->stream));
        bam_sorted_kernel<<<(unsigned)div_up(n, 256), 256, 0, c->stream>>>(c->d_pos.as<int32_t>(), n, c->d_bam_err.as<int>()); NC_LAUNCH_CHECK();
        if ((rc = device_scan(c, c->d_bam_ncig.as<int32_t>() + first, n, c->d_cigar_off.as<int64_t>()))) return rc;
        if ((rc = device_scan(c, c->d_bam_nseq.as<int32_t>() + first, n, c->d_seq_off.as<int64_t>()))) return rc;
        if ((rc = read_i64(c, c->d_cigar_off.as<int64_t>() + n, &n_cig))) return rc;
        if ((rc = read_i64(c, c->d_seq_off.as<int64_t>() + n, &n_seq))) return rc;
        int herr = 0;
        NC_CUDA(cudaMemcpy(&herr, c->d_bam_err.as<int>() + 2, sizeof(int), cudaMemcpyDeviceToHost));
        if (herr) return fail(c, NC_EINVAL, "reads must be coordinate-sorted");
    } else {
        NC_CUDA(cudaMemsetAsync(c->d_cigar_off.p, 0, 8, c->stream)); NC_CUDA(cudaMemsetAsync(c->d_seq_off.p, 0, 8, c->stream));
    }
    NC_CUDA(c->d_cigar.reserve((size_t)std::max<int64_t>(n_cig, 4) * 4));
    NC_CUDA(c->d_seq4.reserve((size_t)std::max<int64_t>(n_seq, 16)));
    if (n > 0) {
        BamFields bf = {c->d_bam_rid.as<int32_t>(), c->d_bam_pos.as<int32_t>(), c->d_bam_flag.as<uint16_t>(), c->d_bam_lseq.as<int32_t>(), c->d_bam_ncig.as<int32_t>(),
                        c->d_bam_nseq.as<int32_t>(), c->d_bam_cigsrc.as<int64_t>(), c->d_bam_seqsrc.as<int64_t>(), c->d_bam_hp.as<int8_t>(), c->d_bam_ps.as<int32_t>()};
        bam_copy_kernel<<<(unsigned)div_up(n * 32, 128), 128, 0, c->stream>>>(c->d_bam.as<uint8_t>(), first, n, bf, c->d_cigar_off.as<int64_t>(), c->d_seq_off.as<int64_t>(),
                                                                             c->d_cigar.as<uint32_t>(), c->d_seq4.as<uint8_t>());
        NC_LAUNCH_CHECK();
    }
    if ((rc = upload(c, c->d_ref, ref, (size_t)ref_len))) return rc;
    c->n_reads = n; c->n_cigar = n_cig; c->n_seq = n_seq; c->ref_start = ref_start; c->ref_len = ref_len;
    c->staged = true; c->tags_staged = true;
    return NC_OK;
}

// ---- development probes (not part of the public header) -------------------------------------------------
int nc_debug_umma(nc_ctx* c, const void* a_img, int a_bytes, const void* b_img, int b_bytes, const void* prog, int n_ops,
                  int N, int ncols, float* out) {
    if (!c || !a_img || !b_img || !prog || !out || a_bytes % 16 || b_bytes % 16 || ncols % 16 || ncols > 128) return fail(c, NC_EINVAL, "nc_debug_umma: bad argument");
    NC_CUDA(cudaSetDevice(c->device));
    DevBuf da, db, dp, dout, derr;
    int rc = NC_OK;
    do {
        if ((rc = upload(c, da, a_img, a_bytes)) || (rc = upload(c, db, b_img, b_bytes)) || (rc = upload(c, dp, prog, (size_t)n_ops * sizeof(MmaOp)))) break;
        if (dout.reserve((size_t)128 * ncols * 4) != cudaSuccess || derr.reserve(16) != cudaSuccess) { rc = NC_ENOMEM; break; }
        cudaMemsetAsync(derr.p, 0, 16, c->stream);
        const int smem = a_bytes + b_bytes;
        if (cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) { rc = fail(c, NC_ECUDA, "probe smem"); break; }
        umma_probe_kernel<<<1, 128, smem, c->stream>>>(da.as<uint8_t>(), a_bytes, db.as<uint8_t>(), b_bytes, dp.as<MmaOp>(), n_ops, N, ncols,
                                                      dout.as<float>(), derr.as<int>());
        c->launches++;
        int herr = 0;
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(out, dout.p, (size_t)128 * ncols * 4, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&herr, derr.p, 4, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = nc_stream_wait(c);
        if (e != cudaSuccess) { rc = fail(c, NC_ECUDA, "nc_debug_umma: %s", cudaGetErrorString(e)); break; }
        if (herr) rc = fail(c, NC_ECUDA, "nc_debug_umma: mbarrier wait timed out");
    } while (0);
    da.release(); db.release(); dp.release(); dout.release(); derr.release();
    return rc;
}

// Runs the tensor-core trunk up to `stage` (1: after conv2, 2: after conv3) on fp32 inputs and returns the raw
// fp16 hi/lo activation image (layouts documented in nc_cnn_tc.cuh) for layer-by-layer parity tests.
int nc_debug_tc_trunk(nc_ctx* c, const float* x, int64_t n, int haploid, int stage, void* raw, size_t raw_bytes) {
    if (!c || !x || !raw || n <= 0 || (stage != 1 && stage != 2)) return fail(c, NC_EINVAL, "nc_debug_tc_trunk: bad argument");
    NC_CUDA(cudaSetDevice(c->device));
    Model& M = c->snp[haploid ? 1 : 0];
    if (!M.loaded || !M.tc.ready) return fail(c, NC_ESTATE, "nc_debug_tc_trunk: no tensor-core SNP model loaded");
    NC_CUDA(c->ws_x.reserve((size_t)n * NC_SNP_SITE_ELEMS * 4));
    NC_CUDA(cudaMemcpyAsync(c->ws_x.p, x, (size_t)n * NC_SNP_SITE_ELEMS * 4, cudaMemcpyHostToDevice, c->stream));
    uint64_t launches = 0;
    int rc = tc_forward_ex(c->stream, M.tc, 0, c->ws_x.p, NC_SNP_SITE_ELEMS, n, nullptr, nullptr, nullptr, nullptr, tail_weights(M),
                           nullptr, nullptr, c->sm_count, &launches, &c->err, stage);
    c->launches += launches;
    if (rc) return rc;
    const size_t have = stage == 1 ? (size_t)((n + 2) / 3) * tcg::C2_GROUP_BYTES : (size_t)((n + 127) / 128) * tcg::C3_TILE_BYTES;
    if (raw_bytes < have) return fail(c, NC_EINVAL, "nc_debug_tc_trunk: output buffer too small (%zu < %zu)", raw_bytes, have);
    NC_CUDA(cudaMemcpyAsync(raw, stage == 1 ? M.tc.c2.p : M.tc.c3.p, have, cudaMemcpyDeviceToHost, c->stream));
    NC_CUDA(nc_stream_wait(c));
    return tc_check(c, M);
}

// Validation switch for the tensor-core SNP kernels: on != 0 runs the instantiations that count activations sitting on the fp16
// saturation value (cvt.rn.satfinite clamps silently); *count (may be NULL) receives the count since the weights were loaded / last reset.
int nc_debug_saturation(nc_ctx* c, int haploid, int on, int reset, int64_t* count) {
    if (!c) return NC_EINVAL;
    NC_CUDA(cudaSetDevice(c->device));
    Model& M = c->snp[haploid ? 1 : 0];
    if (!M.loaded || !M.tc.ready) return fail(c, NC_ESTATE, "nc_debug_saturation: no tensor-core SNP model loaded");
    M.tc.audit = on != 0;
    NC_CUDA(c->pin.reserve(64));
    NC_CUDA(cudaMemcpyAsync(c->pin.p, M.tc.err.as<int>() + 1, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    NC_CUDA(nc_stream_wait(c));
    if (count) *count = *c->pin.as<int>();
    if (reset) NC_CUDA(cudaMemsetAsync(M.tc.err.as<int>() + 1, 0, sizeof(int), c->stream));
    return NC_OK;
}

// Copies the staged contig (nc_stage_reads / nc_bam_device_stage) back to the host: sizes first (any array pointer may be NULL).
int nc_debug_fetch_staged(nc_ctx* c, int64_t* n_reads, int64_t* n_cigar, int64_t* n_seq, int32_t* pos, uint16_t* flag, int64_t* cigar_off, uint32_t* cigar,
                          int64_t* seq_off, int32_t* l_seq, uint8_t* seq4, int8_t* hp, int32_t* ps) {
    if (!c) return NC_EINVAL;
    if (!c->staged) return fail(c, NC_ESTATE, "nc_debug_fetch_staged before staging");
    NC_CUDA(cudaSetDevice(c->device));
    if (n_reads) *n_reads = c->n_reads;
    if (n_cigar) *n_cigar = c->n_cigar;
    if (n_seq) *n_seq = c->n_seq;
    const size_t n = (size_t)c->n_reads;
    auto get = [&](void* dst, const DevBuf& src, size_t bytes) -> cudaError_t { return (dst && bytes) ? cudaMemcpyAsync(dst, src.p, bytes, cudaMemcpyDeviceToHost, c->stream) : cudaSuccess; };
    NC_CUDA(get(pos, c->d_pos, n * 4)); NC_CUDA(get(flag, c->d_flag, n * 2)); NC_CUDA(get(cigar_off, c->d_cigar_off, (n + 1) * 8));
    NC_CUDA(get(cigar, c->d_cigar, (size_t)c->n_cigar * 4)); NC_CUDA(get(seq_off, c->d_seq_off, (n + 1) * 8)); NC_CUDA(get(l_seq, c->d_lseq, n * 4));
    NC_CUDA(get(seq4, c->d_seq4, (size_t)c->n_seq));
    if (c->tags_staged) { NC_CUDA(get(hp, c->d_hp, n)); NC_CUDA(get(ps, c->d_ps, n * 4)); }
    NC_CUDA(nc_stream_wait(c));
    return NC_OK;
}

// Indel tensor-core trunk up to `stage` (1: c2 slabs after conv1 + conv2, 2: c3 after conv3) on fp32 inputs [n][H][128][2]; returns the raw
// fp16 hi/lo activation image (layouts in nc_cnn_tc_indel.cuh).  n <= 32768 (one batch).
int nc_debug_tci_trunk(nc_ctx* c, const float* x, int64_t n, int haploid, int stage, void* raw, size_t raw_bytes) {
    if (!c || !x || !raw || n <= 0 || n > 32768 || (stage != 1 && stage != 2)) return fail(c, NC_EINVAL, "nc_debug_tci_trunk: bad argument");
    NC_CUDA(cudaSetDevice(c->device));
    Model& M = c->indel[haploid ? 1 : 0];
    if (!M.loaded || !M.tc.ready) return fail(c, NC_ESTATE, "nc_debug_tci_trunk: no tensor-core indel model loaded");
    const int64_t site = (int64_t)M.Hin * M.Win * M.Cin;
    NC_CUDA(c->ws_x.reserve((size_t)n * site * 4));
    NC_CUDA(cudaMemcpyAsync(c->ws_x.p, x, (size_t)n * site * 4, cudaMemcpyHostToDevice, c->stream));
    uint64_t launches = 0;
    int rc = tci_forward(c->stream, M.tc, c->ws_x.as<float>(), site, n, tail_weights(M), nullptr, c->sm_count, &launches, &c->err, stage);
    c->launches += launches;
    if (rc) return rc;
    const size_t have = stage == 1 ? (size_t)n * tci::n_slabs(M.Hin) * tci::SLAB_BYTES : (size_t)n * tci::n_pos(M.Hin) * tci::C3_POS_BYTES;
    if (raw_bytes < have) return fail(c, NC_EINVAL, "nc_debug_tci_trunk: output buffer too small (%zu < %zu)", raw_bytes, have);
    NC_CUDA(cudaMemcpyAsync(raw, stage == 1 ? M.tc.c2.p : M.tc.c3.p, have, cudaMemcpyDeviceToHost, c->stream));
    NC_CUDA(nc_stream_wait(c));
    return tc_check(c, M);
}

}  // extern "C"
