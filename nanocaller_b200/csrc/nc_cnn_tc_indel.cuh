// Indel CNN forward on the 5th-generation tensor cores (impl = 0) for Indel_model / haploid_Indel_model
// (model_architect_indel.py:28-48, model_architect_indels_haploid.py:29-48): input [H][128][2] (H = 15: hstack of the hap0 / hap1 /
// all tensors, indelCaller.py:83; H = 5 haploid) -> conv1_{1,2,3} (8 channels each, 'same') -> conv2 2x3 stride (1,2) -> conv3 2x3
// stride (1,2) -> fc1 32 -> fc2 24 -> fc3 4 softmax / 1 sigmoid.
//
// Same arithmetic as the SNP kernels (nc_cnn_tc.cuh): every operand x = hi + lo in fp16, every product hi*hi + lo*hi + hi*lo in
// tcgen05.mma kind::f16 with fp32 accumulation in TMEM; activations as K-major no-swizzle planes [k-group][pixel row][8 x fp16] so
// that a convolution tap is a row shift of the operand descriptor.  Three kernels:
//
//   IA  conv1 + conv2, one site per CTA at a time.  The image is too big for "all of c1 in shared memory" (184 KB), so the site
//       is streamed by image row: conv1 produces one 128-pixel row per M = 128 tile (the 2-channel input is packed four pixels
//       to a K group, so a 5-tap kernel row is ONE K = 16 MMA), its epilogue drops the row into one of two 3-row windows, and as
//       soon as a window holds rows 2j..2j+2 the conv2 tile of output rows 2j, 2j+1 runs on it.  c1 never leaves the SM.
//   IB  conv3.  IA leaves c2 in HBM as "slabs" (4 output rows of conv3 = 5 rows of c2, the shared row stored twice) that ARE the
//       shared-memory plane image, so a slab arrives by one bulk copy; ring of slabs, producer warp + 3 consumer warpgroups.
//   IC  fc1 (K = 19,344 streamed by position, 128 sites per tile) + fc2 + fc3 + softmax / sigmoid.
#pragma once
#include "nc_cnn_tc.cuh"

namespace nc {
namespace tci {
constexpr int W = 128;                          // image width
constexpr int IN_PITCH = 136;                   // padded input pitch: 2 + 128 + 2, + reach of the 8-pixel K = 16 window
constexpr int IN_ROWS = 19 * IN_PITCH;          // H <= 15 image rows + 2 x 2 padding rows
constexpr int IN_PLANE = IN_ROWS * 16;          // 41,344 B per part (hi / lo)
constexpr int IN_COLS = 133;                    // plane rows per image row that can hold a non-zero pixel
constexpr int W1_TILE = 2 * 48 * 16;            // conv1, per kernel row: [k-group 2][w_hi 24 | w_lo 24][16 B]
constexpr int W1_BYTES = 5 * W1_TILE;           // 7,680
constexpr int W2_TILE = 2 * 64 * 16;            // conv2, per K = 16 chunk: [k-group 2][w_hi 32 | w_lo 32][16 B]
constexpr int W2_BYTES = 9 * W2_TILE;           // 18,432
constexpr int WIN_PLANE_ROWS = 3 * 64;          // c1 window: 3 image rows x 64 columns of one parity
constexpr int WIN_PLANE = WIN_PLANE_ROWS * 16;  // 3,072 B per (part, parity, k-group)
constexpr int WIN_BYTES = 12 * WIN_PLANE;       // 36,864 B
constexpr int SLAB_PLANE_ROWS = 5 * 32;         // c2 slab: 5 rows x 32 columns of one parity
constexpr int SLAB_PLANE = SLAB_PLANE_ROWS * 16;   // 2,560 B per (part, parity, k-group)
constexpr int SLAB_BYTES = 16 * SLAB_PLANE;     // 40,960 B
constexpr int W3_TILE = 2 * 96 * 16;            // conv3, per K = 16 chunk: [k-group 2][w_hi 48 | w_lo 48][16 B]
constexpr int W3_BYTES = 12 * W3_TILE;          // 36,864
constexpr int C3_POS_BYTES = 192;               // c3 in HBM: [site][position][part 2][48 channels fp16]
constexpr int WF_TILE = 2 * 64 * 16;            // fc1, per (position, 16-channel chunk): [k-group 2][w_hi 32 | w_lo 32][16 B]
constexpr int WF_POS_BYTES = 3 * WF_TILE;       // 6,144
__host__ __device__ constexpr int n_tiles2(int H) { return (H - 1) / 2; }           // conv2 tiles of two output rows
__host__ __device__ constexpr int n_slabs(int H) { return (H - 2 + 3) / 4; }        // conv3 tiles of four output rows
__host__ __device__ constexpr int n_pos(int H) { return (H - 2) * 31; }             // fc1 positions (conv3 output pixels)
}  // namespace tci

// ---- MMA programs -------------------------------------------------------------------------------------------------------
// conv1, image row h (A start = input plane row (h + kh) * IN_PITCH; K group 1 = four pixels further, LBO = 64 B):
//   a_hi x [w_hi | w_lo] (N = 48)  +  a_lo x (first 32 rows of the same tile: w_hi and 8 rows of w_lo, a harmless a_lo w_lo term)
// accumulator columns [0,24) = (a_hi + a_lo) w_hi, [24,48) = a_hi w_lo (+ lo lo in 24..31)
template <int KH>
__device__ __forceinline__ void issue_iconv1_row(uint32_t in16, uint32_t w16, uint32_t d) {
    const uint32_t a_hi = in16 + KH * tci::IN_PITCH + (4u << 16);
    const uint32_t a_lo = a_hi + tci::IN_PLANE / 16;
    const uint32_t b = (w16 + KH * (tci::W1_TILE / 16)) | (48u << 16);
    umma_f16(d, sdesc16(a_hi), sdesc16(b), make_idesc_f16(128, 48), KH > 0 ? 1u : 0u);
    umma_f16(d, sdesc16(a_lo), sdesc16(b), make_idesc_f16(128, 32), 1u);
}
__device__ __forceinline__ void issue_iconv1_tile(uint32_t in16_row, uint32_t w16, uint32_t d) {
    issue_iconv1_row<0>(in16_row, w16, d); issue_iconv1_row<1>(in16_row, w16, d); issue_iconv1_row<2>(in16_row, w16, d);
    issue_iconv1_row<3>(in16_row, w16, d); issue_iconv1_row<4>(in16_row, w16, d);
}
// conv2 on a 3-row window (tile rows m = h2l * 64 + w2).  Chunks 0..5: tap t = kh * 3 + kw, channels 0..15 (two k-group planes);
// chunks 6..8: channels 16..23 of two taps, the second K group addressed through LBO: (0,0)|(0,1), (0,2)|(1,0), (1,2)|(1,1).
template <int C>
__device__ __forceinline__ void issue_iconv2_chunk(uint32_t win16, uint32_t w16, uint32_t d) {
    constexpr int PR = tci::WIN_PLANE_ROWS;          // 192 rows per plane; plane index = (part * 2 + parity) * 3 + k-group
    uint32_t start, lbo;
    if constexpr (C < 6) {
        constexpr int kh = C / 3, kw = C % 3;
        start = ((kw & 1) * 3) * PR + kh * 64 + (kw >> 1); lbo = PR;
    } else if constexpr (C == 6) { start = (0 * 3 + 2) * PR + 0; lbo = 3 * PR; }              // (0,0) even plane -> (0,1) odd plane
    else if constexpr (C == 7) { start = (0 * 3 + 2) * PR + 1; lbo = 63; }                    // (0,2): row 0, col + 1 -> (1,0): row 1, col 0
    else { start = (0 * 3 + 2) * PR + 64 + 1; lbo = 3 * PR - 1; }                             // (1,2) even plane -> (1,1) odd plane
    const uint32_t a_hi = win16 + start + (lbo << 16);
    const uint32_t a_lo = a_hi + 6 * PR;
    const uint32_t b = (w16 + C * (tci::W2_TILE / 16)) | (64u << 16);
    umma_f16(d, sdesc16(a_hi), sdesc16(b), make_idesc_f16(128, 64), C > 0 ? 1u : 0u);
    umma_f16(d, sdesc16(a_lo), sdesc16(b), make_idesc_f16(128, 32), 1u);
}
template <int... I>
__device__ __forceinline__ void issue_iconv2_seq(uint32_t win16, uint32_t w16, uint32_t d, std::integer_sequence<int, I...>) {
    (issue_iconv2_chunk<I>(win16, w16, d), ...);
}
// conv3 on a slab (tile rows m = h3l * 32 + w3): chunk = (tap, 16-channel half); a_hi x [w_hi | w_lo] (N = 96) + a_lo x w_hi (N = 48)
template <int C>
__device__ __forceinline__ void issue_iconv3_chunk(uint32_t slab16, uint32_t w16, uint32_t d) {
    constexpr int PR = tci::SLAB_PLANE_ROWS;         // 160 rows per plane; plane index = (part * 2 + parity) * 4 + k-group
    constexpr int tap = C / 2, g = C % 2, kh = tap / 3, kw = tap % 3;
    const uint32_t a_hi = slab16 + ((kw & 1) * 4 + 2 * g) * PR + kh * 32 + (kw >> 1) + ((uint32_t)PR << 16);
    const uint32_t a_lo = a_hi + 8 * PR;
    const uint32_t b = (w16 + C * (tci::W3_TILE / 16)) | (96u << 16);
    umma_f16(d, sdesc16(a_hi), sdesc16(b), make_idesc_f16(128, 96), C > 0 ? 1u : 0u);
    umma_f16(d, sdesc16(a_lo), sdesc16(b), make_idesc_f16(128, 48), 1u);
}
template <int... I>
__device__ __forceinline__ void issue_iconv3_seq(uint32_t slab16, uint32_t w16, uint32_t d, std::integer_sequence<int, I...>) {
    (issue_iconv3_chunk<I>(slab16, w16, d), ...);
}

// ------------------------------------------------------------------------------------------------
// IA — conv1 + conv2
// ------------------------------------------------------------------------------------------------
struct IAParams {
    const float* x; int64_t site_stride;      // fp32 [site][H][128][2], stride in floats
    int64_t n_sites; int H;
    const uint8_t* wimg;                      // W1 tiles then W2 tiles
    const float* bias1; const float* bias2;   // 24, 32
    uint8_t* c2_out;                          // [site][slab][SLAB_BYTES]
    int* err;
};
constexpr int IA_THREADS = 2 * 128 + 32;                               // two epilogue warpgroups + the MMA issuer warp
// The kernel is bound by the epilogue's instruction stream (bias, SELU, fp16 split of 584 values per thread and site with ONE warp per
// scheduler: tensor pipe 13 % busy), so two warpgroups share every epilogue by channel groups — both read the same TMEM lanes (a warp's
// lanes are warp % 4), follow the same barrier / mbarrier sequence, and store different 16-byte channel groups: conv1 channels 0-15 | 16-23,
// conv2 channels 0-15 | 16-31, the input conversion by halves.  (Balancing the groups 5 : 5 changed nothing: the rest is the row
// protocol's serial MMA -> epilogue -> barrier chain.)
constexpr int IA_SMEM_W = tci::W1_BYTES + tci::W2_BYTES;               // 26,112
constexpr int IA_SMEM = IA_SMEM_W + 2 * tci::IN_PLANE + 2 * tci::WIN_BYTES + 256 + 2 * 56 * 4 + 64 + 64;
static_assert(IA_SMEM <= 232448, "IA shared memory exceeds the 227 KB per-CTA limit");
constexpr int IA_TMEM_C1 = 64;                                         // per conv1 buffer: 48 columns used
constexpr int IA_TMEM_C2 = 128;                                        // conv2 accumulators at column 128: 64 used

__device__ __forceinline__ void ia_barrier() { asm volatile("bar.sync 1, 288;" ::: "memory"); }

__global__ void __launch_bounds__(IA_THREADS, 1) tci_trunk_a_kernel(const IAParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* s_w = smem;
    uint8_t* s_in = smem + IA_SMEM_W;                                   // hi plane, lo plane
    uint8_t* s_win = s_in + 2 * tci::IN_PLANE;                          // two c1 windows (+ slack: junk rows read one row past the last plane)
    float* s_bias = reinterpret_cast<float*>(s_win + 2 * tci::WIN_BYTES + 256);     // [56] bias1 | bias2, then the same * log2(e)
    float* s_bl = s_bias + 56;
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_bl + 56);           // [0], [1] conv1 buffers, [2] conv2
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 4);

    const int tid = threadIdx.x, warp = tid >> 5, t = tid & 127;
    const int H = P.H, NT2 = tci::n_tiles2(H), NS = tci::n_slabs(H);
    for (int i = tid; i < IA_SMEM_W / 16; i += IA_THREADS) reinterpret_cast<uint4*>(s_w)[i] = __ldg(reinterpret_cast<const uint4*>(P.wimg) + i);
    if (tid < 24) { const float b = P.bias1[tid]; s_bias[tid] = b; s_bl[tid] = b * 1.4426950408889634f; }
    if (tid >= 32 && tid < 64) { const float b = P.bias2[tid - 32]; s_bias[24 + tid - 32] = b; s_bl[24 + tid - 32] = b * 1.4426950408889634f; }
    for (int i = tid; i < (2 * tci::IN_PLANE + 2 * tci::WIN_BYTES + 256) / 16; i += IA_THREADS) reinterpret_cast<uint4*>(s_in)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        for (int i = 0; i < 3; i++) mbar_init(&s_bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(s_tmem, 256);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    const uint32_t in16 = smem_u32(s_in) >> 4, win16 = smem_u32(s_win) >> 4, w16 = smem_u32(s_w) >> 4;

    const int wgi = tid >> 7;                                           // epilogue warpgroup 0 / 1 (2: the issuer warp)
    if (warp == 8) {
        // ===== MMA issuer warp: the epilogue warpgroups hand it work through the 288-thread barrier, in a fixed order
        for (int64_t site = blockIdx.x; site < P.n_sites; site += gridDim.x) {
            ia_barrier();                                                        // input planes written
            if (elect_one()) {
                tc_fence_after();
                issue_iconv1_tile(in16, w16, tmem); umma_commit(&s_bar[0]);
                if (H > 1) { issue_iconv1_tile(in16 + tci::IN_PITCH, w16, tmem + IA_TMEM_C1); umma_commit(&s_bar[1]); }
            }
            __syncwarp();
            for (int h = 0; h < H; h++) {
                ia_barrier();                                                    // accumulators of row h are in registers: the buffer is free
                if (h + 2 < H && elect_one()) {
                    tc_fence_after();
                    issue_iconv1_tile(in16 + (h + 2) * tci::IN_PITCH, w16, tmem + (h & 1) * IA_TMEM_C1); umma_commit(&s_bar[h & 1]);
                }
                __syncwarp();
                if (h >= 2 && !(h & 1)) {
                    ia_barrier();                                                // c1 rows h-2 .. h are in the window
                    if (elect_one()) {
                        tc_fence_after();
                        const int j = (h - 2) >> 1;
                        issue_iconv2_seq(win16 + (j & 1) * (tci::WIN_BYTES / 16), w16 + tci::W1_BYTES / 16, tmem + IA_TMEM_C2, std::make_integer_sequence<int, 9>{});
                        umma_commit(&s_bar[2]);
                    }
                    __syncwarp();
                }
            }
        }
        tc_fence_before();
        __syncthreads();
        return;
    }
    // ===== epilogue warpgroup: thread t owns tile row t (TMEM lane t)
    const uint32_t tmem_lane = tmem + ((uint32_t)(warp & 3) << 21);
    uint32_t ph1[2] = {0, 0}, ph2 = 0;
    bool ok = true;
    // conv2 epilogue of tile j: bias + SELU + split -> HBM slabs [site][slab][part][parity][k-group 4][5 rows x 32][16 B]
    auto conv2_epilogue = [&](int64_t site, int j) {
        ok = mbar_wait(&s_bar[2], ph2) && ok; ph2 ^= 1;
        tc_fence_after();
        float acc[16], lo_part[16];                                       // this warpgroup's 16 channels: columns 16 wgi .., and their w_lo partners
        tmem_ld16_nowait(tmem_lane + IA_TMEM_C2 + 16 * wgi, acc);
        tmem_ld16_nowait(tmem_lane + IA_TMEM_C2 + 32 + 16 * wgi, lo_part);
        tmem_ld_wait();
        tc_fence_before();
        const int h2 = 2 * j + (t >> 6), w2 = t & 63;
        if (w2 < 63 && h2 < H - 1) {
#pragma unroll
            for (int i = 0; i < 16; i++) acc[i] += lo_part[i];
            const int sl = h2 >> 2, r = h2 & 3;
            uint8_t* base = P.c2_out + (site * NS) * (int64_t)tci::SLAB_BYTES + ((w2 & 1) * 4) * tci::SLAB_PLANE + (w2 >> 1) * 16;
#pragma unroll
            for (int kgl = 0; kgl < 2; kgl++) {
                const int kg = 2 * wgi + kgl;
                uint4 hi, lo;
                act_split8(acc + 8 * kgl, s_bias + 24 + 8 * kg, s_bl + 24 + 8 * kg, hi, lo);
                if (sl < NS) {
                    uint8_t* d = base + (int64_t)sl * tci::SLAB_BYTES + kg * tci::SLAB_PLANE + r * 32 * 16;
                    *reinterpret_cast<uint4*>(d) = hi;
                    *reinterpret_cast<uint4*>(d + 8 * tci::SLAB_PLANE) = lo;
                }
                if (r == 0 && sl > 0) {                                   // the row two slabs share: fifth row of the previous one
                    uint8_t* d = base + (int64_t)(sl - 1) * tci::SLAB_BYTES + kg * tci::SLAB_PLANE + 4 * 32 * 16;
                    *reinterpret_cast<uint4*>(d) = hi;
                    *reinterpret_cast<uint4*>(d + 8 * tci::SLAB_PLANE) = lo;
                }
            }
        }
    };
    for (int64_t site = blockIdx.x; site < P.n_sites; site += gridDim.x) {
        // ---- input image -> padded fp16 hi / lo planes; plane row (hp, wp) holds pixels wp-2 .. wp+1 of image row hp-2 (2 channels each)
        {
            const float2* x = reinterpret_cast<const float2*>(P.x + site * P.site_stride);
            for (int idx = t + 128 * wgi; idx < H * tci::IN_COLS; idx += 256) {
                const int hh = idx / tci::IN_COLS, wp = idx - hh * tci::IN_COLS;
                float2 v[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int w = wp - 2 + i;
                    v[i] = (w >= 0 && w < tci::W) ? __ldg(x + hh * tci::W + w) : make_float2(0.f, 0.f);
                }
                uint4 hi, lo;
                split2(v[0].x, v[0].y, hi.x, lo.x); split2(v[1].x, v[1].y, hi.y, lo.y);
                split2(v[2].x, v[2].y, hi.z, lo.z); split2(v[3].x, v[3].y, hi.w, lo.w);
                const int row = (hh + 2) * tci::IN_PITCH + wp;
                *reinterpret_cast<uint4*>(s_in + row * 16) = hi;
                *reinterpret_cast<uint4*>(s_in + tci::IN_PLANE + row * 16) = lo;
            }
        }
        fence_async_smem();
        tc_fence_before();
        ia_barrier();                                                            // -> issuer: conv1 rows 0 and 1
        for (int h = 0; h < H; h++) {
            ok = mbar_wait(&s_bar[h & 1], ph1[h & 1]) && ok; ph1[h & 1] ^= 1;
            tc_fence_after();
            float v[16];                                                         // channel c = column c + column 24 + c; warpgroup 0: channels 0-15, 1: 16-23
            {
                float a[16], b[16];
                const uint32_t col = tmem_lane + (h & 1) * IA_TMEM_C1 + 16 * wgi;
                tmem_ld16_nowait(col, a); tmem_ld16_nowait(col + 24, b);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; i++) v[i] = a[i] + b[i];
            }
            tc_fence_before();
            ia_barrier();                                                        // -> issuer: conv1 row h + 2 into this buffer
            if (h >= 4 && !(h & 1)) conv2_epilogue(site, (h - 4) >> 1);          // also frees the window row h is about to enter
            {
                // c1 row h: pixel w = t, parity plane t & 1, column t >> 1; even rows enter two windows (last row of one, first of the next)
                const int kg0 = 2 * wgi, nkg = 2 - wgi;                          // channel groups of this warpgroup: {0, 1} | {2}
                uint4 hi[2], lo[2];
#pragma unroll
                for (int kgl = 0; kgl < 2; kgl++)
                    if (kgl < nkg) act_split8(v + 8 * kgl, s_bias + 8 * (kg0 + kgl), s_bl + 8 * (kg0 + kgl), hi[kgl], lo[kgl]);
                const int j = h >> 1;
                auto put = [&](int win, int slot) {
                    uint8_t* d = s_win + win * tci::WIN_BYTES + ((t & 1) * 3) * tci::WIN_PLANE + (slot * 64 + (t >> 1)) * 16;
#pragma unroll
                    for (int kgl = 0; kgl < 2; kgl++)
                        if (kgl < nkg) {
                            *reinterpret_cast<uint4*>(d + (kg0 + kgl) * tci::WIN_PLANE) = hi[kgl];
                            *reinterpret_cast<uint4*>(d + (6 + kg0 + kgl) * tci::WIN_PLANE) = lo[kgl];
                        }
                };
                if (h & 1) put(j & 1, 1);
                else {
                    if (j < NT2) put(j & 1, 0);
                    if (j > 0) put((j - 1) & 1, 2);
                }
            }
            if (h >= 2 && !(h & 1)) {
                fence_async_smem();
                tc_fence_before();
                ia_barrier();                                                    // -> issuer: conv2 tile (h - 2) / 2
            }
        }
        // conv2 tiles whose epilogue has not run inside the loop: tile j runs at row 2j + 4, so those with 2j + 4 >= H remain
        for (int j = 0; j < NT2; j++)
            if (2 * j + 4 >= H) conv2_epilogue(site, j);
    }
    if (!ok && t == 0) atomicExch(P.err, 1);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

// ------------------------------------------------------------------------------------------------
// IB — conv3 on c2 slabs
// ------------------------------------------------------------------------------------------------
struct IBParams {
    const uint8_t* c2; int64_t n_sites; int H;
    const uint8_t* wimg; const float* bias;   // 48
    uint8_t* c3_out;                          // [site][position][192 B]
    int* err;
};
#ifndef NC_IB_WGS
#define NC_IB_WGS 2             // as in TB: fewer consumers = more slab copies in flight out of the ring of four
#endif
constexpr int IB_WGS = NC_IB_WGS;
constexpr int IB_THREADS = IB_WGS * 128 + 32;
constexpr int IB_RING = 4;
constexpr int IB_SMEM_RING = IB_RING * tci::SLAB_BYTES + 256;
constexpr int IB_SMEM = tci::W3_BYTES + IB_SMEM_RING + 2 * 48 * 4 + 128 + 64;
static_assert(IB_SMEM <= 232448, "IB shared memory exceeds the 227 KB per-CTA limit");

__global__ void __launch_bounds__(IB_THREADS, 1) tci_trunk_b_kernel(const IBParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* s_w = smem;
    uint8_t* s_ring = smem + tci::W3_BYTES;
    float* s_bias = reinterpret_cast<float*>(s_ring + IB_SMEM_RING);
    float* s_bl = s_bias + 48;
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_bl + 48);       // [wg]
    uint64_t* s_full = s_bar + IB_WGS;
    uint64_t* s_empty = s_full + IB_RING;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_empty + IB_RING);
    const int tid = threadIdx.x, wg = tid >> 7, t = tid & 127, warp = tid >> 5, wq = warp & 3;
    const int H = P.H, NS = tci::n_slabs(H), H3 = H - 2, NPOS = tci::n_pos(H);

    for (int i = tid; i < tci::W3_BYTES / 16; i += IB_THREADS) reinterpret_cast<uint4*>(s_w)[i] = __ldg(reinterpret_cast<const uint4*>(P.wimg) + i);
    if (tid < 48) { const float b = P.bias[tid]; s_bias[tid] = b; s_bl[tid] = b * 1.4426950408889634f; }
    for (int i = tid; i < IB_SMEM_RING / 16; i += IB_THREADS) reinterpret_cast<uint4*>(s_ring)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        for (int i = 0; i < IB_WGS; i++) mbar_init(&s_bar[i], 1);
        for (int i = 0; i < IB_RING; i++) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(s_tmem, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const int64_t n_items = P.n_sites * NS;
    const int64_t gstride = (int64_t)gridDim.x * IB_WGS, gbase = (int64_t)blockIdx.x * IB_WGS;
    bool ok = true;

    if (wg == IB_WGS) {
        if (elect_one()) {
            for (int64_t i = 0;; i++) {
                const int64_t g = (i / IB_WGS) * gstride + gbase + (i % IB_WGS);
                if (g >= n_items) break;
                const int st = (int)(i % IB_RING);
                if (i >= IB_RING) ok = mbar_wait(&s_empty[st], (uint32_t)((i / IB_RING) - 1) & 1u) && ok;
                mbar_expect_tx(&s_full[st], tci::SLAB_BYTES);
                bulk_g2s(s_ring + st * tci::SLAB_BYTES, P.c2 + g * (int64_t)tci::SLAB_BYTES, tci::SLAB_BYTES, &s_full[st]);
            }
            if (!ok) atomicExch(P.err, 1);
        }
        __syncwarp();
        tc_fence_before();
        __syncthreads();
        return;
    }
    const uint32_t tmem = *s_tmem + (uint32_t)wg * 128u;            // columns [0,48) = (a_hi + a_lo) w_hi, [48,96) = a_hi w_lo
    const uint32_t tmem_lane = tmem + ((uint32_t)wq << 21);
    const uint32_t w16 = smem_u32(s_w) >> 4;
    uint32_t phase = 0;
    for (int64_t k = 0;; k++) {
        const int64_t i = k * IB_WGS + wg;
        const int64_t g = k * gstride + gbase + wg;
        if (g >= n_items) break;
        const int64_t site = g / NS;
        const int sl = (int)(g - site * NS);
        const int st = (int)(i % IB_RING);
        const uint32_t slab16 = smem_u32(s_ring + st * tci::SLAB_BYTES) >> 4;
        // the stage's previous user is another warpgroup: first make sure that use is over (see tc_trunk_b_kernel)
        if (i >= IB_RING) ok = mbar_wait(&s_empty[st], (uint32_t)((i / IB_RING) - 1) & 1u) && ok;
        ok = mbar_wait(&s_full[st], (uint32_t)(i / IB_RING) & 1u) && ok;
        tc_fence_before();
        wg_barrier(wg);                                              // the previous item's accumulators have been read by every warp
        if (wq == 0 && elect_one()) {
            tc_fence_after();
            issue_iconv3_seq(slab16, w16, tmem, std::make_integer_sequence<int, 12>{});
            umma_commit(&s_empty[st]);
            umma_commit(&s_bar[wg]);
        }
        __syncwarp();
        ok = mbar_wait(&s_bar[wg], phase) && ok; phase ^= 1;
        tc_fence_after();
        {
            const int h3 = 4 * sl + (t >> 5), w3 = t & 31;
            const bool valid = h3 < H3 && w3 < 31;
            uint8_t* dst = P.c3_out + (site * NPOS + (int64_t)(h3 * 31 + w3)) * tci::C3_POS_BYTES;
#pragma unroll 1
            for (int part3 = 0; part3 < 3; part3++) {               // 16 output channels at a time
                float acc[16], acc2[16];
                tmem_ld16_nowait(tmem_lane + part3 * 16, acc);
                tmem_ld16_nowait(tmem_lane + 48 + part3 * 16, acc2);
                tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int q = 0; q < 16; q++) acc[q] += acc2[q];
#pragma unroll
                    for (int g2 = 0; g2 < 2; g2++) {
                        uint4 hi, lo;
                        const int kg = part3 * 2 + g2;
                        act_split8(acc + 8 * g2, s_bias + 8 * kg, s_bl + 8 * kg, hi, lo);
                        *reinterpret_cast<uint4*>(dst + kg * 16) = hi;
                        *reinterpret_cast<uint4*>(dst + 96 + kg * 16) = lo;
                    }
                }
            }
        }
        tc_fence_before();
    }
    if (!ok && t == 0) atomicExch(P.err, 1);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(*s_tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// IC — fc1 (K streamed by position) + fc2 + fc3 + softmax / sigmoid, 128 sites per CTA
// ------------------------------------------------------------------------------------------------
struct ICParams {
    const uint8_t* c3; int64_t n_sites; int H;
    const uint8_t* wimg;            // [position][3 chunks][WF_TILE]
    const float* bias;              // fc1 bias (32)
    TailW tail; int haploid;
    float* out;                     // [n][4] / [n][1]
    int* err;
};
constexpr int IC_STAGES = 4;
constexpr int IC_STAGE_A = 2 * 6 * 2048;                    // [part][k-group 6][128 sites][16 B]
constexpr int IC_STAGE = IC_STAGE_A + tci::WF_POS_BYTES;    // 30,720
constexpr int IC_SMEM = IC_STAGES * IC_STAGE + 32 * 4 + 64 + 64;
// The tensor core adds every MMA into the fp32 accumulator with truncation, and fc1 chains 6 MMAs per position over 403 positions:
// one accumulator for the whole K range measured 1.7e-4 on the probabilities (the SNP fc1, 216 MMAs, 2e-5).  So K is cut into segments
// of IC_SEG positions that accumulate in two alternating TMEM buffers; finished segments are added up in registers (fp32, round to
// nearest) while the next one runs.
constexpr int IC_SEG = 13;

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

__global__ void __launch_bounds__(128, 1) tci_fc_kernel(const ICParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    float* s_bias = reinterpret_cast<float*>(smem + IC_STAGES * IC_STAGE);
    uint64_t* s_empty = reinterpret_cast<uint64_t*>(s_bias + 32);      // [stage] MMAs that read the stage have completed
    uint64_t* s_seg = s_empty + IC_STAGES;                             // [2] MMAs of the segment in this accumulator buffer have completed
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_seg + 2);
    const int t = threadIdx.x, warp = t >> 5;
    const int NPOS = tci::n_pos(P.H), NSEG = (NPOS + IC_SEG - 1) / IC_SEG;
    if (t < 32) s_bias[t] = P.bias[t];
    if (t == 0) {
        for (int i = 0; i < IC_STAGES; i++) mbar_init(&s_empty[i], 1);
        mbar_init(&s_seg[0], 1); mbar_init(&s_seg[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(s_tmem, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;                                     // per buffer: columns [0,32) = (a_hi + a_lo) w_hi, [32,64) = a_hi w_lo
    const uint32_t tmem_lane = tmem + ((uint32_t)(warp & 3) << 21);
    bool ok = true;
    uint32_t seg_phase[2] = {0, 0};
    int64_t fills = 0;                                                 // stage fills issued by this CTA so far (for the empty-barrier parity)

    for (int64_t tile = blockIdx.x; tile * 128 < P.n_sites; tile += gridDim.x) {
        const int64_t site = (tile * 128 + t < P.n_sites) ? tile * 128 + t : P.n_sites - 1;    // rows past the end re-read the last site (discarded)
        const uint8_t* src = P.c3 + site * (int64_t)NPOS * tci::C3_POS_BYTES;
        float x[32];
#pragma unroll
        for (int i = 0; i < 32; i++) x[i] = 0.f;
        auto flush = [&](int buf) {                                    // finished segment -> registers
            ok = mbar_wait(&s_seg[buf], seg_phase[buf]) && ok; seg_phase[buf] ^= 1;
            tc_fence_after();
            float a[32], b[32];
            tmem_ld16_nowait(tmem_lane + buf * 64, a); tmem_ld16_nowait(tmem_lane + buf * 64 + 16, a + 16);
            tmem_ld16_nowait(tmem_lane + buf * 64 + 32, b); tmem_ld16_nowait(tmem_lane + buf * 64 + 48, b + 16);
            tmem_ld_wait();
            tc_fence_before();
#pragma unroll
            for (int i = 0; i < 32; i++) x[i] += a[i] + b[i];
        };
        auto fill = [&](int pos) {
            const int st = (int)(fills % IC_STAGES);
            if (fills >= IC_STAGES) ok = mbar_wait(&s_empty[st], (uint32_t)((fills / IC_STAGES) - 1) & 1u) && ok;
            uint8_t* dst = smem + st * IC_STAGE;
            const uint8_t* s = src + (int64_t)pos * tci::C3_POS_BYTES;
#pragma unroll
            for (int q = 0; q < 12; q++) cp_async16(dst + (q * 128 + t) * 16, s + q * 16);          // (part, k-group) q of this thread's site
            const uint8_t* wsrc = P.wimg + (int64_t)pos * tci::WF_POS_BYTES;
#pragma unroll
            for (int q = 0; q < 3; q++) cp_async16(dst + IC_STAGE_A + (q * 128 + t) * 16, wsrc + (q * 128 + t) * 16);
            fills++;
        };
        const int64_t fills0 = fills;
        for (int p = 0; p < IC_STAGES - 1 && p < NPOS; p++) { fill(p); cp_async_commit(); }
        for (int pos = 0; pos < NPOS; pos++) {
            const int seg = pos / IC_SEG, buf = seg & 1;
            const bool first = pos - seg * IC_SEG == 0, last = (pos + 1 == NPOS) || (pos + 1 - seg * IC_SEG == IC_SEG);
            if (first && seg >= 2) flush(buf);                         // segment seg - 2 used this buffer; it finished long ago
            if (pos + IC_STAGES - 1 < NPOS) fill(pos + IC_STAGES - 1);
            cp_async_commit();                                         // one group per iteration (possibly empty): position `pos` is complete below
            cp_async_wait_group<IC_STAGES - 1>();
            fence_async_smem();
            __syncthreads();
            if (t == 0) {
                tc_fence_after();
                const int st = (int)((fills0 + pos) % IC_STAGES);
                const uint32_t sb16 = smem_u32(smem + st * IC_STAGE) >> 4;
                const uint32_t d = tmem + buf * 64;
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const uint32_t a_hi = sb16 + (2 * c) * 128 + (128u << 16), a_lo = a_hi + 6 * 128;
                    const uint32_t b = (sb16 + IC_STAGE_A / 16 + c * (tci::WF_TILE / 16)) | (64u << 16);
                    umma_f16(d, sdesc16(a_hi), sdesc16(b), make_idesc_f16(128, 64), (!first || c > 0) ? 1u : 0u);
                    umma_f16(d, sdesc16(a_lo), sdesc16(b), make_idesc_f16(128, 32), 1u);
                }
                umma_commit(&s_empty[st]);
                if (last) umma_commit(&s_seg[buf]);
            }
        }
        if (NSEG >= 2) flush((NSEG - 2) & 1);
        flush((NSEG - 1) & 1);
        {
            const int64_t s = tile * 128 + t;
            if (s < P.n_sites) {
                float h[24];
#pragma unroll
                for (int i = 0; i < 32; i++) x[i] = selu_f(x[i] + s_bias[i]);
                dense_t<32, 24>(x, P.tail.fc2_k, P.tail.fc2_b, h, true);
                if (P.haploid) {
                    float z[1];
                    dense_t<24, 1>(h, P.tail.fc3_k, P.tail.fc3_b, z, false);
                    P.out[s] = 1.f / (1.f + expf(-z[0]));
                } else {
                    float z[4];
                    dense_t<24, 4>(h, P.tail.fc3_k, P.tail.fc3_b, z, false);
                    softmax_t<4>(z);
#pragma unroll
                    for (int j = 0; j < 4; j++) P.out[s * 4 + j] = z[j];
                }
            }
        }
        __syncthreads();                                               // every warp has read its accumulator rows before the next tile's first MMA
    }
    if (!ok) atomicExch(P.err, 1);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// Builds the operand images of an indel model (kind 2 diploid, 3 haploid) into T (wimg_a = conv1 + conv2, wimg_b = conv3,
// wimg_c = fc1, bias = b1(24) b2(32) b3(48) bf(32)).
inline int tci_model_prepare(cudaStream_t stream, TcModel& T, int kind, const float* blob, size_t n_floats, std::string* err) {
    using namespace tci;
    T.ready = false; T.kind = kind;
    const int H = kind == 2 ? 15 : 5, NPOS = n_pos(H);
    const float* w11 = blob; const float* b11 = w11 + 80;            // [1][5][2][8]
    const float* w12 = b11 + 8; const float* b12 = w12 + 80;         // [5][1][2][8]
    const float* w13 = b12 + 8; const float* b13 = w13 + 400;        // [5][5][2][8]
    const float* w2 = b13 + 8; const float* b2 = w2 + 2 * 3 * 24 * 32;
    const float* w3 = b2 + 32; const float* b3 = w3 + 2 * 3 * 32 * 48;
    const float* wf = b3 + 48; const float* bf = wf + (size_t)NPOS * 48 * 32;
    if ((size_t)(bf + 32 - blob) > n_floats) { if (err) *err = "indel weight blob too short"; return NC_EINVAL; }
    auto put = [](std::vector<uint8_t>& img, size_t off, float x, int part) {
        const __half h = __float2half_rn(x);
        const __half l = __float2half_rn(x - __half2float(h));
        memcpy(img.data() + off, part ? &l : &h, 2);
    };
    // element (kslot, n) of a [N rows][K = 16] tile: ((kslot / 8) * N + n) * 16 + (kslot % 8) * 2
    auto off_of = [](int N, int kslot, int n) { return ((size_t)(kslot >> 3) * N + n) * 16 + (size_t)(kslot & 7) * 2; };
    std::vector<uint8_t> w1_img(W1_BYTES, 0);
    for (int kh = 0; kh < 5; kh++)
        for (int kw = 0; kw < 5; kw++)
            for (int ci = 0; ci < 2; ci++)
                for (int part = 0; part < 2; part++)
                    for (int co = 0; co < 24; co++) {
                        float x = 0.f;
                        if (co < 8) { if (kh == 2) x = w11[(kw * 2 + ci) * 8 + co]; }
                        else if (co < 16) { if (kw == 2) x = w12[(kh * 2 + ci) * 8 + (co - 8)]; }
                        else x = w13[((kh * 5 + kw) * 2 + ci) * 8 + (co - 16)];
                        put(w1_img, (size_t)kh * W1_TILE + off_of(48, 2 * kw + ci, part * 24 + co), x, part);
                    }
    // conv2 [kh 2][kw 3][ci 24][co 32]
    std::vector<uint8_t> w2_img(W2_BYTES, 0);
    auto w2_at = [&](int tap, int ci, int co) { return w2[((size_t)tap * 24 + ci) * 32 + co]; };
    static const int pair_tap[3][2] = {{0, 1}, {2, 3}, {5, 4}};          // chunks 6..8: (first K group tap, second K group tap)
    for (int c = 0; c < 9; c++)
        for (int ks = 0; ks < 16; ks++)
            for (int part = 0; part < 2; part++)
                for (int co = 0; co < 32; co++) {
                    int tap, ci;
                    if (c < 6) { tap = c; ci = ks; }
                    else { tap = pair_tap[c - 6][ks >> 3]; ci = 16 + (ks & 7); }
                    put(w2_img, (size_t)c * W2_TILE + off_of(64, ks, part * 32 + co), w2_at(tap, ci, co), part);
                }
    // conv3 [kh 2][kw 3][ci 32][co 48]
    std::vector<uint8_t> w3_img(W3_BYTES, 0);
    for (int c = 0; c < 12; c++)
        for (int ks = 0; ks < 16; ks++)
            for (int part = 0; part < 2; part++)
                for (int co = 0; co < 48; co++)
                    put(w3_img, (size_t)c * W3_TILE + off_of(96, ks, part * 48 + co), w3[((size_t)(c / 2) * 32 + 16 * (c % 2) + ks) * 48 + co], part);
    // fc1 [k = pos * 48 + ch][32]
    std::vector<uint8_t> wf_img((size_t)NPOS * WF_POS_BYTES, 0);
    for (int pos = 0; pos < NPOS; pos++)
        for (int c = 0; c < 3; c++)
            for (int ks = 0; ks < 16; ks++)
                for (int part = 0; part < 2; part++)
                    for (int co = 0; co < 32; co++)
                        put(wf_img, (size_t)pos * WF_POS_BYTES + (size_t)c * WF_TILE + off_of(64, ks, part * 32 + co),
                            wf[((size_t)pos * 48 + 16 * c + ks) * 32 + co], part);
    std::vector<uint8_t> img_a;
    img_a.insert(img_a.end(), w1_img.begin(), w1_img.end());
    img_a.insert(img_a.end(), w2_img.begin(), w2_img.end());
    std::vector<float> bias;
    bias.insert(bias.end(), b11, b11 + 8); bias.insert(bias.end(), b12, b12 + 8); bias.insert(bias.end(), b13, b13 + 8);
    bias.insert(bias.end(), b2, b2 + 32); bias.insert(bias.end(), b3, b3 + 48); bias.insert(bias.end(), bf, bf + 32);
    auto up = [&](DevBuf& d, const void* h, size_t bytes) -> cudaError_t {
        cudaError_t e = d.reserve(bytes);
        if (e != cudaSuccess) return e;
        return cudaMemcpyAsync(d.p, h, bytes, cudaMemcpyHostToDevice, stream);
    };
    cudaError_t e;
    if ((e = up(T.wimg_a, img_a.data(), img_a.size())) != cudaSuccess || (e = up(T.wimg_b, w3_img.data(), w3_img.size())) != cudaSuccess ||
        (e = up(T.wimg_c, wf_img.data(), wf_img.size())) != cudaSuccess || (e = up(T.bias, bias.data(), bias.size() * 4)) != cudaSuccess ||
        (e = T.err.reserve(16)) != cudaSuccess || (e = cudaMemsetAsync(T.err.p, 0, 16, stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(stream)) != cudaSuccess) {
        if (err) *err = std::string("tci_model_prepare: ") + cudaGetErrorString(e);
        return NC_ECUDA;
    }
    T.ready = true;
    return NC_OK;
}

// IA -> IB -> IC over n sites in batches (c2 is 160 KB per site).  x: fp32 [n][H][128][2] on the device, site stride in floats.
// stop_after: 0 full forward, 1 after IA, 2 after IB (debug entry points; then only the first batch runs).
inline int tci_forward(cudaStream_t stream, TcModel& T, const float* x, int64_t site_stride, int64_t n, const TailW& tw, float* out,
                       int sm_count, uint64_t* launches, std::string* err, int stop_after = 0) {
    using namespace tci;
    if (!T.ready || T.kind < 2) return NC_ESTATE;
    if (n <= 0) return NC_OK;
    const int H = T.kind == 2 ? 15 : 5, NS = n_slabs(H), NPOS = n_pos(H);
    auto cuda_fail = [&](cudaError_t e, const char* what) { if (err) *err = std::string(what) + ": " + cudaGetErrorString(e); return NC_ECUDA; };
    cudaError_t e;
    const int64_t NB = std::min<int64_t>(n, 32768);                  // 32 k sites: 5.4 GB of c2, 2.5 GB of c3
    if ((e = T.c2.reserve((size_t)NB * NS * SLAB_BYTES)) != cudaSuccess) return cuda_fail(e, "indel c2 alloc");
    if ((e = T.c3.reserve((size_t)NB * NPOS * C3_POS_BYTES)) != cudaSuccess) return cuda_fail(e, "indel c3 alloc");
    static bool attr_set[64] = {};
    int dev = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    const bool tracked = dev >= 0 && dev < 64;
    if (!tracked || !attr_set[dev]) {
        if ((e = cudaFuncSetAttribute(tci_trunk_a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, IA_SMEM)) != cudaSuccess) return cuda_fail(e, "IA smem attr");
        if ((e = cudaFuncSetAttribute(tci_trunk_b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, IB_SMEM)) != cudaSuccess) return cuda_fail(e, "IB smem attr");
        if ((e = cudaFuncSetAttribute(tci_fc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, IC_SMEM)) != cudaSuccess) return cuda_fail(e, "IC smem attr");
        if (tracked) attr_set[dev] = true;
    }
    const float* bias = T.bias.as<float>();
    const int nout = T.kind == 3 ? 1 : 4;
    for (int64_t b0 = 0; b0 < n; b0 += NB) {
        const int64_t nb = std::min(NB, n - b0);
        IAParams pa = {};
        pa.x = x + b0 * site_stride; pa.site_stride = site_stride; pa.n_sites = nb; pa.H = H; pa.wimg = T.wimg_a.as<uint8_t>();
        pa.bias1 = bias; pa.bias2 = bias + 24; pa.c2_out = T.c2.as<uint8_t>(); pa.err = T.err.as<int>();
        tci_trunk_a_kernel<<<(unsigned)std::min<int64_t>(nb, sm_count), IA_THREADS, IA_SMEM, stream>>>(pa);
        if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail(e, "IA launch");
        (*launches)++;
        if (stop_after == 1) return NC_OK;
        IBParams pb = {};
        pb.c2 = T.c2.as<uint8_t>(); pb.n_sites = nb; pb.H = H; pb.wimg = T.wimg_b.as<uint8_t>(); pb.bias = bias + 56;
        pb.c3_out = T.c3.as<uint8_t>(); pb.err = T.err.as<int>();
        const int64_t items = nb * NS;
        tci_trunk_b_kernel<<<(unsigned)std::min<int64_t>((items + IB_WGS - 1) / IB_WGS, sm_count), IB_THREADS, IB_SMEM, stream>>>(pb);
        if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail(e, "IB launch");
        (*launches)++;
        if (stop_after == 2) return NC_OK;
        ICParams pc = {};
        pc.c3 = T.c3.as<uint8_t>(); pc.n_sites = nb; pc.H = H; pc.wimg = T.wimg_c.as<uint8_t>(); pc.bias = bias + 104;
        pc.tail = tw; pc.haploid = T.kind == 3; pc.out = out + b0 * nout; pc.err = T.err.as<int>();
        tci_fc_kernel<<<(unsigned)std::min<int64_t>((nb + 127) / 128, sm_count), 128, IC_SMEM, stream>>>(pc);
        if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail(e, "IC launch");
        (*launches)++;
    }
    return NC_OK;
}

}  // namespace nc
