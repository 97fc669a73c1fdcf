// CNN forward on the 5th-generation tensor cores (impl = 0): tcgen05.mma kind::f16 with fp32
// accumulation in TMEM, for SNP_model / haploid_SNP_model (model_architect.py:36-64,
// model_architect_SNP_haploid.py:33-53).
//
// Precision.  Single-pass bf16/fp16/tf32 operands miss the 1e-4 probability tolerance by 10-100x
// (measured).  Every operand is therefore split x = hi + lo with hi = fp16(x),
// lo = fp16(x - hi) (22 mantissa bits) and every product is three MMAs: hi*hi + lo*hi + hi*lo.
//
// Convolutions as tap-decomposed implicit GEMMs.  Activations live in shared memory as "planes"
// [k-group of 8 channels][pixel row][8 x fp16 = 16 B]: the canonical K-major no-swizzle UMMA layout
// with SBO = 128 B, so GEMM row r of a core matrix group is simply 16 B further.  A convolution
// tap is then nothing but a different start address (row shift), the two 8-wide K groups of one
// K=16 MMA are addressed through LBO (another plane, or another tap of the same plane), and no
// im2col copy is ever materialised.  Stride-2 layers read from planes split by column parity.
//
//   TA  conv1_{1,2,3} + conv2   one warpgroup per site, two warpgroups per CTA, weights resident in smem
//   TB  conv3                   three sites per 128-row tile
//   TC  fc1 + heads             128 sites per tile, K streamed position by position (double buffered)
#pragma once
#include <cuda_fp16.h>

#include <utility>

#include "nc_common.cuh"
#include "nc_cnn.cuh"

namespace nc {

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle, version 1 (Blackwell) shared-memory matrix descriptor.
__device__ __forceinline__ uint64_t make_sdesc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: D fp32, A/B fp16, both K-major, dense.
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
// Bounded wait: a protocol bug must not hang the GPU box.  Returns false on timeout.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t spin = 0; spin < (1u << 24); spin++) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}\n"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t addr, float* r) {
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(addr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) r[i] = __uint_as_float(v[i]);
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t addr, float* r) {
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(addr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) r[i] = __uint_as_float(v[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xFFFFFFFF;\n\t@px mov.s32 %0, 1;\n\t}\n" : "+r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared through the TMA engine, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wg_barrier(int wg) { asm volatile("bar.sync %0, 128;" :: "r"(wg + 1) : "memory"); }
// the four warps of an epilogue warpgroup plus its MMA issuer warp
__device__ __forceinline__ void wg_issuer_barrier(int wg) { asm volatile("bar.sync %0, 160;" :: "r"(wg + 4) : "memory"); }

// ---- epilogue arithmetic.  The CUDA-core epilogues (bias, SELU, fp16 hi/lo split of ~130 values per thread and site)
// are what bounds TA once the MMA programs are folded, so every instruction counts:
//   * exp through ex2.approx.ftz (no denormal range fix-up: -3 instructions per value), log2(e) folded into an FMA
//     on the raw accumulator (bl = bias * log2 e is precomputed);
//   * hi = x with the low 13 mantissa bits cleared (one LOP3; exactly representable in fp16), lo = fp16(x - hi): 21+
//     significant bits like the round-to-nearest split, without converting hi back to fp32;
//   * cvt.rn.satfinite packs two values per instruction and keeps hi finite without separate clamps.
__device__ __forceinline__ float ex2_ftz(float x) { float e; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x)); return e; }
// SELU(acc + bias): sa*(exp(min(z,0)) - 1) + scale*max(z,0), branch-free, exact 0 from the exponential branch for z > 0
__device__ __forceinline__ float selu_acc(float acc, float bias, float bl) {
    const float scale = 1.0507009873554805f, sa = 1.0507009873554805f * 1.6732632423543772f, l2e = 1.4426950408889634f;
    const float e = ex2_ftz(fminf(fmaf(acc, l2e, bl), 0.f));
    return fmaf(scale, fmaxf(acc + bias, 0.f), fmaf(sa, e, -sa));
}
__device__ __forceinline__ float selu_bf(float x) { return selu_acc(x, 0.f, 0.f); }
// (a -> low half, b -> high half), saturating to the largest finite fp16
__device__ __forceinline__ uint32_t pack_sat(float a, float b) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
// two values -> packed (hi, hi) and (lo, lo) fp16 pairs
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const float ah = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u), bh = __uint_as_float(__float_as_uint(b) & 0xFFFFE000u);
    hi = pack_sat(ah, bh);
    lo = pack_sat(a - ah, b - bh);
}
// Saturation audit (validation builds of the kernels, template flag SAT): cvt.rn.satfinite clamps |x| > 65504 to the largest finite
// fp16 silently; count the hi halves that sit exactly on that value so that a test can assert the clamp never engages.
__device__ __forceinline__ uint32_t sat_halves(const uint4& hi) {
    const uint32_t k = 0x7BFF7BFFu, m = 0x7FFF7FFFu;
    return (__vcmpeq2(hi.x & m, k) | __vcmpeq2(hi.y & m, k) | __vcmpeq2(hi.z & m, k) | __vcmpeq2(hi.w & m, k)) != 0u ? 1u : 0u;
}
// 8 accumulator values -> bias + SELU -> hi / lo 16-byte words.  bias2 = [bias | bias * log2 e] rows of 8 floats: bias2[0..7], bias2[n_bias..]
__device__ __forceinline__ void act_split8_nobias(const float* acc, uint4& hi, uint4& lo) {      // bias already inside the accumulator
    split2(selu_acc(acc[0], 0.f, 0.f), selu_acc(acc[1], 0.f, 0.f), hi.x, lo.x);
    split2(selu_acc(acc[2], 0.f, 0.f), selu_acc(acc[3], 0.f, 0.f), hi.y, lo.y);
    split2(selu_acc(acc[4], 0.f, 0.f), selu_acc(acc[5], 0.f, 0.f), hi.z, lo.z);
    split2(selu_acc(acc[6], 0.f, 0.f), selu_acc(acc[7], 0.f, 0.f), hi.w, lo.w);
}
__device__ __forceinline__ void act_split8(const float* acc, const float* bias, const float* bl, uint4& hi, uint4& lo) {
    const float4 b0 = *reinterpret_cast<const float4*>(bias), b1 = *reinterpret_cast<const float4*>(bias + 4);
    const float4 l0 = *reinterpret_cast<const float4*>(bl), l1 = *reinterpret_cast<const float4*>(bl + 4);
    split2(selu_acc(acc[0], b0.x, l0.x), selu_acc(acc[1], b0.y, l0.y), hi.x, lo.x);
    split2(selu_acc(acc[2], b0.z, l0.z), selu_acc(acc[3], b0.w, l0.w), hi.y, lo.y);
    split2(selu_acc(acc[4], b1.x, l1.x), selu_acc(acc[5], b1.y, l1.y), hi.z, lo.z);
    split2(selu_acc(acc[6], b1.z, l1.z), selu_acc(acc[7], b1.w, l1.w), hi.w, lo.w);
}

// One MMA of a layer's "program": where its A rows start, how its two K groups are spaced, which
// B tile it multiplies and which accumulator columns it adds into.
struct MmaOp { uint32_t a_off, a_lbo, b_off, misc; };     // misc: bits 0-9 D column, bit 14 B LBO = 0, bit 15 accumulate, bits 16-23 N/8 override
constexpr uint32_t kOpAcc = 1u << 15;
constexpr uint32_t kOpBLbo0 = 1u << 14;

__device__ __forceinline__ void issue_program(const MmaOp* prog, int n_ops, uint32_t a_base, uint32_t w_base, uint32_t b_lbo,
                                              uint32_t d_base, uint32_t idesc) {
    for (int i = 0; i < n_ops; i++) {
        const MmaOp op = prog[i];
        const uint32_t n8 = (op.misc >> 16) & 0xFFu;
        const uint32_t id = n8 ? ((idesc & ~(0x3Fu << 17)) | (n8 << 17)) : idesc;
        umma_f16(d_base + (op.misc & 0x3FFu), make_sdesc(a_base + op.a_off, op.a_lbo, 128u),
                 make_sdesc(w_base + op.b_off, (op.misc & kOpBLbo0) ? 0u : b_lbo, 128u), id, (op.misc & kOpAcc) ? 1u : 0u);
    }
}

// ------------------------------------------------------------------------------------------------
// geometry of the SNP trunk
// ------------------------------------------------------------------------------------------------
namespace tcg {
constexpr int WP = 45;                         // padded input width (41 + 2*2)
constexpr int IN_ROWS = 416;                   // 9*45 = 405 padded pixels + slack for junk rows
constexpr int IN_PLANE = IN_ROWS * 16;         // bytes of one input plane (8 channels of fp16 per pixel)
constexpr int C1_PITCH = 21;                   // columns per parity plane of c1 (even: 21 real, odd: 20 real + 1 junk)
constexpr int C1_ROWS = 5 * C1_PITCH;          // 105
constexpr int C1_PLANE = C1_ROWS * 16;         // 1680 B per (part, parity, k-group)
constexpr int C1_BYTES = 2 * 2 * 6 * C1_PLANE + 1024;
constexpr int TILE1_START = 96;                // second conv1 tile covers output rows 96..223 (valid: 128..220), so its first warp has no valid row
constexpr int C2_PITCH = 10;
constexpr int C2_SITE_ROWS = 4 * C2_PITCH;     // 40 rows per site and parity
constexpr int C2_CHUNK = C2_SITE_ROWS * 16;    // 640 B per (part, parity, k-group)
constexpr int C2_SITE_BYTES = 2 * 2 * 4 * C2_CHUNK;   // 10240 B per site in HBM
constexpr int C2_GROUP_BYTES = 3 * C2_SITE_BYTES;      // 30720 B: three sites stored as the exact smem image of TB ([plane 16][site 3][40 rows][16 B])
constexpr int C2_PLANE = 3 * C2_CHUNK;         // three sites stacked: 1920 B
constexpr int C2_SMEM = 2 * 2 * 4 * C2_PLANE + 512;
constexpr int FC_KG = 27 * 8;                  // 216 k-groups of fc1 (27 positions x 64 channels)
constexpr int C3_BLOCK_BYTES = 128 * 128;               // one (position, part) of a 128-site tile: [site][64 channels] fp16, 16 KB
constexpr int C3_TILE_BYTES = 27 * 2 * C3_BLOCK_BYTES;  // 884736 B per 128-site tile in HBM: [position][part][site][64 channels]

// What one MMA costs (measured, tools/umma_rate.cu): M = 128, K = 16, operands in shared memory, any N <= 128:
// 32 + N/4 cycles, i.e. (A bytes + B bytes) / 128 B per clock -- the tensor pipe is never the limit here, the operand
// reads are.  So the layer programs minimise the number of A reads: the hi/lo weight halves sit side by side in one
// wider B tile ([w_hi | w_lo] rows), and conv1 merges its three branches into one tile per 5x5 tap.
//
// conv1: per 5x5 tap (kh, kw) one MMA  (a_hi | a_lo) x B  with B's LBO = 0 (both K groups read the same weight
// group), giving (a_hi + a_lo) w_hi and (a_hi + a_lo) w_lo in adjacent accumulator columns.  Accumulator columns
//   [0,16) 1x5 hi  [16,32) 1x5 lo  [32,48) 5x5 hi  [48,64) 5x5 lo  [64,80) 5x1 hi  [80,96) 5x1 lo
// and the 1x5 / 5x1 branches only have taps on the cross kh == 2 / kw == 2:
//   centre  N = 96 at column 0     row taps (kh == 2)  N = 64 at column 0
//   column taps (kw == 2)  N = 64 at column 32          the other 16 taps  N = 32 at column 32
__host__ __device__ constexpr int c1_tap_n(int t) { return t == 12 ? 96 : (t / 5 == 2 || t % 5 == 2) ? 64 : 32; }
__host__ __device__ constexpr int c1_tap_dcol(int t) { return t / 5 == 2 ? 0 : 32; }
__host__ __device__ constexpr int c1_tap_off16(int t) { int o = 0; for (int i = 0; i < t; i++) o += c1_tap_n(i); return o; }   // 16-byte units
constexpr int W1_BYTES = c1_tap_off16(25) * 16;            // 17920: one 16-byte row (8 input-channel slots) per accumulator column
constexpr int W2_TILE = 2 * 64 * 16;                        // conv2 per (tap, 16-channel chunk): [k-group 2][w_hi 32 | w_lo 32 rows][16 B]
constexpr int W2_BYTES = 18 * W2_TILE;                      // 36864
constexpr int W3_TILE = 2 * 128 * 16;                       // conv3 per (tap, chunk): [k-group 2][w_hi 64 | w_lo 64 rows][16 B]
constexpr int W3_BYTES = 12 * W3_TILE;                      // 49152
constexpr int WF_TILE = 2 * 96 * 16;                        // fc1 per (position, chunk): [k-group 2][w_hi 48 | w_lo 48 rows][16 B]
constexpr int WF_POS_BYTES = 4 * WF_TILE;                   // 12288
}  // namespace tcg

// ---- compile-time MMA programs -----------------------------------------------------------------
// The operand offsets depend only on the layer geometry, so every descriptor is (runtime smem base) +
// (compile-time constant): one MMA costs a handful of integer instructions in the one issuing thread.
// All offsets below are in 16-byte units (the descriptor granularity).  B tile order = tc_model_prepare.
__device__ __forceinline__ uint64_t sdesc16(uint32_t lo) { return (0x4008ull << 32) | (uint64_t)lo; }   // SBO 128 B, version 1

template <int T>
__device__ __forceinline__ void issue_conv1_tap(uint32_t a16, uint32_t w16, uint32_t d, uint32_t acc) {
    constexpr int shift = (T / 5) * tcg::WP + (T % 5);
    constexpr uint32_t idesc = make_idesc_f16(128, tcg::c1_tap_n(T));
    umma_f16(d + tcg::c1_tap_dcol(T), sdesc16(a16 + shift + ((uint32_t)(tcg::IN_PLANE / 16) << 16)),
             sdesc16(w16 + tcg::c1_tap_off16(T)), idesc, acc);                       // B: LBO = 0
}
template <int... I>
__device__ __forceinline__ void issue_conv1_seq(uint32_t a16, uint32_t w16, uint32_t d, std::integer_sequence<int, I...>) {
    issue_conv1_tap<12>(a16, w16, d, 0u);                                            // the centre tap covers all 96 columns: it initialises
    (issue_conv1_tap<(I < 12 ? I : I + 1)>(a16, w16, d, 1u), ...);
}
// one 128-row tile of conv1 (all three branches): 25 MMAs
__device__ __forceinline__ void issue_conv1_tile(uint32_t a16, uint32_t w16, uint32_t d) {
    issue_conv1_seq(a16, w16, d, std::make_integer_sequence<int, 24>{});
}
// conv2 (tap = kh*3 + kw, three 16-channel chunks per tap): a_hi x [w_hi | w_lo] (N = 64) + a_lo x w_hi (N = 32)
template <int IDX>
__device__ __forceinline__ void issue_conv2_step(uint32_t c116, uint32_t w16, uint32_t d) {
    constexpr int tap = IDX / 3, c = IDX % 3, kh = tap / 3, kw = tap % 3, par = kw & 1, shift = kh * tcg::C1_PITCH + (kw >> 1);
    const uint32_t a_hi = c116 + (par * 6 + 2 * c) * tcg::C1_ROWS + shift + ((uint32_t)tcg::C1_ROWS << 16);
    const uint32_t a_lo = a_hi + 12 * tcg::C1_ROWS;
    const uint32_t b = (w16 + (tcg::W1_BYTES + IDX * tcg::W2_TILE) / 16) | (64u << 16);
    umma_f16(d, sdesc16(a_hi), sdesc16(b), make_idesc_f16(128, 64), IDX > 0 ? 1u : 0u);
    umma_f16(d, sdesc16(a_lo), sdesc16(b), make_idesc_f16(128, 32), 1u);
}
template <int... I>
__device__ __forceinline__ void issue_conv2_seq(uint32_t c116, uint32_t w16, uint32_t d, std::integer_sequence<int, I...>) {
    (issue_conv2_step<I>(c116, w16, d), ...);
}
// conv3 (tap = kh*3 + kw, two chunks per tap): a_hi x [w_hi | w_lo] (N = 128) + a_lo x w_hi (N = 64)
template <int IDX>
__device__ __forceinline__ void issue_conv3_step(uint32_t c216, uint32_t w16, uint32_t d) {
    constexpr int PL = tcg::C2_PLANE / 16;          // 120
    constexpr int tap = IDX / 2, c = IDX % 2, kh = tap / 3, kw = tap % 3, par = kw & 1, shift = kh * tcg::C2_PITCH + (kw >> 1);
    const uint32_t a_hi = c216 + (par * 4 + 2 * c) * PL + shift + ((uint32_t)PL << 16);
    const uint32_t a_lo = a_hi + 8 * PL;
    const uint32_t b = (w16 + IDX * (tcg::W3_TILE / 16)) | (128u << 16);
    umma_f16(d, sdesc16(a_hi), sdesc16(b), make_idesc_f16(128, 128), IDX > 0 ? 1u : 0u);
    umma_f16(d, sdesc16(a_lo), sdesc16(b), make_idesc_f16(128, 64), 1u);
}
template <int... I>
__device__ __forceinline__ void issue_conv3_seq(uint32_t c216, uint32_t w16, uint32_t d, std::integer_sequence<int, I...>) {
    (issue_conv3_step<I>(c216, w16, d), ...);
}

struct TAParams {
    const void* in; int in_mode; int64_t in_site_stride;
    const float* scale_f; const double* scale_d;
    int64_t n_sites;
    const uint8_t* wimg;                      // W1 tiles then W2 tiles
    const float* bias1; const float* bias2;
    uint8_t* c2_out;
    int* err;
};

constexpr int TA_WGS = 3;
constexpr int TA_THREADS = TA_WGS * 128 + 128;                                     // three epilogue warpgroups + one warpgroup of MMA issuer warps (one per epilogue warpgroup)
constexpr int TA_SMEM_W = tcg::W1_BYTES + tcg::W2_BYTES;                           // 54784
constexpr int TA_RAW_BYTES = 2176;                                                 // staging of one int16 site image (2064 B)
constexpr int TA_SMEM_WG = 2 * tcg::IN_PLANE + tcg::C1_BYTES + TA_RAW_BYTES;       // 13312 + 41344 + 2176
constexpr int TA_SMEM_MISC = 160 * 4 + 64;                                         // bias[80], bias * log2e [80], mbarriers, TMEM slot
constexpr int TA_SMEM = TA_SMEM_W + TA_WGS * TA_SMEM_WG + TA_SMEM_MISC + 64;
constexpr int TA_TMEM_WG = 160;                                                    // conv1 96 (both tiles, one after the other) + conv2 64
static_assert(TA_SMEM <= 232448, "TA shared memory exceeds the 227 KB per-CTA limit");

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// fp32 debug / drop-in input path: one pixel (5 channels) straight from global memory
__device__ __forceinline__ void ta_load_pixel_f32(const TAParams& P, int64_t site, int px, float* v) {
    const float* src = reinterpret_cast<const float*>(P.in) + site * P.in_site_stride + px * 5;
#pragma unroll
    for (int c = 0; c < 5; c++) v[c] = __ldg(src + c);
}
// 32 accumulator columns [w_hi part 16 | w_lo part 16] -> 16 sums
__device__ __forceinline__ void tmem_ld_pair16(uint32_t addr, float* v) {
    float a[16], b[16];
    tmem_ld16_nowait(addr, a);
    tmem_ld16_nowait(addr + 16, b);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = a[i] + b[i];
}

// TA: conv1_{1,2,3} + conv2.  148 persistent CTAs x 3 epilogue warpgroups, one site per warpgroup at a time, plus one
// MMA issuer warp per warpgroup (tcgen05.mma blocks its issuing thread while the MMA queue is full, so a warp that also
// has epilogue work must not issue).  Per site a warpgroup runs three MMA phases on its own accumulator columns (conv1
// rows 0..127, conv1 rows 97..224, conv2), handed to the issuer through a 160-thread named barrier and tracked by one
// mbarrier; the CUDA-core work is arranged so that most of it runs while this warpgroup's own MMAs are in flight (and
// the other two warpgroups keep the operand pipe busy the rest of the time):
//   conv1 tile 0 in flight : previous site's conv2 epilogue (bias, SELU, fp16 split, HBM stores)
//   conv1 tile 1 in flight : tile 0 epilogue (bias, SELU, split -> c1 planes in shared memory)
//   conv2 in flight        : next site's int16 tensor -> scaled fp16 hi/lo input planes
template <bool SAT>
__global__ void __launch_bounds__(TA_THREADS, 1) tc_trunk_a_kernel(const TAParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* s_w = smem;
    uint8_t* s_wg0 = smem + TA_SMEM_W;
    float* s_bias = reinterpret_cast<float*>(smem + TA_SMEM_W + TA_WGS * TA_SMEM_WG);
    float* s_bl = s_bias + 80;                                                      // bias * log2(e)
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_bl + 80);                       // [wg]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 4);

    const int tid = threadIdx.x, wg = tid >> 7, t = tid & 127, warp = tid >> 5, wq = warp & 3;
    uint8_t* s_in = s_wg0 + wg * TA_SMEM_WG;               // hi plane, then lo plane
    uint8_t* s_c1 = s_in + 2 * tcg::IN_PLANE;
    uint8_t* s_raw = s_c1 + tcg::C1_BYTES;

    for (int i = tid; i < TA_SMEM_W / 16; i += TA_THREADS) reinterpret_cast<uint4*>(s_w)[i] = __ldg(reinterpret_cast<const uint4*>(P.wimg) + i);
    if (tid < 48) { const float b = P.bias1[tid]; s_bias[tid] = b; s_bl[tid] = b * 1.4426950408889634f; }
    if (tid >= 64 && tid < 96) { const float b = P.bias2[tid - 64]; s_bias[48 + tid - 64] = b; s_bl[48 + tid - 64] = b * 1.4426950408889634f; }
    for (int i = tid; i < TA_WGS * TA_SMEM_WG / 16; i += TA_THREADS) reinterpret_cast<uint4*>(s_wg0)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        for (int i = 0; i < TA_WGS; i++) mbar_init(&s_bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(s_tmem, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const int64_t stride = (int64_t)gridDim.x * TA_WGS;
    if (wg == TA_WGS) {
        // ===== MMA issuer warps: warp wq serves epilogue warpgroup wq
        if (wq < TA_WGS) {
            const int ewg = wq;
            const uint32_t tmem = *s_tmem + (uint32_t)ewg * TA_TMEM_WG;
            const uint32_t in16 = smem_u32(s_wg0 + ewg * TA_SMEM_WG) >> 4, c116 = in16 + 2 * tcg::IN_PLANE / 16, w16 = smem_u32(s_w) >> 4;
            uint64_t* bar = &s_bar[ewg];
            for (int64_t site = (int64_t)blockIdx.x * TA_WGS + ewg; site < P.n_sites; site += stride) {
                wg_issuer_barrier(ewg);                                   // input planes of `site` ready, conv1 columns free
                if (elect_one()) { tc_fence_after(); issue_conv1_tile(in16, w16, tmem); umma_commit(bar); }
                __syncwarp();
                wg_issuer_barrier(ewg);                                   // tile 0 accumulators read
                if (elect_one()) { tc_fence_after(); issue_conv1_tile(in16 + tcg::TILE1_START, w16, tmem); umma_commit(bar); }
                __syncwarp();
                wg_issuer_barrier(ewg);                                   // c1 complete
                if (elect_one()) { tc_fence_after(); issue_conv2_seq(c116, w16, tmem + 96, std::make_integer_sequence<int, 18>{}); umma_commit(bar); }
                __syncwarp();
            }
        }
        tc_fence_before();
        __syncthreads();
        return;
    }
    const uint32_t tmem = *s_tmem + (uint32_t)wg * TA_TMEM_WG;
    const uint32_t tmem_lane = tmem + ((uint32_t)wq << 21);                    // lane quarter of this warp
    uint64_t* bar = &s_bar[wg];
    uint32_t phase = 0;
    bool ok = true;
    uint32_t sat = 0;

    int64_t site = (int64_t)blockIdx.x * TA_WGS + wg;
    const bool raw_mode = P.in_mode != 0;
    uint4 pre0 = make_uint4(0, 0, 0, 0), pre1 = make_uint4(0, 0, 0, 0);
    float pre_sf = 1.f; double pre_sd = 1.0;
    auto prefetch = [&](int64_t sidx) {          // 2064 B of int16 -> registers, 16 B per thread (+1 chunk on thread 0), and the site's scale
        const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(P.in) + sidx * P.in_site_stride * 2);
        pre0 = __ldg(src + t);
        if (t == 0) pre1 = __ldg(src + 128);
        if (P.in_mode == 1) pre_sf = __ldg(P.scale_f + sidx);
        if (P.in_mode == 2) pre_sd = __ldg(P.scale_d + sidx);
    };
    // site image (registers) -> padded fp16 hi / lo planes (pixel row = (h+2)*45 + (w+2)); channels 5..7 stay zero
    auto convert = [&](int64_t sidx) {
        const float sc_f = pre_sf; const double sc_d = pre_sd;
        if (raw_mode) {
            *reinterpret_cast<uint4*>(s_raw + t * 16) = pre0;
            if (t == 0) *reinterpret_cast<uint4*>(s_raw + 2048) = pre1;
            wg_barrier(wg);
        }
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int px = t + 128 * k;
            if (px < 205) {
                const int h = px / 41, w = px - h * 41;
                float v[5];
                if (!raw_mode) {
                    ta_load_pixel_f32(P, sidx, px, v);
                } else {
                    const int16_t* rp = reinterpret_cast<const int16_t*>(s_raw) + px * 5;
#pragma unroll
                    for (int c = 0; c < 5; c++) {
                        const int16_t raw = rp[c];
                        float x = __int_as_float(0x4B400000 + (int)raw) - 12582912.f;      // exact int16 -> fp32 without the conversion pipe
                        if (h > 0 && c < 4) x = P.in_mode == 1 ? __fmul_rn(x, sc_f) : (float)((double)raw * sc_d);
                        v[c] = x;
                    }
                }
                uint4 hi = make_uint4(0, 0, 0, 0), lo = make_uint4(0, 0, 0, 0);
                split2(v[0], v[1], hi.x, lo.x);
                split2(v[2], v[3], hi.y, lo.y);
                split2(v[4], 0.f, hi.z, lo.z);
                if (SAT) sat += sat_halves(hi);
                hi.z |= 0x3C000000u;                                       // channel slot 5 = 1.0 on real pixels: multiplies the bias row of the centre tap
                const int row = (h + 2) * tcg::WP + w + 2;
                *reinterpret_cast<uint4*>(s_in + row * 16) = hi;
                *reinterpret_cast<uint4*>(s_in + tcg::IN_PLANE + row * 16) = lo;
            }
        }
        if (raw_mode) wg_barrier(wg);            // s_raw is free for the next image
    };
    // conv1 epilogue of one tile: bias + SELU + split -> c1 planes [part][parity][k-group][h*21 + w/2]
    auto store_c1 = [&](int j, const float* v) {
        const int m = (j ? tcg::TILE1_START : 0) + t;
        const int h = m / tcg::WP, w = m - h * tcg::WP;
        const bool valid = m < 225 && w < 41 && (j == 0 || m >= 128);      // tile 1: t >= 32, warp 0 skips the arithmetic altogether
        if (valid) {
            uint8_t* dst = s_c1 + ((w & 1) * 6) * tcg::C1_PLANE + (h * tcg::C1_PITCH + (w >> 1)) * 16;
#pragma unroll
            for (int kg = 0; kg < 6; kg++) {
                uint4 hi, lo;
                act_split8_nobias(v + 8 * kg, hi, lo);                     // conv1 bias rides on the centre tap (constant channel 5)
                if (SAT) sat += sat_halves(hi);
                *reinterpret_cast<uint4*>(dst + kg * tcg::C1_PLANE) = hi;
                *reinterpret_cast<uint4*>(dst + (12 + kg) * tcg::C1_PLANE) = lo;
            }
        }
    };
    // accumulator columns -> 48 channel values in c1 order [1x5 | 5x1 | 5x5]
    auto load_c1 = [&](float* v) {
        tmem_ld_pair16(tmem_lane + 0, v);
        tmem_ld_pair16(tmem_lane + 64, v + 16);
        tmem_ld_pair16(tmem_lane + 32, v + 32);
    };
    // conv2 epilogue: bias + SELU + split -> HBM c2 [group][part][parity][k-group][site % 3][h2*10 + w2/2][8]
    auto store_c2 = [&](int64_t sidx, const float* acc) {
        const int m = t, h2 = m / tcg::C1_PITCH, w2 = m - h2 * tcg::C1_PITCH;
        if (m < 84 && w2 < 20) {
            const uint32_t grp = (uint32_t)sidx / 3u, sub = (uint32_t)sidx - 3u * grp;       // site counts stay below 2^31
            uint8_t* dst = P.c2_out + (int64_t)grp * tcg::C2_GROUP_BYTES + sub * tcg::C2_CHUNK + ((w2 & 1) * 4) * tcg::C2_PLANE + (h2 * tcg::C2_PITCH + (w2 >> 1)) * 16;
#pragma unroll
            for (int kg = 0; kg < 4; kg++) {
                uint4 hi, lo;
                act_split8(acc + 8 * kg, s_bias + 48 + 8 * kg, s_bl + 48 + 8 * kg, hi, lo);
                if (SAT) sat += sat_halves(hi);
                *reinterpret_cast<uint4*>(dst + kg * tcg::C2_PLANE) = hi;
                *reinterpret_cast<uint4*>(dst + (8 + kg) * tcg::C2_PLANE) = lo;
            }
        }
    };

    if (site < P.n_sites) {
        if (raw_mode) prefetch(site);
        convert(site);
        if (raw_mode && site + stride < P.n_sites) prefetch(site + stride);
    }

#ifdef NC_TA_PROFILE      // development aid: cycles per phase of the site loop, printed by (block 0, warpgroup 0, threads 0 and 64)
    uint32_t prof[13], prof_last = (uint32_t)clock(), prof_n = 0;
    for (int i = 0; i < 13; i++) prof[i] = 0;
#define NC_TA_MARK(i) { const uint32_t now_ = (uint32_t)clock(); prof[i] += now_ - prof_last; prof_last = now_; if (i == 12) prof_n++; }
#else
#define NC_TA_MARK(i)
#endif
    float acc2[32];
    int64_t prev = -1;
    for (; site < P.n_sites; site += stride) {
        NC_TA_MARK(0);
        // ---- conv1, rows 0..127: this site's input planes are written, the previous site's accumulators are in registers
        fence_async_smem();
        tc_fence_before();
        wg_issuer_barrier(wg);
        NC_TA_MARK(1);
        if (prev >= 0) store_c2(prev, acc2);                          // overlaps the MMAs just issued
        NC_TA_MARK(2);
        ok = mbar_wait(bar, phase) && ok; phase ^= 1;
        tc_fence_after();
        NC_TA_MARK(3);
        float v[48];
        load_c1(v);
        tc_fence_before();
        wg_issuer_barrier(wg);                                        // every warp has read its accumulator rows: conv1 rows 97..224 may start
        NC_TA_MARK(4);
        NC_TA_MARK(5);
        store_c1(0, v);                                               // overlaps tile 1
        NC_TA_MARK(6);
        ok = mbar_wait(bar, phase) && ok; phase ^= 1;
        tc_fence_after();
        NC_TA_MARK(7);
        load_c1(v);
        store_c1(1, v);
        fence_async_smem();
        tc_fence_before();
        wg_issuer_barrier(wg);                                        // c1 complete: conv2 may start; input planes and conv1 columns are free
        NC_TA_MARK(8);
        NC_TA_MARK(9);
        const int64_t next = site + stride;
        if (next < P.n_sites) {                                       // overlaps conv2
            convert(next);
            if (raw_mode && next + stride < P.n_sites) prefetch(next + stride);
        }
        NC_TA_MARK(10);
        ok = mbar_wait(bar, phase) && ok; phase ^= 1;
        tc_fence_after();
        NC_TA_MARK(11);
        {
            // columns 96..127 = (a_hi + a_lo) w_hi, 128..159 = a_hi w_lo
            float lo_part[32];
            tmem_ld16_nowait(tmem_lane + 96, acc2);
            tmem_ld16_nowait(tmem_lane + 112, acc2 + 16);
            tmem_ld16_nowait(tmem_lane + 128, lo_part);
            tmem_ld16_nowait(tmem_lane + 144, lo_part + 16);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i++) acc2[i] += lo_part[i];
        }
        prev = site;
        NC_TA_MARK(12);
    }
    if (prev >= 0) store_c2(prev, acc2);
#ifdef NC_TA_PROFILE
    if (blockIdx.x == 0 && wg == 0 && (t == 0 || t == 64) && prof_n > 0) {
        printf("TA profile t=%d sites=%u cycles/site:", t, prof_n);
        for (int i = 0; i < 13; i++) printf(" [%d]%u", i, prof[i] / prof_n);
        printf("\n");
    }
#endif
#undef NC_TA_MARK
    if (!ok && t == 0) atomicExch(P.err, 1);
    if (SAT && sat) atomicAdd(P.err + 1, (int)sat);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(*s_tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// TB — conv3, three sites per tile
// ------------------------------------------------------------------------------------------------
struct TBParams {
    const uint8_t* c2; int64_t n_sites;
    const uint8_t* wimg; const float* bias;
    uint8_t* c3_out; int* err;
};
// Two consumer warpgroups, not three: a stage stays occupied from the arrival of its image until its MMAs have completed, so with
// a ring of five (all that fits beside the weights) three consumers leave two copies in flight per SM and two consumers three —
// the kernel is bound by bytes in flight, not by the tensor pipe (2.7 -> 2.4 ms for configs[1]).
#ifndef NC_TB_WGS
#define NC_TB_WGS 2
#endif
constexpr int TB_WGS = NC_TB_WGS;
constexpr int TB_THREADS = TB_WGS * 128 + 32;               // consumer warpgroups + one TMA producer warp
#ifndef NC_TB_RING
#define NC_TB_RING 5
#endif
constexpr int TB_RING = NC_TB_RING;                         // c2 group images in flight or in use (30,720 B each)
constexpr int TB_SMEM_RING = TB_RING * tcg::C2_GROUP_BYTES + 512;   // + slack: junk GEMM rows read past the last plane
constexpr int TB_SMEM_MISC = 128 * 4 + 128 + 16;            // bias, bias * log2e; mbarriers; TMEM slot
constexpr int TB_SMEM = tcg::W3_BYTES + TB_SMEM_RING + TB_SMEM_MISC + 64;
static_assert(TB_SMEM <= 232448, "TB shared memory exceeds the 227 KB per-CTA limit");

// TB: conv3.  A group = three consecutive sites = one 128-row tile (40 rows per site) whose c2 image TA left in HBM in
// exactly the shared-memory plane layout, so it arrives by ONE bulk copy.  The kernel is bound by the latency of that
// copy (measured: ~8,500 cycles under load against ~1,900 of MMAs and ~2,700 of epilogue per group), so a producer warp
// keeps a ring of TB_RING images in flight for the three consumer warpgroups (full / empty mbarriers per stage; the
// stage is released by a tcgen05.commit, i.e. when the MMAs that read it have completed).
// The CTA's i-th group is global group (i / 3) * 3 * gridDim + 3 * blockIdx + i % 3; warpgroup w consumes i = w mod 3.
template <bool SAT>
__global__ void __launch_bounds__(TB_THREADS, 1) tc_trunk_b_kernel(const TBParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* s_w = smem;
    uint8_t* s_ring = smem + tcg::W3_BYTES;
    float* s_bias = reinterpret_cast<float*>(s_ring + TB_SMEM_RING);
    float* s_bl = s_bias + 64;
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_bl + 64);      // [wg] MMAs of the warpgroup's current group done
    uint64_t* s_full = s_bar + TB_WGS;                              // [stage] image landed
    uint64_t* s_empty = s_full + TB_RING;                           // [stage] MMAs that read the image done
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_empty + TB_RING);
    const int tid = threadIdx.x, wg = tid >> 7, t = tid & 127, warp = tid >> 5, wq = warp & 3;

    for (int i = tid; i < tcg::W3_BYTES / 16; i += TB_THREADS) reinterpret_cast<uint4*>(s_w)[i] = __ldg(reinterpret_cast<const uint4*>(P.wimg) + i);
    if (tid < 64) { const float b = P.bias[tid]; s_bias[tid] = b; s_bl[tid] = b * 1.4426950408889634f; }
    for (int i = tid; i < TB_SMEM_RING / 16; i += TB_THREADS) reinterpret_cast<uint4*>(s_ring)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        for (int i = 0; i < TB_WGS; i++) mbar_init(&s_bar[i], 1);
        for (int i = 0; i < TB_RING; i++) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(s_tmem, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const int64_t n_groups = (P.n_sites + 2) / 3;
    const int64_t gstride = (int64_t)gridDim.x * TB_WGS;
    const int64_t gbase = (int64_t)blockIdx.x * TB_WGS;
    bool ok = true;

    if (wg == TB_WGS) {
        // ===== producer warp: one bulk copy per group, in the CTA's consumption order
        if (elect_one()) {
            for (int64_t i = 0;; i++) {
                const int64_t grp = (i / TB_WGS) * gstride + gbase + (i % TB_WGS);
                if (grp >= n_groups) break;
                const int st = (int)(i % TB_RING);
                if (i >= TB_RING) ok = mbar_wait(&s_empty[st], (uint32_t)((i / TB_RING) - 1) & 1u) && ok;
                mbar_expect_tx(&s_full[st], tcg::C2_GROUP_BYTES);
                bulk_g2s(s_ring + st * tcg::C2_GROUP_BYTES, P.c2 + grp * tcg::C2_GROUP_BYTES, tcg::C2_GROUP_BYTES, &s_full[st]);
            }
            if (!ok) atomicExch(P.err, 1);
        }
        __syncwarp();
        tc_fence_before();
        __syncthreads();
        return;
    }
    const uint32_t tmem = *s_tmem + (uint32_t)wg * 128u;          // columns [0,64) = (a_hi + a_lo) w_hi, [64,128) = a_hi w_lo
    const uint32_t tmem_lane = tmem + ((uint32_t)wq << 21);
    const uint32_t w16 = smem_u32(s_w) >> 4;
    uint32_t phase = 0, sat = 0;
    for (int64_t k = 0;; k++) {
        const int64_t i = k * TB_WGS + wg;
        const int64_t grp = k * gstride + gbase + wg;
        if (grp >= n_groups) break;
        const int64_t s0 = grp * 3;
        const int st = (int)(i % TB_RING);
        const uint32_t c216 = smem_u32(s_ring + st * tcg::C2_GROUP_BYTES) >> 4;
        // ---- the group's c2 image (30,720 B, already in plane layout).  A parity wait is only safe one phase ahead, and the
        // stage's previous user is ANOTHER warpgroup: first make sure that use is over (the condition the producer waited for
        // before refilling), then the stage's full barrier is in this group's phase and the parity is unambiguous.
        if (i >= TB_RING) ok = mbar_wait(&s_empty[st], (uint32_t)((i / TB_RING) - 1) & 1u) && ok;
        ok = mbar_wait(&s_full[st], (uint32_t)(i / TB_RING) & 1u) && ok;
        tc_fence_before();
        wg_barrier(wg);                                              // the previous group's accumulators have been read by every warp
        if (wq == 0 && elect_one()) {
            tc_fence_after();
            issue_conv3_seq(c216, w16, tmem, std::make_integer_sequence<int, 12>{});
            umma_commit(&s_empty[st]);                               // the producer may refill the stage
            umma_commit(&s_bar[wg]);
        }
        __syncwarp();
        ok = mbar_wait(&s_bar[wg], phase) && ok; phase ^= 1;
        tc_fence_after();
        // ---- epilogue: row m = s*40 + h3*10 + w3 -> HBM c3 [tile of 128 sites][pos 27][part 2][site][64 channels]: every thread
        //      writes two full 128-byte lines (scattering 16-byte pieces into a k-group-major layout cost TB a third of its
        //      time); the 16-byte chunks of a line are permuted (chunk ^ site % 8) so that fc1's bulk copy of the 16 KB block
        //      IS the 128-byte-swizzle shared-memory image
        {
            const int s = t / 40, r = t - s * 40, h3 = r / 10, w3 = r - h3 * 10;
            const int64_t site = s0 + s;
            const bool valid = s < 3 && h3 < 3 && w3 < 9 && site < P.n_sites;
            const int pos = h3 * 9 + w3;
            uint8_t* dst = P.c3_out + (site >> 7) * (int64_t)tcg::C3_TILE_BYTES + (int64_t)pos * (2 * tcg::C3_BLOCK_BYTES) + (site & 127) * 128;
#pragma unroll 1
            for (int half = 0; half < 2; half++) {
                float acc[32], acc2[32];
                tmem_ld16_nowait(tmem_lane + half * 32, acc);
                tmem_ld16_nowait(tmem_lane + half * 32 + 16, acc + 16);
                tmem_ld16_nowait(tmem_lane + 64 + half * 32, acc2);
                tmem_ld16_nowait(tmem_lane + 64 + half * 32 + 16, acc2 + 16);
                tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int i2 = 0; i2 < 32; i2++) acc[i2] += acc2[i2];
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        uint4 hi, lo;
                        const int kg = half * 4 + g;
                        act_split8(acc + 8 * g, s_bias + 8 * kg, s_bl + 8 * kg, hi, lo);
                        if (SAT) sat += sat_halves(hi);
                        const int ck = (kg ^ (int)(site & 7)) * 16;                  // chunk position under the 128-byte swizzle
                        *reinterpret_cast<uint4*>(dst + ck) = hi;
                        *reinterpret_cast<uint4*>(dst + tcg::C3_BLOCK_BYTES + ck) = lo;
                    }
                }
            }
        }
        tc_fence_before();
    }
    if (!ok && t == 0) atomicExch(P.err, 1);
    if (SAT && sat) atomicAdd(P.err + 1, (int)sat);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(*s_tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// TC — fc1 (K = 1728 streamed by position, two smem stages) + heads, 128 sites per CTA
// ------------------------------------------------------------------------------------------------
struct TCParams {
    const uint8_t* c3; int64_t n_sites;
    const uint8_t* wimg;            // [27 positions][8 tiles of 1536 B]
    const float* bias;              // fc1 bias (48)
    TailW tail; int haploid;
    const NcSiteMeta* meta; const float* ref4;
    float* out10; float* probs4;
    int* err;
};
constexpr int TC_STAGE_A = 2 * 8 * 2048;                    // 32768: [part][kg 8][128 rows][16 B]
constexpr int TC_STAGE = TC_STAGE_A + tcg::WF_POS_BYTES;    // + 12288 of weights
constexpr int TC_SMEM = 2 * TC_STAGE + 48 * 4 + 96 + 64 + 1024;     // + slack for the 1024-byte alignment of the stages
static_assert(TC_STAGE % 1024 == 0, "stages must keep the 1024-byte alignment");

__device__ __forceinline__ void snp_tail_row(const float* x, int64_t s, const TCParams& P) {
    const TailW& w = P.tail;
    float ref[4];
    if (P.ref4) { for (int j = 0; j < 4; j++) ref[j] = P.ref4[s * 4 + j]; }
    else { const int rc = P.meta[s].ref_code; for (int j = 0; j < 4; j++) ref[j] = (j == rc) ? 1.f : 0.f; }
    if (P.haploid) {
        float in[20], z[4];
        dense_t<48, 16>(x, w.fc2_k, w.fc2_b, in, true);
        for (int j = 0; j < 4; j++) in[16 + j] = ref[j];
        dense_t<20, 4>(in, w.fc3_k, w.fc3_b, z, true);
        softmax_t<4>(z);
        for (int j = 0; j < 4; j++) P.probs4[s * 4 + j] = z[j];
        return;
    }
    float fa[17], fc3in[24];
    dense_t<48, 16>(x, w.fa_k, w.fa_b, fa, true);
    if (P.out10) dense_t<48, 16>(x, w.fc2_k, w.fc2_b, fc3in, true);
    for (int j = 0; j < 4; j++) {
        float z[2];
        fa[16] = ref[j];
        dense_t<17, 2>(fa, w.hk[j], w.hb[j], z, false);
        softmax_t<2>(z);
        fc3in[16 + 2 * j] = z[0]; fc3in[17 + 2 * j] = z[1];
        if (P.out10) { P.out10[s * 10 + 2 * j] = z[0]; P.out10[s * 10 + 2 * j + 1] = z[1]; }
        if (P.probs4) P.probs4[s * 4 + j] = z[1];
    }
    if (P.out10) {
        float fc3[8], gt[2];
        dense_t<24, 8>(fc3in, w.fc3_k, w.fc3_b, fc3, true);
        dense_t<8, 2>(fc3, w.gt_k, w.gt_b, gt, false);
        softmax_t<2>(gt);
        P.out10[s * 10 + 8] = gt[0]; P.out10[s * 10 + 9] = gt[1];
    }
}

// K-major operand in the 128-byte-swizzle layout: rows of 128 B (64 fp16 along K), 16-byte chunk j of row r stored at chunk
// j ^ (r & 7); 8-row groups 1024 B apart (SBO); the K = 16 slice of an MMA is selected by advancing the start address by 32 B.
// The tile base must be 1024-byte aligned.
__device__ __forceinline__ uint64_t sdesc_sw128(uint32_t addr16) {
    return (uint64_t)(addr16 & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// TC: fc1 as a K-streamed GEMM over the 27 positions (K = 64 channels each) + the heads, 128 sites per CTA, 2 CTAs per SM.
// TB leaves c3 in HBM as [position][part][site][64 channels] blocks of 16 KB whose 16-byte chunks are already permuted the
// way the 128-byte swizzle wants them, so a block is ONE bulk copy and lands as a ready UMMA operand.  Warp 0 = TMA
// producer, warp 1 = MMA issuer, two stages (full / empty mbarriers), all four warps run the epilogue.
__global__ void __launch_bounds__(128, 2) tc_fc_kernel(const TCParams P) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // swizzled tiles want 1024-byte alignment
    float* s_bias = reinterpret_cast<float*>(smem + 2 * TC_STAGE);
    uint64_t* s_full = reinterpret_cast<uint64_t*>(s_bias + 48);      // [2] TMA bytes landed
    uint64_t* s_empty = s_full + 2;                                    // [2] MMAs that read the stage have completed
    uint64_t* s_done = s_empty + 2;                                    // every MMA of the tile has completed
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_done + 1);
    const int t = threadIdx.x, warp = t >> 5;
    const int64_t tile = blockIdx.x;

    if (t < 48) s_bias[t] = P.bias[t];
    if (t == 0) {
        mbar_init(&s_full[0], 1); mbar_init(&s_full[1], 1); mbar_init(&s_empty[0], 1); mbar_init(&s_empty[1], 1); mbar_init(s_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(s_tmem, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;                                     // columns [0,48) = (a_hi + a_lo) w_hi, [48,96) = a_hi w_lo
    const uint32_t tmem_lane = tmem + ((uint32_t)(warp & 3) << 21);
    const uint8_t* a_src = P.c3 + tile * (int64_t)tcg::C3_TILE_BYTES;
    bool ok = true;

    if (warp == 0) {
        // ===== TMA producer: per position two 16 KB slabs of activations (hi, lo) and 12 KB of weights
        if (elect_one()) {
            for (int pos = 0; pos < 27; pos++) {
                const int st = pos & 1;
                if (pos >= 2) ok = mbar_wait(&s_empty[st], ((pos >> 1) - 1) & 1) && ok;
                uint8_t* dst = smem + st * TC_STAGE;
                mbar_expect_tx(&s_full[st], TC_STAGE);
                bulk_g2s(dst, a_src + (int64_t)(pos * 2) * tcg::C3_BLOCK_BYTES, 2 * tcg::C3_BLOCK_BYTES, &s_full[st]);   // hi block, lo block
                bulk_g2s(dst + TC_STAGE_A, P.wimg + (int64_t)pos * tcg::WF_POS_BYTES, tcg::WF_POS_BYTES, &s_full[st]);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer
        if (elect_one()) {
            for (int pos = 0; pos < 27; pos++) {
                const int st = pos & 1;
                ok = mbar_wait(&s_full[st], (pos >> 1) & 1) && ok;
                tc_fence_after();
                const uint32_t sb16 = smem_u32(smem + st * TC_STAGE) >> 4;
                const uint32_t b0 = (sb16 + TC_STAGE_A / 16) | (96u << 16);
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const uint32_t a_hi = sb16 + 2 * c, a_lo = a_hi + tcg::C3_BLOCK_BYTES / 16;       // 32 B further along K per chunk
                    umma_f16(tmem, sdesc_sw128(a_hi), sdesc16(b0 + c * (tcg::WF_TILE / 16)), make_idesc_f16(128, 96), (pos > 0 || c > 0) ? 1u : 0u);
                    umma_f16(tmem, sdesc_sw128(a_lo), sdesc16(b0 + c * (tcg::WF_TILE / 16)), make_idesc_f16(128, 48), 1u);
                }
                umma_commit(&s_empty[st]);
            }
            umma_commit(s_done);
        }
        __syncwarp();
    }
    // ===== epilogue (all four warps).  A dedicated barrier: warps 2 and 3 get here before the first MMA, and a parity-1
    // wait on a fresh mbarrier passes immediately.
    ok = mbar_wait(s_done, 0) && ok;
    tc_fence_after();
    {
        float x[48], xl[48];
        tmem_ld16_nowait(tmem_lane, x);
        tmem_ld16_nowait(tmem_lane + 16, x + 16);
        tmem_ld16_nowait(tmem_lane + 32, x + 32);
        tmem_ld16_nowait(tmem_lane + 48, xl);
        tmem_ld16_nowait(tmem_lane + 64, xl + 16);
        tmem_ld16_nowait(tmem_lane + 80, xl + 32);
        tmem_ld_wait();
        const int64_t s = tile * 128 + t;
        if (s < P.n_sites) {
#pragma unroll
            for (int i = 0; i < 48; i++) x[i] = selu_bf(x[i] + xl[i] + s_bias[i]);
            snp_tail_row(x, s, P);
        }
    }
    if (!ok) atomicExch(P.err, 1);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

// ------------------------------------------------------------------------------------------------
// Generic single-tile UMMA probe (tests/test_cuda_umma.py): runs a program on caller-supplied A
// planes and B tiles and returns the fp32 accumulator block — pins the descriptor semantics the
// layer programs rely on (row shifts through the start address, K groups through LBO).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) umma_probe_kernel(const uint8_t* a_img, int a_bytes, const uint8_t* b_img, int b_bytes,
                                                            const MmaOp* prog, int n_ops, int N, int ncols, float* out, int* err) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int t = threadIdx.x, warp = t >> 5;
    for (int i = t; i < a_bytes / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(a_img)[i];
    for (int i = t; i < b_bytes / 16; i += 128) reinterpret_cast<uint4*>(smem + a_bytes)[i] = reinterpret_cast<const uint4*>(b_img)[i];
    if (t == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc(&tmem_slot, 128);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (t == 0) {
        issue_program(prog, n_ops, smem_u32(smem), smem_u32(smem + a_bytes), (uint32_t)N * 16u, tmem, make_idesc_f16(128, N));
        umma_commit(&bar);
    }
    const bool ok = mbar_wait(&bar, 0);
    tc_fence_after();
    for (int cb = 0; cb < ncols / 16; cb++) {
        float acc[16];
        tmem_ld16(tmem + ((uint32_t)(warp & 3) << 21) + cb * 16, acc);
        for (int i = 0; i < 16; i++) out[t * ncols + cb * 16 + i] = acc[i];
    }
    if (!ok && t == 0) atomicExch(err, 1);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

// ------------------------------------------------------------------------------------------------
// host side: weight images, programs, launch
// ------------------------------------------------------------------------------------------------
struct TcModel {
    bool ready = false;
    bool audit = false;                          // run the validation instantiations that count fp16 saturation (err[1])
    int kind = 0;
    DevBuf wimg_a, wimg_b, wimg_c, bias;         // bias: b1(48) b2(32) b3(64) bf(48)
    DevBuf c2, c3, err;
    int64_t tail_off[16][2];
};

inline void tc_model_release(TcModel& T) {
    for (DevBuf* b : {&T.wimg_a, &T.wimg_b, &T.wimg_c, &T.bias, &T.c2, &T.c3, &T.err}) b->release();
    T.ready = false;
}

namespace tcdetail {
inline void put_split(std::vector<uint8_t>& hi_img, std::vector<uint8_t>& lo_img, size_t byte_off, float w) {
    const __half h = __float2half_rn(w);
    const __half l = __float2half_rn(w - __half2float(h));
    memcpy(hi_img.data() + byte_off, &h, 2);
    memcpy(lo_img.data() + byte_off, &l, 2);
}
// B tile [N rows][K = 16] in the K-major blocked layout: element (kslot, n) at ((kslot/8)*N + n)*16 + (kslot%8)*2
inline size_t btile_off(int N, int kslot, int n) { return ((size_t)(kslot >> 3) * N + n) * 16 + (size_t)(kslot & 7) * 2; }
}  // namespace tcdetail

// Builds the tensor-core operand images for an SNP model (kind 0 / 1).  Other kinds: not prepared
// (the fp32 kernels serve them), T.ready stays false.
inline int tc_model_prepare(cudaStream_t stream, TcModel& T, int kind, const float* blob, size_t n_floats, std::string* err) {
    using namespace tcg;
    using namespace tcdetail;
    T.ready = false; T.kind = kind;
    if (kind > 1) return NC_OK;
    (void)n_floats;
    // blob offsets (canonical order, see model_init)
    const float* w11 = blob; const float* b11 = w11 + 400;
    const float* w12 = b11 + 16; const float* b12 = w12 + 400;
    const float* w13 = b12 + 16; const float* b13 = w13 + 2000;
    const float* w2 = b13 + 16; const float* b2 = w2 + 9216;
    const float* w3 = b2 + 32; const float* b3 = w3 + 12288;
    const float* wf = b3 + 64; const float* bf = wf + 82944;

    // 8 input-channel slots of one accumulator column: fp16 hi (part 0) or lo (part 1) of w[ci] for ci < n_ci
    auto put_row = [](std::vector<uint8_t>& img, size_t row_byte_off, const float* w, int ci_stride, int n_ci, int part) {
        for (int ci = 0; ci < n_ci; ci++) {
            const float x = w[(size_t)ci * ci_stride];
            const __half h = __float2half_rn(x);
            const __half l = __float2half_rn(x - __half2float(h));
            memcpy(img.data() + row_byte_off + (size_t)ci * 2, part ? &l : &h, 2);
        }
    };
    // ---------------- conv1: one tile per 5x5 tap, rows = accumulator columns (see tcg), one K group (B LBO = 0)
    std::vector<uint8_t> w1_img(W1_BYTES, 0);
    for (int t = 0; t < 25; t++) {
        const int kh = t / 5, kw = t % 5;
        const size_t base = (size_t)c1_tap_off16(t) * 16;
        int row = 0;
        auto put_branch = [&](const float* wtap, const float* bias16) {   // wtap -> [ci 5][co 16]; 16 hi rows then 16 lo rows
            for (int part = 0; part < 2; part++)
                for (int n = 0; n < 16; n++) {
                    put_row(w1_img, base + (size_t)row * 16, wtap + n, 16, 5, part);
                    if (t == 12) put_row(w1_img, base + (size_t)row * 16 + 10, bias16 + n, 0, 1, part);   // slot 5: the bias, met by the constant 1.0 channel
                    row++;
                }
        };
        if (kh == 2) put_branch(w11 + (size_t)kw * 5 * 16, b11);            // 1x5 branch: columns 0..31
        put_branch(w13 + (size_t)t * 5 * 16, b13);                          // 5x5 branch
        if (kw == 2) put_branch(w12 + (size_t)kh * 5 * 16, b12);            // 5x1 branch: columns 64..95 (32..63 of a column-tap tile)
        if (row != c1_tap_n(t)) { if (err) *err = "conv1 tile size mismatch"; return NC_EINVAL; }
    }
    // ---------------- conv2: [kh 2][kw 3][ci 48][co 32]; per (tap, chunk) [k-group 2][w_hi 32 | w_lo 32][8 slots]
    std::vector<uint8_t> w2_img(W2_BYTES, 0);
    for (int tap = 0; tap < 6; tap++)
        for (int c = 0; c < 3; c++)
            for (int kg = 0; kg < 2; kg++)
                for (int part = 0; part < 2; part++)
                    for (int n = 0; n < 32; n++)
                        put_row(w2_img, (size_t)(tap * 3 + c) * W2_TILE + ((size_t)kg * 64 + part * 32 + n) * 16,
                                w2 + ((size_t)tap * 48 + 16 * c + 8 * kg) * 32 + n, 32, 8, part);
    // ---------------- conv3: [kh 2][kw 3][ci 32][co 64]; per (tap, chunk) [k-group 2][w_hi 64 | w_lo 64][8 slots]
    std::vector<uint8_t> w3_img(W3_BYTES, 0);
    for (int tap = 0; tap < 6; tap++)
        for (int c = 0; c < 2; c++)
            for (int kg = 0; kg < 2; kg++)
                for (int part = 0; part < 2; part++)
                    for (int n = 0; n < 64; n++)
                        put_row(w3_img, (size_t)(tap * 2 + c) * W3_TILE + ((size_t)kg * 128 + part * 64 + n) * 16,
                                w3 + ((size_t)tap * 32 + 16 * c + 8 * kg) * 64 + n, 64, 8, part);
    // ---------------- fc1: [k = pos*64 + ch][n 48]; per (position, chunk) [k-group 2][w_hi 48 | w_lo 48][8 slots]
    std::vector<uint8_t> wf_img((size_t)27 * WF_POS_BYTES, 0);
    for (int pos = 0; pos < 27; pos++)
        for (int c = 0; c < 4; c++)
            for (int kg = 0; kg < 2; kg++)
                for (int part = 0; part < 2; part++)
                    for (int n = 0; n < 48; n++)
                        put_row(wf_img, (size_t)pos * WF_POS_BYTES + (size_t)c * WF_TILE + ((size_t)kg * 96 + part * 48 + n) * 16,
                                wf + ((size_t)pos * 64 + 16 * c + 8 * kg) * 48 + n, 48, 8, part);

    std::vector<uint8_t> img_a;
    img_a.insert(img_a.end(), w1_img.begin(), w1_img.end());
    img_a.insert(img_a.end(), w2_img.begin(), w2_img.end());
    const std::vector<uint8_t>& img_b = w3_img;
    std::vector<float> bias;
    bias.insert(bias.end(), b11, b11 + 16); bias.insert(bias.end(), b12, b12 + 16); bias.insert(bias.end(), b13, b13 + 16);
    bias.insert(bias.end(), b2, b2 + 32); bias.insert(bias.end(), b3, b3 + 64); bias.insert(bias.end(), bf, bf + 48);

    auto up = [&](DevBuf& d, const void* h, size_t bytes) -> cudaError_t {
        cudaError_t e = d.reserve(bytes);
        if (e != cudaSuccess) return e;
        return cudaMemcpyAsync(d.p, h, bytes, cudaMemcpyHostToDevice, stream);
    };
    cudaError_t e;
    if ((e = up(T.wimg_a, img_a.data(), img_a.size())) != cudaSuccess || (e = up(T.wimg_b, img_b.data(), img_b.size())) != cudaSuccess ||
        (e = up(T.wimg_c, wf_img.data(), wf_img.size())) != cudaSuccess ||
        (e = up(T.bias, bias.data(), bias.size() * 4)) != cudaSuccess || (e = T.err.reserve(16)) != cudaSuccess ||
        (e = cudaMemsetAsync(T.err.p, 0, 16, stream)) != cudaSuccess || (e = cudaStreamSynchronize(stream)) != cudaSuccess) {
        if (err) *err = std::string("tc_model_prepare: ") + cudaGetErrorString(e);
        return NC_ECUDA;
    }
    if (img_a.size() != (size_t)TA_SMEM_W) { if (err) *err = "TA weight image size mismatch"; return NC_EINVAL; }
    T.ready = true;
    return NC_OK;
}

// Runs TA -> TB -> TC over n sites.  `blob` = device pointer of the fp32 weight blob (tails).
// stop_after: 0 = full forward, 1 = after TA, 2 = after TB (debug entry points).
inline int tc_forward_ex(cudaStream_t stream, TcModel& T, int in_mode, const void* in_dev, int64_t in_site_stride, int64_t n,
                         const NcSiteMeta* meta, const float* ref4, const float* scale_f, const double* scale_d, const TailW& tw,
                         float* out_full, float* probs, int sm_count, uint64_t* launches, std::string* err, int stop_after,
                         cudaEvent_t ev_after_ta = nullptr) {
    using namespace tcg;
    if (!T.ready) return NC_ESTATE;
    if (n <= 0) return NC_OK;
    auto cuda_fail = [&](cudaError_t e, const char* what) { if (err) *err = std::string(what) + ": " + cudaGetErrorString(e); return NC_ECUDA; };
    cudaError_t e;
    const int64_t n_tiles = (n + 127) / 128;
    if ((e = T.c2.reserve((size_t)((n + 2) / 3) * C2_GROUP_BYTES)) != cudaSuccess) return cuda_fail(e, "c2 alloc");
    if ((e = T.c3.reserve((size_t)n_tiles * C3_TILE_BYTES)) != cudaSuccess) return cuda_fail(e, "c3 alloc");
    static bool attr_set[64] = {};                 // function attributes are per device: one flag per ordinal (one process may open several)
    int dev = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    const bool tracked = dev >= 0 && dev < 64;
    if (!tracked || !attr_set[dev]) {
        if ((e = cudaFuncSetAttribute(tc_trunk_a_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TA_SMEM)) != cudaSuccess) return cuda_fail(e, "TA smem attr");
        if ((e = cudaFuncSetAttribute(tc_trunk_b_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TB_SMEM)) != cudaSuccess) return cuda_fail(e, "TB smem attr");
        if ((e = cudaFuncSetAttribute(tc_trunk_a_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TA_SMEM)) != cudaSuccess) return cuda_fail(e, "TA smem attr");
        if ((e = cudaFuncSetAttribute(tc_trunk_b_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TB_SMEM)) != cudaSuccess) return cuda_fail(e, "TB smem attr");
        if ((e = cudaFuncSetAttribute(tc_fc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM)) != cudaSuccess) return cuda_fail(e, "TC smem attr");
        if (tracked) attr_set[dev] = true;
    }
    const float* bias = T.bias.as<float>();
    TAParams pa = {};
    pa.in = in_dev; pa.in_mode = in_mode; pa.in_site_stride = in_site_stride; pa.scale_f = scale_f; pa.scale_d = scale_d; pa.n_sites = n;
    pa.wimg = T.wimg_a.as<uint8_t>(); pa.bias1 = bias; pa.bias2 = bias + 48;
    pa.c2_out = T.c2.as<uint8_t>(); pa.err = T.err.as<int>();
    const unsigned ga = (unsigned)std::min<int64_t>((n + TA_WGS - 1) / TA_WGS, sm_count);
    if (T.audit) tc_trunk_a_kernel<true><<<ga, TA_THREADS, TA_SMEM, stream>>>(pa);
    else tc_trunk_a_kernel<false><<<ga, TA_THREADS, TA_SMEM, stream>>>(pa);
    if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail(e, "TA launch");
    (*launches)++;
    if (ev_after_ta && (e = cudaEventRecord(ev_after_ta, stream)) != cudaSuccess) return cuda_fail(e, "TA event");
    if (stop_after == 1) return NC_OK;
    TBParams pb = {};
    pb.c2 = T.c2.as<uint8_t>(); pb.n_sites = n; pb.wimg = T.wimg_b.as<uint8_t>(); pb.bias = bias + 80;
    pb.c3_out = T.c3.as<uint8_t>(); pb.err = T.err.as<int>();
    const int64_t n_groups = (n + 2) / 3;
    const unsigned gb = (unsigned)std::min<int64_t>((n_groups + TB_WGS - 1) / TB_WGS, sm_count);
    if (T.audit) tc_trunk_b_kernel<true><<<gb, TB_THREADS, TB_SMEM, stream>>>(pb);
    else tc_trunk_b_kernel<false><<<gb, TB_THREADS, TB_SMEM, stream>>>(pb);
    if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail(e, "TB launch");
    (*launches)++;
    if (stop_after == 2) return NC_OK;
    TCParams pc = {};
    pc.c3 = T.c3.as<uint8_t>(); pc.n_sites = n; pc.wimg = T.wimg_c.as<uint8_t>(); pc.bias = bias + 144;
    pc.tail = tw; pc.haploid = T.kind == 1; pc.meta = meta; pc.ref4 = ref4; pc.out10 = out_full; pc.probs4 = probs; pc.err = T.err.as<int>();
    if (pc.haploid && !pc.probs4) pc.probs4 = out_full;
    tc_fc_kernel<<<(unsigned)n_tiles, 128, TC_SMEM, stream>>>(pc);
    if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail(e, "TC launch");
    (*launches)++;
    return NC_OK;
}

}  // namespace nc
