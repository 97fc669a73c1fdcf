// CNN forward, fp32 CUDA-core path (impl = 1) for the four NanoCaller models:
//   SNP_model            model_architect.py:36-64             (kind 0)
//   haploid_SNP_model    model_architect_SNP_haploid.py:33-53 (kind 1)
//   Indel_model          model_architect_indel.py:28-48       (kind 2)
//   haploid_Indel_model  model_architect_indels_haploid.py:29-48 (kind 3)
// Keras semantics: Conv2D NHWC x HWIO, 'same' = symmetric zero pad (odd kernels, stride 1), 'valid',
// SELU, Flatten in (H, W, C) order, Dense = x @ K + b, softmax / sigmoid, Dropout = identity.
//
// Every conv / fc1 layer is one implicit-GEMM kernel (M = sites x output pixels, N = Cout,
// K = KH*KW*Cin) with bias + SELU fused in the epilogue; the three first-layer branches write into
// channel slices of one concat buffer.  For SNP models the first layer reads the int16 pileup
// tensor directly and applies the coverage scaling of snpCaller.py:90-96 in its operand load.
#pragma once
#include "nc_common.cuh"

namespace nc {

__device__ __forceinline__ float selu_f(float x) {
    const float scale = 1.0507009873554805f, alpha = 1.6732632423543772f;
    return x > 0.f ? scale * x : scale * alpha * expm1f(x);
}

struct ConvArgs {
    const void* in; const float* w; const float* bias; float* out;
    const float* scale_f; const double* scale_d;   // per-site coverage scale (IN_MODE 1 / 2)
    const int32_t* ktab;                           // per k: (dh << 24) | (dw << 16) | ci ; nullptr for 1x1
    int64_t M; int64_t in_site_stride;
    int32_t K, Hin, Win, Cin, SW, PH, PW, Hout, Wout, Cout, out_cstride, out_coff, selu;
};

// IN_MODE 0: fp32 NHWC input.  1: int16 SNP tensor, x[:,1:,:,:4] * fp32 scale (one fp32 rounding).
// 2: int16 SNP tensor, fp32(fp64(x) * fp64 ratio)  (--disable_coverage_normalization).
template <int BN, int IN_MODE>
__global__ void __launch_bounds__(16 * (BN / 4)) conv_f32_kernel(const ConvArgs a) {
    constexpr int BM = 64, BK = 16, T = 16 * (BN / 4);
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN];
    __shared__ int64_t s_base[BM];
    __shared__ int32_t s_h0[BM], s_w0[BM], s_site[BM];

    const int tid = threadIdx.x, tx = tid % (BN / 4), ty = tid / (BN / 4);
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int HW = a.Hout * a.Wout;
    for (int i = tid; i < BM; i += T) {
        const int64_t m = m0 + i;
        if (m < a.M) {
            const int64_t b = m / HW;
            const int r = (int)(m - b * HW), ho = r / a.Wout, wo = r - ho * a.Wout;
            s_base[i] = b * a.in_site_stride; s_site[i] = (int32_t)b;
            s_h0[i] = ho - a.PH; s_w0[i] = wo * a.SW - a.PW;
        } else {
            s_base[i] = 0; s_site[i] = 0; s_h0[i] = -(1 << 20); s_w0[i] = 0;
        }
    }
    __syncthreads();

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < a.K; k0 += BK) {
        for (int e = tid; e < BM * BK; e += T) {
            const int kk = e % BK, i = e / BK, k = k0 + kk;
            float v = 0.f;
            if (k < a.K) {
                int dh = 0, dw = 0, ci = k;
                if (a.ktab) { const int32_t t = __ldg(a.ktab + k); dh = t >> 24; dw = (t >> 16) & 255; ci = t & 0xFFFF; }
                const int hi = s_h0[i] + dh, wi = s_w0[i] + dw;
                if (hi >= 0 && hi < a.Hin && wi >= 0 && wi < a.Win) {
                    const int64_t idx = s_base[i] + (int64_t)(hi * a.Win + wi) * a.Cin + ci;
                    if (IN_MODE == 0) {
                        v = __ldg(reinterpret_cast<const float*>(a.in) + idx);
                    } else {
                        const int16_t raw = __ldg(reinterpret_cast<const int16_t*>(a.in) + idx);
                        v = (float)raw;
                        if (hi > 0 && ci < 4) {
                            if (IN_MODE == 1) v = __fmul_rn(v, __ldg(a.scale_f + s_site[i]));
                            else v = (float)((double)raw * __ldg(a.scale_d + s_site[i]));
                        }
                    }
                }
            }
            As[kk][i] = v;
        }
        for (int e = tid; e < BK * BN; e += T) {
            const int kk = e / BN, n = e - kk * BN, k = k0 + kk;
            Bs[kk][n] = k < a.K ? __ldg(a.w + (int64_t)k * a.Cout + n) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; kk++) {
            const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float am[4] = {av.x, av.y, av.z, av.w}, bn[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(am[i], bn[j], acc[i][j]);
        }
        __syncthreads();
    }
    const float4 bias = *reinterpret_cast<const float4*>(a.bias + tx * 4);
    const float bb[4] = {bias.x, bias.y, bias.z, bias.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int64_t m = m0 + ty * 4 + i;
        if (m < a.M) {
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; j++) { o[j] = acc[i][j] + bb[j]; if (a.selu) o[j] = selu_f(o[j]); }
            *reinterpret_cast<float4*>(a.out + m * a.out_cstride + a.out_coff + tx * 4) = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// tails (everything after fc1), one thread per site
// ------------------------------------------------------------------------------------------------
struct TailW {
    const float *fa_k, *fa_b, *hk[4], *hb[4], *fc2_k, *fc2_b, *fc3_k, *fc3_b, *gt_k, *gt_b;
};

template <int NIN, int NOUT>
__device__ __forceinline__ void dense_t(const float* __restrict__ x, const float* __restrict__ k, const float* __restrict__ b,
                                        float* __restrict__ y, bool act) {
#pragma unroll
    for (int o = 0; o < NOUT; o++) y[o] = 0.f;
#pragma unroll 4
    for (int i = 0; i < NIN; i++) {
        const float xi = x[i];
#pragma unroll
        for (int o = 0; o < NOUT; o++) y[o] = fmaf(xi, __ldg(k + i * NOUT + o), y[o]);
    }
#pragma unroll
    for (int o = 0; o < NOUT; o++) { y[o] += __ldg(b + o); if (act) y[o] = selu_f(y[o]); }
}
template <int N>
__device__ __forceinline__ void softmax_t(float* z) {
    float mx = z[0];
#pragma unroll
    for (int i = 1; i < N; i++) mx = fmaxf(mx, z[i]);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < N; i++) { z[i] = expf(z[i] - mx); s += z[i]; }
#pragma unroll
    for (int i = 0; i < N; i++) z[i] = z[i] / s;
}

// ref: either meta (ref_code -> one-hot) or a float [n][4] array.
// out10 (may be null): [n][10] = out_A out_G out_T out_C out_GT ; probs4 (may be null): [n][4] = P(A),P(G),P(T),P(C).
__global__ void __launch_bounds__(128) snp_tail_kernel(const float* __restrict__ f1, int64_t n, TailW w,
                                                       const NcSiteMeta* __restrict__ meta, const float* __restrict__ ref4,
                                                       float* __restrict__ out10, float* __restrict__ probs4) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    float x[48];
#pragma unroll
    for (int i = 0; i < 48; i += 4) {
        const float4 v = *reinterpret_cast<const float4*>(f1 + s * 48 + i);
        x[i] = v.x; x[i + 1] = v.y; x[i + 2] = v.z; x[i + 3] = v.w;
    }
    float ref[4];
    if (ref4) { for (int j = 0; j < 4; j++) ref[j] = ref4[s * 4 + j]; }
    else { const int rc = meta[s].ref_code; for (int j = 0; j < 4; j++) ref[j] = (j == rc) ? 1.f : 0.f; }
    float fa[17];
    dense_t<48, 16>(x, w.fa_k, w.fa_b, fa, true);
    float fc3in[24];
    dense_t<48, 16>(x, w.fc2_k, w.fc2_b, fc3in, true);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        float z[2];
        fa[16] = ref[j];
        dense_t<17, 2>(fa, w.hk[j], w.hb[j], z, false);
        softmax_t<2>(z);
        fc3in[16 + 2 * j] = z[0]; fc3in[17 + 2 * j] = z[1];
        if (out10) { out10[s * 10 + 2 * j] = z[0]; out10[s * 10 + 2 * j + 1] = z[1]; }
        if (probs4) probs4[s * 4 + j] = z[1];
    }
    if (out10) {
        float fc3[8], gt[2];
        dense_t<24, 8>(fc3in, w.fc3_k, w.fc3_b, fc3, true);
        dense_t<8, 2>(fc3, w.gt_k, w.gt_b, gt, false);
        softmax_t<2>(gt);
        out10[s * 10 + 8] = gt[0]; out10[s * 10 + 9] = gt[1];
    }
}

__global__ void __launch_bounds__(128) snp_hap_tail_kernel(const float* __restrict__ f1, int64_t n, TailW w,
                                                           const NcSiteMeta* __restrict__ meta, const float* __restrict__ ref4,
                                                           float* __restrict__ out4) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    float x[48];
#pragma unroll
    for (int i = 0; i < 48; i += 4) {
        const float4 v = *reinterpret_cast<const float4*>(f1 + s * 48 + i);
        x[i] = v.x; x[i + 1] = v.y; x[i + 2] = v.z; x[i + 3] = v.w;
    }
    float in[20], z[4];
    dense_t<48, 16>(x, w.fc2_k, w.fc2_b, in, true);
    if (ref4) { for (int j = 0; j < 4; j++) in[16 + j] = ref4[s * 4 + j]; }
    else { const int rc = meta[s].ref_code; for (int j = 0; j < 4; j++) in[16 + j] = (j == rc) ? 1.f : 0.f; }
    dense_t<20, 4>(in, w.fc3_k, w.fc3_b, z, true);           // Dense(4, selu) then softmax (:29,:51)
    softmax_t<4>(z);
#pragma unroll
    for (int j = 0; j < 4; j++) out4[s * 4 + j] = z[j];
}

template <bool HAPLOID>
__global__ void __launch_bounds__(128) indel_tail_kernel(const float* __restrict__ f1, int64_t n, TailW w, float* __restrict__ out) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    float x[32], h[24];
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        const float4 v = *reinterpret_cast<const float4*>(f1 + s * 32 + i);
        x[i] = v.x; x[i + 1] = v.y; x[i + 2] = v.z; x[i + 3] = v.w;
    }
    dense_t<32, 24>(x, w.fc2_k, w.fc2_b, h, true);
    if (HAPLOID) {
        float z[1];
        dense_t<24, 1>(h, w.fc3_k, w.fc3_b, z, false);
        out[s] = 1.f / (1.f + expf(-z[0]));
    } else {
        float z[4];
        dense_t<24, 4>(h, w.fc3_k, w.fc3_b, z, false);
        softmax_t<4>(z);
#pragma unroll
        for (int j = 0; j < 4; j++) out[s * 4 + j] = z[j];
    }
}

// per-site coverage scale (snpCaller.py:94-96): normalize -> fp32(train_coverage / chunk depth),
// otherwise the float64 ratio train_coverage / dp.
__global__ void site_scale_kernel(const NcSiteMeta* __restrict__ meta, int64_t n, const double* __restrict__ chunk_depth,
                                  double train_cov, int normalize, float* __restrict__ sf, double* __restrict__ sd) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    if (normalize) sf[s] = (float)(train_cov / chunk_depth[meta[s].chunk]);
    else sd[s] = train_cov / (double)meta[s].dp;
}

}  // namespace nc
