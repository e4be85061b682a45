// vrt_denoise.cu — the reference's present pass (assets/shaders/image.frag:31-79, "sirBirdDenoise") as an sm_100a kernel.
//
// One thread per output texel.  What the fragment shader recomputes for every fragment but depends only on the push
// constants — the golden-angle spiral offsets (:49-53) and the radial part of the sample weight (:52) — is computed once per
// CTA into shared memory (same operations, same order, so the values are the ones every fragment would have got).  The UNORM
// decode byte / 255.0f is done once per texel by a pre-pass into a float4 image (33 MB at 1080p, L2 resident) instead of 12
// times per sample.  The per-sample work left is the bilinear fetch (four 128-bit loads), one sqrt + one divide for
// normalize / length, and two pow.
//
// Fast path (denoise_padded_kernel — every configuration whose largest sample offset stays inside kDenoisePad texels and whose hue
// tolerance is a whole number 2..64, i.e. the reference's defaults): the decoded image carries a kDenoisePad-texel border filled by
// the sampler's repeat addressing, so a bilinear tap is one index computation and four loads at fixed offsets — no wrap selects,
// no second row / column index — and both pow are multiplication chains without the range checks (max(a,0)^n by products gives the
// guarded results by itself: 0 for a < FLT_MIN, inf for inf, NaN for NaN).  Everything else takes denoise_kernel, the general one.
//
// Arithmetic discipline = the oracle's (oracle/vrt_oracle_denoise.cpp header): FP32, no contraction (--fmad=false) except
// the explicit fmaf of det_log2f / det_exp2f, IEEE sqrt and divide, FP32 bilinear weights.  Bound: FP32 / SFU issue, not HBM
// (algorithmic traffic is 4 B read + 4 B written per pixel).
#include <cmath>
#include <cstdint>

#include "../../include/vrt.h"
#include "vrt_kernels.cuh"

namespace vrt {

namespace {

constexpr int kDnBlockX = 32, kDnBlockY = 8;
constexpr int kDnMaxSamples = 255;

__device__ __forceinline__ float dn_max(float a, float b) { return a < b ? b : a; }

// log2 of a finite a >= FLT_MIN: a = 2^e * m, m in [sqrt(1/2), sqrt(2)); ln m = 2 atanh((m-1)/(m+1))
__device__ __forceinline__ float det_log2f(float a) {
    const int ia = __float_as_int(a);
    const int e = (ia - 0x3f3504f3) >> 23;
    const float m = __int_as_float(ia - e * (1 << 23));
    const float f = m - 1.0f;
    // f == 0 (a is a power of two — pow(1, b) is the common case on flat image regions) gives s = 0 / 2 = 0; a zero numerator
    // would send the whole warp through the division's special-case subroutine, so it is divided as 1 / 2 and selected away
    float num = f == 0.0f ? 1.0f : f;
    asm volatile("" : "+f"(num));  // keep the compiler from folding the substitution back into a division of f itself
    const float q = num / (2.0f + f);
    const float s = f == 0.0f ? 0.0f : q;
    const float z = s * s;
    float p = fmaf(z, 0.22222222f, 0.2857143f);
    p = fmaf(p, z, 0.4f);
    p = fmaf(p, z, 0.6666667f);
    const float ln_m = fmaf(s * z, p, s + s);
    return fmaf(ln_m, 1.442695f, (float)e);
}

// 2^y, y clamped to [-126, 126]; NaN stays NaN
__device__ __forceinline__ float det_exp2f(float y) {
    if (y != y) return y;
    y = y < -126.0f ? -126.0f : (y > 126.0f ? 126.0f : y);
    const float n = floorf(y + 0.5f);
    const float r = y - n;
    float p = fmaf(0.0001540353f, r, 0.0013333558f);
    p = fmaf(p, r, 0.009618129f);
    p = fmaf(p, r, 0.05550411f);
    p = fmaf(p, r, 0.2402265f);
    p = fmaf(p, r, 0.6931472f);
    p = fmaf(p, r, 1.0f);
    return p * __int_as_float(((int)n + 127) << 23);
}

// image.frag:29  #define pow(a,b) pow(max(a,0.),b)
// x^n for a whole n >= 1: binary exponentiation, lowest bit first — the oracle's fixed order of exact FP32 products
__device__ __forceinline__ float det_powif(float a, unsigned n) {
    if (n == 20u) {  // the default hue tolerance: the loop below unrolled for 10100b — the same five products in the same order
        const float p2 = a * a, p4 = p2 * p2, p8 = p4 * p4, p16 = p8 * p8;
        return p4 * p16;
    }
    float r = 1.0f, p = a;
    while (n) {
        if (n & 1u) r = r * p;
        n >>= 1;
        if (n) p = p * p;
    }
    return r;
}
__device__ __forceinline__ float gpow(float a, float b) {
    a = dn_max(a, 0.0f);
    if (a != a) return a;
    if (a < 1.17549435e-38f) return 0.0f;
    if (a > 3.4028234e38f) return a;
    if (b >= 1.0f && b <= 64.0f && b == floorf(b)) return det_powif(a, (unsigned)b);  // whole exponents (8, 20): a few multiplications instead of log2 + exp2
    return det_exp2f(b * det_log2f(a));
}

// len = sqrt(d), inv = 1 / len for d >= 0 (or NaN): identical to the IEEE operators, but zero — frequent: black texels — takes
// the fast path with a substituted operand and the exact results (0 and +inf) selected afterwards.
__device__ __forceinline__ void safe_len_inv(float d, float& len, float& inv) {
    const bool zero = d == 0.0f;
    float ds = zero ? 1.0f : d;
    asm volatile("" : "+f"(ds));
    const float l = sqrtf(ds);
    float ls = l;
    asm volatile("" : "+f"(ls));
    const float r = 1.0f / ls;
    len = zero ? 0.0f : l;
    inv = zero ? __int_as_float(0x7f800000) : r;
}

struct Rgb {
    float x, y, z;
};

// texture(imageSampler, uv).rgb: linear filter, repeat addressing (Pipeline.zig:193-212); `unorm` = byte / 255.0f table.
// NEAR: the coordinate is known to lie within one image size of the image (launch_denoise checks the largest sample offset), so
// the repeat wrap is one conditional add / subtract instead of an integer modulo.
template <bool NEAR>
__device__ __forceinline__ Rgb sample_linear_repeat(const float4* __restrict__ img, int w, int h, float fw, float fh, float u, float v) {
    const float x = u * fw - 0.5f, y = v * fh - 0.5f;
    const float fx = floorf(x), fy = floorf(y);
    const float a = x - fx, b = y - fy;
    int x0 = (int)fx, y0 = (int)fy;
    if (NEAR) {
        x0 = x0 < 0 ? x0 + w : (x0 >= w ? x0 - w : x0), y0 = y0 < 0 ? y0 + h : (y0 >= h ? y0 - h : y0);
    } else {
        x0 %= w, y0 %= h;
        x0 = x0 < 0 ? x0 + w : x0, y0 = y0 < 0 ? y0 + h : y0;
    }
    const int x1 = x0 + 1 == w ? 0 : x0 + 1, y1 = y0 + 1 == h ? 0 : y0 + 1;
    const float4* row0 = img + (size_t)y0 * w;
    const float4* row1 = img + (size_t)y1 * w;
    const float4 t00 = __ldg(row0 + x0), t10 = __ldg(row0 + x1), t01 = __ldg(row1 + x0), t11 = __ldg(row1 + x1);
    const float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
    Rgb c;
    c.x = ((w00 * t00.x + w10 * t10.x) + w01 * t01.x) + w11 * t11.x;
    c.y = ((w00 * t00.y + w10 * t10.y) + w01 * t01.y) + w11 * t11.y;
    c.z = ((w00 * t00.z + w10 * t10.z) + w01 * t01.z) + w11 * t11.z;
    return c;
}

// UNORM decode of the whole image, once: texel / 255.0f per channel (IEEE divide, as the oracle's texel())
__global__ void __launch_bounds__(256) decode_unorm_kernel(const uint32_t* __restrict__ img, float4* __restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = __ldg(img + i);
    out[i] = make_float4((float)(p & 255u) / 255.0f, (float)((p >> 8) & 255u) / 255.0f, (float)((p >> 16) & 255u) / 255.0f, (float)(p >> 24) / 255.0f);
}

__device__ __forceinline__ uint32_t dn_unorm8(float c) {
    if (!(c == c)) return 0u;
    c = c < 0.0f ? 0.0f : (c > 1.0f ? 1.0f : c);
    return (uint32_t)(uint8_t)(c * 255.0f + 0.5f);
}

constexpr float kCosGolden = -0.7373688f, kSinGolden = 0.6754904f;  // cos / sin(2.3999632), image.frag:25,29

// What every fragment derives from the push constants alone (:35-37, :45-53) — the rotated, scaled sample offsets and the radial
// weight term of each of the samples + 1 taps — once per launch, with the fragment's own operation order: table[0..255] = offset x,
// [256..511] = offset y, [512..767] = radial term.  (It used to be recomputed by every CTA behind a barrier: 18 % of the pass's stall
// samples, profiles/r02_denoise_1080p_summary.json.)
__global__ void __launch_bounds__(kDnMaxSamples + 1) denoise_table_kernel(int w, int h, const vrt_denoise_params pc, float* __restrict__ table) {
    const int tid = (int)threadIdx.x;
    if (tid <= pc.samples) {
        const float sample_radius = sqrtf((float)pc.samples);                       // :35
        const float sample_true_radius = 0.5f / (sample_radius * sample_radius);    // :36
        float rot_x = 0.0f, rot_y = 1.0f;                                            // :45
        for (int i = 0; i <= tid; i++) {                                             // :49, applied tid + 1 times
            const float nx = rot_x * kCosGolden + rot_y * kSinGolden;
            const float ny = rot_x * (-kSinGolden) + rot_y * kCosGolden;
            rot_x = nx, rot_y = ny;
        }
        const float sq = sqrtf((float)tid);
        const float off_x = ((pc.pixel_multiplier * rot_x) * sq) * 0.5f, off_y = ((pc.pixel_multiplier * rot_y) * sq) * 0.5f;  // :51
        table[2 * (kDnMaxSamples + 1) + tid] = 1.0f - sample_true_radius * gpow(off_x * off_x + off_y * off_y, pc.distribution_bias);  // :52
        table[tid] = off_x * (1.0f / (float)w), table[kDnMaxSamples + 1 + tid] = off_y * (1.0f / (float)h);                             // :37,:53
    }
}

template <bool NEAR>
__global__ void __launch_bounds__(kDnBlockX* kDnBlockY) denoise_kernel(const float4* __restrict__ img, int w, int h, const vrt_denoise_params pc, uint32_t* __restrict__ out,
                                                                        uint32_t out_w, uint32_t out_h, uint32_t bgra, const float* __restrict__ table) {
    __shared__ float s_off_x[kDnMaxSamples + 1], s_off_y[kDnMaxSamples + 1], s_radial[kDnMaxSamples + 1];
    const int tid = threadIdx.y * kDnBlockX + threadIdx.x;
    if (tid <= pc.samples) s_off_x[tid] = table[tid], s_off_y[tid] = table[kDnMaxSamples + 1 + tid], s_radial[tid] = table[2 * (kDnMaxSamples + 1) + tid];
    __syncthreads();

    const uint32_t ox = blockIdx.x * kDnBlockX + threadIdx.x, oy = blockIdx.y * kDnBlockY + threadIdx.y;
    if (ox >= out_w || oy >= out_h) return;
    const float fw = (float)w, fh = (float)h;
    const float uvx = ((float)ox + 0.5f) / (float)out_w, uvy = ((float)oy + 0.5f) / (float)out_h;

    const Rgb center = sample_linear_repeat<NEAR>(img, w, h, fw, fh, uvx, uvy);  // :38
    const float center_sat = sqrtf((center.x * center.x + center.y * center.y) + center.z * center.z);  // :40
    const float center_inv = 1.0f / center_sat;                                                         // :39 normalize
    const float cnx = center.x * center_inv, cny = center.y * center_inv, cnz = center.z * center_inv;
    const float abs_center_sat = fabsf(center_sat);  // length(float), :63
    float dx = 0.0f, dy = 0.0f, dz = 0.0f, influence_sum = 0.0f;
    const int samples = pc.samples;
    for (int k = 0; k <= samples; k++) {  // :47
        const Rgb c = sample_linear_repeat<NEAR>(img, w, h, fw, fh, uvx + s_off_x[k], uvy + s_off_y[k]);  // :55
        float influence = s_radial[k];
        influence *= influence * influence;  // :57
        // a black texel (a shadowed pixel of the traced frame) has length 0 and 1 / 0 = inf: same values as the plain
        // sqrt / divide, computed without their special-case subroutines (safe_len_inv)
        float len, inv;
        safe_len_inv((c.x * c.x + c.y * c.y) + c.z * c.z, len, inv);
        const float d = (cnx * (c.x * inv) + cny * (c.y * inv)) + cnz * (c.z * inv);
        influence *= gpow(0.5f + 0.5f * d, pc.inverse_hue_tolerance) * gpow(1.0f - fabsf(len - abs_center_sat), 8.0f);  // :61-64
        influence_sum += influence;                                                                                     // :66
        dx += c.x * influence, dy += c.y * influence, dz += c.z * influence;                                            // :67
    }
    const uint32_t r = dn_unorm8(dx / influence_sum), g = dn_unorm8(dy / influence_sum), b = dn_unorm8(dz / influence_sum);  // :70, :77
    out[(size_t)oy * out_w + ox] = bgra ? (b | (g << 8) | (r << 16) | 0xff000000u) : (r | (g << 8) | (b << 16) | 0xff000000u);
}

// UNORM decode into the padded image: (w + 2 pad) x (h + 2 pad) texels, padded texel (px, py) = image texel ((px - pad) mod w,
// (py - pad) mod h) — the sampler's repeat addressing (Pipeline.zig:193-212) applied once per border texel instead of per tap.
__global__ void __launch_bounds__(256) decode_unorm_padded_kernel(const uint32_t* __restrict__ img, float4* __restrict__ out, int w, int h, int pad) {
    const int pw = w + 2 * pad, ph = h + 2 * pad;
    const int px = (int)(blockIdx.x * 256 + threadIdx.x), py = (int)blockIdx.y;
    if (px >= pw || py >= ph) return;
    int sx = (px - pad) % w, sy = (py - pad) % h;
    sx = sx < 0 ? sx + w : sx, sy = sy < 0 ? sy + h : sy;
    const uint32_t p = __ldg(img + (size_t)sy * w + sx);
    out[(size_t)py * pw + px] = make_float4((float)(p & 255u) / 255.0f, (float)((p >> 8) & 255u) / 255.0f, (float)((p >> 16) & 255u) / 255.0f, (float)(p >> 24) / 255.0f);
}

// The table of denoise_table_kernel, one float4 per tap: offset x, offset y, the radial term cubed (:57, same two products).
__global__ void __launch_bounds__(kDnMaxSamples + 1) denoise_pack_table_kernel(const float* __restrict__ table, float4* __restrict__ packed, int samples) {
    const int tid = (int)threadIdx.x;
    if (tid <= samples) {
        float influence = table[2 * (kDnMaxSamples + 1) + tid];
        influence *= influence * influence;
        packed[tid] = make_float4(table[tid], table[kDnMaxSamples + 1 + tid], influence, 0.0f);
    }
}

// max(a, 0)^n for a whole n in 2..64 (image.frag:29 with a whole exponent): a NaN stays a NaN through the products, a < FLT_MIN
// squares to 0, inf stays inf — gpow's guarded results without the guards
template <unsigned N>
__device__ __forceinline__ float gpow_whole(float a, unsigned n) {
    a = a < 0.0f ? 0.0f : a;
    return det_powif(a, N ? N : n);
}

// texture(imageSampler, uv).rgb on the padded image: `origin` = padded texel (pad, pad), `stride` = padded row length
__device__ __forceinline__ Rgb sample_padded(const float4* __restrict__ origin, int stride, float fw, float fh, float u, float v) {
    const float x = u * fw - 0.5f, y = v * fh - 0.5f;
    const float fx = floorf(x), fy = floorf(y);
    const float a = x - fx, b = y - fy;
    const float4* p = origin + ((int)fy * stride + (int)fx);
    const float4 t00 = __ldg(p), t10 = __ldg(p + 1), t01 = __ldg(p + stride), t11 = __ldg(p + stride + 1);
    const float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
    Rgb c;
    c.x = ((w00 * t00.x + w10 * t10.x) + w01 * t01.x) + w11 * t11.x;
    c.y = ((w00 * t00.y + w10 * t10.y) + w01 * t01.y) + w11 * t11.y;
    c.z = ((w00 * t00.z + w10 * t10.z) + w01 * t01.z) + w11 * t11.z;
    return c;
}

// HUE = the hue tolerance when it is the default 20 (its product chain unrolled), 0 = any whole exponent 2..64 (hue_n)
template <unsigned HUE>
__global__ void __launch_bounds__(kDnBlockX* kDnBlockY) denoise_padded_kernel(const float4* __restrict__ origin, int stride, int w, int h, int samples, unsigned hue_n,
                                                                               uint32_t* __restrict__ out, uint32_t out_w, uint32_t out_h, uint32_t bgra,
                                                                               const float4* __restrict__ packed) {
    __shared__ float4 s_tap[kDnMaxSamples + 1];
    const int tid = threadIdx.y * kDnBlockX + threadIdx.x;
    if (tid <= samples) s_tap[tid] = packed[tid];
    __syncthreads();

    const uint32_t ox = blockIdx.x * kDnBlockX + threadIdx.x, oy = blockIdx.y * kDnBlockY + threadIdx.y;
    if (ox >= out_w || oy >= out_h) return;
    const float fw = (float)w, fh = (float)h;
    const float uvx = ((float)ox + 0.5f) / (float)out_w, uvy = ((float)oy + 0.5f) / (float)out_h;

    const Rgb center = sample_padded(origin, stride, fw, fh, uvx, uvy);                                 // :38
    const float center_sat = sqrtf((center.x * center.x + center.y * center.y) + center.z * center.z);  // :40
    const float center_inv = 1.0f / center_sat;                                                         // :39 normalize
    const float cnx = center.x * center_inv, cny = center.y * center_inv, cnz = center.z * center_inv;
    const float abs_center_sat = fabsf(center_sat);  // length(float), :63
    float dx = 0.0f, dy = 0.0f, dz = 0.0f, influence_sum = 0.0f;
    for (int k = 0; k <= samples; k++) {  // :47
        const float4 tap = s_tap[k];
        const Rgb c = sample_padded(origin, stride, fw, fh, uvx + tap.x, uvy + tap.y);  // :55
        float len, inv;
        safe_len_inv((c.x * c.x + c.y * c.y) + c.z * c.z, len, inv);
        const float d = (cnx * (c.x * inv) + cny * (c.y * inv)) + cnz * (c.z * inv);
        const float influence = tap.z * (gpow_whole<HUE>(0.5f + 0.5f * d, hue_n) * gpow_whole<8>(1.0f - fabsf(len - abs_center_sat), 8u));  // :57-64
        influence_sum += influence;                                                                                                       // :66
        dx += c.x * influence, dy += c.y * influence, dz += c.z * influence;                                                              // :67
    }
    const uint32_t r = dn_unorm8(dx / influence_sum), g = dn_unorm8(dy / influence_sum), b = dn_unorm8(dz / influence_sum);  // :70, :77
    out[(size_t)oy * out_w + ox] = bgra ? (b | (g << 8) | (r << 16) | 0xff000000u) : (r | (g << 8) | (b << 16) | 0xff000000u);
}

}  // namespace

cudaError_t launch_denoise(const uint32_t* image, float4* decoded, uint32_t width, uint32_t height, const vrt_denoise_params& params, uint32_t* out,
                           uint32_t out_width, uint32_t out_height, bool bgra, cudaStream_t stream, LaunchInfo* info) {
    const size_t n = (size_t)width * height;
    const dim3 block(kDnBlockX, kDnBlockY);
    const dim3 grid((out_width + kDnBlockX - 1) / kDnBlockX, (out_height + kDnBlockY - 1) / kDnBlockY);
    // largest sample offset in input texels (:51): |pixelMultiplier| * sqrt(samples) * 0.5, plus the bilinear footprint and slack
    const float reach = fabsf(params.pixel_multiplier) * sqrtf((float)params.samples) * 0.5f + 3.0f;
    const float hue = params.inverse_hue_tolerance;
    const bool padded = reach <= (float)kDenoisePad && hue >= 2.0f && hue <= 64.0f && hue == floorf(hue) && params.samples >= 0 && params.samples <= kDnMaxSamples;
    if (padded) {
        const int pw = (int)width + 2 * kDenoisePad, ph = (int)height + 2 * kDenoisePad;
        float* table = reinterpret_cast<float*>(decoded + (size_t)pw * ph);  // kDenoiseScratchTail float4 behind the padded image: 3 x 256 floats, then 256 packed taps
        float4* packed = decoded + (size_t)pw * ph + 192;
        decode_unorm_padded_kernel<<<dim3((unsigned)((pw + 255) / 256), (unsigned)ph), 256, 0, stream>>>(image, decoded, (int)width, (int)height, kDenoisePad);
        denoise_table_kernel<<<1, kDnMaxSamples + 1, 0, stream>>>((int)width, (int)height, params, table);
        denoise_pack_table_kernel<<<1, kDnMaxSamples + 1, 0, stream>>>(table, packed, params.samples);
        const float4* origin = decoded + (size_t)kDenoisePad * pw + kDenoisePad;
        if (hue == 20.0f)
            denoise_padded_kernel<20><<<grid, block, 0, stream>>>(origin, pw, (int)width, (int)height, params.samples, 20u, out, out_width, out_height, bgra ? 1u : 0u, packed);
        else
            denoise_padded_kernel<0><<<grid, block, 0, stream>>>(origin, pw, (int)width, (int)height, params.samples, (unsigned)hue, out, out_width, out_height, bgra ? 1u : 0u, packed);
        if (info) info->launches += 4;
        return cudaGetLastError();
    }
    decode_unorm_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(image, decoded, n);
    float* table = reinterpret_cast<float*>(decoded + n);  // 3 x 256 floats behind the decoded image (kDenoiseScratchTail)
    denoise_table_kernel<<<1, kDnMaxSamples + 1, 0, stream>>>((int)width, (int)height, params, table);
    if (info) info->launches += 2;
    const bool near = reach < (float)(width < height ? width : height);  // false also for a NaN multiplier
    if (near)
        denoise_kernel<true><<<grid, block, 0, stream>>>(decoded, (int)width, (int)height, params, out, out_width, out_height, bgra ? 1u : 0u, table);
    else
        denoise_kernel<false><<<grid, block, 0, stream>>>(decoded, (int)width, (int)height, params, out, out_width, out_height, bgra ? 1u : 0u, table);
    if (info) info->launches++;
    return cudaGetLastError();
}

}  // namespace vrt
