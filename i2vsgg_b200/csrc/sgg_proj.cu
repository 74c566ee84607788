// Data movement around the tensor-core FC kernel for the SGG projection (resnet_SGG_emb.py:128-221):
//   im2col        patches of conv_lo's three convolutions (resnet_SGG_emb.py:107-110,182-185) as bf16 rows, so each
//                 convolution is one FC launch whose output rows are already NHWC for the next layer;
//   pair rows     cat(index_select(obj, ix1), index_select(obj, ix2)) of resnet_SGG_emb.py:150-151,169 as bf16 rows.
#include <cuda_bf16.h>

#include "common.cuh"

namespace i2v {
namespace {

__device__ __forceinline__ float to_float(float v) { return v; }
__device__ __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }

// out[(n*OH + oy)*OW + ox][(ky*KW + kx)*C + c] = in[n, c, oy*stride - pad + ky, ox*stride - pad + kx] (0 outside),
// columns [KH*KW*C, ldo) are zero-filled so the row pitch can be padded to the TMA's 16-byte rule.
// `in` is addressed through element strides (sn, sc, sy, sx): NCHW fp32 masks and NHWC bf16 activations both fit.
__device__ __forceinline__ void put_out(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ void put_out(float* p, float v) { *p = v; }

template <typename InT, typename OutT = __nv_bfloat16>
__global__ void __launch_bounds__(256) im2col_kernel(const InT* __restrict__ in, OutT* __restrict__ out,
                                                     int64_t total, int C, int H, int W, int64_t sn, int64_t sc,
                                                     int64_t sy, int64_t sx, int KH, int KW, int stride, int pad, int OH,
                                                     int OW, int64_t ldo) {
    const int K = KH * KW * C;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int col = (int)(idx % ldo);
        int64_t row = idx / ldo;
        float v = 0.f;
        if (col < K) {
            int c = col % C;
            int kx = (col / C) % KW;
            int ky = col / (C * KW);
            int ox = (int)(row % OW);
            int oy = (int)((row / OW) % OH);
            int64_t n = row / ((int64_t)OW * OH);
            int y = oy * stride - pad + ky, x = ox * stride - pad + kx;
            if (y >= 0 && y < H && x >= 0 && x < W) v = to_float(in[n * sn + c * sc + y * sy + x * sx]);
        }
        put_out(out + idx, v);
    }
}

// Same mapping, eight output columns (one 16-byte store) per thread; `ldo` must be a multiple of 8.  When the input is
// NHWC bf16 with C % 8 == 0 the eight columns are eight consecutive channels of one tap: one 16-byte load.
template <typename InT>
__global__ void __launch_bounds__(256) im2col8_kernel(const InT* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                                      int64_t total8, int C, int H, int W, int64_t sn, int64_t sc,
                                                      int64_t sy, int64_t sx, int KH, int KW, int stride, int pad, int OH,
                                                      int OW, int64_t ldo, int vec_in) {
    const int K = KH * KW * C;
    const int chunks = (int)(ldo >> 3);
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total8;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int col0 = (int)(idx % chunks) * 8;
        const int64_t row = idx / chunks;
        const int ox = (int)(row % OW);
        const int oy = (int)((row / OW) % OH);
        const int64_t n = row / ((int64_t)OW * OH);
        uint4 pk = make_uint4(0u, 0u, 0u, 0u);
        if (vec_in) {
            if (col0 < K) {
                const int c = col0 % C, tap = col0 / C;
                const int kx = tap % KW, ky = tap / KW;
                const int y = oy * stride - pad + ky, x = ox * stride - pad + kx;
                if (y >= 0 && y < H && x >= 0 && x < W)
                    pk = *reinterpret_cast<const uint4*>(in + n * sn + c * sc + y * sy + x * sx);
            }
        } else {
            __nv_bfloat16 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int col = col0 + j;
                float f = 0.f;
                if (col < K) {
                    const int c = col % C, tap = col / C;
                    const int kx = tap % KW, ky = tap / KW;
                    const int y = oy * stride - pad + ky, x = ox * stride - pad + kx;
                    if (y >= 0 && y < H && x >= 0 && x < W) f = to_float(in[n * sn + c * sc + y * sy + x * sx]);
                }
                v[j] = __float2bfloat16_rn(f);
            }
            pk = *reinterpret_cast<uint4*>(v);
        }
        *reinterpret_cast<uint4*>(out + row * ldo + col0) = pk;
    }
}

// The same 16-byte-chunk mapping for a contiguous NHWC bf16 input with the geometry known at compile time (conv_lo's
// second layer: 96 channels, 5x5, stride 2, pad 2, 16x16 -> 8x8): every division is by a constant and the row pitch is
// exactly KH*KW*C, which takes the index arithmetic from ~100 to ~25 instructions per chunk.
template <int C, int K, int STRIDE, int PAD, int H, int W>
__global__ void __launch_bounds__(256) im2col8_fixed_kernel(const __nv_bfloat16* __restrict__ in,
                                                            __nv_bfloat16* __restrict__ out, int64_t total8) {
    constexpr int OH = (H + 2 * PAD - K) / STRIDE + 1, OW = (W + 2 * PAD - K) / STRIDE + 1;
    constexpr int CPT = C / 8, TAPS = K * K, CHUNKS = TAPS * CPT;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total8;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int chunk = (int)(idx % CHUNKS);
        const int64_t row = idx / CHUNKS;
        const int c8 = chunk % CPT, tap = chunk / CPT;
        const int kx = tap % K, ky = tap / K;
        const int pos = (int)(row % (OH * OW));
        const int64_t n = row / (OH * OW);
        const int ox = pos % OW, oy = pos / OW;
        const int y = oy * STRIDE - PAD + ky, x = ox * STRIDE - PAD + kx;
        uint4 pk = make_uint4(0u, 0u, 0u, 0u);
        if (y >= 0 && y < H && x >= 0 && x < W)
            pk = *reinterpret_cast<const uint4*>(in + ((n * H + y) * W + x) * C + c8 * 8);
        *reinterpret_cast<uint4*>(out + idx * 8) = pk;
    }
}

__global__ void __launch_bounds__(256) pair_rows_kernel(const float* __restrict__ obj, const int64_t* __restrict__ ixs,
                                                        const int64_t* __restrict__ ixo, __nv_bfloat16* __restrict__ out,
                                                        int64_t total, int num_obj, int E, int64_t ldo) {
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int col = (int)(idx % (2 * E));
        int64_t p = idx / (2 * E);
        int64_t o = col < E ? ixs[p] : ixo[p];
        float v = (o >= 0 && o < num_obj) ? obj[o * E + (col < E ? col : col - E)] : 0.f;
        out[p * ldo + col] = __float2bfloat16_rn(v);
    }
}


// out[p][:] = src[idx[p]][:] for bf16 rows, 16 bytes per thread (cols % 8 == 0, 16-byte aligned pitches).
__global__ void __launch_bounds__(256) gather_rows_kernel(const __nv_bfloat16* __restrict__ src, const int64_t* __restrict__ idx,
                                                          __nv_bfloat16* __restrict__ out, int64_t total8, int num_src,
                                                          int chunks, int64_t lds, int64_t ldo) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (int64_t)gridDim.x * blockDim.x) {
        const int ch = (int)(i % chunks);
        const int64_t p = i / chunks;
        const int64_t r = idx[p];
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (r >= 0 && r < num_src) v = *reinterpret_cast<const uint4*>(src + r * lds + ch * 8);
        *reinterpret_cast<uint4*>(out + p * ldo + ch * 8) = v;
    }
}


// conv_lo[0] for ordered pairs without the im2col of 4032 x 2 masks (resnet_SGG_emb.py:107,182): a pair's two input
// channels are the masks of its subject and its object, and a convolution is linear in its input channels, so
//     conv(pair)[pos][oc] = S[subject][pos][oc] + S[object][pos][C + oc] + bias[oc]
// with S[obj][pos][0..C) / [C..2C) the single-channel convolutions of the object's mask with the subject / object half of
// the kernel (one small FC launch over N objects instead of P pairs).  This kernel does the add, the ReLU and the bf16
// rounding, 8 output channels (16 bytes) per thread; out is NHWC [P, positions, C], or -- `ow` > 0, positions = oh x ow, both
// even -- the parity-split layout [P, 2, 2, oh/2, ow/2, C] (plane (y & 1, x & 1), position (y >> 1, x >> 1)) in which a
// stride-2 convolution reads dense boxes (i2v_conv2d_nhwc_split_forward).
constexpr int kPairRun = 8;     // consecutive pairs per thread: the pair list is subject-major, so a run mostly shares its subject
__global__ void __launch_bounds__(256) pair_conv1_kernel(const float* __restrict__ S, const int64_t* __restrict__ ixs,
                                                         const int64_t* __restrict__ ixo, const float* __restrict__ bias,
                                                         __nv_bfloat16* __restrict__ out, int64_t num_pairs, int num_obj,
                                                         int positions, int C, int relu, int ow) {
    const int chunks = C / 8;
    const int64_t per_pair = (int64_t)positions * chunks;
    const int64_t runs = (num_pairs + kPairRun - 1) / kPairRun;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < runs * per_pair; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t run = i / per_pair;
        const int e = (int)(i - run * per_pair);           // (position, 8-channel chunk) of this thread
        const int pos = e / chunks, ch = e - pos * chunks;
        int64_t dst_in_pair = e;
        if (ow > 0) {
            const int y = pos / ow, x = pos - y * ow;
            const int plane = ((y & 1) << 1) | (x & 1), sub = (y >> 1) * (ow >> 1) + (x >> 1);
            dst_in_pair = ((int64_t)plane * (positions >> 2) + sub) * chunks + ch;
        }
        float base[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) base[j] = bias ? __ldg(bias + ch * 8 + j) : 0.f;
        float sv[8];                                        // bias + the subject's half, kept while the subject stays
#pragma unroll
        for (int j = 0; j < 8; ++j) sv[j] = base[j];        // (what an out-of-range subject index leaves: -1 is one)
        int64_t cur = -1;
        const int64_t p_end = min(num_pairs, (run + 1) * kPairRun);
        for (int64_t p = run * kPairRun; p < p_end; ++p) {
            const int64_t a = ixs[p], b = ixo[p];
            if (a != cur) {
                cur = a;
#pragma unroll
                for (int j = 0; j < 8; ++j) sv[j] = base[j];
                if (a >= 0 && a < num_obj) {
                    const float4* s = reinterpret_cast<const float4*>(S + ((size_t)a * positions + pos) * 2 * C + ch * 8);
                    const float4 s0 = __ldg(s), s1 = __ldg(s + 1);
                    sv[0] += s0.x; sv[1] += s0.y; sv[2] += s0.z; sv[3] += s0.w;
                    sv[4] += s1.x; sv[5] += s1.y; sv[6] += s1.z; sv[7] += s1.w;
                }
            }
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = sv[j];
            if (b >= 0 && b < num_obj) {
                const float4* s = reinterpret_cast<const float4*>(S + ((size_t)b * positions + pos) * 2 * C + C + ch * 8);
                const float4 s0 = __ldg(s), s1 = __ldg(s + 1);
                v[0] += s0.x; v[1] += s0.y; v[2] += s0.z; v[3] += s0.w; v[4] += s1.x; v[5] += s1.y; v[6] += s1.z; v[7] += s1.w;
            }
            __nv_bfloat16 o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = __float2bfloat16_rn(relu ? fmaxf(v[j], 0.f) : v[j]);
            *reinterpret_cast<uint4*>(out + (p * per_pair + dst_in_pair) * 8) = *reinterpret_cast<uint4*>(o);
        }
    }
}

}  // namespace
}  // namespace i2v

using namespace i2v;

extern "C" int i2v_im2col_bf16(const void* in, int in_dtype, int n, int channels, int height, int width,
                               long long stride_n, long long stride_c, long long stride_y, long long stride_x,
                               int kernel_h, int kernel_w, int stride, int pad, void* out, long long ldo,
                               cudaStream_t stream) {
    I2V_REQUIRE(n >= 0 && channels >= 1 && height >= 1 && width >= 1 && kernel_h >= 1 && kernel_w >= 1 && stride >= 1 &&
                    pad >= 0,
                "im2col: bad shape");
    I2V_REQUIRE(in_dtype == I2V_DT_F32 || in_dtype == I2V_DT_BF16, "im2col: in_dtype %d", in_dtype);
    int OH = (height + 2 * pad - kernel_h) / stride + 1, OW = (width + 2 * pad - kernel_w) / stride + 1;
    I2V_REQUIRE(OH >= 1 && OW >= 1, "im2col: kernel larger than the padded input");
    I2V_REQUIRE(ldo >= (long long)kernel_h * kernel_w * channels, "im2col: row pitch smaller than a patch");
    int64_t total = (int64_t)n * OH * OW * ldo;
    if (total == 0) return I2V_OK;
    I2V_REQUIRE(in && out, "im2col: null pointer");
    if (ldo % 8 == 0 && ((uintptr_t)out & 15) == 0) {   // 16-byte stores
        int64_t total8 = total / 8;
        int grid8 = grid_for(total8, 256, 16);
        if (in_dtype == I2V_DT_BF16 && channels == 96 && kernel_h == 5 && kernel_w == 5 && stride == 2 && pad == 2 &&
            height == 16 && width == 16 && stride_c == 1 && stride_x == 96 && stride_y == 96 * 16 &&
            stride_n == 96 * 256 && ldo == 2400 && ((uintptr_t)in & 15) == 0) {
            im2col8_fixed_kernel<96, 5, 2, 2, 16, 16><<<grid8, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(in),
                                                                                 static_cast<__nv_bfloat16*>(out), total8);
            return check_launch("im2col8_fixed_kernel");
        }
        if (in_dtype == I2V_DT_F32) {
            im2col8_kernel<float><<<grid8, 256, 0, stream>>>(static_cast<const float*>(in), static_cast<__nv_bfloat16*>(out),
                                                             total8, channels, height, width, stride_n, stride_c, stride_y,
                                                             stride_x, kernel_h, kernel_w, stride, pad, OH, OW, ldo, 0);
        } else {
            int vec_in = stride_c == 1 && channels % 8 == 0 && stride_n % 8 == 0 && stride_y % 8 == 0 && stride_x % 8 == 0 &&
                         ((uintptr_t)in & 15) == 0;
            im2col8_kernel<__nv_bfloat16><<<grid8, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(in),
                                                                     static_cast<__nv_bfloat16*>(out), total8, channels,
                                                                     height, width, stride_n, stride_c, stride_y, stride_x,
                                                                     kernel_h, kernel_w, stride, pad, OH, OW, ldo, vec_in);
        }
        return check_launch("im2col8_kernel");
    }
    int grid = grid_for(total, 256, 16);
    if (in_dtype == I2V_DT_F32)
        im2col_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(in), static_cast<__nv_bfloat16*>(out),
                                                       total, channels, height, width, stride_n, stride_c, stride_y,
                                                       stride_x, kernel_h, kernel_w, stride, pad, OH, OW, ldo);
    else
        im2col_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(in),
                                                               static_cast<__nv_bfloat16*>(out), total, channels, height,
                                                               width, stride_n, stride_c, stride_y, stride_x, kernel_h,
                                                               kernel_w, stride, pad, OH, OW, ldo);
    return check_launch("im2col_kernel");
}

// fp32 -> the nearest tf32 value (10 mantissa bits), kept in an fp32 word.  tcgen05 kind::tf32 reads the upper 19 bits of
// its operands, i.e. it TRUNCATES; rounding the operands first removes the systematic shrink of every product (about
// 1e-3 per layer, which six layers of the relation head turned into 4e-3 of the feature scale).
__global__ void __launch_bounds__(256) round_tf32_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t rows,
                                                         int64_t cols, int64_t lds, int64_t ldd) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * cols; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols, c = i - r * cols;
        unsigned u;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(src[r * lds + c]));
        dst[r * ldd + c] = __uint_as_float(u);
    }
}
extern "C" int i2v_round_tf32(const float* src, float* dst, long long rows, long long cols, long long lds, long long ldd,
                              cudaStream_t stream) {
    I2V_REQUIRE(rows >= 0 && cols >= 0 && lds >= cols && ldd >= cols, "round_tf32: bad shape");
    if (rows == 0 || cols == 0) return I2V_OK;
    I2V_REQUIRE(src && dst, "round_tf32: null pointer");
    round_tf32_kernel<<<grid_for(rows * cols, 256), 256, 0, stream>>>(src, dst, rows, cols, lds, ldd);
    return check_launch("round_tf32_kernel");
}

// The same patches as fp32 rows (the tf32 precision of the relation head): fp32 or bf16 input, any strides.
extern "C" int i2v_im2col_f32(const void* in, int in_dtype, int n, int channels, int height, int width, long long stride_n,
                              long long stride_c, long long stride_y, long long stride_x, int kernel_h, int kernel_w,
                              int stride, int pad, float* out, long long ldo, cudaStream_t stream) {
    I2V_REQUIRE(n >= 0 && channels >= 1 && height >= 1 && width >= 1 && kernel_h >= 1 && kernel_w >= 1 && stride >= 1 &&
                    pad >= 0,
                "im2col: bad shape");
    I2V_REQUIRE(in_dtype == I2V_DT_F32 || in_dtype == I2V_DT_BF16, "im2col: in_dtype %d", in_dtype);
    int OH = (height + 2 * pad - kernel_h) / stride + 1, OW = (width + 2 * pad - kernel_w) / stride + 1;
    I2V_REQUIRE(OH >= 1 && OW >= 1, "im2col: kernel larger than the padded input");
    I2V_REQUIRE(ldo >= (long long)kernel_h * kernel_w * channels, "im2col: row pitch smaller than a patch");
    int64_t total = (int64_t)n * OH * OW * ldo;
    if (total == 0) return I2V_OK;
    I2V_REQUIRE(in && out, "im2col: null pointer");
    int grid = grid_for(total, 256, 16);
    if (in_dtype == I2V_DT_F32)
        im2col_kernel<float, float><<<grid, 256, 0, stream>>>(static_cast<const float*>(in), out, total, channels, height, width,
                                                              stride_n, stride_c, stride_y, stride_x, kernel_h, kernel_w,
                                                              stride, pad, OH, OW, ldo);
    else
        im2col_kernel<__nv_bfloat16, float><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(in), out, total,
                                                                      channels, height, width, stride_n, stride_c, stride_y,
                                                                      stride_x, kernel_h, kernel_w, stride, pad, OH, OW, ldo);
    return check_launch("im2col_kernel<f32>");
}

extern "C" int i2v_pair_rows_bf16(const float* obj, const int64_t* ixs, const int64_t* ixo, void* out, int num_obj,
                                  int num_pairs, int emb_dim, long long ldo, cudaStream_t stream) {
    I2V_REQUIRE(num_obj >= 0 && num_pairs >= 0 && emb_dim >= 1 && ldo >= 2LL * emb_dim, "pair_rows: bad shape");
    if (num_pairs == 0) return I2V_OK;
    I2V_REQUIRE(obj && ixs && ixo && out, "pair_rows: null pointer");
    int64_t total = (int64_t)num_pairs * 2 * emb_dim;
    pair_rows_kernel<<<grid_for(total, 256), 256, 0, stream>>>(obj, ixs, ixo, static_cast<__nv_bfloat16*>(out), total,
                                                               num_obj, emb_dim, ldo);
    return check_launch("pair_rows_kernel");
}

extern "C" int i2v_gather_rows_bf16(const void* src, const int64_t* idx, void* out, int num_src, int num_out, int cols,
                                    long long lds, long long ldo, cudaStream_t stream) {
    I2V_REQUIRE(num_src >= 0 && num_out >= 0 && cols >= 0 && lds >= cols && ldo >= cols, "gather_rows: bad shape");
    if (num_out == 0 || cols == 0) return I2V_OK;
    I2V_REQUIRE(src && idx && out, "gather_rows: null pointer");
    I2V_REQUIRE(cols % 8 == 0 && lds % 8 == 0 && ldo % 8 == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)out & 15) == 0,
                "gather_rows: rows must be multiples of 16 bytes with 16-byte aligned pitches");
    int64_t total8 = (int64_t)num_out * (cols / 8);
    gather_rows_kernel<<<grid_for(total8, 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(src), idx,
                                                                  static_cast<__nv_bfloat16*>(out), total8, num_src, cols / 8,
                                                                  lds, ldo);
    return check_launch("gather_rows_kernel");
}

static int pair_conv1_impl(const float* obj_maps, const int64_t* ixs, const int64_t* ixo, const float* bias, void* out,
                           int num_obj, int num_pairs, int positions, int channels, int relu, int ow, cudaStream_t stream) {
    I2V_REQUIRE(num_obj >= 0 && num_pairs >= 0 && positions >= 1 && channels >= 8 && channels % 8 == 0,
                "pair_conv1: bad shape (channels must be a multiple of 8)");
    if (num_pairs == 0) return I2V_OK;
    I2V_REQUIRE(obj_maps && ixs && ixo && out, "pair_conv1: null pointer");
    I2V_REQUIRE(((uintptr_t)obj_maps & 15) == 0 && ((uintptr_t)out & 15) == 0, "pair_conv1: 16-byte aligned buffers needed");
    const int64_t threads = (int64_t)ceil_div(num_pairs, kPairRun) * positions * (channels / 8);
    pair_conv1_kernel<<<grid_for(threads, 256, 16), 256, 0, stream>>>(obj_maps, ixs, ixo, bias,
                                                                      static_cast<__nv_bfloat16*>(out), num_pairs, num_obj,
                                                                      positions, channels, relu, ow);
    return check_launch("pair_conv1_kernel");
}

extern "C" int i2v_pair_conv1_bf16(const float* obj_maps, const int64_t* ixs, const int64_t* ixo, const float* bias, void* out,
                                   int num_obj, int num_pairs, int positions, int channels, int relu, cudaStream_t stream) {
    return pair_conv1_impl(obj_maps, ixs, ixo, bias, out, num_obj, num_pairs, positions, channels, relu, 0, stream);
}

// the same rows in the parity-split layout [P, 2, 2, oh/2, ow/2, C] (see pair_conv1_kernel)
extern "C" int i2v_pair_conv1_split_bf16(const float* obj_maps, const int64_t* ixs, const int64_t* ixo, const float* bias,
                                         void* out, int num_obj, int num_pairs, int oh, int ow, int channels, int relu,
                                         cudaStream_t stream) {
    I2V_REQUIRE(oh >= 2 && ow >= 2 && oh % 2 == 0 && ow % 2 == 0, "pair_conv1_split: the map must have even sides");
    return pair_conv1_impl(obj_maps, ixs, ixo, bias, out, num_obj, num_pairs, oh * ow, channels, relu, ow, stream);
}
