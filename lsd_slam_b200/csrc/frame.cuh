// frame.cuh -- per-frame staging kernels: u8 -> f32 image pyramid, gradient pyramid, maxGradients,
// ground-truth depth import and the idepth / idepthVar pyramids.
//
// Replaces (paths relative to lsd_slam_core/src/):
//   Frame::Frame(.., const uchar*)        DataStructures/Frame.cpp:35-54
//   Frame::buildImage                     DataStructures/Frame.cpp:491-630
//   Frame::buildGradients                 DataStructures/Frame.cpp:643-680
//   Frame::buildMaxGradients              DataStructures/Frame.cpp:690-767
//   Frame::setDepthFromGroundTruth        DataStructures/Frame.cpp:245-293
//   Frame::buildIDepthAndIDepthVar        DataStructures/Frame.cpp:775-877
// All of these are streaming, HBM-bound passes over at most 16 B/px; the whole pyramid of one frame is built
// by ONE launch per pass (the reference builds level by level, lazily).
#pragma once
#include "internal.cuh"

struct PyrPtrs {
    float* l[LSD_LEVELS];
};
// One CTA = one 16x16 level-0 tile -> 8x8, 4x4, 2x2, 1x1 on levels 1..4 (w, h are multiples of 16,
// SlamSystem.cpp:55).  The 2x2 box sum is exact in fp32 for u8-origin data and is
// associated like the scalar loop (Frame.cpp:621-624).
// REMAP = true fuses UndistorterPTAM::undistort (util/Undistorter.cpp:380-410) in front: level 0 is gathered bilinearly from
// the RAW (distorted, in_w wide) image through the remap tables, truncated to 8 bits exactly as the reference's
// `data[idx] = float expression` does, and converted to float as Frame::Frame does (Frame.cpp:44-52) -- the undistorted u8 image
// never exists in memory unless `undist` is given (parity hook / host consumers).
__device__ __forceinline__ unsigned char undistortPixel(const uint8_t* __restrict__ raw, int in_w, float xx, float yy)
{
    if (xx < 0) return 0;
    const int xxi = (int)xx, yyi = (int)yy;
    xx -= xxi;
    yy -= yyi;
    const float xxyy = xx * yy;
    const uint8_t* src = raw + xxi + yyi * in_w;
    const float v = xxyy * src[1 + in_w] + (yy - xxyy) * src[in_w] + (xx - xxyy) * src[1] + (1 - xx - yy + xxyy) * src[0];
    return (unsigned char)v;
}
template <bool REMAP>
__global__ void __launch_bounds__(256) k_image_pyramid(const uint8_t* __restrict__ src, PyrPtrs p, int w, int h,
                                                       const float* __restrict__ remapX = nullptr, const float* __restrict__ remapY = nullptr,
                                                       int in_w = 0, uint8_t* __restrict__ undist = nullptr)
{
    __shared__ float s0[16][17], s1[8][9], s2[4][5], s3[2][3];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int bx = blockIdx.x, by = blockIdx.y;
    const int x = bx * 16 + tx, y = by * 16 + ty;
    float v;
    if (REMAP) {
        const unsigned char u = undistortPixel(src, in_w, __ldg(remapX + y * w + x), __ldg(remapY + y * w + x));
        if (undist) undist[y * w + x] = u;
        v = (float)u;
    } else
        v = (float)src[y * w + x];
    p.l[0][y * w + x] = v;
    s0[ty][tx] = v;
    __syncthreads();
    if (tx < 8 && ty < 8) {
        float a = (s0[2 * ty][2 * tx] + s0[2 * ty][2 * tx + 1] + s0[2 * ty + 1][2 * tx] + s0[2 * ty + 1][2 * tx + 1]) * 0.25f;
        s1[ty][tx] = a;
        p.l[1][(by * 8 + ty) * (w >> 1) + bx * 8 + tx] = a;
    }
    __syncthreads();
    if (tx < 4 && ty < 4) {
        float a = (s1[2 * ty][2 * tx] + s1[2 * ty][2 * tx + 1] + s1[2 * ty + 1][2 * tx] + s1[2 * ty + 1][2 * tx + 1]) * 0.25f;
        s2[ty][tx] = a;
        p.l[2][(by * 4 + ty) * (w >> 2) + bx * 4 + tx] = a;
    }
    __syncthreads();
    if (tx < 2 && ty < 2) {
        float a = (s2[2 * ty][2 * tx] + s2[2 * ty][2 * tx + 1] + s2[2 * ty + 1][2 * tx] + s2[2 * ty + 1][2 * tx + 1]) * 0.25f;
        s3[ty][tx] = a;
        p.l[3][(by * 2 + ty) * (w >> 3) + bx * 2 + tx] = a;
    }
    __syncthreads();
    if (tx == 0 && ty == 0) {
        float a = (s3[0][0] + s3[0][1] + s3[1][0] + s3[1][1]) * 0.25f;
        p.l[4][by * (w >> 4) + bx] = a;
    }
}

// Gradients of all levels in one launch (blockIdx.y = level) + maxGradients of level 0 in the same pass.
// The reference sweeps the LINEAR index range [w, w*(h-1)) so x = 0 and x = w-1 wrap across rows
// (Frame.cpp:658-677); rows 0 and h-1 are defined as zero.
// maxGradients (Frame.cpp:708-759): |grad|, then a 3x1 vertical and a 1x3 horizontal max, all over LINEAR index
// ranges, never-written cells defined as zero.  |grad| is recomputed from the image for the 9
// neighbours (same fp32 operations as reading the stored gradients), which saves a launch and a pass over the
// 4.9 MB gradient level.
struct GradPtrs {
    const float* img[LSD_LEVELS];
    float4* grad[LSD_LEVELS];
    int w[LSD_LEVELS], h[LSD_LEVELS];
    float* maxgrad0;
};
__device__ __forceinline__ float absGradImg(const float* __restrict__ img, int j, int w, int h)
{
    if (j < w || j >= w * (h - 1)) return 0.f;
    const float gx = 0.5f * (__ldg(img + j + 1) - __ldg(img + j - 1));
    const float gy = 0.5f * (__ldg(img + j + w) - __ldg(img + j - w));
    return sqrtf(gx * gx + gy * gy);
}
__device__ __forceinline__ float vmax3Img(const float* __restrict__ img, int t, int w, int h)
{
    if (t < w + 1 || t >= w * (h - 1) - 1) return 0.f;
    float g1 = absGradImg(img, t - w, w, h), g2 = absGradImg(img, t, w, h), g3 = absGradImg(img, t + w, w, h);
    if (g1 < g2) g1 = g2;
    return (g1 < g3) ? g3 : g1;
}
// Level 0 is processed in 32x8 tiles (blockIdx.y == 0): |grad| of the tile + 1-cell halo is computed ONCE into shared
// memory (1.33 evaluations per pixel instead of 9), then the vertical and the horizontal 3-max run on shared memory.
// Cells are addressed by LINEAR index t = y*w + x, so the halo column x = -1 / x = w is the neighbouring row's last /
// first pixel exactly as the reference's linear sweeps see it.  Levels 1..4 (blockIdx.y = level) only need the gradient.
#define GR_TW 32
#define GR_TH 8
__global__ void __launch_bounds__(256) k_gradients(const __grid_constant__ GradPtrs p)
{
    pdlWait();                                    // launched early (programmatic dependent launch): wait for the pyramid kernel
    const int lvl = blockIdx.y;
    const int w = p.w[lvl], h = p.h[lvl];
    const float* __restrict__ img = p.img[lvl];
    if (lvl != 0) {
        const int i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= w * h) return;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i >= w && i < w * (h - 1)) {
            g.x = 0.5f * (img[i + 1] - img[i - 1]);
            g.y = 0.5f * (img[i + w] - img[i - w]);
            g.z = img[i];
        }
        p.grad[lvl][i] = g;
        return;
    }
    __shared__ float ag[GR_TH + 2][GR_TW + 2];      // |grad| at linear index (y0 - 1 + r) * w + (x0 - 1 + c)
    __shared__ float vm[GR_TH][GR_TW + 2];          // vertical 3-max (Frame.cpp:721-734)
    const int tilesX = (w + GR_TW - 1) / GR_TW, tilesY = (h + GR_TH - 1) / GR_TH;
    if ((int)blockIdx.x >= tilesX * tilesY) return;
    const int x0 = (blockIdx.x % tilesX) * GR_TW, y0 = (blockIdx.x / tilesX) * GR_TH;
    const int n = w * h;
    for (int c = threadIdx.x; c < (GR_TH + 2) * (GR_TW + 2); c += 256) {
        const int cx = c % (GR_TW + 2), cy = c / (GR_TW + 2);
        const int gy = y0 - 1 + cy, gx = x0 - 1 + cx;
        float v = 0.f;
        if (gy >= 0 && gy < h && gx <= w) {
            const int t = gy * w + gx;
            if (t >= 0 && t < n) v = absGradImg(img, t, w, h);
        }
        ag[cy][cx] = v;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < GR_TH * (GR_TW + 2); c += 256) {
        const int cx = c % (GR_TW + 2), cy = c / (GR_TW + 2);
        const int gy = y0 + cy, gx = x0 - 1 + cx;
        const int t = gy * w + gx;
        float v = 0.f;
        if (gy < h && gx <= w && !(t < w + 1 || t >= w * (h - 1) - 1)) {
            float g1 = ag[cy][cx], g2 = ag[cy + 1][cx], g3 = ag[cy + 2][cx];
            if (g1 < g2) g1 = g2;
            v = (g1 < g3) ? g3 : g1;
        }
        vm[cy][cx] = v;
    }
    __syncthreads();
    const int tx = threadIdx.x % GR_TW, ty = threadIdx.x / GR_TW;
    const int x = x0 + tx, y = y0 + ty;
    if (x >= w || y >= h) return;
    const int i = y * w + x;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i >= w && i < w * (h - 1)) {
        g.x = 0.5f * (img[i + 1] - img[i - 1]);
        g.y = 0.5f * (img[i + w] - img[i - w]);
        g.z = img[i];
    }
    p.grad[0][i] = g;
    float r;
    if (i >= w + 1 && i < w * (h - 1) - 1) {        // horizontal 3-max (Frame.cpp:740-759)
        float g1 = vm[ty][tx], g2 = vm[ty][tx + 1], g3 = vm[ty][tx + 2];
        if (g1 < g2) g1 = g2;
        r = (g1 < g3) ? g3 : g1;
    } else {
        r = (i >= w && i < w * (h - 1)) ? sqrtf(g.x * g.x + g.y * g.y) : 0.f;   // cells w and w*(h-1)-1 keep the raw |grad|
    }
    p.maxgrad0[i] = r;
}

// Frame::setDepthFromGroundTruth, Frame.cpp:264-285
__global__ void __launch_bounds__(256) k_set_depth_gt(const float* __restrict__ depth, const float* __restrict__ maxgrad,
                                                      float* __restrict__ idepth, float* __restrict__ idepthVar,
                                                      int w, int h, float minUseGrad, float cov_scale)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w * h) return;
    const int x = i % w, y = i / w;
    const float d = depth[i];
    if (x > 0 && x < w - 1 && y > 0 && y < h - 1 && maxgrad[i] >= minUseGrad && !isnan(d) && d > 0) {
        idepth[i] = 1.0f / d;
        idepthVar[i] = 0.01f * 0.01f * cov_scale;     // VAR_GT_INIT_INITIAL * cov_scale
    } else {
        idepth[i] = -1;
        idepthVar[i] = -1;
    }
}

// Frame::buildIDepthAndIDepthVar for levels 1..4 in one launch: one CTA = one 16x16 level-0 tile.
__device__ __forceinline__ float2 mergeIdepth4(float2 a, float2 b, float2 c, float2 d)
{   // Frame.cpp:819-871, sources visited in the order idx, idx+1, idx+sw, idx+sw+1; .x = idepth, .y = var
    float idepthSumsSum = 0, ivarSumsSum = 0;
    int num = 0;
    float2 s[4] = { a, b, c, d };
#pragma unroll
    for (int k = 0; k < 4; k++) {
        float var = s[k].y;
        if (var > 0) {
            float ivar = 1.0f / var;
            ivarSumsSum += ivar;
            idepthSumsSum += ivar * s[k].x;
            num++;
        }
    }
    if (num > 0) {
        float depth = ivarSumsSum / idepthSumsSum;
        return make_float2(1.0f / depth, num / ivarSumsSum);
    }
    return make_float2(-1.f, -1.f);
}
__global__ void __launch_bounds__(256) k_idepth_pyramid(PyrPtrs id, PyrPtrs var, int w, int h)
{
    __shared__ float2 s0[16][17], s1[8][9], s2[4][5], s3[2][3];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int bx = blockIdx.x, by = blockIdx.y;
    const int x = bx * 16 + tx, y = by * 16 + ty;
    s0[ty][tx] = make_float2(id.l[0][y * w + x], var.l[0][y * w + x]);
    __syncthreads();
    if (tx < 8 && ty < 8) {
        float2 r = mergeIdepth4(s0[2 * ty][2 * tx], s0[2 * ty][2 * tx + 1], s0[2 * ty + 1][2 * tx], s0[2 * ty + 1][2 * tx + 1]);
        s1[ty][tx] = r;
        int o = (by * 8 + ty) * (w >> 1) + bx * 8 + tx;
        id.l[1][o] = r.x; var.l[1][o] = r.y;
    }
    __syncthreads();
    if (tx < 4 && ty < 4) {
        float2 r = mergeIdepth4(s1[2 * ty][2 * tx], s1[2 * ty][2 * tx + 1], s1[2 * ty + 1][2 * tx], s1[2 * ty + 1][2 * tx + 1]);
        s2[ty][tx] = r;
        int o = (by * 4 + ty) * (w >> 2) + bx * 4 + tx;
        id.l[2][o] = r.x; var.l[2][o] = r.y;
    }
    __syncthreads();
    if (tx < 2 && ty < 2) {
        float2 r = mergeIdepth4(s2[2 * ty][2 * tx], s2[2 * ty][2 * tx + 1], s2[2 * ty + 1][2 * tx], s2[2 * ty + 1][2 * tx + 1]);
        s3[ty][tx] = r;
        int o = (by * 2 + ty) * (w >> 3) + bx * 2 + tx;
        id.l[3][o] = r.x; var.l[3][o] = r.y;
    }
    __syncthreads();
    if (tx == 0 && ty == 0) {
        float2 r = mergeIdepth4(s3[0][0], s3[0][1], s3[1][0], s3[1][1]);
        int o = by * (w >> 4) + bx;
        id.l[4][o] = r.x; var.l[4][o] = r.y;
    }
}

// UndistorterPTAM::undistort alone (u8 -> u8), util/Undistorter.cpp:380-410
__global__ void __launch_bounds__(256) k_undistort(const uint8_t* __restrict__ raw, int in_w, const float* __restrict__ remapX,
                                                   const float* __restrict__ remapY, int n, uint8_t* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = undistortPixel(raw, in_w, __ldg(remapX + i), __ldg(remapY + i));
}
