// interp.cu — reference preparation: border replication and the 16 quarter-pel luma planes.
//
// Stands in for JM's UnifiedOneForthPix ‖ getSubImagesLuma and the
// UMV reference fetch (a9: clamped coordinates = unbounded edge replication).  The arithmetic is
// H.264 8.4.2.2.1: half-pel 6-tap (1,-5,20,20,-5,1) with (x+16)>>5, centre from unrounded
// intermediates with (x+512)>>10, quarter-pel (p+q+1)>>1.
//
// HBM-bound pass: per padded pixel 1 byte read, 16 bytes written (17 B algorithmic).  A CTA
// produces a 128x8 tile of all 16 planes from a (128+6)x(8+6) integer tile staged in shared
// memory; the unrounded horizontal intermediates b1 are staged once and reused by the centre
// filter; every thread packs 4 horizontally adjacent samples and stores one 32-bit word per plane
// (a warp writes 128 contiguous bytes per plane row).
#include "jmme_dev.cuh"

namespace {

constexpr int TW = 128, TH = 8;
constexpr int GW = TW + 8, GH = TH + 6;      // integer tile: cols -2..TW+5, rows -2..TH+3

__device__ __forceinline__ int tap6(int a, int b, int c, int d, int e, int f)
{
    return a - 5 * b + 20 * c + 20 * d - 5 * e + f;
}
__device__ __forceinline__ int clip255(int v) { return min(max(v, 0), 255); }

// src: raw picture w_in x h_in (stride), out: n_planes planes of ps x ph.
__global__ void __launch_bounds__(256) interp_kernel(const uint8_t *__restrict__ src, int w_in, int h_in, int stride,
                                                     int pad, int ps, int ph, int n_planes, int y_begin,
                                                     uint8_t *__restrict__ out)
{
    __shared__ uint8_t Gs[GH][GW];
    __shared__ int16_t B1s[GH][TW + 2];      // unrounded horizontal half-pel, cols 0..TW
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TW, y0 = y_begin + blockIdx.y * TH;   // padded-plane coordinates of the tile

    for (int i = tid; i < GH * GW; i += 256) {
        int r = i / GW, c = i - r * GW;
        int sx = d_clamp(x0 + c - 2 - pad, 0, w_in - 1), sy = d_clamp(y0 + r - 2 - pad, 0, h_in - 1);
        Gs[r][c] = src[(size_t)sy * stride + sx];
    }
    __syncthreads();
    if (n_planes == 1) {                     // integer plane only
        const int tx = tid & 31, ty = tid >> 5;
        const int x = x0 + 4 * tx, y = y0 + ty;
        if (x < ps && y < ph) {
            uint32_t v = Gs[ty + 2][4 * tx + 2] | (Gs[ty + 2][4 * tx + 3] << 8) | (Gs[ty + 2][4 * tx + 4] << 16) |
                         (Gs[ty + 2][4 * tx + 5] << 24);
            *(uint32_t *)(out + (size_t)y * ps + x) = v;
        }
        return;
    }
    for (int i = tid; i < GH * (TW + 1); i += 256) {
        int r = i / (TW + 1), c = i - r * (TW + 1);           // b1 at tile column c, tile row r-2
        const uint8_t *g = &Gs[r][c];                          // Gs column index = tile column + 2
        B1s[r][c] = (int16_t)tap6(g[0], g[1], g[2], g[3], g[4], g[5]);
    }
    __syncthreads();

    const int tx = tid & 31, ty = tid >> 5;
    const int x = x0 + 4 * tx, y = y0 + ty;
    if (x >= ps || y >= ph) return;
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; i++) w[i] = 0;
    int hprev;                                                  // h at column c (rounded)
    {
        const int c = 4 * tx + 2;
        hprev = clip255((tap6(Gs[ty][c], Gs[ty + 1][c], Gs[ty + 2][c], Gs[ty + 3][c], Gs[ty + 4][c], Gs[ty + 5][c]) + 16) >> 5);
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int tc = 4 * tx + q;           // tile column
        const int c = tc + 2;                // Gs column
        const int r = ty + 2;                // Gs / B1s row
        const int g = Gs[r][c], gr = Gs[r][c + 1], gd = Gs[r + 1][c];
        const int b = clip255((B1s[r][tc] + 16) >> 5);
        const int s = clip255((B1s[r + 1][tc] + 16) >> 5);
        const int h = hprev;
        const int m = clip255((tap6(Gs[ty][c + 1], Gs[ty + 1][c + 1], Gs[ty + 2][c + 1], Gs[ty + 3][c + 1],
                                    Gs[ty + 4][c + 1], Gs[ty + 5][c + 1]) + 16) >> 5);
        hprev = m;
        const int j = clip255((tap6(B1s[r - 2][tc], B1s[r - 1][tc], B1s[r][tc], B1s[r + 1][tc], B1s[r + 2][tc],
                                    B1s[r + 3][tc]) + 512) >> 10);
        const int sh = 8 * q;
        // plane index = yfrac*4 + xfrac
        w[0] |= (uint32_t)g << sh;
        w[1] |= (uint32_t)((g + b + 1) >> 1) << sh;            // a (1,0)
        w[2] |= (uint32_t)b << sh;                              // b (2,0)
        w[3] |= (uint32_t)((gr + b + 1) >> 1) << sh;           // c (3,0)
        w[4] |= (uint32_t)((g + h + 1) >> 1) << sh;            // d (0,1)
        w[5] |= (uint32_t)((b + h + 1) >> 1) << sh;            // e (1,1)
        w[6] |= (uint32_t)((b + j + 1) >> 1) << sh;            // f (2,1)
        w[7] |= (uint32_t)((b + m + 1) >> 1) << sh;            // g (3,1)
        w[8] |= (uint32_t)h << sh;                              // h (0,2)
        w[9] |= (uint32_t)((h + j + 1) >> 1) << sh;            // i (1,2)
        w[10] |= (uint32_t)j << sh;                             // j (2,2)
        w[11] |= (uint32_t)((j + m + 1) >> 1) << sh;           // k (3,2)
        w[12] |= (uint32_t)((gd + h + 1) >> 1) << sh;          // n (0,3)
        w[13] |= (uint32_t)((h + s + 1) >> 1) << sh;           // p (1,3)
        w[14] |= (uint32_t)((j + s + 1) >> 1) << sh;           // q (2,3)
        w[15] |= (uint32_t)((m + s + 1) >> 1) << sh;           // r (3,3)
    }
    const size_t psz = (size_t)ps * ph;
    uint8_t *o = out + (size_t)y * ps + x;
#pragma unroll
    for (int i = 0; i < 16; i++) *(uint32_t *)(o + psz * i) = w[i];
}

// replicate the current picture to w16 x h16
__global__ void pad_cur_kernel(const uint8_t *__restrict__ src, int w_in, int h_in, int stride, int w16, int h16,
                               uint8_t *__restrict__ dst)
{
    int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
    if (x >= w16 || y >= h16) return;
    int sy = min(y, h_in - 1);
    uint32_t v = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) v |= (uint32_t)src[(size_t)sy * stride + min(x + q, w_in - 1)] << (8 * q);
    *(uint32_t *)(dst + (size_t)y * w16 + x) = v;
}

}  // namespace

// rows [y_begin, y_end) of the padded planes are produced (y_begin a multiple of 4 is not required)
cudaError_t jmme_launch_interp(const uint8_t *src, int w_in, int h_in, int stride, int pad, int ps, int ph,
                               int n_planes, uint8_t *out, int y_begin, int y_end, cudaStream_t st)
{
    dim3 grid((ps + TW - 1) / TW, (y_end - y_begin + TH - 1) / TH);
    interp_kernel<<<grid, 256, 0, st>>>(src, w_in, h_in, stride, pad, ps, ph, n_planes, y_begin, out);
    return cudaGetLastError();
}

cudaError_t jmme_launch_pad_cur(const uint8_t *src, int w_in, int h_in, int stride, int w16, int h16, uint8_t *dst,
                                cudaStream_t st)
{
    dim3 grid((w16 / 4 + 127) / 128, h16);
    pad_cur_kernel<<<grid, 128, 0, st>>>(src, w_in, h_in, stride, w16, h16, dst);
    return cudaGetLastError();
}
