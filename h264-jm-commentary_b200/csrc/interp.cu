// interp.cu — reference preparation: border replication and the 16 quarter-pel luma planes.
//
// Stands in for JM's UnifiedOneForthPix ‖ getSubImagesLuma and the
// UMV reference fetch (a9: clamped coordinates = unbounded edge replication).  The arithmetic is
// H.264 8.4.2.2.1: half-pel 6-tap (1,-5,20,20,-5,1) with (x+16)>>5, centre from unrounded
// intermediates with (x+512)>>10, quarter-pel (p+q+1)>>1.
//
// HBM-bound pass: per padded pixel 1 byte read, 16 bytes written (17 B algorithmic).  A CTA
// produces a 128x8 tile of all 16 planes from a (128+16)x(8+6) integer tile staged in shared
// memory as words; the unrounded horizontal intermediates b1 are staged once and reused by the
// centre filter; every thread owns 4 horizontally adjacent samples and stores one 32-bit word per
// plane (a warp writes 128 contiguous bytes per plane row).
#include "jmme_dev.cuh"

namespace {

constexpr int TW = 128, TH = 8;
constexpr int GH = TH + 6;                   // integer tile rows -2..TH+3
constexpr int GWW = 36;                      // integer tile words per row: byte k = tile column k-4 (cols -4..139)
constexpr int B1W = TW + 4;                  // b1 row stride (int16), multiple of 4 so that 4 values are one LDS.64

__device__ __forceinline__ int tap6(int a, int b, int c, int d, int e, int f)
{
    return (a + f) - 5 * (b + e) + 20 * (c + d);
}
__device__ __forceinline__ int clip255(int v) { return min(max(v, 0), 255); }
__device__ __forceinline__ int rnd5(int v) { return clip255((v + 16) >> 5); }
__device__ __forceinline__ uint32_t pack4(int a, int b, int c, int d)
{
    return (uint32_t)a | ((uint32_t)b << 8) | ((uint32_t)c << 16) | ((uint32_t)d << 24);
}
// per-byte (a + b + 1) >> 1 of two packed words (5 ALU operations for 4 samples)
__device__ __forceinline__ uint32_t avg4(uint32_t a, uint32_t b) { return __vavgu4(a, b); }

// src: raw picture w_in x h_in (stride), out: n_planes planes of ps x ph; rows [y_begin, ...) are produced.
// Thread (tx, ty) of the 32x8 CTA owns the 4 samples at tile columns 4tx..4tx+3 of tile row ty: everything it
// needs comes from shared memory as aligned 32/64-bit words (integer tile: word tx+1 / tx+2 of six rows;
// unrounded horizontal half-pels b1: four int16 = one LDS.64 of six rows); vertical 6-taps run on 16-bit lane
// pairs (two columns per 32-bit operation); the twelve quarter-pel planes are byte-wise averages of packed words.
__global__ void __launch_bounds__(256) interp_kernel(const uint8_t *__restrict__ src, int w_in, int h_in, int stride,
                                                     int pad, int ps, int ph, int n_planes, int y_begin,
                                                     uint8_t *__restrict__ out)
{
    __shared__ __align__(16) uint32_t Gs[GH][GWW];
    __shared__ __align__(16) int16_t B1s[GH][B1W];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TW, y0 = y_begin + blockIdx.y * TH;   // padded-plane coordinates of the tile
    const int sx0 = x0 - 4 - pad;                                     // picture column of tile byte 0

    // ---- integer tile: rows are clamped (edge replication), columns use aligned words when the whole tile row
    //      lies inside the picture, else clamped bytes --------------------------------------------------------
    const bool words = sx0 >= 0 && sx0 + 4 * GWW <= w_in && !(stride & 3) && !((uintptr_t)src & 3);
    if (words) {
        for (int i = tid; i < GH * GWW; i += 256) {
            const int r = i / GWW, k = i - r * GWW;
            const int sy = d_clamp(y0 + r - 2 - pad, 0, h_in - 1);
            Gs[r][k] = __ldg((const uint32_t *)(src + (size_t)sy * stride + sx0) + k);
        }
    } else {
        uint8_t *gb = (uint8_t *)&Gs[0][0];
        for (int i = tid; i < GH * GWW * 4; i += 256) {
            const int r = i / (GWW * 4), c = i - r * (GWW * 4);
            const int sx = d_clamp(sx0 + c, 0, w_in - 1), sy = d_clamp(y0 + r - 2 - pad, 0, h_in - 1);
            gb[i] = src[(size_t)sy * stride + sx];
        }
    }
    __syncthreads();
    const int tx = tid & 31, ty = tid >> 5;
    const int x = x0 + 4 * tx, y = y0 + ty;
    if (n_planes == 1) {                     // integer plane only
        if (x < ps && y < ph) *(uint32_t *)(out + (size_t)y * ps + x) = Gs[ty + 2][tx + 1];
        return;
    }
    // ---- b1: unrounded horizontal half-pel of tile columns 0..TW-1, all GH rows, four per thread and pass -----
    for (int i = tid; i < GH * (TW / 4); i += 256) {
        const int r = i >> 5, q = i & 31;                    // tile columns 4q..4q+3 need bytes 4q+2..4q+10
        const uint32_t w0 = Gs[r][q], w1 = Gs[r][q + 1], w2 = Gs[r][q + 2];
        int e[9];
        e[0] = (w0 >> 16) & 255; e[1] = w0 >> 24;
        e[2] = w1 & 255; e[3] = (w1 >> 8) & 255; e[4] = (w1 >> 16) & 255; e[5] = w1 >> 24;
        e[6] = w2 & 255; e[7] = (w2 >> 8) & 255; e[8] = (w2 >> 16) & 255;
        int v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = tap6(e[k], e[k + 1], e[k + 2], e[k + 3], e[k + 4], e[k + 5]);
        uint2 pk;
        pk.x = (uint32_t)(v[0] & 0xFFFF) | ((uint32_t)v[1] << 16);
        pk.y = (uint32_t)(v[2] & 0xFFFF) | ((uint32_t)v[3] << 16);
        *(uint2 *)&B1s[r][4 * q] = pk;
    }
    __syncthreads();
    if (x >= ps || y >= ph) return;

    const int r = ty + 2;                                    // tile row ty in Gs / B1s
    // integer words of rows ty..ty+5 (tile rows ty-2..ty+3): columns 4tx..4tx+3 and 4tx+4..4tx+7
    uint32_t wa[6], wb0[6];
#pragma unroll
    for (int k = 0; k < 6; k++) { wa[k] = Gs[ty + k][tx + 1]; wb0[k] = Gs[ty + k][tx + 2] & 255; }
    // vertical 6-tap on 16-bit lane pairs (columns 0|2 and 1|3), column 4 scalar
    int h02, h13, h4;
    {
        int p02[6], p13[6];
#pragma unroll
        for (int k = 0; k < 6; k++) { p02[k] = (int)__byte_perm(wa[k], 0, 0x4240); p13[k] = (int)__byte_perm(wa[k], 0, 0x4341); }
        h02 = tap6(p02[0], p02[1], p02[2], p02[3], p02[4], p02[5]);
        h13 = tap6(p13[0], p13[1], p13[2], p13[3], p13[4], p13[5]);
        h4 = tap6((int)wb0[0], (int)wb0[1], (int)wb0[2], (int)wb0[3], (int)wb0[4], (int)wb0[5]);
    }
    int hv[5];
    {
        const int l02 = (int)(short)h02, l13 = (int)(short)h13;
        hv[0] = rnd5(l02); hv[2] = rnd5((h02 - l02) >> 16);
        hv[1] = rnd5(l13); hv[3] = rnd5((h13 - l13) >> 16);
        hv[4] = rnd5(h4);
    }
    // b1 of rows r-2..r+3 at this thread's four columns
    int b1v[6][4];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        const uint2 p = *(const uint2 *)&B1s[ty + k][4 * tx];
        b1v[k][0] = (int)(short)p.x; b1v[k][1] = (int)p.x >> 16;
        b1v[k][2] = (int)(short)p.y; b1v[k][3] = (int)p.y >> 16;
    }
    int bv[4], sv[4], jv[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        bv[q] = rnd5(b1v[2][q]);
        sv[q] = rnd5(b1v[3][q]);
        jv[q] = clip255((tap6(b1v[0][q], b1v[1][q], b1v[2][q], b1v[3][q], b1v[4][q], b1v[5][q]) + 512) >> 10);
    }
    const uint32_t G = wa[2];
    const uint32_t GR = __funnelshift_r(wa[2], Gs[r][tx + 2], 8);
    const uint32_t GD = wa[3];
    const uint32_t B = pack4(bv[0], bv[1], bv[2], bv[3]), S = pack4(sv[0], sv[1], sv[2], sv[3]);
    const uint32_t H = pack4(hv[0], hv[1], hv[2], hv[3]), M = pack4(hv[1], hv[2], hv[3], hv[4]);
    const uint32_t J = pack4(jv[0], jv[1], jv[2], jv[3]);
    uint32_t w[16];                                          // plane index = yfrac*4 + xfrac
    w[0] = G;           w[1] = avg4(G, B);  w[2] = B;   w[3] = avg4(GR, B);
    w[4] = avg4(G, H);  w[5] = avg4(B, H);  w[6] = avg4(B, J);  w[7] = avg4(B, M);
    w[8] = H;           w[9] = avg4(H, J);  w[10] = J;  w[11] = avg4(J, M);
    w[12] = avg4(GD, H); w[13] = avg4(H, S); w[14] = avg4(J, S); w[15] = avg4(M, S);
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

// chroma ME: a w x h picture replicated into a pw x ph plane with `pad` border samples on every side (pad = 0:
// the current chroma padded to w16/2 x h16/2)
__global__ void pad_plane_kernel(const uint8_t *__restrict__ src, int w, int h, int stride, int pad, int pw, int ph,
                                 uint8_t *__restrict__ dst)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= pw || y >= ph) return;
    dst[(size_t)y * pw + x] = src[(size_t)d_clamp(y - pad, 0, h - 1) * stride + d_clamp(x - pad, 0, w - 1)];
}

}  // namespace

cudaError_t jmme_launch_pad_plane(const uint8_t *src, int w, int h, int stride, int pad, int pw, int ph, uint8_t *dst,
                                  cudaStream_t st)
{
    pad_plane_kernel<<<dim3((pw + 127) / 128, ph), 128, 0, st>>>(src, w, h, stride, pad, pw, ph, dst);
    return cudaGetLastError();
}

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
